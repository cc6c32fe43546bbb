"""Committed golden vectors (tests/golden/small_case.npz, made by tests/golden/make_golden.py from the oracle)."""
import os

import numpy as np
import pytest

from conftest import relerr

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "small_case.npz")


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


def test_oracle_reproduces_golden(gold):
    from oracle import modes as om
    from oracle import windows as ow
    a = om.AnlmModes(float(gold["kmax"]), 500.0, 1000.0)
    c = om.ClnnModes(a)
    assert np.array_equal(c.lnn, gold["lnn"]) and np.array_equal(a.nmax_l, gold["nmax_l"])
    assert np.allclose(a.knl, gold["knl"], rtol=1e-12, equal_nan=True)
    wm = ow.ConfigurationSpaceModes(500.0, 1000.0, int(gold["nr"]), int(gold["win_nside"]))
    W = ow.calc_Wr_lm(gold["win"][:3], 2 * a.lmax, a.nside)
    assert relerr(W, gold["Wr_lm"][:3]) < 1e-13
    M = ow.power_win_mix(gold["win"], gold["win"], wm, c)
    assert relerr(M, gold["M"]) < 1e-12
    assert relerr(ow.win_lnn(gold["win"], wm, c), gold["Wlnn"]) < 1e-12


@pytest.mark.gpu
def test_cuda_reproduces_golden(gold):
    import warnings

    import sfb_b200 as sfb
    a = sfb.AnlmModes(float(gold["kmax"]), 500.0, 1000.0)
    c = sfb.ClnnModes(a)
    assert np.array_equal(c.lnn, gold["lnn"])          # index tables bit-exact
    wm = sfb.ConfigurationSpaceModes(500.0, 1000.0, int(gold["nr"]), int(gold["win_nside"]))
    win = gold["win"]
    assert relerr(sfb.calc_Wr_lm(win, 2 * a.lmax, a.nside), gold["Wr_lm"]) < 1e-10
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        assert relerr(sfb.power_win_mix(win, wm, c), gold["M"]) < 1e-10
        got = sfb.power_win_mix(win, wm, c, div2Lp1=True, interchange_NN=True, lnn_min=7)
        assert relerr(got, gold["M_div_interchange_min7"]) < 1e-10
        wt, v = sfb.bandpower_binning_weights(c, dl=3)
        N = sfb.power_win_mix(win, wt, v, wm, sfb.ClnnBinnedModes(wt, v, c))
        assert relerr(N, gold["N_binned_dl3"]) < 1e-10
        assert relerr(sfb.win_lnn(win, wm, c), gold["Wlnn"]) < 1e-10
