"""calc_wmix on the GPU against the oracle (itself pinned on the reference-held golden element,
tests/test_reference_golden.py) and through the reference's own identities."""
import warnings

import numpy as np
import pytest

from conftest import relerr
from oracle import modes as om
from oracle import windows as ow

pytestmark = pytest.mark.gpu
RTOL = 1e-10


def _window(rng, wm):
    mask = rng.random(wm.npix)
    mask[: wm.npix // 3] *= 0.3
    phi = np.exp(-(wm.r / (0.55 * wm.rmax)) ** 2)
    win = np.outer(phi, mask) * (1 + 0.2 * rng.random((wm.nr, wm.npix)))
    return win / win.max()


@pytest.mark.parametrize("args,nr", [((2, 5, 500.0, 1000.0), 40), ((3, 4, 500.0, 1000.0), 24), ((0.02, 500.0, 1000.0), 60)])
@pytest.mark.parametrize("neg_m", [False, True])
def test_calc_wmix_matches_oracle(args, nr, neg_m):
    import sfb_b200 as sfb
    oa, a = om.AnlmModes(*args), sfb.AnlmModes(*args)
    owm = ow.ConfigurationSpaceModes(500.0, 1000.0, nr, oa.nside)
    wm = sfb.ConfigurationSpaceModes(500.0, 1000.0, nr, a.nside)
    win = _window(np.random.default_rng(nr), owm)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref = ow.calc_wmix(win, owm, oa, neg_m=neg_m)
        got = sfb.calc_wmix(win, wm, a, neg_m=neg_m)
    assert got.shape == ref.shape == (om.getnlmsize(oa),) * 2
    assert np.isfinite(got).all()
    assert relerr(got, ref) < RTOL


def test_power_win_mix_from_wmix_identity():
    """test/test_windows.jl:385-409: power_win_mix(wmix, wmix_negm, cmodes) ≈ power_win_mix(win, wmodes, cmodes), rtol 1e-10:
    the brute-force m-sums over the GPU's calc_wmix_all against the GPU's coupling matrix."""
    import sfb_b200 as sfb
    oa, a = om.AnlmModes(2, 5, 500.0, 1000.0), sfb.AnlmModes(2, 5, 500.0, 1000.0)
    owm = ow.ConfigurationSpaceModes(500.0, 1000.0, 100, oa.nside)
    wm = sfb.ConfigurationSpaceModes(500.0, 1000.0, 100, a.nside)
    oc, c = om.ClnnModes(oa), sfb.ClnnModes(a)
    win = _window(np.random.default_rng(1), owm)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        wmix, wmix_negm = sfb.calc_wmix_all(win, wm, a)
        M = sfb.power_win_mix(win, wm, c)
    brute = ow.power_win_mix_from_wmix(wmix, wmix_negm, oc)
    assert relerr(brute, M) < RTOL


def test_calc_wmix_reference_golden_element():
    """wmix[123,121] of test/test_windows.jl:233-252, computed by the CUDA calc_wmix end to end."""
    import sfb_b200 as sfb
    GOLDEN = -0.025087015337107783 - 1.0170304578086492e-5j
    oa, a = om.AnlmModes(0.019, 500.0, 1000.0), sfb.AnlmModes(0.019, 500.0, 1000.0)
    owm = ow.ConfigurationSpaceModes(500.0, 1000.0, 250, oa.nside)
    wm = sfb.ConfigurationSpaceModes(500.0, 1000.0, 250, a.nside)
    win = ow.make_window(owm, "radial", "ang_sixteenth", "separable", "rotate", "dense")
    wmix = sfb.calc_wmix(win, wm, a)
    assert wmix.shape == (222, 222)
    assert abs(wmix[122, 120] - GOLDEN) <= 1e-10 * abs(GOLDEN), wmix[122, 120]
    # W_lnn' = (2l+1)^-1 Σ_m W_{nlm}^{n'lm}   (test/test_windows.jl:253-255, rtol 1e-3 there)
    oc, c = om.ClnnModes(oa), sfb.ClnnModes(a)
    wlnn = sfb.win_lnn(win, wm, c)
    assert np.allclose(ow.sum_m_lmeqLM(wmix, oc), wlnn, rtol=1e-3)


def test_calc_wmix_full_sky_is_identity():
    """Full-sky window => W_{nlm}^{n'l'm'} = δ up to the radial quadrature (test/test_window_chains.jl:405-415)."""
    import sfb_b200 as sfb
    a = sfb.AnlmModes(2, 6, 500.0, 1000.0)
    wm = sfb.ConfigurationSpaceModes(500.0, 1000.0, 1000, a.nside)
    wmix = sfb.calc_wmix(np.ones((wm.nr, wm.npix)), wm, a)
    assert np.allclose(wmix, np.eye(wmix.shape[0]), atol=1e-4)
