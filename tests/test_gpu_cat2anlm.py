"""SFB transforms next to the window path on the GPU against the oracle (oracle/cat2anlm.py) and the
reference's own tests (test/test_cat2anlm.jl)."""
import math

import numpy as np
import pytest

from conftest import relerr
from oracle import cat2anlm as oc
from oracle import modes as om
from oracle import windows as ow

pytestmark = pytest.mark.gpu
RTOL = 1e-10


def _modes(args, nr, nside=None):
    import sfb_b200 as sfb
    oa, a = om.AnlmModes(*args, nside=nside), sfb.AnlmModes(*args, nside=nside)
    rmin, rmax = args[-2], args[-1]
    owm = ow.ConfigurationSpaceModes(rmin, rmax, nr, oa.nside)
    wm = sfb.ConfigurationSpaceModes(rmin, rmax, nr, a.nside)
    return sfb, oa, a, owm, wm


@pytest.mark.parametrize("args,nr,nside", [((3, 4, 500.0, 1000.0), 30, 8), ((0.02, 500.0, 1000.0), 21, 16),
                                           ((2, 9, 0.0, 1000.0), 12, 12)])
def test_field2anlm_anlm2field_win_rhat_ln_match_oracle(args, nr, nside):
    sfb, oa, a, owm, wm = _modes(args, nr, nside)
    rng = np.random.default_rng(nr)
    f = rng.random((wm.nr, wm.npix))
    ref = oc.field2anlm(f, owm, oa)
    got = sfb.field2anlm(f, wm, a)
    assert got.shape == ref.shape and relerr(got, ref) < RTOL
    fx = sfb.anlm2field(got, wm, a)
    assert fx.shape == (wm.nr, wm.npix) and relerr(fx, oc.anlm2field(ref, owm, oa)) < RTOL
    w = sfb.win_rhat_ln(f, wm, a)
    wref = oc.win_rhat_ln(f, owm, oa)
    assert w.shape == wref.shape and np.array_equal(np.isnan(w), np.isnan(wref))
    assert relerr(np.nan_to_num(w), np.nan_to_num(wref)) < RTOL
    s = sfb.SeparableArray(f[:, 0], f[0])
    ws = sfb.win_rhat_ln(s, wm, a)
    _, wln = oc.win_rhat_ln(ow.SeparableArray(f[:, 0], f[0]), owm, oa)
    assert np.allclose(ws.w_ln.reshape(wln.shape, order="F"), wln, rtol=1e-10, equal_nan=True)


def test_cat2amln_matches_oracle_and_field_route():
    sfb, oa, a, owm, wm = _modes((3, 5, 500.0, 1000.0), 40, 8)
    rng = np.random.default_rng(9)
    ngal = 3000
    rtp = np.stack([rng.uniform(500, 1000, ngal), np.arccos(rng.uniform(-1, 1, ngal)), rng.uniform(0, 2 * math.pi, ngal)])
    wgt = rng.random(ngal)
    win = rng.random((wm.nr, wm.npix))
    wr = sfb.win_rhat_ln(win, wm, a)
    got = sfb.cat2amln(rtp, a, 2e-4, wr, wgt, batch=7)                 # several batches, the last one ragged
    ref = oc.cat2amln(rtp, oa, 2e-4, oc.win_rhat_ln(win, owm, oa), wgt)
    assert relerr(got, ref) < RTOL
    assert np.array_equal(sfb.cat2anlm.ang2pix_ring(8, rtp[1], rtp[2]), oc.ang2pix_ring(8, rtp[1], rtp[2]))
    # test/test_cat2anlm.jl:235-251: the catalogue route with an empty catalogue equals the field route
    f1 = -sfb.cat2amln(np.zeros((3, 0)), a, 1.0, wr, [])
    assert relerr(f1, sfb.field2anlm(win, wm, a)) < 1e-12
    cl = sfb.amln2clnn(got, got, sfb.ClnnModes(a))
    assert relerr(cl, oc.amln2clnn(ref, ref, om.ClnnModes(oa))) < RTOL


def test_round_trips_like_the_reference():
    # test/test_cat2anlm.jl:261-291: nmax = lmax = 10, nside = 32, nr = 100; rtol 1e-4 between successive round trips
    sfb, oa, a, owm, wm = _modes((10, 10, 500.0, 1000.0), 100, 32)
    f1 = np.random.default_rng(0).random((wm.nr, wm.npix))
    a1 = sfb.field2anlm(f1, wm, a)
    f2 = sfb.anlm2field(a1, wm, a)
    a2 = sfb.field2anlm(f2, wm, a)
    f3 = sfb.anlm2field(a2, wm, a)
    a3 = sfb.field2anlm(f3, wm, a)
    assert relerr(f2, f3) < 1e-4 and relerr(a1, a2) < 1e-4 and relerr(a2, a3) < 1e-4


def test_field2anlm_single_voxel_at_reference_resolution():
    # test/test_cat2anlm.jl:153-196 at its own nside = 256: δ voxel -> Δr r² g_nl(r) conj(Y_lm) Ω_p, rtol 1e-5
    from scipy.special import sph_harm_y
    from oracle import healpix as hp
    sfb, oa, a, owm, wm = _modes((2, 3, 500.0, 1000.0), 100, 256)
    r, dr = ow.window_r(owm)
    npix = wm.npix
    f = np.zeros((wm.nr, npix), order="F")
    for (i, j) in ((0, 0), (41, 3 * 256 * 256 + 7), (82, npix - 5)):
        f[i, j] = 1.0
        got = sfb.field2anlm(f, wm, a)
        f[i, j] = 0.0
        th, ph = hp.pix2ang_ring(256, np.array([j]))
        exp = np.array([dr * r[i] ** 2 * oa.basisfunctions(n, l, r[i]) * np.conj(sph_harm_y(l, m, th[0], ph[0])) * 4 * math.pi / npix
                        for (n, l, m) in (om.getnlm(oa, k) for k in range(1, om.getnlmsize(oa) + 1))])
        assert relerr(got, exp) < 1e-5
