"""Oracle for the SFB transforms of SURVEY §8f row 3 against the reference's own tests (test/test_cat2anlm.jl)."""
import math

import numpy as np
from scipy.special import sph_harm_y

from oracle import cat2anlm as oc
from oracle import healpix as hp
from oracle import modes as om
from oracle import windows as ow


def test_ang2pix_inverts_pix2ang():
    for nside in (1, 2, 4, 16):
        pix = np.arange(12 * nside * nside)
        th, ph = hp.pix2ang_ring(nside, pix)
        assert np.array_equal(oc.ang2pix_ring(nside, th, ph), pix)


def test_field2anlm_single_voxel():
    # test/test_cat2anlm.jl:153-196: f = δ at (shell i, pixel j) -> Δr r² g_nl(r) conj(Y_lm(θ,φ)) Ω_p at rtol 1e-5
    # (the Jacobi passes move a single-pixel alm by O(1/nside²): the reference's rtol needs its nside = 256)
    am = om.AnlmModes(2, 3, 500.0, 1000.0, nside=256)
    wm = ow.ConfigurationSpaceModes(500.0, 1000.0, 3, am.nside)
    npix = wm.npix
    r, dr = ow.window_r(wm)
    for (i, j) in ((0, 0), (1, 3 * 256 * 256 + 7), (2, npix - 1)):
        f = np.zeros((wm.nr, npix))
        f[i, j] = 1.0
        got = oc.field2anlm(f, wm, am)
        th, ph = hp.pix2ang_ring(am.nside, np.array([j]))
        exp = np.empty_like(got)
        for idx in range(1, om.getnlmsize(am) + 1):
            n, l, m = om.getnlm(am, idx)
            exp[idx - 1] = dr * r[i] ** 2 * am.basisfunctions(n, l, r[i]) * np.conj(sph_harm_y(l, m, th[0], ph[0])) * 4 * math.pi / npix
        assert np.linalg.norm(got - exp) <= 1e-5 * np.linalg.norm(exp)


def test_catalogue_route_equals_field_route_and_round_trips():
    # test/test_cat2anlm.jl:235-251 (field2anlm_v1 ≈ v2) and :261-291 (round trips, rtol 1e-4)
    rng = np.random.default_rng(4)
    am = om.AnlmModes(3, 4, 500.0, 1000.0, nside=8)
    wm = ow.ConfigurationSpaceModes(500.0, 1000.0, 100, am.nside)      # nr = 100 as in the reference test
    f1 = rng.random((wm.nr, wm.npix))
    a2 = oc.field2anlm(f1, wm, am)
    a1 = -oc.cat2amln(np.zeros((3, 0)), am, 1.0, oc.win_rhat_ln(f1, wm, am), [])
    assert np.linalg.norm(a1 - a2) <= 1e-12 * np.linalg.norm(a2)
    f2 = oc.anlm2field(a2, wm, am)
    a3 = oc.field2anlm(f2, wm, am)
    f3 = oc.anlm2field(a3, wm, am)
    assert np.linalg.norm(f2 - f3) <= 1e-4 * np.linalg.norm(f3)
    assert np.linalg.norm(a2 - a3) <= 1e-4 * np.linalg.norm(a3)


def test_cat2amln_with_galaxies_is_linear_in_the_catalogue():
    rng = np.random.default_rng(5)
    am = om.AnlmModes(2, 3, 500.0, 1000.0, nside=4)
    ngal = 200
    rtp = np.stack([rng.uniform(500, 1000, ngal), np.arccos(rng.uniform(-1, 1, ngal)), rng.uniform(0, 2 * math.pi, ngal)])
    w = rng.random(ngal)
    zero = np.zeros((12 * 16, am.lmax + 1, am.nmax))
    a = oc.cat2amln(rtp, am, 1e-3, zero, w)
    b = oc.cat2amln(rtp[:, :120], am, 1e-3, zero, w[:120]) + oc.cat2amln(rtp[:, 120:], am, 1e-3, zero, w[120:])
    assert np.linalg.norm(a - b) <= 1e-12 * np.linalg.norm(a)
    cl = oc.amln2clnn(a, a, om.ClnnModes(am))
    assert cl.shape == (om.ClnnModes(am).lnn.shape[1],) and np.all(np.isfinite(cl))
