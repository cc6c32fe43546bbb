"""Oracle coupling matrix against the reference's identities and analytic results."""
import math

import numpy as np
import pytest

from conftest import relerr
from oracle import cref
from oracle import modes as om
from oracle import windows as ow


def test_wigner3j000():
    from sympy.physics.wigner import wigner_3j
    for (l, lp, L) in [(0, 0, 0), (1, 1, 2), (3, 5, 4), (6, 4, 2), (10, 10, 20), (2, 3, 4)]:
        assert abs(ow.wigner3j000(l, lp, L) - float(wigner_3j(l, lp, L, 0, 0, 0))) < 1e-14
        assert abs(cref.wigner3j000(l, lp, L) - ow.wigner3j000(l, lp, L)) < 1e-15
    assert ow.wigner3j000(2, 2, 5) == 0.0 and ow.wigner3j000(2, 2, 3) == 0.0
    for (l, L) in [(5, 7), (40, 43), (100, 7)]:   # Σ (2L1+1) w² = 1
        s = sum((2 * L1 + 1) * ow.wigner3j000(l, L, L1) ** 2 for L1 in range(abs(l - L), l + L + 1))
        assert abs(s - 1) < 1e-12


@pytest.fixture(scope="module")
def small():
    a = om.AnlmModes(2, 5, 500.0, 1000.0)
    wm = ow.ConfigurationSpaceModes(500.0, 1000.0, 100, a.nside)
    c = om.ClnnModes(a, dnmax=1)
    rng = np.random.default_rng(5)
    win = ow.make_window(wm, "ang_75", "radial", "separable")
    win.mask = rng.random(win.mask.size)
    win.mask[: win.mask.size // 2] *= 0.5
    return a, wm, c, win


def test_no_window_gives_identity():
    # test/test_windows.jl:181-213 (atol 1e-3)
    a = om.AnlmModes(3, 5, 500.0, 1000.0)
    wm = ow.ConfigurationSpaceModes(500.0, 1000.0, 1000, a.nside)
    c = om.ClnnModes(a, dnmax=1)
    win = np.ones((wm.nr, wm.npix))
    M = ow.power_win_mix(win, win, wm, c)
    assert np.allclose(M, np.eye(M.shape[0]), atol=1e-3)
    W = ow.calc_Wr_lm(win, 2 * a.lmax, a.nside)
    assert np.allclose(W[:, 0], math.sqrt(4 * math.pi), atol=1e-3) and np.abs(W[:, 1:]).max() < 1e-3


def test_literal_blocked_separable_agree(small):
    # test/test_windows.jl:403-409: independent routes agree to rtol 1e-10
    a, wm, c, win = small
    d = win.dense()
    Ml = ow.power_win_mix(d, d, wm, c, literal=True)
    assert relerr(ow.power_win_mix(d, d, wm, c), Ml) < 1e-13
    assert relerr(ow.power_win_mix(win, win, wm, c), Ml) < 1e-10
    for kw in (dict(div2Lp1=True), dict(interchange=True), dict(lnn_min=4)):
        assert relerr(ow.power_win_mix(d, d, wm, c, **kw), ow.power_win_mix(d, d, wm, c, literal=True, **kw)) < 1e-13
    s = 1 + (c.lnn[1] != c.lnn[2])
    K = Ml / ((2 * c.lnn[0] + 1) * s)[None, :]
    assert np.abs(K - K.T).max() / np.abs(K).max() < 1e-13


def test_c_port_matches_numpy(small):
    a, wm, c, win = small
    d = win.dense() * (1 + 0.1 * np.random.default_rng(1).random((wm.nr, wm.npix)))
    LMAX = 2 * a.lmax
    W = ow.optimize_Wr_lm_layout(ow.calc_Wr_lm(d, LMAX, a.nside), LMAX)
    Wl, Wc = ow.calc_Wrl_Wrl(W, W, LMAX), cref.calc_wrl_wrl(W, W, LMAX)
    assert np.abs(Wl - np.transpose(Wc, (0, 2, 1))).max() < 1e-14 * np.abs(Wl).max()
    G = ow.rsdrgnlr(a, wm)
    n = c.lnn.shape[1]
    for kw in (dict(), dict(div2Lp1=True), dict(interchange=True)):
        got = cref.calc_cmix_rows(c.lnn, np.arange(1, n + 1), G, Wc, **kw)
        assert relerr(got, ow.calc_cmix(c, G, Wl, **kw)) < 1e-13
    rows = np.array([2, 7, n])
    assert np.array_equal(cref.calc_cmix_rows(c.lnn, rows, G, Wc, nthreads=2),
                          cref.calc_cmix_rows(c.lnn, np.arange(1, n + 1), G, Wc)[rows - 1])


def test_binned(small):
    # test/test_windows.jl:578-584
    a, wm, c, win = small
    d = win.dense()
    M = ow.power_win_mix(d, d, wm, c)
    wt, v = om.bandpower_binning_weights(c, dl=2)
    bc = om.ClnnBinnedModes(wt, v, c)
    assert relerr(ow.power_win_mix_binned(d, d, wt, v, wm, bc), wt @ M @ v) < 1e-14
    assert relerr(ow.power_win_mix_binned(win, win, wt, None, wm, bc), wt @ M) < 1e-10


def _W0nn_expmrr0(wm, amodes):
    # src/utils.jl:13-41 with r0sign=-1
    rmax = wm.rmax
    r0 = -rmax / 2 / 3
    norm = math.exp(-wm.r[0] / r0)
    eR = math.exp(rmax / r0)
    nmax = amodes.nmax
    out = np.zeros((nmax, nmax))
    for n1 in range(1, nmax + 1):
        for n2 in range(1, nmax + 1):
            k1, k2 = amodes.knl[n1 - 1, 0], amodes.knl[n2 - 1, 0]
            s = (-1) ** (n1 + n2)
            p1 = r0 / (1 + (k1 + k2) ** 2 * r0 ** 2)
            p2 = r0 / (1 + (k1 - k2) ** 2 * r0 ** 2)
            out[n1 - 1, n2 - 1] = norm * (p1 * (s + eR) - p2 * (s - eR)) / rmax
    return out


def test_analytic_ell0_expmrr0():
    # test/test_windows.jl:259-300 + src/utils.jl:141-168: rmin=0, phi = exp(-r/r0), l=L=0 block analytic, rtol 1.1e-6
    a = om.AnlmModes(0.01, 0.0, 1000.0)
    wm = ow.ConfigurationSpaceModes(0.0, 1000.0, 2032, a.nside)
    c = om.ClnnModes(a)
    win = ow.make_window(wm, "radial_expmrr0", "fullsky")
    M = ow.power_win_mix(win, win, wm, c)
    W0 = _W0nn_expmrr0(wm, a)
    idx0 = np.flatnonzero(c.lnn[0] == 0)
    T1 = np.zeros((idx0.size, idx0.size))
    for x, i in enumerate(idx0):
        for y, j in enumerate(idx0):
            n1, n2 = c.lnn[1, i], c.lnn[2, i]
            N1, N2 = c.lnn[1, j], c.lnn[2, j]
            T1[x, y] = W0[n1 - 1, N1 - 1] * W0[N2 - 1, n2 - 1]
            if N1 != N2:
                T1[x, y] += W0[n1 - 1, N2 - 1] * W0[N1 - 1, n2 - 1]
    # like set_T1_ell0_expmrr0!: overwrite the l=L=0 block of a copy and compare the whole matrices
    M0 = M.copy()
    M0[np.ix_(idx0, idx0)] = T1
    assert relerr(M, M0) < 1.1e-6
    assert relerr(M[np.ix_(idx0, idx0)], T1) < 2e-6


def test_win_lnn_analytic_ell0_expmrr0():
    # test/test_windows.jl:259-297: win_lnn's l=0 entries equal the analytic W0nn for phi = exp(-r/r0), rmin=0 (rtol 1e-6)
    a = om.AnlmModes(0.01, 0.0, 1000.0)
    wm = ow.ConfigurationSpaceModes(0.0, 1000.0, 2032, a.nside)
    c = om.ClnnModes(a)
    win = ow.make_window(wm, "radial_expmrr0", "fullsky")
    wlnn = ow.win_lnn(win.dense() if hasattr(win, "dense") else win, wm, c)
    W0 = _W0nn_expmrr0(wm, a)
    got = np.zeros_like(W0)
    for i in np.flatnonzero(c.lnn[0] == 0):      # get_0nn, src/utils.jl:44-54
        n1, n2 = c.lnn[1, i], c.lnn[2, i]
        got[n1 - 1, n2 - 1] = got[n2 - 1, n1 - 1] = wlnn[i]
    assert relerr(got, W0) < 1e-6


def test_win_lnn_full_sky_is_identity():
    # Wr_00 = sqrt(4 pi) for the full sky, so W_lnn' = sum_r r^2 dr g_nl g_n'l = delta_nn' (test/test_gnl.jl:27-41)
    a = om.AnlmModes(3, 4, 500.0, 1000.0)
    wm = ow.ConfigurationSpaceModes(500.0, 1000.0, 1000, a.nside)
    c = om.ClnnModes(a)
    wlnn = ow.win_lnn(np.ones((wm.nr, wm.npix)), wm, c)
    assert np.allclose(wlnn, (c.lnn[1] == c.lnn[2]).astype(float), atol=1e-5)


def test_wigner3j_families_vs_sympy():
    # general-m families (WignerFamilies semantics, src/windows.jl:434-445) against exact sympy values
    from sympy.physics.wigner import wigner_3j
    from oracle.wigner import wigner3j_family
    for (j2, j3, m2, m3) in [(2, 3, 1, -2), (5, 5, 0, 0), (5, 5, -5, 5), (4, 4, 2, -2), (10, 7, -3, 5), (0, 0, 0, 0),
                             (1, 1, 0, 0), (6, 2, -6, 2), (12, 12, 0, 0), (3, 0, 1, 0), (9, 14, 4, -11)]:
        jmin, f = wigner3j_family(j2, j3, m2, m3)
        assert jmin == max(abs(j2 - j3), abs(m2 + m3)) and f.size == j2 + j3 - jmin + 1
        for k, v in enumerate(f):
            assert abs(v - float(wigner_3j(jmin + k, j2, j3, -m2 - m3, m2, m3))) < 1e-14
    assert wigner3j_family(2, 3, 3, 0)[1].size == 0


def test_M_equals_brute_force_from_wmix():
    # test/test_windows.jl:363-409: M from calc_wmix (general-m Gaunt sums, explicit m-sums, src/windows.jl:273-364,
    # 467-525) equals power_win_mix(win, wmodes, cmodes) at rtol 1e-10 -- an independent route to the coupling matrix
    a = om.AnlmModes(2, 5, 500.0, 1000.0)
    wm = ow.ConfigurationSpaceModes(500.0, 1000.0, 100, a.nside)
    c = om.ClnnModes(a, dnmax=1)
    rng = np.random.default_rng(3)
    mask = rng.random(wm.npix)
    mask[: wm.npix // 2] *= 0.5
    win = np.outer(np.exp(-(wm.r / (0.55 * wm.rmax)) ** 2), mask)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        wmix = ow.calc_wmix(win, wm, a)
        wmix_negm = ow.calc_wmix(win, wm, a, neg_m=True)
        M = ow.power_win_mix_from_wmix(wmix, wmix_negm, c)
        M1 = ow.power_win_mix(win, win, wm, c)
        wlnn = ow.win_lnn(win, wm, c)
    assert np.isfinite(wmix).all()
    assert relerr(M, M1) < 1e-10
    # W_lnn' = (1/(2l+1)) Σ_m W_{nlm}^{n'lm} (test/test_windows.jl:250,254: rtol 1e-3 there; exact here up to rounding)
    assert relerr(ow.sum_m_lmeqLM(wmix, c), wlnn) < 1e-10


@pytest.mark.parametrize("kw", [dict(), dict(div2Lp1=True), dict(interchange=True), dict(div2Lp1=True, interchange=True)])
def test_upper_blocks_determine_the_matrix(small, kw):
    # The algebra behind cmix_mirror_fill_kernel / cmix_unpack_mirror_kernel (csrc/cmix_regz.cu): the blocks with
    # l(i) <= l(i') determine the rest through M[i',i] = M[i,i'] f_i / f_i',
    # f = (div2Lp1 ? 1 : 2l+1) (interchange_NN' ? 1 : 1 + [n != n'])   (derivations/sfb.tex:516-517)
    a, wm, c, win = small
    M = ow.power_win_mix(win.dense(), win.dense(), wm, c, **kw)
    ell, n1, n2 = c.lnn
    f = np.ones(ell.size) * (1.0 if kw.get("div2Lp1") else (2.0 * ell + 1.0))
    f = f * (1.0 if kw.get("interchange") else (1.0 + (n1 != n2)))
    # upper-packed storage: column j keeps rows [0, rend(j)), rend = end of j's own l-block
    first = np.concatenate([[0], np.cumsum(np.bincount(ell))])
    rend = first[ell + 1]
    packed = [M[:rend[j], j].copy() for j in range(M.shape[0])]
    R = np.full_like(M, np.nan)
    for j, col in enumerate(packed):
        R[:rend[j], j] = col                                       # direct half
        lower = ell[:rend[j]] < ell[j]
        R[j, :rend[j]][lower] = col[lower] * f[:rend[j]][lower] / f[j]   # mirror image below the block diagonal
    assert np.isfinite(R).all()
    assert relerr(R, M) < 1e-12


def test_wmix_full_sky_is_identity():
    # full sky: W_{nlm}^{n'l'm'} = δ (test/test_window_chains.jl:405-415 pins δ_{ll'}δ_{mm'} at atol 1e-6; the radial part
    # is the g_nl orthonormality of test/test_gnl.jl:27-41), for m >= 0 and for the neg_m table
    a = om.AnlmModes(2, 4, 500.0, 1000.0)
    wm = ow.ConfigurationSpaceModes(500.0, 1000.0, 1000, a.nside)
    win = np.ones((wm.nr, wm.npix))
    w = ow.calc_wmix(win, wm, a)
    assert np.allclose(w, np.eye(w.shape[0]), atol=1e-5)
    wn = ow.calc_wmix(win, wm, a, neg_m=True)
    # neg_m: entry (n,l,-m ; n',l',m') -- only m = m' = 0 survives on the diagonal
    n = w.shape[0]
    expect = np.zeros((n, n))
    for i in range(1, n + 1):
        if om.getnlm(a, i)[2] == 0:
            expect[i - 1, i - 1] = 1.0
    assert np.allclose(wn, expect, atol=1e-5)
