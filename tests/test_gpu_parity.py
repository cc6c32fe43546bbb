"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle on identical inputs.

Tolerance: north_star asks <= 1e-10 relative (FP64), norm-wise like the reference's `≈ rtol=1e-10`
(test/test_windows.jl:408-409,434,440-442); index tables bit-exact.
"""
import numpy as np
import pytest

from conftest import relerr
from oracle import healpix as ohp
from oracle import modes as om
from oracle import windows as ow

pytestmark = pytest.mark.gpu
RTOL = 1e-10


def _setup(nmax=2, lmax=5, nr=100, dnmax=1, seed=5, kmax=None):
    import sfb_b200 as sfb
    if kmax is None:
        oa, a = om.AnlmModes(nmax, lmax, 500.0, 1000.0), sfb.AnlmModes(nmax, lmax, 500.0, 1000.0)
    else:
        oa, a = om.AnlmModes(kmax, 500.0, 1000.0), sfb.AnlmModes(kmax, 500.0, 1000.0)
    owm = ow.ConfigurationSpaceModes(500.0, 1000.0, nr, oa.nside)
    wm = sfb.ConfigurationSpaceModes(500.0, 1000.0, nr, a.nside)
    oc = om.ClnnModes(oa, dnmax=dnmax) if dnmax is not None else om.ClnnModes(oa)
    c = sfb.ClnnModes(a, dnmax=dnmax)
    assert np.array_equal(oc.lnn, c.lnn)
    rng = np.random.default_rng(seed)
    return sfb, oa, a, owm, wm, oc, c, rng


def _random_window(rng, owm, smooth=True):
    mask = rng.random(owm.npix)
    mask[: owm.npix // 2] *= 0.5
    phi = np.exp(-(owm.r / (0.55 * owm.rmax)) ** 2)
    win = np.outer(phi, mask)
    if not smooth:
        win *= 1 + 0.3 * rng.random(win.shape)
    return win / win.max(), phi, mask


# --------------------------------------------------------------------------------------------- stage 1

@pytest.mark.parametrize("nside,lmax,nr", [(4, 8, 3), (8, 16, 10), (8, 32, 8), (16, 40, 17), (32, 64, 5), (1, 2, 2),
                                           (6, 12, 5), (12, 30, 9), (64, 130, 3)])  # 6, 12: not powers of two
def test_calc_wr_lm_matches_oracle(nside, lmax, nr):
    import sfb_b200 as sfb
    rng = np.random.default_rng(nside + lmax)
    win = rng.random((nr, 12 * nside * nside))
    win[:, ::3] = 0.0
    ref = ow.calc_Wr_lm(win, lmax, nside)
    got = sfb.calc_Wr_lm(win, lmax, nside)
    assert got.shape == ref.shape
    assert relerr(got, ref) < RTOL
    got_fast = sfb.calc_Wr_lm(win, lmax, nside, layout=1)
    assert relerr(got_fast, ow.optimize_Wr_lm_layout(ref, lmax)) < RTOL
    assert np.array_equal(sfb.optimize_Wr_lm_layout(got, lmax), got_fast)


@pytest.mark.parametrize("niter", [0, 1, 3])
def test_calc_wr_lm_niter(niter):
    import sfb_b200 as sfb
    rng = np.random.default_rng(11)
    win = rng.random((4, 12 * 8 * 8))
    ref = ow.calc_Wr_lm(win, 16, 8, niter=niter)
    assert relerr(sfb.calc_Wr_lm(win, 16, 8, niter=niter), ref) < RTOL


@pytest.mark.parametrize("nside_in,nside_out", [(4, 16), (16, 8), (8, 8), (2, 32)])
def test_calc_wr_lm_udgrade(nside_in, nside_out):
    import sfb_b200 as sfb
    rng = np.random.default_rng(nside_in * 100 + nside_out)
    win = rng.random((3, 12 * nside_in * nside_in))
    lmax = 2 * nside_out
    ref = ow.calc_Wr_lm(win, lmax, nside_out)
    assert relerr(sfb.calc_Wr_lm(win, lmax, nside_out), ref) < RTOL


def test_calc_wr_lm_full_sky_and_monopole():
    # test/test_windows.jl:19-37: Wr_lm[:,1] ≈ sqrt(4π) mean(win, dims=2), atol 1e-3
    import sfb_b200 as sfb
    nside, nr = 16, 12
    wm = ow.ConfigurationSpaceModes(500.0, 1000.0, nr, nside)
    win = ow.make_window(wm, "ang_quarter", "radial")
    got = sfb.calc_Wr_lm(win, 2 * nside, nside)
    assert np.allclose(got[:, 0].real, np.sqrt(4 * np.pi) * win.mean(axis=1), atol=1e-3)
    full = sfb.calc_Wr_lm(np.ones((2, 12 * nside * nside)), 2 * nside, nside)
    assert abs(full[0, 0] - np.sqrt(4 * np.pi)) < 1e-4 and np.abs(full[:, 1:]).max() < 1e-3


def test_calc_wr_lm_separable_and_errors():
    import sfb_b200 as sfb
    from sfb_b200 import _lib
    rng = np.random.default_rng(3)
    mask, phi = rng.random(12 * 8 * 8), rng.random(6)
    out = sfb.calc_Wr_lm(sfb.SeparableArray(phi, mask), 16, 8)
    _, ref = ow.calc_Wr_lm(ow.SeparableArray(phi, mask), 16, 8)
    assert relerr(out.wlm, ref) < RTOL and np.array_equal(out.phi, phi)
    with pytest.raises(_lib.SFBError, match="poor choice"):   # src/healpix_helpers.jl:60-63
        sfb.calc_Wr_lm(np.ones((2, 12 * 8 * 8)), 4 * 8 + 1, 8)
    with pytest.raises(_lib.SFBError):
        sfb.calc_Wr_lm(np.ones((2, 100)), 4, 8)


# --------------------------------------------------------------------------------------------- stage 2+3

@pytest.mark.parametrize("kw", [dict(), dict(div2Lp1=True), dict(interchange_NN=True), dict(lnn_min=5),
                                dict(div2Lp1=True, interchange_NN=True, lnn_min=3)])
def test_cmix_from_wrlm_matches_oracle(kw):
    sfb, oa, a, owm, wm, oc, c, rng = _setup()
    win, _, _ = _random_window(rng, owm)
    LMAX = 2 * oa.lmax
    W = ow.optimize_Wr_lm_layout(ow.calc_Wr_lm(win, LMAX, oa.nside), LMAX)
    okw = dict(div2Lp1=kw.get("div2Lp1", False), interchange=kw.get("interchange_NN", False), lnn_min=kw.get("lnn_min", 1))
    ref = ow.calc_cmix(oc, ow.rsdrgnlr(oa, owm), ow.calc_Wrl_Wrl(W, W, LMAX), **okw)
    got = sfb.power_win_mix_from_wrlm(W, None, wm, c, layout=1, **kw)
    assert got.shape == ref.shape
    assert relerr(got, ref) < RTOL


# every tile class of the block kernels: nmax_l = 5 / 12 / 20 / 26 / 32 (1..4 row tiles of 8), 24 / 40 / 64 shells
# (register-Z kernel with 4 or 8 radial tiles) and 72 shells (shared-memory-Z kernel), dense lnn tables
# (20, 1, 400): long radial grid (the reference's recommended nr >= 8(n+N)): G_L rows stay in global memory
# (37, 1, 40), (48, 1, 64), (70, 0, 24): nmax_l > 32 on the register-Z kernel — rows as virtual blocks (panel pairs of 16),
# more than 32 column-side N per block; (40, 1, 72): the same on the shared-memory-Z kernel
@pytest.mark.parametrize("nmax,lmax,nr", [(5, 3, 24), (12, 3, 40), (20, 2, 64), (26, 2, 40), (32, 1, 64), (26, 1, 72),
                                          (20, 1, 400), (9, 2, 200), (37, 1, 40), (48, 1, 64), (70, 0, 24), (40, 1, 72)])
@pytest.mark.parametrize("kw", [dict(), dict(div2Lp1=True, interchange_NN=True)])
def test_cmix_tile_classes(nmax, lmax, nr, kw):
    import warnings
    from oracle import cref
    sfb, oa, a, owm, wm, oc, c, rng = _setup(nmax=nmax, lmax=lmax, nr=nr, dnmax=None)
    assert max(oa.nmax_l) == nmax
    LMAX = 2 * oa.lmax
    lmsize = (LMAX + 1) * (LMAX + 2) // 2
    W = rng.standard_normal((nr, lmsize)) + 1j * rng.standard_normal((nr, lmsize))   # m-fast W_lm(r), any values
    for l in range(LMAX + 1):
        W[:, l * (l + 1) // 2] = W[:, l * (l + 1) // 2].real       # m = 0 coefficients of a real map are real
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        G = ow.rsdrgnlr(oa, owm)
        got = sfb.power_win_mix_from_wrlm(W, None, wm, c, layout=1, **kw)
    Wc = cref.calc_wrl_wrl(W, W, LMAX)
    n = oc.lnn.shape[1]
    ref = cref.calc_cmix_rows(oc.lnn, np.arange(1, n + 1), G, Wc, div2Lp1=kw.get("div2Lp1", False),
                              interchange=kw.get("interchange_NN", False))
    assert got.shape == (n, n)
    assert relerr(got, ref) < RTOL


@pytest.mark.parametrize("nmax,lmax,nr", [(3, 4, 37), (35, 1, 24)])   # the second: cross-correlation with nmax_l > 32
@pytest.mark.parametrize("interchange", [False, True])
def test_cmix_two_windows(interchange, nmax, lmax, nr):
    sfb, oa, a, owm, wm, oc, c, rng = _setup(nmax=nmax, lmax=lmax, nr=nr, dnmax=None)
    win1, _, _ = _random_window(rng, owm, smooth=False)
    win2, _, _ = _random_window(rng, owm, smooth=False)
    LMAX = 2 * oa.lmax
    W1 = ow.calc_Wr_lm(win1, LMAX, oa.nside)
    W2 = ow.calc_Wr_lm(win2, LMAX, oa.nside)
    Wl = ow.calc_Wrl_Wrl(ow.optimize_Wr_lm_layout(W1, LMAX), ow.optimize_Wr_lm_layout(W2, LMAX), LMAX)
    ref = ow.calc_cmix(oc, ow.rsdrgnlr(oa, owm), Wl, interchange=interchange)
    got = sfb.power_win_mix_from_wrlm(W1, W2, wm, c, layout=0, interchange_NN=interchange)
    assert relerr(got, ref) < RTOL


def test_power_win_mix_end_to_end_kmax():
    # kmax-selected modes: ragged nmax_l, several a-tile classes
    sfb, oa, a, owm, wm, oc, c, rng = _setup(kmax=0.03, nr=24, dnmax=None)
    win, _, _ = _random_window(rng, owm, smooth=False)
    ref = ow.power_win_mix(win, win, owm, oc)
    with pytest.warns(RuntimeWarning):   # nr < 8(n+N): the reference only warns (src/windows.jl:916-919)
        got = sfb.power_win_mix(win, wm, c)
    assert relerr(got, ref) < RTOL
    s = 1 + (c.lnn[1] != c.lnn[2])
    K = got / ((2 * c.lnn[0] + 1) * s)[None, :]
    assert np.abs(K - K.T).max() / np.abs(K).max() < 1e-11   # symmetry identity, SURVEY §8c.5


def test_win_lnn_matches_oracle():
    # win_lnn (src/windows.jl:382-418) on the Wr_00 of the same stage 1
    from sfb_b200 import _lib
    sfb, oa, a, owm, wm, oc, c, rng = _setup(kmax=0.03, nr=40, dnmax=None)
    win, _, _ = _random_window(rng, owm, smooth=False)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        got = sfb.win_lnn(win, wm, c)
    ref = ow.win_lnn(win, owm, oc)
    assert got.shape == ref.shape == (oc.lnn.shape[1],)
    assert relerr(got, ref) < RTOL
    with pytest.raises(_lib.SFBError, match="DomainError"):      # sqrt of a negative Wr_00 throws in the reference
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            sfb.win_lnn(-win, wm, c)


def test_no_window_is_identity():
    # test/test_windows.jl:181-213: full sky => M ≈ I, atol 1e-3
    sfb, oa, a, owm, wm, oc, c, rng = _setup(nmax=3, lmax=5, nr=1000)
    win = np.ones((owm.nr, owm.npix))
    M = sfb.power_win_mix(win, wm, c)
    assert np.allclose(M, np.eye(M.shape[0]), atol=1e-3)
    assert relerr(M, ow.power_win_mix(win, win, owm, oc)) < RTOL


def test_sep_insep_binned():
    # test/test_windows.jl:365-443: M(dense) ≈ M(separable), binned variants, rtol 1e-10
    sfb, oa, a, owm, wm, oc, c, rng = _setup()
    win, phi, mask = _random_window(rng, owm)
    phi = phi / win.max() if False else phi
    swin = sfb.SeparableArray(phi, mask)
    dense = swin.dense()
    M1 = sfb.power_win_mix(swin, swin, wm, c)
    M2 = sfb.power_win_mix(dense, wm, c)
    ref = ow.power_win_mix(dense, dense, owm, oc)
    assert relerr(M2, ref) < RTOL
    assert relerr(M1, M2) < RTOL
    for kw in (dict(div2Lp1=True), dict(interchange_NN=True)):
        assert relerr(sfb.power_win_mix(swin, swin, wm, c, **kw), sfb.power_win_mix(dense, wm, c, **kw)) < RTOL
    wt, v = sfb.bandpower_binning_weights(c, dl=2)
    bc = sfb.ClnnBinnedModes(wt, v, c)
    N1 = sfb.power_win_mix(swin, wt, v, wm, bc)
    N2 = sfb.power_win_mix(dense, wt, v, wm, bc)
    assert N1.shape == (wt.shape[0], v.shape[1])
    assert relerr(N2, wt @ M2 @ v) < RTOL           # test/test_windows.jl:578-579
    assert relerr(N1, N2) < RTOL
    assert relerr(sfb.power_win_mix(dense, wt, None, wm, bc), wt @ M2) < RTOL     # w̃M, test/test_windows.jl:582
    assert relerr(sfb.power_win_mix(dense, None, v, wm, bc), M2 @ v) < RTOL
    assert relerr(sfb.power_win_mix(swin, wt, None, wm, bc), wt @ M2) < RTOL
    oref = ow.power_win_mix_binned(dense, dense, wt.toarray(), v.toarray(), owm, om.ClnnBinnedModes(wt.toarray(), v.toarray(), oc))
    assert relerr(N2, oref) < RTOL


def test_device_pipeline_row_shards():
    import torch
    from sfb_b200.device import DevicePipeline, shard_rows
    sfb, oa, a, owm, wm, oc, c, rng = _setup(kmax=0.03, nr=24, dnmax=None)
    win, _, _ = _random_window(rng, owm, smooth=False)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        G = sfb.rsdrgnlr(a, wm)
        full = sfb.power_win_mix(win, wm, c)
    pipe = DevicePipeline(wm, c, G)
    d_win = torch.from_numpy(np.ascontiguousarray(win.T)).cuda()
    M = pipe.power_win_mix(d_win).cpu().numpy().T
    assert relerr(M, full) < 1e-13
    ranges = shard_rows(pipe.row_costs, pipe.ell_of_row, 3)
    assert ranges[0][0] == 0 and ranges[-1][1] == pipe.nout
    parts = [pipe.power_win_mix_rows(lo, hi).cpu().numpy().T for lo, hi in ranges]
    assert relerr(np.concatenate(parts, axis=0), M) < 1e-13    # full call uses the mirrored (L >= l) path
    # an arbitrary range that cuts through an l-block
    lo, hi = 7, pipe.nout - 5
    assert relerr(pipe.power_win_mix_rows(lo, hi).cpu().numpy().T, M[lo:hi]) < 1e-13
    cr = shard_rows(pipe.col_costs, pipe.ell_of_row, 3)
    cparts = [pipe.power_win_mix_cols(lo, hi).cpu().numpy().T for lo, hi in cr]     # each (nout, hi-lo)
    assert relerr(np.concatenate(cparts, axis=1), M) < 1e-13
    wc = pipe.wr_lm_complex().cpu().numpy().T
    assert relerr(wc, sfb.calc_Wr_lm(win, 2 * a.lmax, a.nside)) < 1e-13
    # upper-packed exchange format: three column shards into one packed buffer, then unpack + mirror
    off = pipe.packed_offsets()
    assert off[0] == 0 and off[-1] < pipe.nout ** 2 and np.all(np.diff(off) > 0)
    ur = pipe.packed_shard_ranges(3)
    for kw in (dict(), dict(div2Lp1=True), dict(interchange_NN=True)):
        packed = torch.full((int(off[-1]),), float("nan"), dtype=torch.float64, device="cuda")
        for lo, hi in ur:
            pipe.power_win_mix_upper_packed(lo, hi, packed, **kw)
        got = pipe.unpack_mirror(packed, **kw).cpu().numpy().T
        ref = pipe.power_win_mix_rows(0, pipe.nout, **kw).cpu().numpy().T
        assert np.isfinite(got).all()
        assert relerr(got, ref) < 1e-13
        # "L-shaped" shards: upper-packed columns + locally mirrored rows cover every element exactly once
        for world in (1, 3):
            lr = pipe.packed_shard_ranges(world, balance="cost")
            packed = torch.full((int(off[-1]),), float("nan"), dtype=torch.float64, device="cuda")
            rows = [pipe.power_win_mix_lshard(lo, hi, packed, **kw)[1] for lo, hi in lr]
            got = pipe.lshard_assemble_host(packed, rows, lr)
            assert np.isfinite(got).all()
            assert relerr(got, ref) < 1e-13
    pipe.close()


# --------------------------------------------------------------------------------------------- alternative code paths
# the kernels behind the environment switches are exercised here (the switches are read at call
# time), each against the oracle, so none of them is dead code on the GPU box.

@pytest.mark.parametrize("flag", ["SFB_NO_MIRROR", "SFB_CMIX_OLD", "SFB_WHAT_FMA", "SFB_WL_FMA", "SFB_REGZ_CPASYNC",
                                  "SFB_REGZ_FULLDIAG", "SFB_REGZ_ONEBLOCK", "SFB_FILL_OVERLAP", "SFB_FILL_2D"])
@pytest.mark.parametrize("nr", [24, 64])
def test_stage23_alternative_paths(monkeypatch, flag, nr):
    import warnings
    sfb, oa, a, owm, wm, oc, c, rng = _setup(kmax=0.03, nr=nr, dnmax=None)
    win, _, _ = _random_window(rng, owm, smooth=False)
    ref = ow.power_win_mix(win, win, owm, oc)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        monkeypatch.setenv("SFB_REGZ_TMA_STRICT", "1")          # the default TMA staging must not silently fall back
        base = sfb.power_win_mix(win, wm, c)
        monkeypatch.setenv(flag, "1")
        got = sfb.power_win_mix(win, wm, c)
    assert relerr(got, ref) < RTOL
    assert relerr(got, base) < 1e-12


@pytest.mark.parametrize("nside,lmax,nr", [(8, 16, 5), (16, 40, 9), (32, 64, 3), (12, 30, 4)])
def test_stage1_pixel_space_refinement_path(monkeypatch, nside, lmax, nr):
    """SFB_SHT_PIXEL_ITER=1: Jacobi refinement through pixel space (synthesis kernels + residual + re-analysis, the
    literal Healpix.jl iteration) instead of the ring-Fourier alias operator."""
    import sfb_b200 as sfb
    rng = np.random.default_rng(3 * nside + lmax)
    win = rng.random((nr, 12 * nside * nside))
    win[:, ::5] = 0.0
    ref = ow.calc_Wr_lm(win, lmax, nside)
    base = sfb.calc_Wr_lm(win, lmax, nside)
    monkeypatch.setenv("SFB_SHT_PIXEL_ITER", "1")
    got = sfb.calc_Wr_lm(win, lmax, nside)
    assert relerr(got, ref) < RTOL
    assert relerr(got, base) < 1e-11


@pytest.mark.parametrize("flag", ["SFB_SHT_NO_GRAM", "SFB_ALIAS_OLD", "SFB_SHT_NO_MLIM"])
@pytest.mark.parametrize("nside,lmax,nr", [(16, 24, 9), (32, 40, 3), (64, 100, 2)])
def test_stage1_ring_space_variants(monkeypatch, flag, nside, lmax, nr):
    """The ring-Fourier Jacobi pass in its three forms must agree: Gram matrices for the alias-free rings + shared-memory
    alias kernel (default), every ring through synthesis/analysis (SFB_SHT_NO_GRAM, read at plan creation), and the
    per-(ring, m) alias kernel (SFB_ALIAS_OLD)."""
    import sfb_b200 as sfb
    rng = np.random.default_rng(nside + lmax)
    win = rng.random((nr, 12 * nside * nside))
    win[:, ::3] = 0.0
    ref = ow.calc_Wr_lm(win, lmax, nside) if nside <= 32 else None
    base = sfb.calc_Wr_lm(win, lmax, nside)
    monkeypatch.setenv(flag, "1")
    sfb.calc_Wr_lm(win[:1, :12 * 4 * 4].copy(), 4, 4)          # evict the cached plan so that it is rebuilt under the flag
    got = sfb.calc_Wr_lm(win, lmax, nside)
    monkeypatch.delenv(flag)
    sfb.calc_Wr_lm(win[:1, :12 * 4 * 4].copy(), 4, 4)
    if ref is not None:
        assert relerr(base, ref) < RTOL and relerr(got, ref) < RTOL
    assert relerr(got, base) < 1e-12


def test_nmax_l_above_32_every_path():
    """Tables with nmax_l > 32 (VERDICT r1 missing #7): the separable-window coupling matrix and win_lnn go through the plan
    directly; the dense window runs the l-blocks as virtual row blocks of <= 32 basis functions (panel pairs)."""
    import warnings
    from sfb_b200 import _lib
    sfb, oa, a, owm, wm, oc, c, rng = _setup(nmax=34, lmax=1, nr=40, dnmax=None)
    assert max(oa.nmax_l) == 34
    win, phi, mask = _random_window(rng, owm)
    swin = sfb.SeparableArray(phi, mask)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        M = sfb.power_win_mix(swin, swin, wm, c)
        ref = ow.power_win_mix(swin.dense(), swin.dense(), owm, oc)
        assert relerr(M, ref) < RTOL
        assert relerr(sfb.win_lnn(swin.dense(), wm, c), ow.win_lnn(swin.dense(), owm, oc)) < RTOL
        dense = sfb.power_win_mix(swin.dense(), wm, c)
        assert relerr(dense, ref) < RTOL
        # the multi-GPU output forms on the same table: column slabs and L-shaped shards (upper-packed + mirrored rows)
        import torch
        from sfb_b200.device import DevicePipeline, shard_rows
        pipe = DevicePipeline(wm, c, sfb.rsdrgnlr(a, wm))
        pipe.calc_wr_lm(torch.from_numpy(np.ascontiguousarray(swin.dense().T)).cuda())
        cr = shard_rows(pipe.col_costs, pipe.ell_of_row, 2)
        cols = np.concatenate([pipe.power_win_mix_cols(lo, hi).cpu().numpy().T for lo, hi in cr], axis=1)
        assert relerr(cols, dense) < 1e-12
        off = pipe.packed_offsets()
        lr = pipe.packed_shard_ranges(2, balance="lshard")
        packed = torch.full((int(off[-1]),), float("nan"), dtype=torch.float64, device="cuda")
        rows = [pipe.power_win_mix_lshard(lo, hi, packed)[1] for lo, hi in lr]
        assert relerr(pipe.lshard_assemble_host(packed, rows, lr), dense) < 1e-12
        pipe.close()
