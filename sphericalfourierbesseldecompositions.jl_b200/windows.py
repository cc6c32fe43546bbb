"""Host-side mirror of the reference's window -> coupling-matrix interface (module `Windows`, src/windows.jl).

Same names, argument order and error behaviour as the Julia generic functions; every numerical step runs in
the CUDA library behind the C ABI (include/sfb_b200.h).  There is no CPU fallback.

  ConfigurationSpaceModes(rmin, rmax, nr, nside) / (amodes, nr), window_r     src/windows.jl:80-125
  calc_Wr_lm(win, LMAX, Wnside)                                                src/windows.jl:528-545
  optimize_Wr_lm_layout(Wr_lm, LMAX)                                           src/windows.jl:589-610
  precompute_gnlr(amodes, wmodes), check_nsamp                                 src/windows.jl:548-559,901-920
  power_win_mix(win, wmodes, cmodes; ...)                                      src/windows.jl:750
  power_win_mix(win1, win2, wmodes, cmodes; div2Lp1, interchange_NN, lnn_min)  src/windows.jl:781-805,809-814
  power_win_mix(win, w̃, v, wmodes, bcmodes; ...)                               src/windows.jl:751
  power_win_mix(win1, win2, w̃, v, wmodes, bcmodes; ...)                        src/windows.jl:994-1015

Arrays follow Julia's memory layout: `win` is (nr, npix) in Fortran order (a C-ordered array is converted,
which costs a host transpose), results are returned Fortran-ordered.  `None` stands for Julia's `I`.
"""
from __future__ import annotations

import math
import warnings

import numpy as np
from scipy import sparse

from . import _lib
from .modes import AnlmModes, ClnnBinnedModes, ClnnModes, getlmsize, getlnnsize
from .separable import SeparableArray

__all__ = ["ConfigurationSpaceModes", "window_r", "calc_Wr_lm", "optimize_Wr_lm_layout", "precompute_gnlr",
           "check_nsamp", "power_win_mix", "rsdrgnlr", "set_devices", "get_devices", "pinned_empty", "calc_wmix",
           "calc_wmix_all", "solve", "power_win_mix_solve"]

LAYOUT_MMAJOR, LAYOUT_MFAST = 0, 1


def set_devices(n):
    """sfb_set_devices(n): the host-pointer calls (power_win_mix, calc_Wr_lm of a dense window) shard over the GPUs
    0..n-1 of this process — the multi-GPU form of the drop-in (it replaces the reference's pmap gather,
    src/windows.jl:834-861).  n = 1 restores the single-GPU path."""
    _lib.check(_lib.load().sfb_set_devices(int(n)))


def get_devices():
    import ctypes
    n = ctypes.c_int32(0)
    _lib.check(_lib.load().sfb_get_devices(ctypes.byref(n)))
    return int(n.value)


def pinned_empty(shape, dtype=np.float64):
    """Fortran-ordered array in page-locked host memory from sfb_host_alloc (freed with the array): what the Julia shim
    uses for results so that every GPU's device->host copy runs at full PCIe speed."""
    import ctypes
    import weakref
    lib = _lib.load()
    shape = (shape,) if np.isscalar(shape) else tuple(int(x) for x in shape)
    nbytes = int(np.prod(shape, dtype=np.int64)) * np.dtype(dtype).itemsize
    ptr = ctypes.c_void_p()
    _lib.check(lib.sfb_host_alloc(ctypes.byref(ptr), max(nbytes, 1)))
    buf = (ctypes.c_char * max(nbytes, 1)).from_address(ptr.value)
    arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape, dtype=np.int64))).reshape(shape, order="F")
    weakref.finalize(buf, lib.sfb_host_free, ptr)
    return arr


class ConfigurationSpaceModes:
    """Voxelisation scheme: nr radial bins (midpoints r, width Δr) x HEALPix pixels (src/windows.jl:80-114)."""

    def __init__(self, *args):
        if len(args) == 2 and isinstance(args[0], AnlmModes):
            amodes, nr = args
            rmin, rmax, nside = amodes.rmin, amodes.rmax, amodes.nside
        elif len(args) == 4:
            rmin, rmax, nr, nside = args
        else:
            raise TypeError("ConfigurationSpaceModes(rmin, rmax, nr, nside) or ConfigurationSpaceModes(amodes, nr)")
        self.rmin, self.rmax = float(rmin), float(rmax)
        self.nr = int(nr)
        self.dr = (self.rmax - self.rmin) / self.nr
        self.r = np.linspace(self.rmin + self.dr / 2, self.rmax - self.dr / 2, self.nr)
        self.nside = int(nside)
        self.npix = 12 * self.nside ** 2

    Δr = property(lambda self: self.dr)


def window_r(wmodes):
    """r, Δr = window_r(wmodes)   (src/windows.jl:125)"""
    return wmodes.r, wmodes.dr


# --------------------------------------------------------------------------- stage 1

def _as_julia_matrix(win):
    win = np.asarray(win, dtype=np.float64)
    if win.ndim != 2:
        raise ValueError("win must be a (nr, npix) matrix")
    return win if win.flags.f_contiguous else np.asfortranarray(win)


def calc_Wr_lm(win, LMAX, Wnside, niter=3, layout=LAYOUT_MMAJOR):
    """W_lm(r) of every radial shell: (nr, lmsize) ComplexF64, HEALPix (m-major) column order; for a
    SeparableArray, SeparableArray(phi, wlm) with name2 = "wlm"   (src/windows.jl:528-545)."""
    lib = _lib.load()
    lmsize = getlmsize(LMAX)
    if isinstance(win, SeparableArray):
        mask = np.ascontiguousarray(win.mask, dtype=np.float64)
        out = np.empty(lmsize, dtype=np.complex128)
        _lib.check(lib.sfb_calc_wlm_mask(_lib.ptr(mask), mask.size, Wnside, LMAX, niter, _lib.ptr(out)))
        return SeparableArray(win.phi, out, name1="phi", name2="wlm")
    win = _as_julia_matrix(win)
    nr, npix = win.shape
    out = np.empty((nr, lmsize), dtype=np.complex128, order="F")
    _lib.check(lib.sfb_calc_wr_lm(_lib.ptr(win), nr, npix, nr, Wnside, LMAX, niter, layout, _lib.ptr(out)))
    return out


def optimize_Wr_lm_layout(Wr_lm, LMAX):
    """Column permutation m-major -> m-fast; no-op for a SeparableArray   (src/windows.jl:589-610).
    (Inside the library this permutation is fused into the stage-1 output; this host version exists for
    callers that follow the reference's two-step sequence.)"""
    if isinstance(Wr_lm, SeparableArray):
        return Wr_lm
    l = np.repeat(np.arange(LMAX + 1), np.arange(LMAX + 1) + 1)
    m = np.concatenate([np.arange(k + 1) for k in range(LMAX + 1)])
    src = l + ((m * (2 * LMAX + 1 - m)) >> 1)
    return np.asfortranarray(Wr_lm[:, src])


def check_nsamp(amodes, nr):
    """Warn (never raise) when nr < 8(n+N) for some pair   (src/windows.jl:901-920)."""
    nmax = int(np.max(amodes.nmax_l))
    need = 8 * (2 * nmax)
    if need > nr:
        nl = np.asarray(amodes.nmax_l)
        # number of ((l,n),(L,N)) pairs with 8(n+N) > nr, from the histogram of n over the (n,l) modes
        cnt = np.array([np.sum(nl >= k) for k in range(1, nmax + 1)], dtype=np.int64)
        k = np.arange(1, nmax + 1)
        num = int(np.sum(np.outer(cnt, cnt)[8 * (k[:, None] + k[None, :]) > nr]))
        warnings.warn(f"Radial integrals unlikely to converge: num_imprecise={num} max_nr_needed={need} nr={nr}",
                      RuntimeWarning, stacklevel=3)


def precompute_gnlr(amodes, wmodes):
    """gnlr[:, n-1, l] = g_nl(r); NaN where n > nmax_l[l]   (src/windows.jl:548-559).  Fortran order."""
    from scipy import special
    r = wmodes.r
    g = amodes.basisfunctions
    gnlr = np.full((r.size, amodes.nmax, amodes.lmax + 1), np.nan, order="F")
    for l in range(amodes.lmax + 1):
        nl = int(amodes.nmax_l[l])
        q = r[:, None] * g.knl[None, :nl, l]                      # all n of this l at once
        val = g.cnl[None, :nl, l] * special.spherical_jn(l, q)
        d = g.dnl[:nl, l]
        if np.any(d != 0):
            val = val + d[None, :] * special.spherical_yn(l, np.where(q == 0, 1.0, q)) * (d != 0)[None, :]
        gnlr[:, :nl, l] = val
    check_nsamp(amodes, wmodes.nr)
    return gnlr


def rsdrgnlr(amodes, wmodes):
    """r .* √Δr .* precompute_gnlr(amodes, wmodes)   (src/windows.jl:799,1009).
    The table only depends on (amodes, rmin, rmax, nr); it is memoised on the amodes object, playing the role of
    the reference's g_nl cache (AnlmModes(...; cache=true), src/SphericalBesselGNLs.jl:558-579)."""
    r, dr = window_r(wmodes)
    key = (wmodes.rmin, wmodes.rmax, wmodes.nr)
    cache = amodes.__dict__.setdefault("_rsdrgnlr_cache", {})
    if key not in cache:
        cache.clear()
        cache[key] = np.asfortranarray(r[:, None, None] * math.sqrt(dr) * precompute_gnlr(amodes, wmodes))
    else:
        check_nsamp(amodes, wmodes.nr)
    return cache[key]


# --------------------------------------------------------------------------- power_win_mix

def _csc_args(mat, n_rows_expected=None, n_cols_expected=None):
    """(colptr, rowval, nzval, nrows, ncols) 1-based Int64 like Julia's SparseMatrixCSC; None -> I."""
    if mat is None:
        return None, None, None, None, None
    m = sparse.csc_matrix(mat)
    m.sort_indices()
    if n_rows_expected is not None and m.shape[0] != n_rows_expected:
        raise ValueError("binning matrix has the wrong number of rows")
    if n_cols_expected is not None and m.shape[1] != n_cols_expected:
        raise ValueError("binning matrix has the wrong number of columns")
    colptr = (m.indptr.astype(np.int64) + 1)
    rowval = (m.indices.astype(np.int64) + 1)
    nzval = np.ascontiguousarray(m.data, dtype=np.float64)
    return colptr, rowval, nzval, m.shape[0], m.shape[1]


def _result_buffer(shape, out):
    if out is None:
        return np.empty(shape, dtype=np.float64, order="F")
    if out.shape != shape or out.dtype != np.float64 or not out.flags.f_contiguous:
        raise ValueError(f"out must be a Fortran-ordered float64 array of shape {shape}")
    return out


def power_win_mix(*args, div2Lp1=False, interchange_NN=False, lnn_min=1, out=None):
    """Coupling matrix M (unbinned) or N = w̃ M v (binned); see the module docstring for the methods.
    `out` (not in the reference): optional preallocated Fortran-ordered result buffer, e.g. pinned memory."""
    if len(args) == 3:
        win, wmodes, cmodes = args
        args = (win, win, wmodes, cmodes)
    elif len(args) == 5:
        win, wt, v, wmodes, bcmodes = args
        args = (win, win, wt, v, wmodes, bcmodes)
    if len(args) == 4:
        win1, win2, wmodes, cmodes = args
        if not isinstance(cmodes, ClnnModes):
            raise TypeError("power_win_mix(win1, win2, wmodes, cmodes::ClnnModes)")
        if isinstance(win1, SeparableArray) and isinstance(win2, SeparableArray):
            # src/windows.jl:809-814: reuse the binned method with w̃ = v = I
            if lnn_min != 1:
                raise TypeError("power_win_mix(::SeparableArray, ...) got unsupported keyword argument lnn_min")
            return _power_win_mix_binned(win1, win2, None, None, wmodes, ClnnBinnedModes(None, None, cmodes),
                                         div2Lp1, interchange_NN, out)
        return _power_win_mix_dense(win1, win2, wmodes, cmodes, div2Lp1, interchange_NN, lnn_min, out)
    if len(args) == 6:
        win1, win2, wt, v, wmodes, bcmodes = args
        if not isinstance(bcmodes, ClnnBinnedModes):
            raise TypeError("power_win_mix(win1, win2, w̃, v, wmodes, bcmodes::ClnnBinnedModes)")
        if lnn_min != 1:
            raise TypeError("power_win_mix(..., bcmodes) got unsupported keyword argument lnn_min")
        return _power_win_mix_binned(win1, win2, wt, v, wmodes, bcmodes, div2Lp1, interchange_NN, out)
    raise TypeError("no method matching power_win_mix with %d positional arguments" % len(args))


def _mode_tables(cmodes, wmodes):
    amodes = cmodes.amodes
    G = rsdrgnlr(amodes, wmodes)
    lnn = np.asfortranarray(cmodes.lnn, dtype=np.int64)
    return amodes, G, lnn


def _power_win_mix_dense(win1, win2, wmodes, cmodes, div2Lp1, interchange_NN, lnn_min, out=None):
    lib = _lib.load()
    amodes, G, lnn = _mode_tables(cmodes, wmodes)
    w1 = _as_julia_matrix(win1)
    w2 = w1 if win2 is win1 else _as_julia_matrix(win2)
    if w1.shape != (wmodes.nr, w1.shape[1]) or w2.shape != w1.shape:
        raise ValueError("window shape does not match wmodes")
    lnnsize = lnn.shape[1]
    n = lnnsize - lnn_min + 1
    M = _result_buffer((n, n), out)
    _lib.check(lib.sfb_power_win_mix(_lib.ptr(w1), None if w2 is w1 else _lib.ptr(w2), w1.shape[0], w1.shape[1],
                                     w1.shape[0], amodes.nside, _lib.ptr(G), amodes.nmax, amodes.lmax, _lib.ptr(lnn),
                                     lnnsize, lnn_min, int(bool(div2Lp1)), int(bool(interchange_NN)), _lib.ptr(M)))
    return M


def _power_win_mix_binned(win1, win2, wt, v, wmodes, bcmodes, div2Lp1, interchange_NN, out=None):
    lib = _lib.load()
    cmodes = bcmodes.cmodes
    amodes, G, lnn = _mode_tables(cmodes, wmodes)
    lnnsize = lnn.shape[1]
    wc, wr, wv, LNN1, _ = _csc_args(wt, n_cols_expected=lnnsize)
    vc, vr, vv, _, LNN2 = _csc_args(v, n_rows_expected=lnnsize)
    LNN1 = lnnsize if LNN1 is None else LNN1   # src/windows.jl:829-830
    LNN2 = lnnsize if LNN2 is None else LNN2
    N = _result_buffer((LNN1, LNN2), out)
    common = (_lib.ptr(G), amodes.nmax, amodes.lmax, _lib.ptr(lnn), lnnsize, _lib.ptr(wc), _lib.ptr(wr), _lib.ptr(wv),
              LNN1, _lib.ptr(vc), _lib.ptr(vr), _lib.ptr(vv), LNN2, int(bool(div2Lp1)), int(bool(interchange_NN)),
              _lib.ptr(N))
    if isinstance(win1, SeparableArray):
        phi = np.ascontiguousarray(win1.phi, dtype=np.float64)
        mask = np.ascontiguousarray(win1.mask, dtype=np.float64)
        _lib.check(lib.sfb_power_win_mix_separable(_lib.ptr(phi), _lib.ptr(mask), phi.size, mask.size, amodes.nside,
                                                   *common))
    else:
        w1 = _as_julia_matrix(win1)
        _lib.check(lib.sfb_power_win_mix_binned(_lib.ptr(w1), w1.shape[0], w1.shape[1], w1.shape[0], amodes.nside,
                                                *common))
    return N


def solve(N, B):
    """X = N \\ B on the device (LU with partial pivoting, like Julia's `\\` for a square matrix): the deconvolution step
    `bcmix \\ (w̃mat * Cobs)` of docs/src/tutorial_catalog.md:93-97."""
    lib = _lib.load()
    N = np.asfortranarray(N, dtype=np.float64)
    if N.ndim != 2 or N.shape[0] != N.shape[1]:
        raise ValueError("N must be square")
    B = np.asarray(B, dtype=np.float64)
    vec = B.ndim == 1
    B2 = np.asfortranarray(B.reshape(N.shape[0], -1))
    X = np.empty_like(B2, order="F")
    _lib.check(lib.sfb_solve(_lib.ptr(N), N.shape[0], _lib.ptr(B2), B2.shape[1], _lib.ptr(X)))
    return X[:, 0] if vec else X


def power_win_mix_solve(win, wt, v, wmodes, bcmodes, B, div2Lp1=False, interchange_NN=False, return_N=False):
    """`power_win_mix(win, w̃, v, wmodes, bcmodes) \\ B` without bringing the binned coupling matrix to the host
    (docs/src/tutorial_catalog.md:93-97, test/test_windows.jl:583-584).  B: LNN or LNN x nrhs.  Not in the reference as
    one call; returns X (and N with return_N)."""
    lib = _lib.load()
    cmodes = bcmodes.cmodes
    amodes, G, lnn = _mode_tables(cmodes, wmodes)
    lnnsize = lnn.shape[1]
    wc, wr, wv, LNN1, _ = _csc_args(wt, n_cols_expected=lnnsize)
    vc, vr, vv, _, LNN2 = _csc_args(v, n_rows_expected=lnnsize)
    LNN1 = lnnsize if LNN1 is None else LNN1
    LNN2 = lnnsize if LNN2 is None else LNN2
    if LNN1 != LNN2:
        raise ValueError("the binned coupling matrix must be square")
    B = np.asarray(B, dtype=np.float64)
    vec = B.ndim == 1
    B2 = np.asfortranarray(B.reshape(LNN1, -1))
    X = np.empty_like(B2, order="F")
    N = np.empty((LNN1, LNN2), order="F") if return_N else None
    w1 = _as_julia_matrix(win)
    _lib.check(lib.sfb_power_win_mix_binned_solve(
        _lib.ptr(w1), w1.shape[0], w1.shape[1], w1.shape[0], amodes.nside, _lib.ptr(G), amodes.nmax, amodes.lmax,
        _lib.ptr(lnn), lnnsize, _lib.ptr(wc), _lib.ptr(wr), _lib.ptr(wv), LNN1, _lib.ptr(vc), _lib.ptr(vr), _lib.ptr(vv),
        LNN2, int(bool(div2Lp1)), int(bool(interchange_NN)), _lib.ptr(B2), B2.shape[1], _lib.ptr(X), _lib.ptr(N)))
    X = X[:, 0] if vec else X
    return (X, N) if return_N else X


def win_lnn(win, wmodes, cmodes):
    """win_lnn(win, wmodes, cmodes) (src/windows.jl:382-391): shot-noise window W_lnn' from Wr_00(r) of the same
    calc_Wr_lm the coupling matrix uses.  Returns a vector of lnnsize values."""
    lib = _lib.load()
    amodes, G, lnn = _mode_tables(cmodes, wmodes)
    if isinstance(win, SeparableArray):
        win = np.outer(win.phi, win.mask)
    w = _as_julia_matrix(win)
    nr, npix = w.shape
    if nr != wmodes.nr:
        raise ValueError("window has %d shells, wmodes.nr = %d" % (nr, wmodes.nr))
    out = np.empty(lnn.shape[1], dtype=np.float64)
    _lib.check(lib.sfb_win_lnn(_lib.ptr(w), nr, npix, w.strides[1] // 8, amodes.nside, _lib.ptr(G), amodes.nmax,
                               amodes.lmax, _lib.ptr(lnn), lnn.shape[1], _lib.ptr(out)))
    return out


def calc_wmix(win, wmodes, amodes, neg_m=False):
    """calc_wmix(win, wmodes, amodes; neg_m=false) (src/windows.jl:299-364): W_{nlm}^{n'l'm'} for m, m' >= 0 (or m -> -m),
    (nlmsize, nlmsize) ComplexF64 in getidx(amodes, n, l, m) order."""
    from .modes import getnlmsize
    lib = _lib.load()
    G = rsdrgnlr(amodes, wmodes)
    if isinstance(win, SeparableArray):
        win = np.outer(win.phi, win.mask)
    w = _as_julia_matrix(win)
    nr, npix = w.shape
    if nr != wmodes.nr:
        raise ValueError("window has %d shells, wmodes.nr = %d" % (nr, wmodes.nr))
    n = getnlmsize(amodes)
    out = np.empty((n, n), dtype=np.complex128, order="F")
    nmax_l = np.ascontiguousarray(amodes.nmax_l, dtype=np.int64)
    lmax_n = np.ascontiguousarray(amodes.lmax_n, dtype=np.int64)
    _lib.check(lib.sfb_calc_wmix(_lib.ptr(w), nr, npix, w.strides[1] // 8, amodes.nside, _lib.ptr(G), amodes.nmax,
                                 amodes.lmax, _lib.ptr(nmax_l), _lib.ptr(lmax_n), int(bool(neg_m)), _lib.ptr(out)))
    return out


def calc_wmix_all(win, wmodes, amodes):
    """calc_wmix_all (src/window_chains.jl:578-582): (wmix, wmix_negm).  A SeparableArray is materialised (the reference
    takes a different, mathematically equal route through WindowChainsCacheSeparableWmix, :585-589)."""
    return calc_wmix(win, wmodes, amodes), calc_wmix(win, wmodes, amodes, neg_m=True)


def power_win_mix_from_wrlm(W1r_lm, W2r_lm, wmodes, cmodes, layout=LAYOUT_MMAJOR, div2Lp1=False, interchange_NN=False,
                            lnn_min=1):
    """calc_Wrl_Wrl + calc_cmix from precomputed W_lm(r) (src/windows.jl:796-801): the stage-2/3 entry point."""
    lib = _lib.load()
    amodes, G, lnn = _mode_tables(cmodes, wmodes)
    LMAX = 2 * amodes.lmax
    w1 = np.asfortranarray(W1r_lm, dtype=np.complex128)
    w2 = w1 if (W2r_lm is None or W2r_lm is W1r_lm) else np.asfortranarray(W2r_lm, dtype=np.complex128)
    if w1.shape != (wmodes.nr, getlmsize(LMAX)):
        raise ValueError("Wr_lm must be (nr, lmsize(2*lmax))")
    lnnsize = lnn.shape[1]
    n = lnnsize - lnn_min + 1
    M = np.empty((n, n), dtype=np.float64, order="F")
    _lib.check(lib.sfb_power_win_mix_from_wrlm(_lib.ptr(w1), None if w2 is w1 else _lib.ptr(w2), w1.shape[0], LMAX,
                                               layout, _lib.ptr(G), amodes.nmax, amodes.lmax, _lib.ptr(lnn), lnnsize,
                                               lnn_min, int(bool(div2Lp1)), int(bool(interchange_NN)), _lib.ptr(M)))
    return M
