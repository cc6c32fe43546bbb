"""Host-side mirror of the reference's SFB transforms next to the window path (module `Cat2Anlm`, src/cat2anlm.jl, and
`win_rhat_ln`, src/windows.jl:244-270) — SURVEY §8f row 3.  Same names and argument order; the transforms run in the CUDA
library (sfb_field2anlm, sfb_anlm2field, sfb_win_rhat_ln, sfb_cat2amln).  What stays on the host is what stays in Julia in
the real drop-in: evaluating g_nl at the radii, HEALPix pixel lookup of the galaxies, sorting."""
from __future__ import annotations

import math

import numpy as np

from . import _lib
from .modes import getnlmsize
from .separable import SeparableArray
from .windows import _as_julia_matrix, precompute_gnlr, window_r

__all__ = ["field2anlm", "anlm2field", "win_rhat_ln", "cat2amln", "ang2pix_ring", "amln2clnn"]


def _tables(amodes):
    return (np.ascontiguousarray(amodes.nmax_l, dtype=np.int64), np.ascontiguousarray(amodes.lmax_n, dtype=np.int64))


def _gnlr_quiet(amodes, wmodes):
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)      # these callers use check_nsamp_1gnl, not the pair check
        return precompute_gnlr(amodes, wmodes)


def field2anlm(f_xyz, wmodes, amodes):
    """field2anlm(f_xyz, wmodes, amodes) (src/cat2anlm.jl:326-373): SFB coefficients of a real field (nr x npix)."""
    lib = _lib.load()
    f = _as_julia_matrix(f_xyz)
    if f.shape != (wmodes.nr, wmodes.npix):
        raise ValueError("field must be (nr, npix) of wmodes")
    r, dr = window_r(wmodes)
    T = np.asfortranarray(_gnlr_quiet(amodes, wmodes) * (r ** 2 * dr)[:, None, None])
    nmax_l, lmax_n = _tables(amodes)
    out = np.empty(getnlmsize(amodes), dtype=np.complex128)
    _lib.check(lib.sfb_field2anlm(_lib.ptr(f), f.shape[0], f.shape[1], f.shape[0], _lib.ptr(T), amodes.nmax, amodes.lmax,
                                  _lib.ptr(nmax_l), _lib.ptr(lmax_n), _lib.ptr(out)))
    return out


def anlm2field(f_nlm, wmodes, amodes):
    """anlm2field(f_nlm, wmodes, amodes) (src/cat2anlm.jl:385-422): nr x npix real field at amodes.nside."""
    lib = _lib.load()
    f = np.ascontiguousarray(f_nlm, dtype=np.complex128)
    if f.shape != (getnlmsize(amodes),):
        raise ValueError("f_nlm must have getnlmsize(amodes) entries")
    if wmodes.npix != 12 * amodes.nside ** 2:
        raise ValueError("wmodes.npix must match amodes.nside")           # f_xyz[ir,:] .= hpmap would throw
    g = np.asfortranarray(_gnlr_quiet(amodes, wmodes))
    nmax_l, lmax_n = _tables(amodes)
    out = np.empty((wmodes.nr, wmodes.npix), dtype=np.float64, order="F")
    _lib.check(lib.sfb_anlm2field(_lib.ptr(f), _lib.ptr(g), wmodes.nr, amodes.nside, amodes.nmax, amodes.lmax,
                                  _lib.ptr(nmax_l), _lib.ptr(lmax_n), _lib.ptr(out), wmodes.nr))
    return out


def win_rhat_ln(win, wmodes, amodes):
    """win_rhat_ln(win, wmodes, amodes) (src/windows.jl:244-270): (npix, lmax+1, nmax) array, NaN where l > lmax_n[n];
    for a SeparableArray: SeparableArray(mask, W_ln) with names (mask, w_ln) — a few dot products, done on the host."""
    r, dr = window_r(wmodes)
    T = np.asfortranarray(_gnlr_quiet(amodes, wmodes) * (r ** 2 * dr)[:, None, None])
    if isinstance(win, SeparableArray):
        W_ln = np.einsum("r,rnl->ln", np.asarray(win.phi, dtype=np.float64), np.nan_to_num(T))
        W_ln[np.isnan(T[0]).T] = np.nan
        return SeparableArray(win.mask, W_ln.ravel(order="F"), name1="mask", name2="w_ln")
    lib = _lib.load()
    w = _as_julia_matrix(win)
    if w.shape[0] != wmodes.nr:
        raise ValueError("window has %d shells, wmodes.nr = %d" % (w.shape[0], wmodes.nr))
    nmax_l, lmax_n = _tables(amodes)
    out = np.empty((w.shape[1], amodes.lmax + 1, amodes.nmax), dtype=np.float64, order="F")
    _lib.check(lib.sfb_win_rhat_ln(_lib.ptr(w), w.shape[0], w.shape[1], w.shape[0], _lib.ptr(T), amodes.nmax, amodes.lmax,
                                   _lib.ptr(nmax_l), _lib.ptr(lmax_n), _lib.ptr(out)))
    return out


def ang2pix_ring(nside, theta, phi):
    """0-based RING pixel of (θ, φ) (Healpix.ang2pixRing minus one)."""
    theta = np.atleast_1d(np.asarray(theta, dtype=np.float64))
    phi = np.atleast_1d(np.asarray(phi, dtype=np.float64))
    z, tt = np.cos(theta), np.mod(phi, 2 * math.pi) / (math.pi / 2)
    za = np.abs(z)
    nl4, npix, ncap = 4 * nside, 12 * nside * nside, 2 * nside * (nside - 1)
    pix = np.empty(z.shape, dtype=np.int64)
    eq = za <= 2.0 / 3.0
    t1, t2 = nside * (0.5 + tt[eq]), nside * 0.75 * z[eq]
    jp, jm = np.floor(t1 - t2).astype(np.int64), np.floor(t1 + t2).astype(np.int64)
    ir = nside + 1 + jp - jm
    ip = np.mod((jp + jm - nside + (1 - (ir & 1)) + 1) // 2, nl4)
    pix[eq] = ncap + (ir - 1) * nl4 + ip
    po = ~eq
    tp = tt[po] - np.floor(tt[po])
    tmp = nside * np.sqrt(3 * (1 - za[po]))
    jp, jm = np.floor(tp * tmp).astype(np.int64), np.floor((1.0 - tp) * tmp).astype(np.int64)
    ir = jp + jm + 1
    ip = np.mod(np.floor(tt[po] * ir).astype(np.int64), 4 * ir)
    pix[po] = np.where(z[po] > 0, 2 * ir * (ir - 1) + ip, npix - 2 * ir * (ir + 1) + ip)
    return pix


def cat2amln(rtp, amodes, nbar, wrhatln, weight=None, batch=64):
    """cat2amln(rθϕ, amodes, nbar, win_rhat_ln, weights) (src/cat2anlm.jl:257-314).  rtp: 3 x Ngal.  The (n, l) modes are
    transformed in batches of `batch` maps per map2alm call (the reference loops one (n, l) at a time)."""
    lib = _lib.load()
    rtp = np.asarray(rtp, dtype=np.float64).reshape(3, -1)
    ngal = rtp.shape[1]
    weight = np.ones(ngal) if (weight is None or len(weight) == 0) else np.asarray(weight, dtype=np.float64)
    p = np.argsort(rtp[0], kind="stable")                                  # sortout (:48-68): by r
    r, theta, phi = rtp[0, p], rtp[1, p], rtp[2, p]
    weight = weight[p] if ngal else weight
    nside, lmax, nmax = amodes.nside, amodes.lmax, amodes.nmax
    npix = 12 * nside * nside
    wr = np.asfortranarray(wrhatln, dtype=np.float64)
    if wr.shape != (npix, lmax + 1, nmax):
        raise ValueError("win_rhat_ln must be (npix, lmax+1, nmax)")
    pix = ang2pix_ring(nside, theta, phi) if ngal else np.zeros(0, dtype=np.int64)
    order = np.argsort(pix, kind="stable").astype(np.int64)                # galaxies grouped by pixel, catalogue order kept
    pixptr = np.concatenate([[0], np.cumsum(np.bincount(pix, minlength=npix))]).astype(np.int64)
    nmax_l, lmax_n = _tables(amodes)
    modes = [(n, l) for n in range(1, nmax + 1) for l in range(int(amodes.lmax_n[n - 1]) + 1)]
    anlm = np.full(getnlmsize(amodes), np.nan + 0j, dtype=np.complex128)
    g = amodes.basisfunctions
    for b0 in range(0, len(modes), batch):
        mb = modes[b0:b0 + batch]
        mode_n = np.array([m[0] for m in mb], dtype=np.int64)
        mode_l = np.array([m[1] for m in mb], dtype=np.int64)
        gw = np.empty((ngal, len(mb)), dtype=np.float64, order="F")
        for b, (n, l) in enumerate(mb):
            gw[:, b] = weight * g(n, l, r) if ngal else 0.0
        _lib.check(lib.sfb_cat2amln(_lib.ptr(pixptr), _lib.ptr(order), ngal, _lib.ptr(gw), _lib.ptr(mode_n),
                                    _lib.ptr(mode_l), len(mb), float(nbar), _lib.ptr(wr), nside, nmax, lmax,
                                    _lib.ptr(nmax_l), _lib.ptr(lmax_n), _lib.ptr(anlm)))
    if not np.all(np.isfinite(anlm)):
        raise _lib.SFBError("AssertionError: all(isfinite.(anlm))")       # src/cat2anlm.jl:312
    return anlm


def amln2clnn(anlm1, anlm2, cmodes):
    """amln2clnn (…Decompositions.jl:136-148): pseudo-SFB power spectrum (host; a few dot products per mode)."""
    from .modes import getidx
    am = cmodes.amodes
    out = np.empty(cmodes.lnn.shape[1])
    for i in range(out.size):
        l, n1, n2 = (int(x) for x in cmodes.lnn[:, i])
        i1, i2 = getidx(am, n1, l, 0) - 1, getidx(am, n2, l, 0) - 1
        a1, a2 = anlm1[i1:i1 + l + 1], anlm2[i2:i2 + l + 1]
        out[i] = ((a1[0] * np.conj(a2[0])).real + 2 * np.sum((a1[1:] * np.conj(a2[1:])).real)) / (2 * l + 1)
    return out
