"""Host-side mirror of the reference's mode bookkeeping (the inputs of the hot path).

Same names, argument meaning and ordering conventions as src/modes.jl and
src/SphericalBesselGNLs.jl of hsgg/SphericalFourierBesselDecompositions.jl, so that the
tables handed to the C ABI (`lnn`, `G`) are laid out exactly as the Julia structs are:

  AnlmModes(kmax, rmin, rmax) / AnlmModes(nmax, lmax, rmin, rmax)   src/modes.jl:92-175
  getlmsize / getnlmsize / getnlm / getidx                           src/modes.jl:178-232
  ClnnModes(amodes; Δkmax, Δnmax), getlnn, getlnnsize, getidx, getlkk src/modes.jl:280-478,564-589
  ClnnBinnedModes(w̃, v, cmodes), bandpower_binning_weights           src/modes.jl:617-635,727-768
  radial basis g_nl(r) (boundary=potential, cache=false)             src/SphericalBesselGNLs.jl:219-228,309-336,396-421,651-656

Indices stored *inside* tables are 1-based like Julia's (n, and the values returned by
getidx); Python-level positional arguments documented per function.
"""
from __future__ import annotations

import math

import numpy as np
from scipy import sparse, special

__all__ = [
    "AnlmModes", "ClnnModes", "ClnnBinnedModes", "bandpower_binning_weights", "estimate_nside",
    "getlmsize", "getnlmsize", "getnlm", "getidx", "getlnn", "getlnnsize", "getlkk", "SphericalBesselGNL",
]


# --------------------------------------------------------------------------- radial basis

def _zero_fn_potential(k, l, rmin, rmax):
    """knl_zero_function_potential (src/SphericalBesselGNLs.jl:219-228)."""
    jlmax = special.jv(l - 0.5, k * rmax)
    if k * rmin == 0:
        return jlmax
    scale = k ** 3 / (1 + abs(k) ** 3)
    return jlmax * special.yv(l + 1.5, k * rmin) * scale - special.jv(l + 1.5, k * rmin) * special.yv(l - 0.5, k * rmax) * scale


def _refine(fn, a, b, xrtol=1e-10, maxevals=1000):
    """doublesecant_method (src/SphericalBesselGNLs.jl:72-128)."""
    sgn = lambda v: int(v > 0) - int(v < 0)
    fa, fb = fn(a), fn(b)
    if sgn(fa) * sgn(fb) > 0:
        raise ArithmeticError("zero not bracketed")
    for _ in range(maxevals):
        if a == b:
            break
        slope = (fb - fa) / (b - a)
        c = b - fb / slope
        if abs(c - b) <= xrtol * c or abs(c - a) <= xrtol * c:
            return c
        fc = fn(c)
        d = b - (fb + fc) / slope
        if sgn(fa) * sgn(fc) <= 0:
            b, fb = c, fc
        elif sgn(fb) * sgn(fc) <= 0:
            a, fa = c, fc
        else:
            a = b = c
            fa = fb = fc
        if a <= d <= b:
            fd = fn(d)
            if sgn(fa) * sgn(fd) <= 0:
                b, fb = d, fd
            elif sgn(fb) * sgn(fd) <= 0:
                a, fa = d, fd
            else:
                a = b = d
                fa = fb = fd
    raise ArithmeticError("maxevals reached")


def _next_zero(fn, x, step):
    """calc_next_fn_zero (src/SphericalBesselGNLs.jl:131-166)."""
    f = float(fn(x))
    if not math.isfinite(f):
        raise FloatingPointError("Bessel overflow while locating k_nl (the reference switches to ArbFloat here)")
    xn = x + step
    fnew = float(fn(xn))
    while (fnew > 0) - (fnew < 0) == (f > 0) - (f < 0):
        x, f = xn, fnew
        xn += step
        fnew = float(fn(xn))
    return _refine(fn, x, xn)


def _knl_table(rmin, rmax, *, kmax=None, nmax=None, lmax=None):
    """calc_knl_zeros, both methods (src/SphericalBesselGNLs.jl:294-336)."""
    step = math.pi / rmax / 4
    if kmax is None:
        knl = np.full((nmax, lmax + 1), np.nan)
        for l in range(lmax + 1):
            x = (l + 1.5) / rmax
            for n in range(nmax):
                knl[n, l] = _next_zero(lambda k: _zero_fn_potential(k, l, rmin, rmax), x, step)
                x = knl[n, l] + step
        return knl
    nmax = math.ceil(kmax * rmax / math.pi) + 1
    lmax = math.ceil(kmax * rmax)
    knl = np.full((nmax, lmax + 1), np.nan)
    for l in range(lmax + 1):
        x = (l + 1.5) / rmax
        n = 0
        while x <= kmax:
            z = _next_zero(lambda k: _zero_fn_potential(k, l, rmin, rmax), x, step)
            if x <= z <= kmax:
                if n >= nmax:
                    raise ArithmeticError("nmax too small")
                knl[n, l] = z
                n += 1
            x = z + step
    return knl


class SphericalBesselGNL:
    """g_nl(r) = c_nl j_l(k_nl r) + d_nl y_l(k_nl r), potential boundary, exact evaluation (cache=false)."""

    def __init__(self, knl, rmin, rmax):
        self.knl = knl
        self.rmin, self.rmax = float(rmin), float(rmax)
        self.nmax, self.lmax = knl.shape[0], knl.shape[1] - 1
        self.cnl = np.full(knl.shape, np.nan)
        self.dnl = np.full(knl.shape, np.nan)
        for n in range(1, self.nmax + 1):
            for l in range(self.lmax + 1):
                k = knl[n - 1, l]
                if math.isfinite(k):
                    self.cnl[n - 1, l], self.dnl[n - 1, l] = self._cnl_dnl(k, n, l)

    def _cnl_dnl(self, k, n, l):
        """calc_cnl_dnl_potential (src/SphericalBesselGNLs.jl:404-421)."""
        rmin, rmax = self.rmin, self.rmax
        y = special.yv(l + 1.5, k * rmin) if k * rmin != 0 else -math.inf
        dc = -special.jv(l + 1.5, k * rmin) / y

        def g(q):
            v = special.spherical_jn(l, q)
            return v if dc == 0 else v + dc * special.spherical_yn(l, q)

        norm = (rmax ** 3 * g(k * rmax) ** 2 - rmin ** 3 * g(k * rmin) ** 2) / 2
        if not norm >= 0:
            raise ArithmeticError("g_nl normalisation is negative")
        sign = -1.0 if (n + (l > 0) * (n > 1)) % 2 else 1.0
        c = sign / math.sqrt(norm)
        return c, dc * c

    def __call__(self, n, l, r):
        q = self.knl[n - 1, l] * np.asarray(r, dtype=float)
        c, d = self.cnl[n - 1, l], self.dnl[n - 1, l]
        out = c * special.spherical_jn(l, q)
        if d != 0:
            out = out + d * special.spherical_yn(l, q)
        return out


# --------------------------------------------------------------------------- AnlmModes

def estimate_nside(lmax):
    """src/modes.jl:68"""
    return 2 ** max(2, math.ceil(math.log2((2 * lmax + 1) / 2)))


def getlmsize(lmax):
    """src/modes.jl:178-180"""
    return lmax * (lmax + 1) // 2 + lmax + 1


class AnlmModes:
    """AnlmModes(kmax, rmin, rmax; nside=nothing) or AnlmModes(nmax, lmax, rmin, rmax; nside=nothing)."""

    def __init__(self, *args, nside=None):
        if len(args) == 3:
            kmax, rmin, rmax = (float(a) for a in args)
            knl = _knl_table(rmin, rmax, kmax=kmax)
            ok = knl <= kmax
            nmax = int(np.flatnonzero(ok[:, 0])[-1]) + 1
            lmax = int(np.flatnonzero(ok[0, :])[-1])
            self.lmax_n = np.array([np.flatnonzero(ok[n, :])[-1] for n in range(nmax)], dtype=np.int64)
            self.nmax_l = np.array([np.flatnonzero(ok[:, l])[-1] + 1 for l in range(lmax + 1)], dtype=np.int64)
            self.kmax = float(np.nanmax(knl))
            knl = np.ascontiguousarray(knl[:nmax, :lmax + 1])
        elif len(args) == 4:
            nmax, lmax = int(args[0]), int(args[1])
            rmin, rmax = float(args[2]), float(args[3])
            knl = _knl_table(rmin, rmax, nmax=nmax, lmax=lmax)
            self.lmax_n = np.full(nmax, lmax, dtype=np.int64)
            self.nmax_l = np.full(lmax + 1, nmax, dtype=np.int64)
            self.kmax = float(np.nanmax(knl))
        else:
            raise TypeError("AnlmModes(kmax, rmin, rmax) or AnlmModes(nmax, lmax, rmin, rmax)")
        if not rmin < rmax:
            raise ValueError("rmin >= rmax")
        self.rmin, self.rmax = rmin, rmax
        self.nmax, self.lmax = nmax, lmax
        self.knl = knl
        self.basisfunctions = SphericalBesselGNL(knl, rmin, rmax)
        self.nside = int(nside) if nside is not None else estimate_nside(lmax)

    def __repr__(self):
        return (f"AnlmModes(kmax={self.kmax}, rmin={self.rmin}, rmax={self.rmax}, nmax={self.nmax}, lmax={self.lmax}, "
                f"nside={self.nside}, num_modes={int(np.isfinite(self.knl).sum())})")


def getnlmsize(modes, nmax=None):
    """src/modes.jl:183-189"""
    nmax = modes.nmax if nmax is None else nmax
    return int(sum(getlmsize(int(modes.lmax_n[n])) for n in range(nmax)))


def getnlm(modes, idx):
    """1-based idx -> (n, l, m)   (src/modes.jl:203-220)"""
    n = 1
    while idx > getlmsize(int(modes.lmax_n[n - 1])):
        idx -= getlmsize(int(modes.lmax_n[n - 1]))
        n += 1
    l = 0
    while idx > l + 1:
        idx -= l + 1
        l += 1
    return n, l, idx - 1


# --------------------------------------------------------------------------- ClnnModes

class ClnnModes:
    """ClnnModes(amodes; Δkmax=Inf, Δnmax=typemax(Int)) — auto-correlation (S=true) table
    sorted by (l, Δn, n1)  (src/modes.jl:338-393).  `lnn` is a 3 x lnnsize int64 array in
    Fortran order, i.e. the memory image of the Julia Matrix{Int}."""

    def __init__(self, amodes, dkmax=math.inf, dnmax=None):
        self.amodes = self.amodesA = self.amodesB = amodes
        dnmax_in = np.iinfo(np.int64).max if dnmax is None or dnmax == math.inf else int(dnmax)
        cols = []
        for l in range(amodes.lmax + 1):
            nl = int(amodes.nmax_l[l])
            k = amodes.knl[:nl, l]
            n1, n2 = np.triu_indices(nl)
            keep = (np.abs(k[n2] - k[n1]) <= dkmax) & (n2 - n1 <= dnmax_in)
            n1, n2 = n1[keep], n2[keep]
            order = np.lexsort((n1, n2 - n1))
            cols.append(np.stack([np.full(order.size, l), n1[order] + 1, n2[order] + 1]))
        lnn = np.concatenate(cols, axis=1).astype(np.int64)
        self.lnn = np.asfortranarray(lnn)
        k1 = amodes.knl[lnn[1] - 1, lnn[0]]
        k2 = amodes.knl[lnn[2] - 1, lnn[0]]
        self.dkmax = float(np.max(np.abs(k2 - k1))) if lnn.shape[1] else -math.inf
        self.dnmax = int(np.max(lnn[2] - lnn[1])) if lnn.shape[1] else 0
        first = np.zeros(int(lnn[0].max()) + 1, dtype=np.int64)
        ls, pos = np.unique(lnn[0], return_index=True)
        first[ls] = pos + 1
        self.first_ell_idx = first

    def __repr__(self):
        return f"ClnnModes{{S=true}}(Δkmax={self.dkmax}, Δnmax={self.dnmax}, num_lnn={self.lnn.shape[1]})"


def getlnnsize(modes):
    """src/modes.jl:402,638"""
    return modes.LKK.shape[1] if isinstance(modes, ClnnBinnedModes) else modes.lnn.shape[1]


def getlnn(cmodes, idx=None):
    """getlnn(cmodes, idx) with 1-based idx, or the whole table (src/modes.jl:405-412)."""
    if idx is None:
        return cmodes.lnn
    c = cmodes.lnn[:, idx - 1]
    return int(c[0]), int(c[1]), int(c[2])


def getidx(modes, *args):
    """getidx(amodes, n, l, m) (src/modes.jl:223-232) or getidx(cmodes, l, n1, n2) (src/modes.jl:448-478); 1-based."""
    if isinstance(modes, AnlmModes):
        n, l, m = args
        if not (n >= 1 and l >= 0 and m >= 0):
            raise AssertionError("n >= 1, l >= 0, m >= 0")
        return 1 + getnlmsize(modes, n - 1) + getlmsize(l - 1) + m
    l, n1, n2 = args
    if n1 > n2:
        n1, n2 = n2, n1
    nl = int(modes.amodes.nmax_l[l])
    dn = n2 - n1
    idx = int(modes.first_ell_idx[l]) + dn * nl - dn * (dn - 1) // 2 + n1 - 1
    if not 1 <= idx <= getlnnsize(modes):
        raise IndexError("Cannot find index")
    return idx


def getlkk(modes, i=None):
    """src/modes.jl:564-589,640-645"""
    if isinstance(modes, ClnnBinnedModes):
        return modes.LKK if i is None else tuple(modes.LKK[:, i - 1])
    lnn = modes.lnn
    lkk = np.stack([lnn[0].astype(float), modes.amodes.knl[lnn[1] - 1, lnn[0]], modes.amodes.knl[lnn[2] - 1, lnn[0]]])
    return lkk if i is None else tuple(lkk[:, i - 1])


# --------------------------------------------------------------------------- binning

class ClnnBinnedModes:
    """ClnnBinnedModes(w̃, v, cmodes); w̃ / v may be None for Julia's `I`  (src/modes.jl:617-635)."""

    def __init__(self, wtilde, v, cmodes):
        self.cmodes = cmodes
        lkk = getlkk(cmodes)
        if wtilde is None:
            LKK = lkk.copy()
        else:
            wt = sparse.csr_matrix(wtilde)
            if not np.allclose(np.asarray(wt.sum(axis=1)).ravel(), 1):
                raise AssertionError("w̃ rows must sum to 1")
            LKK = (wt @ lkk.T).T
        LKK[1], LKK[2] = np.minimum(LKK[1], LKK[2]), np.maximum(LKK[1], LKK[2])
        self.LKK = LKK


def bandpower_binning_weights(cmodes, dl=1, dn1=1, dn2=1, select="all"):
    """(w̃, v) as scipy CSC matrices (the reference returns SparseMatrixCSC)  (src/modes.jl:727-768).
    Bins are numbered in order of first appearance, like getidx! at src/modes.jl:714-725.  `select`: "all" or a boolean
    mask over the lnn modes — only selected modes get a column, as in the reference's `w̃[1:LNNsize, select]`.
    Built sparse from the start: a dense LNN x lnnsize array is 1 GB at cfg4."""
    lnn = cmodes.lnn
    n_all = lnn.shape[1]
    sel = np.ones(n_all, dtype=bool) if isinstance(select, str) and select == "all" else np.asarray(select, dtype=bool)
    if sel.shape != (n_all,):
        raise ValueError("select must be \"all\" or a boolean mask over the lnn modes")
    keys = np.stack([lnn[0] // dl + 1, (lnn[1] - 1) // dn1 + 1, (lnn[2] - 1) // dn2 + 1], axis=1)[sel]
    # bin index in order of first appearance
    _, first, inv = np.unique(keys, axis=0, return_index=True, return_inverse=True)
    order = np.argsort(first, kind="stable")
    rank = np.empty(order.size, dtype=np.int64)
    rank[order] = np.arange(order.size)
    rows = rank[np.asarray(inv).ravel()]
    nb, n = order.size, keys.shape[0]
    counts = np.bincount(rows, minlength=nb).astype(np.float64)
    cols = np.arange(n)
    wt = sparse.csc_matrix((1.0 / counts[rows], (rows, cols)), shape=(nb, n))
    # v = pinv(w̃).  Every mode falls in exactly one bin, so w̃ w̃ᵀ is diagonal and the pseudo-inverse is the
    # bin indicator: v[i, I] = 1 for i in bin I (what LinearAlgebra.pinv returns up to rounding noise).
    v = sparse.csc_matrix((np.ones(n), (cols, rows)), shape=(n, nb))
    return wt, v
