"""Oracle restatement of the mode / index tables and the radial basis.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Literal, loop-style Python that
follows the reference line by line; 1-based quantities of the Julia source are
kept 1-based in the *values* stored in tables (n, idx), 0-based only in Python
array positions.

Follows (reference paths):
  src/SphericalBesselGNLs.jl:72-128   doublesecant_method
  src/SphericalBesselGNLs.jl:131-166  calc_next_fn_zero
  src/SphericalBesselGNLs.jl:197-214  calc_zeros
  src/SphericalBesselGNLs.jl:219-228  knl_zero_function_potential
  src/SphericalBesselGNLs.jl:294-336  calc_knl_zeros (both methods)
  src/SphericalBesselGNLs.jl:396-421  calc_sphbes_gnl, calc_cnl_dnl_potential
  src/SphericalBesselGNLs.jl:651-656  evaluate (cache=false)
  src/modes.jl:68,122-232             estimate_nside, AnlmModes, getlmsize, getnlmsize, getnlm, getidx
  src/modes.jl:338-478                sort_lnn, calc_lnn, ClnnModes, getlnn, getidx
  src/modes.jl:564-635,727-768        getlkk, ClnnBinnedModes, bandpower_binning_weights
"""
import math

import numpy as np
from scipy import special


# ----------------------------------------------------------------------------
# zero finding  (src/SphericalBesselGNLs.jl:72-214)

def _sign(x):
    return int(x > 0) - int(x < 0)


def doublesecant_method(fn, a, b, maxevals=1000, xrtol=np.finfo(float).eps):
    fa = fn(a)
    fb = fn(b)
    assert _sign(fa) * _sign(fb) <= 0
    neval = 0
    while a != b and neval < maxevals:
        m = (fb - fa) / (b - a)
        c = b - fb / m
        if abs(c - b) <= xrtol * c or abs(c - a) <= xrtol * c:
            return c
        fc = fn(c)
        d = b - (fb + fc) / m
        if _sign(fa) * _sign(fc) <= 0:
            b, fb = c, fc
        elif _sign(fb) * _sign(fc) <= 0:
            a, fa = c, fc
        else:
            a, fa = c, fc
            b, fb = c, fc
        if a <= d <= b:
            fd = fn(d)
            if _sign(fa) * _sign(fd) <= 0:
                b, fb = d, fd
            elif _sign(fb) * _sign(fd) <= 0:
                a, fa = d, fd
            else:
                a, fa = d, fd
                b, fb = d, fd
        neval += 1
    raise RuntimeError("maxevals reached")


def calc_next_fn_zero(func, x, delta, maxevals=1000, xrtol=1e-10):
    f = func(x)
    xnew = x + delta
    fnew = func(xnew)
    while _sign(fnew) == _sign(f):
        x = xnew
        f = fnew
        xnew += delta
        fnew = func(xnew)
    xnew = doublesecant_method(func, x, xnew, maxevals=maxevals, xrtol=xrtol)
    assert math.isfinite(xnew)
    return xnew


def calc_first_n_zeros(func, nmax, delta, xmin):
    xn = np.full(nmax, np.nan)
    for n in range(nmax):
        xn[n] = calc_next_fn_zero(func, xmin, delta)
        xmin = xn[n] + delta
    return xn


def calc_zeros(func, xmin, xmax, delta):
    xn = []
    while xmin <= xmax:
        x = calc_next_fn_zero(func, xmin, delta)
        if xmin <= x <= xmax:
            xn.append(x)
        xmin = x + delta
    return xn


# ----------------------------------------------------------------------------
# k_nl for potential boundary conditions  (src/SphericalBesselGNLs.jl:219-228)

def knl_zero_function_potential(k, l, rmin, rmax):
    scale = k ** 3 / (1 + abs(k) ** 3)
    jlmax = special.jv(l - 0.5, k * rmax)
    if k * rmin == 0:
        return jlmax
    ylmax = special.yv(l - 0.5, k * rmax) * scale
    jlmin = special.jv(l + 1.5, k * rmin)
    ylmin = special.yv(l + 1.5, k * rmin) * scale
    val = jlmax * ylmin - jlmin * ylmax
    if not math.isfinite(val):
        # the reference falls back to ArbFloat{1024} here (AmosException,
        # src/SphericalBesselGNLs.jl:178-184); not needed for the potential
        # boundary at the sizes this oracle is used for.
        raise FloatingPointError("Bessel overflow in knl_zero_function_potential")
    return val


def calc_knl_zeros_nl(nmax, lmax, rmin, rmax):
    """src/SphericalBesselGNLs.jl:294-306"""
    knl = np.full((nmax, lmax + 1), np.nan)
    for l in range(lmax + 1):
        delta = math.pi / rmax / 4
        xmin = (l + 1.5) / rmax
        knl[:, l] = calc_first_n_zeros(
            lambda k: knl_zero_function_potential(k, l, rmin, rmax), nmax, delta, xmin)
    assert np.all(knl > 0)
    return knl


def calc_knl_zeros_kmax(kmax, rmin, rmax, nmax=None, lmax=None):
    """src/SphericalBesselGNLs.jl:309-336"""
    big = 2 ** 62
    nmax = big if nmax is None else nmax
    lmax = big if lmax is None else lmax
    nmax_calc = math.ceil(kmax * rmax / math.pi) + 1
    lmax_calc = math.ceil(kmax * rmax)
    kmax_lim = (lmax > lmax_calc and nmax > nmax_calc)
    nmax = min(nmax, nmax_calc)
    lmax = min(lmax, lmax_calc)
    knl = np.full((nmax, lmax + 1), np.nan)
    for l in range(lmax + 1):
        delta = math.pi / rmax / 4
        kmin = (l + 1.5) / rmax
        kn = calc_zeros(lambda k: knl_zero_function_potential(k, l, rmin, rmax), kmin, kmax, delta)
        if nmax < len(kn):
            raise RuntimeError("nmax too small")
        if kmax_lim and l == lmax:
            assert len(kn) == 0
        for i, kk in enumerate(kn):
            knl[i, l] = kk
    return knl


# ----------------------------------------------------------------------------
# g_nl(r)  (src/SphericalBesselGNLs.jl:396-421, 651-656)

def _sphbes_yl(l, x):
    if x == 0:
        return -math.inf
    return special.spherical_yn(l, x)


def _bes_yl(nu, x):
    if x == 0:
        return -math.inf
    return special.yv(nu, x)


def calc_sphbes_gnl(q, l, c, d):
    jl = special.spherical_jn(l, q)
    if np.all(d == 0):
        return c * jl
    yl = special.spherical_yn(l, q)
    return c * jl + d * yl


def calc_cnl_dnl_potential(knl, n, l, rmin, rmax):
    dc = -special.jv(l + 1.5, knl * rmin) / _bes_yl(l + 1.5, knl * rmin)
    gnl_rmin = calc_sphbes_gnl(knl * rmin, l, 1.0, dc)
    gnl_rmax = calc_sphbes_gnl(knl * rmax, l, 1.0, dc)
    num_one = (rmax ** 3 * gnl_rmax ** 2 - rmin ** 3 * gnl_rmin ** 2) / 2
    assert num_one >= 0
    expo = n + (1 - (1 // (l + 1))) * (1 - (1 // n))
    cnl = (-1) ** expo / math.sqrt(num_one)
    dnl = dc * cnl
    return cnl, dnl


def calc_cnl_dnl(knl, rmin, rmax):
    """src/SphericalBesselGNLs.jl:530-555 (potential boundary)"""
    cnl = np.full(knl.shape, np.nan)
    dnl = np.full(knl.shape, np.nan)
    for n in range(1, knl.shape[0] + 1):
        for l in range(knl.shape[1]):
            if not math.isfinite(knl[n - 1, l]):
                continue
            cnl[n - 1, l], dnl[n - 1, l] = calc_cnl_dnl_potential(knl[n - 1, l], n, l, rmin, rmax)
    return cnl, dnl


class GNL:
    """SphericalBesselGNL with cache=false, boundary=potential."""

    def __init__(self, knl, rmin, rmax):
        self.knl = knl
        self.rmin = float(rmin)
        self.rmax = float(rmax)
        self.nmax = knl.shape[0]
        self.lmax = knl.shape[1] - 1
        self.cnl, self.dnl = calc_cnl_dnl(knl, self.rmin, self.rmax)

    def __call__(self, n, l, r):
        k = self.knl[n - 1, l]
        return calc_sphbes_gnl(k * np.asarray(r, dtype=float), l, self.cnl[n - 1, l], self.dnl[n - 1, l])


# ----------------------------------------------------------------------------
# AnlmModes  (src/modes.jl:68-232)

def estimate_nside(lmax):
    return 2 ** max(2, math.ceil(math.log2((2 * lmax + 1) / 2)))


def getlmsize(lmax):
    return lmax * (lmax + 1) // 2 + lmax + 1


class AnlmModes:
    def __init__(self, *args, nside=None):
        if len(args) == 3:
            kmax, rmin, rmax = args
            self._init_kmax(float(kmax), float(rmin), float(rmax), nside)
        else:
            nmax, lmax, rmin, rmax = args
            self._init_nl(int(nmax), int(lmax), float(rmin), float(rmax), nside)

    def _init_kmax(self, kmax, rmin, rmax, nside):
        """src/modes.jl:122-158"""
        knl = calc_knl_zeros_kmax(kmax, rmin, rmax)
        modes = knl <= kmax  # NaN -> False
        nmax = int(np.flatnonzero(modes[:, 0])[-1]) + 1
        lmax = int(np.flatnonzero(modes[0, :])[-1])
        lmax_n = [int(np.flatnonzero(modes[n, :])[-1]) for n in range(nmax)]
        nmax_l = [int(np.flatnonzero(modes[:, l])[-1]) + 1 for l in range(lmax + 1)]
        assert all(x > 0 for x in nmax_l)
        assert all(x >= 0 for x in lmax_n)
        if nside is None:
            nside = estimate_nside(lmax)
        self.rmin, self.rmax = rmin, rmax
        self.kmax = float(np.nanmax(knl))
        knl = np.array(knl[:nmax, :lmax + 1])
        self.basisfunctions = GNL(knl, rmin, rmax)
        self.nmax, self.lmax = nmax, lmax
        self.nmax_l = np.array(nmax_l, dtype=np.int64)
        self.lmax_n = np.array(lmax_n, dtype=np.int64)
        self.nside = int(nside)
        self.knl = knl

    def _init_nl(self, nmax, lmax, rmin, rmax, nside):
        """src/modes.jl:161-175"""
        knl = calc_knl_zeros_nl(nmax, lmax, rmin, rmax)
        if nside is None:
            nside = estimate_nside(lmax)
        self.rmin, self.rmax = rmin, rmax
        self.kmax = float(np.nanmax(knl))
        self.basisfunctions = GNL(knl, rmin, rmax)
        self.nmax, self.lmax = nmax, lmax
        self.nmax_l = np.full(lmax + 1, nmax, dtype=np.int64)
        self.lmax_n = np.full(nmax, lmax, dtype=np.int64)
        self.nside = int(nside)
        self.knl = knl


def getnlmsize(modes, nmax=None):
    nmax = modes.nmax if nmax is None else nmax
    s = 0
    for n in range(1, nmax + 1):
        s += getlmsize(int(modes.lmax_n[n - 1]))
    return s


def getnlm(modes, idx):
    n = 1
    nmodes = getlmsize(int(modes.lmax_n[n - 1]))
    while idx > nmodes:
        idx -= nmodes
        n += 1
        nmodes = getlmsize(int(modes.lmax_n[n - 1]))
    l = 0
    nmodes = l + 1
    while idx > nmodes:
        idx -= nmodes
        l += 1
        nmodes = l + 1
    m = idx - 1
    return n, l, m


def getidx_nlm(modes, n, l, m):
    assert n >= 1 and l >= 0 and m >= 0
    idx = 1
    idx += getnlmsize(modes, n - 1)
    idx += getlmsize(l - 1)
    idx += m
    return idx


# ----------------------------------------------------------------------------
# ClnnModes  (src/modes.jl:280-478)

class ClnnModes:
    def __init__(self, amodes, dkmax=math.inf, dnmax=2 ** 62):
        """ClnnModes(amodes; Δkmax, Δnmax) with symmetric_kk=true  (src/modes.jl:358-393)"""
        self.amodes = amodes
        lmax = amodes.lmax
        lnn = []
        dkmax_out = -math.inf
        dnmax_out = 0
        for l in range(lmax + 1):
            nAmax = int(amodes.nmax_l[l])
            for nA in range(1, nAmax + 1):
                for nB in range(nA, nAmax + 1):
                    kA = amodes.knl[nA - 1, l]
                    kB = amodes.knl[nB - 1, l]
                    dk = kB - kA
                    if abs(dk) <= dkmax and abs(nB - nA) <= dnmax:
                        lnn.append((l, nA, nB))
                        dkmax_out = max(abs(dk), dkmax_out)
                        dnmax_out = max(abs(nB - nA), dnmax_out)
        # sort_lnn: by l, then Δn, then n1  (src/modes.jl:338-355)
        lnn.sort(key=lambda t: (t[0], t[2] - t[1], t[1]))
        self.lnn = np.array(lnn, dtype=np.int64).T.copy()  # 3 x lnnsize, like the Julia Matrix{Int}
        self.dkmax = dkmax_out
        self.dnmax = dnmax_out
        lnnsize = self.lnn.shape[1]
        lm = int(self.lnn[0].max())
        self.first_ell_idx = np.zeros(lm + 1, dtype=np.int64)  # 1-based values, 0 = unset
        for i in range(1, lnnsize + 1):
            l = int(self.lnn[0, i - 1])
            if self.first_ell_idx[l] == 0:
                self.first_ell_idx[l] = i


def getlnnsize(cmodes):
    if isinstance(cmodes, ClnnBinnedModes):
        return cmodes.LKK.shape[1]
    return cmodes.lnn.shape[1]


def getlnn(cmodes, idx):
    """1-based idx  (src/modes.jl:405-410)"""
    return int(cmodes.lnn[0, idx - 1]), int(cmodes.lnn[1, idx - 1]), int(cmodes.lnn[2, idx - 1])


def getidx_lnn(cmodes, l, n1, n2):
    """Closed form of src/modes.jl:448-478 (1-based result)."""
    if n1 > n2:
        n1, n2 = n2, n1
    lnnsize = getlnnsize(cmodes)
    idx = int(cmodes.first_ell_idx[l])
    for dn in range(1, n2 - n1 + 1):
        idx += int(cmodes.amodes.nmax_l[l]) - dn + 1
    idx += n1 - 1
    if not (1 <= idx <= lnnsize):
        raise IndexError("Cannot find index")
    return idx


def getlkk(cmodes):
    """src/modes.jl:564-582"""
    lnnsize = getlnnsize(cmodes)
    lkk = np.zeros((3, lnnsize))
    for i in range(1, lnnsize + 1):
        l, n1, n2 = getlnn(cmodes, i)
        lkk[0, i - 1] = l
        lkk[1, i - 1] = cmodes.amodes.knl[n1 - 1, l]
        lkk[2, i - 1] = cmodes.amodes.knl[n2 - 1, l]
    return lkk


# ----------------------------------------------------------------------------
# ClnnBinnedModes, bandpower_binning_weights  (src/modes.jl:617-635, 714-768)

class ClnnBinnedModes:
    def __init__(self, wtilde, v, cmodes):
        """wtilde / v may be None, standing for Julia's UniformScaling `I`."""
        self.cmodes = cmodes
        lkk = getlkk(cmodes)
        if wtilde is None:
            LKK = lkk.copy()
        else:
            wt = np.asarray(wtilde)
            assert np.allclose(wt.sum(axis=1), 1)
            LKK = lkk @ wt.T
        for i in range(LKK.shape[1]):  # S=true: ensure k1 <= k2
            lo, hi = min(LKK[1, i], LKK[2, i]), max(LKK[1, i], LKK[2, i])
            LKK[1, i], LKK[2, i] = lo, hi
        self.LKK = LKK


def bandpower_binning_weights(cmodes, dl=1, dn1=1, dn2=1, select="all"):
    """Dense (wtilde, v) — the reference returns the same values as SparseMatrixCSC (src/modes.jl:727-768).
    `select`: "all" or a boolean mask over the lnn modes; only selected modes get a column (w̃[1:LNNsize, select])."""
    lnnsize = getlnnsize(cmodes)
    sel = np.ones(lnnsize, dtype=bool) if isinstance(select, str) and select == "all" else np.asarray(select, dtype=bool)
    assert sel.shape == (lnnsize,)
    iLNN = []
    rows = []
    for i in range(1, lnnsize + 1):
        if not sel[i - 1]:
            continue
        l, n1, n2 = getlnn(cmodes, i)
        key = (l // dl + 1, (n1 - 1) // dn1 + 1, (n2 - 1) // dn2 + 1)
        if key not in iLNN:
            iLNN.append(key)
        rows.append(iLNN.index(key))
    LNNsize = len(iLNN)
    wt = np.zeros((LNNsize, lnnsize))
    for i, I in zip(np.flatnonzero(sel), rows):
        wt[I, i] += 1
    wt = wt[:, sel]
    wt = wt / wt.sum(axis=1, keepdims=True)
    v = np.linalg.pinv(wt)
    assert np.allclose(wt.sum(axis=1), 1)
    return wt, v
