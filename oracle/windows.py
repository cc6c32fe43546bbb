"""Oracle restatement of the window -> coupling-matrix path (src/windows.jl).

TEST INFRASTRUCTURE (see oracle/__init__.py).

Follows (reference paths):
  src/windows.jl:80-125     ConfigurationSpaceModes, window_r
  src/windows.jl:421-431    wigner3j000
  src/windows.jl:528-545    calc_Wr_lm (dense and separable)
  src/windows.jl:548-559    precompute_gnlr
  src/windows.jl:589-610    optimize_Wr_lm_layout
  src/windows.jl:613-627    calc_cmixlnnLNN!
  src/windows.jl:631-647    calc_cmixii!
  src/windows.jl:651-679    calc_cmixii_separable
  src/windows.jl:682-697    calc_Wrl_Wrl
  src/windows.jl:700-746    calc_cmix
  src/windows.jl:781-805    power_win_mix (unbinned)
  src/windows.jl:809-814    power_win_mix (separable, unbinned)
  src/windows.jl:825-862    _power_win_mix (dense, binned)
  src/windows.jl:866-878    calc_angular_mixing_matrix
  src/windows.jl:924-938    calc_radial_mixing
  src/windows.jl:942-990    _power_win_mix (separable, binned)
  src/windows.jl:994-1015   power_win_mix (binned)
  src/SphericalFourierBesselDecompositions.jl:221-234,300-494   gen_mask, make_window (subset of features)
"""
import math

import numpy as np
from scipy.special import gammaln

from . import healpix as hp
from . import modes as om


# ----------------------------------------------------------------------------
class ConfigurationSpaceModes:
    def __init__(self, rmin, rmax, nr, nside):
        self.rmin, self.rmax = float(rmin), float(rmax)
        self.dr = (self.rmax - self.rmin) / nr
        # range(rmin+Δr/2, rmax-Δr/2, length=nr)
        self.r = np.linspace(self.rmin + self.dr / 2, self.rmax - self.dr / 2, nr)
        self.nr = nr
        self.npix = hp.nside2npix(nside)
        self.nside = nside


def window_r(wmodes):
    return wmodes.r, wmodes.dr


class SeparableArray:
    """win[i, p] = phi[i] * mask[p]  (src/SeparableArrays.jl:53-122)."""

    def __init__(self, phi, mask):
        self.phi = np.asarray(phi, dtype=float)
        self.mask = np.asarray(mask, dtype=float)

    def dense(self):
        return np.outer(self.phi, self.mask)


# ----------------------------------------------------------------------------
# synthetic windows

def gen_mask(nside, fsky):
    """Polar cap θ <= acos(1-2 fsky)  (…Decompositions.jl:221-234)."""
    npix = hp.nside2npix(nside)
    theta, _ = hp.pix2ang_ring(nside, np.arange(npix))
    thmax = math.acos(1 - 2 * fsky)
    return (theta <= thmax).astype(float)


def make_window(wmodes, *features):
    """Subset of make_window (…Decompositions.jl:300-494): :fullsky, :ang_75/half/quarter/
    eighth/sixteenth, :radial, :radial_expmrr0, :separable, :dense, :rotate (:277-299,
    378-382; oracle/rotate.py); renormalised to max 1 after every feature like the
    reference (:467-475)."""
    r = wmodes.r
    win = np.ones((wmodes.nr, wmodes.npix))
    fsky = {"ang_75": 0.75, "ang_half": 0.5, "ang_quarter": 0.25, "ang_eighth": 0.125,
            "ang_sixteenth": 1 / 16}

    def normalise(w):
        if isinstance(w, SeparableArray):
            w.mask = w.mask / w.mask.max()
            # maximum(win.phi * win.mask') = the largest of the four extreme products
            w.phi = w.phi / max(a * b for a in (w.phi.max(), w.phi.min()) for b in (w.mask.max(), w.mask.min()))
        else:
            w /= w.max()
        return w

    for feat in features:
        if feat == "fullsky":
            pass
        elif feat in fsky:
            mask = gen_mask(wmodes.nside, fsky[feat])
            if isinstance(win, SeparableArray):
                win.mask = win.mask * mask
            else:
                win = win * mask[None, :]
        elif feat == "radial":
            phi = np.exp(-(r / (wmodes.rmax * 0.55)) ** 2)
            if isinstance(win, SeparableArray):
                win.phi = win.phi * phi
            else:
                win = win * phi[:, None]
        elif feat == "radial_expmrr0":
            phi = np.exp(-r / ((wmodes.rmin + wmodes.rmax) / 2 / 3))
            if isinstance(win, SeparableArray):
                win.phi = win.phi * phi
            else:
                win = win * phi[:, None]
        elif feat == "separable":
            win = SeparableArray(win.mean(axis=1), win.mean(axis=0))
        elif feat == "rotate":
            from . import rotate as rot
            ang = (rot.ROTATE_ALPHA, rot.ROTATE_BETA, rot.ROTATE_GAMMA)
            if isinstance(win, SeparableArray):
                win.mask = rot.rotate_euler(win.mask, *ang)
            else:
                win = np.stack([rot.rotate_euler(win[i], *ang) for i in range(win.shape[0])])
        elif feat == "dense":
            win = win.dense() if isinstance(win, SeparableArray) else np.array(win)
        else:
            raise ValueError(f"Unsupported feature {feat}.")
        win = normalise(win)
    return win


# ----------------------------------------------------------------------------
# stage 1

def calc_Wr_lm(win, LMAX, Wnside, niter=3):
    """src/windows.jl:528-545.  Dense: [nr, lmsize] complex, HEALPix m-major columns.
    Separable: (phi, wlm)."""
    if isinstance(win, SeparableArray):
        mask = hp.udgrade(win.mask, Wnside)
        wlm = hp.SHT(Wnside, LMAX).map2alm(mask[None, :], niter=niter)[0]
        return (win.phi, wlm)
    win = np.asarray(win, dtype=float)
    sht = hp.SHT(Wnside, LMAX)
    W = hp.udgrade(win, Wnside)
    return sht.map2alm(W, niter=niter)


def optimize_Wr_lm_layout(Wr_lm, LMAX):
    """m-major -> m-fast column permutation  (src/windows.jl:589-605)."""
    out = np.empty_like(Wr_lm)
    for l in range(LMAX + 1):
        for m in range(l + 1):
            out[:, hp.lm_index_mfast(l, m)] = Wr_lm[:, hp.lm_index_mmajor(LMAX, l, m)]
    return out


# ----------------------------------------------------------------------------
# stage 2 (reference order)

def calc_Wrl_Wrl(W1r_lm, W2r_lm, LMAX):
    """W[i,j,L1] = Σ_{M1>=0} (2-δ_{M1,0}) Re(W1[i,L1M1] conj(W2[j,L1M1])), m-fast inputs
    (src/windows.jl:682-696).  Returned as [LMAX+1, nr, nr]."""
    nr = W1r_lm.shape[0]
    out = np.empty((LMAX + 1, nr, nr))
    for L1 in range(LMAX + 1):
        b = hp.lm_index_mfast(L1, 0)
        A = W1r_lm[:, b:b + L1 + 1]
        B = W2r_lm[:, b:b + L1 + 1]
        s = np.real(np.outer(A[:, 0], np.conj(B[:, 0])))
        if L1 > 0:
            s = s + 2 * np.real(A[:, 1:] @ np.conj(B[:, 1:]).T)
        out[L1] = s
    return out


def precompute_gnlr(amodes, wmodes):
    """gnlr[:, n-1, l] = g_nl(r), NaN where n > nmax_l[l]  (src/windows.jl:548-559)."""
    r = wmodes.r
    gnlr = np.full((len(r), amodes.nmax, amodes.lmax + 1), np.nan)
    for l in range(amodes.lmax + 1):
        for n in range(1, int(amodes.nmax_l[l]) + 1):
            gnlr[:, n - 1, l] = amodes.basisfunctions(n, l, r)
    return gnlr


def rsdrgnlr(amodes, wmodes):
    """r .* √Δr .* precompute_gnlr  (src/windows.jl:799)."""
    r, dr = window_r(wmodes)
    return r[:, None, None] * math.sqrt(dr) * precompute_gnlr(amodes, wmodes)


# ----------------------------------------------------------------------------
# stage 3 (reference order)

def wigner3j000(l, lp, L):
    """src/windows.jl:421-431"""
    if not (abs(l - lp) <= L <= l + lp):
        return 0.0
    J = l + lp + L
    if J % 2 != 0:
        return 0.0
    w = (-1) ** (J // 2) * math.exp(0.5 * gammaln(1 + J - 2 * l) + 0.5 * gammaln(1 + J - 2 * lp)
                                    + 0.5 * gammaln(1 + J - 2 * L) - 0.5 * gammaln(1 + J + 1)
                                    + gammaln(1 + J // 2)
                                    - gammaln(1 + J // 2 - l) - gammaln(1 + J // 2 - lp)
                                    - gammaln(1 + J // 2 - L))
    return w


def calc_cmixlnnLNN(l, n, n_, L, N, N_, W, G):
    """src/windows.jl:613-627.  W [LMAX+1, nr, nr]; G [nr, nmax, lmax+1]; n,N 1-based."""
    gg1 = G[:, n - 1, l] * G[:, N - 1, L]
    gg2 = G[:, n_ - 1, l] * G[:, N_ - 1, L]
    mix = 0.0
    for L1 in range(abs(l - L), l + L + 1, 2):
        w3j = wigner3j000(l, L, L1)
        mix += w3j ** 2 * (gg1 @ W[L1] @ gg2)
    mix *= (2 * L + 1) / (4 * math.pi)
    return mix


def calc_cmixii(i, i_, cmodes, G, W, div2Lp1, interchange):
    """src/windows.jl:631-647 (1-based i, i')."""
    l, n, n_ = om.getlnn(cmodes, i)
    L, N, N_ = om.getlnn(cmodes, i_)
    if interchange:
        N, N_ = N_, N
    mix = calc_cmixlnnLNN(l, n, n_, L, N, N_, W, G)
    if (not interchange) and N != N_:
        mix += calc_cmixlnnLNN(l, n, n_, L, N_, N, W, G)
    if div2Lp1:
        mix /= (2 * L + 1)
    return mix


def calc_cmix(cmodes, G, W, div2Lp1=False, interchange=False, lnn_min=1):
    """Literal src/windows.jl:700-746.  O(lnnsize² · L · nr²): tiny cases only."""
    lnnsize = om.getlnnsize(cmodes)
    n = lnnsize - lnn_min + 1
    mix = np.empty((n, n))
    for i_ in range(lnn_min, lnnsize + 1):
        for i in range(lnn_min, lnnsize + 1):
            mix[i - lnn_min, i_ - lnn_min] = calc_cmixii(i, i_, cmodes, G, W, div2Lp1, interchange)
    return mix


def calc_cmix_blocked(cmodes, G, W, div2Lp1=False, interchange=False, lnn_min=1):
    """Same numbers as calc_cmix, evaluated per (ℓ,L) block with BLAS (independent
    operation order: per-L1 A·W_{L1}·Aᵀ, no pre-combination over L1)."""
    lnn = cmodes.lnn
    lnnsize = lnn.shape[1]
    ells = np.unique(lnn[0])
    rows_of = {int(l): np.flatnonzero(lnn[0] == l) for l in ells}
    mix = np.zeros((lnnsize, lnnsize))
    for l in ells:
        l = int(l)
        ri = rows_of[l]
        a = int(max(lnn[1, ri].max(), lnn[2, ri].max()))
        for L in ells:
            L = int(L)
            ci = rows_of[L]
            b = int(max(lnn[1, ci].max(), lnn[2, ci].max()))
            A = (G[:, :a, l][:, :, None] * G[:, :b, L][:, None, :]).reshape(G.shape[0], a * b).T  # [(n,N), r]
            T = np.zeros((a * b, a * b))
            for L1 in range(abs(l - L), l + L + 1, 2):
                T += wigner3j000(l, L, L1) ** 2 * (A @ W[L1] @ A.T)
            T = T.reshape(a, b, a, b)  # [n, N, n', N']
            n, n_ = lnn[1, ri] - 1, lnn[2, ri] - 1
            N, N_ = lnn[1, ci] - 1, lnn[2, ci] - 1
            if interchange:
                blk = T[n[:, None], N_[None, :], n_[:, None], N[None, :]]
            else:
                blk = T[n[:, None], N[None, :], n_[:, None], N_[None, :]]
                blk = blk + np.where((N != N_)[None, :], T[n[:, None], N_[None, :], n_[:, None], N[None, :]], 0.0)
            blk = blk * ((1.0 if div2Lp1 else (2 * L + 1)) / (4 * math.pi))
            mix[np.ix_(ri, ci)] = blk
    return mix[lnn_min - 1:, lnn_min - 1:]


def power_win_mix(win1, win2, wmodes, cmodes, div2Lp1=False, interchange=False, lnn_min=1, literal=False):
    """src/windows.jl:781-805 (dense) and :809-814 (separable)."""
    if isinstance(win1, SeparableArray):
        return power_win_mix_binned(win1, win2, None, None, wmodes, om.ClnnBinnedModes(None, None, cmodes),
                                    div2Lp1=div2Lp1, interchange=interchange)
    amodes = cmodes.amodes
    LMAX = 2 * amodes.lmax
    W1 = optimize_Wr_lm_layout(calc_Wr_lm(win1, LMAX, amodes.nside), LMAX)
    W2 = W1 if win2 is win1 else optimize_Wr_lm_layout(calc_Wr_lm(win2, LMAX, amodes.nside), LMAX)
    W = calc_Wrl_Wrl(W1, W2, LMAX)
    G = rsdrgnlr(amodes, wmodes)
    fn = calc_cmix if literal else calc_cmix_blocked
    mix = fn(cmodes, G, W, div2Lp1, interchange, lnn_min=lnn_min)
    assert np.all(np.isfinite(mix))
    return mix


# ----------------------------------------------------------------------------
# separable pieces

def win_lnn(win, wmodes, cmodes):
    """src/windows.jl:382-391 + calc_intr_gg_fn :394-418 (derivative=0).  nodes = r, weights = Δr (:568); the spline
    through (r, Wr_00/√4π) evaluated at its own knots returns the knot values, so fn = Wr_00/√4π.  Literal operation
    order: gnlr *= nodes √weights √fn, then the dot product per (l,n,n')."""
    amodes = cmodes.amodes
    Wr_00 = np.real(calc_Wr_lm(win, 2 * amodes.lmax, amodes.nside)[:, 0])
    r, dr = window_r(wmodes)
    fn = Wr_00 / math.sqrt(4 * math.pi)
    if np.any(fn < 0):
        raise ValueError("DomainError: sqrt of a negative Wr_00")
    gnlr = precompute_gnlr(amodes, wmodes) * (r * math.sqrt(dr) * np.sqrt(fn))[:, None, None]
    lnn = cmodes.lnn
    out = np.empty(lnn.shape[1])
    for i in range(lnn.shape[1]):
        l, n, n_ = lnn[:, i]
        out[i] = gnlr[:, n - 1, l] @ gnlr[:, n_ - 1, l]
    return out


# ----------------------------------------------------------------------------
# calc_wmix and the brute-force coupling matrix from it (src/windows.jl:273-364, 367-378, 434-525): an independent
# route to M through general-m Gaunt sums, which the reference's own test pins against power_win_mix(win, ...) at
# rtol 1e-10 (test/test_windows.jl:403-409).  SURVEY §8f row 1 / §8c identity 3.

def calc_gaunts_L(l, lp, m, mp):
    """src/windows.jl:449-464: gaunt[L] = (l l' L; m m' -m-m') (l l' L; 0 0 0) sqrt((2L+1)(2l+1)(2l'+1)/4π), L = Lmin.."""
    from .wigner import wigner3j_family
    Lmin, w3j = wigner3j_family(l, lp, m, mp)
    Lmin0, w000 = wigner3j_family(l, lp, 0, 0)
    L = Lmin + np.arange(w3j.size)
    return w3j * np.sqrt((2 * L + 1) * (2 * l + 1) * (2 * lp + 1) / (4 * math.pi)) * w000[L - Lmin0], L


def calc_wmix_ii(l, m, lp, mp, gg1, Wr_lm, LMAX):
    """src/windows.jl:273-294 (Wr_lm in HEALPix m-major column order, LMcalcStruct(LMAX))."""
    M = m - mp
    gaunt, L = calc_gaunts_L(l, lp, -m, mp)
    aM = abs(M)
    w_ang = 0.0 + 0.0j
    for j in range(L.size):
        LM = int(L[j]) + (aM * (2 * LMAX + 1 - aM)) // 2     # 0-based LMcalcStruct index, src/LMcalcStructs.jl:13-18
        w_ang += gaunt[j] * (gg1 @ Wr_lm[:, LM])
    if M < 0:
        w_ang = (-1) ** M * np.conj(w_ang)
    return (-1) ** m * w_ang


def calc_wmix(win, wmodes, amodes, neg_m=False, only_nl=None):
    """src/windows.jl:299-364: W_{nlm}^{n'l'm'} for m, m' >= 0 (or m -> -m with neg_m), nlmsize x nlmsize ComplexF64.
    `only_nl` (test aid, not in the reference): restrict both loops to the listed (n,l); other entries stay NaN."""
    nlmsize = om.getnlmsize(amodes)
    LMAX = 2 * amodes.lmax
    Wr_lm = calc_Wr_lm(win, LMAX, amodes.nside)
    G = rsdrgnlr(amodes, wmodes)
    wmix = np.full((nlmsize, nlmsize), np.nan + 0j)
    nl = [(n, l) for n in range(1, amodes.nmax + 1) for l in range(int(amodes.lmax_n[n - 1]) + 1)]
    if only_nl is not None:
        nl = [x for x in nl if x in set(only_nl)]
    for (n, l) in nl:
        ibase = om.getidx_nlm(amodes, n, l, 0) - 1
        for (n_, l_) in nl:
            ibase_ = om.getidx_nlm(amodes, n_, l_, 0) - 1
            gg1 = G[:, n - 1, l] * G[:, n_ - 1, l_]
            for m in range(l + 1):
                for m_ in range(l_ + 1):
                    wmix[ibase + m, ibase_ + m_] = calc_wmix_ii(l, -m if neg_m else m, l_, m_, gg1, Wr_lm, LMAX)
    return wmix


def get_wmix(w, w_, nl, m, NL, M):
    """src/windows.jl:367-378 (nl, NL: 0-based indices of the m = 0 entries)."""
    if m >= 0:
        if M >= 0:
            return w[nl + m, NL + M]
        return (-1) ** (m + M) * np.conj(w_[nl + m, NL - M])
    if M >= 0:
        return w_[nl - m, NL + M]
    return (-1) ** (m - M) * np.conj(w[nl - m, NL - M])


def power_win_mix_ii(lnn, LNN, wmix, wmix_negm, amodes):
    """src/windows.jl:467-509."""
    l, n, n_ = lnn
    L, N, N_ = LNN
    j0 = om.getidx_nlm(amodes, n, l, 0) - 1
    j_0 = om.getidx_nlm(amodes, n_, l, 0) - 1
    J0 = om.getidx_nlm(amodes, N, L, 0) - 1
    J_0 = om.getidx_nlm(amodes, N_, L, 0) - 1
    mix = 0.0
    for m in range(l + 1):
        for M in range(L + 1):
            mix += np.real(wmix[j0 + m, J0 + M] * np.conj(wmix[j_0 + m, J_0 + M]))
            for (sm, sM, cond) in ((-1, 1, m > 0), (1, -1, M > 0), (-1, -1, m > 0 and M > 0)):
                if cond:
                    w1 = get_wmix(wmix, wmix_negm, j0, sm * m, J0, sM * M)
                    w2 = get_wmix(wmix, wmix_negm, j_0, sm * m, J_0, sM * M)
                    mix += np.real(w1 * np.conj(w2))
    return mix / (2 * l + 1)


def power_win_mix_from_wmix(wmix, wmix_negm, cmodes):
    """power_win_mix(wmix, wmix_negm, cmodes), src/windows.jl:512-525."""
    amodes = cmodes.amodes
    n = cmodes.lnn.shape[1]
    out = np.empty((n, n))
    for ip in range(n):
        L, N, N_ = (int(x) for x in cmodes.lnn[:, ip])
        for i in range(n):
            l, a, b = (int(x) for x in cmodes.lnn[:, i])
            v = power_win_mix_ii((l, a, b), (L, N, N_), wmix, wmix_negm, amodes)
            if N != N_:
                v += power_win_mix_ii((l, a, b), (L, N_, N), wmix, wmix_negm, amodes)
            out[i, ip] = v
    return out


def sum_m_lmeqLM(A, cmodes):
    """src/theory.jl:86-118: (1/(2l+1)) Σ_m Re A[(n1,l,m),(n2,l,m)], symmetrised in (n1, n2), per (l,n1,n2) of cmodes."""
    amodes = cmodes.amodes
    out = np.empty(cmodes.lnn.shape[1])
    for i in range(out.size):
        l, n1, n2 = (int(x) for x in cmodes.lnn[:, i])
        acc = 0.0
        for (a, b) in ((n1, n2), (n2, n1)):
            ia = om.getidx_nlm(amodes, a, l, 0) - 1
            ib = om.getidx_nlm(amodes, b, l, 0) - 1
            cl = np.real(A[ia, ib]) + 2 * sum(np.real(A[ia + m, ib + m]) for m in range(1, l + 1))
            acc += cl / (2 * l + 1)
        out[i] = acc / 2
    return out


def calc_angular_mixing_matrix(lmax, w1lm, w2lm):
    """src/windows.jl:866-878 (m-major alm of length lmsize(2 lmax))."""
    Wl = hp.alm2cl(w1lm, w2lm, 2 * lmax)
    ang = np.full((lmax + 1, lmax + 1), np.nan)
    for L in range(lmax + 1):
        for l in range(lmax + 1):
            s = 0.0
            for L1 in range(abs(L - l), L + l + 1, 2):
                s += wigner3j000(l, L, L1) ** 2 * (2 * L1 + 1) * Wl[L1]
            ang[l, L] = s / (4 * math.pi)
    return ang


def calc_radial_mixing(lmax, nmax_l, G, phi):
    """src/windows.jl:924-938 with r=Δr=1 as called at :956-957.  [n, l, N, L]."""
    nmax = int(max(nmax_l))
    out = np.full((nmax, lmax + 1, nmax, lmax + 1), np.nan)
    for L in range(lmax + 1):
        for N in range(1, int(nmax_l[L]) + 1):
            for l in range(lmax + 1):
                for n in range(1, int(nmax_l[l]) + 1):
                    out[n - 1, l, N - 1, L] = np.sum(G[:, n - 1, l] * G[:, N - 1, L] * phi)
    return out


def calc_cmixii_separable(i, i_, cmodes, R1, R2, ang, div2Lp1, interchange):
    """src/windows.jl:651-679"""
    l, n, n_ = om.getlnn(cmodes, i)
    L, N, N_ = om.getlnn(cmodes, i_)
    if interchange:
        N, N_ = N_, N
    mix = ang[l, L] * R1[n - 1, l, N - 1, L] * R2[n_ - 1, l, N_ - 1, L]
    if (not interchange) and N != N_:
        mix += ang[l, L] * R1[n - 1, l, N_ - 1, L] * R2[n_ - 1, l, N - 1, L]
    if not div2Lp1:
        mix *= (2 * L + 1)
    return mix


# ----------------------------------------------------------------------------
# binned

def power_win_mix_binned(win1, win2, wtilde, v, wmodes, bcmodes, div2Lp1=False, interchange=False):
    """src/windows.jl:994-1015 -> :825-862 (dense) / :942-990 (separable).
    wtilde / v dense arrays or None (= Julia `I`).  Reproduces the reference quirk that
    W2r_lm is computed from win1 (:1005-1006)."""
    cmodes = bcmodes.cmodes
    amodes = cmodes.amodes
    lnnsize = om.getlnnsize(cmodes)
    LMAX = 2 * amodes.lmax
    G = rsdrgnlr(amodes, wmodes)
    if isinstance(win1, SeparableArray):
        phi1, w1lm = calc_Wr_lm(win1, LMAX, amodes.nside)
        phi2, w2lm = phi1, w1lm  # quirk: second transform also from win1
        ang = calc_angular_mixing_matrix(amodes.lmax, w1lm, w2lm)
        R1 = calc_radial_mixing(amodes.lmax, amodes.nmax_l, G, phi1)
        R2 = calc_radial_mixing(amodes.lmax, amodes.nmax_l, G, phi2)
        M = np.empty((lnnsize, lnnsize))
        for i_ in range(1, lnnsize + 1):
            for i in range(1, lnnsize + 1):
                M[i - 1, i_ - 1] = calc_cmixii_separable(i, i_, cmodes, R1, R2, ang, div2Lp1, interchange)
    else:
        W1 = optimize_Wr_lm_layout(calc_Wr_lm(win1, LMAX, amodes.nside), LMAX)
        W = calc_Wrl_Wrl(W1, W1, LMAX)
        M = calc_cmix_blocked(cmodes, G, W, div2Lp1, interchange)
    out = M
    if wtilde is not None:
        out = np.asarray(wtilde) @ out
    if v is not None:
        out = out @ np.asarray(v)
    assert np.all(np.isfinite(out))
    return out
