"""The five BASELINE.json configurations as concrete synthetic inputs.

All: boundary=potential, rmin=500, rmax=1000, ClnnModes(amodes) with Δnmax=∞, exact g_nl (cache=false), windows
normalised to max 1 like make_window (src/SphericalFourierBesselDecompositions.jl:467-475), seed 20240517.
Windows are returned in Julia memory order: a Fortran-ordered (nr, npix) float64 array.
"""
from __future__ import annotations

import math
import warnings

import numpy as np

from .modes import AnlmModes, ClnnModes
from .separable import SeparableArray
from .windows import ConfigurationSpaceModes, rsdrgnlr

SEED = 20240517
RMIN, RMAX = 500.0, 1000.0

CONFIGS = {
    1: dict(kmax=0.05, nside=None, win_nside=8, nr=128, kind="fullsky",
            desc="cfg1 kmax=0.05 nside8->64 full-sky nr=128"),
    2: dict(kmax=0.08, nside=64, win_nside=64, nr=30, kind="halfsky_radial",
            desc="cfg2 kmax=0.08 nside=64 half-sky x radial nr=30"),
    3: dict(kmax=0.12, nside=128, win_nside=128, nr=320, kind="separable_survey",
            desc="cfg3 kmax=0.12 nside=128 separable survey mask nr=320"),
    4: dict(kmax=0.15, nside=256, win_nside=256, nr=64, kind="nonseparable",
            desc="cfg4 kmax=0.15 nside=256 non-separable window nr=64"),
    5: dict(kmax=0.20, nside=512, win_nside=512, nr=64, kind="nonseparable",
            desc="cfg5 kmax=0.20 nside=512 non-separable window nr=64"),
    # reduced variants for parity tests at sizes the CPU oracle finishes in seconds
    "2s": dict(kmax=0.03, nside=32, win_nside=32, nr=30, kind="halfsky_radial", desc="cfg2-small"),
    "3s": dict(kmax=0.03, nside=32, win_nside=32, nr=80, kind="separable_survey", desc="cfg3-small"),
    "4s": dict(kmax=0.035, nside=32, win_nside=32, nr=16, kind="nonseparable", desc="cfg4-small"),
    "1s": dict(kmax=0.02, nside=None, win_nside=4, nr=96, kind="fullsky", desc="cfg1-small"),
}


def pix2ang_ring(nside, pix=None):
    """θ, φ of RING pixels (HEALPix definition; 0-based)."""
    npix = 12 * nside * nside
    pix = np.arange(npix, dtype=np.int64) if pix is None else np.asarray(pix, dtype=np.int64)
    ncap = 2 * nside * (nside - 1)
    theta = np.empty(pix.shape)
    phi = np.empty(pix.shape)
    north = pix < ncap
    south = pix >= npix - ncap
    belt = ~(north | south)
    for sel, flip in ((north, False), (south, True)):
        p = pix[sel] if not flip else npix - 1 - pix[sel]
        i = ((1 + np.sqrt(1 + 2 * p.astype(np.float64))) / 2).astype(np.int64)
        i = np.where(2 * i * (i - 1) > p, i - 1, i)
        i = np.where(2 * (i + 1) * i <= p, i + 1, i)
        j = p - 2 * i * (i - 1)
        z = 1.0 - i * i / (3.0 * nside * nside)
        ph = (j + 0.5) * (math.pi / 2) / i
        if flip:
            z = -z
            ph = 2 * math.pi - ph
        theta[sel] = np.arccos(z)
        phi[sel] = ph
    p = pix[belt] - ncap
    i = p // (4 * nside) + nside
    j = p % (4 * nside)
    z = (2 * nside - i) * 2.0 / (3.0 * nside)
    shift = np.where(((i - nside) & 1) == 0, 0.5, 0.0)
    theta[belt] = np.arccos(z)
    phi[belt] = (j + shift) * (math.pi / 2) / nside
    return theta, phi


def _phi_radial(r):
    return np.exp(-(r / (0.55 * RMAX)) ** 2)   # :radial, …Decompositions.jl:393-400


def _cap_mask(nside, fsky):
    theta, _ = pix2ang_ring(nside)
    return (theta <= math.acos(1 - 2 * fsky)).astype(np.float64)   # gen_mask, …Decompositions.jl:221-234


def _survey_mask(nside):
    theta, phi = pix2ang_ring(nside)
    vec = np.stack([np.sin(theta) * np.cos(phi), np.sin(theta) * np.sin(phi), np.cos(theta)], axis=1)

    def unit(t, p):
        return np.array([math.sin(t) * math.cos(p), math.sin(t) * math.sin(p), math.cos(t)])

    c = unit(math.radians(50), math.radians(30))
    mask = (vec @ c >= 1 - 2 * 0.3).astype(np.float64)   # cap with f_sky = 0.3
    rng = np.random.default_rng(SEED)
    for _ in range(20):
        t = math.acos(rng.uniform(-1, 1))
        p = rng.uniform(0, 2 * math.pi)
        mask[vec @ unit(t, p) >= math.cos(math.radians(2.0))] = 0.0
    return mask


def make_window(cfg, wmodes):
    kind, nside = cfg["kind"], cfg["win_nside"]
    r = wmodes.r
    npix = 12 * nside * nside
    if kind == "fullsky":
        return np.ones((r.size, npix), order="F")
    if kind == "halfsky_radial":
        win = np.asfortranarray(_phi_radial(r)[:, None] * _cap_mask(nside, 0.5)[None, :])
        return win / win.max()
    if kind == "separable_survey":
        mask = _survey_mask(nside)
        phi = _phi_radial(r)
        return SeparableArray(phi / phi.max(), mask / mask.max())
    if kind == "nonseparable":
        theta, phi = pix2ang_ring(nside)
        win = np.empty((r.size, npix), order="F")
        ang = 1 + 0.2 * np.cos(3 * phi) * np.sin(theta)
        thmax = np.radians(60 + 20 * (r - RMIN) / (RMAX - RMIN))
        pr = _phi_radial(r)
        for i in range(r.size):
            win[i, :] = np.clip(pr[i] * (theta <= thmax[i]) * ang, 0.0, 1.0)
        return win / win.max()
    raise ValueError(kind)


class Workload:
    """Modes, tables and window of one configuration."""

    def __init__(self, key, with_window=True):
        cfg = CONFIGS[key]
        self.key, self.cfg = key, cfg
        self.amodes = AnlmModes(cfg["kmax"], RMIN, RMAX, nside=cfg["nside"])
        self.cmodes = ClnnModes(self.amodes)
        self.wmodes = ConfigurationSpaceModes(RMIN, RMAX, cfg["nr"], cfg["win_nside"])
        with warnings.catch_warnings():
            warnings.simplefilter("ignore", RuntimeWarning)   # check_nsamp: the reference only warns
            self.G = rsdrgnlr(self.amodes, self.wmodes)
        self.win = make_window(cfg, self.wmodes) if with_window else None
        self.lnnsize = self.cmodes.lnn.shape[1]
        self.nr = cfg["nr"]
        self.LMAX = 2 * self.amodes.lmax

    def describe(self):
        a = self.amodes
        return dict(workload=self.cfg["desc"], kmax=self.cfg["kmax"], nside=a.nside, win_nside=self.cfg["win_nside"],
                    nr=self.nr, lmax=a.lmax, nmax=a.nmax, lnnsize=self.lnnsize, LMAX=self.LMAX)

    # ---- flop / byte models of SURVEY.md §8d -----------------------------------------------------
    def flops_alg_stage23(self):
        """F_alg = Σ_{l,L} [2 nr²(min(l,L)+1) + 2 nn nr² + 2 nn² nr], nn = nmax_l[l] nmax_l[L]."""
        nl = np.asarray(self.amodes.nmax_l, dtype=np.float64)
        l = np.arange(nl.size)
        nn = nl[:, None] * nl[None, :]
        nr = float(self.nr)
        return float(np.sum(2 * nr * nr * (np.minimum(l[:, None], l[None, :]) + 1) + 2 * nn * nr * nr + 2 * nn * nn * nr))

    def flops_bruteforce(self, rows=None):
        """Reference-order count Σ_{i,i'} s (min(l,L)+1)(2nr²+2nr), s = 2 if N≠N' else 1 (rows: 0-based subset)."""
        lnn = self.cmodes.lnn
        l = lnn[0].astype(np.float64)
        s = 1.0 + (lnn[1] != lnn[2])
        nr = float(self.nr)
        lr = l if rows is None else l[rows]
        # Σ_i Σ_i' s_i' (min(l_i, L_i')+1)
        Ls, cnt = np.unique(lnn[0], return_counts=True)
        sw = np.array([s[lnn[0] == L].sum() for L in Ls])
        tot = 0.0
        for li in np.unique(lr):
            tot += np.sum(lr == li) * np.sum(sw * (np.minimum(li, Ls) + 1))
        return tot * (2 * nr * nr + 2 * nr)

    def flops_alg_stage1(self):
        """F1 = nr (1+2 niter) [6 (4nside-1) lmsize + 5 npix log2(4 nside)]   (niter = 3)."""
        ns = self.amodes.nside
        lmsize = (self.LMAX + 1) * (self.LMAX + 2) // 2
        return self.nr * 7.0 * (6.0 * (4 * ns - 1) * lmsize + 5.0 * 12 * ns * ns * math.log2(4 * ns))

    def bytes_alg_stage1(self):
        ns = self.amodes.nside
        lmsize = (self.LMAX + 1) * (self.LMAX + 2) // 2
        return self.nr * 7.0 * (8.0 * 12 * ns * ns + 16.0 * lmsize)
