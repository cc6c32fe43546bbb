"""TEST INFRASTRUCTURE (see oracle/__init__.py): the SFB transforms next to the window path.

Follows (reference paths):
  src/windows.jl:244-270     win_rhat_ln (dense and separable)
  src/cat2anlm.jl:48-68      sortout
  src/cat2anlm.jl:85-99      make_pmu_pmupix, transform_gnl_spmap!
  src/cat2anlm.jl:257-314    cat2amln
  src/cat2anlm.jl:326-362    field2anlm (= field2anlm_v2)
  src/cat2anlm.jl:385-422    anlm2field
  src/SphericalFourierBesselDecompositions.jl:136-148   amln2clnn
Healpix.jl pieces (`map2alm!(map, alm)` with niter = 3, `alm2map!`, `ang2pixRing`) are restated in oracle/healpix.py;
the SHT defaults are pinned by tests/test_reference_golden.py.

Pinned by the reference's own tests: single-voxel field -> Δr r² g_nl(r) conj(Y_lm) Ω_p at rtol 1e-5
(test/test_cat2anlm.jl:153-196), field2anlm_v1 ≈ field2anlm_v2 (:235-251: the catalogue route with an empty catalogue
equals the field route), and the anlm2field/field2anlm round trips at rtol 1e-4 (:261-291).
"""
import math

import numpy as np

from . import healpix as hp
from . import modes as om
from .windows import SeparableArray, window_r


def win_rhat_ln(win, wmodes, amodes):
    """src/windows.jl:244-256 (dense): W_rhat_ln[p, l, n-1] = Δr Σ_r win[r,p] r² g_nl(r); NaN where l > lmax_n[n].
    Separable (:259-270): (mask, W_ln[l, n-1])."""
    r, dr = window_r(wmodes)
    g = amodes.basisfunctions
    if isinstance(win, SeparableArray):
        W_ln = np.full((amodes.lmax + 1, amodes.nmax), np.nan)
        for n in range(1, amodes.nmax + 1):
            for l in range(int(amodes.lmax_n[n - 1]) + 1):
                W_ln[l, n - 1] = dr * np.sum(r ** 2 * g(n, l, r) * win.phi)
        return win.mask, W_ln
    win = np.asarray(win, dtype=float)
    out = np.full((win.shape[1], amodes.lmax + 1, amodes.nmax), np.nan)
    for n in range(1, amodes.nmax + 1):
        for l in range(int(amodes.lmax_n[n - 1]) + 1):
            out[:, l, n - 1] = win.T @ (r ** 2 * g(n, l, r))
    return out * dr


def ang2pix_ring(nside, theta, phi):
    """Healpix ang2pixRing (0-based result): HEALPix ang2pix_ring algorithm (Gorski et al. 2005)."""
    theta = np.atleast_1d(np.asarray(theta, dtype=float))
    phi = np.atleast_1d(np.asarray(phi, dtype=float))
    z = np.cos(theta)
    za = np.abs(z)
    tt = np.mod(phi, 2 * math.pi) / (math.pi / 2)          # in [0, 4)
    nl4 = 4 * nside
    npix = 12 * nside * nside
    ncap = 2 * nside * (nside - 1)
    pix = np.empty(z.shape, dtype=np.int64)
    eq = za <= 2.0 / 3.0
    # equatorial region
    t1 = nside * (0.5 + tt[eq])
    t2 = nside * 0.75 * z[eq]
    jp = np.floor(t1 - t2).astype(np.int64)
    jm = np.floor(t1 + t2).astype(np.int64)
    ir = nside + 1 + jp - jm
    kshift = 1 - (ir & 1)
    ip = (jp + jm - nside + kshift + 1) // 2
    ip = np.mod(ip, nl4)
    pix[eq] = ncap + (ir - 1) * nl4 + ip
    # polar caps
    po = ~eq
    tp = tt[po] - np.floor(tt[po])
    tmp = nside * np.sqrt(3 * (1 - za[po]))
    jp = np.floor(tp * tmp).astype(np.int64)
    jm = np.floor((1.0 - tp) * tmp).astype(np.int64)
    ir = jp + jm + 1
    ip = np.floor(tt[po] * ir).astype(np.int64)
    ip = np.mod(ip, 4 * ir)
    north = z[po] > 0
    pix[po] = np.where(north, 2 * ir * (ir - 1) + ip, npix - 2 * ir * (ir + 1) + ip)
    return pix


def field2anlm(f_xyz, wmodes, amodes, niter=3):
    """src/cat2anlm.jl:326-362: per shell map2alm!(map, alm) (lmax = amodes.lmax, niter = 3) and
    f_nlm[(n,l,m)] += g_nl(r) r² Δr alm[l,m]."""
    f_xyz = np.asarray(f_xyz, dtype=float)
    r, dr = window_r(wmodes)
    lmax = amodes.lmax
    sht = hp.SHT(wmodes.nside, lmax)
    alm = sht.map2alm(f_xyz, niter=niter)                       # [nr, lmsize] m-major
    out = np.zeros(om.getnlmsize(amodes), dtype=complex)
    g = amodes.basisfunctions
    for n in range(1, amodes.nmax + 1):
        for l in range(int(amodes.lmax_n[n - 1]) + 1):
            wgt = g(n, l, r) * r ** 2 * dr
            base = om.getidx_nlm(amodes, n, l, 0) - 1
            for m in range(l + 1):
                out[base + m] = wgt @ alm[:, hp.lm_index_mmajor(lmax, l, m)]
    return out


def anlm2field(f_nlm, wmodes, amodes):
    """src/cat2anlm.jl:385-422: alm_r[l,m] = Σ_n g_nl(r) f_nlm[(n,l,m)], then alm2map per shell."""
    r, _ = window_r(wmodes)
    lmax = amodes.lmax
    g = amodes.basisfunctions
    alm = np.zeros((r.size, hp.getlmsize(lmax)), dtype=complex)
    for n in range(1, amodes.nmax + 1):
        for l in range(int(amodes.lmax_n[n - 1]) + 1):
            gr = g(n, l, r)
            base = om.getidx_nlm(amodes, n, l, 0) - 1
            for m in range(l + 1):
                alm[:, hp.lm_index_mmajor(lmax, l, m)] += gr * f_nlm[base + m]
    return hp.SHT(amodes.nside, lmax).synthesis(alm)


def cat2amln(rtp, amodes, nbar, wrhatln, weight=None, niter=3):
    """src/cat2anlm.jl:257-314.  rtp: 3 x Ngal (r, θ, φ); wrhatln: [npix, lmax+1, nmax] from win_rhat_ln."""
    rtp = np.asarray(rtp, dtype=float).reshape(3, -1)
    ngal = rtp.shape[1]
    weight = np.ones(ngal) if (weight is None or len(weight) == 0) else np.asarray(weight, dtype=float)
    p = np.argsort(rtp[0], kind="stable")                       # sortout: by r
    r, theta, phi, weight = rtp[0, p], rtp[1, p], rtp[2, p], weight[p] if ngal else weight
    nside, lmax = amodes.nside, amodes.lmax
    npix = hp.nside2npix(nside)
    pix = ang2pix_ring(nside, theta, phi) if ngal else np.zeros(0, dtype=np.int64)
    domega = 4 * math.pi / npix
    sht = hp.SHT(nside, lmax)
    g = amodes.basisfunctions
    out = np.full(om.getnlmsize(amodes), np.nan + 0j)
    for n in range(1, amodes.nmax + 1):
        for l in range(int(amodes.lmax_n[n - 1]) + 1):
            m_ = np.zeros(npix)
            if ngal:
                np.add.at(m_, pix, weight * g(n, l, r))
            m_ *= 1 / (nbar * domega)
            m_ = m_ - wrhatln[:, l, n - 1]
            alm = sht.map2alm(m_[None, :], niter=niter)[0]
            base = om.getidx_nlm(amodes, n, l, 0) - 1
            for m in range(l + 1):
                out[base + m] = alm[hp.lm_index_mmajor(lmax, l, m)]
    assert np.all(np.isfinite(out))
    return out


def amln2clnn(anlm1, anlm2, cmodes):
    """…Decompositions.jl:136-148 with alm2cl of the (l, 0..l) slices."""
    am = cmodes.amodes
    out = np.empty(cmodes.lnn.shape[1])
    for i in range(out.size):
        l, n1, n2 = (int(x) for x in cmodes.lnn[:, i])
        i1 = om.getidx_nlm(am, n1, l, 0) - 1
        i2 = om.getidx_nlm(am, n2, l, 0) - 1
        a1, a2 = anlm1[i1:i1 + l + 1], anlm2[i2:i2 + l + 1]
        out[i] = ((a1[0] * np.conj(a2[0])).real + 2 * np.sum((a1[1:] * np.conj(a2[1:])).real)) / (2 * l + 1)
    return out
