"""Oracle restatement of the HEALPix pieces the reference reaches through Healpix.jl.

TEST INFRASTRUCTURE (see oracle/__init__.py).

The arithmetic here lives OUTSIDE /root/reference: Healpix.jl 4.2.2 ->
Libsharp.jl 0.2.0 -> libsharp2_jll 1.0.2+2 (reference Manifest.toml:359-363,
508-512,1012-1016).  It is restated from the published HEALPix definition
(Gorski et al. 2005: RING scheme geometry, NESTED face/bit layout) and the
published libsharp algorithm (Reinecke & Seljebotn 2013: ring DFT + Legendre
step, exact aliasing for short rings), anchored on the reference's call sites:

  src/healpix_helpers.jl:40-45   udgrade(map::Vector, nside)   -> Healpix.udgrade
  src/healpix_helpers.jl:50-71   mymap2alm(map; lmax)          -> Healpix.map2alm(map, lmax=lmax)
                                  niter=3 (default; explicit at :50), uniform
                                  pixel weights 4π/npix (ring weights disabled, :64-69),
                                  error if lmax > 4 nside (:60-63)
  src/windows.jl:867              Healpix.alm2cl
  src/LMcalcStructs.jl:7-18       alm storage order (m-major, m >= 0)

Conventions: Y_lm with Condon-Shortley phase; a_lm = Σ_p f_p conj(Y_lm(p)) 4π/npix;
map2alm(niter): a <- A f ; repeat niter times: a <- a + A (f - S a).
The ring transforms are exact sums over the true pixel longitudes (what
libsharp's FFT + aliasing evaluates, up to rounding).
"""
import math

import numpy as np


# ----------------------------------------------------------------------------
# RING geometry

def nside2npix(nside):
    return 12 * nside * nside


def npix2nside(npix):
    nside = int(round(math.sqrt(npix / 12)))
    assert 12 * nside * nside == npix
    return nside


class RingInfo:
    """Per-ring tables for rings 1..4nside-1 (index 0 = northernmost)."""

    def __init__(self, nside):
        self.nside = nside
        nrings = 4 * nside - 1
        self.nrings = nrings
        self.nphi = np.zeros(nrings, dtype=np.int64)
        self.start = np.zeros(nrings, dtype=np.int64)
        self.z = np.zeros(nrings)
        self.sth = np.zeros(nrings)
        self.phi0 = np.zeros(nrings)
        npix = nside2npix(nside)
        ncap = 2 * nside * (nside - 1)
        for idx in range(nrings):
            i = idx + 1
            northring = i if i <= 2 * nside else 4 * nside - i
            if northring < nside:
                nphi = 4 * northring
                omz = northring * northring / (3.0 * nside * nside)  # 1 - z
                z = 1.0 - omz
                sth = math.sqrt(omz * (1.0 + z))
                phi0 = math.pi / nphi  # (1 - 1/2) * 2π / nphi
                start = 2 * northring * (northring - 1)
            else:
                nphi = 4 * nside
                z = (2 * nside - northring) * 2.0 / (3.0 * nside)
                sth = math.sqrt((1.0 - z) * (1.0 + z))
                shifted = ((northring - nside) & 1) == 0
                phi0 = math.pi / nphi if shifted else 0.0
                start = ncap + (northring - nside) * 4 * nside
            if i > 2 * nside:  # southern mirror
                z = -z
                start = npix - start - nphi
            self.nphi[idx] = nphi
            self.start[idx] = start
            self.z[idx] = z
            self.sth[idx] = sth
            self.phi0[idx] = phi0


def pix2ang_ring(nside, pix):
    """θ, φ of 0-based RING pixel indices (vectorised)."""
    info = RingInfo(nside)
    pix = np.asarray(pix, dtype=np.int64)
    ring = np.searchsorted(info.start, pix, side="right") - 1
    j = pix - info.start[ring]
    theta = np.arctan2(info.sth[ring], info.z[ring])
    phi = info.phi0[ring] + j * (2 * math.pi / info.nphi[ring])
    return theta, phi


# ----------------------------------------------------------------------------
# NESTED <-> RING (for udgrade)

_JRLL = np.array([2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4], dtype=np.int64)
_JPLL = np.array([1, 3, 5, 7, 0, 2, 4, 6, 1, 3, 5, 7], dtype=np.int64)


def _compress_bits(v):
    """Take the even bits of v and pack them."""
    v = v & 0x5555555555555555
    v = (v | (v >> 1)) & 0x3333333333333333
    v = (v | (v >> 2)) & 0x0F0F0F0F0F0F0F0F
    v = (v | (v >> 4)) & 0x00FF00FF00FF00FF
    v = (v | (v >> 8)) & 0x0000FFFF0000FFFF
    v = (v | (v >> 16)) & 0x00000000FFFFFFFF
    return v


def nest2ring(nside, ipnest):
    ipnest = np.asarray(ipnest, dtype=np.int64)
    npface = nside * nside
    npix = 12 * npface
    ncap = 2 * nside * (nside - 1)
    nl4 = 4 * nside
    face = ipnest // npface
    ipf = ipnest % npface
    ix = _compress_bits(ipf)
    iy = _compress_bits(ipf >> 1)
    jrt = ix + iy
    jpt = ix - iy
    jr = _JRLL[face] * nside - jrt - 1
    nr = np.where(jr < nside, jr, np.where(jr > 3 * nside, nl4 - jr, nside))
    n_before = np.where(jr < nside, 2 * nr * (nr - 1),
                        np.where(jr > 3 * nside, npix - 2 * (nr + 1) * nr,
                                 ncap + (jr - nside) * nl4))
    kshift = np.where((jr < nside) | (jr > 3 * nside), 0, (jr - nside) & 1)
    jp = (_JPLL[face] * nr + jpt + 1 + kshift) // 2
    jp = np.where(jp > nl4, jp - nl4, jp)
    jp = np.where(jp < 1, jp + nl4, jp)
    return n_before + jp - 1


def udgrade(map_ring, nside_out):
    """Healpix.udgrade for a RING-ordered map: identity if nside is unchanged,
    up-grade copies the NESTED parent's value into each child, down-grade
    averages the children."""
    map_ring = np.asarray(map_ring, dtype=float)
    nside_in = npix2nside(map_ring.shape[-1])
    if nside_out == nside_in:
        return map_ring.copy()
    r_in = nest2ring(nside_in, np.arange(nside2npix(nside_in)))
    r_out = nest2ring(nside_out, np.arange(nside2npix(nside_out)))
    nest_in = map_ring[..., r_in]
    if nside_out > nside_in:
        ratio = (nside_out // nside_in) ** 2
        nest_out = np.repeat(nest_in, ratio, axis=-1)
    else:
        ratio = (nside_in // nside_out) ** 2
        nest_out = nest_in.reshape(nest_in.shape[:-1] + (-1, ratio)).mean(axis=-1)
    out = np.empty(map_ring.shape[:-1] + (nside2npix(nside_out),))
    out[..., r_out] = nest_out
    return out


# ----------------------------------------------------------------------------
# alm indexing  (src/LMcalcStructs.jl)

def getlmsize(lmax):
    return (lmax + 1) * (lmax + 2) // 2


def lm_index_mmajor(lmax, l, m):
    """0-based position in HEALPix order [(0,0),(1,0),..,(lmax,0),(1,1),..]  (LMcalcStructs.jl:13-18)."""
    return l + ((m * (2 * lmax + 1 - m)) >> 1)


def lm_index_mfast(l, m):
    """0-based position in m-fast order [(0,0),(1,0),(1,1),(2,0),...]  (LMcalcStructs.jl:24-26)."""
    return m + (l * (l + 1)) // 2


# ----------------------------------------------------------------------------
# Legendre tables

def lambda_lm_table(lmax, z, sth):
    """λ_lm(θ) = sqrt((2l+1)/(4π) (l-m)!/(l+m)!) P_lm(cosθ) (with Condon-Shortley phase)
    for all rings: returns array [lmsize(m-major), nrings]."""
    z = np.asarray(z, dtype=float)
    sth = np.asarray(sth, dtype=float)
    out = np.zeros((getlmsize(lmax), z.size))
    lmm = np.full(z.size, math.sqrt(1.0 / (4 * math.pi)))  # λ_00
    for m in range(lmax + 1):
        if m > 0:
            lmm = -lmm * sth * math.sqrt((2 * m + 1) / (2.0 * m))
        out[lm_index_mmajor(lmax, m, m)] = lmm
        if m == lmax:
            break
        l1 = z * math.sqrt(2 * m + 3) * lmm
        out[lm_index_mmajor(lmax, m + 1, m)] = l1
        l2 = lmm
        for l in range(m + 2, lmax + 1):
            a = math.sqrt((4.0 * l * l - 1) / (l * l - m * m))
            b = math.sqrt(((l - 1.0) ** 2 - m * m) / (4.0 * (l - 1) ** 2 - 1))
            lnew = a * (z * l1 - b * l2)
            out[lm_index_mmajor(lmax, l, m)] = lnew
            l2, l1 = l1, lnew
    return out


# ----------------------------------------------------------------------------
# exact-sum SHT, batched over leading "shell" axis

class SHT:
    def __init__(self, nside, lmax):
        if lmax > 4 * nside:
            raise ValueError("lmax > 4*nside is a poor choice")  # src/healpix_helpers.jl:60-63
        self.nside, self.lmax = nside, lmax
        self.info = RingInfo(nside)
        self.npix = nside2npix(nside)
        self.lam = lambda_lm_table(lmax, self.info.z, self.info.sth)  # [lmsize, nrings]
        self.m = np.arange(lmax + 1)

    def _ring_phase(self, ring):
        info = self.info
        phi = info.phi0[ring] + np.arange(info.nphi[ring]) * (2 * math.pi / info.nphi[ring])
        return np.exp(-1j * np.outer(self.m, phi))  # [m, j] = e^{-imφ_j}

    def adjoint_synthesis(self, maps):
        """A f = (4π/npix) Σ_p f_p conj(Y_lm(p));  maps [nshell, npix] -> alm [nshell, lmsize] (m-major)."""
        maps = np.atleast_2d(np.asarray(maps, dtype=float))
        nshell = maps.shape[0]
        info, lmax = self.info, self.lmax
        F = np.zeros((info.nrings, lmax + 1, nshell), dtype=complex)
        for ring in range(info.nrings):
            s, n = info.start[ring], info.nphi[ring]
            F[ring] = self._ring_phase(ring) @ maps[:, s:s + n].T
        alm = np.zeros((nshell, getlmsize(lmax)), dtype=complex)
        w = 4 * math.pi / self.npix
        for m in range(lmax + 1):
            i0 = lm_index_mmajor(lmax, m, m)
            lam_m = self.lam[i0:i0 + lmax + 1 - m]  # [l, ring]
            alm[:, i0:i0 + lmax + 1 - m] = (w * (lam_m @ F[:, m, :])).T
        return alm

    def synthesis(self, alm):
        """S a: alm [nshell, lmsize] -> maps [nshell, npix] (real field, m >= 0 storage)."""
        alm = np.atleast_2d(np.asarray(alm, dtype=complex))
        nshell = alm.shape[0]
        info, lmax = self.info, self.lmax
        G = np.zeros((info.nrings, lmax + 1, nshell), dtype=complex)
        for m in range(lmax + 1):
            i0 = lm_index_mmajor(lmax, m, m)
            lam_m = self.lam[i0:i0 + lmax + 1 - m]  # [l, ring]
            G[:, m, :] = lam_m.T @ alm[:, i0:i0 + lmax + 1 - m].T
        G[:, 1:, :] *= 2.0
        maps = np.zeros((nshell, self.npix))
        for ring in range(info.nrings):
            s, n = info.start[ring], info.nphi[ring]
            E = np.conj(self._ring_phase(ring))  # e^{+imφ_j}, [m, j]
            maps[:, s:s + n] = np.real(E.T @ G[ring]).T
        return maps

    def map2alm(self, maps, niter=3):
        """Healpix.map2alm(map, lmax=lmax; niter=3) with uniform weights."""
        maps = np.atleast_2d(np.asarray(maps, dtype=float))
        alm = self.adjoint_synthesis(maps)
        for _ in range(niter):
            resid = maps - self.synthesis(alm)
            alm = alm + self.adjoint_synthesis(resid)
        return alm


def alm2cl(alm1, alm2, lmax):
    """Healpix.alm2cl on m-major alm  (same formula as
    src/SphericalFourierBesselDecompositions.jl:114-129)."""
    cl = np.zeros(lmax + 1)
    for l in range(lmax + 1):
        c = (alm1[lm_index_mmajor(lmax, l, 0)] * np.conj(alm2[lm_index_mmajor(lmax, l, 0)])).real
        for m in range(1, l + 1):
            i = lm_index_mmajor(lmax, l, m)
            c += 2 * (alm1[i] * np.conj(alm2[i])).real
        cl[l] = c / (2 * l + 1)
    return cl
