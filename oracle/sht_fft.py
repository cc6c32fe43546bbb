"""FFT-based CPU restatement of map2alm(niter) — the libsharp-like algorithm (ring FFT with exact aliasing
+ Legendre step) that the reference reaches through Healpix.jl / libsharp2.

TEST INFRASTRUCTURE / CPU BASELINE (see oracle/__init__.py).  Same numbers as oracle.healpix.SHT (exact sums)
up to rounding; used (a) as the stage-1 leg of bench.py's cpu_baseline (the exact-sum oracle is O(npix·lmsize)
and not representative of the reference's speed) and (b) as the checker at nside >= 64.

Follows the published libsharp ring-helper algorithm (Reinecke & Seljebotn 2013; libsharp2 1.0.2
`ringhelper_ring2phase` / `ringhelper_phase2ring`): F_m = X[m mod nφ] (or its conjugate mirror) · e^{-imφ0}.
"""
import math

import numpy as np

from . import healpix as hp


class FastSHT:
    def __init__(self, nside, lmax):
        if lmax > 4 * nside:
            raise ValueError("lmax > 4*nside is a poor choice")
        self.nside, self.lmax = nside, lmax
        self.info = hp.RingInfo(nside)
        self.npix = hp.nside2npix(nside)
        nhalf = 2 * nside
        self.nhalf = nhalf
        self.lam = hp.lambda_lm_table(lmax, self.info.z[:nhalf], self.info.sth[:nhalf])  # [lmsize, nhalf]
        self.m = np.arange(lmax + 1)
        # groups of rings sharing nφ: (ring indices, nφ)
        groups = {}
        for ring in range(self.info.nrings):
            groups.setdefault(int(self.info.nphi[ring]), []).append(ring)
        self.groups = [(np.array(r), n) for n, r in groups.items()]
        l = np.concatenate([np.arange(m, lmax + 1) for m in range(lmax + 1)])
        mm = np.concatenate([np.full(lmax + 1 - m, m) for m in range(lmax + 1)])
        self.parity = ((l - mm) & 1).astype(bool)  # per lm (m-major)

    def _ring_spectrum_index(self, nphi):
        idx = self.m % nphi
        mirror = idx > nphi // 2
        src = np.where(mirror, nphi - idx, idx)
        return src, mirror

    def _phase(self, rings, sign):
        return np.exp(sign * 1j * np.outer(self.info.phi0[rings], self.m))  # [ring, m]

    def ring_analysis(self, maps):
        """F[ring, m, shell] = Σ_j f_j e^{-imφ_j}"""
        nshell = maps.shape[0]
        info = self.info
        F = np.empty((info.nrings, self.lmax + 1, nshell), dtype=complex)
        for rings, nphi in self.groups:
            cols = (info.start[rings][:, None] + np.arange(nphi)[None, :]).ravel()
            X = np.fft.rfft(maps[:, cols].reshape(nshell, rings.size, nphi), axis=2)  # [shell, ring, k]
            src, mirror = self._ring_spectrum_index(nphi)
            V = X[:, :, src]
            V = np.where(mirror[None, None, :], np.conj(V), V)
            V = V * self._phase(rings, -1.0)[None, :, :]
            F[rings] = np.transpose(V, (1, 2, 0))
        return F

    def ring_synthesis(self, G):
        """maps[shell, pix] = Σ_m (2-δ_m0) Re(G[ring, m, shell] e^{imφ_j})"""
        nshell = G.shape[2]
        info = self.info
        maps = np.empty((nshell, self.npix))
        cm = np.where(self.m == 0, 1.0, 2.0)
        for rings, nphi in self.groups:
            V = G[rings] * (cm[None, :] * self._phase(rings, +1.0))[:, :, None]  # [ring, m, shell]
            Y = np.zeros((rings.size, nphi, nshell), dtype=complex)
            np.add.at(Y, (slice(None), self.m % nphi), V)
            f = np.fft.ifft(Y, axis=1).real * nphi  # [ring, j, shell]
            cols = (info.start[rings][:, None] + np.arange(nphi)[None, :]).ravel()
            maps[:, cols] = np.transpose(f, (2, 0, 1)).reshape(nshell, -1)
        return maps

    def adjoint_synthesis(self, maps):
        maps = np.atleast_2d(np.asarray(maps, dtype=float))
        F = self.ring_analysis(maps)
        nh, nr_ = self.nhalf, self.info.nrings
        Fn = F[:nh]
        Fs = np.zeros_like(Fn)
        Fs[:nh - 1] = F[nr_ - 1:nh - 1:-1]
        Fp, Fm = Fn + Fs, Fn - Fs
        lmax = self.lmax
        alm = np.empty((maps.shape[0], hp.getlmsize(lmax)), dtype=complex)
        w = 4 * math.pi / self.npix
        for m in range(lmax + 1):
            i0 = hp.lm_index_mmajor(lmax, m, m)
            sl = slice(i0, i0 + lmax + 1 - m)
            lam = self.lam[sl]
            odd = self.parity[sl]
            out = np.empty((lmax + 1 - m, maps.shape[0]), dtype=complex)
            out[~odd] = lam[~odd] @ Fp[:, m, :]
            out[odd] = lam[odd] @ Fm[:, m, :]
            alm[:, sl] = (w * out).T
        return alm

    def synthesis(self, alm):
        alm = np.atleast_2d(np.asarray(alm, dtype=complex))
        lmax, nh, nr_ = self.lmax, self.nhalf, self.info.nrings
        G = np.empty((nr_, lmax + 1, alm.shape[0]), dtype=complex)
        for m in range(lmax + 1):
            i0 = hp.lm_index_mmajor(lmax, m, m)
            sl = slice(i0, i0 + lmax + 1 - m)
            lam = self.lam[sl]
            odd = self.parity[sl]
            a = alm[:, sl].T
            E = lam[~odd].T @ a[~odd]
            O = lam[odd].T @ a[odd] if odd.any() else 0.0
            G[:nh, m, :] = E + O
            G[nr_ - 1:nh - 1:-1, m, :] = (E - O)[:nh - 1]
        return self.ring_synthesis(G)

    def map2alm(self, maps, niter=3):
        maps = np.atleast_2d(np.asarray(maps, dtype=float))
        alm = self.adjoint_synthesis(maps)
        for _ in range(niter):
            alm = alm + self.adjoint_synthesis(maps - self.synthesis(alm))
        return alm
