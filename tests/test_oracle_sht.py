"""Pins of the HEALPix / SHT restatement (arithmetic that lives outside the reference tree)."""
import math

import numpy as np
import pytest
from scipy.special import sph_harm_y

from oracle import healpix as hp
from oracle import sht_fft


def test_ring_geometry_known_values():
    # published HEALPix pixel centres (nside=2 RING): pix 0, 4, 12
    th, ph = hp.pix2ang_ring(2, [0, 4, 12])
    assert np.allclose(th, [0.41113786, 0.84106867, 1.23095942], atol=1e-8)
    assert np.allclose(ph, [math.pi / 4, math.pi / 8, 0.0], atol=1e-12)
    th, ph = hp.pix2ang_ring(1, np.arange(12))
    assert np.allclose(np.cos(th[:4]), 2 / 3) and np.allclose(ph[:4], math.pi / 4 + np.arange(4) * math.pi / 2)
    for ns in (1, 2, 8):
        info = hp.RingInfo(ns)
        assert info.nphi.sum() == 12 * ns * ns and np.all(np.diff(info.start) == info.nphi[:-1])
        assert np.allclose(info.z, -info.z[::-1])


def test_nest2ring_known_values_and_hierarchy():
    assert list(hp.nest2ring(2, np.arange(8))) == [13, 5, 4, 0, 15, 7, 6, 1]   # published nside=2 table
    assert list(hp.nest2ring(1, np.arange(12))) == list(range(12))
    for ns in (2, 4, 16):
        assert sorted(hp.nest2ring(ns, np.arange(12 * ns * ns))) == list(range(12 * ns * ns))
    ns = 8
    v = lambda t, p: np.stack([np.sin(t) * np.cos(p), np.sin(t) * np.sin(p), np.cos(t)], -1)
    par = v(*hp.pix2ang_ring(ns, hp.nest2ring(ns, np.arange(12 * ns * ns))))
    chi = v(*hp.pix2ang_ring(2 * ns, hp.nest2ring(2 * ns, np.arange(48 * ns * ns))))
    d = np.arccos(np.clip((np.repeat(par, 4, axis=0) * chi).sum(-1), -1, 1))
    assert d.max() < 0.6 * math.sqrt(4 * math.pi / (12 * ns * ns))   # children sit inside their parent


def test_udgrade():
    rng = np.random.default_rng(0)
    m = rng.random((2, 12 * 4 * 4))
    assert np.array_equal(hp.udgrade(m, 4), m)
    up = hp.udgrade(m, 16)
    assert np.allclose(hp.udgrade(up, 4), m)              # mean of identical children
    assert np.isclose(up.mean(), m.mean())
    assert np.allclose(hp.udgrade(np.ones(12 * 64), 2), 1.0)


def test_lambda_vs_scipy():
    ns, lmax = 8, 24
    info = hp.RingInfo(ns)
    lam = hp.lambda_lm_table(lmax, info.z, info.sth)
    theta = np.arctan2(info.sth, info.z)
    for l in range(0, lmax + 1, 3):
        for m in range(0, l + 1, 2):
            assert np.abs(sph_harm_y(l, m, theta, 0.0).real - lam[hp.lm_index_mmajor(lmax, l, m)]).max() < 1e-13


def test_alm_index_orders():
    lmax = 5
    seen = sorted(hp.lm_index_mmajor(lmax, l, m) for l in range(lmax + 1) for m in range(l + 1))
    assert seen == list(range(hp.getlmsize(lmax)))
    assert [hp.lm_index_mmajor(lmax, l, 0) for l in range(3)] == [0, 1, 2] and hp.lm_index_mmajor(lmax, 1, 1) == lmax + 1
    assert [hp.lm_index_mfast(l, m) for l in range(3) for m in range(l + 1)] == list(range(6))


def test_synthesis_is_ylm_sum_and_full_sky():
    ns, lmax = 8, 16
    sht = hp.SHT(ns, lmax)
    rng = np.random.default_rng(1)
    alm = rng.normal(size=(1, hp.getlmsize(lmax))) + 1j * rng.normal(size=(1, hp.getlmsize(lmax)))
    alm[:, :lmax + 1] = alm[:, :lmax + 1].real
    th, ph = hp.pix2ang_ring(ns, np.arange(sht.npix))
    ref = np.zeros(sht.npix)
    for l in range(lmax + 1):
        for m in range(l + 1):
            ref += (1 if m == 0 else 2) * (alm[0, hp.lm_index_mmajor(lmax, l, m)] * sph_harm_y(l, m, th, ph)).real
    assert np.abs(sht.synthesis(alm)[0] - ref).max() < 1e-12
    a = sht.map2alm(np.ones((1, sht.npix)))
    assert abs(a[0, 0] - math.sqrt(4 * math.pi)) < 1e-4 and np.abs(a[0, 1:]).max() < 1e-3
    # single-pixel map: a_lm ≈ conj(Y_lm(pix)) Ω_pix   (test/test_cat2anlm.jl:153-196, rtol 1e-5 there at nside=256)
    pix = 137
    m1 = np.zeros((1, sht.npix))
    m1[0, pix] = 1.0
    a0 = sht.map2alm(m1, niter=0)[0]
    for (l, m) in ((0, 0), (3, 1), (7, 7), (16, 4)):
        want = np.conj(sph_harm_y(l, m, th[pix], ph[pix])) * 4 * math.pi / sht.npix
        assert abs(a0[hp.lm_index_mmajor(lmax, l, m)] - want) < 1e-14


def test_jacobi_iterations_converge_but_not_by_three():
    # the iteration count is part of the numerical contract
    ns, lmax = 8, 16
    sht = hp.SHT(ns, lmax)
    rng = np.random.default_rng(2)
    alm = rng.normal(size=(1, hp.getlmsize(lmax))) + 1j * rng.normal(size=(1, hp.getlmsize(lmax)))
    alm[:, :lmax + 1] = alm[:, :lmax + 1].real
    m = sht.synthesis(alm)
    err = [np.linalg.norm(sht.map2alm(m, niter=k) - alm) / np.linalg.norm(alm) for k in (0, 1, 2, 3, 4)]
    assert all(e2 < 0.5 * e1 for e1, e2 in zip(err, err[1:]))
    assert 1e-7 < err[3] < 1e-3 and err[0] > 1e-3


@pytest.mark.parametrize("ns,lmax", [(4, 16), (8, 20), (16, 40)])
def test_fft_sht_equals_exact_sum(ns, lmax):
    rng = np.random.default_rng(ns)
    m = rng.random((2, 12 * ns * ns))
    m[:, ::5] = 0
    a, b = hp.SHT(ns, lmax), sht_fft.FastSHT(ns, lmax)
    x, y = a.map2alm(m), b.map2alm(m)
    assert np.linalg.norm(x - y) / np.linalg.norm(x) < 1e-13
    assert np.linalg.norm(a.synthesis(x) - b.synthesis(x)) / np.linalg.norm(m) < 1e-13
    with pytest.raises(ValueError):
        hp.SHT(ns, 4 * ns + 1)


def test_belt_rings_reduce_to_gram_matrices():
    # Design check for DESIGN §9.2: on alias-free rings (nφ > 2 lmax) the composition analysis∘synthesis is the
    # data-independent, m-block-diagonal operator K_m[l][l'] = w nφ Σ_rings λ_lm(θ) λ_l'm(θ) (even and odd l decouple).
    import math
    from oracle import healpix as ohp
    from oracle.sht_fft import FastSHT
    nside, lmax = 8, 12
    sht = FastSHT(nside, lmax)
    info = sht.info
    belt = np.flatnonzero(info.nphi > 2 * lmax)
    assert belt.size >= 2 * nside + 1                      # at least the whole equatorial belt
    rng = np.random.default_rng(4)
    nlm = ohp.getlmsize(lmax)
    alm = rng.standard_normal((1, nlm)) + 1j * rng.standard_normal((1, nlm))
    for l in range(lmax + 1):
        alm[0, ohp.lm_index_mmajor(lmax, l, 0)] = alm[0, ohp.lm_index_mmajor(lmax, l, 0)].real   # real map
    m_syn = sht.synthesis(alm)
    keep = np.zeros(sht.npix, dtype=bool)
    for ring in belt:
        keep[info.start[ring]:info.start[ring] + info.nphi[ring]] = True
    ref = sht.adjoint_synthesis(np.where(keep, m_syn, 0.0))          # A(belt part of S(alm))
    w = 4 * math.pi / sht.npix
    nh = sht.nhalf
    got = np.zeros_like(ref)
    north = np.flatnonzero(info.nphi[:nh] > 2 * lmax)                 # belt rings of the northern half incl. equator
    mult = np.where(north == nh - 1, 1.0, 2.0)                        # N/S pairs count twice, the equator once
    nphi = info.nphi[north].astype(float)
    for m in range(lmax + 1):
        i0 = ohp.lm_index_mmajor(lmax, m, m)
        sl = slice(i0, i0 + lmax + 1 - m)
        lam = sht.lam[sl][:, north]                                   # [l, ring]
        odd = sht.parity[sl]
        K = w * (lam * (nphi * mult)[None, :]) @ lam.T                # [l, l']
        K[np.ix_(~odd, odd)] = 0.0                                    # opposite parities cancel between N and S
        K[np.ix_(odd, ~odd)] = 0.0
        got[0, sl] = K @ alm[0, sl]
    assert np.linalg.norm(got - ref) / np.linalg.norm(ref) < 1e-12


def test_alias_operator_class_sum_form():
    # Design check for DESIGN §9.1: the ring alias operator of ring_alias_kernel (csrc/sht.cu) in O(lmax) per ring.
    # F'_m = nφ σ^q [ (P[ρ][0] + σ P[ρ][1]) + σ^{[ρ≠0]} conj(P[(nφ-ρ) mod nφ][0] + σ P[.][1]) ],  ρ = m mod nφ, q = m div nφ,
    # P[ρ][par] = Σ_{m' ≡ ρ, (m' div nφ) ≡ par} c_{m'} G_{m'},  c_0 = 1/2.
    rng = np.random.default_rng(0)
    lmax = 37
    for n, shift in [(4, 0), (4, 1), (8, 0), (12, 1), (20, 0), (36, 1), (40, 0), (76, 1)]:
        G = rng.standard_normal(lmax + 1) + 1j * rng.standard_normal(lmax + 1)
        sig = -1.0 if shift else 1.0
        F = np.zeros(lmax + 1, complex)                       # the kernel's double loop
        for m in range(lmax + 1):
            re = im = 0.0
            for mp in range(m % n, lmax + 1, n):
                s = (sig if ((mp - m) // n) & 1 else 1.0) * (0.5 if mp == 0 else 1.0)
                re += s * G[mp].real
                im += s * G[mp].imag
            for mp in range((n - m % n) % n, lmax + 1, n):
                s = (sig if ((mp + m) // n) & 1 else 1.0) * (0.5 if mp == 0 else 1.0)
                re += s * G[mp].real
                im -= s * G[mp].imag
            F[m] = n * (re + 1j * im)
        P = np.zeros((n, 2), complex)
        for mp in range(lmax + 1):
            P[mp % n, (mp // n) & 1] += (0.5 if mp == 0 else 1.0) * G[mp]
        F2 = np.zeros(lmax + 1, complex)
        for m in range(lmax + 1):
            rho, q = m % n, m // n
            plus = P[rho, 0] + sig * P[rho, 1]
            rp = (n - rho) % n
            minus = np.conj(P[rp, 0] + sig * P[rp, 1]) * (sig if rho != 0 else 1.0)
            F2[m] = n * (sig ** q) * (plus + minus)
        assert np.abs(F - F2).max() < 1e-13
