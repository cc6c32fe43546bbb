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
