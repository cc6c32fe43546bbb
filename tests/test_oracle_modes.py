"""Oracle mode tables against the reference's own test properties (test/test_modes.jl) and SURVEY sizes."""
import numpy as np
import pytest

from oracle import modes as om


@pytest.fixture(scope="module")
def cfg1():
    a = om.AnlmModes(0.05, 500.0, 1000.0)
    return a, om.ClnnModes(a)


def test_cfg1_sizes(cfg1):
    # lmax=43, nmax=8, nmax_l = [8x14, 7x7, 6x4, 5x3, 4x3, 3x4, 2x4, 1x5], nlsize=237, lnnsize=900
    a, c = cfg1
    assert (a.lmax, a.nmax, a.nside) == (43, 8, 64)
    expect = [8] * 14 + [7] * 7 + [6] * 4 + [5] * 3 + [4] * 3 + [3] * 4 + [2] * 4 + [1] * 5
    assert list(a.nmax_l) == expect
    assert int(a.nmax_l.sum()) == 237 and om.getlnnsize(c) == 900
    assert np.all(a.knl[np.isfinite(a.knl)] <= 0.05)


def test_anlm_index_roundtrip(cfg1):
    # test/test_modes.jl:31-37: getidx ∘ getnlm = id, contiguous
    a, _ = cfg1
    size = om.getnlmsize(a)
    for idx in list(range(1, 200)) + [size - 1, size]:
        n, l, m = om.getnlm(a, idx)
        assert om.getidx_nlm(a, n, l, m) == idx
    assert om.getlmsize(3) == 10


@pytest.mark.parametrize("dnmax", [None, 0, 1, 3])
def test_clnn_index_roundtrip(dnmax):
    # test/test_modes.jl:64-80: getidx(cmodes, getlnn(i)...) == i and idxmax == lnnsize
    a = om.AnlmModes(0.03, 500.0, 1000.0)
    c = om.ClnnModes(a) if dnmax is None else om.ClnnModes(a, dnmax=dnmax)
    n = om.getlnnsize(c)
    lnn = c.lnn
    assert np.all(lnn[1] <= lnn[2])
    key = lnn[0] * 10 ** 6 + (lnn[2] - lnn[1]) * 10 ** 3 + lnn[1]
    assert np.all(np.diff(key) > 0)          # sorted by (l, Δn, n1), no duplicates
    if dnmax is None:
        for i in range(1, n + 1):
            assert om.getidx_lnn(c, *om.getlnn(c, i)) == i
        assert n == int(sum(k * (k + 1) // 2 for k in a.nmax_l))
    else:
        assert int((lnn[2] - lnn[1]).max()) <= dnmax


def test_fixed_nl_modes():
    a = om.AnlmModes(3, 5, 500.0, 1000.0)
    assert a.nside == 8 and a.knl.shape == (3, 6) and np.all(np.diff(a.knl, axis=0) > 0)
    c = om.ClnnModes(a, dnmax=1)
    assert om.getlnnsize(c) == 6 * (3 + 2)


def test_gnl_orthonormal():
    # test/test_gnl.jl:27-41 (atol 1e-6 with Gauss-Legendre there); midpoint rule with many nodes here
    a = om.AnlmModes(0.03, 500.0, 1000.0)
    nr = 4000
    dr = 500.0 / nr
    r = 500.0 + dr * (np.arange(nr) + 0.5)
    for l in (0, 3, a.lmax):
        nl = int(a.nmax_l[l])
        g = np.array([a.basisfunctions(n, l, r) for n in range(1, nl + 1)])
        gram = (g * r ** 2 * dr) @ g.T
        assert np.abs(gram - np.eye(nl)).max() < 1e-6


def test_binning_weights():
    # test/test_modes.jl:141-159,213-214
    a = om.AnlmModes(0.03, 500.0, 1000.0)
    c = om.ClnnModes(a)
    n = om.getlnnsize(c)
    wt, v = om.bandpower_binning_weights(c, dl=1, dn1=1, dn2=1)
    assert np.array_equal(wt, np.eye(n))
    wt, v = om.bandpower_binning_weights(c, dl=3, dn1=2, dn2=2)
    assert wt.shape[1] == n and v.shape == wt.T.shape
    assert np.allclose(wt.sum(axis=1), 1) and np.allclose(wt @ v, np.eye(wt.shape[0]))
    b = om.ClnnBinnedModes(wt, v, c)
    assert b.LKK.shape == (3, wt.shape[0]) and np.all(b.LKK[1] <= b.LKK[2])
