"""Host-side mirror (product code) against the oracle: index tables bit-exact, C-ABI surface, loud failure."""
import os
import re

import numpy as np
import pytest

from oracle import modes as om

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("args", [(0.05, 500.0, 1000.0), (0.03, 500.0, 1000.0), (3, 5, 500.0, 1000.0), (0.01, 0.0, 1000.0)])
def test_mode_tables_bit_exact(sfb, args):
    a, oa = sfb.AnlmModes(*args), om.AnlmModes(*args)
    assert (a.nmax, a.lmax, a.nside) == (oa.nmax, oa.lmax, oa.nside)
    assert np.array_equal(a.nmax_l, oa.nmax_l) and np.array_equal(a.lmax_n, oa.lmax_n)
    assert a.nmax_l.dtype == np.int64
    assert np.allclose(a.knl, oa.knl, rtol=1e-9, atol=0, equal_nan=True)       # xrtol of the zero finder is 1e-10
    for dn in (None, 0, 2):
        c = sfb.ClnnModes(a, dnmax=dn)
        oc = om.ClnnModes(oa) if dn is None else om.ClnnModes(oa, dnmax=dn)
        assert c.lnn.dtype == np.int64 and c.lnn.flags.f_contiguous
        assert np.array_equal(c.lnn, oc.lnn) and np.array_equal(c.first_ell_idx, oc.first_ell_idx)
        assert sfb.getlnnsize(c) == om.getlnnsize(oc)
    c = sfb.ClnnModes(a)
    n = sfb.getlnnsize(c)
    for i in list(range(1, min(n, 60) + 1)) + [n]:
        assert sfb.getidx(c, *sfb.getlnn(c, i)) == i            # test/test_modes.jl:64-80
    for idx in (1, 2, 17, sfb.getnlmsize(a)):
        assert sfb.getidx(a, *sfb.getnlm(a, idx)) == idx         # test/test_modes.jl:31-37
        assert sfb.getnlm(a, idx) == om.getnlm(oa, idx)
    assert sfb.getnlmsize(a) == om.getnlmsize(oa)


def test_gnl_and_rsdrgnlr_match_oracle(sfb):
    from oracle import windows as ow
    a, oa = sfb.AnlmModes(0.03, 500.0, 1000.0), om.AnlmModes(0.03, 500.0, 1000.0)
    wm, owm = sfb.ConfigurationSpaceModes(a, 300), ow.ConfigurationSpaceModes(500.0, 1000.0, 300, oa.nside)
    assert np.array_equal(wm.r, owm.r) and wm.dr == owm.dr and wm.npix == owm.npix
    G, oG = sfb.rsdrgnlr(a, wm), ow.rsdrgnlr(oa, owm)
    assert G.shape == oG.shape and G.flags.f_contiguous
    assert np.array_equal(np.isnan(G), np.isnan(oG))
    assert np.allclose(G, oG, rtol=1e-7, atol=1e-12, equal_nan=True)
    with pytest.warns(RuntimeWarning, match="unlikely to converge"):          # src/windows.jl:916-919
        sfb.rsdrgnlr(a, sfb.ConfigurationSpaceModes(a, 10))


def test_binning_weights_and_binned_modes(sfb):
    a, oa = sfb.AnlmModes(0.03, 500.0, 1000.0), om.AnlmModes(0.03, 500.0, 1000.0)
    c, oc = sfb.ClnnModes(a), om.ClnnModes(oa)
    for kw in (dict(dl=1), dict(dl=4), dict(dl=3, dn1=2, dn2=2)):
        wt, v = sfb.bandpower_binning_weights(c, **kw)
        owt, ov = om.bandpower_binning_weights(oc, **kw)
        assert np.array_equal(wt.toarray(), owt) and np.abs(v.toarray() - ov).max() < 1e-14
        b, ob = sfb.ClnnBinnedModes(wt, v, c), om.ClnnBinnedModes(owt, ov, oc)
        assert np.allclose(b.LKK, ob.LKK, rtol=1e-9) and sfb.getlnnsize(b) == wt.shape[0]
    # select= (src/modes.jl:732-734,757): only the selected modes get a column, bins in order of first appearance
    rng = np.random.default_rng(5)
    mask = rng.random(c.lnn.shape[1]) < 0.6
    wt, v = sfb.bandpower_binning_weights(c, dl=3, dn1=2, select=mask)
    owt, ov = om.bandpower_binning_weights(oc, dl=3, dn1=2, select=mask)
    assert wt.shape == (owt.shape[0], int(mask.sum())) and np.array_equal(wt.toarray(), owt)
    assert np.abs(v.toarray() - ov).max() < 1e-14
    wt, _ = sfb.bandpower_binning_weights(c)
    assert np.array_equal(wt.toarray(), np.eye(wt.shape[0]))                               # test/test_modes.jl:213-214
    bI = sfb.ClnnBinnedModes(None, None, c)
    assert sfb.getlnnsize(bI) == sfb.getlnnsize(c)


def test_separable_array(sfb):
    # src/SeparableArrays.jl semantics used by the path (test/test_separablearrays.jl)
    phi, mask = np.arange(1.0, 4.0), np.arange(1.0, 6.0)
    s = sfb.SeparableArray(phi, mask)
    assert s.shape == (3, 5) and np.array_equal(s.dense(), np.outer(phi, mask))
    assert s[1, 2] == 6.0 and np.array_equal(s[:, 1], phi * 2.0)
    assert s.phi is s.arr1 and np.array_equal(s.mask, mask)
    assert np.allclose(s.mean(axis=1), s.dense().mean(axis=1)) and np.allclose(s.mean(axis=0), s.dense().mean(axis=0))
    w = sfb.SeparableArray(phi, mask.astype(complex), name2="wlm")
    assert np.iscomplexobj(w.wlm)
    with pytest.raises(AttributeError):
        w.mask
    with pytest.raises(ValueError):
        sfb.SeparableArray(np.ones((2, 2)), mask)


def test_layout_permutation(sfb):
    from oracle import windows as ow
    LMAX = 7
    W = np.arange(3 * sfb.getlmsize(LMAX)).reshape(3, -1).astype(complex)
    assert np.array_equal(sfb.optimize_Wr_lm_layout(W, LMAX), ow.optimize_Wr_lm_layout(W, LMAX))


def test_shard_rows_properties(sfb):
    from sfb_b200.device import shard_rows
    rng = np.random.default_rng(0)
    ell = np.repeat(np.arange(40), rng.integers(1, 30, 40))
    cost = rng.random(ell.size) * (50 - ell)
    for world in (1, 2, 3, 8):
        r = shard_rows(cost, ell, world)
        assert len(r) == world and r[0][0] == 0 and r[-1][1] == ell.size
        assert all(a[1] == b[0] for a, b in zip(r, r[1:])) and all(lo <= hi for lo, hi in r)
        for lo, hi in r[:-1]:
            assert hi == ell.size or hi == 0 or ell[hi] != ell[hi - 1]        # cuts only on l-block boundaries
        if world > 1:
            loads = np.array([cost[lo:hi].sum() for lo, hi in r])
            assert loads.max() < 2.0 * cost.sum() / world


def test_c_abi_exports_every_declared_symbol():
    from sfb_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "sfb_b200.h")).read()
    declared = set(re.findall(r"\b(sfb_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name)
    assert lib.sfb_version() == 100
    sass_ok = os.path.exists(_lib.LIB_PATH)
    assert sass_ok


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "sphericalfourierbesseldecompositions.jl_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "oracle/" not in src, f


def test_fails_loudly_without_gpu(sfb):
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    from sfb_b200 import _lib
    with pytest.raises(_lib.SFBError):
        sfb.calc_Wr_lm(np.ones((2, 48)), 4, 2)
    a = sfb.AnlmModes(2, 3, 500.0, 1000.0)
    with pytest.raises(_lib.SFBError):
        sfb.power_win_mix(np.ones((8, 12 * a.nside ** 2)), sfb.ConfigurationSpaceModes(a, 8), sfb.ClnnModes(a))


def test_untrimmed_gnlr_table_is_trimmed_like_the_shim(sfb):
    """the reference's precompute_gnlr returns nr x size(basisfunctions.knl)... (src/windows.jl:551) and for
    kmax-built modes that knl is the UNTRIMMED zero table (SphericalBesselGNLs.jl:313-316), wider than
    amodes.nmax x (amodes.lmax+1).  The C ABI indexes G[r + nr (n + nmax l)], so julia/SFBB200.jl::rsdrgnlr trims it:
    G[:, 1:amodes.nmax, 1:amodes.lmax+1].  Same operation here on a reference-shaped table."""
    from oracle import windows as ow
    kmax, rmin, rmax = 0.03, 500.0, 1000.0
    oa = om.AnlmModes(kmax, rmin, rmax)
    knl_untrimmed = om.calc_knl_zeros_kmax(kmax, rmin, rmax)
    assert knl_untrimmed.shape[0] > oa.nmax and knl_untrimmed.shape[1] > oa.lmax + 1     # the premise of the finding
    owm = ow.ConfigurationSpaceModes(rmin, rmax, 40, oa.nside)
    ref_shaped = np.full((40,) + knl_untrimmed.shape, np.nan)                            # fill(NaN, nr, size(knl)...)
    trimmed = ow.rsdrgnlr(oa, owm)
    ref_shaped[:, :oa.nmax, :oa.lmax + 1] = trimmed
    shim = np.asfortranarray(ref_shaped[:, :oa.nmax, :oa.lmax + 1])                      # what the shim passes
    a = sfb.AnlmModes(kmax, rmin, rmax)
    G = sfb.rsdrgnlr(a, sfb.ConfigurationSpaceModes(a, 40))
    assert shim.shape == G.shape == (40, a.nmax, a.lmax + 1)
    assert np.allclose(shim, G, rtol=1e-7, atol=1e-12, equal_nan=True)
    # the shim source really trims
    src = open(os.path.join(ROOT, "julia", "SFBB200.jl")).read()
    assert "G[:, 1:amodes.nmax, 1:amodes.lmax+1]" in src
