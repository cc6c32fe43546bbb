"""CPU model of the virtual row blocks that let the tiled block kernels handle nmax_l > 32 (csrc/cmix.cu,
cmix_plan_create): an l-block with a > 32 radial modes is run as one block per pair (p <= q) of 16-wide panels of its n
range, whose basis is the two panels concatenated (<= 32 functions) and whose rows are the (n in p, n' in q) modes.  The
model checks that (1) every row of the l-block lands in exactly one virtual block with virtual indices below 32 that point
back at the right basis functions, and (2) evaluating the coupling-matrix block through the virtual blocks reproduces the
direct evaluation M[(n,n'),(N,N')] = c_L (T_NN'[n][n'] + [N != N'] T_NN'[n'][n]) of src/windows.jl:613-627,727-736."""
import numpy as np
import pytest

PANEL = 16


def virtual_blocks(a, rows):
    """rows: list of (n, n') 0-based.  Returns [(basis indices, [(row index, n_virtual, n'_virtual), ...]), ...]."""
    out = []
    npan = -(-a // PANEL)
    for pp in range(npan):
        for qq in range(pp, npan):
            p0, p1 = pp * PANEL, min(a, pp * PANEL + PANEL)
            q0, q1 = qq * PANEL, min(a, qq * PANEL + PANEL)
            basis = list(range(p0, p1)) + ([] if pp == qq else list(range(q0, q1)))
            vrows = []
            for i, (n1, n2) in enumerate(rows):
                n1p, n2p = p0 <= n1 < p1, p0 <= n2 < p1
                n1q, n2q = q0 <= n1 < q1, q0 <= n2 < q1
                if pp == qq:
                    if n1p and n2p:
                        vrows.append((i, n1 - p0, n2 - p0))
                elif n1p and n2q:
                    vrows.append((i, n1 - p0, PANEL + n2 - q0))
                elif n1q and n2p:
                    vrows.append((i, PANEL + n1 - q0, n2 - p0))
            if vrows:
                out.append((basis, vrows))
    return out


def reference_order_rows(a):
    """(n, n') of an l-block in the reference's order: by n' - n, then n (src/modes.jl sort_lnn)."""
    return [(n, n + d) for d in range(a) for n in range(a - d)]


@pytest.mark.parametrize("a", [33, 34, 40, 48, 64, 70])
def test_every_row_in_exactly_one_virtual_block(a):
    rows = reference_order_rows(a)
    blocks = virtual_blocks(a, rows)
    seen = np.zeros(len(rows), dtype=int)
    for basis, vrows in blocks:
        assert len(basis) <= 32
        for i, v1, v2 in vrows:
            assert 0 <= v1 < len(basis) and 0 <= v2 < len(basis) and v1 <= v2
            assert (basis[v1], basis[v2]) == rows[i]
            seen[i] += 1
    assert (seen == 1).all()
    npan = -(-a // PANEL)
    assert len(blocks) == npan * (npan + 1) // 2


@pytest.mark.parametrize("a,b,nr", [(37, 5, 12), (50, 34, 9)])
def test_block_through_virtual_blocks_equals_direct(a, b, nr):
    rng = np.random.default_rng(a * 100 + b)
    Gl, GL = rng.standard_normal((a, nr)), rng.standard_normal((b, nr))
    W = rng.standard_normal((nr, nr))
    W = W + W.T                                    # Ŵ_lL of an auto-correlation is symmetric
    rows = reference_order_rows(a)
    cols = [(N, N2) for N in range(b) for N2 in range(N, b)]
    cL = 0.37

    def block(Gbasis, vrows):
        out = {}
        for (N, N2) in cols:
            Z = (Gbasis * GL[N]) @ W                                   # Z_N[n][r']
            T = Z @ (Gbasis * GL[N2]).T                                # T_NN'[n][n']
            for i, v1, v2 in vrows:
                out[(i, N, N2)] = cL * (T[v1, v2] + (T[v2, v1] if N != N2 else 0.0))
        return out

    direct = block(Gl, [(i, n, n2) for i, (n, n2) in enumerate(rows)])
    via = {}
    for basis, vrows in virtual_blocks(a, rows):
        via.update(block(Gl[basis], vrows))
    assert via.keys() == direct.keys()
    err = max(abs(via[k] - direct[k]) for k in direct) / max(abs(v) for v in direct.values())
    assert err < 1e-13
