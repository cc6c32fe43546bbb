"""CPU model of the register-chained DMMA GEMMs of cmix_regz_kernel (csrc/cmix_regz.cu).

mma.sync.m8n8k4.f64 fragment layout (csrc/common.cuh): lane = 4 g + t holds A[g][t], B[t][g], C[g][2t], C[g][2t+1].
The kernel computes Z = A1 · W (accumulators = C fragments) and feeds the accumulator registers straight back as the A
operand of T = (Z ⊙ s) · G2ᵀ, contracting "k-step (jt, e)" over the permuted index r' = 8 jt + 2 t + e and reading the
B fragments at the same permuted index.  This test replays that data flow lane by lane in numpy and checks it against
plain matrix products: it documents why no shared-memory round trip (and no shuffle) is needed between the two GEMMs.
"""
import numpy as np


def dmma(c, a, b):
    """One warp-wide m8n8k4: c[lane] (2 values) += Σ_k A[g][k] B[k][n], with A[g][t] = a[lane], B[t][g] = b[lane]."""
    A = np.zeros((8, 4))
    B = np.zeros((4, 8))
    for lane in range(32):
        g, t = lane >> 2, lane & 3
        A[g, t] = a[lane]
        B[t, g] = b[lane]
    C = A @ B
    for lane in range(32):
        g, t = lane >> 2, lane & 3
        c[lane, 0] += C[g, 2 * t]
        c[lane, 1] += C[g, 2 * t + 1]


def test_register_chained_gemms_match_matmul():
    rng = np.random.default_rng(0)
    AT, NT = 2, 4                      # 16 rows (n), 32 radial points
    AP, K = 8 * AT, 8 * NT
    Gl = rng.standard_normal((AP, K))  # G_ln[r]
    gN = rng.standard_normal(K)        # G_LN[r]
    gN2 = rng.standard_normal(K)       # G_LN'[r']
    W = rng.standard_normal((K, K))
    W = W + W.T                        # Ŵ_lL is symmetric (auto-correlation)
    lanes = np.arange(32)
    g, t = lanes >> 2, lanes & 3

    # Z phase: Z[i][jt][e] at lane (g,t) = Z_N[8i+g][8jt+2t+e]; k-step (kt,e) contracts r = 8kt+2t+e
    Z = np.zeros((AT, NT, 32, 2))
    for kt in range(NT):
        for e in range(2):
            r = 8 * kt + 2 * t + e                     # per-lane contraction index of this k-step
            for i in range(AT):
                a = Gl[8 * i + g, r] * gN[r]           # A[g][t]
                for jt in range(NT):
                    b = W[8 * jt + g, r]               # B[t][g] = Ŵ[r][r'=8jt+g] read through the symmetry
                    dmma(Z[i, jt], a, b)
    Zref = (Gl * gN) @ W                               # [n][r']
    for i in range(AT):
        for jt in range(NT):
            for e in range(2):
                assert np.allclose(Z[i, jt, :, e], Zref[8 * i + g, 8 * jt + 2 * t + e], rtol=1e-12, atol=1e-12)

    # T phase: the accumulator registers are the A operand; k-step (jt,e) contracts r' = 8jt+2t+e
    T = np.zeros((AT, AT, 32, 2))
    for jt in range(NT):
        for e in range(2):
            rp = 8 * jt + 2 * t + e
            for i in range(AT):
                a = Z[i, jt, :, e] * gN2[rp]           # straight from the Z accumulators of this lane
                for j in range(AT):
                    b = Gl[8 * j + g, rp]              # B[t][g] = G_ln'[r'], n' = 8j+g, same permuted index
                    dmma(T[i, j], a, b)
    Tref = (Zref * gN2) @ Gl.T                         # [n][n']
    for i in range(AT):
        for j in range(AT):
            for e in range(2):
                assert np.allclose(T[i, j, :, e], Tref[8 * i + g, 8 * j + 2 * t + e], rtol=1e-11, atol=1e-11)
