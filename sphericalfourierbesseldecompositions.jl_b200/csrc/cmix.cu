// Stage 2+3 kernels: see cmix.cuh for the reference functions each one replaces.
#include "cmix.cuh"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

namespace sfb {

// =============================================================================================
// (ℓ L L1; 0 0 0)² for L1 = |ℓ-L| + 2k, k = 0..min(ℓ,L), by the two-term ratio recursion in L1
// (derived from the closed form the reference evaluates with loggamma, src/windows.jl:421-431):
//   g=(ℓ+L+L1)/2, a=g-ℓ, b=g-L, c=g-L1:
//   w²(L1+2)/w²(L1) = (2a+1)(2b+1)(g+1)c / [(2c-1)(2g+3)(a+1)(b+1)]
// One warp per (ℓ,L): ratios are independent per lane, the running product is a warp-shuffle scan,
// and the normalisation Σ_{L1} (2L1+1) w² = 1 (3j orthogonality) is a warp-shuffle reduction.
__global__ void w3j000sq_table_kernel(double* __restrict__ w2, int lmax) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const int np = (lmax + 1) * (lmax + 1);
    if (warp >= np) return;
    const int l = warp / (lmax + 1), L = warp % (lmax + 1);
    const int lo = min(l, L), hi = max(l, L);
    const int nk = lo + 1;
    double* out = w2 + (size_t)warp * (lmax + 1);
    double carry = 1.0, sum = 0.0;
    for (int base = 0; base < nk; base += 32) {
        const int k = base + lane;
        double r = 1.0;
        if (k >= 1 && k < nk) {
            const int L1 = hi - lo + 2 * (k - 1);
            const double g = 0.5 * (l + L + L1), a = g - l, b = g - L, c = g - L1;
            r = ((2 * a + 1) * (2 * b + 1) * (g + 1) * c) / ((2 * c - 1) * (2 * g + 3) * (a + 1) * (b + 1));
        }
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const double t = __shfl_up_sync(0xffffffffu, r, off);
            if (lane >= off) r *= t;
        }
        const double w = carry * r;
        carry = __shfl_sync(0xffffffffu, w, 31);
        if (k < nk) {
            out[k] = w;
            sum += (2.0 * (hi - lo + 2 * k) + 1.0) * w;
        }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
    const double inv = 1.0 / sum;
    __syncwarp();
    for (int k = lane; k < lmax + 1; k += 32) out[k] = (k < nk) ? out[k] * inv : 0.0;
}

// =============================================================================================
// W[L1][i][j] = Σ_{M>=0} (2-δ_M0) Re(W1[i,L1M] conj W2[j,L1M])      (src/windows.jl:682-696)
__global__ void __launch_bounds__(256) wl_build_kernel(const double* __restrict__ alm1, const double* __restrict__ alm2,
                                                       double* __restrict__ W, int LMAX, int nrp) {
    // each thread owns a 4 x 4 block of (i, j): 16 loads feed 32 FMAs per M; the M loop is unrolled so that several
    // M's loads are in flight (the longest CTA, L1 = LMAX, is the critical path: heaviest L1 first)
    const int L1 = LMAX - blockIdx.x;
    const int nb = nrp / 4;  // nrp is a multiple of 8
    for (int blk = threadIdx.x; blk < nb * nb; blk += blockDim.x) {
        const int i0 = (blk / nb) * 4, j0 = (blk % nb) * 4;
        double s[4][4];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) s[a][b] = 0.0;
#pragma unroll 4
        for (int M = 0; M <= L1; ++M) {
            const size_t lm = (size_t)L1 + ((size_t)M * (2 * LMAX + 1 - M)) / 2;
            const double* a1 = alm1 + lm * 2 * nrp;
            const double* a2 = alm2 + lm * 2 * nrp;
            const double c = (M == 0) ? 1.0 : 2.0;
            double r1[4], m1[4], r2[4], m2[4];
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                r1[a] = c * a1[i0 + a];
                m1[a] = c * a1[nrp + i0 + a];
                r2[a] = a2[j0 + a];
                m2[a] = a2[nrp + j0 + a];
            }
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) s[a][b] = fma(r1[a], r2[b], fma(m1[a], m2[b], s[a][b]));
        }
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) W[(size_t)L1 * nrp * nrp + (size_t)(i0 + a) * nrp + j0 + b] = s[a][b];
    }
}

// The same contraction on the FP64 tensor cores: for one L1 it is the rank-2(L1+1) product
//   W_{L1} = A_{L1}ᵀ B_{L1},  A[k][i] = c_M (Re|Im) W1[i, L1 M],  B[k][j] = (Re|Im) W2[j, L1 M],  k = (M, re/im),
// whose operand rows are the contiguous shell vectors of the planar alm.  CTA = (L1, 64 x 64 tile of (i, j)), 8 warps as
// 4 x 2, k-chunks of 16 M's (32 rows) staged in shared memory with the next chunk's loads in flight in registers.
__global__ void __launch_bounds__(256) wl_build_dmma_kernel(const double* __restrict__ alm1, const double* __restrict__ alm2,
                                                            double* __restrict__ W, int LMAX, int nrp) {
    constexpr int LD = 72;   // 64 + 8: At/B fragment loads (t * LD + g) hit distinct banks
    __shared__ double As[32 * LD];
    __shared__ double Bs[32 * LD];
    const int L1 = LMAX - blockIdx.x;                 // heaviest L1 first
    const int i0 = blockIdx.y * 64, j0 = blockIdx.z * 64;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int wm = warp >> 1, wn = warp & 1;
    double acc[2][4][2];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    const int nk = 2 * (L1 + 1);                      // k rows: (M, comp)
    double ra[8], rb[8];
    auto prefetch = [&](int k0) {
        const int c = tid & 63;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int k = k0 + (tid >> 6) + 4 * q;
            ra[q] = rb[q] = 0.0;
            if (k < nk) {
                const int M = k >> 1, comp = k & 1;
                const size_t row = ((size_t)L1 + ((size_t)M * (2 * LMAX + 1 - M)) / 2) * 2 + comp;
                const double cm = (M == 0) ? 1.0 : 2.0;
                if (i0 + c < nrp) ra[q] = cm * alm1[row * nrp + i0 + c];
                if (j0 + c < nrp) rb[q] = alm2[row * nrp + j0 + c];
            }
        }
    };
    prefetch(0);
    for (int k0 = 0; k0 < nk; k0 += 32) {
        {
            const int c = tid & 63;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int kk = (tid >> 6) + 4 * q;
                As[kk * LD + c] = ra[q];
                Bs[kk * LD + c] = rb[q];
            }
        }
        __syncthreads();
        if (k0 + 32 < nk) prefetch(k0 + 32);
        warp_gemm_ts<2, 4>(acc, As + wm * 16, LD, Bs + wn * 32, LD, 32);
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int row = i0 + wm * 16 + i * 8 + g;
        if (row >= nrp) continue;
        double* dst = W + (size_t)L1 * nrp * nrp + (size_t)row * nrp + j0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int col = wn * 32 + j * 8 + 2 * t;
            if (j0 + col < nrp) *reinterpret_cast<double2*>(dst + col) = make_double2(acc[i][j][0], acc[i][j][1]);
        }
    }
}

// =============================================================================================
// Ŵ_{ℓL}[r][r'] = Σ_{L1} (ℓ L L1;000)² W_{L1}[r][r']   — the L1 loop of src/windows.jl:619-622 hoisted out of
// the per-element work (the quadratic form is linear in W).  One CTA = one ℓ and four L of equal parity, so
// every W_{L1} element loaded from L2 feeds four accumulators.
constexpr int kWhatGroup = 8;  // L values (same parity) per CTA: every W_{L1} element loaded from L2 feeds 8 accumulators

__global__ void __launch_bounds__(256) what_build_kernel(const double* __restrict__ W, const double* __restrict__ w2,
                                                         double* __restrict__ What, const int* __restrict__ ells,
                                                         int ell0, int lmax, int nrp, int Llo, int Lhi, int mirror) {
    extern __shared__ double wsm[];  // [kWhatGroup][lmax+1]
    const int ell = ells[blockIdx.y];
    const int gi = blockIdx.x, p = gi & 1, base = (gi >> 1) * (2 * kWhatGroup) + p;
    if (base > Lhi || base + 2 * (kWhatGroup - 1) < Llo) return;
    if (mirror && base + 2 * (kWhatGroup - 1) < ell) return;  // only L >= l blocks are formed  // no L of this group is needed by the column shard
    const int par = (ell + p) & 1;
    const int KW = lmax + 1;
    for (int x = threadIdx.x; x < kWhatGroup * KW; x += blockDim.x) wsm[x] = 0.0;
    __syncthreads();
    int L1lo = 1 << 30, L1hi = -1;
    for (int q = 0; q < kWhatGroup; ++q) {
        const int L = base + 2 * q;
        if (L > lmax) break;
        const int lo = min(ell, L), d = abs(ell - L);
        L1lo = min(L1lo, d);
        L1hi = max(L1hi, ell + L);
        const double* src = w2 + ((size_t)ell * (lmax + 1) + L) * (lmax + 1);
        for (int k = threadIdx.x; k <= lo; k += blockDim.x) wsm[q * KW + ((d + 2 * k - par) >> 1)] = src[k];
    }
    __syncthreads();
    if (L1hi < 0) return;
    const int n2 = nrp * nrp;
    // two elements per thread and four L1 terms per iteration: 8 independent L2 loads in flight per thread
    for (int e = threadIdx.x; e < n2; e += 2 * blockDim.x) {
        const int e2 = e + blockDim.x;
        const bool has2 = e2 < n2;
        double acc[kWhatGroup], acc2[kWhatGroup];
#pragma unroll
        for (int q = 0; q < kWhatGroup; ++q) acc[q] = acc2[q] = 0.0;
        int L1 = L1lo;
        for (; L1 + 6 <= L1hi; L1 += 8) {
            double x[4], y[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                x[u] = __ldg(W + (size_t)(L1 + 2 * u) * n2 + e);
                y[u] = has2 ? __ldg(W + (size_t)(L1 + 2 * u) * n2 + e2) : 0.0;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int h = (L1 + 2 * u - par) >> 1;
#pragma unroll
                for (int q = 0; q < kWhatGroup; ++q) {
                    const double wv = wsm[q * KW + h];
                    acc[q] = fma(wv, x[u], acc[q]);
                    acc2[q] = fma(wv, y[u], acc2[q]);
                }
            }
        }
        for (; L1 <= L1hi; L1 += 2) {
            const double x = __ldg(W + (size_t)L1 * n2 + e);
            const double y = has2 ? __ldg(W + (size_t)L1 * n2 + e2) : 0.0;
            const int h = (L1 - par) >> 1;
#pragma unroll
            for (int q = 0; q < kWhatGroup; ++q) {
                acc[q] = fma(wsm[q * KW + h], x, acc[q]);
                acc2[q] = fma(wsm[q * KW + h], y, acc2[q]);
            }
        }
#pragma unroll
        for (int q = 0; q < kWhatGroup; ++q) {
            const int L = base + 2 * q;
            if (L <= lmax) {
                double* dst = What + ((size_t)(ell - ell0) * (lmax + 1) + L) * n2;
                dst[e] = acc[q];
                if (has2) dst[e2] = acc2[q];
            }
        }
    }
}

// =============================================================================================
// Ŵ on the FP64 tensor cores: for one ℓ the L1 sum is the banded GEMM  Ŵ_ℓ[L][e] = Σ_{L1} A_ℓ[L][L1] · W[L1][e],
// A_ℓ[L][L1] = (ℓ L L1;000)² (zero outside the triangle and for the wrong parity), e = (r, r') flattened.
// CTA = (ℓ, 32 L's of one parity = 4 DMMA row tiles, 1024 e's: 4 warps x 4 tiles of 64, three CTAs per SM); every W element fetched from L2 feeds 32 L's (the FMA
// kernel above: 8), and the contraction runs over the union of the 32 triangles only, in steps of four L1 of the right
// parity.  A lives in shared memory (row stride ≡ 4 mod 16: conflict-free fragment loads), B fragments come straight
// from L2 with the next k-step's eight loads in flight.
constexpr int kWhatRows = 32;   // L values (same parity) per CTA
constexpr int kWhatEPW = 4;     // 64-wide e tiles per warp
__global__ void __launch_bounds__(128, 3) what_build_dmma_kernel(const double* __restrict__ W, const double* __restrict__ w2,
                                                              double* __restrict__ What, const int* __restrict__ ells,
                                                              int ell0, int lmax, int nrp, int Llo, int Lhi, int mirror,
                                                              int SW) {
    extern __shared__ double wsm[];  // [kWhatRows][SW], column h <-> L1 = par + 2h
    const int ell = ells[blockIdx.y];
    const int gi = blockIdx.x, p = gi & 1, base = (gi >> 1) * (2 * kWhatRows) + p;
    const int last = base + 2 * (kWhatRows - 1);
    if (base > Lhi || last < Llo || base > lmax) return;
    if (mirror && last < ell) return;  // only L >= l blocks are formed
    const int par = (ell + p) & 1;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    for (int x = tid; x < kWhatRows * SW; x += blockDim.x) wsm[x] = 0.0;
    __syncthreads();
    int L1lo = 1 << 30, L1hi = -1;
    for (int q = 0; q < kWhatRows; ++q) {
        const int L = base + 2 * q;
        if (L > lmax) break;
        if ((mirror && L < ell) || L < Llo || L > Lhi) continue;   // rows nobody reads stay zero
        const int lo = min(ell, L), d = abs(ell - L);
        L1lo = min(L1lo, d);
        L1hi = max(L1hi, ell + L);
        const double* src = w2 + ((size_t)ell * (lmax + 1) + L) * (lmax + 1);
        for (int k = tid; k <= lo; k += blockDim.x) wsm[q * SW + ((d + 2 * k - par) >> 1)] = src[k];
    }
    __syncthreads();
    if (L1hi < 0) return;
    const int n2 = nrp * nrp;
    const int hlo = (L1lo - par) >> 1, hhi = (L1hi - par) >> 1;
    const int nks = (hhi - hlo + 4) / 4;   // k-steps of four L1 values (h, h+1, h+2, h+3)
    // row tiles that hold at least one wanted L, and the k-steps that meet their band: the eight rows of a tile span
    // L1 in [min |l-L|, max l+L], l + 8 values of the right parity, against l + 32 for the union over the CTA's rows
    bool live[4];
    int kslo[4], kshi[4];
#pragma unroll
    for (int mt = 0; mt < 4; ++mt) {
        const int Lf = base + 2 * (mt * 8), Ll = min(base + 2 * (mt * 8 + 7), lmax - ((lmax - base) & 1));
        live[mt] = (Lf <= lmax) && !(mirror && Ll < ell) && !(Ll < Llo || Lf > Lhi);
        const int dmin = (ell >= Lf && ell <= Ll) ? ((ell - Lf) & 1) : min(abs(ell - Lf), abs(ell - Ll));
        const int h0 = (dmin - par) >> 1, h1 = (ell + Ll - par) >> 1;      // wsm columns of the tile's band
        kslo[mt] = max(0, (h0 - hlo) / 4);
        kshi[mt] = min(nks - 1, (h1 - hlo) / 4);
    }
    // each warp sweeps kWhatEPW column tiles of 64 e's: the 3j tile in shared memory is built once per 1024 e's
    for (int ep = 0; ep < kWhatEPW; ++ep) {
        const int e0 = ((blockIdx.z * 4 + warp) * kWhatEPW + ep) * 64;
        if (e0 >= n2) break;
        double acc[4][8][2];
#pragma unroll
        for (int mt = 0; mt < 4; ++mt)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[mt][j][0] = acc[mt][j][1] = 0.0;
        const double* Wb = W + e0 + g;
        auto loadB = [&](int ks, double (&b)[8]) {
            const int L1 = par + 2 * (hlo + 4 * ks + t);
            const bool ok = (L1 <= 2 * lmax);
            const double* src = Wb + (size_t)(ok ? L1 : 0) * n2;
#pragma unroll
            for (int j = 0; j < 8; ++j) b[j] = ok ? __ldg(src + 8 * j) : 0.0;
        };
        double bcur[8], bnxt[8];
        loadB(0, bcur);
        for (int ks = 0; ks < nks; ++ks) {
            if (ks + 1 < nks) loadB(ks + 1, bnxt);
            const double* arow = wsm + g * SW + hlo + 4 * ks + t;
#pragma unroll
            for (int mt = 0; mt < 4; ++mt) {
                if (!live[mt] || ks < kslo[mt] || ks > kshi[mt]) continue;
                const double a = arow[mt * 8 * SW];
#pragma unroll
                for (int j = 0; j < 8; ++j) dmma884(acc[mt][j], a, bcur[j]);
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) bcur[j] = bnxt[j];
        }
#pragma unroll
        for (int mt = 0; mt < 4; ++mt) {
            if (!live[mt]) continue;
            const int L = base + 2 * (mt * 8 + g);
            if (L > lmax) continue;
            double* dst = What + ((size_t)(ell - ell0) * (lmax + 1) + L) * n2 + e0 + 2 * t;
#pragma unroll
            for (int j = 0; j < 8; ++j)
                *reinterpret_cast<double2*>(dst + 8 * j) = make_double2(acc[mt][j][0], acc[mt][j][1]);
        }
    }
}

// =============================================================================================
// Coupling-matrix block kernel.  One CTA = (ℓ, L, chunk of N values [N0,N1)), 8 warps:
//   Z_N[n][r']   = Σ_r  G_ℓn[r] G_LN[r] Ŵ_ℓL[r][r']   for every N of the chunk     (DMMA, kept in shared memory)
//   T_N,N'[n][n'] = Σ_r' Z_N[n][r'] G_LN'[r'] G_ℓn'[r']  for N' >= N               (DMMA, a × a × nr tiles; the
//                   (N,N') tiles of the chunk are dealt round-robin to the warps, two N' per pass share loads)
//   M[(ℓ,n,n'),(L,N,N')] = c_L · ( T[n][n'] + [N≠N'] T[n'][n] )                     (src/windows.jl:727-736)
// With win1 ≢ win2 (SYM=false) Ŵ is not symmetric and the N<->N' partner uses Zt_N = (G_ℓ ⊙ G_LN)ᵀ Ŵᵀ instead.
struct CmixArgs {
    const double* G;        // [ell][nmax][nrp]
    const double* What;     // [ell - ell0][L][nrp][nrp]
    const int* ell_list;    // ells handled by this launch (blockIdx.y)
    const int* ell_ptr;     // CSR rows per ell
    const int* row_out;     // output row (relative to the shard) or -1
    const int* row_n;       // 0-based n
    const int* row_n2;      // 0-based n'
    const int* a_of_ell;
    const int* real_ell;    // row block -> l (virtual row blocks of an l with nmax_l > 32)
    const int* pairidx;     // [L][nmax][nmax] -> output column or -1
    const int* ch_L;        // blockIdx.x -> (L, N0, N1)
    const int* ch_N0;
    const int* ch_N1;
    double* M[8];           // output matrix on this device ([0]) and, for the fused all-gather, on every peer
    int npeers;
    long long ldM;
    int ell0, lmax, nmax, nrp, S, NC;
    int col_lo, col_hi;     // output columns [col_lo, col_hi) are written, relative to col_lo
    int div2Lp1, interchange;
    int mirror;             // 1: only blocks with L >= l are formed; each tile also fills M[(L,N,N'),(l,n,n')]
    int gl_global;          // 1: G_L rows are read from global memory (L1/L2) instead of being staged: long radial grids
    int dbg;                // profiling aid (SFB_CMIX_DBG): 1 skip epilogue stores, 2 skip Z phase, 4 skip T-phase DMMA
};

constexpr int kCmixThreads = 128;   // 4 warps; two CTAs per SM so that one CTA's prologue/tail overlaps the other's DMMA phase
constexpr int kCmixWarps = kCmixThreads / 32;

__host__ __device__ constexpr int cmix_tld(int AP) { return (AP % 16 == 8) ? AP : AP + 8; }  // ≡ 8 (mod 16)

template <int AT, bool SYM>
__global__ void __launch_bounds__(kCmixThreads, 2) cmix_block_kernel(CmixArgs p) {
    extern __shared__ double sm[];
    const int ell = p.ell_list[blockIdx.y];       // row block (G rows, row table); rl = its l (Ŵ index, factors)
    const int rl = p.real_ell[ell];
    const int L = p.ch_L[blockIdx.x], N0 = p.ch_N0[blockIdx.x], N1 = p.ch_N1[blockIdx.x];
    const int a = p.a_of_ell[ell], b = p.a_of_ell[L];
    if (a == 0 || N0 >= b) return;
    if (p.mirror && L < rl) return;  // obtained from block (L, l) by the symmetry of the un-symmetrised kernel
    constexpr int AP = AT * 8;
    constexpr int TLD = cmix_tld(AP);       // staging leading dimension: double2 stores are conflict free
    constexpr int P = SYM ? 2 : 1;          // N' tiles per warp pass (shares the Z and G_l fragment loads)
    constexpr int NZ = SYM ? 1 : 2;
    const int S = p.S, nrp = p.nrp, nN = N1 - N0;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int r0 = p.ell_ptr[ell], nrows = p.ell_ptr[ell + 1] - r0;

    double* Gl = sm;                            // [AP][S]        G_ln[r]
    double* GLs = Gl + AP * S;                  // [nmax][S]      G_LN[r]  (absent with gl_global)
    const bool glg = p.gl_global != 0;
    const double* GL = glg ? p.G + (size_t)L * p.nmax * nrp : GLs;
    const int gS = glg ? nrp : S;               // row stride of GL
    double* Zs = GLs + (glg ? 0 : p.nmax * S);  // [NC][AP][S]
    double* Zt = Zs + (size_t)p.NC * AP * S;    // [NC][AP][S] (only if !SYM)
    double* Ts = Zs + (size_t)NZ * p.NC * AP * S;  // [warps][NZ][AP][TLD]
    int* coltab = reinterpret_cast<int*>(Ts + kCmixWarps * NZ * AP * TLD);  // [NC][nmax] output column of (N, N') or -1
    int* rowtab = coltab + p.NC * p.nmax;                                   // [nrows] orow, [nrows] n | n'<<16

    // ---- stage G_l, G_L rows (zero padded) and this l-block's row table ------------------------------
    for (int n = warp; n < AP; n += kCmixWarps) {
        const double* src = p.G + ((size_t)ell * p.nmax + n) * nrp;
        for (int r = lane; r < S; r += 32) Gl[n * S + r] = (n < a && r < nrp) ? src[r] : 0.0;
    }
    for (int n = warp; n < b && !glg; n += kCmixWarps) {
        const double* src = p.G + ((size_t)L * p.nmax + n) * nrp;
        for (int r = lane; r < S; r += 32) GLs[n * S + r] = (r < nrp) ? src[r] : 0.0;
    }
    for (int x = tid; x < nrows; x += kCmixThreads) {
        rowtab[x] = p.row_out[r0 + x];
        rowtab[nrows + x] = p.row_n[r0 + x] | (p.row_n2[r0 + x] << 16);
    }
    for (int x = tid; x < nN * p.nmax; x += kCmixThreads) {
        const int Nloc = x / p.nmax, N2 = x - Nloc * p.nmax;
        int c = (N2 < b) ? p.pairidx[((size_t)L * p.nmax + N0 + Nloc) * p.nmax + N2] : -1;
        coltab[x] = (c >= p.col_lo && c < p.col_hi) ? c - p.col_lo : -1;
    }
    __syncthreads();

    // ---- Z phase: each warp owns a pair of 8-wide r' tiles (B fragments of Ŵ) and sweeps the N of the chunk --
    const double* Wh = p.What + ((size_t)(rl - p.ell0) * (p.lmax + 1) + L) * nrp * nrp;
    const int ntile = nrp / 8, njp = (ntile + 1) / 2;
    const int wgrp = min(njp, kCmixWarps), nlane = kCmixWarps / wgrp;
    const int wj = warp % wgrp, wn = warp / wgrp;
    if (wn < nlane && !(p.dbg & 2)) {
        for (int jtp = wj; jtp < njp; jtp += wgrp) {
            const int jt = 2 * jtp;
            const bool two = (jt + 1 < ntile);
            if (SYM && nrp <= 64) {
                // fast path: all B fragments of this tile pair live in registers, so the L2 latency of Ŵ is paid
                // once per warp instead of once per k-step
                double bw[16][2];
#pragma unroll
                for (int ks = 0; ks < 16; ++ks) {
                    bw[ks][0] = bw[ks][1] = 0.0;
                    if (4 * ks < nrp) {
                        const double* wrow = Wh + (size_t)(4 * ks + t) * nrp + jt * 8 + g;
                        bw[ks][0] = __ldg(wrow);
                        if (two) bw[ks][1] = __ldg(wrow + 8);
                    }
                }
                for (int Nloc = wn; Nloc < nN; Nloc += nlane) {
                    const double* glN = GL + (N0 + Nloc) * gS;
                    double acc[AT][2][2];
#pragma unroll
                    for (int i = 0; i < AT; ++i) acc[i][0][0] = acc[i][0][1] = acc[i][1][0] = acc[i][1][1] = 0.0;
#pragma unroll
                    for (int ks = 0; ks < 16; ++ks) {
                        if (4 * ks < nrp) {
                            const double sN = glN[4 * ks + t];
#pragma unroll
                            for (int i = 0; i < AT; ++i) {
                                const double av = Gl[(i * 8 + g) * S + 4 * ks + t] * sN;
                                dmma884(acc[i][0], av, bw[ks][0]);
                                dmma884(acc[i][1], av, bw[ks][1]);
                            }
                        }
                    }
                    double* zdst = Zs + (size_t)Nloc * AP * S;
#pragma unroll
                    for (int i = 0; i < AT; ++i)
#pragma unroll
                        for (int j = 0; j < 2; ++j) {
                            if (j == 1 && !two) continue;
                            const int off = (i * 8 + g) * S + (jt + j) * 8 + 2 * t;
                            *reinterpret_cast<double2*>(zdst + off) = make_double2(acc[i][j][0], acc[i][j][1]);
                        }
                }
            } else {
                for (int Nloc = wn; Nloc < nN; Nloc += nlane) {
                    const double* glN = GL + (N0 + Nloc) * gS;
                    double acc[AT][2][2], acct[AT][2][2];
#pragma unroll
                    for (int i = 0; i < AT; ++i)
#pragma unroll
                        for (int j = 0; j < 2; ++j) acc[i][j][0] = acc[i][j][1] = acct[i][j][0] = acct[i][j][1] = 0.0;
#pragma unroll 2
                    for (int k0 = 0; k0 < nrp; k0 += 4) {
                        const double sN = glN[k0 + t];
                        const double* wrow = Wh + (size_t)(k0 + t) * nrp + jt * 8 + g;
                        const double bw0 = __ldg(wrow), bw1 = two ? __ldg(wrow + 8) : 0.0;
                        double bt0 = 0.0, bt1 = 0.0;
                        if (!SYM) {
                            bt0 = __ldg(Wh + (size_t)(jt * 8 + g) * nrp + k0 + t);
                            bt1 = two ? __ldg(Wh + (size_t)(jt * 8 + 8 + g) * nrp + k0 + t) : 0.0;
                        }
#pragma unroll
                        for (int i = 0; i < AT; ++i) {
                            const double av = Gl[(i * 8 + g) * S + k0 + t] * sN;
                            dmma884(acc[i][0], av, bw0);
                            dmma884(acc[i][1], av, bw1);
                            if (!SYM) {
                                dmma884(acct[i][0], av, bt0);
                                dmma884(acct[i][1], av, bt1);
                            }
                        }
                    }
                    double* zdst = Zs + (size_t)Nloc * AP * S;
                    double* ztdst = Zt + (size_t)Nloc * AP * S;
#pragma unroll
                    for (int i = 0; i < AT; ++i)
#pragma unroll
                        for (int j = 0; j < 2; ++j) {
                            if (j == 1 && !two) continue;
                            const int off = (i * 8 + g) * S + (jt + j) * 8 + 2 * t;
                            *reinterpret_cast<double2*>(zdst + off) = make_double2(acc[i][j][0], acc[i][j][1]);
                            if (!SYM)
                                *reinterpret_cast<double2*>(ztdst + off) = make_double2(acct[i][j][0], acct[i][j][1]);
                        }
                }
            }
        }
    }
    __syncthreads();

    // ---- T phase: (N, N') tiles of the chunk dealt round-robin, P consecutive N' per pass -----------------
    const double scale = (p.div2Lp1 ? 1.0 : (2.0 * L + 1.0)) * 0.07957747154594767;  // 1/(4π)
    const double scale_l = (p.div2Lp1 ? 1.0 : (2.0 * rl + 1.0)) * 0.07957747154594767;
    const bool domirror = p.mirror && (L > rl);
    double* Tw = Ts + warp * NZ * AP * TLD;
    int cnt = 0;
    for (int Nloc = 0; Nloc < nN; ++Nloc) {
        const int N = N0 + Nloc;
        const int npass = (b - N + P - 1) / P;
        const int* prow = coltab + Nloc * p.nmax;
        const double* Zn = Zs + (size_t)Nloc * AP * S;
        const double* Ztn = Zt + (size_t)Nloc * AP * S;
        for (int pi = (warp - cnt) & (kCmixWarps - 1); pi < npass; pi += kCmixWarps) {
            const int N2 = N + P * pi;
            int col[P];
            bool any = false;
#pragma unroll
            for (int q = 0; q < P; ++q) {
                col[q] = (N2 + q < b) ? prow[N2 + q] : -1;
                any = any || (col[q] >= 0);
            }
            if (!any) continue;  // warp-uniform
            double acc[P][AT][AT][2], acc2[SYM ? 1 : AT][SYM ? 1 : AT][2];
#pragma unroll
            for (int q = 0; q < P; ++q)
#pragma unroll
                for (int i = 0; i < AT; ++i)
#pragma unroll
                    for (int j = 0; j < AT; ++j) acc[q][i][j][0] = acc[q][i][j][1] = 0.0;
            if (!SYM) {
#pragma unroll
                for (int i = 0; i < (SYM ? 1 : AT); ++i)
#pragma unroll
                    for (int j = 0; j < (SYM ? 1 : AT); ++j) acc2[i][j][0] = acc2[i][j][1] = 0.0;
            }
            const double* glA = GL + N2 * gS;
            const double* glB = GL + ((N2 + 1 < b) ? N2 + 1 : N2) * gS;
#pragma unroll 2
            for (int k0 = 0; k0 < ((p.dbg & 4) ? 4 : nrp); k0 += 4) {
                double sc[P], zv[AT], gv[AT];
                sc[0] = glA[k0 + t];
                if (P > 1) sc[P - 1] = glB[k0 + t];
#pragma unroll
                for (int i = 0; i < AT; ++i) zv[i] = Zn[(i * 8 + g) * S + k0 + t];
#pragma unroll
                for (int j = 0; j < AT; ++j) gv[j] = Gl[(j * 8 + g) * S + k0 + t];
#pragma unroll
                for (int q = 0; q < P; ++q)
#pragma unroll
                    for (int i = 0; i < AT; ++i) {
                        const double av = zv[i] * sc[q];
#pragma unroll
                        for (int j = 0; j < AT; ++j) dmma884(acc[q][i][j], av, gv[j]);
                    }
                if (!SYM) {
#pragma unroll
                    for (int i = 0; i < (SYM ? 1 : AT); ++i) {
                        const double av = Ztn[(i * 8 + g) * S + k0 + t] * sc[0];
#pragma unroll
                        for (int j = 0; j < (SYM ? 1 : AT); ++j) dmma884(acc2[i][j], av, gv[j]);
                    }
                }
            }
#pragma unroll
            for (int q = 0; q < P; ++q) {
                if (col[q] < 0 || (p.dbg & 1)) continue;  // warp-uniform
#pragma unroll
                for (int i = 0; i < AT; ++i)
#pragma unroll
                    for (int j = 0; j < AT; ++j) {
                        const int off = (i * 8 + g) * TLD + j * 8 + 2 * t;
                        *reinterpret_cast<double2*>(Tw + off) = make_double2(acc[q][i][j][0], acc[q][i][j][1]);
                        if (!SYM)
                            *reinterpret_cast<double2*>(Tw + AP * TLD + off) =
                                make_double2(acc2[SYM ? 0 : i][SYM ? 0 : j][0], acc2[SYM ? 0 : i][SYM ? 0 : j][1]);
                    }
                __syncwarp();
                const double* T1 = Tw;
                const double* T2 = SYM ? Tw : Tw + AP * TLD;
                const size_t coff = (size_t)col[q] * p.ldM;
                const bool offdiag = (N2 + q != N);
                // four rows per lane per iteration: independent load chains hide the shared-memory latency
                for (int idx0 = lane; idx0 < nrows; idx0 += 128) {
                    int orow[4], code[4];
                    double v[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int idx = idx0 + 32 * u;
                        orow[u] = (idx < nrows) ? rowtab[idx] : -1;
                        code[u] = (idx < nrows) ? rowtab[nrows + idx] : 0;
                    }
                    double vm[4];  // mirrored element M[(L,N,N'),(l,n,n')] = c_l (A + [n≠n'] B), A = T[n][n'], B = T[n'][n]
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int n = code[u] & 0xffff, n2 = code[u] >> 16;
                        const double A = T1[n * TLD + n2], B = T2[n2 * TLD + n];
                        if (p.interchange) {
                            v[u] = B;
                            vm[u] = B;
                        } else {
                            v[u] = offdiag ? A + B : A;
                            vm[u] = (n != n2) ? A + B : A;
                        }
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        if (orow[u] < 0) continue;
                        const double val = v[u] * scale;
                        p.M[0][coff + orow[u]] = val;
                        if (p.npeers > 1) {
#pragma unroll
                            for (int d = 1; d < 8; ++d)
                                if (d < p.npeers) p.M[d][coff + orow[u]] = val;
                        }
                        if (domirror) p.M[0][(size_t)orow[u] * p.ldM + col[q]] = vm[u] * scale_l;
                    }
                }
                __syncwarp();
            }
        }
        cnt += npass;
    }
}

template <int AT, bool SYM>
static size_t cmix_smem_bytes(int S, int nmax, int NC, int max_rows, bool glg) {
    const int AP = AT * 8, NZ = SYM ? 1 : 2;
    return sizeof(double) * ((size_t)AP * S + (glg ? 0 : (size_t)nmax * S) + (size_t)NZ * NC * AP * S +
                             (size_t)kCmixWarps * NZ * AP * cmix_tld(AP)) +
           sizeof(int) * (2 * (size_t)max_rows + (size_t)NC * nmax);
}

template <int AT, bool SYM>
static int launch_cmix(const CmixArgs& args, int nchunks, int nells, int nmax, int max_rows, cudaStream_t stream) {
    const size_t smem = cmix_smem_bytes<AT, SYM>(args.S, nmax, args.NC, max_rows, args.gl_global != 0);
    SFB_REQUIRE(smem <= 227 * 1024, "cmix_block_kernel: shared memory footprint exceeds 227 KB (nr * nmax too large)");
    SFB_CUDA_OK(cudaFuncSetAttribute(cmix_block_kernel<AT, SYM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(nchunks, nells);
    cmix_block_kernel<AT, SYM><<<grid, kCmixThreads, smem, stream>>>(args);
    SFB_CUDA_OK(cudaGetLastError());
    return 0;
}

template <bool SYM>
static int launch_cmix_at(int AT, const CmixArgs& args, int nchunks, int nells, int nmax, int max_rows,
                          cudaStream_t stream) {
    switch (AT) {
        case 1: return launch_cmix<1, SYM>(args, nchunks, nells, nmax, max_rows, stream);
        case 2: return launch_cmix<2, SYM>(args, nchunks, nells, nmax, max_rows, stream);
        case 3: return launch_cmix<3, SYM>(args, nchunks, nells, nmax, max_rows, stream);
        case 4: return launch_cmix<4, SYM>(args, nchunks, nells, nmax, max_rows, stream);
        default: break;
    }
    set_error("cmix: nmax_l > 32 is not supported by this build");
    return 2;
}

// number of N values whose Z fits next to the fixed buffers in ~110 KB of shared memory (two CTAs per SM)
static int cmix_chunk_size_for(size_t budget, int AT, bool sym, int S, int nmax, int max_rows, bool glg = false) {
    const int AP = AT * 8, NZ = sym ? 1 : 2;
    const size_t fixed = sizeof(double) * ((size_t)AP * S + (glg ? 0 : (size_t)nmax * S) +
                                           (size_t)kCmixWarps * NZ * AP * cmix_tld(AP)) +
                         sizeof(int) * 2 * (size_t)max_rows;
    const size_t per = sizeof(double) * (size_t)NZ * AP * S + sizeof(int) * (size_t)nmax;
    if (fixed + per > budget) return 0;
    return (int)std::min<size_t>(nmax, (budget - fixed) / per);
}
static int cmix_chunk_size(int AT, bool sym, int S, int nmax, int max_rows, bool* glg) {
    // prefer two CTAs per SM (~110 KB each); long radial grids need the whole 220 KB for one CTA, and beyond that the
    // G_L rows stay in global memory (e.g. cfg4's modes on the reference's recommended nr = 8(n+N) = 384 grid)
    *glg = false;
    int nc = cmix_chunk_size_for(110 * 1024, AT, sym, S, nmax, max_rows);
    if (nc >= 1) return nc;
    nc = cmix_chunk_size_for(220 * 1024, AT, sym, S, nmax, max_rows);
    if (nc >= 1) return nc;
    *glg = true;
    return cmix_chunk_size_for(220 * 1024, AT, sym, S, nmax, max_rows, true);
}

// =============================================================================================
// plan

int cmix_plan_create(CmixPlan** out, const int64_t* lnn, int64_t lnnsize, int64_t lnn_min, const double* G,
                     int64_t nr, int64_t nmax, int64_t lmax) {
    SFB_REQUIRE(out && lnn && G, "cmix_plan_create: null pointer");
    SFB_REQUIRE(lnnsize >= 1 && nr >= 1 && nmax >= 1 && lmax >= 0, "cmix_plan_create: bad sizes");
    SFB_REQUIRE(lnn_min >= 1 && lnn_min <= lnnsize, "cmix_plan_create: lnn_min out of range");
    SFB_REQUIRE(lnnsize < (int64_t(1) << 31), "cmix_plan_create: lnnsize too large");
    auto* p = new CmixPlan();
    p->lmax = (int)lmax;
    p->nmax = (int)nmax;
    p->LMAX = 2 * (int)lmax;
    p->nr = (int)nr;
    p->nrp = (int)round_up(nr, 8);
    p->S = (int)round_up(p->nrp, 16) + 4;
    p->lnnsize = lnnsize;
    p->lnn_min = lnn_min;
    p->nout = lnnsize - lnn_min + 1;

    // --- index tables (bit-exact consumers of the caller's lnn table) ---
    std::vector<int> a(lmax + 1, 0), cnt(lmax + 2, 0);
    for (int64_t i = lnn_min - 1; i < lnnsize; ++i) {
        const int64_t l = lnn[3 * i], n1 = lnn[3 * i + 1], n2 = lnn[3 * i + 2];
        if (l < 0 || l > lmax || n1 < 1 || n2 < 1 || n1 > nmax || n2 > nmax) {
            delete p;
            set_error("cmix_plan_create: lnn entry out of range");
            return 2;
        }
        a[l] = std::max(a[l], (int)std::max(n1, n2));
        cnt[l + 1]++;
    }
    p->a_of_ell = a;
    p->ell_ptr.assign(lmax + 2, 0);
    for (int l = 0; l <= lmax; ++l) p->ell_ptr[l + 1] = p->ell_ptr[l] + cnt[l + 1];
    const int nrows = p->ell_ptr[lmax + 1];
    p->h_row_out.assign(nrows, -1);
    p->h_row_n.assign(nrows, 0);
    p->h_row_n2.assign(nrows, 0);
    std::vector<int> fill(p->ell_ptr.begin(), p->ell_ptr.end() - 1);
    std::vector<int> pairidx((size_t)(lmax + 1) * nmax * nmax, -1);
    for (int64_t i = lnn_min - 1; i < lnnsize; ++i) {
        const int l = (int)lnn[3 * i], n1 = (int)lnn[3 * i + 1] - 1, n2 = (int)lnn[3 * i + 2] - 1;
        const int slot = fill[l]++;
        p->h_row_out[slot] = (int)(i - (lnn_min - 1));
        p->h_row_n[slot] = n1;
        p->h_row_n2[slot] = n2;
        pairidx[((size_t)l * nmax + n1) * nmax + n2] = (int)(i - (lnn_min - 1));
    }
    // row blocks: the l-blocks, then the virtual blocks of every l with more than 32 radial modes
    p->nblk = (int)lmax + 1;
    p->blk_real.resize(lmax + 1);
    p->blk_of_ell.assign(lmax + 1, std::vector<int>());
    for (int l = 0; l <= lmax; ++l) {
        p->blk_real[l] = l;
        p->blk_of_ell[l].push_back(l);
    }
    struct VPanel { int l, p0, np, q0, nq; };
    std::vector<VPanel> vpanels;
    constexpr int kPanel = 16;
    for (int l = 0; l <= lmax; ++l) {
        if (a[l] <= 32) continue;
        p->blk_of_ell[l].clear();
        const int npan = (a[l] + kPanel - 1) / kPanel;
        for (int pp = 0; pp < npan; ++pp)
            for (int qq = pp; qq < npan; ++qq) {
                const int p0 = pp * kPanel, p1 = std::min(a[l], p0 + kPanel), q0 = qq * kPanel, q1 = std::min(a[l], q0 + kPanel);
                std::vector<int> ro, rn, rn2;
                for (int s2 = p->ell_ptr[l]; s2 < p->ell_ptr[l + 1]; ++s2) {
                    const int n1 = p->h_row_n[s2], n2 = p->h_row_n2[s2];
                    const bool n1p = n1 >= p0 && n1 < p1, n2p = n2 >= p0 && n2 < p1;
                    const bool n1q = n1 >= q0 && n1 < q1, n2q = n2 >= q0 && n2 < q1;
                    int v1 = -1, v2 = -1;
                    if (pp == qq) {
                        if (n1p && n2p) v1 = n1 - p0, v2 = n2 - p0;
                    } else if (n1p && n2q) {
                        v1 = n1 - p0, v2 = kPanel + n2 - q0;
                    } else if (n1q && n2p) {
                        v1 = kPanel + n1 - q0, v2 = n2 - p0;
                    }
                    if (v1 < 0) continue;
                    ro.push_back(p->h_row_out[s2]);
                    rn.push_back(v1);
                    rn2.push_back(v2);
                }
                if (ro.empty()) continue;
                const int v = p->nblk++;
                p->blk_real.push_back(l);
                p->blk_of_ell[l].push_back(v);
                a.push_back(pp == qq ? p1 - p0 : kPanel + (q1 - q0));
                p->ell_ptr.push_back(p->ell_ptr.back() + (int)ro.size());
                p->h_row_out.insert(p->h_row_out.end(), ro.begin(), ro.end());
                p->h_row_n.insert(p->h_row_n.end(), rn.begin(), rn.end());
                p->h_row_n2.insert(p->h_row_n2.end(), rn2.begin(), rn2.end());
                vpanels.push_back({l, p0, p1 - p0, q0, pp == qq ? 0 : q1 - q0});
                (void)v;
            }
    }
    p->a_of_ell = a;
    // per output index: l | [n≠n'] << 30 (mirror fill), and whether l is non-decreasing in output order
    std::vector<int> es((size_t)p->nout, 0);
    p->ell_sorted = true;
    for (int64_t i = lnn_min - 1; i < lnnsize; ++i) {
        es[i - (lnn_min - 1)] = (int)lnn[3 * i] | ((lnn[3 * i + 1] != lnn[3 * i + 2]) ? (1 << 30) : 0);
        if (i > lnn_min - 1 && lnn[3 * i] < lnn[3 * (i - 1)]) p->ell_sorted = false;
    }
    std::vector<int> nl_L, nl_N;
    int amax = 0;
    for (int L = 0; L <= lmax; ++L) {
        amax = std::max(amax, a[L]);
        for (int N = 0; N < a[L]; ++N) {
            bool any = false;
            for (int N2 = N; N2 < a[L] && !any; ++N2) any = pairidx[((size_t)L * nmax + N) * nmax + N2] >= 0;
            if (any) {
                nl_L.push_back(L);
                nl_N.push_back(N);
            }
        }
    }
    p->nl = (int)nl_L.size();
    p->amax_tiles = std::min(4, (amax + 7) / 8);   // row tiles of the launched row blocks (virtual blocks hold <= 32 functions)

    // --- G: [nr][nmax][lmax+1] column-major -> [ell][n][nrp], NaN / padding -> 0 where unused ---
    std::vector<double> Gh((size_t)p->nblk * nmax * p->nrp, 0.0);
    for (int l = 0; l <= lmax; ++l)
        for (int n = 0; n < a[l]; ++n)
            for (int r = 0; r < nr; ++r) {
                const double v = G[(size_t)r + (size_t)nr * (n + (size_t)nmax * l)];
                if (!std::isfinite(v)) {
                    delete p;
                    set_error("cmix_plan_create: non-finite radial basis value for a mode listed in lnn");
                    return 2;
                }
                Gh[((size_t)l * nmax + n) * p->nrp + r] = v;
            }

    for (size_t k = 0; k < vpanels.size(); ++k) {   // basis of the virtual blocks: panel p, then panel q
        const VPanel& vp = vpanels[k];
        const size_t v = (size_t)lmax + 1 + k;
        for (int h = 0; h < 2; ++h) {
            const int n0 = h ? vp.q0 : vp.p0, cnt2 = h ? vp.nq : vp.np, dst0 = h ? kPanel : 0;
            for (int n = 0; n < cnt2; ++n)
                std::copy(Gh.begin() + ((size_t)vp.l * nmax + n0 + n) * p->nrp, Gh.begin() + ((size_t)vp.l * nmax + n0 + n + 1) * p->nrp,
                          Gh.begin() + (v * nmax + dst0 + n) * p->nrp);
        }
    }

    auto up = [&](auto& dbuf, const auto& hvec) -> int {
        SFB_TRY(dbuf.alloc(hvec.size()));
        SFB_CUDA_OK(cudaMemcpy(dbuf.p, hvec.data(), hvec.size() * sizeof(hvec[0]), cudaMemcpyHostToDevice));
        return 0;
    };
    int rc = 0;
    rc = rc ? rc : up(p->d_G, Gh);
    rc = rc ? rc : up(p->d_ell_ptr, p->ell_ptr);
    rc = rc ? rc : up(p->d_row_n, p->h_row_n);
    rc = rc ? rc : up(p->d_row_n2, p->h_row_n2);
    rc = rc ? rc : up(p->d_a, p->a_of_ell);
    rc = rc ? rc : up(p->d_blk_real, p->blk_real);
    rc = rc ? rc : up(p->d_pairidx, pairidx);
    rc = rc ? rc : up(p->d_nl_L, nl_L);
    rc = rc ? rc : up(p->d_nl_N, nl_N);
    rc = rc ? rc : up(p->d_es, es);
    {   // output index -> CSR slot (n, n' of the row)
        std::vector<int> slot_of_out((size_t)p->nout, 0);
        for (size_t sl = 0; sl < (size_t)p->ell_ptr[lmax + 1]; ++sl)   // slots of the l-blocks (virtual blocks renumber n)
            if (p->h_row_out[sl] >= 0) slot_of_out[p->h_row_out[sl]] = (int)sl;
        rc = rc ? rc : up(p->d_slot_of_out, slot_of_out);
    }
    if (p->ell_sorted) {  // upper-packed storage: column j keeps rows [0, rend(j)), rend = end of j's own l-block
        std::vector<int64_t> lend(lmax + 1, 0);
        for (int64_t o = 0; o < p->nout; ++o) lend[es[o] & 0x3fffffff] = o + 1;
        p->h_colbase.assign(p->nout + 1, 0);
        // column lengths rounded up to 4 doubles: every column starts 32-byte aligned (16-byte loads in the unpack kernel)
        for (int64_t o = 0; o < p->nout; ++o)
            p->h_colbase[o + 1] = p->h_colbase[o] + round_up(lend[es[o] & 0x3fffffff], 4);
        rc = rc ? rc : up(p->d_colbase, p->h_colbase);
    }
    rc = rc ? rc : p->d_row_out.alloc(p->h_row_out.size());
    rc = rc ? rc : p->d_ell_list.alloc(p->nblk);
    rc = rc ? rc : p->d_w2.alloc((size_t)(lmax + 1) * (lmax + 1) * (lmax + 1));
    rc = rc ? rc : p->d_W.alloc((size_t)(p->LMAX + 1) * p->nrp * p->nrp);
    if (rc) {
        delete p;
        return rc;
    }
    // the 3j table depends on lmax only: build it once
    {
        const int np = (lmax + 1) * (lmax + 1);
        w3j000sq_table_kernel<<<(unsigned)ceil_div((int64_t)np * 32, 128), 128>>>(p->d_w2.p, (int)lmax);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) {
            set_error(std::string("w3j000sq_table_kernel: ") + cudaGetErrorString(e));
            delete p;
            return 1;
        }
    }
    *out = p;
    return 0;
}

void cmix_plan_destroy(CmixPlan* p) {
    if (p) {
        for (auto& e : p->pend_ev) cudaEventDestroy(e);
        for (auto& e : p->pend_fill_ev) cudaEventDestroy(e);
        if (p->side_stream) cudaStreamDestroy(p->side_stream);
    }
    delete p;
}

int cmix_run(CmixPlan* p, const double* d_alm1, const double* d_alm2, int div2Lp1, int interchange, int64_t row_lo,
             int64_t row_hi, int64_t col_lo, int64_t col_hi, double* d_M, int64_t ldM, cudaStream_t stream,
             double* const* peers, int npeers, bool reuse_wl, bool mirror, bool upper_packed) {
    SFB_REQUIRE(p && d_alm1 && d_alm2 && d_M, "cmix_run: null pointer");
    SFB_REQUIRE(npeers >= 0 && npeers <= 7, "cmix_run: at most 7 peers");
    SFB_REQUIRE(0 <= row_lo && row_lo <= row_hi && row_hi <= p->nout, "cmix_run: bad row range");
    SFB_REQUIRE(0 <= col_lo && col_lo <= col_hi && col_hi <= p->nout, "cmix_run: bad column range");
    SFB_REQUIRE(ldM >= row_hi - row_lo, "cmix_run: ldM smaller than the row shard");
    p->t_wl = p->t_what = p->t_block = 0;
    p->flops_executed = 0;
    p->launches = 0;
    if (row_hi == row_lo || col_hi == col_lo) return 0;
    const bool sym = (d_alm1 == d_alm2);
    if (mirror) {
        // mirror mode fills the FULL matrix: d_M is its base, rows are an l-aligned range, all columns
        SFB_REQUIRE(sym && npeers == 0 && col_lo == 0 && col_hi == p->nout && ldM >= p->nout,
                    "cmix_run: mirror mode needs the auto-correlation path and the full matrix");
    }
    // mirror mode relies on the l-blocks being contiguous and ordered (every table ClnnModes builds is); for any other
    // table all blocks are formed directly.  With the full row and column range the two modes address d_M identically.
    if (mirror && !p->ell_sorted && row_lo == 0 && row_hi == p->nout) mirror = false;
    SFB_REQUIRE(!mirror || p->ell_sorted, "cmix_run: mirror mode needs an lnn table sorted by l");
    const int lmax = p->lmax, nrp = p->nrp;
    // register-Z kernel (cmix_regz.cu) for the auto-correlation path with nr <= 64; in mirror mode it forms the L >= l
    // blocks only and cmix_mirror_fill writes the blocks below the block diagonal afterwards
    const bool regz = cmix_regz_eligible(p, sym, npeers);
    if (upper_packed) {
        SFB_REQUIRE(regz && p->ell_sorted && !mirror && row_lo == 0 && row_hi == p->nout,
                    "cmix_run: upper-packed output needs the auto-correlation path with nr <= 64, an l-sorted lnn table "
                    "and the full row range");
    }
    const bool upper = mirror || upper_packed;  // only the blocks with L >= l are formed
    if (p->pending) SFB_TRY(cmix_resolve_times(p));   // a previous asynchronous run nobody asked the times of

    // rows of the shard -> output row index relative to row_lo, -1 elsewhere; ells touched by the shard
    std::vector<int> row_out(p->h_row_out.size());
    std::vector<char> ell_used(lmax + 1, 0);
    for (int blk = 0; blk < p->nblk; ++blk)
        for (int s = p->ell_ptr[blk]; s < p->ell_ptr[blk + 1]; ++s) {
            const int o = p->h_row_out[s];
            const bool in = (o >= row_lo && o < row_hi);
            row_out[s] = in ? (int)(mirror ? o : o - row_lo) : -1;
            if (in) ell_used[p->blk_real[blk]] = 1;
        }
    SFB_CUDA_OK(cudaMemcpyAsync(p->d_row_out.p, row_out.data(), row_out.size() * sizeof(int), cudaMemcpyHostToDevice,
                                stream));
    // L-blocks with at least one column in [col_lo, col_hi) (rows and columns share the same index set)
    std::vector<char> L_used(lmax + 1, 0);
    int Llo = lmax + 1, Lhi = -1;
    for (int L = 0; L <= lmax; ++L)
        for (int s = p->ell_ptr[L]; s < p->ell_ptr[L + 1]; ++s) {
            const int o = p->h_row_out[s];
            if (o >= col_lo && o < col_hi) {
                L_used[L] = 1;
                Llo = std::min(Llo, L);
                Lhi = std::max(Lhi, L);
            }
        }
    if (Lhi < 0) return 0;
    cudaEvent_t ev[4];
    for (auto& e : ev) SFB_CUDA_OK(cudaEventCreate(&e));
    std::vector<cudaEvent_t> chunk_ev;

    // ---- W_{L1} ----
    SFB_CUDA_OK(cudaEventRecord(ev[0], stream));
    if (!reuse_wl) {
        if (getenv("SFB_WL_FMA"))
            wl_build_kernel<<<p->LMAX + 1, 256, 0, stream>>>(d_alm1, d_alm2, p->d_W.p, p->LMAX, nrp);
        else
            wl_build_dmma_kernel<<<dim3(p->LMAX + 1, (unsigned)ceil_div(nrp, 64), (unsigned)ceil_div(nrp, 64)), 256, 0,
                                   stream>>>(d_alm1, d_alm2, p->d_W.p, p->LMAX, nrp);
        SFB_CUDA_OK(cudaGetLastError());
        p->launches = 1;
    }
    SFB_CUDA_OK(cudaEventRecord(ev[1], stream));

    // ---- ℓ chunks: Ŵ then blocks ----
    const size_t per_ell = (size_t)(lmax + 1) * nrp * nrp * sizeof(double);
    int chunk = (int)std::max<size_t>(1, p->what_budget_bytes / per_ell);
    chunk = std::min(chunk, lmax + 1);
    SFB_TRY(p->d_What.alloc((size_t)chunk * (lmax + 1) * nrp * nrp));

    CmixArgs args;
    args.G = p->d_G.p;
    args.What = p->d_What.p;
    args.ell_list = p->d_ell_list.p;
    args.ell_ptr = p->d_ell_ptr.p;
    args.row_out = p->d_row_out.p;
    args.row_n = p->d_row_n.p;
    args.row_n2 = p->d_row_n2.p;
    args.a_of_ell = p->d_a.p;
    args.real_ell = p->d_blk_real.p;
    args.pairidx = p->d_pairidx.p;
    size_t chunk_fill = 0;
    args.M[0] = d_M;
    for (int d = 0; d < 7; ++d) args.M[d + 1] = (d < npeers) ? peers[d] : nullptr;
    args.npeers = npeers + 1;
    args.ldM = ldM;
    args.lmax = lmax;
    args.nmax = p->nmax;
    args.nrp = nrp;
    args.S = p->S;
    args.div2Lp1 = div2Lp1;
    args.interchange = interchange;
    args.col_lo = (int)col_lo;
    args.col_hi = (int)col_hi;
    args.mirror = mirror ? 1 : 0;
    args.gl_global = 0;
    args.dbg = getenv("SFB_CMIX_DBG") ? atoi(getenv("SFB_CMIX_DBG")) : 0;

    double flops = 0.0;
    const int ngroups = 2 * (int)ceil_div(lmax + 1, 2 * kWhatGroup);
    std::vector<int> ell_list_all;  // device ell_list is filled per launch at distinct offsets
    // l-chunks of the Ŵ memory budget.  SFB_FILL_OVERLAP=1 (mirror mode with the register-Z kernel; an ablation that LOST, off
    // by default): the block kernel of a chunk runs in a few sub-chunks of equal cost and the mirror fill of sub-chunk i
    // (HBM-bound, on a second stream) runs under the block kernel of sub-chunk i+1 — the fill of the columns of sub-chunk i
    // reads only blocks (l in sub-chunk i, L >= l) and writes below the block diagonal, where no block kernel writes.
    // Measured at cfg4: the fill disappears from the critical path (0.78 -> 0.15 ms exposed) but the block kernels, whose
    // operand tiles stream from HBM, slow from 2.73 to 3.5 ms under the fill's traffic: 5.96 -> 6.09 ms per step.
    const bool overlap_fill = regz && mirror && getenv("SFB_FILL_OVERLAP") != nullptr;
    std::vector<std::pair<int, int>> lchunks;
    for (int ell0 = 0; ell0 <= lmax; ell0 += chunk) lchunks.emplace_back(ell0, std::min(lmax + 1, ell0 + chunk));
    std::vector<char> sub_end(lmax + 1, 0);   // 1: a block-kernel sub-chunk ends after this l
    if (overlap_fill) {
        int K = getenv("SFB_FILL_CHUNKS") ? atoi(getenv("SFB_FILL_CHUNKS")) : 4;
        K = std::max(1, std::min(K, 16));
        std::vector<double> cl(lmax + 1, 0.0);
        double total = 0;
        for (int l = 0; l <= lmax; ++l) {
            if (!ell_used[l] || p->a_of_ell[l] == 0) continue;
            const double ap = 8.0 * ((p->a_of_ell[l] + 7) / 8);
            for (int L = l; L <= lmax; ++L) {
                if (!L_used[L]) continue;
                const double b = p->a_of_ell[L];
                cl[l] += 2.0 * ap * nrp * nrp * b + ap * ap * nrp * b * (b + 1);
            }
            total += cl[l];
        }
        double acc = 0;
        int made = 0;
        for (int l = 0; l < lmax; ++l) {
            acc += cl[l];
            if (made < K - 1 && acc >= total * (made + 1) / K) sub_end[l] = 1, ++made;
        }
    }
    // output index range of every l (mirror mode: the table is l-sorted, the rows of an l are consecutive)
    std::vector<int> out_lo(lmax + 1, 1 << 30), out_hi(lmax + 1, -1);
    if (overlap_fill)
        for (int l = 0; l <= lmax; ++l)
            for (int s = p->ell_ptr[l]; s < p->ell_ptr[l + 1]; ++s)
                if (row_out[s] >= 0) {
                    out_lo[l] = std::min(out_lo[l], row_out[s]);
                    out_hi[l] = std::max(out_hi[l], row_out[s] + 1);
                }
    if (overlap_fill && !p->side_stream) SFB_CUDA_OK(cudaStreamCreateWithFlags(&p->side_stream, cudaStreamNonBlocking));
    std::vector<cudaEvent_t> fill_ev;
    for (const auto& lc : lchunks) {
        const int ell0 = lc.first, ell1 = lc.second;
        bool any = false;
        for (int l = ell0; l < ell1; ++l) any = any || (ell_used[l] && p->a_of_ell[l] > 0);
        if (!any) continue;
        cudaEvent_t e0, e1, e2;
        SFB_CUDA_OK(cudaEventCreate(&e0));
        SFB_CUDA_OK(cudaEventCreate(&e1));
        SFB_CUDA_OK(cudaEventCreate(&e2));
        SFB_CUDA_OK(cudaEventRecord(e0, stream));
        // only the l-blocks this row shard touches
        std::vector<int> wells;
        for (int l = ell0; l < ell1; ++l)
            if (ell_used[l] && p->a_of_ell[l] > 0) wells.push_back(l);
        SFB_TRY(p->d_what_ells.alloc(lmax + 1));
        SFB_CUDA_OK(cudaMemcpyAsync(p->d_what_ells.p, wells.data(), wells.size() * sizeof(int), cudaMemcpyHostToDevice,
                                    stream));
        if (getenv("SFB_WHAT_FMA") == nullptr && (nrp * nrp) % 64 == 0) {
            const int SW = (int)round_up(lmax + 4, 16) + 4;   // >= lmax + 4 columns, ≡ 4 (mod 16)
            const size_t wsm_bytes = (size_t)kWhatRows * SW * sizeof(double);
            SFB_REQUIRE(wsm_bytes <= 200 * 1024, "what_build: lmax too large for the shared-memory 3j tile");
            SFB_CUDA_OK(cudaFuncSetAttribute(what_build_dmma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)wsm_bytes));
            const int ng32 = 2 * (int)ceil_div(lmax + 1, 2 * kWhatRows);
            what_build_dmma_kernel<<<dim3(ng32, (unsigned)wells.size(), (unsigned)ceil_div(nrp * nrp, 256 * kWhatEPW)), 128, wsm_bytes,
                                     stream>>>(p->d_W.p, p->d_w2.p, p->d_What.p, p->d_what_ells.p, ell0, lmax, nrp, Llo,
                                               Lhi, upper ? 1 : 0, SW);
        } else {
            what_build_kernel<<<dim3(ngroups, (unsigned)wells.size()), 256, kWhatGroup * (lmax + 1) * sizeof(double),
                                stream>>>(p->d_W.p, p->d_w2.p, p->d_What.p, p->d_what_ells.p, ell0, lmax, nrp, Llo, Lhi,
                                          upper ? 1 : 0);
        }
        SFB_CUDA_OK(cudaGetLastError());
        p->launches++;
        SFB_CUDA_OK(cudaEventRecord(e1, stream));
        args.ell0 = ell0;
        chunk_fill = 0;  // the previous l-chunk's kernels have completed (event sync below)
        if (regz) {
            // sub-chunks of this memory chunk (one, unless the fill is overlapped)
            size_t w0 = 0;
            while (w0 < wells.size()) {
                size_t w1 = w0;
                while (w1 < wells.size()) {
                    const int l = wells[w1++];
                    bool cut = false;
                    for (int q = l; q < (w1 < wells.size() ? wells[w1] : l + 1) && !cut; ++q) cut = sub_end[q] != 0;
                    if (cut) break;
                }
                std::vector<int> blocks;
                for (size_t wi = w0; wi < w1; ++wi) {
                    const int l = wells[wi];
                    for (int v : p->blk_of_ell[l])            // the l-block itself, or its virtual row blocks
                        for (int L = upper ? l : 0; L <= lmax; ++L) {
                            if (!L_used[L] || p->a_of_ell[L] == 0) continue;
                            const int desc[8] = {v, L, p->a_of_ell[v], p->a_of_ell[L], p->ell_ptr[v],
                                                 p->ell_ptr[v + 1] - p->ell_ptr[v], (l - ell0) * (lmax + 1) + L, 0};
                            blocks.insert(blocks.end(), desc, desc + 8);
                        }
                }
                SFB_TRY(cmix_regz_run(p, blocks, p->d_What.p, div2Lp1, interchange, col_lo, col_hi, d_M, ldM,
                                      upper_packed ? p->d_colbase.p + col_lo : nullptr, stream, &flops, &p->launches));
                if (overlap_fill) {
                    int c0 = 1 << 30, c1 = -1;
                    for (size_t wi = w0; wi < w1; ++wi)
                        c0 = std::min(c0, out_lo[wells[wi]]), c1 = std::max(c1, out_hi[wells[wi]]);
                    if (c1 > c0) {
                        cudaEvent_t kd, f0, f1;
                        SFB_CUDA_OK(cudaEventCreateWithFlags(&kd, cudaEventDisableTiming));
                        SFB_CUDA_OK(cudaEventCreate(&f0));
                        SFB_CUDA_OK(cudaEventCreate(&f1));
                        SFB_CUDA_OK(cudaEventRecord(kd, stream));
                        SFB_CUDA_OK(cudaStreamWaitEvent(p->side_stream, kd, 0));
                        SFB_CUDA_OK(cudaEventDestroy(kd));   // released once the wait has been satisfied
                        SFB_CUDA_OK(cudaEventRecord(f0, p->side_stream));
                        SFB_TRY(cmix_mirror_fill(p, c0, c1, div2Lp1, interchange, d_M, ldM, p->side_stream));
                        p->launches++;
                        SFB_CUDA_OK(cudaEventRecord(f1, p->side_stream));
                        fill_ev.push_back(f0);
                        fill_ev.push_back(f1);
                    }
                }
                w0 = w1;
            }
        }
        for (int AT = regz ? 0 : p->amax_tiles; AT >= 1; --AT) {
            std::vector<int> ells;   // row blocks of this tile class
            for (int l = ell0; l < ell1; ++l)
                if (ell_used[l] && p->a_of_ell[l] > 0)
                    for (int v : p->blk_of_ell[l])
                        if ((p->a_of_ell[v] + 7) / 8 == AT) ells.push_back(v);
            if (ells.empty()) continue;
            const size_t off = ell_list_all.size();
            ell_list_all.insert(ell_list_all.end(), ells.begin(), ells.end());
            SFB_REQUIRE(ell_list_all.size() <= (size_t)p->nblk, "cmix_run: internal ell list overflow");
            SFB_CUDA_OK(cudaMemcpyAsync(p->d_ell_list.p + off, ells.data(), ells.size() * sizeof(int),
                                        cudaMemcpyHostToDevice, stream));
            args.ell_list = p->d_ell_list.p + off;
            int max_rows = 0;
            for (int l : ells) max_rows = std::max(max_rows, p->ell_ptr[l + 1] - p->ell_ptr[l]);
            // chunk list (L, N0, N1) for this tile class
            bool glg = false;
            const int NC = cmix_chunk_size(AT, sym, p->S, p->nmax, max_rows, &glg);
            SFB_REQUIRE(NC >= 1, "cmix: nr too large for the shared-memory tiling of this build (nr * 8 * ceil(nmax_l/8) "
                                 "doubles must fit twice in 220 KB)");
            args.gl_global = glg ? 1 : 0;
            std::vector<int> chL, chN0, chN1;
            for (int L = 0; L <= lmax; ++L)
                for (int n0 = 0; L_used[L] && n0 < p->a_of_ell[L]; n0 += NC) {
                    chL.push_back(L);
                    chN0.push_back(n0);
                    chN1.push_back(std::min(n0 + NC, p->a_of_ell[L]));
                }
            const size_t coff = chunk_fill;
            chunk_fill += chL.size();
            SFB_TRY(p->d_chunks.alloc(3 * (size_t)(lmax + 1) * p->nmax * 4));
            SFB_REQUIRE(3 * chunk_fill <= p->d_chunks.n, "cmix_run: internal chunk list overflow");
            int* dch = p->d_chunks.p + 3 * coff;
            SFB_CUDA_OK(cudaMemcpyAsync(dch, chL.data(), chL.size() * sizeof(int), cudaMemcpyHostToDevice, stream));
            SFB_CUDA_OK(cudaMemcpyAsync(dch + chL.size(), chN0.data(), chL.size() * sizeof(int), cudaMemcpyHostToDevice,
                                        stream));
            SFB_CUDA_OK(cudaMemcpyAsync(dch + 2 * chL.size(), chN1.data(), chL.size() * sizeof(int),
                                        cudaMemcpyHostToDevice, stream));
            args.ch_L = dch;
            args.ch_N0 = dch + chL.size();
            args.ch_N1 = dch + 2 * chL.size();
            args.NC = NC;
            if (sym)
                SFB_TRY(launch_cmix_at<true>(AT, args, (int)chL.size(), (int)ells.size(), p->nmax, max_rows, stream));
            else
                SFB_TRY(launch_cmix_at<false>(AT, args, (int)chL.size(), (int)ells.size(), p->nmax, max_rows, stream));
            p->launches++;
            for (int l : ells)
                for (int L = 0; L <= lmax; ++L) {
                    if (!L_used[L] || (mirror && L < p->blk_real[l])) continue;
                    const double b = p->a_of_ell[L], ap = AT * 8.0;
                    flops += (sym ? 1.0 : 2.0) * (2.0 * ap * nrp * nrp * b + 2.0 * ap * ap * nrp * b * (b + 1) / 2);
                }
        }
        SFB_CUDA_OK(cudaEventRecord(e2, stream));
        chunk_ev.push_back(e0);
        chunk_ev.push_back(e1);
        chunk_ev.push_back(e2);
    }
    p->t_fill = 0;
    if (regz && mirror) {
        SFB_CUDA_OK(cudaEventRecord(ev[2], stream));
        if (overlap_fill) {
            if (!fill_ev.empty()) SFB_CUDA_OK(cudaStreamWaitEvent(stream, fill_ev.back(), 0));   // join the fill stream
        } else {
            SFB_TRY(cmix_mirror_fill(p, row_lo, row_hi, div2Lp1, interchange, d_M, ldM, stream));
            p->launches++;
        }
        SFB_CUDA_OK(cudaEventRecord(ev[3], stream));
    }
    for (auto& e : p->pend_fill_ev) cudaEventDestroy(e);
    p->pend_fill_ev = fill_ev;
    p->flops_executed = flops;
    p->pend_ev.clear();
    p->pend_ev.push_back(ev[0]);
    p->pend_ev.push_back(ev[1]);
    p->pend_ev.insert(p->pend_ev.end(), chunk_ev.begin(), chunk_ev.end());
    p->pend_ev.push_back(ev[2]);
    p->pend_ev.push_back(ev[3]);
    p->pend_fill = regz && mirror;
    p->pending = true;
    if (p->async_times) return 0;
    SFB_TRY(cmix_resolve_times(p));
    SFB_CUDA_OK(cudaStreamSynchronize(stream));
    return 0;
}

// Wait for the last run's events and fill t_wl / t_what / t_block (incl. t_fill) / t_fill.
int cmix_resolve_times(CmixPlan* p) {
    if (!p || !p->pending) return 0;
    auto& ev = p->pend_ev;
    const size_t n = ev.size();
    float t_what = 0.f, t_block = 0.f;
    p->t_fill = 0.f;
    int rc = 0;
    if (n >= 4) {
        const size_t nch = (n - 4) / 3;   // (begin, Ŵ end, blocks end) per l-chunk, after the two W_L1 events
        const size_t last = p->pend_fill ? n - 1 : (nch ? 2 + 3 * (nch - 1) + 2 : 1);
        if (cudaEventSynchronize(ev[last]) != cudaSuccess) rc = 1;
        if (!rc) {
            cudaEventElapsedTime(&p->t_wl, ev[0], ev[1]);
            for (size_t i = 0; i < nch; ++i) {
                const size_t c = 2 + 3 * i;
                float a_ms = 0, b_ms = 0;
                cudaEventElapsedTime(&a_ms, ev[c], ev[c + 1]);
                cudaEventElapsedTime(&b_ms, ev[c + 1], ev[c + 2]);
                t_what += a_ms;
                t_block += b_ms;
            }
            p->t_k3 = t_block;
            if (p->pend_fill) {
                // serial fill: the span between the last two events; overlapped fill: that span is only the exposed tail
                // (block kernels done, fill stream still draining) and t_fill sums the fill launches of the side stream
                float tail = 0;
                cudaEventElapsedTime(&tail, ev[n - 2], ev[n - 1]);
                t_block += tail;
                p->t_fill = tail;
                if (!p->pend_fill_ev.empty()) {
                    p->t_fill = 0;
                    for (size_t i = 0; i + 1 < p->pend_fill_ev.size(); i += 2) {
                        float f = 0;
                        cudaEventElapsedTime(&f, p->pend_fill_ev[i], p->pend_fill_ev[i + 1]);
                        p->t_fill += f;
                    }
                }
            }
        }
    }
    for (auto& e : ev) cudaEventDestroy(e);
    ev.clear();
    for (auto& e : p->pend_fill_ev) cudaEventDestroy(e);
    p->pend_fill_ev.clear();
    p->pending = false;
    p->t_what = t_what;
    p->t_block = t_block;
    if (rc) {
        set_error("cmix_resolve_times: event synchronisation failed");
        return 1;
    }
    return 0;
}

// Split the columns into at most `k` contiguous ranges of roughly equal cost, cut on L-block boundaries.
// Falls back to a single range when the caller's table does not keep each L-block contiguous.
std::vector<std::pair<int64_t, int64_t>> cmix_col_chunks(const CmixPlan* p, int k) {
    std::vector<std::pair<int64_t, int64_t>> out;
    const int lmax = p->lmax;
    std::vector<int64_t> first(lmax + 2, -1);
    std::vector<double> cost(lmax + 1, 0.0);
    bool contiguous = true;
    int64_t expect = 0;
    for (int L = 0; L <= lmax; ++L) {
        const int c0 = p->ell_ptr[L], c1 = p->ell_ptr[L + 1];
        first[L] = expect;
        for (int s = c0; s < c1; ++s)
            if (p->h_row_out[s] != expect++) contiguous = false;
        const double b = p->a_of_ell[L];
        for (int l = 0; l <= lmax && c1 > c0; ++l) {
            const double ap = 8.0 * ((p->a_of_ell[l] + 7) / 8);
            cost[L] += 2.0 * ap * p->nrp * p->nrp * b + 2.0 * ap * ap * p->nrp * b * (b + 1) / 2;
        }
    }
    first[lmax + 1] = expect;
    if (!contiguous || k <= 1) {
        out.emplace_back(0, p->nout);
        return out;
    }
    double total = 0;
    for (double c : cost) total += c;
    double acc = 0;
    int64_t start = 0;
    int made = 0;
    for (int L = 0; L <= lmax; ++L) {
        acc += cost[L];
        if (acc >= total * (made + 1) / k && first[L + 1] > start && made < k - 1) {
            out.emplace_back(start, first[L + 1]);
            start = first[L + 1];
            ++made;
        }
    }
    if (start < p->nout) out.emplace_back(start, p->nout);
    return out;
}

// The same split restricted to the columns [lo, hi): at most k ranges of roughly equal full-height cost, cut on L-block
// boundaries inside the range (lo and hi themselves may lie anywhere).
std::vector<std::pair<int64_t, int64_t>> cmix_col_chunks_range(const CmixPlan* p, int64_t lo, int64_t hi, int k) {
    std::vector<std::pair<int64_t, int64_t>> out;
    if (hi <= lo) return out;
    if (lo == 0 && hi == p->nout) return cmix_col_chunks(p, k);
    if (!p->ell_sorted || k <= 1) {
        out.emplace_back(lo, hi);
        return out;
    }
    const int lmax = p->lmax;
    // l-sorted table: block L covers the output columns [first[L], first[L+1])
    std::vector<int64_t> first(lmax + 2, 0);
    for (int L = 0; L <= lmax; ++L) first[L + 1] = first[L] + (p->ell_ptr[L + 1] - p->ell_ptr[L]);
    std::vector<double> cost(lmax + 1, 0.0);
    double total = 0;
    for (int L = 0; L <= lmax; ++L) {
        const int64_t c0 = std::max(first[L], lo), c1 = std::min(first[L + 1], hi);
        if (c1 <= c0) continue;
        const double b = p->a_of_ell[L];
        double c = 0;
        for (int l = 0; l <= lmax; ++l) {
            if (p->ell_ptr[l + 1] == p->ell_ptr[l]) continue;
            const double ap = 8.0 * ((p->a_of_ell[l] + 7) / 8);
            c += 2.0 * ap * p->nrp * p->nrp * b + 2.0 * ap * ap * p->nrp * b * (b + 1) / 2 +
                 2.0 * p->nrp * p->nrp * (std::min(l, L) + 1);
        }
        cost[L] = c;    // a partially covered block still costs (almost) the whole block
        total += c;
    }
    double acc = 0;
    int64_t start = lo;
    int made = 0;
    for (int L = 0; L <= lmax; ++L) {
        if (cost[L] == 0) continue;
        acc += cost[L];
        const int64_t end = std::min(first[L + 1], hi);
        if (acc >= total * (made + 1) / k && end > start && made < k - 1 && end < hi) {
            out.emplace_back(start, end);
            start = end;
            ++made;
        }
    }
    if (start < hi) out.emplace_back(start, hi);
    return out;
}

// Column bounds for `ndev` devices straight from the caller's mode table (no plan needed): bounds[0..ndev], cut where l
// changes (when the table is sorted by l) so that every device forms whole (l, L) blocks, balanced on the full-height cost.
int cmix_col_bounds_from_lnn(const int64_t* lnn, int64_t lnnsize, int64_t lnn_min, int64_t nr, int64_t nmax, int64_t lmax,
                             int ndev, int64_t* bounds) {
    SFB_REQUIRE(lnn && bounds && ndev >= 1, "cmix_col_bounds_from_lnn: bad arguments");
    SFB_REQUIRE(lnn_min >= 1 && lnn_min <= lnnsize, "lnn_min out of range");
    const int64_t n = lnnsize - lnn_min + 1;
    std::vector<int> a(lmax + 1, 0), cols(lmax + 1, 0);
    bool sorted = true;
    for (int64_t i = lnn_min - 1; i < lnnsize; ++i) {
        const int64_t l = lnn[3 * i], n1 = lnn[3 * i + 1], n2 = lnn[3 * i + 2];
        SFB_REQUIRE(l >= 0 && l <= lmax && n1 >= 1 && n2 >= 1 && n1 <= nmax && n2 <= nmax, "lnn entry out of range");
        a[l] = std::max(a[l], (int)std::max(n1, n2));
        cols[l]++;
        if (i > lnn_min - 1 && l < lnn[3 * (i - 1)]) sorted = false;
    }
    const double nrp = (double)round_up(nr, 8);
    std::vector<double> colcost(lmax + 1, 0.0);
    for (int L = 0; L <= lmax; ++L) {
        if (!cols[L]) continue;
        const double b = a[L];
        double c = 0;
        for (int l = 0; l <= lmax; ++l) {
            if (!cols[l]) continue;
            const double ap = 8.0 * ((a[l] + 7) / 8);
            c += 2.0 * ap * nrp * nrp * b + 2.0 * ap * ap * nrp * b * (b + 1) / 2 + 2.0 * nrp * nrp * (std::min(l, L) + 1);
        }
        colcost[L] = c / cols[L];
    }
    std::vector<double> cum((size_t)n + 1, 0.0);
    for (int64_t o = 0; o < n; ++o) cum[o + 1] = cum[o] + colcost[lnn[3 * (o + lnn_min - 1)]];
    bounds[0] = 0;
    for (int g = 1; g < ndev; ++g) {
        const double target = cum[n] * g / ndev;
        int64_t c = std::lower_bound(cum.begin(), cum.end(), target) - cum.begin();
        c = std::min<int64_t>(std::max<int64_t>(c, 0), n);
        if (sorted && c > 0 && c < n) {   // snap to the nearest l-block boundary
            int64_t up = c, dn = c;
            while (up < n && lnn[3 * (up + lnn_min - 1)] == lnn[3 * (up + lnn_min - 2)]) ++up;
            while (dn > 0 && dn < n && lnn[3 * (dn + lnn_min - 1)] == lnn[3 * (dn + lnn_min - 2)]) --dn;
            c = (cum[up] - target <= target - cum[dn]) ? up : dn;
        }
        bounds[g] = std::max(c, bounds[g - 1]);
    }
    bounds[ndev] = n;
    return 0;
}

// l-block aligned ranges of rows (== columns: same index set) with roughly equal mirror-mode cost
// (block (l, L) is only formed for L >= l).  One range when the table does not keep l-blocks contiguous.
std::vector<std::pair<int64_t, int64_t>> cmix_row_chunks_mirror(const CmixPlan* p, int k) {
    std::vector<std::pair<int64_t, int64_t>> out;
    const int lmax = p->lmax;
    std::vector<int64_t> first(lmax + 2, 0);
    std::vector<double> cost(lmax + 1, 0.0);
    bool contiguous = true;
    int64_t expect = 0;
    for (int l = 0; l <= lmax; ++l) {
        first[l] = expect;
        for (int s = p->ell_ptr[l]; s < p->ell_ptr[l + 1]; ++s)
            if (p->h_row_out[s] != expect++) contiguous = false;
        if (p->ell_ptr[l + 1] == p->ell_ptr[l]) continue;
        const double ap = 8.0 * ((p->a_of_ell[l] + 7) / 8);
        for (int L = l; L <= lmax; ++L) {
            const double b = p->a_of_ell[L];
            cost[l] += 2.0 * ap * p->nrp * p->nrp * b + 2.0 * ap * ap * p->nrp * b * (b + 1) / 2;
        }
    }
    first[lmax + 1] = expect;
    if (!contiguous || k <= 1) {
        out.emplace_back(0, p->nout);
        return out;
    }
    double total = 0;
    for (double c : cost) total += c;
    double acc = 0;
    int64_t start = 0;
    int made = 0;
    for (int l = 0; l <= lmax; ++l) {
        acc += cost[l];
        if (acc >= total * (made + 1) / k && first[l + 1] > start && made < k - 1) {
            out.emplace_back(start, first[l + 1]);
            start = first[l + 1];
            ++made;
        }
    }
    if (start < p->nout) out.emplace_back(start, p->nout);
    return out;
}

// =============================================================================================
// win_lnn (src/windows.jl:382-418): W_lnn' = Σ_r [r √Δr g_nl(r)] [r √Δr g_n'l(r)] W_00(r)/√(4π), one warp per (l,n,n').
// W_00(r) is the (l,m) = (0,0) coefficient of stage 1: the first re-plane of the planar alm buffer.
__global__ void __launch_bounds__(256) win_lnn_kernel(const double* __restrict__ G, const double* __restrict__ alm,
                                                      const int* __restrict__ es, const int* __restrict__ slot_of_out,
                                                      const int* __restrict__ row_n, const int* __restrict__ row_n2,
                                                      int nmax, int nrp, int nout, double* __restrict__ out) {
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (i >= nout) return;
    const int l = es[i] & 0x3fffffff, s = slot_of_out[i];
    const double* g1 = G + ((size_t)l * nmax + row_n[s]) * nrp;
    const double* g2 = G + ((size_t)l * nmax + row_n2[s]) * nrp;
    double acc = 0.0;
    for (int r = lane; r < nrp; r += 32) acc = fma(g1[r] * g2[r], alm[r], acc);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    if (lane == 0) out[i] = acc * 0.28209479177387814;  // 1/√(4π)
}

int win_lnn_run(CmixPlan* p, const double* d_alm, double* d_out, cudaStream_t stream) {
    SFB_REQUIRE(p && d_alm && d_out, "win_lnn: null pointer");
    const int n = (int)p->nout;
    win_lnn_kernel<<<(unsigned)ceil_div((int64_t)n * 32, 256), 256, 0, stream>>>(
        p->d_G.p, d_alm, p->d_es.p, p->d_slot_of_out.p, p->d_row_n.p, p->d_row_n2.p, p->nmax, p->nrp, n, d_out);
    SFB_CUDA_OK(cudaGetLastError());
    return 0;
}

// =============================================================================================
// host complex Wr_lm -> planar device alm

__global__ void alm_planar_kernel(const double* __restrict__ src, double* __restrict__ dst, int nr, int nrp, int lmax2,
                                  int layout) {
    // src: [lm_src][r] complex interleaved (column-major nr x lmsize); dst: [lm m-major][comp][nrp]
    const int lmsize = (lmax2 + 1) * (lmax2 + 2) / 2;
    const long long total = (long long)lmsize * nrp;
    for (long long x = blockIdx.x * (long long)blockDim.x + threadIdx.x; x < total;
         x += (long long)gridDim.x * blockDim.x) {
        const int lm = (int)(x / nrp), r = (int)(x - (long long)lm * nrp);
        // decode m-major index -> (l, m)
        int m = 0, base = 0;
        while (base + (lmax2 + 1 - m) <= lm) {
            base += lmax2 + 1 - m;
            ++m;
        }
        const int l = m + (lm - base);
        const long long s = layout ? ((long long)l * (l + 1) / 2 + m) : lm;
        double re = 0.0, im = 0.0;
        if (r < nr) {
            re = src[2 * (s * nr + r)];
            im = src[2 * (s * nr + r) + 1];
        }
        dst[((size_t)lm * 2) * nrp + r] = re;
        dst[((size_t)lm * 2 + 1) * nrp + r] = im;
    }
}

int alm_from_host(const double* h_wrlm, int64_t nr, int lmax2, int layout, DevBuf<double>& d_alm, int nrp,
                  cudaStream_t stream) {
    const size_t lmsize = (size_t)(lmax2 + 1) * (lmax2 + 2) / 2;
    DevBuf<double> tmp;
    SFB_TRY(tmp.alloc(lmsize * nr * 2));
    SFB_CUDA_OK(cudaMemcpyAsync(tmp.p, h_wrlm, lmsize * nr * 2 * sizeof(double), cudaMemcpyHostToDevice, stream));
    SFB_TRY(d_alm.alloc(lmsize * 2 * nrp));
    alm_planar_kernel<<<1024, 256, 0, stream>>>(tmp.p, d_alm.p, (int)nr, nrp, lmax2, layout);
    SFB_CUDA_OK(cudaGetLastError());
    SFB_CUDA_OK(cudaStreamSynchronize(stream));
    return 0;
}

}  // namespace sfb
