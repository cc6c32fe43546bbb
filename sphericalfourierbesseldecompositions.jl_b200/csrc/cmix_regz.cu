// Coupling-matrix block kernel, register-resident Z variant (auto-correlation, nr <= 64).
//
// Replaces calc_cmixlnnLNN! + calc_cmix (reference src/windows.jl:613-627, 700-746) for one (l, L) block per CTA:
//   Z_N[n][r']      = Σ_r  G_ln[r] G_LN[r] Ŵ_lL[r][r']                (DMMA; accumulators STAY in registers)
//   T_NN'[n][n']    = Σ_r' Z_N[n][r'] G_LN'[r'] G_ln'[r']   N' >= N   (DMMA; Z_N is the A operand straight from
//                                                                      the accumulator registers)
//   M[(l,n,n'),(L,N,N')] = c_L (T[n][n'] + [N≠N'] T[n'][n])            (src/windows.jl:727-736)
// In mirror mode only the blocks with L >= l are formed; cmix_mirror_fill_kernel then fills the blocks below the
// block diagonal from M[i',i] = M[i,i'] f_i / f_i', f = c_l (1 + [n≠n'])  (the un-symmetrised kernel is
// symmetric).  Storing the mirror image from the tile itself (one 8-byte store per 32-byte sector) made the kernel
// store-bound: 5.2 ms against 3.0 ms + 0.x ms for the separate coalesced pass (cfg4, tools/ablate_regz.sh).
//
// Chaining the two GEMMs through registers works because the contraction index of an m8n8k4 DMMA can be permuted
// freely as long as A and B agree: the C fragment of lane (g,t) holds columns r' = 8jt+2t+{0,1}, so the T phase
// contracts "k-step (jt,e)" over r' = 8jt+2t+e and reads its B fragments (G_ln'[r']) at the same permuted index.
// The same permutation is used for the r contraction of the Z phase, which makes every operand a 16-byte
// shared-memory load (row stride ≡ 8 mod 16 doubles: conflict free) feeding two k-steps.
//
// A warp owns one N at a time (grabbed from a shared-memory counter, heaviest first), so after the block's operands
// are staged (TMA tensor copies, or cp.async) there is no CTA-wide synchronisation: each warp runs Z -> tiles -> epilogue
// on its own.
#include "cmix.cuh"

#include <cuda.h>   // CUtensorMap (types only: cuTensorMapEncodeTiled is fetched through cudaGetDriverEntryPoint)

#include <algorithm>
#include <cstdlib>
#include <cstring>

namespace sfb {

// ---- TMA staging (default; SFB_REGZ_CPASYNC=1 selects the cp.async variant): the three operand tiles of a block arrive by
// cp.async.bulk.tensor.2d, completion on an mbarrier (measured at cfg4: 2.897 ms vs 2.935 ms with cp.async, same bits).  The tensor maps describe G as {nrp, (lmax+1)·nmax} and the Ŵ chunk as {nrp, nblocks·nrp} doubles with boxes
// whose inner extent is the PADDED shared-memory row (K + 8 doubles): the out-of-bounds columns [nrp, K+8) are zero-filled
// by the copy engine, so the tile lands directly in the conflict-free padded layout the DMMA fragment loads need.
struct RegzTmaps {
    CUtensorMap gl;   // box {K+8, AP}   rows of the row-side l (rows >= a belong to n >= nmax_l or the next l: never read)
    CUtensorMap gL;   // box {K+8, nmax} rows of the column-side L
    CUtensorMap w;    // box {K+8, nrp}  one Ŵ_lL block
};

__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
    const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(a), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(a), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "SFB_MBAR_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra SFB_MBAR_DONE;\n"
        "bra SFB_MBAR_WAIT;\n"
        "SFB_MBAR_DONE:\n"
        "}\n" ::"r"(a),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, int c0, int c1, unsigned long long* bar) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    const unsigned b = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n" ::"r"(d),
                 "l"(map), "r"(c0), "r"(c1), "r"(b)
                 : "memory");
}

struct RegzArgs {
    const double* G;       // [ell][nmax][nrp]
    const double* What;    // [what_index][nrp][nrp]
    const int* blocks;     // 8 ints per block: rowell, colell, a, b, r0, nrows, what_index, unused
    const int* row_out;    // output row per CSR slot (absolute in mirror mode, relative to the shard otherwise) or -1
    const int* row_n;
    const int* row_n2;
    const int* pairidx;    // [L][nmax][nmax] -> output column or -1
    double* M;
    long long ldM;
    const long long* colbase;  // optional: element offset of output column c (relative to col_lo) instead of c * ldM
    int nmax, nrp;
    int col_lo, col_hi;
    int div2Lp1, interchange;
    int dbg;
    int force_rowtab;      // SFB_REGZ_FULLDIAG: form all sub-tiles of the N' = N tiles
};

constexpr int kRegzNoRow = 0x3FFFFF;   // packed row-table marker: row not in the shard

// per-warp tile staging stride ≡ 8 (mod 16): conflict-free 16-byte tile stores; the epilogue reads run along the diagonals
// of the tile (the reference orders an l-block by n'-n, then n: src/modes.jl getidx), i.e. with stride TLD + 1 (odd)
// (AP = 32: 34 instead of 40 — two-way conflicts on the tile stores, but the 12 KB saved let the persistent kernel's two
//  operand stages fit next to eight staging tiles: cfg5's four-tile blocks then run persistent as well)
__host__ __device__ constexpr int regz_tld(int AP) { return (AP % 16 == 8) ? AP : (AP == 32 ? 34 : AP + 8); }

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gsrc));
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

// Epilogue rows idx0 + 32u, u < U, of the current tile: M[row] = T[n][n'] (+ T[n'][n] off the N diagonal) read from the
// warp's staged tile through the packed row table (the tile is already scaled: the factor c_L sits in Z)
template <int U, int TLD>
__device__ __forceinline__ void regz_store_rows(const double* Tw, const int* rowtab, int idx0, int nrows, double* Mc,
                                                bool offdiag, int interchange) {
    int pk[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const int idx = idx0 + 32 * u;
        pk[u] = (idx < nrows) ? rowtab[idx] : kRegzNoRow;
    }
    double v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const int n = (pk[u] >> 22) & 31, n2 = (pk[u] >> 27) & 31;
        const double A = Tw[n * TLD + n2], B = Tw[n2 * TLD + n];
        v[u] = interchange ? B : (offdiag ? A + B : A);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const int orow = pk[u] & kRegzNoRow;
        if (orow != kRegzNoRow) Mc[orow] = v[u];
    }
}

// T phase of one (N, N') tile: acc[i][j] = Σ_r' Z[i][.][r'] G_LN'[r'] G_l(8j+.)[r'].  UPPER: only the sub-tiles with i <= j
// (T_NN is symmetric, so the N' = N tile needs no more)
template <int AT, int NT, int S, bool UPPER>
__device__ __forceinline__ void regz_tile(double (&acc)[AT][AT][2], const double (&Z)[AT][NT][2], const double* glS,
                                          const double* glB) {
#pragma unroll
    for (int i = 0; i < AT; ++i)
#pragma unroll
        for (int j = 0; j < AT; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
#pragma unroll
    for (int jt = 0; jt < NT; ++jt) {
        const double2 s2 = *reinterpret_cast<const double2*>(glS + 8 * jt);
        double2 gv[AT];
#pragma unroll
        for (int j = 0; j < AT; ++j) gv[j] = *reinterpret_cast<const double2*>(glB + j * 8 * S + 8 * jt);
#pragma unroll
        for (int i = 0; i < AT; ++i) {
            const double a0 = Z[i][jt][0] * s2.x, a1 = Z[i][jt][1] * s2.y;
#pragma unroll
            for (int j = UPPER ? i : 0; j < AT; ++j) {
                dmma884(acc[i][j], a0, gv[j].x);
                dmma884(acc[i][j], a1, gv[j].y);
            }
        }
    }
}

// Work of one warp on a staged (l, L) block: grab an N, Z phase, tiles N' >= N with their epilogue; repeat until the block's
// N counter runs out.  Shared by the one-block-per-CTA kernel and the persistent kernel.
template <int AT, int NT>
__device__ __forceinline__ void regz_warp_work(const RegzArgs& p, const double* Gl, const double* GL, const double* Ws,
                                               double* Tw, const int* rowtab, int* counter, int colell, int a, int b,
                                               int nrows, bool rows_upper) {
    constexpr int AP = AT * 8;
    constexpr int K = NT * 8;
    constexpr int S = K + 8;
    constexpr int TLD = regz_tld(AP);
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int nmax = p.nmax;
    (void)a;
    const double inv4pi = 0.07957747154594767;
    const double scale = (p.div2Lp1 ? 1.0 : (2.0 * colell + 1.0)) * inv4pi;

    for (;;) {
        int N = 0;
        if (lane == 0) N = atomicAdd(counter, 1);
        N = __shfl_sync(0xffffffffu, N, 0);
        if (N >= b) break;
        // output columns of (N, N'), N' = N + 32 c + lane (c-th group of 32; one group unless nmax_L > 32)
        auto cols_of = [&](int n2base) {
            int mc = -1;
            if (n2base + lane < b) {
                const int c = __ldg(p.pairidx + ((size_t)colell * nmax + N) * nmax + n2base + lane);
                mc = (c >= p.col_lo && c < p.col_hi) ? c - p.col_lo : -1;
            }
            return mc;
        };
        int mycol = cols_of(N);
        {
            unsigned any = __ballot_sync(0xffffffffu, mycol >= 0);
            for (int n2b = N + 32; n2b < b && !any; n2b += 32) any = __ballot_sync(0xffffffffu, cols_of(n2b) >= 0);
            if (any == 0u) continue;
        }

        // ---- Z phase: Z[i][jt][e] = Z_N[8i+g][8jt+2t+e] ------------------------------------------------------
        double Z[AT][NT][2];
#pragma unroll
        for (int i = 0; i < AT; ++i)
#pragma unroll
            for (int jt = 0; jt < NT; ++jt) Z[i][jt][0] = Z[i][jt][1] = 0.0;
        if (!(p.dbg & 2)) {
            const double* glN = GL + N * S + 2 * t;
            const double* glA = Gl + g * S + 2 * t;
            const double* wsB = Ws + g * S + 2 * t;
#pragma unroll 2
            for (int kt = 0; kt < NT; ++kt) {
                const double2 sN = *reinterpret_cast<const double2*>(glN + 8 * kt);
                double av0[AT], av1[AT];
#pragma unroll
                for (int i = 0; i < AT; ++i) {
                    const double2 x = *reinterpret_cast<const double2*>(glA + i * 8 * S + 8 * kt);
                    av0[i] = x.x * sN.x;
                    av1[i] = x.y * sN.y;
                }
#pragma unroll
                for (int jt = 0; jt < NT; ++jt) {
                    const double2 w = *reinterpret_cast<const double2*>(wsB + jt * 8 * S + 8 * kt);
#pragma unroll
                    for (int i = 0; i < AT; ++i) {
                        dmma884(Z[i][jt], av0[i], w.x);
                        dmma884(Z[i][jt], av1[i], w.y);
                    }
                }
            }
        }

        // the factor c_L = (2L+1)/(4π) of the output rides on Z (once per N instead of once per stored element)
#pragma unroll
        for (int i = 0; i < AT; ++i)
#pragma unroll
            for (int jt = 0; jt < NT; ++jt) Z[i][jt][0] *= scale, Z[i][jt][1] *= scale;

        // ---- T phase over N' >= N -------------------------------------------------------------------------
        for (int N2 = N; N2 < b; ++N2) {
            if (N2 > N && ((N2 - N) & 31) == 0) mycol = cols_of(N2);   // next group of 32 columns
            const int col = __shfl_sync(0xffffffffu, mycol, (N2 - N) & 31);
            if (col < 0) continue;  // warp-uniform
            const bool offdiag = (N2 != N);
            const bool upper_only = rows_upper && !offdiag && !p.interchange;
            double acc[AT][AT][2];
            const double* glS = GL + N2 * S + 2 * t;
            const double* glB = Gl + g * S + 2 * t;
            if (p.dbg & 4) {
#pragma unroll
                for (int i = 0; i < AT; ++i)
#pragma unroll
                    for (int j = 0; j < AT; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
            } else if (upper_only) {
                regz_tile<AT, NT, S, true>(acc, Z, glS, glB);
            } else {
                regz_tile<AT, NT, S, false>(acc, Z, glS, glB);
            }
            if (p.dbg & 1) continue;
            const size_t coff = p.colbase ? (size_t)p.colbase[col] : (size_t)col * p.ldM;
            // ---- epilogue: tile -> per-warp shared staging -> rows of this l-block ----------------------------
            // (software-pipelining these row stores into the next tile's DMMA stream was tried: the extra live values spill
            //  at the 168-register cap and the block kernel slows from 2.73 to 2.95 ms at cfg4)
#pragma unroll
            for (int i = 0; i < AT; ++i)
#pragma unroll
                for (int j = 0; j < AT; ++j)
                    *reinterpret_cast<double2*>(Tw + (i * 8 + g) * TLD + j * 8 + 2 * t) =
                        make_double2(acc[i][j][0], acc[i][j][1]);
            __syncwarp();
            // U rows per lane per pass, all table and tile loads of a pass independent; U follows the rows that are left
            double* Mc = p.M + coff;
            int idx0 = lane;
            for (int left = nrows; left > 0;) {
                const int per = (left + 31) >> 5;
                if (per >= 7) {
                    regz_store_rows<8, TLD>(Tw, rowtab, idx0, nrows, Mc, offdiag, p.interchange);
                    idx0 += 256, left -= 256;
                } else if (per >= 5) {
                    regz_store_rows<6, TLD>(Tw, rowtab, idx0, nrows, Mc, offdiag, p.interchange);
                    idx0 += 192, left -= 192;
                } else if (per >= 3) {
                    regz_store_rows<4, TLD>(Tw, rowtab, idx0, nrows, Mc, offdiag, p.interchange);
                    idx0 += 128, left -= 128;
                } else {
                    regz_store_rows<2, TLD>(Tw, rowtab, idx0, nrows, Mc, offdiag, p.interchange);
                    idx0 += 64, left -= 64;
                }
            }
            __syncwarp();
        }
    }
}

template <int AT, int NT, int NW, int MINB, bool TMA>
__global__ void __launch_bounds__(NW * 32, MINB) cmix_regz_kernel(RegzArgs p, const __grid_constant__ RegzTmaps tm) {
    extern __shared__ __align__(128) double sm[];
    constexpr int AP = AT * 8;
    constexpr int K = NT * 8;          // padded radial length
    constexpr int S = K + 8;           // row stride ≡ 8 (mod 16): conflict-free 16-byte fragment loads
    constexpr int TLD = regz_tld(AP);
    constexpr int NTHR = NW * 32;
    const int tid = threadIdx.x, warp = tid >> 5;

    const int4 d0 = reinterpret_cast<const int4*>(p.blocks)[2 * blockIdx.x];
    const int4 d1 = reinterpret_cast<const int4*>(p.blocks)[2 * blockIdx.x + 1];
    const int rowell = d0.x, colell = d0.y, a = d0.z, b = d0.w;
    const int r0 = d1.x, nrows = d1.y, widx = d1.z;
    const int nrp = p.nrp, nmax = p.nmax;

    const int nmaxe = nmax + (nmax & 1);   // even row count: every tile starts 128-byte aligned (TMA destination)
    double* Gl = sm;                       // [AP][S]   G_ln[r], rows >= a zero
    double* GL = Gl + AP * S;              // [nmaxe][S] G_LN[r]
    double* Ws = GL + nmaxe * S;           // [K][S]    Ŵ_lL (symmetric)
    double* Ts = Ws + K * S;               // [NW][AP][TLD]
    int* counter = reinterpret_cast<int*>(Ts + NW * AP * TLD);
    int* rowtab = counter + 2;                                    // [nrows] orow | n << 22 | n' << 27

    // every row (n, n') of this l-block has n <= n' (true of every table ClnnModes builds): the N' = N tile, which is symmetric
    // and enters the output as is, then needs only its sub-tiles on and above the diagonal
    bool rows_upper = !p.force_rowtab;
    if (TMA) {
        // ---- stage operands with three tensor copies (one elected thread), completion on an mbarrier -------------
        __shared__ __align__(8) unsigned long long bar;
        if (tid == 0) mbar_init(&bar, 1);
        __syncthreads();
        if (tid == 0) {
            mbar_expect_tx(&bar, (unsigned)((AP + nmax + nrp) * S * sizeof(double)));
            tma_load_2d(Gl, &tm.gl, 0, rowell * nmax, &bar);
            tma_load_2d(GL, &tm.gL, 0, colell * nmax, &bar);
            tma_load_2d(Ws, &tm.w, 0, widx * nrp, &bar);
        }
        if (nrp < K) {   // rows of Ŵ beyond nrp (the columns beyond nrp are zero-filled by the copy engine)
            const int padc = K - nrp;
            for (int x = tid; x < padc * K; x += NTHR) Ws[(nrp + x / K) * S + x % K] = 0.0;
        }
        for (int x = tid; x < nrows; x += NTHR) {
            const int o = p.row_out[r0 + x], n = p.row_n[r0 + x], n2 = p.row_n2[r0 + x];
            rowtab[x] = (o < 0 ? kRegzNoRow : o) | (n << 22) | (n2 << 27);
            rows_upper = rows_upper && (n <= n2);
        }
        if (tid == 0) *counter = 0;
        mbar_wait(&bar, 0);
        rows_upper = __syncthreads_and(rows_upper) != 0;
    } else {
    // ---- stage operands: one round of 16-byte async copies, zero fill of the padding --------------------------
    const int cpr = nrp / 2;  // 16-byte chunks per row
    {
        const double* src = p.G + (size_t)rowell * nmax * nrp;
        for (int x = tid; x < a * cpr; x += NTHR) {
            const int n = x / cpr, c = x - n * cpr;
            cp_async16(Gl + n * S + 2 * c, src + (size_t)n * nrp + 2 * c);
        }
        src = p.G + (size_t)colell * nmax * nrp;
        for (int x = tid; x < b * cpr; x += NTHR) {
            const int n = x / cpr, c = x - n * cpr;
            cp_async16(GL + n * S + 2 * c, src + (size_t)n * nrp + 2 * c);
        }
        src = p.What + (size_t)widx * nrp * nrp;
        for (int x = tid; x < nrp * cpr; x += NTHR) {
            const int r = x / cpr, c = x - r * cpr;
            cp_async16(Ws + r * S + 2 * c, src + (size_t)r * nrp + 2 * c);
        }
    }
    for (int x = tid; x < (AP - a) * K; x += NTHR) {
        const int n = a + x / K, c = x % K;
        Gl[n * S + c] = 0.0;
    }
    if (nrp < K) {
        const int padc = K - nrp;
        for (int x = tid; x < a * padc; x += NTHR) Gl[(x / padc) * S + nrp + x % padc] = 0.0;
        for (int x = tid; x < b * padc; x += NTHR) GL[(x / padc) * S + nrp + x % padc] = 0.0;
        for (int x = tid; x < nrp * padc; x += NTHR) Ws[(x / padc) * S + nrp + x % padc] = 0.0;
        for (int x = tid; x < padc * K; x += NTHR) Ws[(nrp + x / K) * S + x % K] = 0.0;
    }
    for (int x = tid; x < nrows; x += NTHR) {
        const int o = p.row_out[r0 + x], n = p.row_n[r0 + x], n2 = p.row_n2[r0 + x];
        rowtab[x] = (o < 0 ? kRegzNoRow : o) | (n << 22) | (n2 << 27);
        rows_upper = rows_upper && (n <= n2);
    }
    if (tid == 0) *counter = 0;
    cp_async_wait_all();
    rows_upper = __syncthreads_and(rows_upper) != 0;
    }
    regz_warp_work<AT, NT>(p, Gl, GL, Ws, Ts + warp * AP * TLD, rowtab, counter, colell, a, b, nrows, rows_upper);
}

// ---- Persistent variant (default when the tensor maps encode and two operand stages fit in shared memory) ----------------
// One CTA per SM walks a global queue of (l, L) blocks (heaviest first).  The operands of the NEXT block are staged by TMA
// into the second buffer while the warps still work on the current one, and a warp that finds the current block's N
// counter exhausted moves straight on to the next block: neither the staging latency nor the end-of-block imbalance
// (N is a coarse work item: 9 % of the warp time at cfg4 with one block per CTA) idles a warp.  The last warp to leave a
// stage refills it (row table by the warp, tiles by TMA); completion is an mbarrier per stage.
struct RegzStage {
    int desc[8];       // rowell, colell, a, b, r0, nrows, what_index, valid
    int counter;       // next N of the block
    int done;          // warps that have left the stage
    int rows_upper;
    int pad;
};

__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
    const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(a) : "memory");
}

template <int AT, int NT, int NW>
__global__ void __launch_bounds__(NW * 32, 1) cmix_regz_persist_kernel(RegzArgs p, const __grid_constant__ RegzTmaps tm,
                                                                       int nblocks, int max_rows, int* queue) {
    extern __shared__ __align__(128) double sm[];
    constexpr int AP = AT * 8;
    constexpr int K = NT * 8;
    constexpr int S = K + 8;
    constexpr int TLD = regz_tld(AP);
    constexpr int NTHR = NW * 32;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nrp = p.nrp, nmax = p.nmax;
    const int nmaxe = nmax + (nmax & 1);
    const int stage_doubles = (AP + nmaxe + K) * S;          // Gl | GL | Ws, every tile 128-byte aligned
    double* Ts = sm + 2 * stage_doubles;                      // [NW][AP][TLD]
    int* rowtabs = reinterpret_cast<int*>(Ts + NW * AP * TLD);   // [2][max_rows]
    RegzStage* st = reinterpret_cast<RegzStage*>(rowtabs + 2 * max_rows + ((2 * max_rows) & 1));
    __shared__ __align__(8) unsigned long long full[2];

    auto tiles = [&](int s) { return sm + s * stage_doubles; };
    // refill stage s with the next block of the queue (one warp; every other warp has left the stage)
    auto produce = [&](int s) {
        int idx = 0;
        if (lane == 0) idx = atomicAdd(queue, 1);
        idx = __shfl_sync(0xffffffffu, idx, 0);
        if (idx >= nblocks) {
            if (lane == 0) {
                st[s].desc[7] = 0;
                st[s].done = 0;
                __threadfence_block();
                mbar_arrive(&full[s]);
            }
            return;
        }
        const int4 d0 = reinterpret_cast<const int4*>(p.blocks)[2 * idx];
        const int4 d1 = reinterpret_cast<const int4*>(p.blocks)[2 * idx + 1];
        const int r0 = d1.x, nrows = d1.y;
        int* rowtab = rowtabs + s * max_rows;
        bool up = !p.force_rowtab;
        for (int x = lane; x < nrows; x += 32) {
            const int o = p.row_out[r0 + x], n = p.row_n[r0 + x], n2 = p.row_n2[r0 + x];
            rowtab[x] = (o < 0 ? kRegzNoRow : o) | (n << 22) | (n2 << 27);
            up = up && (n <= n2);
        }
        up = __all_sync(0xffffffffu, up);
        __syncwarp();
        if (lane == 0) {
            st[s].desc[0] = d0.x, st[s].desc[1] = d0.y, st[s].desc[2] = d0.z, st[s].desc[3] = d0.w;
            st[s].desc[4] = d1.x, st[s].desc[5] = d1.y, st[s].desc[6] = d1.z, st[s].desc[7] = 1;
            st[s].counter = 0;
            st[s].done = 0;
            st[s].rows_upper = up ? 1 : 0;
            __threadfence_block();
            asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
            double* T0 = tiles(s);
            mbar_expect_tx(&full[s], (unsigned)((AP + nmax + nrp) * S * sizeof(double)));
            tma_load_2d(T0, &tm.gl, 0, d0.x * nmax, &full[s]);
            tma_load_2d(T0 + AP * S, &tm.gL, 0, d0.y * nmax, &full[s]);
            tma_load_2d(T0 + (AP + nmaxe) * S, &tm.w, 0, d1.z * nrp, &full[s]);
        }
    };

    if (tid == 0) {
        mbar_init(&full[0], 1);
        mbar_init(&full[1], 1);
    }
    // rows of Ŵ beyond nrp are never written by the copy engine (box = nrp rows): zero them once in both stages; so is the odd
    // padding row of GL
    for (int s = 0; s < 2; ++s) {
        double* Ws = tiles(s) + (AP + nmaxe) * S;
        for (int x = tid; x < (K - nrp) * S; x += NTHR) Ws[nrp * S + x] = 0.0;
        double* GLp = tiles(s) + AP * S;
        for (int x = tid; x < (nmaxe - nmax) * S; x += NTHR) GLp[nmax * S + x] = 0.0;
    }
    __syncthreads();
    if (warp == 0) {
        produce(0);
        produce(1);
    }
    for (int k = 0;; ++k) {
        const int s = k & 1;
        mbar_wait(&full[s], (unsigned)((k >> 1) & 1));
        if (!st[s].desc[7]) break;
        const int colell = st[s].desc[1], a = st[s].desc[2], b = st[s].desc[3], nrows = st[s].desc[5];
        const bool rows_upper = st[s].rows_upper != 0;
        double* T0 = tiles(s);
        regz_warp_work<AT, NT>(p, T0, T0 + AP * S, T0 + (AP + nmaxe) * S, Ts + warp * AP * TLD, rowtabs + s * max_rows,
                               &st[s].counter, colell, a, b, nrows, rows_upper);
        __syncwarp();
        int last = 0;
        if (lane == 0) {
            __threadfence_block();
            last = (atomicAdd(&st[s].done, 1) == NW - 1) ? 1 : 0;
            __threadfence_block();
        }
        last = __shfl_sync(0xffffffffu, last, 0);
        if (last) produce(s);
    }
}

// =============================================================================================
// Mirror fill: for the output columns c in [c0, c1) (an l-aligned range) write every element below the block diagonal,
//   M[r, c] = M[c, r] · f_c / f_r   for l(r) > l(c),   f = (div2Lp1 ? 1 : 2l+1) · (interchange ? 1 : 1 + [n≠n']),
// from the directly formed element M[c, r] (block (l(c), l(r)), L >= l).  64 x 64 tiles through shared memory: both
// the read (along the source column) and the write (along the destination column) are coalesced.
// es[o] = l | [n≠n'] << 30 per output index.  `sorted` = l is non-decreasing in output order, so tiles entirely
// above the diagonal can be skipped without looking at the table.
constexpr int kFillT = 64;  // tile edge: 16 independent 8-byte loads in flight per thread

template <bool VEC>
__global__ void __launch_bounds__(256) cmix_mirror_fill_kernel(double* __restrict__ M, long long ld, int n, int c0,
                                                               int c1, int rbase, const int* __restrict__ es,
                                                               int div2Lp1, int interchange, int sorted, int tri) {
    __shared__ double tile[kFillT][kFillT + 1];
    __shared__ int esr[kFillT], esc[kFillT];
    int bx = blockIdx.x, by = blockIdx.y;
    if (tri) {   // 1-D grid over the tiles on and below the tile diagonal only (by >= bx): no empty CTAs
        const unsigned k = blockIdx.x;
        by = (int)((sqrtf(8.0f * (float)k + 1.0f) - 1.0f) * 0.5f);
        while ((unsigned)(by + 1) * (unsigned)(by + 2) / 2 <= k) ++by;
        while ((unsigned)by * (unsigned)(by + 1) / 2 > k) --by;
        bx = (int)(k - (unsigned)by * (unsigned)(by + 1) / 2);
    }
    const int cb = c0 + bx * kFillT, rb = rbase + by * kFillT;
    if (sorted && rb + kFillT - 1 <= cb) return;
    const int x = threadIdx.x & 31, y0 = threadIdx.x >> 5;
    if (threadIdx.x < kFillT) esr[threadIdx.x] = (rb + threadIdx.x < n) ? es[rb + threadIdx.x] : -1;
    else if (threadIdx.x < 2 * kFillT)
        esc[threadIdx.x - kFillT] = (cb + threadIdx.x - kFillT < c1) ? es[cb + threadIdx.x - kFillT] : -1;
    __syncthreads();
    // whole tile strictly above / on the block diagonal: nothing to fill
    int lrmax = -1, lcmin = 1 << 30;
    if (sorted) {
        lrmax = esr[kFillT - 1] >= 0 ? (esr[kFillT - 1] & 0x3fffffff) : (1 << 29);
        lcmin = esc[0] & 0x3fffffff;
        if (lrmax <= lcmin) return;
    }
    // source elements (rows cb+2x, cb+2x+1, column rb+y) -> tile[y][.]; 16-byte accesses when VEC (ld, c0 even and M
    // 16-byte aligned).  l is non-decreasing along rows and columns, which orders the two predicates of a pair.
    constexpr int NK = kFillT / 8;
    const int x2 = 2 * x;
    const int ec0 = esc[x2], ec1 = esc[x2 + 1];
    const int lc0 = ec0 & 0x3fffffff, lc1 = ec1 & 0x3fffffff;
    bool any = false;
#pragma unroll
    for (int k = 0; k < NK; ++k) {
        const int y = y0 + 8 * k;
        const int er = esr[y];
        const int lr = er & 0x3fffffff;
        const bool n0 = (er >= 0 && ec0 >= 0 && lr > lc0), n1 = (er >= 0 && ec1 >= 0 && lr > lc1);
        const double* srcp = M + (size_t)(rb + y) * ld + cb + x2;
        if (n0 && n1) {
            if (VEC) {
                const double2 v = *reinterpret_cast<const double2*>(srcp);
                tile[y][x2] = v.x;
                tile[y][x2 + 1] = v.y;
            } else {
                tile[y][x2] = srcp[0];
                tile[y][x2 + 1] = srcp[1];
            }
        } else if (n0) {
            tile[y][x2] = srcp[0];
        } else if (n1) {
            tile[y][x2 + 1] = srcp[1];
        }
        any = any || n0 || n1;
    }
    if (!__syncthreads_or(any)) return;
    // destination elements (rows rb+2x, rb+2x+1, column cb+y) = tile[.][y] f_c / f_r
    const int er0 = esr[x2], er1 = esr[x2 + 1];
    const int lr0 = er0 & 0x3fffffff, lr1 = er1 & 0x3fffffff;
    const double ifr0 = 1.0 / ((div2Lp1 ? 1.0 : 2.0 * lr0 + 1.0) * ((!interchange && (er0 >> 30)) ? 2.0 : 1.0));
    const double ifr1 = 1.0 / ((div2Lp1 ? 1.0 : 2.0 * lr1 + 1.0) * ((!interchange && (er1 >> 30)) ? 2.0 : 1.0));
#pragma unroll
    for (int k = 0; k < NK; ++k) {
        const int y = y0 + 8 * k;
        const int e2 = esc[y];
        if (e2 < 0) continue;
        const int lc = e2 & 0x3fffffff;
        const double fc = (div2Lp1 ? 1.0 : 2.0 * lc + 1.0) * ((!interchange && (e2 >> 30)) ? 2.0 : 1.0);
        const bool d0 = (er0 >= 0 && lr0 > lc), d1 = (er1 >= 0 && lr1 > lc);
        double* dst = M + (size_t)(cb + y) * ld + rb + x2;
        if (d0 && d1) {
            const double w0 = tile[x2][y] * (fc * ifr0), w1 = tile[x2 + 1][y] * (fc * ifr1);
            if (VEC) *reinterpret_cast<double2*>(dst) = make_double2(w0, w1);
            else { dst[0] = w0; dst[1] = w1; }
        } else if (d0) {
            dst[0] = tile[x2][y] * (fc * ifr0);
        } else if (d1) {
            dst[1] = tile[x2 + 1][y] * (fc * ifr1);
        }
    }
}

int cmix_mirror_fill(CmixPlan* p, int64_t c0, int64_t c1, int div2Lp1, int interchange, double* d_M, int64_t ldM,
                     cudaStream_t stream) {
    if (c1 <= c0) return 0;
    const int n = (int)p->nout;
    const int rbase = p->ell_sorted ? (int)(c0 / kFillT) * kFillT : 0;
    dim3 grid((unsigned)ceil_div(c1 - c0, kFillT), (unsigned)ceil_div(n - rbase, kFillT));
    // whole sorted matrix: the tiles above the tile diagonal hold nothing to fill, so only the others are launched
    const int tri = (p->ell_sorted && c0 == 0 && c1 == n && !getenv("SFB_FILL_2D")) ? 1 : 0;
    if (tri) grid = dim3((unsigned)((size_t)grid.y * (grid.y + 1) / 2), 1);
    const bool vec = p->ell_sorted && (ldM % 2 == 0) && (c0 % 2 == 0) && (reinterpret_cast<uintptr_t>(d_M) % 16 == 0);
    if (vec)
        cmix_mirror_fill_kernel<true><<<grid, 256, 0, stream>>>(d_M, ldM, n, (int)c0, (int)c1, rbase, p->d_es.p, div2Lp1,
                                                                interchange, 1, tri);
    else
        cmix_mirror_fill_kernel<false><<<grid, 256, 0, stream>>>(d_M, ldM, n, (int)c0, (int)c1, rbase, p->d_es.p,
                                                                 div2Lp1, interchange, p->ell_sorted ? 1 : 0, tri);
    SFB_CUDA_OK(cudaGetLastError());
    return 0;
}

// =============================================================================================
// Upper-packed storage (multi-GPU exchange format): output column j keeps only the rows of the blocks with
// l <= l(j), i.e. rows [0, rend(j)), at element offset colbase[j]; a rank's column range is one contiguous slab
// (half the bytes of the full matrix cross NVLink).  cmix_unpack_mirror_kernel expands it into the full
// column-major matrix: M[r, j] = P[colbase[j] + r] for r < rend(j), and below the block diagonal
// M[j, r] = M[r, j] · f_r / f_j for l(r) < l(j)  (same identity as cmix_mirror_fill_kernel).
// With several ranks the packed buffer of the column's OWNER is read directly (peer-mapped memory, NVLink loads): the
// all-gather of the packed slabs is fused into this kernel and overlaps its HBM writes.
struct PackedSrc {
    const double* base[8];  // packed buffer of rank g (own buffer included), peer-mapped
    int col_start[9];       // rank g owns the columns [col_start[g], col_start[g+1])
    int nranks;
    const int* jt_order;    // column-tile visiting order (null: natural).  Every rank starts with a different owner's
                            // columns and proceeds cyclically, so that at any time the readers pull from different
                            // GPUs (no NVLink egress hot spot); its own (local) tiles are interleaved evenly
};

template <bool VEC>
__global__ void __launch_bounds__(256) cmix_unpack_mirror_kernel(PackedSrc src,
                                                                 const long long* __restrict__ colbase,
                                                                 double* __restrict__ M, long long ld, int n,
                                                                 const int* __restrict__ es, int div2Lp1,
                                                                 int interchange) {
    __shared__ double tile[kFillT][kFillT + 1];
    __shared__ int esr[kFillT], esc[kFillT];
    __shared__ const double* cptr[kFillT];
    // row tiles fastest: one wave of CTAs works on a few column tiles, i.e. on ONE owner's columns
    const int nt = (n + kFillT - 1) / kFillT;
    const int jq = (int)(blockIdx.x / nt), rt = (int)(blockIdx.x % nt);
    const int jt = src.jt_order ? src.jt_order[jq] : jq;
    const int jb = jt * kFillT, rb = rt * kFillT;
    const int x = threadIdx.x & 31, y0 = threadIdx.x >> 5;
    if (threadIdx.x < kFillT) esr[threadIdx.x] = (rb + threadIdx.x < n) ? es[rb + threadIdx.x] : -1;
    else if (threadIdx.x < 2 * kFillT) {
        const int j = jb + threadIdx.x - kFillT;
        esc[threadIdx.x - kFillT] = (j < n) ? es[j] : -1;
        int g = 0;
        while (g + 1 < src.nranks && j >= src.col_start[g + 1]) ++g;
        cptr[threadIdx.x - kFillT] = src.base[g] + ((j < n) ? colbase[j] : 0);
    }
    __syncthreads();
    {   // l is sorted: a tile whose smallest row l exceeds its largest column l holds no source element
        const int jlast = min(kFillT, n - jb) - 1;
        if ((esr[0] & 0x3fffffff) > (esc[jlast] & 0x3fffffff)) return;
    }
    // source elements (rows rb+2x, rb+2x+1, column jb+y): direct write, and into tile[y][.] for the mirror image.
    // 16-byte accesses (columns start on even offsets in packed storage, VEC: ld even and M 16-byte aligned); all
    // eight loads of a thread are issued before the first store (they may cross NVLink: keep them in flight).
    constexpr int NK = kFillT / 8;
    const int x2 = 2 * x;
    const int er0 = esr[x2], er1 = esr[x2 + 1];
    double2 v[NK];
    unsigned ok0 = 0, ok1 = 0;
#pragma unroll
    for (int k = 0; k < NK; ++k) {
        const int y = y0 + 8 * k;
        const int ec = esc[y];
        const bool a0 = (er0 >= 0 && ec >= 0 && (er0 & 0x3fffffff) <= (ec & 0x3fffffff));
        const bool a1 = (er1 >= 0 && ec >= 0 && (er1 & 0x3fffffff) <= (ec & 0x3fffffff));
        v[k] = make_double2(0.0, 0.0);
        if (a0) v[k] = *reinterpret_cast<const double2*>(cptr[y] + rb + x2);   // row rb+2x+1 < padded rend(j)
        ok0 |= (a0 ? 1u : 0u) << k;
        ok1 |= (a1 ? 1u : 0u) << k;   // l sorted: a1 implies a0
    }
#pragma unroll
    for (int k = 0; k < NK; ++k) {
        const int y = y0 + 8 * k;
        double* dst = M + (size_t)(jb + y) * ld + rb + x2;
        if ((ok1 >> k) & 1u) {
            if (VEC) *reinterpret_cast<double2*>(dst) = v[k];
            else { dst[0] = v[k].x; dst[1] = v[k].y; }
            tile[y][x2] = v[k].x;
            tile[y][x2 + 1] = v[k].y;
        } else if ((ok0 >> k) & 1u) {
            dst[0] = v[k].x;
            tile[y][x2] = v[k].x;
        }
    }
    __syncthreads();
    // mirror image: destination (rows jb+2x, jb+2x+1, column rb+y) = tile[.][y] f_r / f_j  for l(r) < l(j)
    const int ej0 = esc[x2], ej1 = esc[x2 + 1];
    const int lj0 = ej0 & 0x3fffffff, lj1 = ej1 & 0x3fffffff;
    const double ifj0 = 1.0 / ((div2Lp1 ? 1.0 : 2.0 * lj0 + 1.0) * ((!interchange && (ej0 >> 30)) ? 2.0 : 1.0));
    const double ifj1 = 1.0 / ((div2Lp1 ? 1.0 : 2.0 * lj1 + 1.0) * ((!interchange && (ej1 >> 30)) ? 2.0 : 1.0));
#pragma unroll
    for (int k = 0; k < NK; ++k) {
        const int y = y0 + 8 * k;
        const int e2 = esr[y];
        if (e2 < 0) continue;
        const int lr = e2 & 0x3fffffff;
        const double fr = (div2Lp1 ? 1.0 : 2.0 * lr + 1.0) * ((!interchange && (e2 >> 30)) ? 2.0 : 1.0);
        const bool b0 = (ej0 >= 0 && lr < lj0), b1 = (ej1 >= 0 && lr < lj1);
        double* dst = M + (size_t)(rb + y) * ld + jb + x2;
        if (b0 && b1) {
            const double w0 = tile[x2][y] * (fr * ifj0), w1 = tile[x2 + 1][y] * (fr * ifj1);
            if (VEC) *reinterpret_cast<double2*>(dst) = make_double2(w0, w1);
            else { dst[0] = w0; dst[1] = w1; }
        } else if (b0) {
            dst[0] = tile[x2][y] * (fr * ifj0);
        } else if (b1) {
            dst[1] = tile[x2 + 1][y] * (fr * ifj1);
        }
    }
}

int cmix_unpack_mirror(CmixPlan* p, const double* const* bases, const int64_t* col_bounds, int nranks, int my_rank,
                       int div2Lp1, int interchange, double* d_M, int64_t ldM, cudaStream_t stream) {
    SFB_REQUIRE(p && bases && d_M && nranks >= 1 && nranks <= 8, "cmix_unpack_mirror: bad arguments");
    PackedSrc src;
    for (int g = 0; g < 8; ++g) src.base[g] = (g < nranks) ? bases[g] : nullptr;
    for (int g = 0; g <= 8; ++g) src.col_start[g] = (int)((g <= nranks && col_bounds) ? col_bounds[std::min(g, nranks)] : p->nout);
    if (!col_bounds) src.col_start[0] = 0;
    src.nranks = nranks;
    src.jt_order = nullptr;
    const int ntile = (int)ceil_div(p->nout, kFillT);
    if (nranks > 1 && my_rank >= 0 && my_rank < nranks) {
        std::vector<int> key(src.col_start, src.col_start + 9);
        key.push_back(nranks);
        key.push_back(my_rank);
        if (key != p->h_jt_key) {
            // a tile belongs to the owner of its first column
            std::vector<int> remote, local, order;
            const int t0 = (int)ceil_div(src.col_start[(my_rank + 1) % nranks], kFillT);
            for (int q = 0; q < ntile; ++q) {
                const int jt = (t0 + q) % ntile;
                const int j = jt * kFillT;
                (j >= src.col_start[my_rank] && j < src.col_start[my_rank + 1] ? local : remote).push_back(jt);
            }
            size_t il = 0, ir = 0;
            for (int q = 0; q < ntile; ++q) {
                const bool take_local = ((q + 1) * local.size()) / ntile > (q * local.size()) / ntile;
                if ((take_local && il < local.size()) || ir >= remote.size()) order.push_back(local[il++]);
                else order.push_back(remote[ir++]);
            }
            SFB_TRY(p->d_jt_order.alloc(ntile));
            SFB_CUDA_OK(cudaMemcpyAsync(p->d_jt_order.p, order.data(), ntile * sizeof(int), cudaMemcpyHostToDevice, stream));
            SFB_CUDA_OK(cudaStreamSynchronize(stream));
            p->h_jt_key = key;
        }
        src.jt_order = p->d_jt_order.p;
    }
    for (int g = 0; g < nranks; ++g) SFB_REQUIRE(bases[g], "cmix_unpack_mirror: null packed buffer");
    SFB_REQUIRE(p->ell_sorted, "cmix_unpack_mirror: the lnn table must be sorted by l");
    SFB_REQUIRE(ldM >= p->nout, "cmix_unpack_mirror: ldM smaller than the matrix");
    const int n = (int)p->nout;
    const unsigned nt = (unsigned)ceil_div(n, kFillT);
    dim3 grid(nt * nt);
    const bool vec = (ldM % 2 == 0) && (reinterpret_cast<uintptr_t>(d_M) % 16 == 0);
    if (vec)
        cmix_unpack_mirror_kernel<true><<<grid, 256, 0, stream>>>(src, p->d_colbase.p, d_M, ldM, n, p->d_es.p, div2Lp1,
                                                                  interchange);
    else
        cmix_unpack_mirror_kernel<false><<<grid, 256, 0, stream>>>(src, p->d_colbase.p, d_M, ldM, n, p->d_es.p, div2Lp1,
                                                                   interchange);
    SFB_CUDA_OK(cudaGetLastError());
    return 0;
}

// =============================================================================================
// Row mirror of an upper-packed column range ("L-shaped" shards, the multi-GPU output without redundant flops): the owner
// of the columns [j_lo, j_hi) (whole L-blocks, blocks l <= L in packed storage) also owns the ROWS [j_lo, j_hi) of the
// part below the block diagonal, which it forms locally from its own packed columns,
//   M[j, r] = M[r, j] · f_r / f_j   for l(r) < l(j),   j in [j_lo, j_hi)
// (same identity as cmix_mirror_fill_kernel), into the compact row slab R[r * ldR + (j - j_lo)], r in [0, j_hi).
// Every element of M then lives on exactly one device: (i, j) with max(l_i, l_j) in the device's L range.
__global__ void __launch_bounds__(256) cmix_mirror_rows_kernel(const double* __restrict__ P,
                                                               const long long* __restrict__ colbase,
                                                               double* __restrict__ R, long long ldR, int j_lo, int j_hi,
                                                               const int* __restrict__ es, int div2Lp1, int interchange) {
    __shared__ double tile[kFillT][kFillT + 1];
    __shared__ int esr[kFillT], esc[kFillT];
    __shared__ long long cbase[kFillT];
    const int jb = j_lo + blockIdx.x * kFillT, rb = blockIdx.y * kFillT;
    const int x = threadIdx.x & 31, y0 = threadIdx.x >> 5;
    if (threadIdx.x < kFillT) esr[threadIdx.x] = (rb + threadIdx.x < j_hi) ? es[rb + threadIdx.x] : -1;
    else if (threadIdx.x < 2 * kFillT) {
        const int j = jb + threadIdx.x - kFillT;
        esc[threadIdx.x - kFillT] = (j < j_hi) ? es[j] : -1;
        cbase[threadIdx.x - kFillT] = (j < j_hi) ? colbase[j] : 0;
    }
    __syncthreads();
    {   // l is sorted: nothing to do unless the tile's smallest row l is below its largest column l
        const int jlast = min(kFillT, j_hi - jb) - 1;
        if ((esr[0] & 0x3fffffff) >= (esc[jlast] & 0x3fffffff)) return;
    }
    constexpr int NK = kFillT / 8;
    const int x2 = 2 * x;
    const int er0 = esr[x2], er1 = esr[x2 + 1];
    // source elements (rows rb+2x, rb+2x+1 of packed column jb+y), rows with l(r) < l(j) only
#pragma unroll
    for (int k = 0; k < NK; ++k) {
        const int y = y0 + 8 * k;
        const int ec = esc[y];
        const bool a0 = (er0 >= 0 && ec >= 0 && (er0 & 0x3fffffff) < (ec & 0x3fffffff));
        const bool a1 = (er1 >= 0 && ec >= 0 && (er1 & 0x3fffffff) < (ec & 0x3fffffff));
        const double* src = P + cbase[y] + rb + x2;
        if (a1) {   // l sorted: a1 implies a0; packed columns start 32-byte aligned and rb + x2 is even
            const double2 v = *reinterpret_cast<const double2*>(src);
            tile[y][x2] = v.x;
            tile[y][x2 + 1] = v.y;
        } else if (a0) {
            tile[y][x2] = src[0];
        }
    }
    __syncthreads();
    // destination: R[(rb+y) * ldR + (jb + 2x - j_lo)], rows j = jb+2x, jb+2x+1 of the slab
    const int ej0 = esc[x2], ej1 = esc[x2 + 1];
    const int lj0 = ej0 & 0x3fffffff, lj1 = ej1 & 0x3fffffff;
    const double ifj0 = 1.0 / ((div2Lp1 ? 1.0 : 2.0 * lj0 + 1.0) * ((!interchange && (ej0 >> 30)) ? 2.0 : 1.0));
    const double ifj1 = 1.0 / ((div2Lp1 ? 1.0 : 2.0 * lj1 + 1.0) * ((!interchange && (ej1 >> 30)) ? 2.0 : 1.0));
#pragma unroll
    for (int k = 0; k < NK; ++k) {
        const int y = y0 + 8 * k;
        const int e2 = esr[y];
        if (e2 < 0) continue;
        const int lr = e2 & 0x3fffffff;
        const double fr = (div2Lp1 ? 1.0 : 2.0 * lr + 1.0) * ((!interchange && (e2 >> 30)) ? 2.0 : 1.0);
        const bool b0 = (ej0 >= 0 && lr < lj0), b1 = (ej1 >= 0 && lr < lj1);
        double* dst = R + (size_t)(rb + y) * ldR + (jb + x2 - j_lo);
        if (b0) dst[0] = tile[x2][y] * (fr * ifj0);
        if (b1) dst[1] = tile[x2 + 1][y] * (fr * ifj1);
    }
}

int cmix_mirror_rows(CmixPlan* p, const double* d_packed, int64_t j_lo, int64_t j_hi, int div2Lp1, int interchange,
                     double* d_rows, int64_t ldR, cudaStream_t stream) {
    SFB_REQUIRE(p && d_packed && d_rows, "cmix_mirror_rows: null pointer");
    SFB_REQUIRE(p->ell_sorted && (int64_t)p->h_colbase.size() == p->nout + 1,
                "cmix_mirror_rows: upper-packed storage needs an lnn table sorted by l");
    SFB_REQUIRE(0 <= j_lo && j_lo <= j_hi && j_hi <= p->nout && ldR >= j_hi - j_lo, "cmix_mirror_rows: bad range");
    if (j_hi == j_lo) return 0;
    dim3 grid((unsigned)ceil_div(j_hi - j_lo, kFillT), (unsigned)ceil_div(j_hi, kFillT));
    cmix_mirror_rows_kernel<<<grid, 256, 0, stream>>>(d_packed, p->d_colbase.p, d_rows, ldR, (int)j_lo, (int)j_hi, p->d_es.p,
                                                      div2Lp1, interchange);
    SFB_CUDA_OK(cudaGetLastError());
    return 0;
}

template <int AT, int NT, int NW>
static size_t regz_smem_bytes(int nmax, int max_rows) {
    constexpr int AP = AT * 8, K = NT * 8, S = K + 8;
    const int nmaxe = nmax + (nmax & 1);
    return sizeof(double) * ((size_t)AP * S + (size_t)nmaxe * S + (size_t)K * S + (size_t)NW * AP * regz_tld(AP)) +
           sizeof(int) * ((size_t)max_rows + 4);
}

// tensor maps for the TMA staging variant; `ok` false (and no error) when the driver entry point is unavailable
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode_tiled() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(f);
    }
    return fn;
}
static bool encode_2d(CUtensorMap* m, const double* base, int inner, long long rows, int box_inner, int box_rows) {
    EncodeTiledFn fn = get_encode_tiled();
    if (!fn || rows < 1 || box_rows < 1 || box_rows > 256 || box_inner > 256) return false;
    const cuuint64_t gdim[2] = {(cuuint64_t)inner, (cuuint64_t)rows};
    const cuuint64_t gstr[1] = {(cuuint64_t)inner * sizeof(double)};
    const cuuint32_t box[2] = {(cuuint32_t)box_inner, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    return fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(base), gdim, gstr, box, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

struct RegzLaunchCtx {
    const double* G;
    long long g_rows;      // (lmax+1) * nmax
    const double* What;
    long long w_rows;      // rows of the Ŵ chunk buffer
    bool want_tma;
    bool want_persist;
    int* queue;            // zeroed head of this launch's block queue (persistent kernel)
    int num_sms;
};

template <int AT, int NT, int NW>
static size_t regz_persist_smem_bytes(int nmax, int max_rows) {
    constexpr int AP = AT * 8, K = NT * 8, S = K + 8;
    const int nmaxe = nmax + (nmax & 1);
    return sizeof(double) * (2 * (size_t)(AP + nmaxe + K) * S + (size_t)NW * AP * regz_tld(AP)) +
           sizeof(int) * (2 * (size_t)max_rows + 2) + 2 * sizeof(RegzStage) + 64;
}

template <int AT, int NT, int NW, int MINB, int NWP>
static int launch_regz(const RegzArgs& args, const RegzLaunchCtx& ctx, int nblocks, int nmax, int max_rows,
                       cudaStream_t stream) {
    constexpr int AP = AT * 8, S = NT * 8 + 8;
    const size_t smem = regz_smem_bytes<AT, NT, NW>(nmax, max_rows);
    // dynamic + static shared memory of a block must fit the 227 KB opt-in limit: leave 1 KB for the static part (mbarriers)
    constexpr size_t kSmemLimit = 227 * 1024 - 1024;
    SFB_REQUIRE(smem <= kSmemLimit, "cmix_regz_kernel: shared memory footprint exceeds 226 KB");
    RegzTmaps tm;
    std::memset(&tm, 0, sizeof(tm));
    bool tma = ctx.want_tma;
    if (tma)
        tma = encode_2d(&tm.gl, ctx.G, args.nrp, ctx.g_rows, S, AP) && encode_2d(&tm.gL, ctx.G, args.nrp, ctx.g_rows, S, nmax) &&
              encode_2d(&tm.w, ctx.What, args.nrp, ctx.w_rows, S, args.nrp);
    const size_t psmem = regz_persist_smem_bytes<AT, NT, NWP>(nmax, max_rows);
    if (tma && ctx.want_persist && ctx.queue && psmem <= kSmemLimit) {
        const int grid = std::min(nblocks, ctx.num_sms);
        SFB_CUDA_OK(cudaFuncSetAttribute(cmix_regz_persist_kernel<AT, NT, NWP>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)psmem));
        cmix_regz_persist_kernel<AT, NT, NWP><<<grid, NWP * 32, psmem, stream>>>(args, tm, nblocks, max_rows, ctx.queue);
        SFB_CUDA_OK(cudaGetLastError());
        return 0;
    }
    if (tma) {
        SFB_CUDA_OK(cudaFuncSetAttribute(cmix_regz_kernel<AT, NT, NW, MINB, true>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        cmix_regz_kernel<AT, NT, NW, MINB, true><<<nblocks, NW * 32, smem, stream>>>(args, tm);
    } else {
        SFB_REQUIRE(!(ctx.want_tma && getenv("SFB_REGZ_TMA_STRICT")), "cmix_regz_kernel: tensor maps could not be encoded");
        SFB_CUDA_OK(cudaFuncSetAttribute(cmix_regz_kernel<AT, NT, NW, MINB, false>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        cmix_regz_kernel<AT, NT, NW, MINB, false><<<nblocks, NW * 32, smem, stream>>>(args, tm);
    }
    SFB_CUDA_OK(cudaGetLastError());
    return 0;
}

template <int NT>
static int launch_regz_at(int AT, const RegzArgs& args, const RegzLaunchCtx& ctx, int nblocks, int nmax, int max_rows,
                          cudaStream_t stream) {
    switch (AT) {
        case 1: return launch_regz<1, NT, 4, 4, 12>(args, ctx, nblocks, nmax, max_rows, stream);
        case 2: return launch_regz<2, NT, 6, 2, 12>(args, ctx, nblocks, nmax, max_rows, stream);
        case 3: return launch_regz<3, NT, 6, 2, 12>(args, ctx, nblocks, nmax, max_rows, stream);
        case 4: return launch_regz<4, NT, 8, 1, 8>(args, ctx, nblocks, nmax, max_rows, stream);
        default: break;
    }
    set_error("cmix: nmax_l > 32 is not supported by this build");
    return 2;
}

bool cmix_regz_eligible(const CmixPlan* p, bool sym, int npeers) {
    if (getenv("SFB_CMIX_OLD")) return false;
    // nmax > 32: rows run as virtual blocks of <= 32 functions (cmix_plan_create); the column side only needs its G_L tile
    // (nmax rows) in shared memory next to the other operands
    return sym && npeers == 0 && p->nrp <= 64 && p->amax_tiles <= 4 && p->nmax <= 144 && p->nout < kRegzNoRow;
}

// Launch the register-Z kernel over the listed (row-ell, col-ell) blocks of one Ŵ chunk.
// `blocks` holds 8 ints per block (see RegzArgs); they are grouped by the row side's tile count here.
int cmix_regz_run(CmixPlan* p, const std::vector<int>& blocks, const double* d_What, int div2Lp1, int interchange,
                  int64_t col_lo, int64_t col_hi, double* d_M, int64_t ldM, const long long* d_colbase,
                  cudaStream_t stream, double* flops, int* launches) {
    const int nb = (int)(blocks.size() / 8);
    if (nb == 0) return 0;
    const int NT = (p->nrp <= 32) ? 4 : 8;
    // order: tile class (descending), then cost (descending) so the heavy CTAs start first
    std::vector<int> order(nb);
    for (int i = 0; i < nb; ++i) order[i] = i;
    auto at_of = [&](int i) { return (blocks[8 * i + 2] + 7) / 8; };
    auto cost_of = [&](int i) {
        const double ap = 8.0 * at_of(i), b = blocks[8 * i + 3];
        return ap * b + ap * ap * b * (b + 1) / (2.0 * 8 * NT);
    };
    std::stable_sort(order.begin(), order.end(), [&](int x, int y) {
        if (at_of(x) != at_of(y)) return at_of(x) > at_of(y);
        return cost_of(x) > cost_of(y);
    });
    std::vector<int> sorted((size_t)nb * 8);
    for (int i = 0; i < nb; ++i) std::copy(blocks.begin() + 8 * order[i], blocks.begin() + 8 * order[i] + 8,
                                           sorted.begin() + 8 * (size_t)i);
    SFB_TRY(p->d_regz_blocks.alloc(sorted.size()));
    SFB_CUDA_OK(cudaMemcpyAsync(p->d_regz_blocks.p, sorted.data(), sorted.size() * sizeof(int), cudaMemcpyHostToDevice,
                                stream));
    RegzArgs args;
    args.G = p->d_G.p;
    args.What = d_What;
    args.row_out = p->d_row_out.p;
    args.row_n = p->d_row_n.p;
    args.row_n2 = p->d_row_n2.p;
    args.pairidx = p->d_pairidx.p;
    args.M = d_M;
    args.ldM = ldM;
    args.colbase = d_colbase;
    args.nmax = p->nmax;
    args.nrp = p->nrp;
    args.col_lo = (int)col_lo;
    args.col_hi = (int)col_hi;
    args.div2Lp1 = div2Lp1;
    args.interchange = interchange;
    args.dbg = getenv("SFB_CMIX_DBG") ? atoi(getenv("SFB_CMIX_DBG")) : 0;
    args.force_rowtab = getenv("SFB_REGZ_FULLDIAG") ? 1 : 0;
    // persistent launches pull their blocks from a queue head each (zeroed here, in stream order)
    const bool want_persist = getenv("SFB_REGZ_ONEBLOCK") == nullptr;
    static int num_sms = 0;
    if (!num_sms) {
        int dev = 0;
        SFB_CUDA_OK(cudaGetDevice(&dev));
        SFB_CUDA_OK(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
    }
    if (want_persist) {
        SFB_TRY(p->d_regz_queue.alloc(8));
        SFB_CUDA_OK(cudaMemsetAsync(p->d_regz_queue.p, 0, 8 * sizeof(int), stream));
    }
    int nlaunch = 0;
    int i0 = 0;
    while (i0 < nb) {
        const int AT = (sorted[8 * (size_t)i0 + 2] + 7) / 8;
        int i1 = i0, max_rows = 0;
        while (i1 < nb && (sorted[8 * (size_t)i1 + 2] + 7) / 8 == AT) {
            max_rows = std::max(max_rows, sorted[8 * (size_t)i1 + 5]);
            const double b = sorted[8 * (size_t)i1 + 3];
            // executed DMMA flops: Z = AT*NT*2NT DMMAs per N, T = 2NT DMMAs per sub-tile: AT*AT per (N < N') tile,
            // AT(AT+1)/2 per N' = N tile
            const double diag_sub = (interchange || args.force_rowtab) ? AT * AT : AT * (AT + 1) / 2;
            *flops += 512.0 * (AT * NT * 2.0 * NT * b + 2.0 * NT * (AT * AT * b * (b - 1) / 2 + diag_sub * b));
            ++i1;
        }
        args.blocks = p->d_regz_blocks.p + 8 * (size_t)i0;
        RegzLaunchCtx ctx;
        ctx.G = p->d_G.p;
        ctx.g_rows = (long long)p->nblk * p->nmax;
        ctx.What = d_What;
        ctx.w_rows = (long long)(p->d_What.n / (size_t)p->nrp);
        ctx.want_tma = getenv("SFB_REGZ_CPASYNC") == nullptr;
        ctx.want_persist = want_persist && nlaunch < 8;
        ctx.queue = want_persist ? p->d_regz_queue.p + nlaunch : nullptr;
        ctx.num_sms = num_sms;
        ++nlaunch;
        if (NT == 4)
            SFB_TRY(launch_regz_at<4>(AT, args, ctx, i1 - i0, p->nmax, max_rows, stream));
        else
            SFB_TRY(launch_regz_at<8>(AT, args, ctx, i1 - i0, p->nmax, max_rows, stream));
        ++*launches;
        i0 = i1;
    }
    return 0;
}

}  // namespace sfb
