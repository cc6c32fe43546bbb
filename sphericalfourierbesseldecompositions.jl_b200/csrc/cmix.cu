// Stage 2+3 kernels: see cmix.cuh for the reference functions each one replaces.
#include "cmix.cuh"

#include <algorithm>
#include <cmath>
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
__global__ void wl_build_kernel(const double* __restrict__ alm1, const double* __restrict__ alm2,
                                double* __restrict__ W, int LMAX, int nrp) {
    const int L1 = blockIdx.x;
    const int n2 = nrp * nrp;
    for (int e = threadIdx.x; e < n2; e += blockDim.x) {
        const int i = e / nrp, j = e - i * nrp;
        double s = 0.0;
        for (int M = 0; M <= L1; ++M) {
            const size_t lm = (size_t)L1 + ((size_t)M * (2 * LMAX + 1 - M)) / 2;
            const double* a1 = alm1 + lm * 2 * nrp;
            const double* a2 = alm2 + lm * 2 * nrp;
            const double v = a1[i] * a2[j] + a1[nrp + i] * a2[nrp + j];
            s += (M == 0) ? v : 2.0 * v;
        }
        W[(size_t)L1 * n2 + e] = s;
    }
}

// =============================================================================================
// Ŵ_{ℓL}[r][r'] = Σ_{L1} (ℓ L L1;000)² W_{L1}[r][r']   — the L1 loop of src/windows.jl:619-622 hoisted out of
// the per-element work (the quadratic form is linear in W).  One CTA = one ℓ and four L of equal parity, so
// every W_{L1} element loaded from L2 feeds four accumulators.
__global__ void __launch_bounds__(256) what_build_kernel(const double* __restrict__ W, const double* __restrict__ w2,
                                                         double* __restrict__ What, int ell0, int lmax, int nrp) {
    extern __shared__ double wsm[];  // [4][lmax+1]
    const int ell = ell0 + blockIdx.y;
    const int gi = blockIdx.x, p = gi & 1, base = (gi >> 1) * 8 + p;
    const int par = (ell + p) & 1;
    const int KW = lmax + 1;
    for (int x = threadIdx.x; x < 4 * KW; x += blockDim.x) wsm[x] = 0.0;
    __syncthreads();
    int L1lo = 1 << 30, L1hi = -1;
    for (int q = 0; q < 4; ++q) {
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
    for (int e = threadIdx.x; e < n2; e += blockDim.x) {
        double acc[4] = {0.0, 0.0, 0.0, 0.0};
        for (int L1 = L1lo; L1 <= L1hi; L1 += 2) {
            const double x = __ldg(W + (size_t)L1 * n2 + e);
            const int h = (L1 - par) >> 1;
#pragma unroll
            for (int q = 0; q < 4; ++q) acc[q] = fma(wsm[q * KW + h], x, acc[q]);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int L = base + 2 * q;
            if (L <= lmax) What[((size_t)blockIdx.y * (lmax + 1) + L) * n2 + e] = acc[q];
        }
    }
}

// =============================================================================================
// Coupling-matrix block kernel.  One CTA = (ℓ, L, N):
//   Z_N[n][r']   = Σ_r  G_ℓn[r] G_LN[r] Ŵ_ℓL[r][r']                       (DMMA,  a × nr × nr)
//   T_N'[n][n']  = Σ_r' Z_N[n][r'] G_LN'[r'] G_ℓn'[r']   for N' >= N       (DMMA,  a × a × nr, one warp per N')
//   M[(ℓ,n,n'),(L,N,N')] = c_L · ( T_N'[n][n'] + [N≠N'] T_N'[n'][n] )      (src/windows.jl:727-736)
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
    const int* pairidx;     // [L][nmax][nmax] -> output column or -1
    const int* nl_L;
    const int* nl_N;
    double* M;
    long long ldM;
    int ell0, lmax, nmax, nrp, S;
    int div2Lp1, interchange;
};

constexpr int kCmixThreads = 128;
constexpr int kCmixWarps = kCmixThreads / 32;

template <int AT, bool SYM>
__global__ void __launch_bounds__(kCmixThreads) cmix_block_kernel(CmixArgs p) {
    extern __shared__ double sm[];
    const int ell = p.ell_list[blockIdx.y];
    const int L = p.nl_L[blockIdx.x], N = p.nl_N[blockIdx.x];
    const int a = p.a_of_ell[ell], b = p.a_of_ell[L];
    if (a == 0 || N >= b) return;
    constexpr int AP = AT * 8;
    const int S = p.S, nrp = p.nrp;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;

    double* Gl = sm;                   // [AP][S]   G_ℓn[r]
    double* GL = Gl + AP * S;          // [nmax][S] G_LN[r]
    double* Zs = GL + p.nmax * S;      // [AP][S]
    double* Zt = Zs + AP * S;          // [AP][S] (only if !SYM)
    double* Ts = Zt + (SYM ? 0 : AP * S);  // [warps][(SYM?1:2)][AP][AP+1]

    // ---- stage G_ℓ and G_L rows (zero-padded) ------------------------------------------------
    for (int x = tid; x < AP * S; x += kCmixThreads) {
        const int n = x / S, r = x - n * S;
        Gl[x] = (n < a && r < nrp) ? p.G[((size_t)ell * p.nmax + n) * nrp + r] : 0.0;
    }
    for (int x = tid; x < b * S; x += kCmixThreads) {
        const int n = x / S, r = x - n * S;
        GL[x] = (r < nrp) ? p.G[((size_t)L * p.nmax + n) * nrp + r] : 0.0;
    }
    __syncthreads();

    // ---- Z phase -------------------------------------------------------------------------------
    const double* Wh = p.What + ((size_t)(ell - p.ell0) * (p.lmax + 1) + L) * nrp * nrp;
    const double* glN = GL + N * S;
    for (int jt = warp; jt < nrp / 8; jt += kCmixWarps) {
        double acc[AT][2], acct[AT][2];
#pragma unroll
        for (int i = 0; i < AT; ++i) acc[i][0] = acc[i][1] = acct[i][0] = acct[i][1] = 0.0;
        for (int k0 = 0; k0 < nrp; k0 += 4) {
            const double s = glN[k0 + t];
            const double bw = __ldg(Wh + (size_t)(k0 + t) * nrp + jt * 8 + g);
            double bt = 0.0;
            if (!SYM) bt = __ldg(Wh + (size_t)(jt * 8 + g) * nrp + k0 + t);
#pragma unroll
            for (int i = 0; i < AT; ++i) {
                const double av = Gl[(i * 8 + g) * S + k0 + t] * s;
                dmma884(acc[i], av, bw);
                if (!SYM) dmma884(acct[i], av, bt);
            }
        }
#pragma unroll
        for (int i = 0; i < AT; ++i) {
            double* z = Zs + (i * 8 + g) * S + jt * 8 + 2 * t;
            z[0] = acc[i][0];
            z[1] = acc[i][1];
            if (!SYM) {
                double* zt = Zt + (i * 8 + g) * S + jt * 8 + 2 * t;
                zt[0] = acct[i][0];
                zt[1] = acct[i][1];
            }
        }
    }
    __syncthreads();

    // ---- T phase: one warp per N' ----------------------------------------------------------------
    const int r0 = p.ell_ptr[ell], nrows = p.ell_ptr[ell + 1] - r0;
    const double scale = (p.div2Lp1 ? 1.0 : (2.0 * L + 1.0)) * 0.07957747154594767;  // 1/(4π)
    double* Tw = Ts + warp * (SYM ? 1 : 2) * AP * (AP + 1);
    const int* prow = p.pairidx + ((size_t)L * p.nmax + N) * p.nmax;
    for (int N2 = N + warp; N2 < b; N2 += kCmixWarps) {
        const int col = prow[N2];
        if (col < 0) continue;  // warp-uniform
        const double* glN2 = GL + N2 * S;
        double acc[AT][AT][2];
#pragma unroll
        for (int i = 0; i < AT; ++i)
#pragma unroll
            for (int j = 0; j < AT; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
#pragma unroll 1
        for (int pass = 0; pass < (SYM ? 1 : 2); ++pass) {
            const double* Zsrc = pass ? Zt : Zs;
            for (int k0 = 0; k0 < nrp; k0 += 4) {
                const double s = glN2[k0 + t];
                double av[AT], bv[AT];
#pragma unroll
                for (int i = 0; i < AT; ++i) av[i] = Zsrc[(i * 8 + g) * S + k0 + t] * s;
#pragma unroll
                for (int j = 0; j < AT; ++j) bv[j] = Gl[(j * 8 + g) * S + k0 + t];
#pragma unroll
                for (int i = 0; i < AT; ++i)
#pragma unroll
                    for (int j = 0; j < AT; ++j) dmma884(acc[i][j], av[i], bv[j]);
            }
            double* Tp = Tw + pass * AP * (AP + 1);
#pragma unroll
            for (int i = 0; i < AT; ++i)
#pragma unroll
                for (int j = 0; j < AT; ++j) {
                    double* q = Tp + (i * 8 + g) * (AP + 1) + j * 8 + 2 * t;
                    q[0] = acc[i][j][0];
                    q[1] = acc[i][j][1];
                    acc[i][j][0] = acc[i][j][1] = 0.0;
                }
        }
        __syncwarp();
        const double* T1 = Tw;
        const double* T2 = SYM ? Tw : Tw + AP * (AP + 1);
        double* Mcol = p.M + (size_t)col * p.ldM;
        const bool offdiag = (N2 != N);
        for (int idx = lane; idx < nrows; idx += 32) {
            const int orow = p.row_out[r0 + idx];
            if (orow < 0) continue;
            const int n = p.row_n[r0 + idx], n2 = p.row_n2[r0 + idx];
            double v;
            if (p.interchange) {
                v = T2[n2 * (AP + 1) + n];
            } else {
                v = T1[n * (AP + 1) + n2];
                if (offdiag) v += T2[n2 * (AP + 1) + n];
            }
            Mcol[orow] = scale * v;
        }
        __syncwarp();
    }
}

template <int AT, bool SYM>
static int launch_cmix(const CmixArgs& args, int nl, int nells, int nmax, cudaStream_t stream) {
    const int AP = AT * 8;
    const size_t smem = sizeof(double) * ((size_t)AP * args.S * (SYM ? 2 : 3) + (size_t)nmax * args.S +
                                          (size_t)kCmixWarps * (SYM ? 1 : 2) * AP * (AP + 1));
    SFB_REQUIRE(smem <= 227 * 1024, "cmix_block_kernel: shared memory footprint exceeds 227 KB (nr * nmax too large)");
    SFB_CUDA_OK(cudaFuncSetAttribute(cmix_block_kernel<AT, SYM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(nl, nells);
    cmix_block_kernel<AT, SYM><<<grid, kCmixThreads, smem, stream>>>(args);
    SFB_CUDA_OK(cudaGetLastError());
    return 0;
}

template <bool SYM>
static int launch_cmix_at(int AT, const CmixArgs& args, int nl, int nells, int nmax, cudaStream_t stream) {
    switch (AT) {
        case 1: return launch_cmix<1, SYM>(args, nl, nells, nmax, stream);
        case 2: return launch_cmix<2, SYM>(args, nl, nells, nmax, stream);
        case 3: return launch_cmix<3, SYM>(args, nl, nells, nmax, stream);
        case 4: return launch_cmix<4, SYM>(args, nl, nells, nmax, stream);
        default: break;
    }
    set_error("cmix: nmax_l > 32 is not supported by this build");
    return 2;
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
    p->amax_tiles = (amax + 7) / 8;
    if (p->amax_tiles > 4) {
        delete p;
        set_error("cmix_plan_create: nmax_l > 32 is not supported by this build");
        return 2;
    }

    // --- G: [nr][nmax][lmax+1] column-major -> [ell][n][nrp], NaN / padding -> 0 where unused ---
    std::vector<double> Gh((size_t)(lmax + 1) * nmax * p->nrp, 0.0);
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
    rc = rc ? rc : up(p->d_pairidx, pairidx);
    rc = rc ? rc : up(p->d_nl_L, nl_L);
    rc = rc ? rc : up(p->d_nl_N, nl_N);
    rc = rc ? rc : p->d_row_out.alloc(nrows);
    rc = rc ? rc : p->d_ell_list.alloc(lmax + 1);
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

void cmix_plan_destroy(CmixPlan* p) { delete p; }

int cmix_run(CmixPlan* p, const double* d_alm1, const double* d_alm2, int div2Lp1, int interchange, int64_t row_lo,
             int64_t row_hi, double* d_M, int64_t ldM, cudaStream_t stream) {
    SFB_REQUIRE(p && d_alm1 && d_alm2 && d_M, "cmix_run: null pointer");
    SFB_REQUIRE(0 <= row_lo && row_lo <= row_hi && row_hi <= p->nout, "cmix_run: bad row range");
    SFB_REQUIRE(ldM >= row_hi - row_lo, "cmix_run: ldM smaller than the row shard");
    if (row_hi == row_lo) return 0;
    const bool sym = (d_alm1 == d_alm2);
    const int lmax = p->lmax, nrp = p->nrp;
    cudaEvent_t ev[4];
    for (auto& e : ev) SFB_CUDA_OK(cudaEventCreate(&e));

    // rows of the shard -> output row index relative to row_lo, -1 elsewhere; ells touched by the shard
    std::vector<int> row_out(p->h_row_out.size());
    std::vector<char> ell_used(lmax + 1, 0);
    for (int l = 0; l <= lmax; ++l)
        for (int s = p->ell_ptr[l]; s < p->ell_ptr[l + 1]; ++s) {
            const int o = p->h_row_out[s];
            const bool in = (o >= row_lo && o < row_hi);
            row_out[s] = in ? (int)(o - row_lo) : -1;
            if (in) ell_used[l] = 1;
        }
    SFB_CUDA_OK(cudaMemcpyAsync(p->d_row_out.p, row_out.data(), row_out.size() * sizeof(int), cudaMemcpyHostToDevice,
                                stream));

    // ---- W_{L1} ----
    SFB_CUDA_OK(cudaEventRecord(ev[0], stream));
    wl_build_kernel<<<p->LMAX + 1, 256, 0, stream>>>(d_alm1, d_alm2, p->d_W.p, p->LMAX, nrp);
    SFB_CUDA_OK(cudaGetLastError());
    SFB_CUDA_OK(cudaEventRecord(ev[1], stream));
    p->launches = 1;

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
    args.pairidx = p->d_pairidx.p;
    args.nl_L = p->d_nl_L.p;
    args.nl_N = p->d_nl_N.p;
    args.M = d_M;
    args.ldM = ldM;
    args.lmax = lmax;
    args.nmax = p->nmax;
    args.nrp = nrp;
    args.S = p->S;
    args.div2Lp1 = div2Lp1;
    args.interchange = interchange;

    float t_what = 0.f, t_block = 0.f;
    double flops = 0.0;
    const int ngroups = 2 * (int)ceil_div(lmax + 1, 8);
    std::vector<int> ell_list_all;  // device ell_list is filled per launch at distinct offsets
    for (int ell0 = 0; ell0 <= lmax; ell0 += chunk) {
        const int ell1 = std::min(lmax + 1, ell0 + chunk);
        bool any = false;
        for (int l = ell0; l < ell1; ++l) any = any || (ell_used[l] && p->a_of_ell[l] > 0);
        if (!any) continue;
        cudaEvent_t e0, e1, e2;
        SFB_CUDA_OK(cudaEventCreate(&e0));
        SFB_CUDA_OK(cudaEventCreate(&e1));
        SFB_CUDA_OK(cudaEventCreate(&e2));
        SFB_CUDA_OK(cudaEventRecord(e0, stream));
        what_build_kernel<<<dim3(ngroups, ell1 - ell0), 256, 4 * (lmax + 1) * sizeof(double), stream>>>(
            p->d_W.p, p->d_w2.p, p->d_What.p, ell0, lmax, nrp);
        SFB_CUDA_OK(cudaGetLastError());
        p->launches++;
        SFB_CUDA_OK(cudaEventRecord(e1, stream));
        args.ell0 = ell0;
        for (int AT = p->amax_tiles; AT >= 1; --AT) {
            std::vector<int> ells;
            for (int l = ell0; l < ell1; ++l)
                if (ell_used[l] && p->a_of_ell[l] > 0 && (p->a_of_ell[l] + 7) / 8 == AT) ells.push_back(l);
            if (ells.empty()) continue;
            const size_t off = ell_list_all.size();
            ell_list_all.insert(ell_list_all.end(), ells.begin(), ells.end());
            SFB_REQUIRE(ell_list_all.size() <= (size_t)lmax + 1, "cmix_run: internal ell list overflow");
            SFB_CUDA_OK(cudaMemcpyAsync(p->d_ell_list.p + off, ells.data(), ells.size() * sizeof(int),
                                        cudaMemcpyHostToDevice, stream));
            args.ell_list = p->d_ell_list.p + off;
            if (sym)
                SFB_TRY(launch_cmix_at<true>(AT, args, p->nl, (int)ells.size(), p->nmax, stream));
            else
                SFB_TRY(launch_cmix_at<false>(AT, args, p->nl, (int)ells.size(), p->nmax, stream));
            p->launches++;
            for (int l : ells)
                for (int L = 0; L <= lmax; ++L) {
                    const double b = p->a_of_ell[L], ap = AT * 8.0;
                    flops += (sym ? 1.0 : 2.0) * (2.0 * ap * nrp * nrp * b + 2.0 * ap * ap * nrp * b * (b + 1) / 2);
                }
        }
        SFB_CUDA_OK(cudaEventRecord(e2, stream));
        SFB_CUDA_OK(cudaEventSynchronize(e2));
        float a_ms = 0, b_ms = 0;
        cudaEventElapsedTime(&a_ms, e0, e1);
        cudaEventElapsedTime(&b_ms, e1, e2);
        t_what += a_ms;
        t_block += b_ms;
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
        cudaEventDestroy(e2);
    }
    SFB_CUDA_OK(cudaStreamSynchronize(stream));
    cudaEventElapsedTime(&p->t_wl, ev[0], ev[1]);
    p->t_what = t_what;
    p->t_block = t_block;
    p->flops_executed = flops;
    for (auto& e : ev) cudaEventDestroy(e);
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
