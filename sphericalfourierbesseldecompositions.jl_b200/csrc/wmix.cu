// calc_wmix on sm_100a: the full window mixing matrix W_{nlm}^{n'l'm'}.
//
// Replaces (reference hsgg/SphericalFourierBesselDecompositions.jl):
//   calc_wmix        src/windows.jl:299-364   loop over (nl, n'l') blocks, (m, m') inside
//   calc_wmix_ii     src/windows.jl:273-294   w = (-1)^m [M<0: (-1)^M conj] Σ_L gaunt_L · Σ_r G_nl G_n'l' W_{L|M|}(r), M = m - m'
//   calc_gaunts_L    src/windows.jl:449-464   gaunt_L = (l l' L; -m m' M)(l l' L; 0 0 0) sqrt((2L+1)(2l+1)(2l'+1)/4π)
//   calc_w3j_f       src/windows.jl:434-446   WignerFamilies.wigner3j_f!: all L of a family by the Schulten-Gordon recursion
//
// One CTA = (l, l', tile of 64 (n, n') pairs).  For every |M| in turn:
//   1. overlap integrals  I[(n,n')][(L, re/im)] = Σ_r (G_nl ⊙ G_n'l')[r] W_{L|M|}(r)  for the L of the (l, l') triangle with
//      l + l' + L even — a real GEMM [64 pairs] x [2 nL] x [nr] on the FP64 tensor cores (DMMA), operands staged in
//      shared memory in k-chunks of 32 shells, result kept in shared memory;
//   2. the 3j families of all (m, m') with |m - m'| = |M| (or m + m' = |M| for neg_m), one family per warp pass:
//      the recursion coefficients A(j), B(j) are evaluated by all lanes in parallel, lanes 0 and 1 run the forward and
//      the backward three-term chain, the match index (largest product of the normalised runs) and the normalisation
//      Σ (2j+1) f² = 1 are warp-shuffle reductions; then the lanes take the (n, n') pairs and sum gaunt_L · I over L.
#include "wmix.cuh"

#include <algorithm>
#include <cmath>
#include <vector>

namespace sfb {

constexpr int kWmT = 256;
constexpr int kWmLdA = 36, kWmLdB = 68;

struct WmixArgs {
    const double* G;        // [l][nmax][nrp]
    const double* alm;      // planar [lm (m-major, LMAX = 2 lmax)][re,im][nrp]
    const int* nmax_l;      // [lmax+1]
    const long long* nbase; // [nmax]: 0-based index of (n, l=0, m=0)
    double* out;            // nlmsize x nlmsize complex, column-major
    long long nlmsize;
    int lmax, nmax, nr, nrp, neg_m;
    int ldi;                // leading dimension of the I tile in shared memory (>= 2 (lmax+1), ≡ 4 mod 16)
    int jcap;               // capacity of the family arrays (2 lmax + 2)
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, off));
    return v;
}

// (j j2 j3; m1 m2 m3), j = jmin..jmax, m1 = -m2-m3, into f[0..n) (warp-private shared memory; Aj, Bj, fw, bw scratch of
// the same capacity).  The WignerFamilies algorithm (Schulten-Gordon / Luscombe-Luban two-sided recursion): same steps as the CPU checker.
__device__ void w3j_family_warp(int j2, int j3, int m2, int m3, double* Aj, double* Bj, double* fw, double* bw, double* f,
                                int* jmin_out, int* n_out) {
    const int lane = threadIdx.x & 31;
    const int m1 = -m2 - m3;
    const int jmin = max(abs(j2 - j3), abs(m1)), jmax = j2 + j3;
    const int n = jmax - jmin + 1;
    *jmin_out = jmin;
    *n_out = n;
    const double sign = ((j2 - j3 - m1) & 1) ? -1.0 : 1.0;
    if (n == 1) {
        if (lane == 0) f[0] = sign / sqrt(2.0 * jmin + 1.0);      // positive value, then the sign rule
        __syncwarp();
        return;
    }
    // coefficients, all lanes: A(j) for j = jmin..jmax+1, B(j) for j = jmin..jmax
    const double d23 = (double)(j2 - j3) * (j2 - j3), s23 = (double)(j2 + j3 + 1) * (j2 + j3 + 1), mm = (double)m1 * m1;
    const double c2 = (double)j2 * (j2 + 1) * m1 - (double)j3 * (j3 + 1) * m1, dm = (double)(m3 - m2);
    for (int k = lane; k <= n; k += 32) {
        const double j = jmin + k, jj = j * j;
        Aj[k] = sqrt(fmax(0.0, (jj - d23) * (s23 - jj) * (jj - mm)));
        if (k < n) Bj[k] = -(2.0 * j + 1.0) * (c2 - j * (j + 1.0) * dm);
    }
    __syncwarp();
    // two three-term chains with one instruction stream: lane 0 forward from jmin, lane 1 backward from jmax
    if (lane < 2) {
        const int dir = lane;
        double* x = dir ? bw : fw;
        double prev = 0.0, cur = 1.0;
        x[dir ? n - 1 : 0] = 1.0;
        for (int k = 0; k < n - 1; ++k) {
            const int kk = dir ? n - 1 - k : k;            // index of `cur`
            const double j = jmin + kk;
            const double c1 = dir ? j * Aj[kk + 1] : (j + 1.0) * Aj[kk];
            const double d = dir ? (j + 1.0) * Aj[kk] : j * Aj[kk + 1];
            const double nxt = -(Bj[kk] * cur + c1 * prev) / d;
            x[dir ? kk - 1 : kk + 1] = nxt;
            prev = cur;
            cur = nxt;
        }
    }
    __syncwarp();
    const bool fw_ok = jmin > 0;   // for jmin = 0 the first forward step is 0/0: the backward run alone is used
    int km = 0;
    if (fw_ok) {
        double mf = 0.0, mb = 0.0;
        for (int k = lane; k < n; k += 32) {
            mf = fmax(mf, fabs(fw[k]));
            mb = fmax(mb, fabs(bw[k]));
        }
        mf = warp_max(mf);
        mb = warp_max(mb);
        double best = -1.0;
        int bi = 1 << 30;
        for (int k = lane; k < n; k += 32) {
            const double sc = fabs(fw[k]) / mf * fabs(bw[k]) / mb;
            if (sc > best) {   // first maximum within the lane's ascending k
                best = sc;
                bi = k;
            }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const double ob = __shfl_xor_sync(0xffffffffu, best, off);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
            if (ob > best || (ob == best && oi < bi)) {
                best = ob;
                bi = oi;
            }
        }
        km = bi;
    }
    const double scale = fw_ok ? bw[km] / fw[km] : 1.0;
    double ss = 0.0;
    for (int k = lane; k < n; k += 32) {
        const double v = (fw_ok && k < km) ? fw[k] * scale : bw[k];
        f[k] = v;
        ss += (2.0 * (jmin + k) + 1.0) * v * v;
    }
    ss = warp_sum(ss);
    __syncwarp();
    double nrm = 1.0 / sqrt(ss);
    if (f[n - 1] * sign < 0.0) nrm = -nrm;
    __syncwarp();
    for (int k = lane; k < n; k += 32) f[k] *= nrm;
    __syncwarp();
}

// (l l' L; 0 0 0) for L = |l-l'| + 2k, k = 0..min(l,l'), signed, by the ratio recursion of w3j000sq_table_kernel
// (cmix.cu) + the sum rule; one warp, result in w0[k].
__device__ void w3j000_signed_warp(int l, int lp, double* w0) {
    const int lane = threadIdx.x & 31;
    const int lo = min(l, lp), hi = max(l, lp), nk = lo + 1;
    double carry = 1.0, sum = 0.0;
    for (int base = 0; base < nk; base += 32) {
        const int k = base + lane;
        double r = 1.0;
        if (k >= 1 && k < nk) {
            const int L1 = hi - lo + 2 * (k - 1);
            const double g = 0.5 * (l + lp + L1), a = g - l, b = g - lp, c = g - L1;
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
            w0[k] = w;
            sum += (2.0 * (hi - lo + 2 * k) + 1.0) * w;
        }
    }
    sum = warp_sum(sum);
    __syncwarp();
    for (int k = lane; k < nk; k += 32) {
        const int g = (l + lp + (hi - lo + 2 * k)) / 2;
        const double v = sqrt(w0[k] / sum);
        w0[k] = (g & 1) ? -v : v;
    }
    __syncwarp();
}

__global__ void __launch_bounds__(kWmT) wmix_kernel(WmixArgs a) {
    extern __shared__ double sm[];
    const int l = blockIdx.x / (a.lmax + 1), lp = blockIdx.x % (a.lmax + 1);
    const int al = a.nmax_l[l], alp = a.nmax_l[lp];
    const int npairs = al * alp, p0 = blockIdx.y * 64;
    if (p0 >= npairs) return;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int wm = warp >> 1, wn = warp & 1;
    const int LMAX = 2 * a.lmax, nrp = a.nrp, ldi = a.ldi, jcap = a.jcap;
    // shared memory carve-up
    double* As = sm;                         // [64][kWmLdA]
    double* Bs = As + 64 * kWmLdA;           // [32][kWmLdB]
    double* Is = Bs + 32 * kWmLdB;           // [64][ldi]
    double* w0 = Is + 64 * ldi;              // [lmax+1] signed (l l' L;000)
    double* fam = w0 + (a.lmax + 1);         // per warp: Aj, Bj, fw, bw, f  (5 x jcap)
    double* myfam = fam + (size_t)warp * 5 * jcap;

    if (warp == 0) w3j000_signed_warp(l, lp, w0);
    const int dll = abs(l - lp), par = (l + lp) & 1;
    const int Mmax = a.neg_m ? l + lp : max(l, lp);
    const double pref = sqrt((2.0 * l + 1.0) * (2.0 * lp + 1.0) / (4.0 * 3.14159265358979323846));
    const double* Gl = a.G + (size_t)l * a.nmax * nrp;
    const double* Glp = a.G + (size_t)lp * a.nmax * nrp;
    __syncthreads();

    for (int M = 0; M <= Mmax; ++M) {
        // L of the triangle with the parity of l + l' and L >= M
        int Lmin = max(dll, M);
        if ((Lmin & 1) != par) ++Lmin;
        const int nL = (Lmin <= l + lp) ? (l + lp - Lmin) / 2 + 1 : 0;
        if (nL > 0) {
            const int ncol = 2 * nL;
            // ---- 1. I = (G_l ⊙ G_l') · W_{L M}: 64 pairs x ncol, k over shells ----
            for (int cb = 0; cb < ncol; cb += 64) {
                double acc[2][4][2];
#pragma unroll
                for (int i = 0; i < 2; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
                for (int r0 = 0; r0 < nrp; r0 += 32) {
                    {
                        const int kk = tid & 31;
#pragma unroll
                        for (int q = 0; q < 8; ++q) {
                            const int row = (tid >> 5) + 8 * q, pr = p0 + row;
                            double v = 0.0;
                            if (pr < npairs && r0 + kk < nrp) {
                                const int n = pr / alp, np_ = pr % alp;
                                v = Gl[(size_t)n * nrp + r0 + kk] * Glp[(size_t)np_ * nrp + r0 + kk];
                            }
                            As[row * kWmLdA + kk] = v;
                        }
                        // B[k][col]: col = (L index, comp); rows of the planar alm are contiguous in r
#pragma unroll
                        for (int q = 0; q < 8; ++q) {
                            const int col = (tid >> 5) + 8 * q + 0;      // 8 columns per sweep, 64 in 8 sweeps
                            const int c = cb + col;
                            double v = 0.0;
                            if (c < ncol && r0 + kk < nrp) {
                                const int L = Lmin + 2 * (c >> 1), comp = c & 1;
                                const size_t lm = (size_t)L + ((size_t)M * (2 * LMAX + 1 - M)) / 2;
                                v = a.alm[(lm * 2 + comp) * nrp + r0 + kk];
                            }
                            Bs[kk * kWmLdB + col] = v;
                        }
                    }
                    __syncthreads();
                    warp_gemm_ss<2, 4>(acc, As + wm * 16 * kWmLdA, kWmLdA, Bs + wn * 32, kWmLdB, 32);
                    __syncthreads();
                }
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const int row = wm * 16 + i * 8 + g;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int col = cb + wn * 32 + j * 8 + 2 * t;
                        if (col < ncol) {
                            Is[row * ldi + col] = acc[i][j][0];
                            Is[row * ldi + col + 1] = acc[i][j][1];
                        }
                    }
                }
            }
            __syncthreads();
            // ---- 2. families with |Mtrue| = M, dealt round-robin to the warps ----
            // neg_m = 0: (m, m' = m - M) and, for M > 0, (m, m' = m + M);   neg_m = 1: m + m' = M
            const int nfamA = a.neg_m ? max(0, min(l, M) - max(0, M - lp) + 1) : max(0, min(l, lp + M) - M + 1);
            const int nfamB = (a.neg_m || M == 0) ? 0 : max(0, min(l, lp - M) + 1);
            for (int fi = warp; fi < nfamA + nfamB; fi += kWmT / 32) {
                int m, mp;
                if (a.neg_m) {
                    m = max(0, M - lp) + fi;
                    mp = M - m;
                } else if (fi < nfamA) {
                    m = M + fi;
                    mp = m - M;
                } else {
                    m = fi - nfamA;
                    mp = m + M;
                }
                // family (L l l'; Mtrue, m2, m3) with (m2, m3) = (-m, m') or, for neg_m, (m, m')
                const int m2 = a.neg_m ? m : -m, m3 = mp;
                const int Mtrue = -m2 - m3;                       // = m - m'  (neg_m: -m - m')
                int jmin, n;
                double* f = myfam + 4 * jcap;
                w3j_family_warp(l, lp, m2, m3, myfam, myfam + jcap, myfam + 2 * jcap, myfam + 3 * jcap, f, &jmin, &n);
                // gaunt_L for the L of the right parity (in place)
                for (int k = lane; k < n; k += 32) {
                    const int L = jmin + k;
                    double gv = 0.0;
                    if (((L & 1) == par) && L >= dll) gv = f[k] * w0[(L - dll) >> 1] * pref * sqrt(2.0 * L + 1.0);
                    f[k] = gv;
                }
                __syncwarp();
                const double sgn_m = (m & 1) ? -1.0 : 1.0;
                const double sgn_M = (Mtrue & 1) ? -1.0 : 1.0;
                for (int row = lane; row < 64; row += 32) {
                    const int pr = p0 + row;
                    if (pr >= npairs) continue;
                    double re = 0.0, im = 0.0;
                    for (int L = Lmin; L <= l + lp; L += 2) {
                        const double gv = f[L - jmin];
                        const int c = (L - Lmin);                 // 2 * column pair index
                        re = fma(gv, Is[row * ldi + c], re);
                        im = fma(gv, Is[row * ldi + c + 1], im);
                    }
                    if (Mtrue < 0) {                              // (-1)^M conj
                        re *= sgn_M;
                        im *= -sgn_M;
                    }
                    re *= sgn_m;
                    im *= sgn_m;
                    const int nn = pr / alp, np_ = pr % alp;      // 0-based n, n'
                    const long long i = a.nbase[nn] + (long long)l * (l + 1) / 2 + m;
                    const long long ip = a.nbase[np_] + (long long)lp * (lp + 1) / 2 + mp;
                    double* dst = a.out + 2 * (i + a.nlmsize * ip);
                    dst[0] = re;
                    dst[1] = im;
                }
                __syncwarp();
            }
        }
        __syncthreads();   // Is is rewritten by the next M
    }
}

int wmix_run(const double* d_alm, int nrp_alm, const double* G, int64_t nr, int64_t nmax, int64_t lmax,
             const int64_t* nmax_l, const int64_t* lmax_n, int neg_m, double* d_out, int64_t* nlmsize_out,
             cudaStream_t stream) {
    SFB_REQUIRE(d_alm && G && nmax_l && lmax_n && d_out, "calc_wmix: null pointer");
    SFB_REQUIRE(nr >= 1 && nmax >= 1 && lmax >= 0, "calc_wmix: bad sizes");
    const int nrp = (int)round_up(nr, 8);
    SFB_REQUIRE(nrp == nrp_alm, "calc_wmix: alm padding mismatch");
    // index tables  (src/modes.jl:178-232): idx(n,l,m) = 1 + Σ_{k<n} lmsize(lmax_n[k]) + l(l+1)/2 + m
    std::vector<long long> nbase(nmax, 0);
    long long acc = 0;
    for (int64_t n = 0; n < nmax; ++n) {
        nbase[n] = acc;
        SFB_REQUIRE(lmax_n[n] >= 0 && lmax_n[n] <= lmax, "calc_wmix: lmax_n out of range");
        acc += (lmax_n[n] + 1) * (lmax_n[n] + 2) / 2;
    }
    const long long nlmsize = acc;
    if (nlmsize_out) *nlmsize_out = nlmsize;
    std::vector<int> nl(lmax + 1);
    int amax = 0;
    for (int64_t l = 0; l <= lmax; ++l) {
        SFB_REQUIRE(nmax_l[l] >= 1 && nmax_l[l] <= nmax, "calc_wmix: nmax_l out of range");
        nl[l] = (int)nmax_l[l];
        amax = std::max(amax, nl[l]);
        // the (n, l) index set must be the staircase the index formula assumes
        for (int64_t n = 0; n < nmax; ++n)
            SFB_REQUIRE((n < nmax_l[l]) == (l <= lmax_n[n]), "calc_wmix: nmax_l and lmax_n are inconsistent");
    }
    std::vector<double> Gh((size_t)(lmax + 1) * nmax * nrp, 0.0);
    for (int64_t l = 0; l <= lmax; ++l)
        for (int n = 0; n < nl[l]; ++n)
            for (int64_t r = 0; r < nr; ++r) {
                const double v = G[(size_t)r + (size_t)nr * (n + (size_t)nmax * l)];
                SFB_REQUIRE(std::isfinite(v), "calc_wmix: non-finite radial basis value");
                Gh[((size_t)l * nmax + n) * nrp + r] = v;
            }
    DevBuf<double> dG;
    DevBuf<int> dnl;
    DevBuf<long long> dnb;
    SFB_TRY(dG.alloc(Gh.size()));
    SFB_TRY(dnl.alloc(nl.size()));
    SFB_TRY(dnb.alloc(nbase.size()));
    SFB_CUDA_OK(cudaMemcpyAsync(dG.p, Gh.data(), Gh.size() * sizeof(double), cudaMemcpyHostToDevice, stream));
    SFB_CUDA_OK(cudaMemcpyAsync(dnl.p, nl.data(), nl.size() * sizeof(int), cudaMemcpyHostToDevice, stream));
    SFB_CUDA_OK(cudaMemcpyAsync(dnb.p, nbase.data(), nbase.size() * sizeof(long long), cudaMemcpyHostToDevice, stream));
    WmixArgs a;
    a.G = dG.p;
    a.alm = d_alm;
    a.nmax_l = dnl.p;
    a.nbase = dnb.p;
    a.out = d_out;
    a.nlmsize = nlmsize;
    a.lmax = (int)lmax;
    a.nmax = (int)nmax;
    a.nr = (int)nr;
    a.nrp = nrp;
    a.neg_m = neg_m ? 1 : 0;
    a.ldi = (int)round_up(2 * (lmax + 1), 16) + 4;
    a.jcap = 2 * (int)lmax + 2;
    const size_t smem = sizeof(double) * ((size_t)64 * kWmLdA + 32 * kWmLdB + (size_t)64 * a.ldi + (lmax + 1) +
                                          (size_t)(kWmT / 32) * 5 * a.jcap);
    SFB_REQUIRE(smem <= 227 * 1024, "calc_wmix: lmax too large for the shared-memory tiling of this build");
    SFB_CUDA_OK(cudaFuncSetAttribute(wmix_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)((lmax + 1) * (lmax + 1)), (unsigned)ceil_div((int64_t)amax * amax, 64));
    wmix_kernel<<<grid, kWmT, smem, stream>>>(a);
    SFB_CUDA_OK(cudaGetLastError());
    SFB_CUDA_OK(cudaStreamSynchronize(stream));   // the staging buffers above are released on return
    return 0;
}

}  // namespace sfb
