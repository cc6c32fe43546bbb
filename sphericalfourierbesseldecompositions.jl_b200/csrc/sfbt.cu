// SFB transforms next to the window path on sm_100a: they reuse the batched HEALPix kernels of stage 1.
//
// Replaces (reference hsgg/SphericalFourierBesselDecompositions.jl):
//   field2anlm (= field2anlm_v2)   src/cat2anlm.jl:326-362   per shell map2alm! (niter 3), f_nlm += g_nl(r) r² Δr a_lm(r)
//   anlm2field                     src/cat2anlm.jl:385-422   a_lm(r) = Σ_n g_nl(r) f_nlm, alm2map! per shell
//   win_rhat_ln (dense)            src/windows.jl:244-256    W[p, l, n] = Δr Σ_r win[r,p] r² g_nl(r): a GEMM (DMMA)
//   cat2amln                       src/cat2anlm.jl:257-314   per (n,l): map = Σ_gal w g_nl(r_gal)/(n̄ ΔΩ) − W[:,l,n], map2alm,
//                                                            keep a_{l m}, m = 0..l  (batched over (n,l) as "shells")
#include "sfbt.cuh"

#include <algorithm>
#include <cmath>
#include <vector>

namespace sfb {

// ---- index tables shared by the kernels: idx(n,l,m) = nbase[n] + l(l+1)/2 + m  (src/modes.jl:222-232) ----
struct NlmTabs {
    std::vector<long long> nbase;
    std::vector<int> nl;      // nmax_l
    long long nlmsize = 0;
};

static int make_nlm_tabs(NlmTabs* t, int64_t nmax, int64_t lmax, const int64_t* nmax_l, const int64_t* lmax_n) {
    SFB_REQUIRE(nmax_l && lmax_n && nmax >= 1 && lmax >= 0, "bad mode tables");
    t->nbase.assign(nmax, 0);
    long long acc = 0;
    for (int64_t n = 0; n < nmax; ++n) {
        SFB_REQUIRE(lmax_n[n] >= 0 && lmax_n[n] <= lmax, "lmax_n out of range");
        t->nbase[n] = acc;
        acc += (lmax_n[n] + 1) * (lmax_n[n] + 2) / 2;
    }
    t->nlmsize = acc;
    t->nl.resize(lmax + 1);
    for (int64_t l = 0; l <= lmax; ++l) {
        SFB_REQUIRE(nmax_l[l] >= 1 && nmax_l[l] <= nmax, "nmax_l out of range");
        t->nl[l] = (int)nmax_l[l];
        for (int64_t n = 0; n < nmax; ++n)
            SFB_REQUIRE((n < nmax_l[l]) == (l <= lmax_n[n]), "nmax_l and lmax_n are inconsistent");
    }
    return 0;
}

// host table T[r, n, l] (nr x nmax x (lmax+1), column-major, NaN where n > nmax_l) -> device [l][n][nrp], zero padded
static int upload_radial_table(const double* T, int64_t nr, int64_t nmax, int64_t lmax, const std::vector<int>& nl, int nrp,
                               DevBuf<double>& d, cudaStream_t st) {
    std::vector<double> h((size_t)(lmax + 1) * nmax * nrp, 0.0);
    for (int64_t l = 0; l <= lmax; ++l)
        for (int n = 0; n < nl[l]; ++n)
            for (int64_t r = 0; r < nr; ++r) {
                const double v = T[(size_t)r + (size_t)nr * (n + (size_t)nmax * l)];
                SFB_REQUIRE(std::isfinite(v), "non-finite radial basis value for a mode of the table");
                h[((size_t)l * nmax + n) * nrp + r] = v;
            }
    SFB_TRY(d.alloc(h.size()));
    SFB_CUDA_OK(cudaMemcpyAsync(d.p, h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice, st));
    SFB_CUDA_OK(cudaStreamSynchronize(st));
    return 0;
}

template <typename T>
static int upload_vec(DevBuf<T>& d, const std::vector<T>& h, cudaStream_t st) {
    SFB_TRY(d.alloc(std::max<size_t>(1, h.size())));
    if (!h.empty()) SFB_CUDA_OK(cudaMemcpyAsync(d.p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice, st));
    SFB_CUDA_OK(cudaStreamSynchronize(st));
    return 0;
}

__device__ __forceinline__ size_t lm_mm(int lmax, int l, int m) { return (size_t)l + ((size_t)m * (2 * lmax + 1 - m)) / 2; }

// f_nlm[(n,l,m)] = Σ_r T[l][n][r] a_lm(r)      one warp per (n, l, m)
__global__ void __launch_bounds__(256) nlm_contract_kernel(const double* __restrict__ alm, const double* __restrict__ T,
                                                           const int* __restrict__ nl, const long long* __restrict__ nbase,
                                                           int lmax, int nmax, int nrp, double* __restrict__ out) {
    const int lm = blockIdx.x;                       // m-fast enumeration: lm = l(l+1)/2 + m
    int l = (int)((sqrt(8.0 * lm + 1.0) - 1.0) * 0.5);
    while ((l + 1) * (l + 2) / 2 <= lm) ++l;
    while (l * (l + 1) / 2 > lm) --l;
    const int m = lm - l * (l + 1) / 2;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const double* a = alm + lm_mm(lmax, l, m) * 2 * nrp;
    for (int n = warp; n < nl[l]; n += blockDim.x >> 5) {
        const double* t = T + ((size_t)l * nmax + n) * nrp;
        double re = 0.0, im = 0.0;
        for (int r = lane; r < nrp; r += 32) {
            re = fma(t[r], a[r], re);
            im = fma(t[r], a[nrp + r], im);
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            re += __shfl_xor_sync(0xffffffffu, re, off);
            im += __shfl_xor_sync(0xffffffffu, im, off);
        }
        if (lane == 0) {
            const long long i = nbase[n] + (long long)l * (l + 1) / 2 + m;
            out[2 * i] = re;
            out[2 * i + 1] = im;
        }
    }
}

// a_lm(r) = Σ_n g[l][n][r] f_nlm[(n,l,m)]     CTA per (l, m), threads over r
__global__ void __launch_bounds__(128) nlm_expand_kernel(const double* __restrict__ f, const double* __restrict__ g,
                                                         const int* __restrict__ nl, const long long* __restrict__ nbase,
                                                         int lmax, int nmax, int nrp, double* __restrict__ alm) {
    const int l = blockIdx.x, m = blockIdx.y;
    if (m > l) return;
    double* a = alm + lm_mm(lmax, l, m) * 2 * nrp;
    for (int r = threadIdx.x; r < nrp; r += blockDim.x) {
        double re = 0.0, im = 0.0;
        for (int n = 0; n < nl[l]; ++n) {
            const long long i = nbase[n] + (long long)l * (l + 1) / 2 + m;
            const double gv = g[((size_t)l * nmax + n) * nrp + r];
            re = fma(gv, f[2 * i], re);
            im = fma(gv, f[2 * i + 1], im);
        }
        a[r] = re;
        a[nrp + r] = im;
    }
}

// out[p + npix (l + (lmax+1) n)] = Σ_r win[p][r] T[l][n][r]  (NaN where n >= nmax_l[l])      64 pixels x 64 (l,n) columns per CTA
__global__ void __launch_bounds__(256) win_rhat_ln_kernel(const double* __restrict__ win, long long ldw, long long npix,
                                                          const double* __restrict__ T, const int* __restrict__ nl,
                                                          int lmax, int nmax, int nr, int nrp, double* __restrict__ out) {
    constexpr int LDA = 36, LDB = 68;
    __shared__ double As[64 * LDA];
    __shared__ double Bs[32 * LDB];
    const long long p0 = (long long)blockIdx.x * 64;
    const int c0 = blockIdx.y * 64, ncols = (lmax + 1) * nmax;   // column c = l + (lmax+1) n
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int wm = warp >> 1, wn = warp & 1;
    double acc[2][4][2];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    for (int r0 = 0; r0 < nr; r0 += 32) {
        const int kk = tid & 31;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int row = (tid >> 5) + 8 * q;
            const long long p = p0 + row;
            As[row * LDA + kk] = (p < npix && r0 + kk < nr) ? win[p * ldw + r0 + kk] : 0.0;
            const int c = c0 + row;
            double v = 0.0;
            if (c < ncols && r0 + kk < nr) {
                const int l = c % (lmax + 1), n = c / (lmax + 1);
                if (n < nl[l]) v = T[((size_t)l * nmax + n) * nrp + r0 + kk];
            }
            Bs[kk * LDB + row] = v;
        }
        __syncthreads();
        warp_gemm_ss<2, 4>(acc, As + wm * 16 * LDA, LDA, Bs + wn * 32, LDB, 32);
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const long long p = p0 + wm * 16 + i * 8 + g;
        if (p >= npix) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int c = c0 + wn * 32 + j * 8 + 2 * t + e;
                if (c >= ncols) continue;
                const int l = c % (lmax + 1), n = c / (lmax + 1);
                out[p + npix * (size_t)c] = (n < nl[l]) ? acc[i][j][e] : nan("");
            }
    }
}

// maps[pix][b] = Σ_{gal in pix} gw[gal, b] / (n̄ ΔΩ) − W[pix, l_b, n_b]: thread per (pixel, mode); the galaxies of a pixel are
// summed in catalogue order (deterministic, the order of transform_gnl_spmap!, src/cat2anlm.jl:92-99)
__global__ void cat_maps_kernel(const long long* __restrict__ pixptr, const long long* __restrict__ gidx,
                                const double* __restrict__ gw, long long ngal, const double* __restrict__ wrhatln,
                                const int* __restrict__ mode_l, const int* __restrict__ mode_n, int nb, int nbp,
                                long long npix, int lmax, double inv_nbar_domega, double* __restrict__ maps) {
    const long long total = npix * nbp;
    for (long long x = blockIdx.x * (long long)blockDim.x + threadIdx.x; x < total; x += (long long)gridDim.x * blockDim.x) {
        const long long pix = x / nbp;
        const int b = (int)(x - pix * nbp);
        double v = 0.0;
        if (b < nb) {
            double s = 0.0;
            for (long long k = pixptr[pix]; k < pixptr[pix + 1]; ++k) s += gw[gidx[k] + ngal * b];
            v = s * inv_nbar_domega - wrhatln[pix + npix * ((size_t)mode_l[b] + (size_t)(lmax + 1) * mode_n[b])];
        }
        maps[x] = v;
    }
}

// anlm[(n_b, l_b, m)] = alm_b[l_b, m], m = 0..l_b
__global__ void cat_extract_kernel(const double* __restrict__ alm, const int* __restrict__ mode_l,
                                   const int* __restrict__ mode_n, const long long* __restrict__ nbase, int nb, int nbp,
                                   int lmax, double* __restrict__ out) {
    const int b = blockIdx.x;
    if (b >= nb) return;
    const int l = mode_l[b], n = mode_n[b];
    for (int m = threadIdx.x; m <= l; m += blockDim.x) {
        const size_t lm = lm_mm(lmax, l, m);
        const long long i = nbase[n] + (long long)l * (l + 1) / 2 + m;
        out[2 * i] = alm[lm * 2 * nbp + b];
        out[2 * i + 1] = alm[lm * 2 * nbp + nbp + b];
    }
}

// ------------------------------------------------------------------------------------------------------------------
int sfbt_field2anlm(ShtPlan* sp, const double* d_field, int64_t ldw, const double* T, int64_t nmax, int64_t lmax,
                    const int64_t* nmax_l, const int64_t* lmax_n, double* d_alm, double* out_host, cudaStream_t st) {
    SFB_REQUIRE(sp && d_field && T && d_alm && out_host, "field2anlm: null pointer");
    SFB_REQUIRE(sp->lmax == lmax, "field2anlm: plan lmax mismatch");
    NlmTabs tb;
    SFB_TRY(make_nlm_tabs(&tb, nmax, lmax, nmax_l, lmax_n));
    DevBuf<double> dT, dout;
    DevBuf<int> dnl;
    DevBuf<long long> dnb;
    SFB_TRY(upload_radial_table(T, sp->nr, nmax, lmax, tb.nl, sp->nrp, dT, st));
    SFB_TRY(upload_vec(dnl, tb.nl, st));
    SFB_TRY(upload_vec(dnb, tb.nbase, st));
    SFB_TRY(dout.alloc((size_t)tb.nlmsize * 2));
    SFB_TRY(sht_map2alm(sp, d_field, ldw, 3, d_alm, st));      // map2alm!(map, alm): niter = 3
    nlm_contract_kernel<<<(unsigned)((lmax + 1) * (lmax + 2) / 2), 256, 0, st>>>(d_alm, dT.p, dnl.p, dnb.p, (int)lmax,
                                                                                  (int)nmax, sp->nrp, dout.p);
    SFB_CUDA_OK(cudaGetLastError());
    SFB_CUDA_OK(cudaMemcpyAsync(out_host, dout.p, (size_t)tb.nlmsize * 2 * sizeof(double), cudaMemcpyDeviceToHost, st));
    SFB_CUDA_OK(cudaStreamSynchronize(st));
    return 0;
}

int sfbt_anlm2field(ShtPlan* sp, const double* f_nlm_host, const double* g, int64_t nmax, int64_t lmax,
                    const int64_t* nmax_l, const int64_t* lmax_n, double* d_alm, double* d_map, double* out_host,
                    int64_t ld_out, cudaStream_t st) {
    SFB_REQUIRE(sp && f_nlm_host && g && d_alm && d_map && out_host, "anlm2field: null pointer");
    SFB_REQUIRE(sp->lmax == lmax && sp->nside_in == sp->nside, "anlm2field: plan mismatch");
    SFB_REQUIRE(ld_out >= sp->nr, "anlm2field: leading dimension of the field < nr");
    NlmTabs tb;
    SFB_TRY(make_nlm_tabs(&tb, nmax, lmax, nmax_l, lmax_n));
    DevBuf<double> dg, df;
    DevBuf<int> dnl;
    DevBuf<long long> dnb;
    SFB_TRY(upload_radial_table(g, sp->nr, nmax, lmax, tb.nl, sp->nrp, dg, st));
    SFB_TRY(upload_vec(dnl, tb.nl, st));
    SFB_TRY(upload_vec(dnb, tb.nbase, st));
    SFB_TRY(df.alloc((size_t)tb.nlmsize * 2));
    SFB_CUDA_OK(cudaMemcpyAsync(df.p, f_nlm_host, (size_t)tb.nlmsize * 2 * sizeof(double), cudaMemcpyHostToDevice, st));
    nlm_expand_kernel<<<dim3((unsigned)(lmax + 1), (unsigned)(lmax + 1)), 128, 0, st>>>(df.p, dg.p, dnl.p, dnb.p, (int)lmax,
                                                                                        (int)nmax, sp->nrp, d_alm);
    SFB_CUDA_OK(cudaGetLastError());
    SFB_TRY(sht_alm2map(sp, d_alm, d_map, st));
    // device [pixel][nrp] -> host nr x npix column-major (leading dimension ld_out)
    SFB_CUDA_OK(cudaMemcpy2DAsync(out_host, ld_out * sizeof(double), d_map, sp->nrp * sizeof(double), sp->nr * sizeof(double),
                                  (size_t)sp->npix, cudaMemcpyDeviceToHost, st));
    SFB_CUDA_OK(cudaStreamSynchronize(st));
    return 0;
}

int sfbt_win_rhat_ln(const double* d_win, int64_t ldw, int64_t npix, int64_t nr, const double* T, int64_t nmax,
                     int64_t lmax, const int64_t* nmax_l, const int64_t* lmax_n, double* d_out, cudaStream_t st) {
    SFB_REQUIRE(d_win && T && d_out, "win_rhat_ln: null pointer");
    NlmTabs tb;
    SFB_TRY(make_nlm_tabs(&tb, nmax, lmax, nmax_l, lmax_n));
    const int nrp = (int)round_up(nr, 8);
    DevBuf<double> dT;
    DevBuf<int> dnl;
    SFB_TRY(upload_radial_table(T, nr, nmax, lmax, tb.nl, nrp, dT, st));
    SFB_TRY(upload_vec(dnl, tb.nl, st));
    const int ncols = (int)((lmax + 1) * nmax);
    win_rhat_ln_kernel<<<dim3((unsigned)ceil_div(npix, 64), (unsigned)ceil_div(ncols, 64)), 256, 0, st>>>(
        d_win, ldw, npix, dT.p, dnl.p, (int)lmax, (int)nmax, (int)nr, nrp, d_out);
    SFB_CUDA_OK(cudaGetLastError());
    SFB_CUDA_OK(cudaStreamSynchronize(st));
    return 0;
}

int sfbt_cat2amln_batch(ShtPlan* sp, const int64_t* pixptr, const int64_t* gidx, int64_t ngal, const double* gw,
                        const int64_t* mode_n, const int64_t* mode_l, int64_t nb, double nbar, const double* d_wrhatln,
                        int64_t nmax, int64_t lmax, const int64_t* nmax_l, const int64_t* lmax_n, double* d_anlm,
                        cudaStream_t st) {
    SFB_REQUIRE(sp && pixptr && mode_n && mode_l && d_wrhatln && d_anlm, "cat2amln: null pointer");
    SFB_REQUIRE(ngal == 0 || (gidx && gw), "cat2amln: null catalogue arrays");
    SFB_REQUIRE(sp->lmax == lmax && sp->nr == nb && sp->nside_in == sp->nside, "cat2amln: plan mismatch");
    SFB_REQUIRE(nbar > 0, "cat2amln: nbar must be positive");
    NlmTabs tb;
    SFB_TRY(make_nlm_tabs(&tb, nmax, lmax, nmax_l, lmax_n));
    const int64_t npix = sp->npix;
    std::vector<int> ml(nb), mn(nb);
    for (int64_t b = 0; b < nb; ++b) {
        SFB_REQUIRE(mode_l[b] >= 0 && mode_l[b] <= lmax && mode_n[b] >= 1 && mode_n[b] <= tb.nl[mode_l[b]],
                    "cat2amln: (n, l) outside the mode set");
        ml[b] = (int)mode_l[b];
        mn[b] = (int)mode_n[b] - 1;
    }
    std::vector<long long> pp(pixptr, pixptr + npix + 1), gi(gidx, gidx + ngal);
    SFB_REQUIRE(pp[0] == 0 && pp[npix] == ngal, "cat2amln: pixel CSR does not cover the catalogue");
    DevBuf<long long> dpp, dgi, dnb;
    DevBuf<int> dml, dmn;
    DevBuf<double> dgw, dmaps, dalm;
    SFB_TRY(upload_vec(dpp, pp, st));
    SFB_TRY(upload_vec(dgi, gi, st));
    SFB_TRY(upload_vec(dnb, tb.nbase, st));
    SFB_TRY(upload_vec(dml, ml, st));
    SFB_TRY(upload_vec(dmn, mn, st));
    SFB_TRY(dgw.alloc((size_t)std::max<int64_t>(1, ngal * nb)));
    if (ngal > 0)
        SFB_CUDA_OK(cudaMemcpyAsync(dgw.p, gw, (size_t)ngal * nb * sizeof(double), cudaMemcpyHostToDevice, st));
    const int nbp = sp->nrp;
    SFB_TRY(dmaps.alloc((size_t)npix * nbp));
    SFB_TRY(dalm.alloc(sp->lmsize * 2 * nbp));
    const double inv = 1.0 / (nbar * (4.0 * 3.14159265358979323846 / (double)npix));
    cat_maps_kernel<<<2048, 256, 0, st>>>(dpp.p, dgi.p, dgw.p, ngal, d_wrhatln, dml.p, dmn.p, (int)nb, nbp, npix, (int)lmax,
                                          inv, dmaps.p);
    SFB_CUDA_OK(cudaGetLastError());
    SFB_TRY(sht_map2alm(sp, dmaps.p, nbp, 3, dalm.p, st));      // mymap2alm!(map, alm): lmax = amodes.lmax, niter = 3
    cat_extract_kernel<<<(unsigned)nb, 64, 0, st>>>(dalm.p, dml.p, dmn.p, dnb.p, (int)nb, nbp, (int)lmax, d_anlm);
    SFB_CUDA_OK(cudaGetLastError());
    SFB_CUDA_OK(cudaStreamSynchronize(st));
    return 0;
}

int64_t sfbt_nlmsize(int64_t nmax, const int64_t* lmax_n) {
    int64_t acc = 0;
    for (int64_t n = 0; n < nmax; ++n) acc += (lmax_n[n] + 1) * (lmax_n[n] + 2) / 2;
    return acc;
}

}  // namespace sfb
