// See binned.cuh for the reference functions replaced.
#include "binned.cuh"

#include <algorithm>
#include <vector>

namespace sfb {

// N[I, m] = Σ_{i' ∈ nz(v[:,m])} v[i',m] Σ_{i ∈ nz(w̃[I,:])} w̃[I,i] M[i,i']     (src/windows.jl:847-853)
// thread = (I, m); with disjoint bins every element of M is read exactly once.
__global__ void binned_product_kernel(const double* __restrict__ M, long long n, const int* __restrict__ wptr,
                                      const int* __restrict__ wcol, const double* __restrict__ wval, int LNN1,
                                      const int* __restrict__ vptr, const int* __restrict__ vrow,
                                      const double* __restrict__ vval, int LNN2, double* __restrict__ N, long long ldN) {
    const int I = blockIdx.x * blockDim.x + threadIdx.x;
    const int m = blockIdx.y;
    if (I >= LNN1) return;
    double c = 0.0;
    const int w0 = wptr ? wptr[I] : I, w1 = wptr ? wptr[I + 1] : I + 1;
    const int v0 = vptr ? vptr[m] : m, v1 = vptr ? vptr[m + 1] : m + 1;
    for (int a = w0; a < w1; ++a) {
        const long long i = wptr ? wcol[a] : a;
        const double wv = wptr ? wval[a] : 1.0;
        if (wv == 0.0) continue;
        for (int b = v0; b < v1; ++b) {
            const long long ip = vptr ? vrow[b] : b;
            const double vv = vptr ? vval[b] : 1.0;
            if (vv == 0.0) continue;
            c += wv * vv * M[i + ip * n];
        }
    }
    N[I + (size_t)m * ldN] = c;
}

__global__ void nonfinite_flag_kernel(const double* __restrict__ x, size_t n, int* flag) {
    int bad = 0;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        bad |= !isfinite(x[i]);
    if (bad) atomicOr(flag, 1);
}

__global__ void nonfinite_flag_2d_kernel(const double* __restrict__ x, long long rows, long long cols, long long ld,
                                         int* flag) {
    int bad = 0;
    const long long n = rows * cols;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        bad |= !isfinite(x[(i % rows) + (i / rows) * ld]);
    if (bad) atomicOr(flag, 1);
}

// Device scratch of the binned products, kept per device across calls: cudaMalloc / cudaFree of the 246 MB result and the
// binning tables on every call made the host-pointer binned call erratic (10 - 700 ms for the same work at cfg4).
struct BinScratch {
    DevBuf<double> d_N, d_wval, d_vval;
    DevBuf<int> d_wptr, d_wcol, d_vptr, d_vrow;
};
static int bin_scratch(BinScratch** out) {
    static BinScratch scratch[64];
    int dev = 0;
    SFB_CUDA_OK(cudaGetDevice(&dev));
    SFB_REQUIRE(dev >= 0 && dev < 64, "binned_product: device index out of range");
    *out = &scratch[dev];
    return 0;
}

int binned_product_to_host(const double* d_M, int64_t n, const int64_t* wt_colptr, const int64_t* wt_rowval,
                           const double* wt_nzval, int64_t LNN1, const int64_t* v_colptr, const int64_t* v_rowval,
                           const double* v_nzval, int64_t LNN2, double* N_out, float* ms,
                           int (*d2h)(void* dst, const void* src, size_t bytes)) {
    SFB_REQUIRE(d_M && N_out, "binned_product: null pointer");
    if (!wt_colptr) SFB_REQUIRE(LNN1 == n, "w̃ = I requires LNN1 == lnnsize");
    if (!v_colptr) SFB_REQUIRE(LNN2 == n, "v = I requires LNN2 == lnnsize");
    if (!wt_colptr && !v_colptr) {
        SFB_CUDA_OK(cudaMemcpy(N_out, d_M, (size_t)n * n * sizeof(double), cudaMemcpyDeviceToHost));
        if (ms) *ms = 0;
        return 0;
    }
    BinScratch* bs = nullptr;
    SFB_TRY(bin_scratch(&bs));
    DevBuf<double>& d_N = bs->d_N;
    SFB_TRY(d_N.alloc((size_t)LNN1 * LNN2));
    SFB_TRY(binned_product_device(d_M, n, wt_colptr, wt_rowval, wt_nzval, LNN1, v_colptr, v_rowval, v_nzval, LNN2, d_N.p,
                                  LNN1, ms));
    if (d2h) return d2h(N_out, d_N.p, (size_t)LNN1 * LNN2 * sizeof(double));
    SFB_CUDA_OK(cudaMemcpy(N_out, d_N.p, (size_t)LNN1 * LNN2 * sizeof(double), cudaMemcpyDeviceToHost));
    return 0;
}

// N = w̃ M v left on the device: d_N is LNN1 x LNN2 column-major with leading dimension ldN (the non-finite check of
// src/windows.jl:1013 included).  w̃ = v = I copies M.
int binned_product_device(const double* d_M, int64_t n, const int64_t* wt_colptr, const int64_t* wt_rowval,
                          const double* wt_nzval, int64_t LNN1, const int64_t* v_colptr, const int64_t* v_rowval,
                          const double* v_nzval, int64_t LNN2, double* d_Nout, int64_t ldN, float* ms) {
    SFB_REQUIRE(d_M && d_Nout && ldN >= LNN1, "binned_product: bad arguments");
    if (!wt_colptr) SFB_REQUIRE(LNN1 == n, "w̃ = I requires LNN1 == lnnsize");
    if (!v_colptr) SFB_REQUIRE(LNN2 == n, "v = I requires LNN2 == lnnsize");
    if (!wt_colptr && !v_colptr) {
        SFB_CUDA_OK(cudaMemcpy2D(d_Nout, ldN * sizeof(double), d_M, n * sizeof(double), n * sizeof(double), (size_t)n,
                                 cudaMemcpyDeviceToDevice));
        if (ms) *ms = 0;
        return 0;
    }
    BinScratch* bs = nullptr;
    SFB_TRY(bin_scratch(&bs));
    DevBuf<int>&d_wptr = bs->d_wptr, &d_wcol = bs->d_wcol, &d_vptr = bs->d_vptr, &d_vrow = bs->d_vrow;
    DevBuf<double>&d_wval = bs->d_wval, &d_vval = bs->d_vval;
    if (wt_colptr) {
        // CSC (columns i) -> CSR (rows I), keeping ascending i within a row like the reference's nzind order
        SFB_REQUIRE(wt_rowval && wt_nzval, "w̃: null rowval/nzval");
        const int64_t nnz = wt_colptr[n] - 1;
        std::vector<int> ptr(LNN1 + 1, 0), col(nnz);
        std::vector<double> val(nnz);
        for (int64_t k = 0; k < nnz; ++k) {
            const int64_t I = wt_rowval[k] - 1;
            SFB_REQUIRE(I >= 0 && I < LNN1, "w̃: row index out of range");
            ptr[I + 1]++;
        }
        for (int64_t I = 0; I < LNN1; ++I) ptr[I + 1] += ptr[I];
        std::vector<int> fill(ptr.begin(), ptr.end() - 1);
        for (int64_t i = 0; i < n; ++i)
            for (int64_t k = wt_colptr[i] - 1; k < wt_colptr[i + 1] - 1; ++k) {
                const int slot = fill[wt_rowval[k] - 1]++;
                col[slot] = (int)i;
                val[slot] = wt_nzval[k];
            }
        SFB_TRY(d_wptr.alloc(ptr.size()));
        SFB_TRY(d_wcol.alloc(col.size()));
        SFB_TRY(d_wval.alloc(val.size()));
        SFB_CUDA_OK(cudaMemcpy(d_wptr.p, ptr.data(), ptr.size() * sizeof(int), cudaMemcpyHostToDevice));
        SFB_CUDA_OK(cudaMemcpy(d_wcol.p, col.data(), col.size() * sizeof(int), cudaMemcpyHostToDevice));
        SFB_CUDA_OK(cudaMemcpy(d_wval.p, val.data(), val.size() * sizeof(double), cudaMemcpyHostToDevice));
    }
    if (v_colptr) {
        SFB_REQUIRE(v_rowval && v_nzval, "v: null rowval/nzval");
        const int64_t nnz = v_colptr[LNN2] - 1;
        std::vector<int> ptr(LNN2 + 1), row(nnz);
        for (int64_t m = 0; m <= LNN2; ++m) ptr[m] = (int)(v_colptr[m] - 1);
        for (int64_t k = 0; k < nnz; ++k) {
            SFB_REQUIRE(v_rowval[k] >= 1 && v_rowval[k] <= n, "v: row index out of range");
            row[k] = (int)(v_rowval[k] - 1);
        }
        SFB_TRY(d_vptr.alloc(ptr.size()));
        SFB_TRY(d_vrow.alloc(row.size()));
        SFB_TRY(d_vval.alloc(nnz));
        SFB_CUDA_OK(cudaMemcpy(d_vptr.p, ptr.data(), ptr.size() * sizeof(int), cudaMemcpyHostToDevice));
        SFB_CUDA_OK(cudaMemcpy(d_vrow.p, row.data(), row.size() * sizeof(int), cudaMemcpyHostToDevice));
        SFB_CUDA_OK(cudaMemcpy(d_vval.p, v_nzval, nnz * sizeof(double), cudaMemcpyHostToDevice));
    }
    cudaEvent_t e0, e1;
    SFB_CUDA_OK(cudaEventCreate(&e0));
    SFB_CUDA_OK(cudaEventCreate(&e1));
    SFB_CUDA_OK(cudaEventRecord(e0));
    SFB_REQUIRE(LNN2 <= 65535 * 32768LL, "LNN2 too large");
    // grid.y is limited to 65535: loop over column slabs
    for (int64_t m0 = 0; m0 < LNN2; m0 += 65535) {
        const int64_t mc = std::min<int64_t>(65535, LNN2 - m0);
        binned_product_kernel<<<dim3((unsigned)ceil_div(LNN1, 128), (unsigned)mc), 128>>>(
            d_M, n, wt_colptr ? d_wptr.p : nullptr, d_wcol.p, d_wval.p, (int)LNN1,
            v_colptr ? d_vptr.p + m0 : nullptr, d_vrow.p, d_vval.p, (int)LNN2, d_Nout + (size_t)m0 * ldN, ldN);
        SFB_CUDA_OK(cudaGetLastError());
        SFB_REQUIRE(v_colptr || m0 == 0, "v = I with LNN2 > 65535 is not supported");
    }
    SFB_CUDA_OK(cudaEventRecord(e1));
    SFB_CUDA_OK(cudaEventSynchronize(e1));
    float t = 0;
    cudaEventElapsedTime(&t, e0, e1);
    if (ms) *ms = t;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    {  // @assert all(isfinite.(mix))  src/windows.jl:1013 — checked on the device
        DevBuf<int> flag;
        SFB_TRY(flag.alloc(1));
        SFB_CUDA_OK(cudaMemset(flag.p, 0, sizeof(int)));
        nonfinite_flag_2d_kernel<<<512, 256>>>(d_Nout, LNN1, LNN2, ldN, flag.p);
        int h = 0;
        SFB_CUDA_OK(cudaMemcpy(&h, flag.p, sizeof(int), cudaMemcpyDeviceToHost));
        if (h) {
            set_error("AssertionError: all(isfinite.(mix))");
            return 4;
        }
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------
// column-restricted form (multi-device runs: every device bins the column slab of M it formed)

int bin_tables_build(BinTables* t, int64_t n, const int64_t* wt_colptr, const int64_t* wt_rowval, const double* wt_nzval,
                     int64_t LNN1, const int64_t* v_colptr, const int64_t* v_rowval, const double* v_nzval, int64_t LNN2) {
    t->n = n;
    t->LNN1 = LNN1;
    t->LNN2 = LNN2;
    t->has_w = wt_colptr != nullptr;
    t->has_v = v_colptr != nullptr;
    if (!t->has_w) SFB_REQUIRE(LNN1 == n, "w̃ = I requires LNN1 == lnnsize");
    if (!t->has_v) SFB_REQUIRE(LNN2 == n, "v = I requires LNN2 == lnnsize");
    if (t->has_w) {
        SFB_REQUIRE(wt_rowval && wt_nzval, "w̃: null rowval/nzval");
        const int64_t nnz = wt_colptr[n] - 1;
        t->wptr.assign(LNN1 + 1, 0);
        t->wcol.resize(nnz);
        t->wval.resize(nnz);
        for (int64_t k = 0; k < nnz; ++k) {
            const int64_t I = wt_rowval[k] - 1;
            SFB_REQUIRE(I >= 0 && I < LNN1, "w̃: row index out of range");
            t->wptr[I + 1]++;
        }
        for (int64_t I = 0; I < LNN1; ++I) t->wptr[I + 1] += t->wptr[I];
        std::vector<int> fill(t->wptr.begin(), t->wptr.end() - 1);
        for (int64_t i = 0; i < n; ++i)
            for (int64_t k = wt_colptr[i] - 1; k < wt_colptr[i + 1] - 1; ++k) {
                const int slot = fill[wt_rowval[k] - 1]++;
                t->wcol[slot] = (int)i;
                t->wval[slot] = wt_nzval[k];
            }
    }
    if (t->has_v) {
        SFB_REQUIRE(v_rowval && v_nzval, "v: null rowval/nzval");
        const int64_t nnz = v_colptr[LNN2] - 1;
        t->vptr.resize(LNN2 + 1);
        t->vrow.resize(nnz);
        t->vval.assign(v_nzval, v_nzval + nnz);
        for (int64_t m = 0; m <= LNN2; ++m) t->vptr[m] = (int)(v_colptr[m] - 1);
        for (int64_t k = 0; k < nnz; ++k) {
            SFB_REQUIRE(v_rowval[k] >= 1 && v_rowval[k] <= n, "v: row index out of range");
            t->vrow[k] = (int)(v_rowval[k] - 1);
        }
    }
    return 0;
}

template <typename T>
static int up_async(DevBuf<T>& d, const std::vector<T>& h, cudaStream_t st) {
    SFB_TRY(d.alloc(std::max<size_t>(1, h.size())));
    if (!h.empty()) SFB_CUDA_OK(cudaMemcpyAsync(d.p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice, st));
    return 0;
}

int bin_tables_upload(const BinTables& t, BinDev& d, cudaStream_t st) {
    SFB_TRY(up_async(d.wptr, t.wptr, st));
    SFB_TRY(up_async(d.wcol, t.wcol, st));
    SFB_TRY(up_async(d.wval, t.wval, st));
    SFB_TRY(up_async(d.vptr, t.vptr, st));
    SFB_TRY(up_async(d.vrow, t.vrow, st));
    SFB_TRY(up_async(d.vval, t.vval, st));
    return 0;
}

void bin_needed_cols(const BinTables& t, int64_t J0, int64_t J1, int64_t* c0, int64_t* c1) {
    if (!t.has_v) {
        *c0 = J0;
        *c1 = J1;
        return;
    }
    int64_t lo = t.n, hi = 0;
    for (int k = t.vptr[J0]; k < t.vptr[J1]; ++k) {
        lo = std::min<int64_t>(lo, t.vrow[k]);
        hi = std::max<int64_t>(hi, t.vrow[k] + 1);
    }
    if (hi <= lo) lo = hi = 0;
    *c0 = lo;
    *c1 = hi;
}

int binned_product_range(const double* d_Mslab, int64_t c0, const BinTables& t, const BinDev& d, int64_t J0, int64_t J1,
                         double* d_N, int64_t ldN, cudaStream_t st) {
    if (J1 <= J0) return 0;
    SFB_REQUIRE(d_Mslab && d_N && ldN >= t.LNN1, "binned_product_range: bad arguments");
    const double* Mbase = d_Mslab - c0 * t.n;      // virtual base of the full matrix: only columns >= c0 are dereferenced
    for (int64_t m0 = J0; m0 < J1; m0 += 65535) {
        const int64_t mc = std::min<int64_t>(65535, J1 - m0);
        if (t.has_v) {
            binned_product_kernel<<<dim3((unsigned)ceil_div(t.LNN1, 128), (unsigned)mc), 128, 0, st>>>(
                Mbase, t.n, t.has_w ? d.wptr.p : nullptr, d.wcol.p, d.wval.p, (int)t.LNN1, d.vptr.p + m0, d.vrow.p, d.vval.p,
                (int)t.LNN2, d_N + (size_t)(m0 - J0) * ldN, ldN);
        } else {   // v = I: output column J reads M column J
            binned_product_kernel<<<dim3((unsigned)ceil_div(t.LNN1, 128), (unsigned)mc), 128, 0, st>>>(
                Mbase + m0 * t.n, t.n, t.has_w ? d.wptr.p : nullptr, d.wcol.p, d.wval.p, (int)t.LNN1, nullptr, nullptr, nullptr,
                (int)t.LNN2, d_N + (size_t)(m0 - J0) * ldN, ldN);
        }
        SFB_CUDA_OK(cudaGetLastError());
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------
// separable window

// C_L = (Re a_L0 conj a_L0 + 2 Σ_{M>=1} |a_LM|²)/(2L+1)  (Healpix.alm2cl; auto-spectrum of the mask)
__global__ void alm2cl_kernel(const double* __restrict__ wlm, int nrp_s, int LMAX, double* __restrict__ cl) {
    const int L = blockIdx.x * blockDim.x + threadIdx.x;
    if (L > LMAX) return;
    double c = 0.0;
    for (int M = 0; M <= L; ++M) {
        const size_t lm = (size_t)L + ((size_t)M * (2 * LMAX + 1 - M)) / 2;
        const double re = wlm[lm * 2 * nrp_s], im = wlm[lm * 2 * nrp_s + nrp_s];
        const double v = re * re + im * im;
        c += (M == 0) ? v : 2.0 * v;
    }
    cl[L] = c / (2.0 * L + 1.0);
}

// ang[l,L] = 1/(4π) Σ_{L1} (l L L1;000)² (2L1+1) C_{L1}      (src/windows.jl:866-878)
__global__ void ang_mix_kernel(const double* __restrict__ w2, const double* __restrict__ cl, int lmax,
                               double* __restrict__ ang) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int np = (lmax + 1) * (lmax + 1);
    if (x >= np) return;
    const int l = x / (lmax + 1), L = x % (lmax + 1);
    const int lo = min(l, L), d = abs(l - L);
    const double* w = w2 + (size_t)x * (lmax + 1);
    double s = 0.0;
    for (int k = 0; k <= lo; ++k) {
        const int L1 = d + 2 * k;
        s += w[k] * (2.0 * L1 + 1.0) * cl[L1];
    }
    ang[x] = s * 0.07957747154594767;
}

// R[(l,n),(L,N)] = Σ_r G_ln[r] G_LN[r] phi[r]                 (src/windows.jl:924-938 with r=Δr=1)
__global__ void radial_mix_kernel(const double* __restrict__ G, const double* __restrict__ phi, int nln, int nrp,
                                  double* __restrict__ R) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= nln) return;
    const double* gx = G + (size_t)x * nrp;
    const double* gy = G + (size_t)y * nrp;
    double s = 0.0;
    for (int r = 0; r < nrp; ++r) s += gx[r] * gy[r] * phi[r];
    R[(size_t)y * nln + x] = s;
}

// M[i,i'] per src/windows.jl:651-679
__global__ void separable_cmix_kernel(const double* __restrict__ R, const double* __restrict__ ang,
                                      const int* __restrict__ out_l, const int* __restrict__ out_n,
                                      const int* __restrict__ out_n2, int nout, int lmax, int nmax, int div2Lp1,
                                      int interchange, double* __restrict__ M) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x, ip = blockIdx.y;
    if (i >= nout) return;
    const int l = out_l[i], n = out_n[i], n2 = out_n2[i];
    const int L = out_l[ip];
    int N = out_n[ip], N2 = out_n2[ip];
    if (interchange) {
        const int tmp = N;
        N = N2;
        N2 = tmp;
    }
    const size_t nln = (size_t)(lmax + 1) * nmax;
    const size_t a = (size_t)l * nmax + n, a2 = (size_t)l * nmax + n2;
    const size_t b = (size_t)L * nmax + N, b2 = (size_t)L * nmax + N2;
    const double mang = ang[(size_t)l * (lmax + 1) + L];
    double mix = mang * R[b * nln + a] * R[b2 * nln + a2];
    if (!interchange && N != N2) mix += mang * R[b2 * nln + a] * R[b * nln + a2];
    if (!div2Lp1) mix *= (2.0 * L + 1.0);
    M[i + (size_t)ip * nout] = mix;
}

int separable_cmix(CmixPlan* p, const double* d_wlm, int nrp_s, const double* phi, int div2Lp1, int interchange,
                   double* d_M) {
    SFB_REQUIRE(p && d_wlm && phi && d_M, "separable_cmix: null pointer");
    const int lmax = p->lmax, nmax = p->nmax, LMAX = p->LMAX;
    const int nln = (lmax + 1) * nmax;
    const int nout = (int)p->nout;
    SFB_REQUIRE(nout <= 65535 * 1024, "separable_cmix: matrix too large");
    std::vector<double> phip(p->nrp, 0.0);
    for (int r = 0; r < p->nr; ++r) phip[r] = phi[r];
    std::vector<int> ol(nout), on(nout), on2(nout);
    for (int l = 0; l <= lmax; ++l)
        for (int s = p->ell_ptr[l]; s < p->ell_ptr[l + 1]; ++s) {
            const int o = p->h_row_out[s];
            ol[o] = l;
            on[o] = p->h_row_n[s];
            on2[o] = p->h_row_n2[s];
        }
    DevBuf<double> d_phi, d_cl, d_ang, d_R;
    DevBuf<int> d_ol, d_on, d_on2;
    SFB_TRY(d_phi.alloc(p->nrp));
    SFB_TRY(d_cl.alloc(LMAX + 1));
    SFB_TRY(d_ang.alloc((size_t)(lmax + 1) * (lmax + 1)));
    SFB_TRY(d_R.alloc((size_t)nln * nln));
    SFB_TRY(d_ol.alloc(nout));
    SFB_TRY(d_on.alloc(nout));
    SFB_TRY(d_on2.alloc(nout));
    SFB_CUDA_OK(cudaMemcpy(d_phi.p, phip.data(), p->nrp * sizeof(double), cudaMemcpyHostToDevice));
    SFB_CUDA_OK(cudaMemcpy(d_ol.p, ol.data(), nout * sizeof(int), cudaMemcpyHostToDevice));
    SFB_CUDA_OK(cudaMemcpy(d_on.p, on.data(), nout * sizeof(int), cudaMemcpyHostToDevice));
    SFB_CUDA_OK(cudaMemcpy(d_on2.p, on2.data(), nout * sizeof(int), cudaMemcpyHostToDevice));
    cudaEvent_t e0, e1;
    SFB_CUDA_OK(cudaEventCreate(&e0));
    SFB_CUDA_OK(cudaEventCreate(&e1));
    SFB_CUDA_OK(cudaEventRecord(e0));
    alm2cl_kernel<<<(unsigned)ceil_div(LMAX + 1, 128), 128>>>(d_wlm, nrp_s, LMAX, d_cl.p);
    ang_mix_kernel<<<(unsigned)ceil_div((lmax + 1) * (lmax + 1), 128), 128>>>(p->d_w2.p, d_cl.p, lmax, d_ang.p);
    radial_mix_kernel<<<dim3((unsigned)ceil_div(nln, 128), nln), 128>>>(p->d_G.p, d_phi.p, nln, p->nrp, d_R.p);
    for (int c0 = 0; c0 < nout; c0 += 65535) {
        // grid.y limit: slabs of columns (kernel indexes columns through blockIdx.y + implicit offset)
        SFB_REQUIRE(c0 == 0, "separable_cmix: lnnsize > 65535 not supported");
        separable_cmix_kernel<<<dim3((unsigned)ceil_div(nout, 128), (unsigned)std::min(nout, 65535)), 128>>>(
            d_R.p, d_ang.p, d_ol.p, d_on.p, d_on2.p, nout, lmax, nmax, div2Lp1, interchange, d_M);
    }
    SFB_CUDA_OK(cudaGetLastError());
    SFB_CUDA_OK(cudaEventRecord(e1));
    SFB_CUDA_OK(cudaEventSynchronize(e1));
    cudaEventElapsedTime(&p->t_block, e0, e1);
    p->t_wl = p->t_what = 0;
    p->launches = 4;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return 0;
}

}  // namespace sfb
