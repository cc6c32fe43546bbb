// On-device window deconvolution: X = N \ B for the (binned) coupling matrix N left in HBM.
//
// Replaces the host step of the reference's pipeline
//   C    = bcmix \ (w̃mat * Cobs)            docs/src/tutorial_catalog.md:93-97   (Julia `\` = LU with partial pivoting)
//   wmat = inv(Nmix) * w̃M                   test/test_windows.jl:583-584
// so that only the band powers (LNN x nrhs) cross PCIe instead of the matrix.
//
// Blocked right-looking LU with partial pivoting on the augmented matrix [N | B] (column-major, in place), block size 64:
//   panel      one CTA per 8-column sub-panel: pivot search (block reduction), row swap inside it, scaling, rank-1 updates;
//              the other columns of the 64-column panel are updated by the parallel kernels below (two-level blocking)
//   swap       the panel's row interchanges applied to all other columns (left of the panel and right of it, B included)
//   trsm_l     U12 = L11^{-1} A12 (unit lower 64 x 64 in shared memory, one thread per column)
//   gemm       A22 -= L21 U12: the O(n³) part, FP64 tensor cores (DMMA), 64 x 64 tiles, K = 64
// then the back substitution U X = Y by 64-row blocks from the bottom (trsm_u + the same DMMA GEMM).
#include "lusolve.cuh"

#include <algorithm>
#include <vector>

namespace sfb {

constexpr int kNB = 64;

// ---- panel factorisation: columns [k0, k0+nb) and rows [k0, n) of A (ld = lda); piv[k0 + j] = pivot row of column k0 + j ----
__global__ void __launch_bounds__(1024) lu_panel_kernel(double* __restrict__ A, long long lda, int n, int k0, int nb,
                                                        int* __restrict__ piv, int* __restrict__ info) {
    __shared__ double s_val[32];
    __shared__ int s_idx[32];
    __shared__ double s_row[kNB];     // pivot row of the panel (columns j+1.. of it)
    __shared__ int s_p;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double* P = A + (size_t)k0 * lda;   // panel base: element (r, j) at P[r + j*lda]
    for (int j = 0; j < nb; ++j) {
        const int c = k0 + j;
        // 1. pivot: max |P[r, j]|, r in [c, n) (first maximum, like LAPACK's idamax)
        double best = -1.0;
        int bi = c;
        for (int r = c + tid; r < n; r += blockDim.x) {
            const double v = fabs(P[r + (size_t)j * lda]);
            if (v > best) {
                best = v;
                bi = r;
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
        if (lane == 0) {
            s_val[warp] = best;
            s_idx[warp] = bi;
        }
        __syncthreads();
        if (warp == 0) {
            best = (lane < (blockDim.x >> 5)) ? s_val[lane] : -1.0;
            bi = (lane < (blockDim.x >> 5)) ? s_idx[lane] : (1 << 30);
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                const double ob = __shfl_xor_sync(0xffffffffu, best, off);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
                if (ob > best || (ob == best && oi < bi)) {
                    best = ob;
                    bi = oi;
                }
            }
            if (lane == 0) {
                s_p = bi;
                piv[c] = bi;
                if (!(best > 0.0)) atomicMax(info, c + 1);   // exactly singular (or NaN) column
            }
        }
        __syncthreads();
        const int pr = s_p;
        // 2. swap rows c <-> pr inside the panel, keep the new pivot row in shared memory
        if (tid < nb) {
            const double a = P[c + (size_t)tid * lda], b = P[pr + (size_t)tid * lda];
            P[c + (size_t)tid * lda] = b;
            P[pr + (size_t)tid * lda] = a;
            s_row[tid] = b;
        }
        __syncthreads();
        const double inv = 1.0 / s_row[j];
        // 3. scale the column and update the remaining panel columns: thread per row
        for (int r = c + 1 + tid; r < n; r += blockDim.x) {
            const double lij = P[r + (size_t)j * lda] * inv;
            P[r + (size_t)j * lda] = lij;
            for (int jj = j + 1; jj < nb; ++jj) P[r + (size_t)jj * lda] -= lij * s_row[jj];
        }
        __syncthreads();
    }
}

// the panel's interchanges on the columns [c_lo, c_hi) (outside the panel): thread per column, swaps in order
__global__ void lu_swap_kernel(double* __restrict__ A, long long lda, const int* __restrict__ piv, int k0, int nb, int c_lo,
                               int c_hi) {
    const int c = c_lo + blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= c_hi) return;
    double* col = A + (size_t)c * lda;
    for (int j = 0; j < nb; ++j) {
        const int r = k0 + j, p = piv[r];
        if (p != r) {
            const double t = col[r];
            col[r] = col[p];
            col[p] = t;
        }
    }
}

// U12 = L11^{-1} A12: columns [c_lo, c_hi), rows [k0, k0+nb); L11 unit lower triangular in A[k0.., k0..]
__global__ void __launch_bounds__(128) lu_trsm_lower_kernel(double* __restrict__ A, long long lda, int k0, int nb, int c_lo,
                                                            int c_hi) {
    __shared__ double L[kNB][kNB + 1];
    for (int x = threadIdx.x; x < nb * nb; x += blockDim.x) {
        const int i = x % nb, j = x / nb;
        L[i][j] = A[(size_t)(k0 + i) + (size_t)(k0 + j) * lda];
    }
    __syncthreads();
    const int c = c_lo + blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= c_hi) return;
    double* col = A + (size_t)c * lda + k0;
    double x[kNB];
#pragma unroll
    for (int i = 0; i < kNB; ++i) x[i] = (i < nb) ? col[i] : 0.0;
#pragma unroll
    for (int i = 0; i < kNB; ++i) {
        if (i >= nb) break;
        const double xi = x[i];
#pragma unroll
        for (int r = i + 1; r < kNB; ++r)
            if (r < nb) x[r] -= L[r][i] * xi;
    }
#pragma unroll
    for (int i = 0; i < kNB; ++i)
        if (i < nb) col[i] = x[i];
}

// X_k = U_kk^{-1} Y_k: rows [k0, k0+nb) of the columns [c_lo, c_hi); U_kk upper triangular (non-unit) in A[k0.., k0..]
__global__ void __launch_bounds__(128) lu_trsm_upper_kernel(const double* __restrict__ A, long long lda, int k0, int nb,
                                                            double* __restrict__ B, long long ldb, int c_lo, int c_hi) {
    __shared__ double U[kNB][kNB + 1];
    for (int x = threadIdx.x; x < nb * nb; x += blockDim.x) {
        const int i = x % nb, j = x / nb;
        U[i][j] = A[(size_t)(k0 + i) + (size_t)(k0 + j) * lda];
    }
    __syncthreads();
    const int c = c_lo + blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= c_hi) return;
    double* col = B + (size_t)c * ldb + k0;
    double x[kNB];
#pragma unroll
    for (int i = 0; i < kNB; ++i) x[i] = (i < nb) ? col[i] : 0.0;
#pragma unroll
    for (int ii = 0; ii < kNB; ++ii) {
        const int i = kNB - 1 - ii;
        if (i >= nb) continue;
        const double xi = x[i] / U[i][i];
        x[i] = xi;
#pragma unroll
        for (int r = 0; r < kNB; ++r)
            if (r < i) x[r] -= U[r][i] * xi;
    }
#pragma unroll
    for (int i = 0; i < kNB; ++i)
        if (i < nb) col[i] = x[i];
}

// C[m0.., n0..] -= A[m0.., ka..ka+kc) * B[kb..kb+kc, n0..]   (all column-major)      CTA = 64 x 64 tile, 8 warps as 4 x 2, DMMA
__global__ void __launch_bounds__(256) lu_gemm_kernel(const double* __restrict__ Amat, long long lda,
                                                      const double* __restrict__ Bmat, long long ldb,
                                                      double* __restrict__ C, long long ldc, int m, int n, int kc) {
    constexpr int LDA = 72, LDB = 68;     // At[k][m] (transposed operand), Bs[k][n]
    __shared__ double As[32 * LDA];
    __shared__ double Bs[32 * LDB];
    const int m0 = blockIdx.x * 64, n0 = blockIdx.y * 64;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int wm = warp >> 1, wn = warp & 1;
    double acc[2][4][2];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    for (int k0 = 0; k0 < kc; k0 += 32) {
        {
            // A tile: At[kk][mm] = A[m0+mm, k0+kk]  (contiguous in mm)
            const int mm = tid & 63;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int kk = (tid >> 6) + 4 * q;
                As[kk * LDA + mm] = (m0 + mm < m && k0 + kk < kc) ? Amat[(size_t)(m0 + mm) + (size_t)(k0 + kk) * lda] : 0.0;
            }
            // B tile: Bs[kk][nn] = B[k0+kk, n0+nn]  (contiguous in kk)
            const int kk = tid & 31;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int nn = (tid >> 5) + 8 * q;
                Bs[kk * LDB + nn] = (n0 + nn < n && k0 + kk < kc) ? Bmat[(size_t)(k0 + kk) + (size_t)(n0 + nn) * ldb] : 0.0;
            }
        }
        __syncthreads();
        warp_gemm_ts<2, 4>(acc, As + wm * 16, LDA, Bs + wn * 32, LDB, 32);
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int row = m0 + wm * 16 + i * 8 + g;
        if (row >= m) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int col = n0 + wn * 32 + j * 8 + 2 * t + e;
                if (col < n) C[(size_t)row + (size_t)col * ldc] -= acc[i][j][e];
            }
    }
}

// [N | B] on the device: d_A is n x (n + nrhs) column-major with leading dimension lda >= n.  On return the first n columns
// hold the LU factors, the last nrhs columns hold X.  info > 0: column info-1 has no nonzero pivot (singular matrix).
int lu_solve_inplace(double* d_A, int64_t lda, int64_t n, int64_t nrhs, int* info_host, cudaStream_t st) {
    SFB_REQUIRE(d_A && n >= 1 && nrhs >= 0 && lda >= n && n < (int64_t(1) << 30), "lu_solve: bad arguments");
    DevBuf<int> d_piv, d_info;
    SFB_TRY(d_piv.alloc((size_t)n));
    SFB_TRY(d_info.alloc(1));
    SFB_CUDA_OK(cudaMemsetAsync(d_info.p, 0, sizeof(int), st));
    const int N = (int)n, NC = (int)(n + nrhs);
    // two-level blocking: the 64-column outer panel is factored in sub-panels of kNBi columns by the single-CTA kernel (whose
    // traffic is what one SM can pull from L2), the rest of the outer panel is updated by the parallel kernels; the O(n³)
    // trailing update outside the panel then runs with K = 64.
    constexpr int kNBi = 8;
    for (int k0 = 0; k0 < N; k0 += kNB) {
        const int nb = std::min(kNB, N - k0);
        const int pend = k0 + nb;                                  // end of the outer panel
        for (int i0 = k0; i0 < pend; i0 += kNBi) {
            const int ib = std::min(kNBi, pend - i0);
            lu_panel_kernel<<<1, 1024, 0, st>>>(d_A, lda, N, i0, ib, d_piv.p, d_info.p);
            // interchanges + updates on the other columns of the outer panel
            if (i0 > k0) lu_swap_kernel<<<(unsigned)ceil_div(i0 - k0, 64), 64, 0, st>>>(d_A, lda, d_piv.p, i0, ib, k0, i0);
            const int c1 = i0 + ib;
            if (c1 < pend) {
                lu_swap_kernel<<<(unsigned)ceil_div(pend - c1, 64), 64, 0, st>>>(d_A, lda, d_piv.p, i0, ib, c1, pend);
                lu_trsm_lower_kernel<<<(unsigned)ceil_div(pend - c1, 128), 128, 0, st>>>(d_A, lda, i0, ib, c1, pend);
                const int m = N - c1;
                if (m > 0)
                    lu_gemm_kernel<<<dim3((unsigned)ceil_div(m, 64), (unsigned)ceil_div(pend - c1, 64)), 256, 0, st>>>(
                        d_A + c1 + (size_t)i0 * lda, lda, d_A + i0 + (size_t)c1 * lda, lda, d_A + c1 + (size_t)c1 * lda, lda, m,
                        pend - c1, ib);
            }
        }
        if (k0 > 0) lu_swap_kernel<<<(unsigned)ceil_div(k0, 256), 256, 0, st>>>(d_A, lda, d_piv.p, k0, nb, 0, k0);
        const int c_lo = pend;
        if (c_lo < NC) {
            lu_swap_kernel<<<(unsigned)ceil_div(NC - c_lo, 256), 256, 0, st>>>(d_A, lda, d_piv.p, k0, nb, c_lo, NC);
            lu_trsm_lower_kernel<<<(unsigned)ceil_div(NC - c_lo, 128), 128, 0, st>>>(d_A, lda, k0, nb, c_lo, NC);
            const int m = N - c_lo;
            if (m > 0)
                lu_gemm_kernel<<<dim3((unsigned)ceil_div(m, 64), (unsigned)ceil_div(NC - c_lo, 64)), 256, 0, st>>>(
                    d_A + c_lo + (size_t)k0 * lda, lda, d_A + k0 + (size_t)c_lo * lda, lda, d_A + c_lo + (size_t)c_lo * lda,
                    lda, m, NC - c_lo, nb);
        }
        SFB_CUDA_OK(cudaGetLastError());
    }
    // back substitution U X = Y on the right-hand-side columns, 64-row blocks from the bottom
    if (nrhs > 0) {
        double* B = d_A + (size_t)n * lda;
        const int R = (int)nrhs;
        const int last = ((N - 1) / kNB) * kNB;
        for (int k0 = last; k0 >= 0; k0 -= kNB) {
            const int nb = std::min(kNB, N - k0);
            lu_trsm_upper_kernel<<<(unsigned)ceil_div(R, 128), 128, 0, st>>>(d_A, lda, k0, nb, B, lda, 0, R);
            if (k0 > 0)   // Y[0:k0) -= U[0:k0, k0:k0+nb) X_k
                lu_gemm_kernel<<<dim3((unsigned)ceil_div(k0, 64), (unsigned)ceil_div(R, 64)), 256, 0, st>>>(
                    d_A + (size_t)k0 * lda, lda, B + k0, lda, B, lda, k0, R, nb);
            SFB_CUDA_OK(cudaGetLastError());
        }
    }
    int info = 0;
    SFB_CUDA_OK(cudaMemcpyAsync(&info, d_info.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    SFB_CUDA_OK(cudaStreamSynchronize(st));
    if (info_host) *info_host = info;
    if (info > 0) {   // LinearAlgebra.SingularException(info) in Julia
        set_error("SingularException(" + std::to_string(info) + "): the coupling matrix is singular to working precision");
        return 5;
    }
    return 0;
}

}  // namespace sfb
