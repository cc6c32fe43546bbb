// Shared device/host helpers for the sm_100a SuperFaB window coupling-matrix library.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <string>

namespace sfb {

// ---------------------------------------------------------------------------------------------
// error plumbing: every entry point returns 0 on success, nonzero + message otherwise
// (mirrors the reference's error()/@assert conventions: src/windows.jl:803,1013, src/healpix_helpers.jl:60-63)
void set_error(const std::string& msg);

#define SFB_CUDA_OK(expr)                                                                     \
    do {                                                                                      \
        cudaError_t e__ = (expr);                                                             \
        if (e__ != cudaSuccess) {                                                             \
            ::sfb::set_error(std::string(#expr) + ": " + cudaGetErrorString(e__) + " (" +     \
                             __FILE__ + ":" + std::to_string(__LINE__) + ")");                \
            return 1;                                                                         \
        }                                                                                     \
    } while (0)

#define SFB_REQUIRE(cond, msg)                                                                \
    do {                                                                                      \
        if (!(cond)) {                                                                        \
            ::sfb::set_error(std::string(msg));                                               \
            return 2;                                                                         \
        }                                                                                     \
    } while (0)

#define SFB_TRY(expr)                                                                         \
    do {                                                                                      \
        int r__ = (expr);                                                                     \
        if (r__ != 0) return r__;                                                             \
    } while (0)

// ---------------------------------------------------------------------------------------------
// simple owning device buffer
template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() { release(); }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        n = 0;
    }
    int alloc(size_t count) {
        if (count <= n && p) return 0;
        release();
        if (count == 0) count = 1;
        SFB_CUDA_OK(cudaMalloc(&p, count * sizeof(T)));
        n = count;
        return 0;
    }
};

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline int64_t round_up(int64_t a, int64_t b) { return ceil_div(a, b) * b; }

// ---------------------------------------------------------------------------------------------
// FP64 tensor-core MMA (DMMA.8x8x4 on sm_100a).
//   A 8x4 row-major: lane holds A[lane>>2][lane&3]
//   B 4x8 col-major: lane holds B[lane&3][lane>>2]
//   C 8x8:           lane holds C[lane>>2][2*(lane&3) + {0,1}]
#ifdef __CUDACC__
__device__ __forceinline__ void dmma884(double (&c)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c[0]), "+d"(c[1])
                 : "d"(a), "d"(b));
}

// Warp-level GEMM over shared-memory operands:
//   acc[i][j] (8x8 tiles) += A[(i*8..)+rows, k] * B[k, (j*8..)+cols],  k in [0, kt), kt % 4 == 0
//   A row-major with leading dimension lda, B stored [k][n] with leading dimension ldb.
//   lda, ldb ≡ 4 (mod 16) doubles makes both fragment loads bank-conflict free.
template <int MI, int NI>
__device__ __forceinline__ void warp_gemm_ss(double (&acc)[MI][NI][2], const double* __restrict__ As, int lda,
                                             const double* __restrict__ Bs, int ldb, int kt) {
    const int lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;
    const double* ap = As + g * lda + t;
    const double* bp = Bs + t * ldb + g;
#pragma unroll 2
    for (int k0 = 0; k0 < kt; k0 += 4) {
        double a[MI], b[NI];
#pragma unroll
        for (int i = 0; i < MI; ++i) a[i] = ap[i * 8 * lda + k0];
#pragma unroll
        for (int j = 0; j < NI; ++j) b[j] = bp[k0 * ldb + j * 8];
#pragma unroll
        for (int i = 0; i < MI; ++i)
#pragma unroll
            for (int j = 0; j < NI; ++j) dmma884(acc[i][j], a[i], b[j]);
    }
}

// Same, with A stored transposed in shared memory: A[m][k] = At[k * lda + m].
template <int MI, int NI>
__device__ __forceinline__ void warp_gemm_ts(double (&acc)[MI][NI][2], const double* __restrict__ At, int lda,
                                             const double* __restrict__ Bs, int ldb, int kt) {
    const int lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;
    const double* ap = At + t * lda + g;
    const double* bp = Bs + t * ldb + g;
#pragma unroll 2
    for (int k0 = 0; k0 < kt; k0 += 4) {
        double a[MI], b[NI];
#pragma unroll
        for (int i = 0; i < MI; ++i) a[i] = ap[k0 * lda + i * 8];
#pragma unroll
        for (int j = 0; j < NI; ++j) b[j] = bp[k0 * ldb + j * 8];
#pragma unroll
        for (int i = 0; i < MI; ++i)
#pragma unroll
            for (int j = 0; j < NI; ++j) dmma884(acc[i][j], a[i], b[j]);
    }
}
#endif

}  // namespace sfb
