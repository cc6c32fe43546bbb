// On-device deconvolution X = N \\ B: see lusolve.cu.
#pragma once
#include "common.cuh"

namespace sfb {

// d_A: n x (n + nrhs) column-major (leading dimension lda), [N | B] on entry, [LU | X] on return.
int lu_solve_inplace(double* d_A, int64_t lda, int64_t n, int64_t nrhs, int* info_host, cudaStream_t stream);

}  // namespace sfb
