// calc_wmix: see wmix.cu for the reference functions replaced.
#pragma once
#include "common.cuh"

namespace sfb {

// W_{nlm}^{n'l'm'} for m, m' >= 0 (neg_m: m -> -m) from the planar W_lm(r) of stage 1 (LMAX = 2 lmax, padded to
// nrp_alm shells).  G: host, nr x nmax x (lmax+1) column-major (rsdrgnlr); nmax_l[lmax+1], lmax_n[nmax]: the AnlmModes
// tables (src/modes.jl:92-105).  d_out: device, nlmsize x nlmsize ComplexF64 column-major, index (n,l,m) as
// getidx(amodes, n, l, m) (src/modes.jl:222-232).
int wmix_run(const double* d_alm, int nrp_alm, const double* G, int64_t nr, int64_t nmax, int64_t lmax,
             const int64_t* nmax_l, const int64_t* lmax_n, int neg_m, double* d_out, int64_t* nlmsize_out,
             cudaStream_t stream);

}  // namespace sfb
