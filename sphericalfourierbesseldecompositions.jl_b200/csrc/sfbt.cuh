// SFB transforms next to the window path: see sfbt.cu for the reference functions replaced.
#pragma once
#include "common.cuh"
#include "sht.cuh"

namespace sfb {

// All radial tables are host arrays nr x nmax x (lmax+1), column-major, NaN where n > nmax_l[l] (never read).
// sp: stage-1 plan with lmax = amodes.lmax and nr shells (cat2amln: nr = number of (n,l) modes in the batch).
int sfbt_field2anlm(ShtPlan* sp, const double* d_field, int64_t ldw, const double* T, int64_t nmax, int64_t lmax,
                    const int64_t* nmax_l, const int64_t* lmax_n, double* d_alm, double* out_host, cudaStream_t st);
int sfbt_anlm2field(ShtPlan* sp, const double* f_nlm_host, const double* g, int64_t nmax, int64_t lmax,
                    const int64_t* nmax_l, const int64_t* lmax_n, double* d_alm, double* d_map, double* out_host,
                    int64_t ld_out, cudaStream_t st);
int sfbt_win_rhat_ln(const double* d_win, int64_t ldw, int64_t npix, int64_t nr, const double* T, int64_t nmax,
                     int64_t lmax, const int64_t* nmax_l, const int64_t* lmax_n, double* d_out, cudaStream_t st);
int sfbt_cat2amln_batch(ShtPlan* sp, const int64_t* pixptr, const int64_t* gidx, int64_t ngal, const double* gw,
                        const int64_t* mode_n, const int64_t* mode_l, int64_t nb, double nbar, const double* d_wrhatln,
                        int64_t nmax, int64_t lmax, const int64_t* nmax_l, const int64_t* lmax_n, double* d_anlm,
                        cudaStream_t st);
int64_t sfbt_nlmsize(int64_t nmax, const int64_t* lmax_n);

}  // namespace sfb
