// Band-power binning epilogue N = w̃ · M · v and the separable-window coupling matrix.
//
// Replaces (reference hsgg/SphericalFourierBesselDecompositions.jl):
//   _power_win_mix (dense)      src/windows.jl:825-862   sparse double sum over nz(w̃[n,:]) x nz(v[:,m])
//   _power_win_mix (separable)  src/windows.jl:942-990
//   calc_angular_mixing_matrix  src/windows.jl:866-878
//   calc_radial_mixing          src/windows.jl:924-938
//   calc_cmixii_separable       src/windows.jl:651-679
#pragma once
#include "cmix.cuh"
#include "common.cuh"

namespace sfb {

// d_M: n x n (device, column-major).  w̃ / v: host CSC (1-based colptr/rowval, Julia SparseMatrixCSC);
// NULL colptr = UniformScaling I.  N_out: host, LNN1 x LNN2 column-major.
int binned_product_to_host(const double* d_M, int64_t n, const int64_t* wt_colptr, const int64_t* wt_rowval,
                           const double* wt_nzval, int64_t LNN1, const int64_t* v_colptr, const int64_t* v_rowval,
                           const double* v_nzval, int64_t LNN2, double* N_out, float* ms);

int binned_product_device(const double* d_M, int64_t n, const int64_t* wt_colptr, const int64_t* wt_rowval,
                          const double* wt_nzval, int64_t LNN1, const int64_t* v_colptr, const int64_t* v_rowval,
                          const double* v_nzval, int64_t LNN2, double* d_N, int64_t ldN, float* ms);

// Coupling matrix of a separable window W(r, n̂) = phi(r) mask(n̂): d_wlm planar alm of the mask (1 shell,
// padded to nrp_s), phi host vector of length nr.  Writes the full nout x nout matrix to d_M.
int separable_cmix(CmixPlan* p, const double* d_wlm, int nrp_s, const double* phi, int div2Lp1, int interchange,
                   double* d_M);

}  // namespace sfb
