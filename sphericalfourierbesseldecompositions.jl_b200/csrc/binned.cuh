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
#include <vector>

namespace sfb {

// d_M: n x n (device, column-major).  w̃ / v: host CSC (1-based colptr/rowval, Julia SparseMatrixCSC);
// NULL colptr = UniformScaling I.  N_out: host, LNN1 x LNN2 column-major.
// d2h (optional): device -> host copy to use for the result (the library-staged copy for pageable arrays)
int binned_product_to_host(const double* d_M, int64_t n, const int64_t* wt_colptr, const int64_t* wt_rowval,
                           const double* wt_nzval, int64_t LNN1, const int64_t* v_colptr, const int64_t* v_rowval,
                           const double* v_nzval, int64_t LNN2, double* N_out, float* ms,
                           int (*d2h)(void* dst, const void* src, size_t bytes) = nullptr);

int binned_product_device(const double* d_M, int64_t n, const int64_t* wt_colptr, const int64_t* wt_rowval,
                          const double* wt_nzval, int64_t LNN1, const int64_t* v_colptr, const int64_t* v_rowval,
                          const double* v_nzval, int64_t LNN2, double* d_N, int64_t ldN, float* ms);

// ---- the same product restricted to output columns [J0, J1) from a column slab of M (multi-device runs) ----
struct BinTables {            // host: w̃ as CSR over its rows I, v as CSC over its columns J (0-based), built once per call
    std::vector<int> wptr, wcol, vptr, vrow;
    std::vector<double> wval, vval;
    bool has_w = false, has_v = false;
    int64_t n = 0, LNN1 = 0, LNN2 = 0;
};
struct BinDev {               // the tables on one device
    DevBuf<int> wptr, wcol, vptr, vrow;
    DevBuf<double> wval, vval;
};
int bin_tables_build(BinTables* t, int64_t n, const int64_t* wt_colptr, const int64_t* wt_rowval, const double* wt_nzval,
                     int64_t LNN1, const int64_t* v_colptr, const int64_t* v_rowval, const double* v_nzval, int64_t LNN2);
int bin_tables_upload(const BinTables& t, BinDev& d, cudaStream_t st);
// M columns [*c0, *c1) that the output columns [J0, J1) read
void bin_needed_cols(const BinTables& t, int64_t J0, int64_t J1, int64_t* c0, int64_t* c1);
// d_N (LNN1 x (J1-J0), leading dimension ldN) = w̃ · M[:, c0..] · v[c0.., J0:J1]; d_Mslab holds the columns of M from c0 on
int binned_product_range(const double* d_Mslab, int64_t c0, const BinTables& t, const BinDev& d, int64_t J0, int64_t J1,
                         double* d_N, int64_t ldN, cudaStream_t st);

// Coupling matrix of a separable window W(r, n̂) = phi(r) mask(n̂): d_wlm planar alm of the mask (1 shell,
// padded to nrp_s), phi host vector of length nr.  Writes the full nout x nout matrix to d_M.
int separable_cmix(CmixPlan* p, const double* d_wlm, int nrp_s, const double* phi, int div2Lp1, int interchange,
                   double* d_M);

}  // namespace sfb
