// Stage 1: batched HEALPix map2alm (niter Jacobi refinements) for all radial shells at once, sm_100a FP64.
//
// Replaces (reference hsgg/SphericalFourierBesselDecompositions.jl):
//   calc_Wr_lm                      src/windows.jl:528-545
//   udgrade(::Vector, nside)        src/healpix_helpers.jl:40-45   (Healpix.jl udgrade)
//   mymap2alm(map; lmax)            src/healpix_helpers.jl:59-71   (Healpix.jl map2alm -> libsharp2)
//   optimize_Wr_lm_layout           src/windows.jl:589-605         (folded into the output conversion)
#pragma once
#include "common.cuh"
#include <vector>

namespace sfb {

struct ShtPlan {
    int nside_in = 0, nside = 0, lmax = 0, nr = 0, nrp = 0;
    int64_t npix_in = 0, npix = 0;
    int nrings = 0, nhalf = 0;
    size_t lmsize = 0;
    int ntiles = 0;  // synthesis (ring, pixel-chunk) tiles

    DevBuf<int> d_nphi, d_start, d_shift, d_twoff;
    DevBuf<int> d_tile_ring, d_tile_j0;
    DevBuf<int> d_gemm_rings, d_fft_rings;  // rings transformed by DFT-as-GEMM (polar caps) / by the smem FFT (belt)
    DevBuf<int> d_cap_rings, d_ctile_ring, d_ctile_q0;  // polar-cap rings: four-fold folded DFT-as-GEMM
    int n_cap_rings = 0, n_ctiles = 0;
    bool use_fft = false;
    int n_gemm_rings = 0, n_fft_rings = 0, log2n = 0, fft_sch = 1;
    DevBuf<double2> d_tw;    // (cos, sin)(π t / nφ), t in [0, 2nφ), one block per distinct ring length
    DevBuf<double> d_lam;    // λ_lm(θ_k): [lm (m-major)][nhalf]
    DevBuf<double> d_FG;     // ring-space intermediates [m][ring][2*nrp]
    DevBuf<double> d_map;    // udgraded maps [npix][nrp] (only if nside_in != nside)
    DevBuf<double> d_resid;  // residual maps [npix][nrp] (pixel-space refinement only)
    DevBuf<double> d_F2;     // aliased ring-Fourier coefficients of the synthesised field (ring-space refinement)
    DevBuf<double> d_a0;     // A f, the first-pass alm
    DevBuf<int> d_alias_rings;   // rings with nφ <= 2 lmax (both hemispheres): the alias pass visits only these
    int n_alias_rings = 0;
    // m-cutoff (SFB_SHT_NO_MLIM disables): beyond mlim_ring[k] every λ_lm(θ_k) is < 1e-30
    DevBuf<int> d_mlim_ring;     // [nhalf] north rings, monotone towards the equator
    DevBuf<int> d_mlim_tile;     // [ceil(nhalf/64)] max over a synthesis tile of 64 rings
    DevBuf<int> d_kbeg_of_m;     // [lmax+1] first ring chunk (multiple of 32) the Legendre analysis of m has to visit
    int kpolar = 0;          // north rings [0, kpolar) keep synthesis -> alias -> analysis in a Jacobi pass ...
    DevBuf<double> d_gram;   // ... the alias-free rings beyond act through Gram matrices K_m (see sht.cu)
    DevBuf<long long> d_gram_off;
    DevBuf<double> d_t;      // A f - K alm

    float t_total = 0;
    int launches = 0;
    // async_times: sht_map2alm only records its two timing events and returns without a host sync; t_total is filled by
    // sht_resolve_times (the device-resident API uses this so that the host can enqueue stage 2+3 behind stage 1)
    bool async_times = false, pending = false;
    cudaEvent_t tev0 = nullptr, tev1 = nullptr;
    // independent kernels of one transform run side by side (SFB_SHT_SERIAL=1: one stream): the belt FFT next to the cap
    // DFT, and the Gram-matrix product next to the polar synthesis + alias pass of a Jacobi iteration
    cudaStream_t side = nullptr;
    cudaEvent_t fork_ev = nullptr, join_ev = nullptr;
};

int sht_plan_create(ShtPlan** out, int64_t nside_in, int64_t nside_out, int64_t lmax, int64_t nr);
void sht_plan_destroy(ShtPlan* p);

// d_win: [pixel][shell], pixel stride ldw (>= nr).  d_alm: planar [lm (m-major)][re,im][nrp].
int sht_map2alm(ShtPlan* p, const double* d_win, int64_t ldw, int niter, double* d_alm, cudaStream_t stream);
int sht_resolve_times(ShtPlan* p);
// alm2map of every shell: planar alm -> d_out[pixel][nrp] (RING order, nside of the plan)
int sht_alm2map(ShtPlan* p, const double* d_alm, double* d_out, cudaStream_t stream);
// planar -> ComplexF64 nr x lmsize (device), layout 0 = m-major, 1 = m-fast
int sht_alm_to_complex(const ShtPlan* p, const double* d_alm, int layout, double* d_out, cudaStream_t stream);
// the same conversion without a plan (multi-device runs assemble W_lm(r) from shards of several plans)
int alm_planar_to_complex(const double* d_alm, int lmax, int nr, int nrp, int layout, double* d_out, cudaStream_t stream);

}  // namespace sfb
