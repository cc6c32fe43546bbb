/*
 * sfb_b200.h — C ABI of the B200-native (sm_100a) window coupling-matrix library.
 *
 * Drop-in boundary for the `power_win_mix` / `calc_Wr_lm` hot path of
 * hsgg/SphericalFourierBesselDecompositions.jl (v0.5.19).  The reference has no FFI layer of its own
 * (100 % Julia); these entry points are what thin Julia method bodies `ccall` (see INTEGRATION.md and
 * julia/SFBB200.jl).  Citations are file:line in the reference checkout.
 *
 * Conventions
 *   - every function returns 0 on success; nonzero => sfb_last_error() holds the message
 *     (the Julia shim rethrows it as ErrorException, mirroring error()/@assert at
 *     src/windows.jl:803,1013 and src/healpix_helpers.jl:60-63)
 *   - all arrays are column-major (Julia), Int64 indices, Float64 / ComplexF64 (interleaved re,im)
 *   - the caller owns every buffer; nothing is retained past return (wrap calls in GC.@preserve)
 *   - functions without `_dev` take HOST pointers and do their own H2D/D2H;
 *     `_dev` functions take DEVICE pointers on the current device and a cudaStream_t passed as void*
 *   - the library never falls back to a CPU path: without a CUDA device every compute call fails
 */
#ifndef SFB_B200_H
#define SFB_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SFB_B200_VERSION 100

/* Wr_lm column orders (src/LMcalcStructs.jl:7-26) */
#define SFB_LAYOUT_MMAJOR 0 /* HEALPix order, what calc_Wr_lm returns        (LMcalcStruct)      */
#define SFB_LAYOUT_MFAST 1  /* what optimize_Wr_lm_layout returns             (LMcalcStructMfast) */

typedef struct sfb_sht_plan sfb_sht_plan;   /* stage 1: ring tables, Legendre tables, workspaces */
typedef struct sfb_cmix_plan sfb_cmix_plan; /* stage 2+3: mode tables, radial basis, 3j table     */

/* ---- housekeeping ------------------------------------------------------------------------------ */
int32_t sfb_version(void);
const char* sfb_last_error(void);
int32_t sfb_device_count(int32_t* count);
int32_t sfb_set_device(int32_t device);
/* Multi-GPU behind the host-pointer boundary.  After sfb_set_devices(n), n in 1..8, the host-pointer
 * entry points sfb_power_win_mix and sfb_calc_wr_lm run on devices 0..n-1 of THIS process (one worker thread per device
 * for the duration of the call, peer access over NVLink required): every GPU uploads 1/n of the window over its own
 * PCIe link, transforms nr/n shells, forms a contiguous full-height column range of M and copies it straight into the
 * caller's M_out.  This replaces the reference's parallel gather, the pmap over output rows inside _power_win_mix
 * (src/windows.jl:834-861), and the serial shell loop of calc_Wr_lm (src/windows.jl:531-535).  n = 1 restores the
 * single-device path on the current device.  The other entry points keep using the current device.                 */
int32_t sfb_set_devices(int32_t n);
int32_t sfb_get_devices(int32_t* n);
/* Page-locked host memory for callers that want full PCIe speed (a pageable Julia Matrix is staged by the driver at a
 * fraction of it): the Julia shim allocates the result with sfb_host_alloc and wraps it (unsafe_wrap + finalizer), or
 * pins an existing array for the lifetime of several calls with sfb_host_register / sfb_host_unregister.          */
int32_t sfb_host_alloc(void** ptr, int64_t bytes);
int32_t sfb_host_free(void* ptr);
int32_t sfb_host_register(void* ptr, int64_t bytes);
int32_t sfb_host_unregister(void* ptr);
/* times (ms, CUDA events) of the last call on this thread's plans:
 *   [0] stage 1 total, [1] W_{L1} build, [2] mirror fill, [3] Ŵ_{ℓL} build, [4] block kernel + exposed fill,
 *   [5] executed DMMA flops of [4], [6] kernel launches of the last power_win_mix, [7] binned products,
 *   [8] block-kernel launches alone ([4] = [8] + the mirror-fill time not hidden under block kernels; [2] = the mirror
 *       fill launches, which run on a second stream under the block kernels of the next l-chunk) */
int32_t sfb_get_timings(double* out, int32_t n);

/* FP64 tensor-pipe (DMMA) throughput of the current device in TFLOP/s, measured by an in-register probe:
 * the roofline denominator for the DMMA kernels (bench.py) */
int32_t sfb_probe_dmma_tflops(double* tflops);

/* ---- host-pointer entry points (what the Julia methods ccall) ------------------------------- */

/* calc_Wr_lm(win, LMAX, Wnside)                                       src/windows.jl:528-537
 *   win  : nr x npix_in Float64, leading dimension ld_win (>= nr); shell i is the strided row win[i,:]
 *   per shell: udgrade to nside_out (src/healpix_helpers.jl:40-45), then
 *   Healpix.map2alm(lmax=mmax=lmax, niter, uniform weights 4π/npix)   src/healpix_helpers.jl:59-71
 *   out  : nr x lmsize ComplexF64, lmsize=(lmax+1)(lmax+2)/2, columns in `layout` order
 *   errors: lmax > 4*nside_out ("lmax > 4*nside is a poor choice", src/healpix_helpers.jl:60-63)   */
int32_t sfb_calc_wr_lm(const double* win, int64_t nr, int64_t npix_in, int64_t ld_win, int64_t nside_out,
                       int64_t lmax, int64_t niter, int32_t layout, double* out);

/* calc_Wr_lm(win::SeparableArray, LMAX, Wnside)                       src/windows.jl:540-545
 *   one map2alm of the angular mask; out: lmsize ComplexF64 in m-major order (an `Alm`)            */
int32_t sfb_calc_wlm_mask(const double* mask, int64_t npix_in, int64_t nside_out, int64_t lmax, int64_t niter,
                          double* out);

/* calc_Wrl_Wrl + calc_cmix given W_lm(r)                              src/windows.jl:682-746
 *   w1r_lm, w2r_lm : nr x lmsize(LMAX) ComplexF64 in `layout` order; may alias (auto-correlation)
 *   G              : rsdrgnlr, nr x nmax x (lmax+1) Float64 (NaN where n > nmax_l[l+1])  src/windows.jl:799
 *   lnn            : 3 x lnnsize Int64 (ClnnModes.lnn, src/modes.jl:280-288), consumed as given
 *   M_out          : (lnnsize-lnn_min+1)^2 Float64, [i,i'] = (l,n,n') row, (L,N,N') column         */
int32_t sfb_power_win_mix_from_wrlm(const double* w1r_lm, const double* w2r_lm, int64_t nr, int64_t LMAX,
                                    int32_t layout, const double* G, int64_t nmax, int64_t lmax,
                                    const int64_t* lnn, int64_t lnnsize, int64_t lnn_min, int32_t div2Lp1,
                                    int32_t interchange_NN, double* M_out);

/* power_win_mix(win1, win2, wmodes, cmodes; div2Lp1, interchange_NN′, lnn_min)   src/windows.jl:781-805
 *   win2 == NULL or win2 == win1 means win2 === win1.  nside = amodes.nside.                       */
int32_t sfb_power_win_mix(const double* win1, const double* win2, int64_t nr, int64_t npix_in, int64_t ld_win,
                          int64_t nside, const double* G, int64_t nmax, int64_t lmax, const int64_t* lnn,
                          int64_t lnnsize, int64_t lnn_min, int32_t div2Lp1, int32_t interchange_NN,
                          double* M_out);

/* power_win_mix(win1, win2, w̃mat, vmat, wmodes, bcmodes; ...)         src/windows.jl:994-1015, 825-862
 *   w̃ (LNN1 x lnnsize) and v (lnnsize x LNN2) as SparseMatrixCSC triplets (1-based colptr/rowval);
 *   a NULL colptr stands for UniformScaling `I` (then LNN = lnnsize, src/windows.jl:829-830).
 *   Like the reference (src/windows.jl:1005-1006) both transforms are taken from win1.
 *   N_out : LNN1 x LNN2 Float64                                                                   */
int32_t sfb_power_win_mix_binned(const double* win1, int64_t nr, int64_t npix_in, int64_t ld_win, int64_t nside,
                                 const double* G, int64_t nmax, int64_t lmax, const int64_t* lnn,
                                 int64_t lnnsize, const int64_t* wt_colptr, const int64_t* wt_rowval,
                                 const double* wt_nzval, int64_t LNN1, const int64_t* v_colptr,
                                 const int64_t* v_rowval, const double* v_nzval, int64_t LNN2, int32_t div2Lp1,
                                 int32_t interchange_NN, double* N_out);

/* On-device window deconvolution.  The reference finishes on the host with
 *   C = bcmix \ (w̃mat * Cobs)   (docs/src/tutorial_catalog.md:93-97)   and   wmat = inv(Nmix) * w̃M   (test/test_windows.jl:583-584);
 * Julia's `\` on a square matrix is an LU factorisation with partial pivoting.  Here the same factorisation runs on the
 * device (blocked LU, trailing updates on DMMA), so only LNN x nrhs numbers cross PCIe instead of the matrix.
 *   sfb_solve: X = N \ B for host arrays N (n x n) and B (n x nrhs), column-major.
 *   sfb_power_win_mix_binned_solve: N = w̃ M v exactly as sfb_power_win_mix_binned, kept on the device, then X = N \ B
 *     (B: LNN x nrhs); N_out (LNN x LNN) is filled too unless NULL.  A singular matrix is an error (SingularException). */
int32_t sfb_solve(const double* N, int64_t n, const double* B, int64_t nrhs, double* X_out);
int32_t sfb_power_win_mix_binned_solve(const double* win1, int64_t nr, int64_t npix_in, int64_t ld_win, int64_t nside,
                                       const double* G, int64_t nmax, int64_t lmax, const int64_t* lnn, int64_t lnnsize,
                                       const int64_t* wt_colptr, const int64_t* wt_rowval, const double* wt_nzval,
                                       int64_t LNN1, const int64_t* v_colptr, const int64_t* v_rowval,
                                       const double* v_nzval, int64_t LNN2, int32_t div2Lp1, int32_t interchange_NN,
                                       const double* B, int64_t nrhs, double* X_out, double* N_out);

/* win_lnn(win, wmodes, cmodes) (src/windows.jl:382-391, calc_intr_gg_fn :394-418): the shot-noise
 * window W_lnn' = sum_r r^2 dr g_nl(r) g_n'l(r) Wr_00(r)/sqrt(4 pi), with Wr_00 = calc_Wr_lm(win, 2 lmax, nside)[:,1]
 * ("need to be consistent", :386).  G = rsdrgnlr as for power_win_mix; Wlnn_out has lnnsize entries.  A negative
 * Wr_00 is an error, as the reference takes its square root (:400).                                              */
int32_t sfb_win_lnn(const double* win, int64_t nr, int64_t npix_in, int64_t ld_win, int64_t nside, const double* G,
                    int64_t nmax, int64_t lmax, const int64_t* lnn, int64_t lnnsize, double* Wlnn_out);
/* calc_wmix(win, wmodes, amodes; neg_m) (src/windows.jl:299-364, calc_wmix_ii :273-294, calc_gaunts_L :449-464;
 * SURVEY §8f row 1): the full window mixing matrix W_{nlm}^{n'l'm'} for m, m' >= 0 (neg_m != 0: m -> -m), nlmsize x
 * nlmsize ComplexF64 in getidx(amodes, n, l, m) order (src/modes.jl:222-232).  Stage 1 as for power_win_mix
 * (calc_Wr_lm(win, 2 lmax, nside)), overlap integrals on DMMA, general-m 3j families by a warp-cooperative
 * Schulten-Gordon recursion (WignerFamilies semantics, src/windows.jl:434-446).  nmax_l (lmax+1 entries) and lmax_n
 * (nmax entries) are the AnlmModes tables; calc_wmix_all (src/window_chains.jl:578-582) = the two calls neg_m = 0, 1. */
int32_t sfb_calc_wmix(const double* win, int64_t nr, int64_t npix_in, int64_t ld_win, int64_t nside, const double* G,
                      int64_t nmax, int64_t lmax, const int64_t* nmax_l, const int64_t* lmax_n, int32_t neg_m,
                      double* wmix_out);
/* ---- SFB transforms next to the window path; they reuse the batched HEALPix kernels of stage 1.
 * Radial tables are nr x amodes.nmax x (amodes.lmax+1) Float64 (NaN where n > nmax_l[l+1], never read); nmax_l / lmax_n are
 * the AnlmModes tables; f_nlm vectors are nlmsize ComplexF64 in getidx(amodes, n, l, m) order (src/modes.jl:222-232).
 *
 * field2anlm(f_xyz, wmodes, amodes) (src/cat2anlm.jl:326-362): per shell map2alm!(map, alm) with lmax = amodes.lmax,
 *   niter = 3, then f_nlm = Σ_r T[r,n,l] a_lm(r) with T = g_nl(r) r² Δr.  f_xyz: nr x npix (leading dimension ld).     */
int32_t sfb_field2anlm(const double* f_xyz, int64_t nr, int64_t npix, int64_t ld, const double* T, int64_t nmax,
                       int64_t lmax, const int64_t* nmax_l, const int64_t* lmax_n, double* f_nlm_out);
/* anlm2field(f_nlm, wmodes, amodes) (src/cat2anlm.jl:385-422): a_lm(r) = Σ_n g[r,n,l] f_nlm, alm2map! per shell at
 *   amodes.nside; f_xyz_out: nr x 12 nside² (leading dimension ld_out).  g = g_nl(r).                                  */
int32_t sfb_anlm2field(const double* f_nlm, const double* g, int64_t nr, int64_t nside, int64_t nmax, int64_t lmax,
                       const int64_t* nmax_l, const int64_t* lmax_n, double* f_xyz_out, int64_t ld_out);
/* win_rhat_ln(win, wmodes, amodes) (src/windows.jl:244-256): out[p, l+1, n] = Σ_r win[r,p] T[r,n,l], T = r² g_nl(r) Δr;
 *   out: npix x (lmax+1) x nmax, NaN where l > lmax_n[n] (like the reference's NaN-filled array).                      */
int32_t sfb_win_rhat_ln(const double* win, int64_t nr, int64_t npix, int64_t ld_win, const double* T, int64_t nmax,
                        int64_t lmax, const int64_t* nmax_l, const int64_t* lmax_n, double* out);
/* cat2amln(rθϕ, amodes, nbar, win_rhat_ln, weights) (src/cat2anlm.jl:257-314), one batch of nb (n,l) modes: for mode b
 *   map = Σ_{gal} gw[gal,b] δ_{pix(gal)} / (nbar ΔΩ_pix) − win_rhat_ln[:, l_b+1, n_b], mymap2alm!(map, alm) (lmax = amodes.lmax,
 *   niter = 3), anlm[(n_b, l_b, m)] = a_{l_b m}.  The shim evaluates gw[gal,b] = weight g_{n_b l_b}(r_gal) (splines stay in
 *   Julia) and groups the galaxies by RING pixel: pixptr (npix+1, 0-based CSR) and gidx (galaxy indices, catalogue order
 *   within a pixel, as transform_gnl_spmap! sums them :92-99).  mode_n is 1-based.  anlm_inout: the full nlmsize vector;
 *   only the entries of the batch are written.                                                                         */
int32_t sfb_cat2amln(const int64_t* pixptr, const int64_t* gidx, int64_t ngal, const double* gw, const int64_t* mode_n,
                     const int64_t* mode_l, int64_t nb, double nbar, const double* win_rhat_ln, int64_t nside,
                     int64_t nmax, int64_t lmax, const int64_t* nmax_l, const int64_t* lmax_n, double* anlm_inout);
/* device-resident form: d_alm = planar W_lm(r) of sfb_calc_wr_lm_dev, d_Wlnn = nout doubles */
int32_t sfb_win_lnn_dev(sfb_cmix_plan* plan, const double* d_alm, double* d_Wlnn, void* stream);

/* separable window: power_win_mix(win1::SeparableArray, ..., w̃, v, ...)  src/windows.jl:809-814, 942-990
 *   (calc_angular_mixing_matrix :866-878, calc_radial_mixing :924-938, calc_cmixii_separable :651-679)
 *   phi : nr Float64, mask : npix_in Float64; binning arguments as above (NULL colptr = I)         */
int32_t sfb_power_win_mix_separable(const double* phi, const double* mask, int64_t nr, int64_t npix_in,
                                    int64_t nside, const double* G, int64_t nmax, int64_t lmax,
                                    const int64_t* lnn, int64_t lnnsize, const int64_t* wt_colptr,
                                    const int64_t* wt_rowval, const double* wt_nzval, int64_t LNN1,
                                    const int64_t* v_colptr, const int64_t* v_rowval, const double* v_nzval,
                                    int64_t LNN2, int32_t div2Lp1, int32_t interchange_NN, double* N_out);

/* ---- device-resident entry points (plans; used for row-sharded multi-GPU runs and the bench) --- */

/* stage-1 plan for maps of nside_in transformed at nside_out with lmax, batches of up to nr shells */
int32_t sfb_sht_plan_create(sfb_sht_plan** plan, int64_t nside_in, int64_t nside_out, int64_t lmax, int64_t nr);
int32_t sfb_sht_plan_destroy(sfb_sht_plan* plan);
/* number of doubles of the planar W_lm(r) buffer: lmsize * 2 * nrp, nrp = nr rounded up to 8 */
int64_t sfb_sht_alm_doubles(const sfb_sht_plan* plan);
/* d_win: device, [pixel][shell] with pixel stride ld_win (the Julia nr x npix array as is);
 * d_alm: device planar W_lm(r) [lm (m-major)][re,im][nrp]                                          */
int32_t sfb_calc_wr_lm_dev(sfb_sht_plan* plan, const double* d_win, int64_t ld_win, int64_t niter, double* d_alm,
                           void* stream);
/* planar device W_lm(r) -> nr x lmsize ComplexF64 (device) in `layout` order */
int32_t sfb_alm_to_complex_dev(const sfb_sht_plan* plan, const double* d_alm, int32_t layout, double* d_out,
                               void* stream);

/* planar W_lm(r) of all nr shells (nr rounded up to 8 per row) <- shell shards: shard g holds the shells
 * [shell_bounds[g], shell_bounds[g+1]) as [lm][re,im][strides[g]] at shard_ptrs[g] (device pointers on the current device,
 * e.g. slices of an all-gather receive buffer): the placement step of the shell-sharded stage 1 in one kernel          */
int32_t sfb_alm_gather_shards_dev(const double* const* shard_ptrs, const int64_t* shell_bounds, const int64_t* strides,
                                  int32_t nshards, int64_t LMAX, int64_t nr, double* d_alm, void* stream);

int32_t sfb_cmix_plan_create(sfb_cmix_plan** plan, const int64_t* lnn, int64_t lnnsize, int64_t lnn_min,
                             const double* G, int64_t nr, int64_t nmax, int64_t lmax);
int32_t sfb_cmix_plan_destroy(sfb_cmix_plan* plan);
/* rows [row_lo,row_hi) (0-based) of M into d_M (column-major, leading dimension ldM >= row_hi-row_lo).
 * d_alm2 == d_alm1 selects the auto-correlation path.                                              */
int32_t sfb_power_win_mix_dev(sfb_cmix_plan* plan, const double* d_alm1, const double* d_alm2, int32_t div2Lp1,
                              int32_t interchange_NN, int64_t row_lo, int64_t row_hi, double* d_M, int64_t ldM,
                              void* stream);
/* device buffers shareable between the per-GPU processes of one node (cudaIpc*); handle64 is 64 bytes */
int32_t sfb_ipc_alloc(void** dptr, int64_t bytes, void* handle64);
int32_t sfb_ipc_open(const void* handle64, void** dptr);
int32_t sfb_ipc_close(void* dptr);
int32_t sfb_ipc_free(void* dptr);
int32_t sfb_memcpy_dev(void* dst, const void* src, int64_t bytes, void* stream);

/* general block: rows [row_lo,row_hi) x columns [col_lo,col_hi), element (row_lo, col_lo) at d_M[0].  A column
 * shard (all rows, a range of (L,N,N')) is a contiguous slab of the column-major matrix.              */
int32_t sfb_power_win_mix_block_dev(sfb_cmix_plan* plan, const double* d_alm1, const double* d_alm2, int32_t div2Lp1,
                                    int32_t interchange_NN, int64_t row_lo, int64_t row_hi, int64_t col_lo,
                                    int64_t col_hi, double* d_M, int64_t ldM, void* stream);
/* cost models used to balance shards: cost[i] for each of the nout rows / columns (host array) */
int32_t sfb_cmix_row_costs(const sfb_cmix_plan* plan, double* cost, int64_t n);
int32_t sfb_cmix_col_costs(const sfb_cmix_plan* plan, double* cost, int64_t n);

/* Multi-GPU exchange format (auto-correlation, lnn sorted by l, nr <= 64): "upper-packed" storage keeps, for output
 * column j, only the rows of the blocks with l <= l(j), i.e. rows [0, rend(j)), contiguously at element offset
 * offsets[j] (offsets[nout] = total length).  A rank forms the blocks with l <= L of its column range
 * [col_lo,col_hi) -- a contiguous slab [offsets[col_lo], offsets[col_hi]) of the packed buffer -- the slabs are
 * all-gathered in place (half the bytes of the full matrix), and sfb_cmix_unpack_mirror_dev expands the packed buffer
 * into the full column-major matrix: M[r,j] = P[offsets[j]+r], and below the block diagonal M[j,r] = M[r,j] f_r/f_j,
 * f = (div2Lp1 ? 1 : 2l+1)(interchange_NN ? 1 : 1+[n != n']): the un-symmetrised kernel of src/windows.jl:613-627 is
 * symmetric under (l,n,n') <-> (L,N,N') (derivations/sfb.tex:516-517; the reference's own to-do, src/windows.jl:22-33). */
int32_t sfb_cmix_packed_offsets(const sfb_cmix_plan* plan, int64_t* offsets, int64_t n_plus_1);
int32_t sfb_cmix_col_costs_upper(const sfb_cmix_plan* plan, double* cost, int64_t n);
int32_t sfb_power_win_mix_upper_packed_dev(sfb_cmix_plan* plan, const double* d_alm, int32_t div2Lp1,
                                           int32_t interchange_NN, int64_t col_lo, int64_t col_hi, double* d_packed,
                                           void* stream);
/* "L-shaped" shards (multi-GPU output without redundant work; the reference has no counterpart, its gather is the pmap at
 * src/windows.jl:841-861): the device that formed the upper-packed columns [col_lo, col_hi) (whole L-blocks) also forms
 * the rows [col_lo, col_hi) of the part of M below the block diagonal, locally, from the symmetry of the un-symmetrised
 * kernel: d_rows[r * ld_rows + (j - col_lo)] = M[j, r] for l(r) < l(j).  Every element (i, i') of M then lives on exactly
 * one device — the one whose L range holds max(l_i, l_i').                                                            */
int32_t sfb_cmix_mirror_rows_dev(sfb_cmix_plan* plan, const double* d_packed, int64_t col_lo, int64_t col_hi,
                                 int32_t div2Lp1, int32_t interchange_NN, double* d_rows, int64_t ld_rows, void* stream);
int32_t sfb_cmix_unpack_mirror_dev(sfb_cmix_plan* plan, const double* d_packed, int32_t div2Lp1, int32_t interchange_NN,
                                   double* d_M, int64_t ldM, void* stream);
/* Fused exchange: packed_of_rank[g] is rank g's packed buffer (sfb_ipc_alloc / sfb_ipc_open, own buffer included), rank g
 * owning the columns [col_bounds[g], col_bounds[g+1]).  The kernel pulls every column from its owner over NVLink while
 * writing the local full matrix, so no separate all-gather of M runs; rank my_rank starts with the columns of rank
 * my_rank+1 and proceeds cyclically, so the readers never gang up on one owner.  Callers must order it after every rank's
 * sfb_power_win_mix_upper_packed_dev (a stream-ordered barrier, e.g. a 1-element NCCL all-reduce).            */
int32_t sfb_cmix_unpack_mirror_peers_dev(sfb_cmix_plan* plan, const double* const* packed_of_rank,
                                         const int64_t* col_bounds, int32_t nranks, int32_t my_rank,
                                         int32_t div2Lp1, int32_t interchange_NN, double* d_M, int64_t ldM,
                                         void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SFB_B200_H */
