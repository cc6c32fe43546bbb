// Stage 2+3 of the window coupling matrix on sm_100a (FP64, DMMA).
//
// Replaces (reference hsgg/SphericalFourierBesselDecompositions.jl):
//   calc_Wrl_Wrl        src/windows.jl:682-696   -> wl_build_kernel
//   wigner3j000         src/windows.jl:421-431   -> w3j000sq_table_kernel (warp-cooperative recursion)
//   calc_cmixlnnLNN!    src/windows.jl:613-627   -> what_build_kernel + cmix_block_kernel
//   calc_cmix           src/windows.jl:700-746   -> cmix_block_kernel epilogue (N<->N' partner, flags)
//   _power_win_mix      src/windows.jl:825-862   -> csr/csc products in binned.cu
//   win_lnn             src/windows.jl:382-418   -> win_lnn_kernel
#pragma once
#include "common.cuh"
#include <utility>
#include <vector>

namespace sfb {

struct CmixPlan {
    // problem sizes
    int lmax = 0, nmax = 0, LMAX = 0;
    int nr = 0, nrp = 0, S = 0;  // shells, padded to 8, smem row stride (≡ 4 mod 16)
    int64_t lnnsize = 0;         // size of the caller's lnn table
    int64_t nout = 0;            // lnnsize - lnn_min + 1 : rows == cols of M
    int64_t lnn_min = 1;
    int nl = 0;                  // number of (L,N) pairs with a column
    int amax_tiles = 0;

    // host tables.  Row blocks: [0, lmax] are the l-blocks of the table; an l with nmax_l > 32 (more than four 8-row
    // tiles) is run as "virtual" row blocks appended behind them, one per pair (p <= q) of 16-wide panels of its n range,
    // whose basis is the concatenation of the two panels (<= 32 functions) and whose rows are the (n in p, n' in q)
    // modes: the block kernels never see more than 32 row-side functions.  a_of_ell / ell_ptr / G cover all blocks.
    int nblk = 0;
    std::vector<int> blk_real;                 // the l of every row block
    std::vector<std::vector<int>> blk_of_ell;  // row blocks the tiled kernels launch for l ({l} unless virtualised)
    std::vector<int> a_of_ell;        // max n (1-based) used by rows/cols of this ell (or virtual block), 0 if none
    std::vector<int> ell_ptr;         // CSR over row blocks -> rows
    std::vector<int> h_row_out, h_row_n, h_row_n2;

    // device tables
    DevBuf<double> d_G;               // [ell][n][nrp]
    DevBuf<int> d_ell_ptr, d_row_out, d_row_n, d_row_n2, d_a, d_blk_real;
    DevBuf<int> d_pairidx;            // [L][N][N'] -> output column or -1
    DevBuf<int> d_nl_L, d_nl_N;       // grid.x -> (L, N)
    DevBuf<double> d_w2;              // [(lmax+1)^2][lmax+1] squared 3j symbols
    DevBuf<double> d_W;               // [LMAX+1][nrp][nrp]
    DevBuf<double> d_What;            // [ell chunk][L][nrp][nrp]
    DevBuf<int> d_ell_list;           // per-launch list of ells
    DevBuf<int> d_chunks;             // per-launch (L, N0, N1) chunk lists
    DevBuf<int> d_what_ells;          // l-blocks whose Ŵ is built for the current row shard
    DevBuf<int> d_slot_of_out;        // output index -> CSR slot
    DevBuf<int> d_es;                 // per output index: l | [n≠n'] << 30 (mirror fill)
    bool ell_sorted = false;          // l non-decreasing in output order
    std::vector<long long> h_colbase; // upper-packed storage: element offset of column j (nout + 1 entries), ell_sorted only
    DevBuf<long long> d_colbase;
    DevBuf<int> d_jt_order;           // column-tile visiting order of the fused pull (cmix_unpack_mirror)
    std::vector<int> h_jt_key;        // (column bounds, nranks, rank) the order was built for
    DevBuf<int> d_regz_blocks;        // per-launch block descriptors of the register-Z kernel (cmix_regz.cu)
    DevBuf<int> d_regz_queue;         // block-queue heads of the persistent register-Z launches (one int per launch)
    size_t what_budget_bytes = size_t(2) << 30;

    // last-run stage times (ms): wl, w3j, what, block
    float t_wl = 0, t_fill = 0, t_what = 0, t_block = 0;  // t_block = t_k3 + the fill time not hidden under block kernels
    float t_k3 = 0;                                        // block-kernel launches alone
    cudaStream_t side_stream = nullptr;                    // mirror fills of finished l-chunks run here
    std::vector<cudaEvent_t> pend_fill_ev;                 // (begin, end) per fill launch of the side stream
    double flops_executed = 0;
    int launches = 0;
    // async_times: cmix_run records its timing events and returns without a host sync; cmix_resolve_times fills t_*.
    // pend_ev = [wl begin, wl end, (chunk begin, Ŵ end, blocks end)*, fill begin, fill end]
    bool async_times = false, pending = false, pend_fill = false;
    std::vector<cudaEvent_t> pend_ev;
};

// Build the plan from the caller's tables (host pointers).
//   lnn: 3 x lnnsize Int64, column-major (Julia Matrix{Int}), 1-based n
//   G:   nr x nmax x (lmax+1) Float64 column-major (rsdrgnlr, NaN where n > nmax_l)
int cmix_plan_create(CmixPlan** out, const int64_t* lnn, int64_t lnnsize, int64_t lnn_min, const double* G,
                     int64_t nr, int64_t nmax, int64_t lmax);
void cmix_plan_destroy(CmixPlan* p);
int cmix_resolve_times(CmixPlan* p);

// alm layout (device): planar [lm (m-major, lmax2 = 2*lmax)][comp (re,im)][nrp], padded shells zero.
// Writes the block rows [row_lo,row_hi) x columns [col_lo,col_hi) (0-based, of the nout x nout matrix) into d_M
// (column-major, leading dim ldM, element (row_lo, col_lo) at offset 0).
// mirror = true (auto-correlation only): d_M is the base of the FULL matrix; for the l-blocks of the row range only
// the blocks with L >= l are formed and every tile also fills the mirrored block, using
//   M[(L,N,N'),(l,n,n')] = c_l (A + [n≠n'] B),  M[(l,n,n'),(L,N,N')] = c_L (A + [N≠N'] B)   (same A, B).
// upper_packed = true (auto-correlation, register-Z kernel, l-sorted table): only the blocks with l <= L of the column
// range [col_lo, col_hi) are formed and written in upper-packed storage (column j at d_M + colbase[j], rows
// [0, rend(j))); d_M is the base of the whole packed buffer.  cmix_unpack_mirror expands it.
// `peers` (optional): up to 7 more device pointers (peer-mapped, same offset/ldM semantics as d_M); every element
// is also stored there, which fuses the all-gather of row shards into the kernel epilogue.
int cmix_run(CmixPlan* p, const double* d_alm1, const double* d_alm2, int div2Lp1, int interchange,
             int64_t row_lo, int64_t row_hi, int64_t col_lo, int64_t col_hi, double* d_M, int64_t ldM,
             cudaStream_t stream, double* const* peers = nullptr, int npeers = 0, bool reuse_wl = false,
             bool mirror = false, bool upper_packed = false);
// Register-resident-Z block kernel (cmix_regz.cu): auto-correlation, nr <= 64, no peer stores.
bool cmix_regz_eligible(const CmixPlan* p, bool sym, int npeers);
int cmix_regz_run(CmixPlan* p, const std::vector<int>& blocks, const double* d_What, int div2Lp1, int interchange,
                  int64_t col_lo, int64_t col_hi, double* d_M, int64_t ldM, const long long* d_colbase,
                  cudaStream_t stream, double* flops, int* launches);
// upper-packed storage -> full column-major matrix (direct copy of the l <= L blocks + mirror image below them)
// bases[g] = packed buffer of rank g (peer-mapped; nranks = 1: the local buffer), rank g owning the columns
// [col_bounds[g], col_bounds[g+1]) (col_bounds may be null for nranks = 1)
int cmix_unpack_mirror(CmixPlan* p, const double* const* bases, const int64_t* col_bounds, int nranks, int my_rank,
                       int div2Lp1, int interchange, double* d_M, int64_t ldM, cudaStream_t stream);
int cmix_mirror_fill(CmixPlan* p, int64_t c0, int64_t c1, int div2Lp1, int interchange, double* d_M, int64_t ldM,
                     cudaStream_t stream);
// rows [j_lo, j_hi) of the part of M below the block diagonal, from the upper-packed columns [j_lo, j_hi) of the same
// device: d_rows[r * ldR + (j - j_lo)] = M[j, r] for l(r) < l(j)  ("L-shaped" multi-GPU shards)
int cmix_mirror_rows(CmixPlan* p, const double* d_packed, int64_t j_lo, int64_t j_hi, int div2Lp1, int interchange,
                     double* d_rows, int64_t ldR, cudaStream_t stream);
// l-block aligned row ranges of roughly equal cost for the mirrored, pipelined host path
std::vector<std::pair<int64_t, int64_t>> cmix_row_chunks_mirror(const CmixPlan* p, int k);
// column ranges of roughly equal cost (L-block aligned) for pipelining compute with the D2H of finished slabs
std::vector<std::pair<int64_t, int64_t>> cmix_col_chunks(const CmixPlan* p, int k);
std::vector<std::pair<int64_t, int64_t>> cmix_col_chunks_range(const CmixPlan* p, int64_t lo, int64_t hi, int k);
// column bounds[0..ndev] for a multi-device run, from the caller's mode table alone (l-block aligned, cost balanced)
int cmix_col_bounds_from_lnn(const int64_t* lnn, int64_t lnnsize, int64_t lnn_min, int64_t nr, int64_t nmax, int64_t lmax,
                             int ndev, int64_t* bounds);

// win_lnn (src/windows.jl:382-418) from the planar alm of stage 1: out[i] = Σ_r G_ln G_ln' W_00(r)/√(4π), nout values
int win_lnn_run(CmixPlan* p, const double* d_alm, double* d_out, cudaStream_t stream);

// Host complex (nr x lmsize, column-major, interleaved) -> device planar alm; layout 0 = m-major, 1 = m-fast.
int alm_from_host(const double* h_wrlm, int64_t nr, int lmax2, int layout, DevBuf<double>& d_alm, int nrp,
                  cudaStream_t stream);

}  // namespace sfb
