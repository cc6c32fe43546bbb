// extern "C" entry points declared in include/sfb_b200.h.
#include "../../include/sfb_b200.h"

#include "binned.cuh"
#include "cmix.cuh"
#include "common.cuh"
#include "sht.cuh"

#include <chrono>
#include <cstdlib>
#include <cstring>
#include <mutex>

namespace sfb {
static thread_local std::string g_err;
void set_error(const std::string& msg) { g_err = msg; }

// SFB_TRACE=1: wall-clock of the phases of the host-pointer entry points on stderr
struct Trace {
    bool on = getenv("SFB_TRACE") != nullptr;
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    void mark(const char* what) {
        if (!on) return;
        cudaDeviceSynchronize();
        auto t1 = std::chrono::steady_clock::now();
        fprintf(stderr, "[sfb] %-28s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(t1 - t0).count());
        t0 = t1;
    }
};

static std::mutex g_mutex;  // one call at a time (the Julia side calls from one task and blocks)
static double g_times[8] = {0, 0, 0, 0, 0, 0, 0, 0};
// plans whose last (device-resident, asynchronous) run has unread timing events: resolved by sfb_get_timings
static ShtPlan* g_pend_sht = nullptr;
static CmixPlan* g_pend_cmix = nullptr;

// ---- cached stage-1 plan (tables depend only on nside/lmax/nr) ----
static ShtPlan* g_sht = nullptr;
static int g_sht_dev = -1;
static int get_sht_plan(ShtPlan** out, int64_t nside_in, int64_t nside_out, int64_t lmax, int64_t nr) {
    int dev = 0;
    SFB_CUDA_OK(cudaGetDevice(&dev));
    if (g_sht && g_sht_dev == dev && g_sht->nside_in == nside_in && g_sht->nside == nside_out && g_sht->lmax == lmax &&
        g_sht->nr == nr) {
        *out = g_sht;
        return 0;
    }
    if (g_sht) {
        sht_plan_destroy(g_sht);
        g_sht = nullptr;
    }
    SFB_TRY(sht_plan_create(&g_sht, nside_in, nside_out, lmax, nr));
    g_sht_dev = dev;
    *out = g_sht;
    return 0;
}

static int npix2nside(int64_t npix, int64_t* nside) {
    int64_t ns = (int64_t)llround(std::sqrt((double)npix / 12.0));
    SFB_REQUIRE(ns >= 1 && 12 * ns * ns == npix, "npix is not 12*nside^2");
    *nside = ns;
    return 0;
}

// H2D of the Julia array win (nr x npix, leading dimension ld) -> device [pixel][nr]
static int upload_win(const double* win, int64_t nr, int64_t npix, int64_t ld, DevBuf<double>& d) {
    SFB_REQUIRE(win, "win is null");
    SFB_REQUIRE(ld >= nr, "ld_win < nr");
    SFB_TRY(d.alloc((size_t)nr * npix));
    SFB_CUDA_OK(cudaMemcpy2D(d.p, nr * sizeof(double), win, ld * sizeof(double), nr * sizeof(double), npix,
                             cudaMemcpyHostToDevice));
    return 0;
}

__global__ void finite_check_kernel(const double* __restrict__ x, size_t n, int* flag) {
    int bad = 0;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        bad |= !isfinite(x[i]);
    if (bad) atomicOr(flag, 1);
}

static int check_finite(const double* d, size_t n, const char* what) {
    DevBuf<int> flag;
    SFB_TRY(flag.alloc(1));
    SFB_CUDA_OK(cudaMemset(flag.p, 0, sizeof(int)));
    finite_check_kernel<<<1024, 256>>>(d, n, flag.p);
    int h = 0;
    SFB_CUDA_OK(cudaMemcpy(&h, flag.p, sizeof(int), cudaMemcpyDeviceToHost));
    if (h) {  // @assert all(isfinite.(mix))  src/windows.jl:803,1013
        set_error(std::string("AssertionError: all(isfinite.(") + what + "))");
        return 4;
    }
    return 0;
}

// FP64 tensor-pipe probe: independent DMMA chains, no memory traffic.  Gives the roofline denominator for the
// DMMA kernels (MEASURED_PEAKS.json only carries HBM and bf16 figures).
__global__ void __launch_bounds__(256) dmma_probe_kernel(double* out, int iters, double seed) {
    double acc[8][2];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i][0] = acc[i][1] = 0.0;
    double a = seed + threadIdx.x * 1e-9, b = 1.0 - seed;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) dmma884(acc[i], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += acc[i][0] + acc[i][1];
    if (s == 12345.678) out[0] = s;  // keep the chain alive
}

static void record_cmix_times(const CmixPlan* p) {
    g_times[1] = p->t_wl;
    g_times[2] = p->t_fill;
    g_times[3] = p->t_what;
    g_times[4] = p->t_block;
    g_times[5] = p->flops_executed;
    g_times[6] += p->launches;
}

// Device-resident entry points: run without a host sync and leave the timing events to sfb_get_timings.
static bool async_enabled() { return getenv("SFB_SYNC_TIMINGS") == nullptr; }   // SFB_SYNC_TIMINGS=1: sync per call
#define SFB_CMIX_RUN_ASYNC(plan_, call_)            \
    do {                                            \
        (plan_)->async_times = async_enabled();     \
        const int rc__ = (call_);                   \
        (plan_)->async_times = false;               \
        if (rc__ != 0) return rc__;                 \
        g_pend_cmix = (plan_);                      \
    } while (0)

// Device workspace reused across host-pointer calls (cudaMalloc / cudaFree of multi-GB buffers is slow).
struct Workspace {
    DevBuf<double> win, alm1, alm2, slab[2], M;
    DevBuf<int> flag;
    cudaStream_t copy = nullptr;
    cudaEvent_t computed[2] = {nullptr, nullptr}, copied[2] = {nullptr, nullptr};
    int dev = -1;
    int init() {
        int d = 0;
        SFB_CUDA_OK(cudaGetDevice(&d));
        if (copy && d == dev) return 0;
        SFB_CUDA_OK(cudaStreamCreateWithFlags(&copy, cudaStreamNonBlocking));
        for (int i = 0; i < 2; ++i) {
            SFB_CUDA_OK(cudaEventCreateWithFlags(&computed[i], cudaEventDisableTiming));
            SFB_CUDA_OK(cudaEventCreateWithFlags(&copied[i], cudaEventDisableTiming));
        }
        dev = d;
        return 0;
    }
};
static Workspace g_ws;

// Stage 2+3 for all columns, pipelined with the device->host copy: the matrix is produced in column slabs
// (contiguous in column-major order); while slab k is copied to the caller's buffer, slab k+1 is computed.
static bool mirror_enabled() { return getenv("SFB_NO_MIRROR") == nullptr; }

static int cmix_to_host_pipelined(CmixPlan* p, const double* a1, const double* a2, int div2Lp1, int interchange,
                                  double* M_out) {
    SFB_TRY(g_ws.init());
    const int64_t n = p->nout;
    const bool mirror = (a1 == a2) && mirror_enabled();
    // mirror mode: step k forms the blocks (l in chunk k, L >= l) and their mirror images in the full device
    // matrix, after which the columns of chunk k are complete and can leave; otherwise plain column slabs.
    const auto chunks = mirror ? cmix_row_chunks_mirror(p, 8) : cmix_col_chunks(p, 8);
    int64_t maxc = 0;
    for (auto& c : chunks) maxc = std::max(maxc, c.second - c.first);
    if (mirror) {
        SFB_TRY(g_ws.M.alloc((size_t)n * n));
    } else {
        for (int i = 0; i < 2 && i < (int)chunks.size(); ++i) SFB_TRY(g_ws.slab[i].alloc((size_t)maxc * n));
    }
    SFB_TRY(g_ws.flag.alloc(1));
    SFB_CUDA_OK(cudaMemsetAsync(g_ws.flag.p, 0, sizeof(int), 0));
    float t_wl = 0, t_what = 0, t_block = 0;
    double flops = 0;
    int launches = 0;
    for (size_t k = 0; k < chunks.size(); ++k) {
        const int b = (int)(k & 1);
        const int64_t c0 = chunks[k].first, c1 = chunks[k].second;
        double* src = nullptr;
        if (mirror) {
            SFB_TRY(cmix_run(p, a1, a2, div2Lp1, interchange, c0, c1, 0, n, g_ws.M.p, n, 0, nullptr, 0, k > 0, true));
            src = g_ws.M.p + c0 * n;
        } else {
            if (k >= 2) SFB_CUDA_OK(cudaEventSynchronize(g_ws.copied[b]));  // slab free again
            SFB_TRY(cmix_run(p, a1, a2, div2Lp1, interchange, 0, n, c0, c1, g_ws.slab[b].p, n, 0, nullptr, 0, k > 0));
            src = g_ws.slab[b].p;
        }
        t_wl += p->t_wl;
        t_what += p->t_what;
        t_block += p->t_block;
        flops += p->flops_executed;
        launches += p->launches;
        finite_check_kernel<<<512, 256, 0, 0>>>(src, (size_t)(c1 - c0) * n, g_ws.flag.p);
        SFB_CUDA_OK(cudaEventRecord(g_ws.computed[b], 0));
        SFB_CUDA_OK(cudaStreamWaitEvent(g_ws.copy, g_ws.computed[b], 0));
        SFB_CUDA_OK(cudaMemcpyAsync(M_out + c0 * n, src, (size_t)(c1 - c0) * n * sizeof(double), cudaMemcpyDeviceToHost,
                                    g_ws.copy));
        SFB_CUDA_OK(cudaEventRecord(g_ws.copied[b], g_ws.copy));
    }
    SFB_CUDA_OK(cudaStreamSynchronize(g_ws.copy));
    p->t_wl = t_wl;
    p->t_what = t_what;
    p->t_block = t_block;
    p->flops_executed = flops;
    p->launches = launches + (int)chunks.size();
    int h = 0;
    SFB_CUDA_OK(cudaMemcpy(&h, g_ws.flag.p, sizeof(int), cudaMemcpyDeviceToHost));
    if (h) {  // @assert all(isfinite.(mix))  src/windows.jl:803
        set_error("AssertionError: all(isfinite.(mix))");
        return 4;
    }
    return 0;
}

// Host-pointer entry points reuse the stage-2/3 plan (index tables, radial basis, 3j table, Ŵ workspace) when the
// caller passes the same tables again: keyed by the sizes and an FNV-1a hash of the lnn and G bytes.
static uint64_t fnv1a(const void* data, size_t bytes, uint64_t h = 1469598103934665603ull) {
    const uint64_t* w = static_cast<const uint64_t*>(data);
    for (size_t i = 0; i < bytes / 8; ++i) {
        h ^= w[i];
        h *= 1099511628211ull;
    }
    return h;
}
struct PlanCache {
    CmixPlan* p = nullptr;
    uint64_t key = 0;
    int64_t dims[5] = {0, 0, 0, 0, 0};
    int dev = -1;
};
static PlanCache g_plan_cache;

struct PlanGuard {  // non-owning handle on the cached plan
    CmixPlan* p = nullptr;
};

static int get_cmix_plan(CmixPlan** out, const int64_t* lnn, int64_t lnnsize, int64_t lnn_min, const double* G,
                         int64_t nr, int64_t nmax, int64_t lmax) {
    SFB_REQUIRE(lnn && G && lnnsize >= 1 && nr >= 1 && nmax >= 1 && lmax >= 0, "bad mode tables");
    int dev = 0;
    SFB_CUDA_OK(cudaGetDevice(&dev));
    uint64_t key = fnv1a(lnn, (size_t)lnnsize * 3 * sizeof(int64_t));
    key = fnv1a(G, (size_t)nr * nmax * (lmax + 1) * sizeof(double), key);
    const int64_t dims[5] = {lnnsize, lnn_min, nr, nmax, lmax};
    PlanCache& c = g_plan_cache;
    if (c.p && c.dev == dev && c.key == key && std::memcmp(c.dims, dims, sizeof(dims)) == 0) {
        *out = c.p;
        return 0;
    }
    if (c.p) {
        cmix_plan_destroy(c.p);
        c.p = nullptr;
    }
    SFB_TRY(cmix_plan_create(&c.p, lnn, lnnsize, lnn_min, G, nr, nmax, lmax));
    c.key = key;
    c.dev = dev;
    std::memcpy(c.dims, dims, sizeof(dims));
    *out = c.p;
    return 0;
}

// W_lm(r) of one or two windows on the device (planar), shared by the end-to-end entry points
static int windows_to_alm(const double* win1, const double* win2, int64_t nr, int64_t npix_in, int64_t ld_win,
                          int64_t nside, int64_t LMAX, DevBuf<double>& alm1, DevBuf<double>& alm2, bool* same) {
    int64_t nside_in = 0;
    SFB_TRY(npix2nside(npix_in, &nside_in));
    ShtPlan* sp = nullptr;
    SFB_TRY(get_sht_plan(&sp, nside_in, nside, LMAX, nr));
    const size_t nalm = sp->lmsize * 2 * sp->nrp;
    DevBuf<double>& d_win = g_ws.win;
    SFB_TRY(upload_win(win1, nr, npix_in, ld_win, d_win));
    SFB_TRY(alm1.alloc(nalm));
    SFB_TRY(sht_map2alm(sp, d_win.p, nr, 3, alm1.p, 0));
    g_times[0] = sp->t_total;
    g_times[6] = sp->launches;
    *same = (win2 == nullptr || win2 == win1);
    if (!*same) {
        SFB_TRY(upload_win(win2, nr, npix_in, ld_win, d_win));
        SFB_TRY(alm2.alloc(nalm));
        SFB_TRY(sht_map2alm(sp, d_win.p, nr, 3, alm2.p, 0));
        g_times[0] += sp->t_total;
        g_times[6] += sp->launches;
    }
    return 0;
}
}  // namespace sfb

using namespace sfb;

extern "C" {

int32_t sfb_version(void) { return SFB_B200_VERSION; }
const char* sfb_last_error(void) { return g_err.c_str(); }

int32_t sfb_device_count(int32_t* count) {
    SFB_REQUIRE(count, "count is null");
    int n = 0;
    SFB_CUDA_OK(cudaGetDeviceCount(&n));
    *count = n;
    return 0;
}

int32_t sfb_set_device(int32_t device) {
    std::lock_guard<std::mutex> lk(g_mutex);
    SFB_CUDA_OK(cudaSetDevice(device));
    return 0;
}

int32_t sfb_get_timings(double* out, int32_t n) {
    std::lock_guard<std::mutex> lk(g_mutex);
    SFB_REQUIRE(out && n >= 0, "bad arguments");
    if (g_pend_sht) {
        SFB_TRY(sht_resolve_times(g_pend_sht));
        g_times[0] = g_pend_sht->t_total;
        g_pend_sht = nullptr;
    }
    if (g_pend_cmix) {
        SFB_TRY(cmix_resolve_times(g_pend_cmix));
        const double launches = g_times[6];
        record_cmix_times(g_pend_cmix);
        g_times[6] = launches;
        g_pend_cmix = nullptr;
    }
    for (int i = 0; i < n && i < 8; ++i) out[i] = g_times[i];
    return 0;
}

// ------------------------------------------------------------------------------------------------
int32_t sfb_calc_wr_lm(const double* win, int64_t nr, int64_t npix_in, int64_t ld_win, int64_t nside_out,
                       int64_t lmax, int64_t niter, int32_t layout, double* out) {
    std::lock_guard<std::mutex> lk(g_mutex);
    SFB_REQUIRE(win && out, "null pointer");
    SFB_REQUIRE(nr >= 1, "nr < 1");
    SFB_REQUIRE(layout == 0 || layout == 1, "bad layout");
    int64_t nside_in = 0;
    SFB_TRY(npix2nside(npix_in, &nside_in));
    ShtPlan* sp = nullptr;
    SFB_TRY(get_sht_plan(&sp, nside_in, nside_out, lmax, nr));
    DevBuf<double> d_win, d_alm, d_out;
    SFB_TRY(upload_win(win, nr, npix_in, ld_win, d_win));
    SFB_TRY(d_alm.alloc(sp->lmsize * 2 * sp->nrp));
    SFB_TRY(d_out.alloc(sp->lmsize * 2 * nr));
    SFB_TRY(sht_map2alm(sp, d_win.p, nr, (int)niter, d_alm.p, 0));
    g_times[0] = sp->t_total;
    g_times[6] = sp->launches;
    SFB_TRY(sht_alm_to_complex(sp, d_alm.p, layout, d_out.p, 0));
    SFB_CUDA_OK(cudaMemcpy(out, d_out.p, sp->lmsize * 2 * nr * sizeof(double), cudaMemcpyDeviceToHost));
    return 0;
}

int32_t sfb_calc_wlm_mask(const double* mask, int64_t npix_in, int64_t nside_out, int64_t lmax, int64_t niter,
                          double* out) {
    return sfb_calc_wr_lm(mask, 1, npix_in, 1, nside_out, lmax, niter, SFB_LAYOUT_MMAJOR, out);
}

int32_t sfb_power_win_mix_from_wrlm(const double* w1r_lm, const double* w2r_lm, int64_t nr, int64_t LMAX,
                                    int32_t layout, const double* G, int64_t nmax, int64_t lmax,
                                    const int64_t* lnn, int64_t lnnsize, int64_t lnn_min, int32_t div2Lp1,
                                    int32_t interchange_NN, double* M_out) {
    std::lock_guard<std::mutex> lk(g_mutex);
    SFB_REQUIRE(w1r_lm && M_out, "null pointer");
    SFB_REQUIRE(LMAX == 2 * lmax, "LMAX must equal 2*lmax (src/windows.jl:788)");
    SFB_REQUIRE(layout == 0 || layout == 1, "bad layout");
    PlanGuard pg;
    SFB_TRY(get_cmix_plan(&pg.p, lnn, lnnsize, lnn_min, G, nr, nmax, lmax));
    DevBuf<double> a1, a2, dM;
    SFB_TRY(alm_from_host(w1r_lm, nr, (int)LMAX, layout, a1, pg.p->nrp, 0));
    const bool same = (w2r_lm == nullptr || w2r_lm == w1r_lm);
    if (!same) SFB_TRY(alm_from_host(w2r_lm, nr, (int)LMAX, layout, a2, pg.p->nrp, 0));
    const int64_t n = pg.p->nout;
    SFB_TRY(dM.alloc((size_t)n * n));
    g_times[0] = 0;
    g_times[6] = 0;
    SFB_TRY(cmix_run(pg.p, a1.p, same ? a1.p : a2.p, div2Lp1, interchange_NN, 0, n, 0, n, dM.p, n, 0, nullptr, 0, false,
                     same && mirror_enabled()));
    record_cmix_times(pg.p);
    SFB_TRY(check_finite(dM.p, (size_t)n * n, "mix"));
    SFB_CUDA_OK(cudaMemcpy(M_out, dM.p, (size_t)n * n * sizeof(double), cudaMemcpyDeviceToHost));
    return 0;
}

int32_t sfb_power_win_mix(const double* win1, const double* win2, int64_t nr, int64_t npix_in, int64_t ld_win,
                          int64_t nside, const double* G, int64_t nmax, int64_t lmax, const int64_t* lnn,
                          int64_t lnnsize, int64_t lnn_min, int32_t div2Lp1, int32_t interchange_NN,
                          double* M_out) {
    std::lock_guard<std::mutex> lk(g_mutex);
    SFB_REQUIRE(win1 && M_out, "null pointer");
    Trace tr;
    PlanGuard pg;
    SFB_TRY(get_cmix_plan(&pg.p, lnn, lnnsize, lnn_min, G, nr, nmax, lmax));
    tr.mark("plan");
    DevBuf<double>&a1 = g_ws.alm1, &a2 = g_ws.alm2;
    bool same = true;
    SFB_TRY(windows_to_alm(win1, win2, nr, npix_in, ld_win, nside, 2 * lmax, a1, a2, &same));
    tr.mark("H2D + stage 1");
    SFB_TRY(cmix_to_host_pipelined(pg.p, a1.p, same ? a1.p : a2.p, div2Lp1, interchange_NN, M_out));
    record_cmix_times(pg.p);
    tr.mark("stage 2+3 pipelined with D2H");
    return 0;
}

int32_t sfb_power_win_mix_binned(const double* win1, int64_t nr, int64_t npix_in, int64_t ld_win, int64_t nside,
                                 const double* G, int64_t nmax, int64_t lmax, const int64_t* lnn,
                                 int64_t lnnsize, const int64_t* wt_colptr, const int64_t* wt_rowval,
                                 const double* wt_nzval, int64_t LNN1, const int64_t* v_colptr,
                                 const int64_t* v_rowval, const double* v_nzval, int64_t LNN2, int32_t div2Lp1,
                                 int32_t interchange_NN, double* N_out) {
    std::lock_guard<std::mutex> lk(g_mutex);
    SFB_REQUIRE(win1 && N_out, "null pointer");
    Trace tr;
    PlanGuard pg;
    SFB_TRY(get_cmix_plan(&pg.p, lnn, lnnsize, 1, G, nr, nmax, lmax));
    tr.mark("binned: plan");
    DevBuf<double>&a1 = g_ws.alm1, &a2 = g_ws.alm2, &dM = g_ws.M;
    bool same = true;
    // like the reference, W2r_lm is computed from win1 as well (src/windows.jl:1005-1006)
    SFB_TRY(windows_to_alm(win1, nullptr, nr, npix_in, ld_win, nside, 2 * lmax, a1, a2, &same));
    tr.mark("binned: H2D + stage 1");
    const int64_t n = pg.p->nout;
    SFB_TRY(dM.alloc((size_t)n * n));
    SFB_TRY(cmix_run(pg.p, a1.p, a1.p, div2Lp1, interchange_NN, 0, n, 0, n, dM.p, n, 0, nullptr, 0, false, mirror_enabled()));
    record_cmix_times(pg.p);
    tr.mark("binned: stage 2+3");
    float t_bin = 0;
    SFB_TRY(binned_product_to_host(dM.p, n, wt_colptr, wt_rowval, wt_nzval, LNN1, v_colptr, v_rowval, v_nzval, LNN2,
                                   N_out, &t_bin));
    g_times[7] = t_bin;
    tr.mark("binned: w~ M v + D2H");
    return 0;
}

int32_t sfb_win_lnn(const double* win, int64_t nr, int64_t npix_in, int64_t ld_win, int64_t nside, const double* G,
                    int64_t nmax, int64_t lmax, const int64_t* lnn, int64_t lnnsize, double* Wlnn_out) {
    std::lock_guard<std::mutex> lk(g_mutex);
    SFB_REQUIRE(win && Wlnn_out, "null pointer");
    PlanGuard pg;
    SFB_TRY(get_cmix_plan(&pg.p, lnn, lnnsize, 1, G, nr, nmax, lmax));
    DevBuf<double>&a1 = g_ws.alm1, &a2 = g_ws.alm2;
    bool same = true;
    SFB_TRY(windows_to_alm(win, nullptr, nr, npix_in, ld_win, nside, 2 * lmax, a1, a2, &same));
    // Wr_00 / √(4π) enters under a square root in the reference (src/windows.jl:400): negative values throw there
    std::vector<double> w00((size_t)nr);
    SFB_CUDA_OK(cudaMemcpy(w00.data(), a1.p, (size_t)nr * sizeof(double), cudaMemcpyDeviceToHost));
    for (double v : w00) {
        if (!(v >= 0.0)) {
            set_error("DomainError: sqrt of a negative Wr_00 in win_lnn (src/windows.jl:400)");
            return 4;
        }
    }
    DevBuf<double> d_out;
    SFB_TRY(d_out.alloc((size_t)pg.p->nout));
    SFB_TRY(win_lnn_run(pg.p, a1.p, d_out.p, 0));
    g_times[6] += 1;
    SFB_TRY(check_finite(d_out.p, (size_t)pg.p->nout, "Wlnn"));
    SFB_CUDA_OK(cudaMemcpy(Wlnn_out, d_out.p, (size_t)pg.p->nout * sizeof(double), cudaMemcpyDeviceToHost));
    return 0;
}
int32_t sfb_win_lnn_dev(sfb_cmix_plan* plan, const double* d_alm, double* d_Wlnn, void* stream) {
    std::lock_guard<std::mutex> lk(g_mutex);
    return win_lnn_run(reinterpret_cast<CmixPlan*>(plan), d_alm, d_Wlnn, (cudaStream_t)stream);
}

int32_t sfb_power_win_mix_separable(const double* phi, const double* mask, int64_t nr, int64_t npix_in,
                                    int64_t nside, const double* G, int64_t nmax, int64_t lmax,
                                    const int64_t* lnn, int64_t lnnsize, const int64_t* wt_colptr,
                                    const int64_t* wt_rowval, const double* wt_nzval, int64_t LNN1,
                                    const int64_t* v_colptr, const int64_t* v_rowval, const double* v_nzval,
                                    int64_t LNN2, int32_t div2Lp1, int32_t interchange_NN, double* N_out) {
    std::lock_guard<std::mutex> lk(g_mutex);
    SFB_REQUIRE(phi && mask && N_out, "null pointer");
    int64_t nside_in = 0;
    SFB_TRY(npix2nside(npix_in, &nside_in));
    PlanGuard pg;
    SFB_TRY(get_cmix_plan(&pg.p, lnn, lnnsize, 1, G, nr, nmax, lmax));
    // one map2alm of the mask (src/windows.jl:540-545)
    ShtPlan* sp = nullptr;
    SFB_TRY(get_sht_plan(&sp, nside_in, nside, 2 * lmax, 1));
    DevBuf<double> d_mask, d_wlm, dM;
    SFB_TRY(upload_win(mask, 1, npix_in, 1, d_mask));
    SFB_TRY(d_wlm.alloc(sp->lmsize * 2 * sp->nrp));
    SFB_TRY(sht_map2alm(sp, d_mask.p, 1, 3, d_wlm.p, 0));
    g_times[0] = sp->t_total;
    g_times[6] = sp->launches;
    const int64_t n = pg.p->nout;
    SFB_TRY(dM.alloc((size_t)n * n));
    SFB_TRY(separable_cmix(pg.p, d_wlm.p, sp->nrp, phi, div2Lp1, interchange_NN, dM.p));
    SFB_TRY(check_finite(dM.p, (size_t)n * n, "mix"));
    float t_bin = 0;
    SFB_TRY(binned_product_to_host(dM.p, n, wt_colptr, wt_rowval, wt_nzval, LNN1, v_colptr, v_rowval, v_nzval, LNN2,
                                   N_out, &t_bin));
    g_times[7] = t_bin;
    return 0;
}

int32_t sfb_probe_dmma_tflops(double* tflops) {
    std::lock_guard<std::mutex> lk(g_mutex);
    SFB_REQUIRE(tflops, "null pointer");
    int dev = 0, sms = 0;
    SFB_CUDA_OK(cudaGetDevice(&dev));
    SFB_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    DevBuf<double> d;
    SFB_TRY(d.alloc(1));
    const int blocks = sms * 4, iters = 20000;
    cudaEvent_t e0, e1;
    SFB_CUDA_OK(cudaEventCreate(&e0));
    SFB_CUDA_OK(cudaEventCreate(&e1));
    double best = 0.0;
    for (int rep = 0; rep < 5; ++rep) {
        SFB_CUDA_OK(cudaEventRecord(e0));
        dmma_probe_kernel<<<blocks, 256>>>(d.p, iters, 0.25);
        SFB_CUDA_OK(cudaEventRecord(e1));
        SFB_CUDA_OK(cudaEventSynchronize(e1));
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        const double fl = (double)blocks * 8 /*warps*/ * iters * 8.0 * 512.0;
        if (rep > 0) best = std::max(best, fl / (ms * 1e-3) / 1e12);
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *tflops = best;
    return 0;
}

// ------------------------------------------------------------------------------------------------
// device-resident API

int32_t sfb_sht_plan_create(sfb_sht_plan** plan, int64_t nside_in, int64_t nside_out, int64_t lmax, int64_t nr) {
    std::lock_guard<std::mutex> lk(g_mutex);
    return sht_plan_create(reinterpret_cast<ShtPlan**>(plan), nside_in, nside_out, lmax, nr);
}
int32_t sfb_sht_plan_destroy(sfb_sht_plan* plan) {
    std::lock_guard<std::mutex> lk(g_mutex);
    if (g_pend_sht == reinterpret_cast<ShtPlan*>(plan)) g_pend_sht = nullptr;
    sht_plan_destroy(reinterpret_cast<ShtPlan*>(plan));
    return 0;
}
int64_t sfb_sht_alm_doubles(const sfb_sht_plan* plan) {
    const auto* p = reinterpret_cast<const ShtPlan*>(plan);
    return p ? (int64_t)(p->lmsize * 2 * p->nrp) : 0;
}
int32_t sfb_calc_wr_lm_dev(sfb_sht_plan* plan, const double* d_win, int64_t ld_win, int64_t niter, double* d_alm,
                           void* stream) {
    std::lock_guard<std::mutex> lk(g_mutex);
    auto* p = reinterpret_cast<ShtPlan*>(plan);
    SFB_REQUIRE(p, "null plan");
    p->async_times = async_enabled();   // no host sync: the caller can enqueue stage 2+3 behind stage 1
    const int rc = sht_map2alm(p, d_win, ld_win, (int)niter, d_alm, (cudaStream_t)stream);
    p->async_times = false;
    SFB_TRY(rc);
    g_pend_sht = p;
    g_times[6] = p->launches;
    return 0;
}
int32_t sfb_alm_to_complex_dev(const sfb_sht_plan* plan, const double* d_alm, int32_t layout, double* d_out,
                               void* stream) {
    std::lock_guard<std::mutex> lk(g_mutex);
    return sht_alm_to_complex(reinterpret_cast<const ShtPlan*>(plan), d_alm, layout, d_out, (cudaStream_t)stream);
}

int32_t sfb_cmix_plan_create(sfb_cmix_plan** plan, const int64_t* lnn, int64_t lnnsize, int64_t lnn_min,
                             const double* G, int64_t nr, int64_t nmax, int64_t lmax) {
    std::lock_guard<std::mutex> lk(g_mutex);
    return cmix_plan_create(reinterpret_cast<CmixPlan**>(plan), lnn, lnnsize, lnn_min, G, nr, nmax, lmax);
}
int32_t sfb_cmix_plan_destroy(sfb_cmix_plan* plan) {
    std::lock_guard<std::mutex> lk(g_mutex);
    if (g_pend_cmix == reinterpret_cast<CmixPlan*>(plan)) g_pend_cmix = nullptr;
    cmix_plan_destroy(reinterpret_cast<CmixPlan*>(plan));
    return 0;
}
int32_t sfb_power_win_mix_dev(sfb_cmix_plan* plan, const double* d_alm1, const double* d_alm2, int32_t div2Lp1,
                              int32_t interchange_NN, int64_t row_lo, int64_t row_hi, double* d_M, int64_t ldM,
                              void* stream) {
    std::lock_guard<std::mutex> lk(g_mutex);
    auto* p = reinterpret_cast<CmixPlan*>(plan);
    SFB_REQUIRE(p, "null plan");
    const double t0 = g_times[6];
    const bool mirror = (d_alm1 == d_alm2) && row_lo == 0 && row_hi == p->nout && ldM >= p->nout && mirror_enabled();
    SFB_CMIX_RUN_ASYNC(p, cmix_run(p, d_alm1, d_alm2, div2Lp1, interchange_NN, row_lo, row_hi, 0, p->nout, d_M, ldM,
                                   (cudaStream_t)stream, nullptr, 0, false, mirror));
    g_times[6] = t0 + p->launches;
    return 0;
}
int32_t sfb_power_win_mix_block_dev(sfb_cmix_plan* plan, const double* d_alm1, const double* d_alm2, int32_t div2Lp1,
                                    int32_t interchange_NN, int64_t row_lo, int64_t row_hi, int64_t col_lo,
                                    int64_t col_hi, double* d_M, int64_t ldM, void* stream) {
    std::lock_guard<std::mutex> lk(g_mutex);
    auto* p = reinterpret_cast<CmixPlan*>(plan);
    SFB_REQUIRE(p, "null plan");
    const double t0 = g_times[6];
    SFB_CMIX_RUN_ASYNC(p, cmix_run(p, d_alm1, d_alm2, div2Lp1, interchange_NN, row_lo, row_hi, col_lo, col_hi, d_M, ldM,
                                   (cudaStream_t)stream));
    g_times[6] = t0 + p->launches;
    return 0;
}
int32_t sfb_power_win_mix_dev_peers(sfb_cmix_plan* plan, const double* d_alm1, const double* d_alm2, int32_t div2Lp1,
                                    int32_t interchange_NN, int64_t row_lo, int64_t row_hi, double* d_M_full,
                                    double* const* peer_M_full, int32_t npeers, int64_t ldM, void* stream) {
    std::lock_guard<std::mutex> lk(g_mutex);
    auto* p = reinterpret_cast<CmixPlan*>(plan);
    SFB_REQUIRE(p && d_M_full, "null pointer");
    SFB_REQUIRE(npeers >= 0 && npeers <= 7, "at most 7 peers");
    SFB_REQUIRE(ldM >= p->nout, "ldM must be the leading dimension of the full matrix");
    double* peers[7] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    for (int i = 0; i < npeers; ++i) {
        SFB_REQUIRE(peer_M_full && peer_M_full[i], "null peer pointer");
        peers[i] = peer_M_full[i] + row_lo;
    }
    const double t0 = g_times[6];
    SFB_CMIX_RUN_ASYNC(p, cmix_run(p, d_alm1, d_alm2, div2Lp1, interchange_NN, row_lo, row_hi, 0, p->nout,
                                   d_M_full + row_lo, ldM, (cudaStream_t)stream, peers, npeers));
    g_times[6] = t0 + p->launches;
    return 0;
}

// Push rows [row_lo,row_hi) of this device's full matrix into the same rows of every peer's full matrix with one
// pitched peer-to-peer copy per peer (copy engines over NVLink), each on its own stream, ordered after `stream`.
int32_t sfb_push_rows_to_peers(const double* d_M_full, double* const* peer_M_full, int32_t npeers, int64_t row_lo,
                               int64_t row_hi, int64_t ncols, int64_t ldM, void* stream) {
    std::lock_guard<std::mutex> lk(g_mutex);
    SFB_REQUIRE(d_M_full && npeers >= 0 && npeers <= 7, "bad arguments");
    SFB_REQUIRE(0 <= row_lo && row_lo <= row_hi && row_hi <= ldM, "bad row range");
    if (row_hi == row_lo || npeers == 0) return 0;
    static cudaStream_t ps[7] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    static cudaEvent_t ready = nullptr, done[7];
    static int ps_dev = -1;
    int dev = 0;
    SFB_CUDA_OK(cudaGetDevice(&dev));
    if (!ready || ps_dev != dev) {
        SFB_CUDA_OK(cudaEventCreateWithFlags(&ready, cudaEventDisableTiming));
        for (int i = 0; i < 7; ++i) {
            SFB_CUDA_OK(cudaStreamCreateWithFlags(&ps[i], cudaStreamNonBlocking));
            SFB_CUDA_OK(cudaEventCreateWithFlags(&done[i], cudaEventDisableTiming));
        }
        ps_dev = dev;
    }
    cudaStream_t st = (cudaStream_t)stream;
    SFB_CUDA_OK(cudaEventRecord(ready, st));
    const size_t pitch = (size_t)ldM * sizeof(double), width = (size_t)(row_hi - row_lo) * sizeof(double);
    for (int i = 0; i < npeers; ++i) {
        SFB_REQUIRE(peer_M_full && peer_M_full[i], "null peer pointer");
        SFB_CUDA_OK(cudaStreamWaitEvent(ps[i], ready, 0));
        SFB_CUDA_OK(cudaMemcpy2DAsync(peer_M_full[i] + row_lo, pitch, d_M_full + row_lo, pitch, width, (size_t)ncols,
                                      cudaMemcpyDeviceToDevice, ps[i]));
        SFB_CUDA_OK(cudaEventRecord(done[i], ps[i]));
        SFB_CUDA_OK(cudaStreamWaitEvent(st, done[i], 0));
    }
    return 0;
}

// Column slabs are contiguous in a column-major matrix: one plain peer copy per peer.
int32_t sfb_push_cols_to_peers(const double* d_M_full, double* const* peer_M_full, int32_t npeers, int64_t col_lo,
                               int64_t col_hi, int64_t ldM, void* stream) {
    SFB_REQUIRE(d_M_full && npeers >= 0 && npeers <= 7 && col_lo >= 0 && col_lo <= col_hi, "bad arguments");
    if (col_hi == col_lo || npeers == 0) return 0;
    double* shifted[7];
    for (int i = 0; i < npeers; ++i) {
        SFB_REQUIRE(peer_M_full && peer_M_full[i], "null peer pointer");
        shifted[i] = peer_M_full[i] + col_lo * ldM;
    }
    // the slab is one contiguous run of (col_hi-col_lo)*ldM doubles == a single "row" of a pitched copy
    const int64_t run = (col_hi - col_lo) * ldM;
    return sfb_push_rows_to_peers(d_M_full + col_lo * ldM, shifted, npeers, 0, run, 1, run, stream);
}

int32_t sfb_ipc_alloc(void** dptr, int64_t bytes, void* handle64) {
    std::lock_guard<std::mutex> lk(g_mutex);
    SFB_REQUIRE(dptr && handle64 && bytes > 0, "bad arguments");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    SFB_CUDA_OK(cudaMalloc(dptr, (size_t)bytes));
    cudaIpcMemHandle_t h;
    SFB_CUDA_OK(cudaIpcGetMemHandle(&h, *dptr));
    std::memcpy(handle64, &h, 64);
    return 0;
}
int32_t sfb_ipc_open(const void* handle64, void** dptr) {
    std::lock_guard<std::mutex> lk(g_mutex);
    SFB_REQUIRE(dptr && handle64, "bad arguments");
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle64, 64);
    SFB_CUDA_OK(cudaIpcOpenMemHandle(dptr, h, cudaIpcMemLazyEnablePeerAccess));
    return 0;
}
int32_t sfb_ipc_close(void* dptr) {
    std::lock_guard<std::mutex> lk(g_mutex);
    SFB_CUDA_OK(cudaIpcCloseMemHandle(dptr));
    return 0;
}
int32_t sfb_ipc_free(void* dptr) {
    std::lock_guard<std::mutex> lk(g_mutex);
    SFB_CUDA_OK(cudaFree(dptr));
    return 0;
}
int32_t sfb_memcpy_dev(void* dst, const void* src, int64_t bytes, void* stream) {
    SFB_CUDA_OK(cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return 0;
}

int32_t sfb_power_win_mix_upper_packed_dev(sfb_cmix_plan* plan, const double* d_alm, int32_t div2Lp1,
                                           int32_t interchange_NN, int64_t col_lo, int64_t col_hi, double* d_packed,
                                           void* stream) {
    std::lock_guard<std::mutex> lk(g_mutex);
    auto* p = reinterpret_cast<CmixPlan*>(plan);
    SFB_REQUIRE(p && d_alm && d_packed, "null pointer");
    const double t0 = g_times[6];
    SFB_CMIX_RUN_ASYNC(p, cmix_run(p, d_alm, d_alm, div2Lp1, interchange_NN, 0, p->nout, col_lo, col_hi, d_packed, p->nout,
                                   (cudaStream_t)stream, nullptr, 0, false, false, true));
    g_times[6] = t0 + p->launches;
    return 0;
}
int32_t sfb_cmix_unpack_mirror_dev(sfb_cmix_plan* plan, const double* d_packed, int32_t div2Lp1, int32_t interchange_NN,
                                   double* d_M, int64_t ldM, void* stream) {
    std::lock_guard<std::mutex> lk(g_mutex);
    auto* p = reinterpret_cast<CmixPlan*>(plan);
    const double* bases[1] = {d_packed};
    SFB_TRY(cmix_unpack_mirror(p, bases, nullptr, 1, 0, div2Lp1, interchange_NN, d_M, ldM, (cudaStream_t)stream));
    g_times[6] += 1;
    return 0;
}
int32_t sfb_cmix_unpack_mirror_peers_dev(sfb_cmix_plan* plan, const double* const* packed_of_rank,
                                         const int64_t* col_bounds, int32_t nranks, int32_t my_rank,
                                         int32_t div2Lp1, int32_t interchange_NN, double* d_M, int64_t ldM,
                                         void* stream) {
    std::lock_guard<std::mutex> lk(g_mutex);
    auto* p = reinterpret_cast<CmixPlan*>(plan);
    SFB_REQUIRE(col_bounds, "null pointer");
    SFB_TRY(cmix_unpack_mirror(p, packed_of_rank, col_bounds, nranks, my_rank, div2Lp1, interchange_NN, d_M, ldM,
                               (cudaStream_t)stream));
    g_times[6] += 1;
    return 0;
}
int32_t sfb_cmix_packed_offsets(const sfb_cmix_plan* plan, int64_t* offsets, int64_t n_plus_1) {
    const auto* p = reinterpret_cast<const CmixPlan*>(plan);
    SFB_REQUIRE(p && offsets && n_plus_1 == p->nout + 1, "sfb_cmix_packed_offsets: bad arguments");
    SFB_REQUIRE(p->ell_sorted && (int64_t)p->h_colbase.size() == p->nout + 1,
                "sfb_cmix_packed_offsets: upper-packed storage needs an lnn table sorted by l");
    for (int64_t i = 0; i <= p->nout; ++i) offsets[i] = p->h_colbase[i];
    return 0;
}
// cost of each output column when only the blocks with l <= L are formed (upper-packed / mirror mode)
int32_t sfb_cmix_col_costs_upper(const sfb_cmix_plan* plan, double* cost, int64_t n) {
    const auto* p = reinterpret_cast<const CmixPlan*>(plan);
    SFB_REQUIRE(p && cost && n == p->nout, "sfb_cmix_col_costs_upper: bad arguments");
    for (int L = 0; L <= p->lmax; ++L) {
        const int cols = p->ell_ptr[L + 1] - p->ell_ptr[L];
        if (!cols) continue;
        const double b = p->a_of_ell[L];
        double c = 0;
        for (int l = 0; l <= L; ++l) {
            if (p->ell_ptr[l + 1] == p->ell_ptr[l]) continue;
            const double ap = 8.0 * ((p->a_of_ell[l] + 7) / 8);
            // Ŵ_{lL} is built with plain FMAs at 6.5 TFLOP/s against 29 for the DMMA block kernel (fit to per-rank timings, cfg4)
            c += 2.0 * ap * p->nrp * p->nrp * b + 2.0 * ap * ap * p->nrp * b * (b + 1) / 2 +
                 4.5 * 2.0 * p->nrp * p->nrp * (l + 1);
        }
        for (int s = p->ell_ptr[L]; s < p->ell_ptr[L + 1]; ++s) cost[p->h_row_out[s]] = c / cols;
    }
    return 0;
}

// cost of each output COLUMN (L,N,N'): the block (l,L) costs the same whichever way it is assigned, so the
// column cost is the transpose of the row model
int32_t sfb_cmix_col_costs(const sfb_cmix_plan* plan, double* cost, int64_t n) {
    const auto* p = reinterpret_cast<const CmixPlan*>(plan);
    SFB_REQUIRE(p && cost && n == p->nout, "sfb_cmix_col_costs: bad arguments");
    for (int L = 0; L <= p->lmax; ++L) {
        const int cols = p->ell_ptr[L + 1] - p->ell_ptr[L];
        if (!cols) continue;
        const double b = p->a_of_ell[L];
        double c = 0;
        for (int l = 0; l <= p->lmax; ++l) {
            if (p->ell_ptr[l + 1] == p->ell_ptr[l]) continue;
            const double ap = 8.0 * ((p->a_of_ell[l] + 7) / 8);
            c += 2.0 * ap * p->nrp * p->nrp * b + 2.0 * ap * ap * p->nrp * b * (b + 1) / 2 +
                 2.0 * p->nrp * p->nrp * (std::min(l, L) + 1);
        }
        for (int s = p->ell_ptr[L]; s < p->ell_ptr[L + 1]; ++s) cost[p->h_row_out[s]] = c / cols;
    }
    return 0;
}

int32_t sfb_cmix_row_costs(const sfb_cmix_plan* plan, double* cost, int64_t n) {
    const auto* p = reinterpret_cast<const CmixPlan*>(plan);
    SFB_REQUIRE(p && cost && n == p->nout, "sfb_cmix_row_costs: bad arguments");
    for (int l = 0; l <= p->lmax; ++l) {
        const int rows = p->ell_ptr[l + 1] - p->ell_ptr[l];
        if (!rows) continue;
        const double ap = 8.0 * ((p->a_of_ell[l] + 7) / 8);
        double c = 0;
        for (int L = 0; L <= p->lmax; ++L) {
            const double b = p->a_of_ell[L];
            c += 2.0 * ap * p->nrp * p->nrp * b + 2.0 * ap * ap * p->nrp * b * (b + 1) / 2 +
                 2.0 * p->nrp * p->nrp * (std::min(l, L) + 1);
        }
        for (int s = p->ell_ptr[l]; s < p->ell_ptr[l + 1]; ++s) cost[p->h_row_out[s]] = c / rows;
    }
    return 0;
}

}  // extern "C"
