// extern "C" entry points declared in include/sfb_b200.h.
#include "../../include/sfb_b200.h"

#include "binned.cuh"
#include "cmix.cuh"
#include "common.cuh"
#include "lusolve.cuh"
#include "sfbt.cuh"
#include "sht.cuh"
#include "wmix.cuh"

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

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
static double g_times[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
// plans whose last (device-resident, asynchronous) run has unread timing events: resolved by sfb_get_timings
static ShtPlan* g_pend_sht = nullptr;
static CmixPlan* g_pend_cmix = nullptr;

// ---- per-device state: cached plans and the reusable workspace are keyed by the CUDA device they live on ----
constexpr int kMaxDev = 8;        // devices driven by one call (sfb_set_devices)
constexpr int kMaxDevId = 64;     // highest CUDA device ordinal + 1 this library keeps state for
static int g_ndev = 1;            // sfb_set_devices(n): the host-pointer entry points shard over devices 0..n-1

struct ShtCache {
    ShtPlan* p = nullptr;
};
static ShtCache g_sht_cache[kMaxDevId];

static int current_device(int* dev) {
    SFB_CUDA_OK(cudaGetDevice(dev));
    SFB_REQUIRE(*dev >= 0 && *dev < kMaxDevId, "CUDA device ordinal out of range");
    return 0;
}

// cached stage-1 plan of the current device (tables depend only on nside/lmax/nr)
static int get_sht_plan(ShtPlan** out, int64_t nside_in, int64_t nside_out, int64_t lmax, int64_t nr) {
    int dev = 0;
    SFB_TRY(current_device(&dev));
    ShtPlan*& g = g_sht_cache[dev].p;
    if (g && g->nside_in == nside_in && g->nside == nside_out && g->lmax == lmax && g->nr == nr) {
        *out = g;
        return 0;
    }
    if (g) {
        sht_plan_destroy(g);
        g = nullptr;
    }
    SFB_TRY(sht_plan_create(&g, nside_in, nside_out, lmax, nr));
    *out = g;
    return 0;
}

static int npix2nside(int64_t npix, int64_t* nside) {
    int64_t ns = (int64_t)llround(std::sqrt((double)npix / 12.0));
    SFB_REQUIRE(ns >= 1 && 12 * ns * ns == npix, "npix is not 12*nside^2");
    *nside = ns;
    return 0;
}

// ---- pageable host buffers ----------------------------------------------------------------------------------------------
// cudaMemcpyAsync to or from pageable memory is staged by the driver through its own pinned buffers with a single-threaded
// memcpy (≈ 18 GB/s for cfg4's 3.75 GB matrix against 56 GB/s of the link).  When a host-pointer entry point is handed a
// pageable array (a plain Julia Matrix{Float64}), the library stages it itself: pieces of 32 MB alternate between two
// pinned buffers, the DMA of piece i runs while a small pool of host threads moves piece i-1 between its pinned buffer and
// the caller's array.  Page-locked buffers (sfb_host_alloc / sfb_host_register) take the direct path.
// copy with non-temporal stores: the destination is written once and not read again by this thread, so the
// read-for-ownership of a cached store (a third of the memory traffic of a plain memcpy) is avoided
#if defined(__x86_64__)
#include <immintrin.h>
__attribute__((target("avx2"))) static void stream_copy_avx2(char* d, const char* s2, size_t n) {
    size_t i = 0;
    for (; i + 128 <= n; i += 128) {
        const __m256i a = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(s2 + i));
        const __m256i b = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(s2 + i + 32));
        const __m256i c = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(s2 + i + 64));
        const __m256i e = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(s2 + i + 96));
        _mm256_stream_si256(reinterpret_cast<__m256i*>(d + i), a);
        _mm256_stream_si256(reinterpret_cast<__m256i*>(d + i + 32), b);
        _mm256_stream_si256(reinterpret_cast<__m256i*>(d + i + 64), c);
        _mm256_stream_si256(reinterpret_cast<__m256i*>(d + i + 96), e);
    }
    _mm_sfence();
    if (i < n) std::memcpy(d + i, s2 + i, n - i);
}
#endif
static void host_copy(char* d, const char* s2, size_t n) {
#if defined(__x86_64__)
    static const bool avx2 = __builtin_cpu_supports("avx2") && getenv("SFB_NO_NT_COPY") == nullptr;
    if (avx2 && n >= 4096) {
        const size_t head = (32 - (reinterpret_cast<uintptr_t>(d) & 31)) & 31;   // align the destination to 32 bytes
        if (head) std::memcpy(d, s2, head);
        stream_copy_avx2(d + head, s2 + head, n - head);
        return;
    }
#endif
    std::memcpy(d, s2, n);
}

class HostPool {
public:
    explicit HostPool(int n) : n_(n) {
        for (int i = 0; i < n_; ++i) th_.emplace_back([this, i] { run(i); });
    }
    ~HostPool() {
        {
            std::lock_guard<std::mutex> lk(m_);
            stop_ = true;
        }
        cv_.notify_all();
        for (auto& t : th_) t.join();
    }
    void copy(void* dst, const void* src, size_t bytes) {   // blocking, split over the pool
        if (bytes < (size_t(1) << 20) || n_ == 0) {
            std::memcpy(dst, src, bytes);
            return;
        }
        std::unique_lock<std::mutex> lk(m_);
        dst_ = static_cast<char*>(dst), src_ = static_cast<const char*>(src), bytes_ = bytes;
        pending_ = n_;
        ++gen_;
        cv_.notify_all();
        done_.wait(lk, [this] { return pending_ == 0; });
    }
private:
    void run(int i) {
        unsigned long long seen = 0;
        for (;;) {
            std::unique_lock<std::mutex> lk(m_);
            cv_.wait(lk, [&] { return stop_ || gen_ != seen; });
            if (stop_) return;
            seen = gen_;
            char* d = dst_;
            const char* s2 = src_;
            const size_t b = bytes_;
            lk.unlock();
            const size_t per = ((b / n_) + 4095) & ~size_t(4095), o = std::min(b, per * i), e = std::min(b, o + per);
            if (e > o) host_copy(d + o, s2 + o, e - o);
            lk.lock();
            if (--pending_ == 0) done_.notify_one();
        }
    }
    int n_;
    std::vector<std::thread> th_;
    std::mutex m_;
    std::condition_variable cv_, done_;
    bool stop_ = false;
    unsigned long long gen_ = 0;
    char* dst_ = nullptr;
    const char* src_ = nullptr;
    size_t bytes_ = 0;
    int pending_ = 0;
};

struct HostStage {
    static constexpr size_t kPiece = size_t(32) << 20;
    void* buf[2] = {nullptr, nullptr};
    cudaEvent_t evdev[64][2] = {};      // events belong to a device: one pair per device that used the ring
    cudaEvent_t* ev = nullptr;          // the pair of the current device (set by init)
    HostPool* pool = nullptr;
    int init() {
        int dev = 0;
        SFB_CUDA_OK(cudaGetDevice(&dev));
        SFB_REQUIRE(dev >= 0 && dev < 64, "device index out of range");
        if (!evdev[dev][0])
            for (int i = 0; i < 2; ++i) SFB_CUDA_OK(cudaEventCreateWithFlags(&evdev[dev][i], cudaEventDisableTiming));
        ev = evdev[dev];
        if (pool) return 0;
        for (int i = 0; i < 2; ++i) SFB_CUDA_OK(cudaHostAlloc(&buf[i], kPiece, cudaHostAllocPortable));
        const unsigned hc = std::thread::hardware_concurrency();
        int nt = (int)std::max(1u, std::min(8u, hc ? hc : 4u));   // 8 threads saturate the host copy (16: no gain)
        if (getenv("SFB_COPY_THREADS")) nt = std::max(1, atoi(getenv("SFB_COPY_THREADS")));
        pool = new HostPool(nt);
        return 0;
    }
};
static HostStage g_stage;

static bool is_pageable(const void* p) {
    if (getenv("SFB_NO_STAGING")) return false;
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return true;
    }
    return a.type == cudaMemoryTypeUnregistered;
}

// device -> pageable host, blocking; `st` must already be ordered behind the producer of `src`
static int d2h_staged(void* dst, const void* src, size_t bytes, cudaStream_t st) {
    SFB_TRY(g_stage.init());
    char* d = static_cast<char*>(dst);
    const char* s2 = static_cast<const char*>(src);
    const size_t P = HostStage::kPiece;
    const size_t np = (bytes + P - 1) / P;
    for (size_t i = 0; i <= np; ++i) {
        if (i < np) {   // DMA of piece i into its pinned buffer (free: piece i-2 left it in the previous iteration)
            const size_t o = i * P, n = std::min(P, bytes - o);
            SFB_CUDA_OK(cudaMemcpyAsync(g_stage.buf[i & 1], s2 + o, n, cudaMemcpyDeviceToHost, st));
            SFB_CUDA_OK(cudaEventRecord(g_stage.ev[i & 1], st));
        }
        if (i >= 1) {   // piece i-1 out of its pinned buffer while piece i is in flight
            const size_t o = (i - 1) * P, n = std::min(P, bytes - o);
            SFB_CUDA_OK(cudaEventSynchronize(g_stage.ev[(i - 1) & 1]));
            g_stage.pool->copy(d + o, g_stage.buf[(i - 1) & 1], n);
        }
    }
    return 0;
}

// pageable host -> device, blocking
static int h2d_staged(void* dst, const void* src, size_t bytes, cudaStream_t st) {
    SFB_TRY(g_stage.init());
    char* d = static_cast<char*>(dst);
    const char* s2 = static_cast<const char*>(src);
    const size_t P = HostStage::kPiece;
    const size_t np = (bytes + P - 1) / P;
    for (size_t i = 0; i < np; ++i) {
        const size_t o = i * P, n = std::min(P, bytes - o);
        if (i >= 2) SFB_CUDA_OK(cudaEventSynchronize(g_stage.ev[i & 1]));   // the DMA of piece i-2 has drained the buffer
        g_stage.pool->copy(g_stage.buf[i & 1], s2 + o, n);
        SFB_CUDA_OK(cudaMemcpyAsync(d + o, g_stage.buf[i & 1], n, cudaMemcpyHostToDevice, st));
        SFB_CUDA_OK(cudaEventRecord(g_stage.ev[i & 1], st));
    }
    SFB_CUDA_OK(cudaStreamSynchronize(st));
    return 0;
}

// blocking device -> host copy on the legacy stream: staged by the library when the destination is pageable
static int d2h_auto(void* dst, const void* src, size_t bytes) {
    if (is_pageable(dst)) {
        SFB_CUDA_OK(cudaDeviceSynchronize());
        SFB_TRY(d2h_staged(dst, src, bytes, 0));
        return 0;
    }
    SFB_CUDA_OK(cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost));
    return 0;
}

// H2D of the Julia array win (nr x npix, leading dimension ld) -> device [pixel][nr]
static int h2d_rows(double* dst, const double* src, int64_t nr, int64_t npix, int64_t ld, cudaStream_t st,
                    bool stage_ok = false) {   // stage_ok: single-device host path (the staging ring is not shared)
    if (ld == nr && stage_ok && is_pageable(src))
        SFB_TRY(h2d_staged(dst, src, (size_t)nr * npix * sizeof(double), st));
    else if (ld == nr)
        SFB_CUDA_OK(cudaMemcpyAsync(dst, src, (size_t)nr * npix * sizeof(double), cudaMemcpyHostToDevice, st));
    else
        SFB_CUDA_OK(cudaMemcpy2DAsync(dst, nr * sizeof(double), src, ld * sizeof(double), nr * sizeof(double), (size_t)npix,
                                      cudaMemcpyHostToDevice, st));
    return 0;
}
static int upload_win(const double* win, int64_t nr, int64_t npix, int64_t ld, DevBuf<double>& d, cudaStream_t st) {
    SFB_REQUIRE(win, "win is null");
    SFB_REQUIRE(ld >= nr, "ld_win < nr");
    SFB_TRY(d.alloc((size_t)nr * npix));
    return h2d_rows(d.p, win, nr, npix, ld, st, true);
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
    g_times[8] = p->t_k3;
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

// Device workspace reused across host-pointer calls (cudaMalloc / cudaFree of multi-GB buffers is slow); one per
// device, so that a call after sfb_set_device(other) never touches buffers of the previous device.
struct Workspace {
    DevBuf<double> win, alm1, alm2, slab[2], M;
    DevBuf<double> slice[2], shard, almshard[2];   // multi-device runs: pixel slice of win1/win2, shell shard, alm shard
    DevBuf<int> flag;
    cudaStream_t main = nullptr, copy = nullptr;
    cudaEvent_t computed[2] = {nullptr, nullptr}, copied[2] = {nullptr, nullptr};
    cudaEvent_t ev_slice[2] = {nullptr, nullptr}, ev_alm[2] = {nullptr, nullptr};
    bool ready = false;
    int init() {   // on the current device
        if (ready) return 0;
        SFB_CUDA_OK(cudaStreamCreateWithFlags(&main, cudaStreamNonBlocking));
        SFB_CUDA_OK(cudaStreamCreateWithFlags(&copy, cudaStreamNonBlocking));
        for (int i = 0; i < 2; ++i) {
            SFB_CUDA_OK(cudaEventCreateWithFlags(&computed[i], cudaEventDisableTiming));
            SFB_CUDA_OK(cudaEventCreateWithFlags(&copied[i], cudaEventDisableTiming));
            SFB_CUDA_OK(cudaEventCreateWithFlags(&ev_slice[i], cudaEventDisableTiming));
            SFB_CUDA_OK(cudaEventCreateWithFlags(&ev_alm[i], cudaEventDisableTiming));
        }
        ready = true;
        return 0;
    }
};
static Workspace g_ws_dev[kMaxDevId];
static int get_ws(Workspace** out) {
    int dev = 0;
    SFB_TRY(current_device(&dev));
    SFB_TRY(g_ws_dev[dev].init());
    *out = &g_ws_dev[dev];
    return 0;
}

// Stage 2+3 pipelined with the device->host copy: the matrix is produced in column slabs (contiguous in column-major
// order); while slab k is copied to the caller's buffer, slab k+1 is computed.
static bool mirror_enabled() { return getenv("SFB_NO_MIRROR") == nullptr; }

struct CmixTimes {
    float wl = 0, what = 0, block = 0, fill = 0, k3 = 0;
    double flops = 0;
    int launches = 0;
    void add(const CmixPlan* p) {
        wl += p->t_wl;
        what += p->t_what;
        block += p->t_block;
        fill += p->t_fill;
        k3 += p->t_k3;
        flops += p->flops_executed;
        launches += p->launches;
    }
    void store(CmixPlan* p) const {
        p->t_wl = wl;
        p->t_what = what;
        p->t_block = block;
        p->t_fill = fill;
        p->t_k3 = k3;
        p->flops_executed = flops;
        p->launches = launches;
    }
};

// Columns [col_lo, col_hi) (all rows) of M -> M_out + col_lo * n on the host.  Whole matrix + auto-correlation: mirror
// mode (step k forms the blocks (l in chunk k, L >= l) and their mirror images in the full device matrix, after which
// the columns of chunk k are complete and can leave); otherwise plain full-height column slabs.
static int cmix_cols_to_host(CmixPlan* p, Workspace& ws, const double* a1, const double* a2, int div2Lp1, int interchange,
                             int64_t col_lo, int64_t col_hi, int nchunks, double* M_out, bool stage_ok = false) {
    // stage_ok: single-device host path only — the staging ring and its host threads are one per process, and the workers
    // of a multi-device call run concurrently on other devices (their pageable copies are left to the driver)
    const int64_t n = p->nout;
    if (col_hi <= col_lo) return 0;
    const bool whole = (col_lo == 0 && col_hi == n);
    const bool mirror = whole && (a1 == a2) && mirror_enabled();
    const auto chunks = mirror ? cmix_row_chunks_mirror(p, nchunks) : cmix_col_chunks_range(p, col_lo, col_hi, nchunks);
    int64_t maxc = 0;
    for (auto& c : chunks) maxc = std::max(maxc, c.second - c.first);
    if (mirror) {
        SFB_TRY(ws.M.alloc((size_t)n * n));
    } else {
        for (int i = 0; i < 2 && i < (int)chunks.size(); ++i) SFB_TRY(ws.slab[i].alloc((size_t)maxc * n));
    }
    SFB_TRY(ws.flag.alloc(1));
    cudaStream_t st = ws.main;
    SFB_CUDA_OK(cudaMemsetAsync(ws.flag.p, 0, sizeof(int), st));
    CmixTimes tt;
    const bool staged = stage_ok && is_pageable(M_out);
    bool have_prev = false;
    int64_t prev_c0 = 0;
    const double* prev_src = nullptr;
    size_t prev_bytes = 0;
    for (size_t k = 0; k < chunks.size(); ++k) {
        const int b = (int)(k & 1);
        const int64_t c0 = chunks[k].first, c1 = chunks[k].second;
        double* src = nullptr;
        if (mirror) {
            SFB_TRY(cmix_run(p, a1, a2, div2Lp1, interchange, c0, c1, 0, n, ws.M.p, n, st, nullptr, 0, k > 0, true));
            src = ws.M.p + c0 * n;
        } else {
            if (k >= 2) SFB_CUDA_OK(cudaEventSynchronize(ws.copied[b]));  // slab free again
            SFB_TRY(cmix_run(p, a1, a2, div2Lp1, interchange, 0, n, c0, c1, ws.slab[b].p, n, st, nullptr, 0, k > 0));
            src = ws.slab[b].p;
        }
        tt.add(p);
        finite_check_kernel<<<512, 256, 0, st>>>(src, (size_t)(c1 - c0) * n, ws.flag.p);
        SFB_CUDA_OK(cudaEventRecord(ws.computed[b], st));
        if (staged) {
            // pageable result: the blocking staged copy of slab k-1 runs here, after slab k's kernels have been enqueued, so
            // the device keeps computing under it; the copy stream then waits for slab k (copied in the next iteration)
            if (have_prev) {
                SFB_TRY(d2h_staged(M_out + prev_c0 * n, prev_src, prev_bytes, ws.copy));
                SFB_CUDA_OK(cudaEventRecord(ws.copied[b ^ 1], ws.copy));
            }
            prev_c0 = c0, prev_src = src, prev_bytes = (size_t)(c1 - c0) * n * sizeof(double), have_prev = true;
            SFB_CUDA_OK(cudaStreamWaitEvent(ws.copy, ws.computed[b], 0));
            continue;
        }
        SFB_CUDA_OK(cudaStreamWaitEvent(ws.copy, ws.computed[b], 0));
        SFB_CUDA_OK(cudaMemcpyAsync(M_out + c0 * n, src, (size_t)(c1 - c0) * n * sizeof(double), cudaMemcpyDeviceToHost,
                                    ws.copy));
        SFB_CUDA_OK(cudaEventRecord(ws.copied[b], ws.copy));
    }
    if (staged && have_prev) SFB_TRY(d2h_staged(M_out + prev_c0 * n, prev_src, prev_bytes, ws.copy));
    SFB_CUDA_OK(cudaStreamSynchronize(ws.copy));
    tt.launches += (int)chunks.size();
    tt.store(p);
    int h = 0;
    SFB_CUDA_OK(cudaMemcpyAsync(&h, ws.flag.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    SFB_CUDA_OK(cudaStreamSynchronize(st));
    if (h) {  // @assert all(isfinite.(mix))  src/windows.jl:803
        set_error("AssertionError: all(isfinite.(mix))");
        return 4;
    }
    return 0;
}

// Host-pointer entry points reuse the stage-2/3 plan (index tables, radial basis, 3j table, Ŵ workspace) when the
// caller passes the same tables again: keyed by the sizes and an FNV-1a hash of the lnn and G bytes, per device.
static uint64_t fnv1a(const void* data, size_t bytes, uint64_t h = 1469598103934665603ull) {
    const uint64_t* w = static_cast<const uint64_t*>(data);
    for (size_t i = 0; i < bytes / 8; ++i) {
        h ^= w[i];
        h *= 1099511628211ull;
    }
    return h;
}
struct PlanKey {
    uint64_t hash = 0;
    int64_t dims[5] = {0, 0, 0, 0, 0};
    bool operator==(const PlanKey& o) const { return hash == o.hash && std::memcmp(dims, o.dims, sizeof(dims)) == 0; }
};
struct PlanCache {
    CmixPlan* p = nullptr;
    PlanKey key;
};
static PlanCache g_plan_cache[kMaxDevId];

struct PlanGuard {  // non-owning handle on the cached plan
    CmixPlan* p = nullptr;
};

static int make_plan_key(PlanKey* k, const int64_t* lnn, int64_t lnnsize, int64_t lnn_min, const double* G, int64_t nr,
                         int64_t nmax, int64_t lmax) {
    SFB_REQUIRE(lnn && G && lnnsize >= 1 && nr >= 1 && nmax >= 1 && lmax >= 0, "bad mode tables");
    k->hash = fnv1a(lnn, (size_t)lnnsize * 3 * sizeof(int64_t));
    k->hash = fnv1a(G, (size_t)nr * nmax * (lmax + 1) * sizeof(double), k->hash);
    const int64_t dims[5] = {lnnsize, lnn_min, nr, nmax, lmax};
    std::memcpy(k->dims, dims, sizeof(dims));
    return 0;
}

// cached stage-2/3 plan of the current device
static int get_cmix_plan_keyed(CmixPlan** out, const PlanKey& key, const int64_t* lnn, int64_t lnnsize, int64_t lnn_min,
                               const double* G, int64_t nr, int64_t nmax, int64_t lmax) {
    int dev = 0;
    SFB_TRY(current_device(&dev));
    PlanCache& c = g_plan_cache[dev];
    if (c.p && c.key == key) {
        *out = c.p;
        return 0;
    }
    if (c.p) {
        cmix_plan_destroy(c.p);
        c.p = nullptr;
    }
    SFB_TRY(cmix_plan_create(&c.p, lnn, lnnsize, lnn_min, G, nr, nmax, lmax));
    c.key = key;
    *out = c.p;
    return 0;
}
static int get_cmix_plan(CmixPlan** out, const int64_t* lnn, int64_t lnnsize, int64_t lnn_min, const double* G,
                         int64_t nr, int64_t nmax, int64_t lmax) {
    PlanKey key;
    SFB_TRY(make_plan_key(&key, lnn, lnnsize, lnn_min, G, nr, nmax, lmax));
    return get_cmix_plan_keyed(out, key, lnn, lnnsize, lnn_min, G, nr, nmax, lmax);
}

// W_lm(r) of one or two windows on the current device (planar), shared by the single-device end-to-end entry points
static int windows_to_alm(Workspace& ws, const double* win1, const double* win2, int64_t nr, int64_t npix_in,
                          int64_t ld_win, int64_t nside, int64_t LMAX, bool* same) {
    int64_t nside_in = 0;
    SFB_TRY(npix2nside(npix_in, &nside_in));
    ShtPlan* sp = nullptr;
    SFB_TRY(get_sht_plan(&sp, nside_in, nside, LMAX, nr));
    const size_t nalm = sp->lmsize * 2 * sp->nrp;
    SFB_TRY(upload_win(win1, nr, npix_in, ld_win, ws.win, ws.main));
    SFB_TRY(ws.alm1.alloc(nalm));
    SFB_TRY(sht_map2alm(sp, ws.win.p, nr, 3, ws.alm1.p, ws.main));
    g_times[0] = sp->t_total;
    g_times[6] = sp->launches;
    *same = (win2 == nullptr || win2 == win1);
    if (!*same) {
        SFB_TRY(upload_win(win2, nr, npix_in, ld_win, ws.win, ws.main));
        SFB_TRY(ws.alm2.alloc(nalm));
        SFB_TRY(sht_map2alm(sp, ws.win.p, nr, 3, ws.alm2.p, ws.main));
        g_times[0] += sp->t_total;
        g_times[6] += sp->launches;
    }
    return 0;
}

// =================================================================================================================
// Multi-device host path (sfb_set_devices(n), n > 1): ONE process drives n GPUs, one worker thread per device for the
// duration of the call (nothing outlives it).  It replaces the reference's only parallel gather, the pmap over rows
// inside _power_win_mix (src/windows.jl:834-861), and the serial shell loop of calc_Wr_lm (:531-535):
//   A  every device uploads a contiguous PIXEL slice of the window (1/n of the bytes over its own PCIe link);
//   B  every device gathers ITS SHELLS of all pixels from the peers' slices over NVLink (peer loads in a kernel),
//      transforms them (stage 1 on nr/n shells) and publishes its W_lm(r) shard;
//   C  every device gathers all W_lm(r) shards (peer loads), forms a full-height COLUMN range of M — a contiguous slab
//      of the caller's column-major matrix — and copies it straight into the caller's buffer over its own PCIe link
//      (sub-slabs pipelined with the compute).  No all-gather of M ever runs.
struct Gang {
    int n = 1;
    std::mutex m;
    std::condition_variable cv;
    int count = 0;
    unsigned gen = 0;
    std::atomic<int> failed{0};
    int rc[kMaxDev] = {0, 0, 0, 0, 0, 0, 0, 0};
    std::string err[kMaxDev];
    void sync() {
        std::unique_lock<std::mutex> lk(m);
        const unsigned g = gen;
        if (++count == n) {
            count = 0;
            ++gen;
            cv.notify_all();
        } else {
            cv.wait(lk, [&] { return gen != g; });
        }
    }
    void fail(int d, int code) {
        rc[d] = code;
        err[d] = g_err;
        failed.store(1);
    }
};

struct PtrTable8 {
    const double* p[kMaxDev];
    long long bound[kMaxDev + 1];
    int stride[kMaxDev];
    int n;
};

// dst[pix][j] (j < cnt, row stride ldd) = window[pix][s_lo + j], pixel `pix` living in the slice of the device that
// uploaded it: slice g holds pixels [bound[g], bound[g+1]) as [pixel - bound[g]][nr]
__global__ void win_shard_gather_kernel(PtrTable8 t, long long npix, int nr, int s_lo, int cnt, int ldd,
                                        double* __restrict__ dst) {
    const long long total = npix * cnt;
    for (long long x = blockIdx.x * (long long)blockDim.x + threadIdx.x; x < total; x += (long long)gridDim.x * blockDim.x) {
        const long long pix = x / cnt;
        const int j = (int)(x - pix * cnt);
        int g = 0;
        while (g + 1 < t.n && pix >= t.bound[g + 1]) ++g;
        dst[pix * ldd + j] = t.p[g][(pix - t.bound[g]) * nr + s_lo + j];
    }
}

// full planar alm [row = lm*2+comp][nrp] <- shards [row][stride[g]], shard g holding shells [bound[g], bound[g+1])
__global__ void alm_shard_gather_kernel(PtrTable8 t, long long rows, int nr, int nrp, double* __restrict__ dst) {
    const long long total = rows * nrp;
    for (long long x = blockIdx.x * (long long)blockDim.x + threadIdx.x; x < total; x += (long long)gridDim.x * blockDim.x) {
        const long long row = x / nrp;
        const int r = (int)(x - row * nrp);
        double v = 0.0;
        if (r < nr) {
            int g = 0;
            while (g + 1 < t.n && r >= t.bound[g + 1]) ++g;
            v = t.p[g][row * t.stride[g] + (r - t.bound[g])];
        }
        dst[x] = v;
    }
}

struct MdJob {
    // inputs of the call
    const double* win[2] = {nullptr, nullptr};
    int nwin = 1;
    int64_t nr = 0, npix_in = 0, ld_win = 0, nside_in = 0, nside = 0, LMAX = 0, niter = 3;
    // stage 2+3 (absent for calc_Wr_lm)
    bool want_cmix = false;
    PlanKey key;
    const double* G = nullptr;
    int64_t nmax = 0, lmax = 0, lnnsize = 0, lnn_min = 1;
    const int64_t* lnn = nullptr;
    int div2Lp1 = 0, interchange = 0;
    double* M_out = nullptr;
    // binned output N = w̃ M v (every device bins the column slab of M it forms and returns its columns of N)
    const BinTables* bin = nullptr;
    double* N_out = nullptr;
    int64_t J[kMaxDev + 1] = {0};                          // output-column bounds per device
    int64_t bcol0[kMaxDev] = {0}, bcol1[kMaxDev] = {0};   // M columns each device needs for them (may overlap)
    double t_bin[kMaxDev] = {0};
    // calc_Wr_lm output (device 0 gathers and converts)
    double* wr_out = nullptr;
    int layout = 0;
    // sharding
    int ndev = 1;
    int64_t shell[kMaxDev + 1] = {0}, pix[kMaxDev + 1] = {0}, col[kMaxDev + 1] = {0};
    // published by the workers
    Workspace* ws[kMaxDev] = {nullptr};
    int nrp_shard[kMaxDev] = {0};
    CmixPlan* cplan[kMaxDev] = {nullptr};
    double t_stage1[kMaxDev] = {0};
    int launches1[kMaxDev] = {0};
    Gang gang;
};

static int md_phase_upload(MdJob& J, int d) {
    SFB_CUDA_OK(cudaSetDevice(d));
    Workspace* ws = nullptr;
    SFB_TRY(get_ws(&ws));
    J.ws[d] = ws;
    if (J.want_cmix) SFB_TRY(get_cmix_plan_keyed(&J.cplan[d], J.key, J.lnn, J.lnnsize, J.lnn_min, J.G, J.nr, J.nmax, J.lmax));
    const int64_t p0 = J.pix[d], p1 = J.pix[d + 1];
    for (int w = 0; w < J.nwin; ++w) {
        SFB_TRY(ws->slice[w].alloc((size_t)std::max<int64_t>(1, p1 - p0) * J.nr));
        if (p1 > p0) SFB_TRY(h2d_rows(ws->slice[w].p, J.win[w] + p0 * J.ld_win, J.nr, p1 - p0, J.ld_win, ws->main));
        SFB_CUDA_OK(cudaEventRecord(ws->ev_slice[w], ws->main));
    }
    return 0;
}

static int md_phase_stage1(MdJob& J, int d) {
    Workspace& ws = *J.ws[d];
    const int64_t s0 = J.shell[d], s1 = J.shell[d + 1], cnt = s1 - s0;
    J.nrp_shard[d] = (int)round_up(std::max<int64_t>(cnt, 1), 8);
    if (cnt <= 0) {
        for (int w = 0; w < J.nwin; ++w) SFB_CUDA_OK(cudaEventRecord(ws.ev_alm[w], ws.main));
        return 0;
    }
    ShtPlan* sp = nullptr;
    SFB_TRY(get_sht_plan(&sp, J.nside_in, J.nside, J.LMAX, cnt));
    SFB_TRY(ws.shard.alloc((size_t)J.npix_in * cnt));
    PtrTable8 t;
    t.n = J.ndev;
    for (int w = 0; w < J.nwin; ++w) {
        for (int g = 0; g < J.ndev; ++g) {
            t.p[g] = J.ws[g]->slice[w].p;
            t.bound[g] = J.pix[g];
            t.stride[g] = (int)J.nr;
            SFB_CUDA_OK(cudaStreamWaitEvent(ws.main, J.ws[g]->ev_slice[w], 0));
        }
        t.bound[J.ndev] = J.pix[J.ndev];
        win_shard_gather_kernel<<<2048, 256, 0, ws.main>>>(t, J.npix_in, (int)J.nr, (int)s0, (int)cnt, (int)cnt, ws.shard.p);
        SFB_CUDA_OK(cudaGetLastError());
        SFB_TRY(ws.almshard[w].alloc(sp->lmsize * 2 * sp->nrp));
        SFB_TRY(sht_map2alm(sp, ws.shard.p, cnt, (int)J.niter, ws.almshard[w].p, ws.main));   // synchronous
        J.t_stage1[d] += sp->t_total;
        J.launches1[d] += sp->launches + 1;
        SFB_CUDA_OK(cudaEventRecord(ws.ev_alm[w], ws.main));
    }
    return 0;
}

// full planar W_lm(r) of window w on device d (gathered from every device's shard)
static int md_gather_alm(MdJob& J, int d, int w, DevBuf<double>& dst, int nrp) {
    Workspace& ws = *J.ws[d];
    const size_t lmsize = (size_t)(J.LMAX + 1) * (J.LMAX + 2) / 2;
    SFB_TRY(dst.alloc(lmsize * 2 * nrp));
    PtrTable8 t;
    t.n = 0;
    for (int g = 0; g < J.ndev; ++g) {
        if (J.shell[g + 1] <= J.shell[g]) continue;
        t.p[t.n] = J.ws[g]->almshard[w].p;
        t.bound[t.n] = J.shell[g];
        t.stride[t.n] = J.nrp_shard[g];
        ++t.n;
        SFB_CUDA_OK(cudaStreamWaitEvent(ws.main, J.ws[g]->ev_alm[w], 0));
    }
    t.bound[t.n] = J.nr;
    alm_shard_gather_kernel<<<1024, 256, 0, ws.main>>>(t, (long long)lmsize * 2, (int)J.nr, nrp, dst.p);
    SFB_CUDA_OK(cudaGetLastError());
    return 0;
}

static int md_phase_cmix(MdJob& J, int d) {
    Workspace& ws = *J.ws[d];
    if (J.wr_out) {   // calc_Wr_lm: device 0 assembles, converts to ComplexF64 in the requested column order, copies out
        if (d != 0) return 0;
        const int nrp = (int)round_up(J.nr, 8);
        SFB_TRY(md_gather_alm(J, d, 0, ws.alm1, nrp));
        const size_t lmsize = (size_t)(J.LMAX + 1) * (J.LMAX + 2) / 2;
        SFB_TRY(ws.M.alloc(lmsize * 2 * J.nr));
        SFB_TRY(alm_planar_to_complex(ws.alm1.p, (int)J.LMAX, (int)J.nr, nrp, J.layout, ws.M.p, ws.main));
        SFB_CUDA_OK(cudaMemcpyAsync(J.wr_out, ws.M.p, lmsize * 2 * J.nr * sizeof(double), cudaMemcpyDeviceToHost, ws.main));
        SFB_CUDA_OK(cudaStreamSynchronize(ws.main));
        return 0;
    }
    CmixPlan* p = J.cplan[d];
    SFB_TRY(md_gather_alm(J, d, 0, ws.alm1, p->nrp));
    if (J.nwin == 2) SFB_TRY(md_gather_alm(J, d, 1, ws.alm2, p->nrp));
    const double* a1 = ws.alm1.p;
    const double* a2 = J.nwin == 2 ? ws.alm2.p : a1;
    if (J.bin) {   // N[:, J0:J1) = w̃ · M[:, c0:c1) · v[c0:c1, J0:J1): the slab of M never leaves the device
        const int64_t J0 = J.J[d], J1 = J.J[d + 1], c0 = J.bcol0[d], c1 = J.bcol1[d], n = p->nout, L1 = J.bin->LNN1;
        if (J1 <= J0) return 0;
        BinDev bd;
        SFB_TRY(bin_tables_upload(*J.bin, bd, ws.main));
        SFB_TRY(ws.slab[0].alloc((size_t)std::max<int64_t>(1, c1 - c0) * n));
        if (c1 > c0)
            SFB_TRY(cmix_run(p, a1, a2, J.div2Lp1, J.interchange, 0, n, c0, c1, ws.slab[0].p, n, ws.main));   // synchronous
        SFB_TRY(ws.slab[1].alloc((size_t)L1 * (J1 - J0)));
        cudaEvent_t e0, e1;
        SFB_CUDA_OK(cudaEventCreate(&e0));
        SFB_CUDA_OK(cudaEventCreate(&e1));
        SFB_CUDA_OK(cudaEventRecord(e0, ws.main));
        SFB_TRY(binned_product_range(ws.slab[0].p, c0, *J.bin, bd, J0, J1, ws.slab[1].p, L1, ws.main));
        SFB_CUDA_OK(cudaEventRecord(e1, ws.main));
        SFB_TRY(ws.flag.alloc(1));
        SFB_CUDA_OK(cudaMemsetAsync(ws.flag.p, 0, sizeof(int), ws.main));
        finite_check_kernel<<<256, 256, 0, ws.main>>>(ws.slab[1].p, (size_t)L1 * (J1 - J0), ws.flag.p);
        SFB_CUDA_OK(cudaMemcpyAsync(J.N_out + J0 * L1, ws.slab[1].p, (size_t)L1 * (J1 - J0) * sizeof(double),
                                    cudaMemcpyDeviceToHost, ws.main));
        int h = 0;
        SFB_CUDA_OK(cudaMemcpyAsync(&h, ws.flag.p, sizeof(int), cudaMemcpyDeviceToHost, ws.main));
        SFB_CUDA_OK(cudaStreamSynchronize(ws.main));
        float t = 0;
        cudaEventElapsedTime(&t, e0, e1);
        J.t_bin[d] = t;
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
        if (h) {   // @assert all(isfinite.(mix))  src/windows.jl:1013
            set_error("AssertionError: all(isfinite.(mix))");
            return 4;
        }
        return 0;
    }
    SFB_TRY(cmix_cols_to_host(p, ws, a1, a2, J.div2Lp1, J.interchange, J.col[d], J.col[d + 1], 4, J.M_out));
    return 0;
}

static void md_worker(MdJob* Jp, int d) {
    MdJob& J = *Jp;
    int (*phases[3])(MdJob&, int) = {md_phase_upload, md_phase_stage1, md_phase_cmix};
    static const char* names[3] = {"md: plans + H2D enqueue", "md: shell gather + stage 1", "md: alm gather + stage 2/3 + D2H"};
    const bool trace = (d == 0) && getenv("SFB_TRACE") != nullptr;
    auto t0 = std::chrono::steady_clock::now();
    for (int ph = 0; ph < 3; ++ph) {
        if (!J.gang.failed.load()) {
            const int rc = phases[ph](J, d);
            if (rc) J.gang.fail(d, rc);
        }
        J.gang.sync();   // every worker arrives at every barrier, failed or not
        if (trace) {
            const auto t1 = std::chrono::steady_clock::now();
            fprintf(stderr, "[sfb] %-36s %8.3f ms (device 0 + barrier)\n", names[ph],
                    std::chrono::duration<double, std::milli>(t1 - t0).count());
            t0 = t1;
        }
    }
    // nobody may reuse a slice / alm shard while a peer still reads it: drain this device before the call returns
    if (J.ws[d]) {
        cudaSetDevice(d);
        cudaStreamSynchronize(J.ws[d]->main);
        cudaStreamSynchronize(J.ws[d]->copy);
    }
}

static int md_run(MdJob& J) {
    const int n = J.ndev;
    J.gang.n = n;
    // shells and pixels: equal contiguous ranges
    for (int g = 0; g <= n; ++g) {
        J.shell[g] = std::min<int64_t>(J.nr, ceil_div(J.nr, n) * g);
        J.pix[g] = J.npix_in * g / n;
    }
    int dev0 = 0;
    SFB_CUDA_OK(cudaGetDevice(&dev0));
    std::vector<std::thread> th;
    for (int d = 1; d < n; ++d) th.emplace_back(md_worker, &J, d);
    md_worker(&J, 0);
    for (auto& t : th) t.join();
    cudaSetDevice(dev0);
    for (int d = 0; d < n; ++d)
        if (J.gang.rc[d]) {
            set_error("device " + std::to_string(d) + ": " + J.gang.err[d]);
            return J.gang.rc[d];
        }
    // timings: the slowest device per stage
    g_times[0] = 0;
    g_times[6] = 0;
    for (int d = 0; d < n; ++d) {
        g_times[0] = std::max(g_times[0], J.t_stage1[d]);
        g_times[6] = std::max(g_times[6], (double)J.launches1[d]);
    }
    if (J.want_cmix) {
        double t[6] = {0, 0, 0, 0, 0, 0}, launches = 0;
        for (int d = 0; d < n; ++d) {
            const CmixPlan* p = J.cplan[d];
            if (!p || (J.bin ? J.bcol1[d] <= J.bcol0[d] : J.col[d + 1] <= J.col[d])) continue;
            t[0] = std::max<double>(t[0], p->t_wl);
            t[1] = std::max<double>(t[1], p->t_fill);
            t[2] = std::max<double>(t[2], p->t_what);
            t[3] = std::max<double>(t[3], p->t_block);
            t[4] += p->flops_executed;
            t[5] = std::max<double>(t[5], p->t_k3);
            launches = std::max<double>(launches, p->launches);
        }
        g_times[1] = t[0];
        g_times[2] = t[1];
        g_times[3] = t[2];
        g_times[4] = t[3];
        g_times[5] = t[4];
        g_times[8] = t[5];
        g_times[6] += launches;
        g_times[7] = 0;
        for (int d = 0; d < n; ++d) g_times[7] = std::max(g_times[7], J.t_bin[d]);
    }
    return 0;
}

// number of devices a call of this size is spread over
static int md_devices_for(int64_t nr) { return (int)std::max<int64_t>(1, std::min<int64_t>(g_ndev, nr)); }
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

int32_t sfb_set_devices(int32_t n) {
    std::lock_guard<std::mutex> lk(g_mutex);
    int have = 0;
    SFB_CUDA_OK(cudaGetDeviceCount(&have));
    SFB_REQUIRE(n >= 1 && n <= kMaxDev, "sfb_set_devices: n must be in 1..8");
    if (n > have) {
        set_error("sfb_set_devices: " + std::to_string(n) + " devices requested, " + std::to_string(have) + " visible");
        return 2;
    }
    int dev0 = 0;
    SFB_CUDA_OK(cudaGetDevice(&dev0));
    // the workers read each other's buffers in kernels: every pair needs peer access (NVLink / NVSwitch)
    for (int i = 0; i < n && n > 1; ++i) {
        SFB_CUDA_OK(cudaSetDevice(i));
        for (int j = 0; j < n; ++j) {
            if (i == j) continue;
            int can = 0;
            SFB_CUDA_OK(cudaDeviceCanAccessPeer(&can, i, j));
            if (!can) {
                cudaSetDevice(dev0);
                set_error("sfb_set_devices: device " + std::to_string(i) + " cannot access device " + std::to_string(j) +
                          " (peer access is required)");
                return 2;
            }
            const cudaError_t e = cudaDeviceEnablePeerAccess(j, 0);
            if (e == cudaErrorPeerAccessAlreadyEnabled) {
                cudaGetLastError();
            } else if (e != cudaSuccess) {
                cudaSetDevice(dev0);
                set_error(std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e));
                return 1;
            }
        }
    }
    SFB_CUDA_OK(cudaSetDevice(n > 1 ? 0 : dev0));
    g_ndev = n;
    return 0;
}
int32_t sfb_get_devices(int32_t* n) {
    SFB_REQUIRE(n, "null pointer");
    *n = g_ndev;
    return 0;
}

int32_t sfb_host_alloc(void** ptr, int64_t bytes) {
    SFB_REQUIRE(ptr && bytes > 0, "sfb_host_alloc: bad arguments");
    SFB_CUDA_OK(cudaHostAlloc(ptr, (size_t)bytes, cudaHostAllocPortable));
    return 0;
}
int32_t sfb_host_free(void* ptr) {
    if (ptr) SFB_CUDA_OK(cudaFreeHost(ptr));
    return 0;
}
int32_t sfb_host_register(void* ptr, int64_t bytes) {
    SFB_REQUIRE(ptr && bytes > 0, "sfb_host_register: bad arguments");
    SFB_CUDA_OK(cudaHostRegister(ptr, (size_t)bytes, cudaHostRegisterPortable));
    return 0;
}
int32_t sfb_host_unregister(void* ptr) {
    SFB_REQUIRE(ptr, "sfb_host_unregister: null pointer");
    SFB_CUDA_OK(cudaHostUnregister(ptr));
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
    for (int i = 0; i < n && i < 9; ++i) out[i] = g_times[i];
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
    SFB_REQUIRE(ld_win >= nr, "ld_win < nr");
    SFB_REQUIRE(niter >= 0, "niter < 0");
    if (md_devices_for(nr) > 1) {   // shells sharded over the devices of sfb_set_devices
        MdJob J;
        J.win[0] = win;
        J.nr = nr, J.npix_in = npix_in, J.ld_win = ld_win, J.nside_in = nside_in, J.nside = nside_out, J.LMAX = lmax;
        J.niter = niter;
        J.wr_out = out, J.layout = layout;
        J.ndev = md_devices_for(nr);
        return md_run(J);
    }
    ShtPlan* sp = nullptr;
    SFB_TRY(get_sht_plan(&sp, nside_in, nside_out, lmax, nr));
    Workspace* ws = nullptr;
    SFB_TRY(get_ws(&ws));
    DevBuf<double> d_out;
    SFB_TRY(upload_win(win, nr, npix_in, ld_win, ws->win, ws->main));
    SFB_TRY(ws->alm1.alloc(sp->lmsize * 2 * sp->nrp));
    SFB_TRY(d_out.alloc(sp->lmsize * 2 * nr));
    SFB_TRY(sht_map2alm(sp, ws->win.p, nr, (int)niter, ws->alm1.p, ws->main));
    g_times[0] = sp->t_total;
    g_times[6] = sp->launches;
    SFB_TRY(sht_alm_to_complex(sp, ws->alm1.p, layout, d_out.p, ws->main));
    SFB_CUDA_OK(cudaMemcpyAsync(out, d_out.p, sp->lmsize * 2 * nr * sizeof(double), cudaMemcpyDeviceToHost, ws->main));
    SFB_CUDA_OK(cudaStreamSynchronize(ws->main));
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
    if (md_devices_for(nr) > 1) {   // sfb_set_devices(n): shells, then columns of M, sharded over n GPUs
        MdJob J;
        SFB_TRY(make_plan_key(&J.key, lnn, lnnsize, lnn_min, G, nr, nmax, lmax));
        SFB_REQUIRE(lnn_min >= 1 && lnn_min <= lnnsize, "lnn_min out of range");
        SFB_REQUIRE(ld_win >= nr, "ld_win < nr");
        J.win[0] = win1;
        J.nwin = (win2 == nullptr || win2 == win1) ? 1 : 2;
        J.win[1] = win2;
        J.nr = nr, J.npix_in = npix_in, J.ld_win = ld_win, J.nside = nside, J.LMAX = 2 * lmax;
        SFB_TRY(npix2nside(npix_in, &J.nside_in));
        J.want_cmix = true;
        J.G = G, J.nmax = nmax, J.lmax = lmax, J.lnn = lnn, J.lnnsize = lnnsize, J.lnn_min = lnn_min;
        J.div2Lp1 = div2Lp1, J.interchange = interchange_NN, J.M_out = M_out;
        J.ndev = md_devices_for(nr);
        // column ranges (L-block aligned, balanced on the full-height cost): needs the mode tables only
        SFB_TRY(cmix_col_bounds_from_lnn(lnn, lnnsize, lnn_min, nr, nmax, lmax, J.ndev, J.col));
        SFB_TRY(md_run(J));
        tr.mark("multi-device power_win_mix");
        return 0;
    }
    PlanGuard pg;
    SFB_TRY(get_cmix_plan(&pg.p, lnn, lnnsize, lnn_min, G, nr, nmax, lmax));
    tr.mark("plan");
    Workspace* ws = nullptr;
    SFB_TRY(get_ws(&ws));
    bool same = true;
    SFB_TRY(windows_to_alm(*ws, win1, win2, nr, npix_in, ld_win, nside, 2 * lmax, &same));
    tr.mark("H2D + stage 1");
    SFB_TRY(cmix_cols_to_host(pg.p, *ws, ws->alm1.p, same ? ws->alm1.p : ws->alm2.p, div2Lp1, interchange_NN, 0,
                              pg.p->nout, 8, M_out, true));
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
    if (md_devices_for(nr) > 1) {   // sfb_set_devices(n): output columns of N, hence column slabs of M, sharded over n GPUs
        MdJob J;
        SFB_TRY(make_plan_key(&J.key, lnn, lnnsize, 1, G, nr, nmax, lmax));
        SFB_REQUIRE(ld_win >= nr, "ld_win < nr");
        BinTables bt;
        SFB_TRY(bin_tables_build(&bt, lnnsize, wt_colptr, wt_rowval, wt_nzval, LNN1, v_colptr, v_rowval, v_nzval, LNN2));
        J.win[0] = win1;      // like the reference, W2r_lm is computed from win1 as well (src/windows.jl:1005-1006)
        J.nr = nr, J.npix_in = npix_in, J.ld_win = ld_win, J.nside = nside, J.LMAX = 2 * lmax;
        SFB_TRY(npix2nside(npix_in, &J.nside_in));
        J.want_cmix = true;
        J.G = G, J.nmax = nmax, J.lmax = lmax, J.lnn = lnn, J.lnnsize = lnnsize, J.lnn_min = 1;
        J.div2Lp1 = div2Lp1, J.interchange = interchange_NN;
        J.bin = &bt, J.N_out = N_out;
        J.ndev = md_devices_for(nr);
        // output columns in ndev contiguous ranges of equal v-nonzero count (a proxy for the M columns behind them)
        const int64_t total = bt.has_v ? (int64_t)bt.vptr[LNN2] : LNN2;
        J.J[0] = 0;
        for (int g = 1; g <= J.ndev; ++g) {
            int64_t j = J.J[g - 1];
            const int64_t target = total * g / J.ndev;
            while (j < LNN2 && (bt.has_v ? (int64_t)bt.vptr[j] : j) < target) ++j;
            J.J[g] = (g == J.ndev) ? LNN2 : j;
        }
        for (int g = 0; g < J.ndev; ++g) bin_needed_cols(bt, J.J[g], J.J[g + 1], &J.bcol0[g], &J.bcol1[g]);
        SFB_TRY(md_run(J));
        tr.mark("multi-device binned power_win_mix");
        return 0;
    }
    PlanGuard pg;
    SFB_TRY(get_cmix_plan(&pg.p, lnn, lnnsize, 1, G, nr, nmax, lmax));
    tr.mark("binned: plan");
    Workspace* ws = nullptr;
    SFB_TRY(get_ws(&ws));
    DevBuf<double>&a1 = ws->alm1, &dM = ws->M;
    bool same = true;
    // like the reference, W2r_lm is computed from win1 as well (src/windows.jl:1005-1006)
    SFB_TRY(windows_to_alm(*ws, win1, nullptr, nr, npix_in, ld_win, nside, 2 * lmax, &same));
    tr.mark("binned: H2D + stage 1");
    const int64_t n = pg.p->nout;
    SFB_TRY(dM.alloc((size_t)n * n));
    SFB_TRY(cmix_run(pg.p, a1.p, a1.p, div2Lp1, interchange_NN, 0, n, 0, n, dM.p, n, 0, nullptr, 0, false, mirror_enabled()));
    record_cmix_times(pg.p);
    tr.mark("binned: stage 2+3");
    float t_bin = 0;
    SFB_TRY(binned_product_to_host(dM.p, n, wt_colptr, wt_rowval, wt_nzval, LNN1, v_colptr, v_rowval, v_nzval, LNN2,
                                   N_out, &t_bin, d2h_auto));
    g_times[7] = t_bin;
    tr.mark("binned: w~ M v + D2H");
    return 0;
}

// ---- on-device deconvolution ----
int32_t sfb_solve(const double* N, int64_t n, const double* B, int64_t nrhs, double* X_out) {
    std::lock_guard<std::mutex> lk(g_mutex);
    SFB_REQUIRE(N && B && X_out && n >= 1 && nrhs >= 1, "sfb_solve: bad arguments");
    Workspace* ws = nullptr;
    SFB_TRY(get_ws(&ws));
    DevBuf<double> A;
    SFB_TRY(A.alloc((size_t)n * (n + nrhs)));
    SFB_CUDA_OK(cudaMemcpyAsync(A.p, N, (size_t)n * n * sizeof(double), cudaMemcpyHostToDevice, ws->main));
    SFB_CUDA_OK(cudaMemcpyAsync(A.p + (size_t)n * n, B, (size_t)n * nrhs * sizeof(double), cudaMemcpyHostToDevice, ws->main));
    int info = 0;
    SFB_TRY(lu_solve_inplace(A.p, n, n, nrhs, &info, ws->main));
    SFB_TRY(check_finite(A.p + (size_t)n * n, (size_t)n * nrhs, "solution"));
    SFB_CUDA_OK(cudaMemcpy(X_out, A.p + (size_t)n * n, (size_t)n * nrhs * sizeof(double), cudaMemcpyDeviceToHost));
    return 0;
}

int32_t sfb_power_win_mix_binned_solve(const double* win1, int64_t nr, int64_t npix_in, int64_t ld_win, int64_t nside,
                                       const double* G, int64_t nmax, int64_t lmax, const int64_t* lnn, int64_t lnnsize,
                                       const int64_t* wt_colptr, const int64_t* wt_rowval, const double* wt_nzval,
                                       int64_t LNN1, const int64_t* v_colptr, const int64_t* v_rowval,
                                       const double* v_nzval, int64_t LNN2, int32_t div2Lp1, int32_t interchange_NN,
                                       const double* B, int64_t nrhs, double* X_out, double* N_out) {
    std::lock_guard<std::mutex> lk(g_mutex);
    SFB_REQUIRE(win1 && B && X_out && nrhs >= 1, "null pointer");
    SFB_REQUIRE(LNN1 == LNN2, "the binned coupling matrix must be square to be inverted");
    Trace tr;
    PlanGuard pg;
    SFB_TRY(get_cmix_plan(&pg.p, lnn, lnnsize, 1, G, nr, nmax, lmax));
    Workspace* ws = nullptr;
    SFB_TRY(get_ws(&ws));
    bool same = true;
    SFB_TRY(windows_to_alm(*ws, win1, nullptr, nr, npix_in, ld_win, nside, 2 * lmax, &same));
    const int64_t n = pg.p->nout, L = LNN1;
    SFB_TRY(ws->M.alloc((size_t)n * n));
    SFB_TRY(cmix_run(pg.p, ws->alm1.p, ws->alm1.p, div2Lp1, interchange_NN, 0, n, 0, n, ws->M.p, n, 0, nullptr, 0, false,
                     mirror_enabled()));
    record_cmix_times(pg.p);
    tr.mark("solve: stage 1-3");
    DevBuf<double> A;                         // [N | B], L x (L + nrhs)
    SFB_TRY(A.alloc((size_t)L * (L + nrhs)));
    float t_bin = 0;
    SFB_TRY(binned_product_device(ws->M.p, n, wt_colptr, wt_rowval, wt_nzval, LNN1, v_colptr, v_rowval, v_nzval, LNN2, A.p, L,
                                  &t_bin));
    g_times[7] = t_bin;
    if (N_out) SFB_CUDA_OK(cudaMemcpy(N_out, A.p, (size_t)L * L * sizeof(double), cudaMemcpyDeviceToHost));
    SFB_CUDA_OK(cudaMemcpy(A.p + (size_t)L * L, B, (size_t)L * nrhs * sizeof(double), cudaMemcpyHostToDevice));
    tr.mark("solve: N = w~ M v");
    int info = 0;
    SFB_TRY(lu_solve_inplace(A.p, L, L, nrhs, &info, 0));
    SFB_TRY(check_finite(A.p + (size_t)L * L, (size_t)L * nrhs, "solution"));
    SFB_CUDA_OK(cudaMemcpy(X_out, A.p + (size_t)L * L, (size_t)L * nrhs * sizeof(double), cudaMemcpyDeviceToHost));
    tr.mark("solve: LU + substitution + D2H");
    return 0;
}

int32_t sfb_win_lnn(const double* win, int64_t nr, int64_t npix_in, int64_t ld_win, int64_t nside, const double* G,
                    int64_t nmax, int64_t lmax, const int64_t* lnn, int64_t lnnsize, double* Wlnn_out) {
    std::lock_guard<std::mutex> lk(g_mutex);
    SFB_REQUIRE(win && Wlnn_out, "null pointer");
    PlanGuard pg;
    SFB_TRY(get_cmix_plan(&pg.p, lnn, lnnsize, 1, G, nr, nmax, lmax));
    Workspace* ws = nullptr;
    SFB_TRY(get_ws(&ws));
    DevBuf<double>& a1 = ws->alm1;
    bool same = true;
    SFB_TRY(windows_to_alm(*ws, win, nullptr, nr, npix_in, ld_win, nside, 2 * lmax, &same));
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
int32_t sfb_calc_wmix(const double* win, int64_t nr, int64_t npix_in, int64_t ld_win, int64_t nside, const double* G,
                      int64_t nmax, int64_t lmax, const int64_t* nmax_l, const int64_t* lmax_n, int32_t neg_m,
                      double* wmix_out) {
    std::lock_guard<std::mutex> lk(g_mutex);
    SFB_REQUIRE(win && G && nmax_l && lmax_n && wmix_out, "null pointer");
    Workspace* ws = nullptr;
    SFB_TRY(get_ws(&ws));
    bool same = true;
    SFB_TRY(windows_to_alm(*ws, win, nullptr, nr, npix_in, ld_win, nside, 2 * lmax, &same));
    int64_t nlmsize = 0;
    for (int64_t n = 0; n < nmax; ++n) nlmsize += (lmax_n[n] + 1) * (lmax_n[n] + 2) / 2;
    SFB_REQUIRE(nlmsize >= 1, "calc_wmix: empty mode set");
    DevBuf<double> d_out;
    SFB_TRY(d_out.alloc((size_t)nlmsize * nlmsize * 2));
    SFB_TRY(wmix_run(ws->alm1.p, (int)round_up(nr, 8), G, nr, nmax, lmax, nmax_l, lmax_n, neg_m, d_out.p, nullptr,
                     ws->main));
    g_times[6] += 1;
    SFB_TRY(check_finite(d_out.p, (size_t)nlmsize * nlmsize * 2, "wmix"));
    SFB_CUDA_OK(cudaMemcpy(wmix_out, d_out.p, (size_t)nlmsize * nlmsize * 2 * sizeof(double), cudaMemcpyDeviceToHost));
    return 0;
}

// ---- SFB transforms next to the window path ----
int32_t sfb_field2anlm(const double* f_xyz, int64_t nr, int64_t npix, int64_t ld, const double* T, int64_t nmax,
                       int64_t lmax, const int64_t* nmax_l, const int64_t* lmax_n, double* f_nlm_out) {
    std::lock_guard<std::mutex> lk(g_mutex);
    SFB_REQUIRE(f_xyz && T && f_nlm_out, "null pointer");
    int64_t nside = 0;
    SFB_TRY(npix2nside(npix, &nside));
    ShtPlan* sp = nullptr;
    SFB_TRY(get_sht_plan(&sp, nside, nside, lmax, nr));
    Workspace* ws = nullptr;
    SFB_TRY(get_ws(&ws));
    SFB_TRY(upload_win(f_xyz, nr, npix, ld, ws->win, ws->main));
    SFB_TRY(ws->alm1.alloc(sp->lmsize * 2 * sp->nrp));
    SFB_TRY(sfbt_field2anlm(sp, ws->win.p, nr, T, nmax, lmax, nmax_l, lmax_n, ws->alm1.p, f_nlm_out, ws->main));
    g_times[0] = sp->t_total;
    g_times[6] = sp->launches + 1;
    return 0;
}

int32_t sfb_anlm2field(const double* f_nlm, const double* g, int64_t nr, int64_t nside, int64_t nmax, int64_t lmax,
                       const int64_t* nmax_l, const int64_t* lmax_n, double* f_xyz_out, int64_t ld_out) {
    std::lock_guard<std::mutex> lk(g_mutex);
    SFB_REQUIRE(f_nlm && g && f_xyz_out, "null pointer");
    ShtPlan* sp = nullptr;
    SFB_TRY(get_sht_plan(&sp, nside, nside, lmax, nr));
    Workspace* ws = nullptr;
    SFB_TRY(get_ws(&ws));
    SFB_TRY(ws->alm1.alloc(sp->lmsize * 2 * sp->nrp));
    SFB_TRY(ws->win.alloc((size_t)sp->npix * sp->nrp));
    SFB_TRY(sfbt_anlm2field(sp, f_nlm, g, nmax, lmax, nmax_l, lmax_n, ws->alm1.p, ws->win.p, f_xyz_out, ld_out, ws->main));
    g_times[6] = sp->launches + 1;
    return 0;
}

int32_t sfb_win_rhat_ln(const double* win, int64_t nr, int64_t npix, int64_t ld_win, const double* T, int64_t nmax,
                        int64_t lmax, const int64_t* nmax_l, const int64_t* lmax_n, double* out) {
    std::lock_guard<std::mutex> lk(g_mutex);
    SFB_REQUIRE(win && T && out, "null pointer");
    Workspace* ws = nullptr;
    SFB_TRY(get_ws(&ws));
    SFB_TRY(upload_win(win, nr, npix, ld_win, ws->win, ws->main));
    const size_t nout = (size_t)npix * (lmax + 1) * nmax;
    SFB_TRY(ws->M.alloc(nout));
    SFB_TRY(sfbt_win_rhat_ln(ws->win.p, nr, npix, nr, T, nmax, lmax, nmax_l, lmax_n, ws->M.p, ws->main));
    SFB_CUDA_OK(cudaMemcpyAsync(out, ws->M.p, nout * sizeof(double), cudaMemcpyDeviceToHost, ws->main));
    SFB_CUDA_OK(cudaStreamSynchronize(ws->main));
    g_times[6] = 1;
    return 0;
}

int32_t sfb_cat2amln(const int64_t* pixptr, const int64_t* gidx, int64_t ngal, const double* gw, const int64_t* mode_n,
                     const int64_t* mode_l, int64_t nb, double nbar, const double* win_rhat_ln, int64_t nside,
                     int64_t nmax, int64_t lmax, const int64_t* nmax_l, const int64_t* lmax_n, double* anlm_inout) {
    std::lock_guard<std::mutex> lk(g_mutex);
    SFB_REQUIRE(pixptr && mode_n && mode_l && win_rhat_ln && nmax_l && lmax_n && anlm_inout && nb >= 1, "bad arguments");
    ShtPlan* sp = nullptr;
    SFB_TRY(get_sht_plan(&sp, nside, nside, lmax, nb));
    Workspace* ws = nullptr;
    SFB_TRY(get_ws(&ws));
    const size_t nw = (size_t)sp->npix * (lmax + 1) * nmax;
    SFB_TRY(ws->M.alloc(nw));
    SFB_CUDA_OK(cudaMemcpyAsync(ws->M.p, win_rhat_ln, nw * sizeof(double), cudaMemcpyHostToDevice, ws->main));
    const int64_t nlmsize = sfbt_nlmsize(nmax, lmax_n);
    DevBuf<double> d_anlm;
    SFB_TRY(d_anlm.alloc((size_t)nlmsize * 2));
    SFB_CUDA_OK(cudaMemcpyAsync(d_anlm.p, anlm_inout, (size_t)nlmsize * 2 * sizeof(double), cudaMemcpyHostToDevice, ws->main));
    SFB_TRY(sfbt_cat2amln_batch(sp, pixptr, gidx, ngal, gw, mode_n, mode_l, nb, nbar, ws->M.p, nmax, lmax, nmax_l, lmax_n,
                                d_anlm.p, ws->main));
    SFB_CUDA_OK(cudaMemcpyAsync(anlm_inout, d_anlm.p, (size_t)nlmsize * 2 * sizeof(double), cudaMemcpyDeviceToHost, ws->main));
    SFB_CUDA_OK(cudaStreamSynchronize(ws->main));
    g_times[0] = sp->t_total;
    g_times[6] = sp->launches + 2;
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
    SFB_TRY(upload_win(mask, 1, npix_in, 1, d_mask, 0));
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

// full planar W_lm(r) <- the shell shards of several ranks (all pointers on the current device, e.g. the receive buffer of
// an NCCL all-gather): one placement kernel instead of a per-rank strided copy
int32_t sfb_alm_gather_shards_dev(const double* const* shard_ptrs, const int64_t* shell_bounds, const int64_t* strides,
                                  int32_t nshards, int64_t LMAX, int64_t nr, double* d_alm, void* stream) {
    SFB_REQUIRE(shard_ptrs && shell_bounds && strides && d_alm && nshards >= 1 && nshards <= kMaxDev, "bad arguments");
    PtrTable8 t;
    t.n = 0;
    for (int g = 0; g < nshards; ++g) {
        if (shell_bounds[g + 1] <= shell_bounds[g]) continue;
        SFB_REQUIRE(shard_ptrs[g], "null shard pointer");
        t.p[t.n] = shard_ptrs[g];
        t.bound[t.n] = shell_bounds[g];
        t.stride[t.n] = (int)strides[g];
        ++t.n;
    }
    SFB_REQUIRE(t.n >= 1 && shell_bounds[nshards] == nr, "shell bounds do not cover [0, nr)");
    t.bound[t.n] = nr;
    const long long rows = (LMAX + 1) * (LMAX + 2);   // lmsize * 2
    alm_shard_gather_kernel<<<1024, 256, 0, (cudaStream_t)stream>>>(t, rows, (int)nr, (int)round_up(nr, 8), d_alm);
    SFB_CUDA_OK(cudaGetLastError());
    return 0;
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
int32_t sfb_cmix_mirror_rows_dev(sfb_cmix_plan* plan, const double* d_packed, int64_t col_lo, int64_t col_hi,
                                 int32_t div2Lp1, int32_t interchange_NN, double* d_rows, int64_t ld_rows, void* stream) {
    std::lock_guard<std::mutex> lk(g_mutex);
    auto* p = reinterpret_cast<CmixPlan*>(plan);
    SFB_TRY(cmix_mirror_rows(p, d_packed, col_lo, col_hi, div2Lp1, interchange_NN, d_rows, ld_rows, (cudaStream_t)stream));
    g_times[6] += 1;
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
            // the Ŵ build (banded DMMA GEMM, L2-bound) runs ≈ 3x slower per flop than the block kernel (per-rank timings, cfg4)
            c += 2.0 * ap * p->nrp * p->nrp * b + 2.0 * ap * ap * p->nrp * b * (b + 1) / 2 +
                 3.0 * 2.0 * p->nrp * p->nrp * (std::min(l, L) + 1);
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
