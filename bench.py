#!/usr/bin/env python
"""bench.py — coupling-matrix elements/s of power_win_mix (BASELINE.json metric) on N B200s.

  python bench.py --gpus N --steps K --warmup W            # CUDA arm (one rank per GPU under torchrun for N > 1)
  python bench.py --impl reference --gpus N --steps K ...  # CPU arm: the reference's algorithm on the host cores

A step = one full power_win_mix(win, wmodes, cmodes) over the configuration's window:
stage 1 (W_lm(r) of all shells, map2alm niter=3) + stage 2 (W_{L1}, Ŵ_{lL}) + stage 3 (all lnnsize² elements).
`value`   : elements/s with the window already resident in HBM and M left in HBM (CUDA events, max over ranks).
`e2e`     : the same call through the public host API (`sfb_b200.power_win_mix`, i.e. the C ABI with HOST
            buffers): pinned host window in, full matrix out, H2D/D2H inside the timed region.
Default workload: cfg4 (nside=256, kmax=0.15, 64 shells) — the configuration north_star's target names and the
largest that is quoted for one GPU; cfg5 needs `--config 5`.
"""
from __future__ import annotations

import argparse
import contextlib
import io
import json
import os
import statistics
import subprocess
import sys
import threading
import time
import warnings

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "coupling_matrix_elements_per_sec"
UNIT = "elements/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="4")
    ap.add_argument("--cpu-seconds", type=float, default=20.0, help="CPU work budget of the cpu_baseline sample")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    return ap.parse_args()


def cfg_key(s):
    return int(s) if s.isdigit() else s


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
                power.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_reference_run(wl, seconds, shells=2):
    """The reference's algorithm in the reference's operation order on the host cores (oracle port; Julia is not
    installed so the reference itself cannot run).  Bounded sample: `shells` shells of stage 1 (FFT-based
    map2alm niter=3), all of stage 2, and a strided row subset x all columns of stage 3 sized for ~`seconds`.
    Returns (elements_per_s_whole_job_estimate, details)."""
    from oracle import cref, sht_fft
    from oracle import healpix as ohp
    from oracle import windows as ow
    from sfb_b200.separable import SeparableArray

    a = wl.amodes
    LMAX, nr, n = wl.LMAX, wl.nr, wl.lnnsize
    cores = cref.max_threads()
    win = wl.win
    t0 = time.perf_counter()
    if isinstance(win, SeparableArray):
        dense_rows = np.outer(win.phi[:shells], win.mask)
    else:
        dense_rows = np.ascontiguousarray(win[:shells, :])
    sht = sht_fft.FastSHT(a.nside, LMAX)
    t_plan = time.perf_counter() - t0
    t0 = time.perf_counter()
    maps = ohp.udgrade(dense_rows, a.nside)
    alm_s = sht.map2alm(maps, niter=3)
    t1_sample = time.perf_counter() - t0
    t1_full = t1_sample * nr / shells + t_plan
    # stage 3 needs W_{L1} of all shells: tile the sampled shells' alm (timing of stage 2/3 is data independent)
    reps = -(-nr // shells)
    Wr = ow.optimize_Wr_lm_layout(np.tile(alm_s, (reps, 1))[:nr], LMAX)
    t0 = time.perf_counter()
    Wc = cref.calc_wrl_wrl(Wr, Wr, LMAX)
    t2 = time.perf_counter() - t0
    # calibrate stage 3 on a few rows, then size the sample
    rng = np.random.default_rng(1)
    probe = np.sort(rng.choice(n, size=min(n, 2 * cores), replace=False)) + 1
    t0 = time.perf_counter()
    cref.calc_cmix_rows(wl.cmodes.lnn, probe, wl.G, Wc)
    tp = time.perf_counter() - t0
    nrows = int(max(2 * cores, min(n, seconds / (tp / probe.size))))
    rows = np.unique(np.linspace(0, n - 1, nrows).astype(np.int64)) + 1
    t0 = time.perf_counter()
    cref.calc_cmix_rows(wl.cmodes.lnn, rows, wl.G, Wc)
    t3_sample = time.perf_counter() - t0
    fl_sample = wl.flops_bruteforce(rows - 1)
    fl_full = wl.flops_bruteforce()
    t3_full = t3_sample * fl_full / fl_sample
    total = t1_full + t2 + t3_full
    details = {
        "cores": cores, "kind": "port",
        "sample": (f"stage1: {shells}/{nr} shells (FFT-based map2alm niter=3, numpy) x{nr / shells:.0f}; stage2: full "
                   f"calc_Wrl_Wrl; stage3: {rows.size}/{n} rows x all columns in the reference's per-(element,L1) "
                   f"gemv order (OpenMP dynamic), extrapolated by the brute-force flop model"),
        "stage_s": {"stage1_est": t1_full, "stage2": t2, "stage3_sample": t3_sample, "stage3_est": t3_full},
        "stage3_gflops": fl_sample / t3_sample / 1e9,
        "cpu_seconds_measured": t1_sample + t2 + t3_sample + tp,
        "wall_s_est": total,
    }
    return n * n / total, details


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from sfb_b200 import configs
    wl = configs.Workload(cfg_key(args.config))
    vals, det = [], None
    for it in range(args.warmup + args.steps):
        v, det = cpu_reference_run(wl, args.cpu_seconds / max(1, args.steps))
        if it >= args.warmup:
            vals.append(v)
        if it == 0 and args.warmup > 0 and det["cpu_seconds_measured"] > 60:
            args.warmup = 1  # keep the whole run within a few minutes
    value = float(np.mean(vals))
    n = wl.lnnsize
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * n * n / value, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": wl.describe(),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": det["cores"], "kind": det["kind"],
                         "sample": det["sample"], "stage_s": det["stage_s"], "stage3_gflops": det["stage3_gflops"]},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "whole-job elements/s estimated from a bounded sample (see cpu_baseline.sample); Julia is not "
                "installed, so this is the oracle port of the reference's algorithm, all host threads",
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------ CUDA arm
def run_b200(args):
    import torch
    import torch.distributed as dist

    import sfb_b200 as sfb
    from sfb_b200 import _lib, configs
    from sfb_b200.device import DevicePipeline, PeerMatrix, shard_rows
    from sfb_b200.separable import SeparableArray

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the CUDA arm has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    lib = _lib.load()
    _lib.check(lib.sfb_set_device(local_rank))

    wl = configs.Workload(cfg_key(args.config))
    n = wl.lnnsize
    win = wl.win
    if isinstance(win, SeparableArray):
        raise SystemExit("bench.py times the dense path; cfg3 (separable) is a parity-test case")
    pipe = DevicePipeline(wl.wmodes, wl.cmodes, wl.G)
    # window resident in HBM in Julia memory order: (npix, nr) C-contiguous == (nr, npix) column-major
    d_win = torch.from_numpy(np.ascontiguousarray(win.T)).cuda()
    # exchange variants (default first): cols = column slabs + in-place NCCL send/recv; cols_dma = column slabs +
    # IPC copy-engine pushes; dma / stores = row shards via pitched P2P copies / in-kernel P2P stores;
    # nccl = row slabs + padded all_gather + placement
    # packed (default) = column shards of the l <= L blocks in upper-packed storage + in-place NCCL all-gather of the
    # packed slabs (half the matrix bytes) + local unpack/mirror pass
    # pull (default) = packed slabs left in place in peer-mapped buffers, the expansion kernel pulls every column from
    # its owner over NVLink (exchange fused into the kernel)
    xmode = os.environ.get("SFB_BENCH_EXCHANGE", "pull")
    if xmode in ("packed", "pull"):
        off = pipe.packed_offsets()
        ranges = pipe.packed_shard_ranges(world)
    else:
        ranges = shard_rows(pipe.col_costs if xmode.startswith("cols") else pipe.row_costs, pipe.ell_of_row, world)
    lo, hi = ranges[rank]
    # N = 1: the matrix stays in a device buffer; N > 1: every rank holds the full matrix, rows stored into all
    # copies by the block kernel itself (all-gather fused into the epilogue over NVLink peer memory)
    slab = torch.empty((pipe.nout, hi - lo), dtype=torch.float64, device="cuda") if world == 1 else None
    pm = PeerMatrix(pipe.nout) if (world > 1 and xmode in ("cols_dma", "dma", "stores")) else None
    full_t = torch.empty((pipe.nout, pipe.nout), dtype=torch.float64, device="cuda") if (world > 1 and xmode in ("cols", "packed", "pull")) else None
    packed_t = torch.empty(int(off[-1]), dtype=torch.float64, device="cuda") if (world > 1 and xmode == "packed") else None
    from sfb_b200.device import PeerBuffer, stream_barrier
    pb = PeerBuffer(int(off[-1])) if (world > 1 and xmode == "pull") else None

    fused = xmode in ("dma", "stores")
    from sfb_b200.device import allgather_col_slabs, allgather_packed_slabs
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
    if world > 1 and not fused:
        from sfb_b200.device import gather_row_slabs
        slab = torch.empty((pipe.nout, hi - lo), dtype=torch.float64, device="cuda")

    def step():
        evs[0].record()
        pipe.calc_wr_lm_sharded(d_win)
        evs[1].record()
        if world > 1 and xmode == "pull":
            if hi > lo:
                pipe.power_win_mix_upper_packed(lo, hi, pb.tensor)
            evs[3].record()
            stream_barrier(pipe)
            evs[4].record()
            out = pipe.unpack_mirror_pull(pb, ranges, full_t)
            stream_barrier(pipe)
        elif world > 1 and xmode == "packed":
            if hi > lo:
                pipe.power_win_mix_upper_packed(lo, hi, packed_t)
            evs[3].record()
            allgather_packed_slabs(packed_t, [(int(off[l]), int(off[h])) for l, h in ranges])
            evs[4].record()
            out = pipe.unpack_mirror(packed_t, full_t)
        elif world > 1 and xmode == "cols":
            if hi > lo:
                pipe.power_win_mix_cols(lo, hi, out=full_t[lo:hi])
            allgather_col_slabs(full_t, ranges)
            out = full_t
        elif world > 1 and xmode == "cols_dma":
            _lib.check(lib.sfb_power_win_mix_block_dev(pipe._cmix, pipe.alm.data_ptr(), pipe.alm.data_ptr(), 0, 0, 0,
                                                       pipe.nout, lo, hi, pm.ptr.value + 8 * lo * pipe.nout, pipe.nout,
                                                       pipe._stream()))
            _lib.check(lib.sfb_push_cols_to_peers(pm.ptr, pm.peer_array, len(pm.peer_ptrs), lo, hi, pipe.nout,
                                                  pipe._stream()))
            out = pm.tensor
        elif world > 1 and fused:
            _lib.check(lib.sfb_power_win_mix_dev_peers(pipe._cmix, pipe.alm.data_ptr(), pipe.alm.data_ptr(), 0, 0, lo, hi,
                                                       pm.ptr, pm.peer_array,
                                                       len(pm.peer_ptrs) if xmode == "stores" else 0, pipe.nout,
                                                       pipe._stream()))
            if xmode == "dma":
                _lib.check(lib.sfb_push_rows_to_peers(pm.ptr, pm.peer_array, len(pm.peer_ptrs), lo, hi, pipe.nout,
                                                      pipe.nout, pipe._stream()))
            out = pm.tensor
        else:
            out = pipe.power_win_mix_rows(lo, hi, out=slab)
            if world > 1:
                out = gather_row_slabs(slab, ranges, pipe.nout)
        evs[2].record()
        return out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        full = step()
    barrier()
    tim = _lib.timings()
    stage_ms = {"stage1": [], "wl": [], "what": [], "block": [], "fill": [], "stage1_incl_gather": [], "stage23_incl_exchange": []}
    if world > 1 and xmode in ("packed", "pull"):
        stage_ms.update({"exchange": [], "unpack": []})
    launches_per_step = 0
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        full = step()
        tim = _lib.timings()
        stage_ms["stage1"].append(tim["stage1_ms"])
        stage_ms["wl"].append(tim["wl_ms"])
        stage_ms["what"].append(tim["what_ms"])
        stage_ms["block"].append(tim["block_ms"])
        stage_ms["fill"].append(tim["fill_ms"])
        torch.cuda.synchronize()
        stage_ms["stage1_incl_gather"].append(evs[0].elapsed_time(evs[1]))
        stage_ms["stage23_incl_exchange"].append(evs[1].elapsed_time(evs[2]))
        if world > 1 and xmode in ("packed", "pull"):
            stage_ms["exchange"].append(evs[3].elapsed_time(evs[4]))
            stage_ms["unpack"].append(evs[4].elapsed_time(evs[2]))
        launches_per_step = int(tim["launches"])
    ev1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms = ev0.elapsed_time(ev1) / args.steps
    if world > 1:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = n * n / (ms * 1e-3)
    per_rank = None
    if world > 1:
        mine = {"rank": rank, "rows": [int(lo), int(hi)], **{k: round(statistics.mean(v), 3) for k, v in stage_ms.items()}}
        per_rank = [None] * world
        dist.all_gather_object(per_rank, mine)

    # ---- roofline of the dominant kernel (measured live: CUDA events on the launching stream, inside the lib) ----
    block_ms = statistics.mean(stage_ms["block"])
    fill_ms = statistics.mean(stage_ms["fill"])
    kern_ms = block_ms - fill_ms            # the DMMA block kernel(s) alone; block_ms includes the mirror-fill pass
    stage1_ms = statistics.mean(stage_ms["stage1"])
    dmma = np.zeros(1)
    _lib.check(lib.sfb_probe_dmma_tflops(_lib.ptr(dmma)))
    peaks, traffic = {}, {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))["kernels"]
    except (OSError, KeyError):
        pass
    # algorithmic flops of the block kernel's share of F_alg: the two GEMM terms, for this rank's rows
    nl = np.asarray(wl.amodes.nmax_l, dtype=np.float64)
    ells_mine = np.unique(wl.cmodes.lnn[0, lo:hi])      # l-blocks (row shard) or L-blocks (column shard): same model
    nn = nl[ells_mine][:, None] * nl[None, :]
    f_alg_block = float(np.sum(2 * nn * wl.nr ** 2 + 2 * nn * nn * wl.nr))
    exec_tflops = tim["block_flops"] / (kern_ms * 1e-3) / 1e12
    # ncu dram traffic of the block kernel launches (cfg4, N = 1 capture): only comparable for that workload
    regz_traffic = sum(v["traffic_bytes"] for k, v in traffic.items() if k.startswith("cmix_regz_kernel")) or None
    if world > 1 or str(args.config) != "4":
        regz_traffic = None
    roofline = {
        "kernel": "cmix_regz_kernel (stage 2+3 block GEMMs, FP64 DMMA, all tile classes of one step)", "bound": "tensor",
        "achieved": exec_tflops, "peak": float(dmma[0]), "unit": "TFLOP/s", "frac": exec_tflops / float(dmma[0]),
        "traffic": regz_traffic,
        "traffic_note": "dram__bytes_read.sum + dram__bytes_write.sum of the block-kernel launches of one step from the "
                        "ncu --set full capture in profiles/r01_traffic.json; algorithmic bytes of these launches = the "
                        "directly formed half of M (8·n²/2) + the Ŵ blocks read once",
        "algorithmic_bytes": 8.0 * (hi - lo) * n / (2 if world == 1 else 1),
        "peak_source": "FP64 DMMA probe measured in this run (MEASURED_PEAKS.json has no FP64 figure; nominal B200 "
                       "FP64 tensor peak is 37-40 TFLOP/s)",
        "algorithmic_tflops": f_alg_block / (kern_ms * 1e-3) / 1e12,
        "algorithmic_flops": f_alg_block, "executed_flops": tim["block_flops"], "launch_ms": kern_ms,
        "note": "achieved/frac count the DMMA flops the kernel EXECUTES (n padded to multiples of 8, +9% at cfg4) over its "
                "own launch time.  SURVEY §8d's algorithmic count (2·nn·nr² + 2·nn²·nr per (l,L) block) is larger than the "
                "executed one because only N<=N' tiles are formed and, at N=1, only the L>=l blocks (the mirror-fill pass "
                "writes the other half): algorithmic_tflops can therefore exceed the peak and is not a utilisation figure",
        "mirror_fill": {"ms": fill_ms, "bound": "hbm", "algorithmic_bytes": 8.0 * n * n if world == 1 else 0.0,
                        "achieved_gbs": (8.0 * n * n / (fill_ms * 1e-3) / 1e9) if fill_ms > 0 else None,
                        "hbm_peak_gbs": peaks.get("hbm_gbs"),
                        "frac": (8.0 * n * n / (fill_ms * 1e-3) / 1e9 / peaks["hbm_gbs"])
                        if (fill_ms > 0 and peaks.get("hbm_gbs")) else None,
                        "traffic": (sum(v["traffic_bytes"] for k, v in traffic.items()
                                        if k.startswith("cmix_mirror_fill_kernel")) or None) if world == 1 else None},
        "stage1": {"ms": stage1_ms, "alg_tflops": wl.flops_alg_stage1() / (stage1_ms * 1e-3) / 1e12,
                   "alg_gbs": wl.bytes_alg_stage1() / (stage1_ms * 1e-3) / 1e9,
                   "hbm_peak_gbs": peaks.get("hbm_gbs"), "hbm_frac": (wl.bytes_alg_stage1() / (stage1_ms * 1e-3) / 1e9 /
                                                                      peaks["hbm_gbs"]) if peaks.get("hbm_gbs") else None},
        "stage3_write_gbs": 8.0 * (hi - lo) * n / (block_ms * 1e-3) / 1e9,
        "stage_ms": {k: statistics.mean(v) for k, v in stage_ms.items()},
    }

    # ---- e2e through the public host API: pinned host window in, matrix out ----
    e2e = None
    if not args.no_e2e and world == 1:
        nr, npix = win.shape
        host_win_t = torch.empty((npix, nr), dtype=torch.float64).pin_memory()
        host_win_t.copy_(torch.from_numpy(np.ascontiguousarray(win.T)))
        host_win = host_win_t.numpy().T            # (nr, npix) Fortran-ordered view of pinned memory
        out_t = torch.empty((n, n), dtype=torch.float64).pin_memory()
        out = out_t.numpy().T                      # Fortran-ordered (n, n) view
        del pipe, slab, full
        torch.cuda.empty_cache()
        with warnings.catch_warnings():
            warnings.simplefilter("ignore", RuntimeWarning)
            for _ in range(min(args.warmup, 2)):
                sfb.power_win_mix(host_win, wl.wmodes, wl.cmodes, out=out)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            k = max(1, min(args.steps, 3))
            for _ in range(k):
                M = sfb.power_win_mix(host_win, wl.wmodes, wl.cmodes, out=out)
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t0) / k
        e2e = {"value": n * n / dt, "unit": UNIT, "h2d_bytes_per_step": int(win.nbytes + wl.G.nbytes + wl.cmodes.lnn.nbytes),
               "d2h_bytes_per_step": int(8 * n * n), "ms_per_step": dt * 1e3,
               "api": "sfb_b200.power_win_mix(win, wmodes, cmodes) -> sfb_power_win_mix (C ABI, host pointers)",
               "checksum": float(M[:: max(1, n // 97), :: max(1, n // 89)].sum())}
        # the binned call of cfg4 (BASELINE.json: "binned ClnnBinnedModes output"): N = w̃ M v, Δl = 4
        try:
            wt, vv = sfb.bandpower_binning_weights(wl.cmodes, dl=4)
            bc = sfb.ClnnBinnedModes(wt, vv, wl.cmodes)
            outN_t = torch.empty((vv.shape[1], wt.shape[0]), dtype=torch.float64).pin_memory()
            outN = outN_t.numpy().T
            with warnings.catch_warnings():
                warnings.simplefilter("ignore", RuntimeWarning)
                sfb.power_win_mix(host_win, wt, vv, wl.wmodes, bc, out=outN)
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                for _ in range(k):
                    sfb.power_win_mix(host_win, wt, vv, wl.wmodes, bc, out=outN)
                torch.cuda.synchronize()
                dtb = (time.perf_counter() - t0) / k
            e2e["binned"] = {"ms_per_step": dtb * 1e3, "value": n * n / dtb, "unit": UNIT, "LNN": int(wt.shape[0]),
                             "d2h_bytes_per_step": int(8 * wt.shape[0] * vv.shape[1]),
                             "note": "power_win_mix(win, w̃, v, wmodes, bcmodes) with Δl=4: all lnnsize² elements of M are "
                                     "formed on the device, only N = w̃Mv returns to the host"}
        except Exception as exc:  # noqa: BLE001
            e2e["binned"] = {"error": str(exc)}
    elif world > 1 and not args.no_e2e:
        # N > 1: every rank uploads the shells it transforms from pinned host memory, the sharded step runs, and rank 0
        # reads the assembled matrix back into pinned host memory (what a user of the one-process-per-GPU pipeline
        # does to obtain M on the host).  Falls back to a note if anything in this leg fails.
        e2e = {"value": None, "unit": UNIT, "h2d_bytes_per_step": None, "d2h_bytes_per_step": None,
               "note": "e2e is measured at N=1 through the host C ABI; multi-GPU runs keep shards device-resident"}
        from sfb_b200.device import shard_shells
        s_lo, s_hi = shard_shells(wl.nr, world)[rank]
        ok, host_shard, host_out, err = 1.0, None, None, ""
        try:    # the only steps that can fail on one rank alone: host allocations
            host_shard = torch.empty((d_win.shape[0], max(1, s_hi - s_lo)), dtype=torch.float64).pin_memory()
            if s_hi > s_lo:
                host_shard.copy_(d_win[:, s_lo:s_hi].cpu())
            host_out = torch.empty((n, n), dtype=torch.float64).pin_memory() if rank == 0 else None
        except Exception as exc:  # noqa: BLE001
            ok, err = 0.0, str(exc)
        flag = torch.tensor([ok], device="cuda", dtype=torch.float64)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)     # every rank takes the same branch: no collective can hang
        if float(flag.item()) == 1.0:
            def e2e_step():
                if s_hi > s_lo:
                    d_win[:, s_lo:s_hi].copy_(host_shard, non_blocking=True)
                out = step()
                if rank == 0:
                    host_out.copy_(out, non_blocking=True)
                torch.cuda.synchronize()

            e2e_step()
            barrier()
            t0 = time.perf_counter()
            k = max(1, min(args.steps, 3))
            for _ in range(k):
                e2e_step()
            barrier()
            dt = torch.tensor([(time.perf_counter() - t0) / k], device="cuda", dtype=torch.float64)
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
            dt = float(dt.item())
            e2e = {"value": n * n / dt, "unit": UNIT, "h2d_bytes_per_step": int(win.nbytes),
                   "d2h_bytes_per_step": int(8 * n * n), "ms_per_step": dt * 1e3,
                   "api": "DevicePipeline (one process per GPU): pinned host shells in on every rank, sharded step, "
                          "assembled matrix out to pinned host memory on rank 0"}
        elif err:
            e2e["error"] = err
    elif world > 1:
        e2e = {"value": None, "unit": UNIT, "h2d_bytes_per_step": None, "d2h_bytes_per_step": None,
               "note": "--no-e2e"}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        v, det = cpu_reference_run(wl, args.cpu_seconds)
        cpu = {"value": v, "unit": UNIT, "cores": det["cores"], "kind": det["kind"], "sample": det["sample"],
               "stage_s": det["stage_s"], "stage3_gflops": det["stage3_gflops"]}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": dict(wl.describe(), l2="working set (win 0.4 GB + ring buffers 0.3 GB + M %.1f GB) exceeds the "
                                             "126 MB L2, no explicit flush" % (8e-9 * n * n),
                           parallelism=f"sharded x{world}" + (" (stage 1 shell-sharded + NCCL all-gather of W_lm(r); "
                                                                   "M sharded by columns (L,N,N'): each rank forms the "
                                                                   "l<=L blocks of its columns in upper-packed storage "
                                                                   "(half the matrix bytes); pull: the unpack+mirror "
                                                                   "kernel reads every column from its owner's peer-"
                                                                   "mapped buffer over NVLink (exchange fused into the "
                                                                   "kernel); packed: in-place NCCL all-gather of the "
                                                                   f"slabs first; exchange={xmode})"
                                                                   if world > 1 else "")),
            "roofline": roofline, "per_rank": per_rank, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches_per_step * args.steps,
            "clocks": clocks,
        }
        print(json.dumps(line))
    if world > 1:
        if pm is not None:
            pm.close()
        if pb is not None:
            pb.close()
        dist.destroy_process_group()


def main():
    args = parse_args()
    # stdout carries exactly ONE JSON line: anything a library prints there meanwhile (NCCL's version banner, ...)
    # is sent to stderr by pointing fd 1 at fd 2 until the result is ready
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    out = io.StringIO()
    try:
        with contextlib.redirect_stdout(out):
            if args.impl == "reference":
                run_reference(args)
            else:
                run_b200(args)
    finally:
        sys.stdout.flush()
        os.dup2(saved, 1)
        os.close(saved)
    lines = [ln for ln in out.getvalue().splitlines() if ln.startswith("{")]
    rest = [ln for ln in out.getvalue().splitlines() if not ln.startswith("{")]
    if rest:
        print("\n".join(rest), file=sys.stderr)
    if lines:
        print(lines[-1], flush=True)


if __name__ == "__main__":
    main()
