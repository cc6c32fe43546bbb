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
    ap.add_argument("--no-assembled", action="store_true", help="N > 1: skip the full-matrix-on-every-GPU variant")
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
    # torchrun exports OMP_NUM_THREADS=1 to its workers: the CPU arm must still use every host core
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
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
    cref.calc_cmix_rows(wl.cmodes.lnn, probe, wl.G, Wc, nthreads=cores)
    tp = time.perf_counter() - t0
    nrows = int(max(2 * cores, min(n, seconds / (tp / probe.size))))
    rows = np.unique(np.linspace(0, n - 1, nrows).astype(np.int64)) + 1
    t0 = time.perf_counter()
    cref.calc_cmix_rows(wl.cmodes.lnn, rows, wl.G, Wc, nthreads=cores)
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
def strided_checksum_cols(cols_t, col_lo, n):
    """Σ M[i, j] over i ∈ 0::n//97, j ∈ 0::n//89 restricted to the columns [col_lo, col_lo + cols) this tensor holds
    (tensor[j - col_lo, i] = M[i, j]).  Summed over all shards it is the same number for every N."""
    import torch
    si, sj = max(1, n // 97), max(1, n // 89)
    hi = col_lo + cols_t.shape[0]
    j0 = -(-col_lo // sj) * sj
    if j0 >= hi:
        return 0.0
    jj = torch.arange(j0, hi, sj, device=cols_t.device) - col_lo
    return float(cols_t[jj][:, ::si].sum().item())


def run_b200(args):
    import torch
    import torch.distributed as dist

    import sfb_b200 as sfb
    from sfb_b200 import _lib, configs
    from sfb_b200.device import DevicePipeline, PeerBuffer, shard_rows, stream_barrier
    from sfb_b200.separable import SeparableArray

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the CUDA arm has no CPU fallback")
    torch.cuda.set_device(local_rank)
    cpu_group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        # host-side barrier for the e2e leg: an NCCL barrier would keep a spinning kernel on every idle rank's GPU while
        # rank 0 drives all GPUs through the C ABI
        cpu_group = dist.new_group(backend="gloo")
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

    # Output modes of the device-timed step (`value`):
    #   N = 1   : the whole matrix in HBM (mirror mode: L >= l blocks + fill pass)
    #   N > 1   : "L-shaped shards" (default) — every element of M is formed exactly once, by the rank whose L range holds
    #             max(l_i, l_i'): the blocks l <= L of its columns in upper-packed storage plus, mirrored locally from them,
    #             its rows of the part below the block diagonal.  No redundant flops, no exchange; stage 1 is shell-sharded +
    #             NCCL all-gather of W_lm(r).
    #             "column_slabs" (also timed) — rank g forms a full-height column range of M (what the per-GPU host copy of
    #             sfb_set_devices and the binned product want; twice the block flops).
    #             "assembled" (also timed) — the full matrix on EVERY GPU: upper-packed column shards pulled from their
    #             owners over NVLink inside the unpack+mirror kernel.
    ranges = shard_rows(pipe.col_costs, pipe.ell_of_row, world)          # column slabs
    lo, hi = ranges[rank]
    ell = np.asarray(pipe.ell_of_row)
    if world > 1:
        off = pipe.packed_offsets()
        lranges = pipe.packed_shard_ranges(world, balance="lshard")      # L-shaped shards
        llo, lhi = lranges[rank]
        slab = torch.zeros(max(1, int(off[lhi] - off[llo])), dtype=torch.float64, device="cuda")
        rows_t = torch.zeros((max(1, lhi), max(1, lhi - llo)), dtype=torch.float64, device="cuda")
        out_t = None
    else:
        out_t = torch.empty((max(1, hi - lo), pipe.nout), dtype=torch.float64, device="cuda")
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(3)]

    def step():
        evs[0].record()
        pipe.calc_wr_lm_sharded(d_win)
        evs[1].record()
        if world == 1:
            out = pipe.power_win_mix_rows(0, pipe.nout, out=out_t)
        else:
            out = pipe.power_win_mix_lshard(llo, lhi, packed_slab=slab, rows=rows_t) if lhi > llo else (slab, rows_t)
        evs[2].record()
        return out

    def lshard_sample(i_idx, j_idx):
        """Values M[i, j] of the listed elements this rank owns (others 0), read from its packed slab / mirrored rows."""
        i_idx, j_idx = np.asarray(i_idx), np.asarray(j_idx)
        up = (ell[i_idx] <= ell[j_idx]) & (j_idx >= llo) & (j_idx < lhi)
        dn = (ell[i_idx] > ell[j_idx]) & (i_idx >= llo) & (i_idx < lhi)
        vals = torch.zeros(i_idx.size, dtype=torch.float64, device="cuda")
        if up.any():
            pos = torch.from_numpy((off[j_idx[up]] - off[llo] + i_idx[up]).astype(np.int64)).cuda()
            vals[torch.from_numpy(np.flatnonzero(up)).cuda()] = slab[pos]
        if dn.any():
            pos = torch.from_numpy((j_idx[dn] * (lhi - llo) + (i_idx[dn] - llo)).astype(np.int64)).cuda()
            vals[torch.from_numpy(np.flatnonzero(dn)).cuda()] = rows_t.view(-1)[pos]
        return vals, up | dn

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, collect=None):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
            if collect:
                collect()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1) / steps
        if world > 1:
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    for _ in range(args.warmup):
        step()
    barrier()
    stage_ms = {"stage1": [], "wl": [], "what": [], "block": [], "fill": [], "k3": [], "stage1_incl_gather": [], "stage23": []}
    launches = [0]

    def collect():
        tim = _lib.timings()
        for k, key in (("stage1", "stage1_ms"), ("wl", "wl_ms"), ("what", "what_ms"), ("block", "block_ms"), ("fill", "fill_ms"),
                       ("k3", "k3_ms")):
            stage_ms[k].append(tim[key])
        torch.cuda.synchronize()
        stage_ms["stage1_incl_gather"].append(evs[0].elapsed_time(evs[1]))
        stage_ms["stage23"].append(evs[1].elapsed_time(evs[2]))
        launches[0] = int(tim["launches"])
        collect.flops = tim["block_flops"]

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms = timed(step, args.steps, collect)
    clocks = sampler.stop() if rank == 0 else None
    value = n * n / (ms * 1e-3)
    launches_per_step = launches[0]
    sm = {k: statistics.mean(v) for k, v in stage_ms.items()}

    # ---- correctness carried by the bench line: checksum (same formula for every N) and parity against a 1-rank recompute
    full = step()
    torch.cuda.synchronize()
    column_slabs = None
    if world == 1:
        checksum = strided_checksum_cols(full, 0, n)
        parity = None
    else:
        si, sj = max(1, n // 97), max(1, n // 89)
        ii, jj = np.meshgrid(np.arange(0, n, si), np.arange(0, n, sj), indexing="ij")
        vals, mine = lshard_sample(ii.ravel(), jj.ravel())
        cs = torch.stack([vals.sum(), torch.tensor(float(mine.sum()), device="cuda", dtype=torch.float64)])
        dist.all_reduce(cs)
        checksum = float(cs[0].item())
        if int(round(float(cs[1].item()))) != ii.size:
            raise SystemExit("bench.py: the L-shaped shards do not cover every sampled element exactly once")
        # every rank recomputes, on its own GPU alone (full stage 1, no collective), a few of its columns and rows
        alm_sharded = pipe.alm.clone()
        pipe.calc_wr_lm(d_win)
        torch.cuda.synchronize()
        err = float(((pipe.alm - alm_sharded).norm() / pipe.alm.norm()).item())
        if lhi > llo:
            for c0 in sorted({llo, (llo + lhi) // 2, max(llo, lhi - 8)}):
                c1 = min(lhi, c0 + 8)
                ref = pipe.power_win_mix_cols(c0, c1)                 # (c1-c0, nout): ref[j-c0, i] = M[i, j]
                ref_r = pipe.power_win_mix_rows(c0, c1)               # (nout, c1-c0): ref_r[j, i-c0] = M[i, j]
                torch.cuda.synchronize()
                jc, ic = np.meshgrid(np.arange(c0, c1), np.arange(n), indexing="ij")
                keep = ell[ic] <= ell[jc]
                got, own = lshard_sample(ic[keep], jc[keep])
                assert own.all()
                want = ref[torch.from_numpy(jc[keep] - c0).cuda(), torch.from_numpy(ic[keep]).cuda()]
                err = max(err, float(((got - want).norm() / want.norm()).item()))
                ir, jr = np.meshgrid(np.arange(c0, c1), np.arange(n), indexing="ij")
                keep = ell[jr] < ell[ir]
                if keep.any():
                    got, own = lshard_sample(ir[keep], jr[keep])
                    assert own.all()
                    want = ref_r[torch.from_numpy(jr[keep]).cuda(), torch.from_numpy(ir[keep] - c0).cuda()]
                    err = max(err, float(((got - want).norm() / want.norm()).item()))
        e = torch.tensor([err], device="cuda", dtype=torch.float64)
        dist.all_reduce(e, op=dist.ReduceOp.MAX)
        parity = float(e.item())
        if not parity < 1e-10:
            raise SystemExit(f"bench.py: sharded result differs from the single-GPU recompute (rel err {parity:.3e})")
        pipe.alm.copy_(alm_sharded)

        # ---- column slabs (full-height column range per rank), timed the same way
        out_c = torch.empty((max(1, hi - lo), pipe.nout), dtype=torch.float64, device="cuda")

        def step_cols():
            pipe.calc_wr_lm_sharded(d_win)
            return pipe.power_win_mix_cols(lo, hi, out=out_c) if hi > lo else out_c

        for _ in range(args.warmup):
            step_cols()
        cms = timed(step_cols, max(3, args.steps // 2))
        ccs = torch.tensor([strided_checksum_cols(out_c, lo, n) if hi > lo else 0.0], device="cuda", dtype=torch.float64)
        dist.all_reduce(ccs)
        column_slabs = {"ms_per_step": cms, "value": n * n / (cms * 1e-3), "unit": UNIT, "checksum": float(ccs.item()),
                        "cols_of_rank": [list(map(int, r)) for r in ranges],
                        "note": "rank g forms a full-height column range of M (a contiguous slab of the column-major matrix): "
                                "the form the host copy of sfb_set_devices and the binned product use; twice the block flops"}
        del out_c

    per_rank = None
    if world > 1:
        mine = {"rank": rank, "L_range_cols": [int(llo), int(lhi)], **{k: round(v, 3) for k, v in sm.items()}}
        per_rank = [None] * world
        dist.all_gather_object(per_rank, mine)

    # ---- N > 1: the assembled mode (full matrix on every GPU), timed the same way and checked the same way
    assembled = None
    if world > 1 and not args.no_assembled:
        off = pipe.packed_offsets()
        pranges = pipe.packed_shard_ranges(world)
        plo, phi_ = pranges[rank]
        pb = PeerBuffer(int(off[-1]))
        full_t = torch.empty((pipe.nout, pipe.nout), dtype=torch.float64, device="cuda")

        def step_assembled():
            pipe.calc_wr_lm_sharded(d_win)
            if phi_ > plo:
                pipe.power_win_mix_upper_packed(plo, phi_, pb.tensor)
            stream_barrier(pipe)
            pipe.unpack_mirror_pull(pb, pranges, full_t)
            stream_barrier(pipe)

        for _ in range(args.warmup):
            step_assembled()
        ams = timed(step_assembled, max(3, args.steps // 2))
        acs = strided_checksum_cols(full_t, 0, n)
        aerr = 0.0
        if lhi > llo:   # against this rank's packed columns (full_t[j, i] = M[i, j])
            jc, ic = np.meshgrid(np.arange(llo, lhi, max(1, (lhi - llo) // 16)), np.arange(0, n, 7), indexing="ij")
            keep = ell[ic] <= ell[jc]
            got, _ = lshard_sample(ic[keep], jc[keep])
            want = full_t[torch.from_numpy(jc[keep]).cuda(), torch.from_numpy(ic[keep]).cuda()]
            aerr = float(((got - want).norm() / want.norm()).item())
        e = torch.tensor([aerr], device="cuda", dtype=torch.float64)
        dist.all_reduce(e, op=dist.ReduceOp.MAX)
        assembled = {"ms_per_step": ams, "value": n * n / (ams * 1e-3), "unit": UNIT, "checksum": acs,
                     "rel_err_vs_sharded": float(e.item()),
                     "note": "full matrix on EVERY GPU: upper-packed column shards (l <= L blocks) pulled from their owners "
                             "over NVLink inside the unpack+mirror kernel"}
        del full_t
        pb.close()

    # ---- roofline of the dominant kernel (measured live: CUDA events on the launching stream, inside the lib) ----
    # the DMMA block kernel launches alone (block_ms adds the part of the mirror fill that is not hidden under them; at N = 1
    # the fill of an l-chunk runs on a second stream under the block kernel of the next chunk, and its launches sum to fill_ms)
    kern_ms = sm["k3"]
    dmma = np.zeros(1)
    _lib.check(lib.sfb_probe_dmma_tflops(_lib.ptr(dmma)))
    peaks, traffic = {}, {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    try:
        tfile = [f for f in ("r02_final_traffic.json", "r02_traffic.json") if os.path.exists(os.path.join(ROOT, "profiles", f))]
        traffic = json.load(open(os.path.join(ROOT, "profiles", tfile[0])))["kernels"]
    except (OSError, KeyError):
        pass
    nl = np.asarray(wl.amodes.nmax_l, dtype=np.float64)
    if world > 1:
        lo, hi = llo, lhi
    ells_mine = np.unique(wl.cmodes.lnn[0, lo:hi])
    nn = nl[ells_mine][:, None] * nl[None, :]
    f_alg_block = float(np.sum(2 * nn * wl.nr ** 2 + 2 * nn * nn * wl.nr))
    exec_tflops = collect.flops / (kern_ms * 1e-3) / 1e12
    regz_traffic = sum(v["traffic_bytes"] for k, v in traffic.items() if k.startswith(("cmix_regz", "sfb::cmix_regz"))) or None
    if world > 1 or str(args.config) != "4":
        regz_traffic = None
    hbm = peaks.get("hbm_gbs")
    roofline = {
        "kernel": "cmix_regz_persist_kernel (stage 2+3 block GEMMs, FP64 DMMA, all tile classes of one step)", "bound": "tensor",
        "achieved": exec_tflops, "peak": float(dmma[0]), "unit": "TFLOP/s", "frac": exec_tflops / float(dmma[0]),
        "traffic": regz_traffic,
        "algorithmic_bytes": 8.0 * (hi - lo) * n / (2 if world == 1 else 1),
        "peak_source": "FP64 DMMA probe measured in this run (MEASURED_PEAKS.json has no FP64 figure; nominal B200 "
                       "FP64 tensor peak is 37-40 TFLOP/s)",
        "algorithmic_tflops": f_alg_block / (kern_ms * 1e-3) / 1e12,
        "algorithmic_flops": f_alg_block, "executed_flops": collect.flops, "launch_ms": kern_ms,
        "note": "achieved/frac = DMMA flops the kernel EXECUTES (n padded to multiples of 8) over its own launch time; "
                "SURVEY §8d's algorithmic count is larger because only N<=N' tiles (and at N=1 only L>=l blocks) are formed",
    }
    fill_ms, s1_ms = sm["fill"], sm["stage1"]
    roofline_fill = {"kernel": "cmix_mirror_fill_kernel", "ms": fill_ms, "bound": "hbm",
                     "achieved": (8.0 * n * n / (fill_ms * 1e-3) / 1e9) if fill_ms > 0 else None, "peak": hbm, "unit": "GB/s",
                     "frac": (8.0 * n * n / (fill_ms * 1e-3) / 1e9 / hbm) if (fill_ms > 0 and hbm) else None,
                     "exposed_ms": sm["block"] - sm["k3"],
                     "note": "ms = sum of the fill launches (second stream, each under the block kernel of the next l-chunk, "
                             "so slower than alone); exposed_ms = what the step still waits for after its last block kernel"}
    # useful Legendre-GEMM flops of one map2alm(niter=3) on this rank's shells: 7 passes (4 analyses + 3 syntheses) of
    # Σ_m [l x ring] x [ring x 2 shells] with north/south rings folded: 2 · lmsize · 2nside · 2·shells per pass
    shells_here = -(-wl.nr // world)
    lmsize = (wl.LMAX + 1) * (wl.LMAX + 2) // 2
    s1_exec = 7 * 2.0 * lmsize * (2 * wl.amodes.nside) * 2 * shells_here
    s1_traffic = None
    if world == 1 and str(args.config) == "4" and traffic:
        s1_traffic = sum(v["traffic_bytes"] for k, v in traffic.items()
                         if k.startswith(("cap_analysis", "belt_analysis", "legendre_", "gram_apply", "ring_alias"))
                         and not (k.startswith(("cap_analysis", "belt_analysis")) and "#" in k)) or None
    roofline_stage1 = {"ms": s1_ms, "bound": "tensor", "traffic": s1_traffic,
                       "executed_flops": s1_exec, "achieved": s1_exec / (s1_ms * 1e-3) / 1e12 if s1_ms > 0 else None,
                       "peak": float(dmma[0]), "unit": "TFLOP/s",
                       "frac": s1_exec / (s1_ms * 1e-3) / 1e12 / float(dmma[0]) if s1_ms > 0 else None,
                       "alg_flops": wl.flops_alg_stage1() / world, "alg_bytes": wl.bytes_alg_stage1() / world,
                       "alg_gbs": wl.bytes_alg_stage1() / world / (s1_ms * 1e-3) / 1e9 if s1_ms > 0 else None,
                       "hbm_frac": (wl.bytes_alg_stage1() / world / (s1_ms * 1e-3) / 1e9 / hbm) if (hbm and s1_ms > 0) else None,
                       "note": "executed_flops = useful (unpadded) DMMA flops of the 7 Legendre GEMM passes of one "
                               "map2alm(niter=3) on this rank's shells, over the whole stage-1 time (ring FFT/DFT and alias "
                               "passes included in the time, not in the flops); alg_* = SURVEY §8d's F1, B1"}

    # ---- e2e through the reference-facing host API (C ABI with HOST buffers) on `world` GPUs of this process ----
    e2e = None
    if not args.no_e2e:
        del full
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier(group=cpu_group)
        if rank == 0:
            e2e = run_e2e(sfb, wl, world, args)
        if world > 1:
            dist.barrier(group=cpu_group)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        v, det = cpu_reference_run(wl, args.cpu_seconds)
        cpu = {"value": v, "unit": UNIT, "cores": det["cores"], "kind": det["kind"], "sample": det["sample"],
               "stage_s": det["stage_s"], "stage3_gflops": det["stage3_gflops"]}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": wl.describe(),
            "l2": "working set (win 0.4 GB + ring buffers 0.3 GB + M %.1f GB) exceeds the 126 MB L2, no explicit flush"
                  % (8e-9 * n * n),
            "output_mode": ("full matrix in HBM" if world == 1 else
                            f"L-shaped shards x{world}: stage 1 by shells + NCCL all-gather of W_lm(r); every element of M formed "
                            "exactly once, on the rank whose L range holds max(l_i, l_i') (blocks l <= L of its columns in "
                            "upper-packed storage + its rows below the block diagonal mirrored locally), left in its owner's "
                            "HBM; see `column_slabs` and `assembled` for the other two forms"),
            "checksum": checksum, "parity_vs_n1": parity,
            "roofline": roofline, "roofline_stage1": roofline_stage1, "roofline_mirror_fill": roofline_fill,
            "stage_ms": sm, "per_rank": per_rank, "column_slabs": column_slabs, "assembled": assembled,
            "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches_per_step * args.steps, "clocks": clocks,
        }
        print(json.dumps(line))
    pipe.close()
    if world > 1:
        dist.destroy_process_group()


def run_e2e(sfb, wl, ndev, args):
    """The reference-facing call `power_win_mix(win, wmodes, cmodes)` -> sfb_power_win_mix (C ABI, HOST buffers) with
    sfb_set_devices(ndev): window in host memory in, full matrix in host memory out, every copy inside the timed region.
    Page-locked buffers (sfb_host_alloc, what the Julia shim allocates) are the headline; the pageable variant is
    reported next to it."""
    n = wl.lnnsize
    win = wl.win
    nr, npix = win.shape
    res = {"value": None, "unit": UNIT, "h2d_bytes_per_step": int(win.nbytes + wl.G.nbytes + wl.cmodes.lnn.nbytes),
           "d2h_bytes_per_step": int(8 * n * n), "n_gpus": ndev,
           "api": f"sfb_set_devices({ndev}); sfb_b200.power_win_mix(win, wmodes, cmodes) -> sfb_power_win_mix (C ABI, "
                  "host pointers, one process)"}
    try:
        sfb.set_devices(ndev)
        host_win = sfb.pinned_empty((nr, npix))
        host_win[...] = win
        out = sfb.pinned_empty((n, n))
        k = max(1, min(args.steps, 5))

        def run(w, o):
            with warnings.catch_warnings():
                warnings.simplefilter("ignore", RuntimeWarning)
                for _ in range(2):
                    sfb.power_win_mix(w, wl.wmodes, wl.cmodes, out=o)
                t0 = time.perf_counter()
                for _ in range(k):
                    M = sfb.power_win_mix(w, wl.wmodes, wl.cmodes, out=o)
                return (time.perf_counter() - t0) / k, M

        dt, M = run(host_win, out)
        from sfb_b200 import _lib
        res.update({"value": n * n / dt, "ms_per_step": dt * 1e3, "buffers": "page-locked (sfb_host_alloc)",
                    "checksum": float(M[:: max(1, n // 97), :: max(1, n // 89)].sum()), "stage_ms": _lib.timings()})
        if not np.isfinite(M).all():
            res["error"] = "non-finite result"
        # pageable buffers: a plain Julia Matrix{Float64} for the result and the window
        try:
            if 8.0 * n * n > 6e9:
                raise RuntimeError("skipped: matrix larger than 6 GB")
            pout = np.empty((n, n), order="F")
            dtp, Mp = run(win, pout)
            res["pageable"] = {"ms_per_step": dtp * 1e3, "value": n * n / dtp,
                               "max_abs_diff_vs_pinned": float(np.abs(Mp[::97, ::89] - M[::97, ::89]).max())}
            del pout, Mp
        except Exception as exc:  # noqa: BLE001
            res["pageable"] = {"error": str(exc)}
        # the binned call of cfg4 (BASELINE.json: "binned ClnnBinnedModes output"): N = w̃ M v, Δl = 4, on the same devices
        try:
            wt, vv = sfb.bandpower_binning_weights(wl.cmodes, dl=4)
            bc = sfb.ClnnBinnedModes(wt, vv, wl.cmodes)
            outN = sfb.pinned_empty((wt.shape[0], vv.shape[1]))
            with warnings.catch_warnings():
                warnings.simplefilter("ignore", RuntimeWarning)
                sfb.power_win_mix(host_win, wt, vv, wl.wmodes, bc, out=outN)
                t0 = time.perf_counter()
                for _ in range(k):
                    sfb.power_win_mix(host_win, wt, vv, wl.wmodes, bc, out=outN)
                dtb = (time.perf_counter() - t0) / k
            tb = _lib.timings()
            hbm = None
            try:
                hbm = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs")
            except OSError:
                pass
            bms = tb["binned_ms"]
            res["binned"] = {"ms_per_step": dtb * 1e3, "value": n * n / dtb, "unit": UNIT, "LNN": int(wt.shape[0]),
                             "d2h_bytes_per_step": int(8 * wt.shape[0] * vv.shape[1]), "n_gpus": ndev, "binned_ms": bms,
                             "checksum": float(outN[::53, ::47].sum()),
                             "roofline": {"kernel": "binned_product_kernel", "bound": "hbm", "unit": "GB/s",
                                          "achieved": 8.0 * n * n / ndev / (bms * 1e-3) / 1e9 if bms > 0 else None,
                                          "peak": hbm,
                                          "frac": (8.0 * n * n / ndev / (bms * 1e-3) / 1e9 / hbm) if (hbm and bms > 0) else None,
                                          "algorithmic_bytes": 8.0 * n * n / ndev,
                                          "note": "per device: every element of its column slab of M is read once"},
                             "note": "power_win_mix(win, w̃, v, wmodes, bcmodes) with Δl=4: all lnnsize² elements of M are "
                                     "formed on the device, only N = w̃Mv returns to the host"}
        except Exception as exc:  # noqa: BLE001
            res["binned"] = {"error": str(exc)}
    except Exception as exc:  # noqa: BLE001
        res["error"] = str(exc)
    finally:
        try:
            sfb.set_devices(1)
        except Exception:  # noqa: BLE001
            pass
    return res


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
