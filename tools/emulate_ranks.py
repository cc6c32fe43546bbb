"""Time, on ONE GPU, the work each rank of an N-GPU sharded step does (stage 1 on its shells, stage 2+3 on its column
range) — the per-rank critical path without the collectives.  Lets the N = 8 balance be tuned without an 8-GPU lease.
Usage: python tools/emulate_ranks.py [--world 8] [--config 4] [--reps 5]"""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--world", type=int, default=8)
    ap.add_argument("--config", default="4")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--mode", default="lshard", choices=["lshard", "cols"],
                    help="lshard: upper-packed columns + locally mirrored rows; cols: full-height column slabs")
    args = ap.parse_args()
    import torch

    import sfb_b200  # noqa: F401
    from sfb_b200 import _lib, configs
    from sfb_b200.device import DevicePipeline, shard_rows, shard_shells

    lib = _lib.load()
    wl = configs.Workload(int(args.config) if args.config.isdigit() else args.config)
    pipe = DevicePipeline(wl.wmodes, wl.cmodes, wl.G)
    d_win = torch.from_numpy(np.ascontiguousarray(wl.win.T)).cuda()
    W = args.world
    pipe.calc_wr_lm(d_win)
    torch.cuda.synchronize()
    off = pipe.packed_offsets()
    ranges = (pipe.packed_shard_ranges(W, balance="lshard") if args.mode == "lshard"
              else shard_rows(pipe.col_costs, pipe.ell_of_row, W))
    shells = shard_shells(pipe.nr, W)

    def ev():
        return torch.cuda.Event(enable_timing=True)

    res = []
    maxcols = max(h - l for l, h in ranges)
    out = torch.empty((maxcols, pipe.nout), dtype=torch.float64, device="cuda") if args.mode == "cols" else None
    for g in range(W):
        lo, hi = ranges[g]
        slo, shi = shells[g]
        cnt = shi - slo
        rec = {"rank": g, "cols": [int(lo), int(hi)], "shells": cnt}
        if cnt > 0:
            plan = C.c_void_p()
            _lib.check(lib.sfb_sht_plan_create(C.byref(plan), pipe.nside_in, pipe.amodes.nside, pipe.LMAX, cnt))
            nrp = 8 * (-(-cnt // 8))
            alm = torch.zeros(pipe.lmsize * 2 * nrp, dtype=torch.float64, device="cuda")
            ts = []
            for _ in range(args.reps + 2):
                e0, e1 = ev(), ev()
                e0.record()
                _lib.check(lib.sfb_calc_wr_lm_dev(plan, d_win.data_ptr() + 8 * slo, pipe.nr, 3, alm.data_ptr(),
                                                  pipe._stream()))
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            rec["stage1_ms"] = round(float(np.median(ts[2:])), 4)
            lib.sfb_sht_plan_destroy(plan)
        if hi > lo:
            ts, det = [], None
            if args.mode == "lshard":
                slab = torch.empty(int(off[hi] - off[lo]), dtype=torch.float64, device="cuda")
                rows = torch.empty((hi, hi - lo), dtype=torch.float64, device="cuda")
            for _ in range(args.reps + 2):
                e0, e1 = ev(), ev()
                e0.record()
                if args.mode == "lshard":
                    pipe.power_win_mix_lshard(lo, hi, packed_slab=slab, rows=rows)
                else:
                    pipe.power_win_mix_cols(lo, hi, out=out[: hi - lo])
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
                det = _lib.timings()
            rec["stage23_ms"] = round(float(np.median(ts[2:])), 4)
            rec.update({k: round(det[k], 4) for k in ("wl_ms", "what_ms", "block_ms", "fill_ms")})
            if args.mode == "lshard":
                del slab, rows
        res.append(rec)
        print(json.dumps(rec), flush=True)
    crit = max(r.get("stage1_ms", 0) for r in res) + max(r.get("stage23_ms", 0) for r in res)
    print(json.dumps({"world": W, "critical_path_ms_without_gather": round(crit, 4)}))


if __name__ == "__main__":
    main()
