#!/usr/bin/env python
"""Per-kernel device time of the power_win_mix steps in an `ncu --metrics gpu__time_duration.sum --csv` launch list.
Usage: python profiles/launch_shares.py gpurun_out/launches.csv [steps_captured] > profiles/X_launch_shares.txt
Plan/probe kernels (tables built once per plan, the DMMA peak probe) are listed separately and excluded from the shares."""
import collections
import csv
import sys

ONCE = ("twiddle_table_kernel", "lambda_table_kernel", "w3j000sq_table_kernel", "dmma_probe_kernel")


def main(path, steps):
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    ix = {h: i for i, h in enumerate(rows[0])}
    agg = collections.OrderedDict()
    for r in rows[1:]:
        if r[ix["Metric Name"]] != "gpu__time_duration.sum":
            continue
        v = float(r[ix["Metric Value"]].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0}.get(r[ix["Metric Unit"]], 1e-6)
        a = agg.setdefault(r[ix["Kernel Name"]][:58], [0, 0.0])
        a[0] += 1
        a[1] += v
    step = {k: a for k, a in agg.items() if not k.startswith(ONCE)}
    tot = sum(a[1] for a in step.values())
    print(f"# per-kernel device time over {steps} power_win_mix step(s) at cfg4 ({path}; ncu gpu__time_duration.sum, cold cache, serialised)")
    for k, a in step.items():
        print(f"{k:60s} launches={a[0]:3d} total {a[1]:8.3f} ms  per-launch {a[1] / a[0]:7.3f} ms  share {100 * a[1] / tot:5.1f}%")
    print(f"{'sum':60s}              total {tot:8.3f} ms  = {tot / steps:.3f} ms per step")
    print("# once per plan / probe (not part of a step):")
    for k, a in agg.items():
        if k.startswith(ONCE):
            print(f"{k:60s} launches={a[0]:3d} total {a[1]:8.3f} ms")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 1)
