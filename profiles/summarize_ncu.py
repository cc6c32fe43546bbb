#!/usr/bin/env python
"""Summarise an Nsight Compute report (ncu -i X.ncu-rep --page raw --csv) into the handful of metrics the
roofline discussion in DESIGN.md uses.  Usage: python profiles/summarize_ncu.py gpurun_out/prof.ncu-rep > profiles/X.txt"""
import csv
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe (DMMA) active %"),
    ("sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active", "DMMA inst % of peak"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64 (non-tensor) pipe active %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem/block"),
    ("launch__shared_mem_per_block_static", "static smem/block"),
    ("launch__grid_size", "grid"),
    ("launch__waves_per_multiprocessor", "waves/SM"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("smsp__inst_executed.sum", "warp instructions"),
]


def main(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    print(f"# ncu --set full --clock-control none summary of {path}")
    for n, r in enumerate(data):
        print(f"\n[{n}] {r[idx['Kernel Name']]}")
        for key, label in WANT:
            if key in idx:
                print(f"    {label:34s} {r[idx[key]]} {units[idx[key]]}")


if __name__ == "__main__":
    main(sys.argv[1])
