#!/bin/bash
# Evidence capture on one B200 (run under gpurun): launch list, full ncu capture of every kernel of a step, summaries.
# Usage: tools/capture_profiles.sh TAG      -> gpurun_out/TAG_launches.csv, TAG_launch_shares.txt, TAG_full_ncu_summary.txt
TAG=${1:-r02_final}
OUT=gpurun_out
mkdir -p $OUT
KERN='regex:cap_|belt_|legendre_|gram_apply|ring_alias|wl_build|what_build|cmix_|udgrade'
# 1. launch list (per-launch durations; cold cache, serialised): 4 steps
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e > $OUT/${TAG}_launches_bench.log 2>&1
python profiles/launch_shares.py $OUT/${TAG}_launches.csv 4 > $OUT/${TAG}_launch_shares.txt 2>&1
# 2. full capture: the kernels of the warm-up step and of the timed step (the summary keeps both; the second is warm)
ncu --set full --clock-control none --import-source on -k "$KERN" -c 48 -o $OUT/${TAG}_full \
    python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e > $OUT/${TAG}_full_bench.log 2>&1
python profiles/summarize_ncu.py $OUT/${TAG}_full.ncu-rep > $OUT/${TAG}_full_ncu_summary.txt 2>&1
ls -la $OUT/${TAG}_full.ncu-rep
rm -f $OUT/${TAG}_full.ncu-rep     # too large to travel back; the summary is what is kept
tail -3 $OUT/${TAG}_launch_shares.txt
grep -c "^\[" $OUT/${TAG}_full_ncu_summary.txt
