#!/bin/bash
# stage 1 with independent kernels forked onto a second stream (default) against one stream (SFB_SHT_SERIAL=1)
run() { python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e 2>&1 | tail -1 | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(round(d["ms_per_step"],3), {k: round(v,3) for k,v in d["stage_ms"].items()}, "frac", round(d["roofline"]["frac"],3), "checksum", d["checksum"])'; }
echo -n "forked "; run
echo -n "serial "; SFB_SHT_SERIAL=1 run
