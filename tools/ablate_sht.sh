#!/bin/bash
# phase ablation of legendre_analysis_kernel (timing only; results are wrong by construction when SFB_SHT_DBG != 0)
# bits: 1 skip the DMMA GEMM, 2 skip the global loads
for D in 0 1 2 3; do
  echo -n "dbg $D stage1_ms "
  SFB_SHT_DBG=$D python bench.py --steps 3 --warmup 2 --no-cpu --no-e2e 2>&1 | tail -1 | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(d["stage_ms"]["stage1"])'
done
