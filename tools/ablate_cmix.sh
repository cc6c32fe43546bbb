#!/bin/bash
# phase ablation of the shared-memory-Z kernel cmix_block_kernel (run with SFB_CMIX_OLD=1; tools/ablate_regz.sh covers the
# register-Z kernel).  Timing only: results are wrong by construction when SFB_CMIX_DBG != 0.
for D in 0 1 2 3 4 7; do
  echo -n "dbg $D block_ms "
  SFB_CMIX_OLD=1 SFB_CMIX_DBG=$D python bench.py --steps 3 --warmup 2 --no-cpu --no-e2e 2>&1 | tail -1 | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(d["stage_ms"]["block"])'
done
