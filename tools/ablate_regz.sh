#!/bin/bash
# phase ablation of cmix_regz_kernel (timing only; results are wrong by construction when SFB_CMIX_DBG != 0)
# bits: 1 skip epilogue, 2 skip Z phase, 4 skip T-phase DMMA
for D in 0 1 2 4 7; do
  echo -n "dbg $D "
  SFB_CMIX_DBG=$D python bench.py --steps 3 --warmup 2 --no-cpu --no-e2e 2>&1 | tail -1 | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(d["stage_ms"], d["roofline"]["frac"])'
done
echo -n "no-mirror "
SFB_NO_MIRROR=1 python bench.py --steps 3 --warmup 2 --no-cpu --no-e2e 2>&1 | tail -1 | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(d["stage_ms"])'
echo -n "full-diag-tiles "
SFB_REGZ_FULLDIAG=1 python bench.py --steps 3 --warmup 2 --no-cpu --no-e2e 2>&1 | tail -1 | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(d["stage_ms"])'
echo -n "one-block-per-CTA "
SFB_REGZ_ONEBLOCK=1 python bench.py --steps 3 --warmup 2 --no-cpu --no-e2e 2>&1 | tail -1 | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(d["stage_ms"], d["roofline"]["frac"])'
