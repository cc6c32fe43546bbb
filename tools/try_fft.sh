#!/bin/bash
for E in 4096 2048 8192; do
  echo -n "fft elems $E stage1_ms "
  SFB_FFT_ELEMS=$E python bench.py --steps 3 --warmup 2 --no-cpu --no-e2e 2>&1 | tail -1 | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(d["stage_ms"]["stage1"])'
done
