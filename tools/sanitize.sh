#!/bin/bash
# compute-sanitizer over the stage 2+3 paths on a small configuration (memcheck, then racecheck of the shared-memory
# protocols: TMA stages + mbarriers of the persistent block kernel, per-warp staging tiles, mirror kernels)
T='tests/test_gpu_parity.py -m gpu -x -q -k "device_pipeline or alternative_paths"'
echo "== memcheck"; eval timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest $T 2>&1 | tail -6
echo "== racecheck"; eval timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "device_pipeline" 2>&1 | tail -8
