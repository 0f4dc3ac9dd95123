#!/bin/bash
# wide step iteration: parity tests, batch sweep, per-kernel durations (ncu launch list)
mkdir -p gpurun_out/r2w
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "wide" 2>&1 | tail -8 | tee gpurun_out/r2w/wide_tests.log
: > gpurun_out/r2w/batch_sweep_wide.txt
for B in 1024 4096 16384 65536; do
  SWEEP_STEPS=50 timeout 120 python scripts/batch_sweep.py $B 2>&1 | grep "^B=" | tee -a gpurun_out/r2w/batch_sweep_wide.txt
done
bash scripts/r2_wide_ncu.sh
