#!/bin/bash
# first GPU contact of the wide step: parity tests under a watchdog, then the batch sweep with it (default from B = 1024)
mkdir -p gpurun_out/r2w
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "wide" 2>&1 | tail -25 | tee gpurun_out/r2w/wide_tests.log
for B in 1024 4096 16384 65536; do
  SWEEP_STEPS=50 timeout 120 python scripts/batch_sweep.py $B 2>&1 | tail -2 | tee -a gpurun_out/r2w/batch_sweep_wide.txt
done
