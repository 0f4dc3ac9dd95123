#!/bin/bash
# ncu --set full of one launch of each wide kernel at B = 65536 (import-source: per-line stall samples)
mkdir -p gpurun_out/r2w
for K in fwd bwd wgrad; do
  SWEEP_STEPS=2 timeout 600 ncu --set full --clock-control none --import-source on --kernel-name regex:k_wide_$K --launch-skip 3 -c 1 \
    -f -o gpurun_out/r2w/wide_$K python scripts/batch_sweep.py ${1:-65536} > gpurun_out/r2w/ncu_full_$K.log 2>&1
  tail -3 gpurun_out/r2w/ncu_full_$K.log
done
ls -la gpurun_out/r2w
