#!/bin/bash
# ncu --set full of one launch of ONE wide kernel: $1 = kernel suffix (fwd / bwd / wgrad ...), $2 = batch
mkdir -p gpurun_out/r2w
SWEEP_STEPS=2 timeout 600 ncu --set full --clock-control none --import-source on --kernel-name regex:k_wide_$1 --launch-skip 3 -c 1 \
  -f -o gpurun_out/r2w/wide_$1 python scripts/batch_sweep.py ${2:-65536} > gpurun_out/r2w/ncu_full_$1.log 2>&1
tail -2 gpurun_out/r2w/ncu_full_$1.log | cut -c1-200
