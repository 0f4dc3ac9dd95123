#!/bin/bash
# (GPU box) evidence of the wide step: full captures of the three tensor-core kernels at B = 65536, launch list, batch sweep
mkdir -p gpurun_out/r2w
for K in fwd loss bwd wgrad; do bash scripts/r2_wide_ncu_one.sh $K 65536; done
bash scripts/r2_wide_ncu.sh > gpurun_out/r2w/launch_list_wide.txt 2>&1
: > gpurun_out/r2w/batch_sweep_wide.txt
for B in 256 1024 2048 4096 16384 65536; do
  SWEEP_STEPS=50 timeout 120 python scripts/batch_sweep.py $B 2>&1 | grep "^B=" | tee -a gpurun_out/r2w/batch_sweep_wide.txt
done
