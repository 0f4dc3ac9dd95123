#!/bin/bash
# ncu evidence for profiles/: (1) launch list of the default bench command, (2) one --set full capture of the
# persistent step kernel (256 steps in one launch) and of the sweep kernels, (3) the same for the opt-in cluster
# step kernel, (4) cfg3 (RACER + LSTM(64)).  Run on the GPU box:   gpurun -- bash scripts/ncu_capture.sh
# Outputs land in gpurun_out/; scripts/ncu_summary.py turns them into the text files kept under profiles/<round>/.
set -u
mkdir -p gpurun_out
M="gpu__time_duration.sum"
ncu --metrics $M --clock-control none -c 4000 --csv --log-file gpurun_out/launches_bench.csv \
    python bench.py --steps 2000 --warmup 100 --no-cpu-baseline --no-batch-sweep > gpurun_out/bench_under_ncu.log 2>&1
PROF_RANGE=1 PROF_STEPS=256 PROF_SWEEPS=1 timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:'k_steps_persistent|k_sweep|k_moments' -o gpurun_out/full_steps -f python scripts/profile_steps.py > gpurun_out/ncu_full.log 2>&1 || true
tail -5 gpurun_out/ncu_full.log
ncu -i gpurun_out/full_steps.ncu-rep --page raw --csv > gpurun_out/full_steps_raw.csv 2>/dev/null
# the cluster step kernel (opt-in)
SMB200_CLUSTER=1 PROF_RANGE=1 PROF_STEPS=256 timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:'k_steps_cluster' -o gpurun_out/full_cluster -f python scripts/profile_steps.py > gpurun_out/ncu_cluster.log 2>&1 || true
ncu -i gpurun_out/full_cluster.ncu-rep --page raw --csv > gpurun_out/full_cluster_raw.csv 2>/dev/null
# cfg3 (RACER + LSTM): the persistent kernel with the tcgen05 weight-gradient contraction — tensor-pipe evidence
PROF_RANGE=1 PROF_STEPS=64 timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:'k_steps_persistent' -o gpurun_out/full_cfg3 -f python scripts/cfg3_probe.py > gpurun_out/ncu_cfg3.log 2>&1 || true
ncu -i gpurun_out/full_cfg3.ncu-rep --page raw --csv > gpurun_out/full_cfg3_raw.csv 2>/dev/null
ls -la gpurun_out/*.ncu-rep gpurun_out/*.csv
