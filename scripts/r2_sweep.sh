mkdir -p gpurun_out/r2
for B in 256 1024 4096 16384 65536; do SWEEP_STEPS=100 timeout -s KILL 200 python scripts/batch_sweep.py $B 2>&1 | grep "^B="; done > gpurun_out/r2/cluster_sweep.log 2>&1
cat gpurun_out/r2/cluster_sweep.log
