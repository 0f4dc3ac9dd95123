#!/bin/bash
# 8-GPU evidence: multi-rank parity tests (2 / 4 / 8 ranks; tile kernel, LSTM tensor-core path, cluster kernel), weak and strong
# scaling bench lines at 4 and 8 ranks, phase report of the fused exchange at 8 ranks.   gpurun --gpus 8 -- bash scripts/r2_mgpu8.sh
set -u
OUT=gpurun_out/r2; mkdir -p $OUT
timeout -s KILL 900 python -m pytest tests/test_gpu_multirank.py -x -q --timeout 400 -rs > $OUT/mgpu_tests_n8.log 2>&1; echo "rc=$?" >> $OUT/mgpu_tests_n8.log; tail -4 $OUT/mgpu_tests_n8.log
for G in 4 8; do
  for SC in weak strong; do
    timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 2970$G bench.py \
      --gpus $G --steps 10000 --warmup 500 --scaling $SC --no-cpu-baseline --no-batch-sweep > $OUT/bench_${SC}_${G}gpu.json 2> $OUT/bench_${SC}_${G}gpu.err
    echo "$SC $G rc=$?: $(cut -c1-330 $OUT/bench_${SC}_${G}gpu.json)"
  done
done
SMB200_PROFILE=1 timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29799 scripts/phase_report.py > $OUT/phase_report_8gpu_v5.txt 2>&1
grep -A45 "rank 0 of 8" $OUT/phase_report_8gpu_v5.txt | head -50
