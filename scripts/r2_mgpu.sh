#!/bin/bash
# Multi-GPU evidence: the multi-rank parity tests and the scaling benches (weak: 1M transitions and batch 256 per GPU;
# strong: global batch 256 and an 8M-transition buffer split over the ranks).  usage: gpurun --gpus N -- bash scripts/r2_mgpu.sh N
set -u
N=${1:-2}
OUT=gpurun_out/r2; mkdir -p $OUT
timeout -s KILL 900 python -m pytest tests/test_gpu_multirank.py -x -q --timeout 400 -rs > $OUT/mgpu_tests_n$N.log 2>&1; echo "rc=$?" >> $OUT/mgpu_tests_n$N.log; tail -6 $OUT/mgpu_tests_n$N.log
for G in $(seq 1 $N); do
  case $G in 1|2|4|8) ;; *) continue;; esac
  for SC in weak strong; do
    if [ $G -eq 1 ]; then CMD="python bench.py"; else CMD="python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 2960$G bench.py"; fi
    timeout -s KILL 900 $CMD --gpus $G --steps 10000 --warmup 500 --scaling $SC --no-cpu-baseline --no-batch-sweep > $OUT/bench_${SC}_${G}gpu.json 2> $OUT/bench_${SC}_${G}gpu.err
    echo "$SC $G rc=$?: $(cut -c1-400 $OUT/bench_${SC}_${G}gpu.json)"
  done
done
