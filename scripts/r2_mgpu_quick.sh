#!/bin/bash
# quick multi-rank check after host-side changes: the multi-rank parity tests and one driver-style scaling line at N ranks
set -u
N=${1:-2}
OUT=gpurun_out/r2c; mkdir -p $OUT
timeout -s KILL 900 python -m pytest tests/test_gpu_multirank.py -x -q --timeout 400 -rs > $OUT/mgpu_tests_n$N.log 2>&1; echo "rc=$?" >> $OUT/mgpu_tests_n$N.log; tail -6 $OUT/mgpu_tests_n$N.log
timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N --steps 20 --warmup 5 > $OUT/bench_driver_${N}gpu.json 2> $OUT/bench_driver_${N}gpu.err
echo "driver-args $N rc=$?: $(cut -c1-700 $OUT/bench_driver_${N}gpu.json)"
