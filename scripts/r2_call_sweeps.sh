#!/bin/bash
set -u
OUT=gpurun_out/r2; mkdir -p $OUT
timeout -s KILL 600 python -m pytest tests/test_gpu_sweeps.py -x -q --timeout 200 > $OUT/sweeps_tests.log 2>&1; echo "rc=$?" >> $OUT/sweeps_tests.log; tail -12 $OUT/sweeps_tests.log
timeout -s KILL 900 python scripts/sweep_scaling.py > $OUT/sweep_scaling.jsonl 2> $OUT/sweep_scaling.err; tail -3 $OUT/sweep_scaling.err; cat $OUT/sweep_scaling.jsonl
