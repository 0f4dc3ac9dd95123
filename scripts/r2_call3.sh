#!/bin/bash
set -u
OUT=gpurun_out/r2; mkdir -p $OUT
for i in 1 2 3; do timeout -s KILL 200 python scripts/batch_sweep.py 256 2>/dev/null | head -1; done > $OUT/tile_b256_redstore.log; cat $OUT/tile_b256_redstore.log
timeout -s KILL 300 python scripts/cfg3_probe.py 2>&1 | head -2 | tee $OUT/cfg3_probe.log
bash scripts/ncu_capture.sh > $OUT/ncu_capture.log 2>&1; tail -12 $OUT/ncu_capture.log
