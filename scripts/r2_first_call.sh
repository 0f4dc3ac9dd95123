#!/bin/bash
# First gpurun call of the next round: everything that was written after round 1's GPU budget was spent, in one call
# (≈ 6-8 minutes of box time).  Every step is bounded by its own timeout; outputs land in gpurun_out/r2/.
#   gpurun --timeout 1200 -- bash scripts/r2_first_call.sh
# 1. the pending device cases (x86 `Uint += float` count with one action component; retraceExplore sweep)
# 2. the whole GPU suite (regression check of round 1's state on a fresh box)
# 3. batch-size sweep of the MLP step (256 ... 65 536), one process per batch size
# 4. cluster-split layer microbenchmark
# 5. the default bench line
set -u
OUT=gpurun_out/r2
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks_throttle_reasons.active --format=csv > $OUT/gpu.txt 2>&1

# the pending cases in the open (under pytest they are non-strict xfails whose report is hidden); the full suite below
# lists them again as XPASS / XFAIL (-rxX)
for case in vracer_da1 vracer_explore vracer_b1024; do
  SMB200_UNVERIFIED=1 timeout 200 python - "$case" > $OUT/pending_$case.log 2>&1 <<'EOF'
import os, sys
root = os.getcwd()
sys.path[:0] = [root, os.path.join(root, "oracle"), os.path.join(root, "tests")]
import test_gpu_zz_pending as t
exec(compile(t.CHILD.format(root=root, oracle=os.path.join(root, "oracle"), tests=os.path.join(root, "tests"), case=sys.argv[1]), "child", "exec"))
EOF
  echo "$case rc=$?" >> $OUT/pending_$case.log
done

: > $OUT/pending_full_size.log
for fx in cfg2_full_props.npz cfg3_full_props.npz; do
  timeout 300 python - "$fx" >> $OUT/pending_full_size.log 2>&1 <<'EOF'
import os, sys
root = os.getcwd()
sys.path[:0] = [root, os.path.join(root, "tests")]
import test_gpu_zz_pending as t
exec(compile(t.CHILD_FULL.format(root=root, oracle=os.path.join(root, "oracle"), tests=os.path.join(root, "tests"), fixture=sys.argv[1]), "child", "exec"))
EOF
  echo "full size $fx rc=$?" >> $OUT/pending_full_size.log
done

timeout 900 python -m pytest tests -m gpu -x -q -rxX > $OUT/gpu_tests.log 2>&1
echo "gpu tests rc=$?" >> $OUT/gpu_tests.log

: > $OUT/batch_sweep.log
for B in 256 1024 4096 16384 65536; do
  timeout 120 python scripts/batch_sweep.py $B >> $OUT/batch_sweep.log 2>&1 || { echo "B=$B FAILED rc=$?" >> $OUT/batch_sweep.log; break; }
done

nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o scripts/micro/cluster_layer scripts/micro/cluster_layer.cu > $OUT/cluster_layer.txt 2>&1 \
  && timeout 60 scripts/micro/cluster_layer >> $OUT/cluster_layer.txt 2>&1
echo "cluster_layer rc=$?" >> $OUT/cluster_layer.txt

timeout 600 python bench.py > $OUT/bench_1gpu.json 2> $OUT/bench_1gpu.err
echo "bench rc=$?" >> $OUT/bench_1gpu.err
tail -3 $OUT/pending_vracer_da1.log $OUT/pending_vracer_explore.log $OUT/pending_vracer_b1024.log $OUT/pending_full_size.log $OUT/gpu_tests.log $OUT/batch_sweep.log $OUT/cluster_layer.txt
