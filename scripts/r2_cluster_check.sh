#!/bin/bash
# First runs of the cluster step kernel: one small golden case in a bounded child, then the parity file, then timing.
set -u
OUT=gpurun_out/r2
mkdir -p $OUT
SMB200_CLUSTER=1 SMB200_DEBUG=1 timeout -s KILL 120 python -u - > $OUT/cluster_first.log 2>&1 <<'PY'
import os, sys
root = os.getcwd()
sys.path[:0] = [root, os.path.join(root, "oracle"), os.path.join(root, "tests")]
import numpy as np
from parity_utils import Golden, make_learner, relerr
for case in ("vracer_small", "vracer_cfg2mini"):
    g = Golden(case)
    print("create", case, flush=True)
    L = make_learner(g)
    print("created", flush=True)
    R = g.ref
    for s in range(g.steps):
        st = L.train_steps(1)[0]
        O, gg, X = L.get_last_batch()
        pre = f"s{s}"
        print(case, s, "X", np.array_equal(X, R[pre + "/S"]), "O", np.abs(O - R[pre + "/O"]).max(), "g", relerr(gg, R[pre + "/g"]),
              "G", relerr(L.get_grad(), R[pre + "/gradSum"]) if pre + "/gradSum" in R else None,
              "W", np.abs(L.get_weights() - R[pre + "/weights"]).max() if pre + "/weights" in R else None,
              "nfar", st["n_far_policy"], int(g.refer(pre + "/post")[3]), flush=True)
    L.close()
print("ok")
PY
echo "first rc=$?" >> $OUT/cluster_first.log
tail -25 $OUT/cluster_first.log
if grep -q "^ok" $OUT/cluster_first.log; then
  timeout -s KILL 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_size.py -x -q --timeout 150 > $OUT/cluster_parity.log 2>&1
  echo "parity rc=$?" >> $OUT/cluster_parity.log
  tail -15 $OUT/cluster_parity.log
  SMB200_CLUSTER=1 timeout -s KILL 300 python scripts/batch_sweep.py 256 > $OUT/cluster_b256.log 2>&1; tail -3 $OUT/cluster_b256.log
  SMB200_CLUSTER=1 timeout -s KILL 300 python scripts/phase_report_cluster.py > $OUT/cluster_phases.log 2>&1; cat $OUT/cluster_phases.log
  timeout -s KILL 300 python scripts/batch_sweep.py 256 > $OUT/tile_b256.log 2>&1; tail -3 $OUT/tile_b256.log
fi
