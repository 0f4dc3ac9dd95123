#!/bin/bash
# Device actor inference: forward_seq parity, the three drop-in runs with SMARTIES_B200_ACTORS=1, and host-actor vs
# device-actor wall clock of a recurrent run (where every action costs a window forward pass).
set -u
OUT=gpurun_out/r2a
mkdir -p $OUT
timeout -s KILL 400 python -m pytest tests/test_gpu_parity.py -x -q --timeout 200 -k "forward" > $OUT/forward_parity.log 2>&1
echo "rc=$?" >> $OUT/forward_parity.log; tail -5 $OUT/forward_parity.log
timeout -s KILL 900 python -m pytest tests/test_gpu_dropin.py -x -q --timeout 400 -k "actors" > $OUT/dropin_actors.log 2>&1
echo "rc=$?" >> $OUT/dropin_actors.log; tail -15 $OUT/dropin_actors.log
timeout -s KILL 600 python - > $OUT/actors_compare.log 2>&1 <<'PY'
import os, sys, json
sys.path.insert(0, "scripts")
from dropin_run import run_arm
S = {"learner": "RACER", "nnType": "LSTM", "nnLayerSizes": [64], "nnBPTTseq": 32, "clipImpWeight": 4, "explNoise": 0.1, "gamma": 0.99,
     "epsAnneal": 0, "nnLambda": 1e-6, "maxTotObsNum": 16384, "minTotObsNum": 4096}
for name, env in (("host actors", {}), ("device actors", {"SMARTIES_B200_ACTORS": "1"})):
    r = run_arm("b200", steps=4000, threads=4, seed=7, settings=S, timeout=280, extra_env=env)
    print(name, json.dumps({k: r.get(k) for k in ("rc", "wall_s", "avgR_last", "beta_last", "b200_lines")}), flush=True)
for name, env in (("host actors", {}), ("device actors", {"SMARTIES_B200_ACTORS": "1"})):
    r = run_arm("b200", steps=20000, threads=4, seed=3, timeout=280, app="synth_env", envs=16, extra_env=env)
    print("synth16", name, json.dumps({k: r.get(k) for k in ("rc", "wall_s", "avgR_last", "beta_last", "b200_lines")}), flush=True)
PY
cat $OUT/actors_compare.log
