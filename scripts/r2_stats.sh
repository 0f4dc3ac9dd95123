#!/bin/bash
# launch-resident statistics: invariance test, golden step tests, and the 8 M-transition buffer on one GPU (8000 episodes:
# the statistics CTA used to bound the step) with and without it
set -u
OUT=gpurun_out/r2s; mkdir -p $OUT
timeout -s KILL 900 python -m pytest tests/test_gpu_parity.py -x -q --timeout 400 -k "launch_resident or learner_steps_match or sample_ahead or persistent" > $OUT/stats_tests.log 2>&1
echo "rc=$?" >> $OUT/stats_tests.log; tail -6 $OUT/stats_tests.log
for FULL in 1 0; do
  SMB200_STATS_FULL=$FULL timeout -s KILL 600 python bench.py --gpus 1 --steps 4000 --warmup 200 --scaling strong --no-cpu-baseline --no-batch-sweep > $OUT/bench_strong_1gpu_full$FULL.json 2> $OUT/bench_strong_1gpu_full$FULL.err
  python - <<PY
import json
for l in open("$OUT/bench_strong_1gpu_full$FULL.json"):
    if l.startswith("{"):
        d = json.loads(l); print("strong 8M 1gpu STATS_FULL=$FULL:", round(d["value"]/1e6, 2), "e2e", round(d["e2e"]["value"]/1e6, 2), "us/step", round(d["ms_per_step"]*1e3, 2), d["final_stats"])
PY
done
for FULL in 1 0; do
  SMB200_STATS_FULL=$FULL python bench.py --gpus 1 --steps 4000 --warmup 200 --no-cpu-baseline --no-batch-sweep 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('weak 1M 1gpu STATS_FULL=$FULL:', round(d['value']/1e6,2), 'e2e', round(d['e2e']['value']/1e6,2), d['final_stats'])"
done
