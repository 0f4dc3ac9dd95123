#!/bin/bash
# sample-ahead queue: invariance test, the goldens (one call per step: every step is served from the queue), e2e at the driver's arguments
set -u
OUT=gpurun_out/r2b
mkdir -p $OUT
timeout -s KILL 600 python -m pytest tests/test_gpu_parity.py -x -q --timeout 300 -k "sample_ahead or sampler or learner_steps_match" > $OUT/ahead_tests.log 2>&1
echo "rc=$?" >> $OUT/ahead_tests.log; tail -6 $OUT/ahead_tests.log
for i in 1 2; do
python bench.py --gpus 1 --steps 20 --warmup 5 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('ahead  20/5:', round(d['value']/1e6,2), 'e2e', round(d['e2e']['value']/1e6,2))"
SMB200_NO_SAMPLE_AHEAD=1 python bench.py --gpus 1 --steps 20 --warmup 5 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('no-ahead 20/5:', round(d['value']/1e6,2), 'e2e', round(d['e2e']['value']/1e6,2))"
done
python bench.py 2>/dev/null > $OUT/bench_default.json; python -c "
import json
for l in open('$OUT/bench_default.json'):
    if l.startswith('{'):
        d = json.loads(l); print('default:', d['steps'], round(d['value']/1e6,2), 'e2e', round(d['e2e']['value']/1e6,2))"
