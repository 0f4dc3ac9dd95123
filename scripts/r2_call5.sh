#!/bin/bash
set -u
OUT=gpurun_out/r2; mkdir -p $OUT
timeout -s KILL 1500 python -m pytest tests -m gpu -x -q --timeout 400 > $OUT/gpu_tests_e.log 2>&1; echo "rc=$?" >> $OUT/gpu_tests_e.log; tail -8 $OUT/gpu_tests_e.log
timeout -s KILL 600 python scripts/dropin_run.py --app synth_env --envs 64 --steps 200000 --threads 16 --arms b200 > $OUT/dropin_synth_env_64envs.jsonl 2>&1; tail -2 $OUT/dropin_synth_env_64envs.jsonl | cut -c1-700
