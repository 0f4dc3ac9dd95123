#!/bin/bash
# End-of-round evidence on one GPU: the whole GPU suite, the bench lines (driver arguments, reference arm, default, cfg3),
# the ncu launch list of the bench command and one --set full capture of the step kernel.
set -u
OUT=gpurun_out/r2f; mkdir -p $OUT
( time timeout -s KILL 2400 python -m pytest tests -q -m gpu --timeout 600 -rs > $OUT/gpu_tests_full_suite.log 2>&1 ) 2>&1 | grep real
echo "rc=$?" >> $OUT/gpu_tests_full_suite.log; tail -8 $OUT/gpu_tests_full_suite.log
( time python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $OUT/bench_ref_driver.json 2> $OUT/bench_ref_driver.err ) 2>&1 | grep real
( time python bench.py --gpus 1 --steps 20 --warmup 5 > $OUT/bench_driver.json 2> $OUT/bench_driver.err ) 2>&1 | grep real
( time python bench.py --workload cfg3 --steps 2000 --warmup 100 > $OUT/bench_cfg3.json 2> $OUT/bench_cfg3.err ) 2>&1 | grep real
( time python bench.py > $OUT/bench_default.json 2> $OUT/bench_default.err ) 2>&1 | grep real
python - <<'PY'
import json
for f in ("bench_ref_driver", "bench_driver", "bench_cfg3", "bench_default"):
    try:
        d = json.load(open(f"gpurun_out/r2f/{f}.json"))
        print(f, "value %.3e e2e %.3e ms/step %.5f" % (d["value"], d["e2e"]["value"], d["ms_per_step"]), "roof", d.get("roofline", {}).get("frac"), "cpu", d.get("cpu_baseline", {}).get("value"),
              "sweep", [(x["batch"], round(x["us_per_step"], 1)) for x in d.get("roofline_batch_sweep", [])])
    except Exception as e:
        print(f, "ERR", e)
PY
M="gpu__time_duration.sum"
ncu --metrics $M --clock-control none -c 4000 --csv --log-file $OUT/launches_bench.csv \
    python bench.py --steps 2000 --warmup 100 --no-cpu-baseline --no-batch-sweep > $OUT/bench_under_ncu.log 2>&1
PROF_RANGE=1 PROF_STEPS=256 PROF_SWEEPS=1 timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:'k_steps_persistent|k_sweep|k_moments' -o $OUT/full_steps -f python scripts/profile_steps.py > $OUT/ncu_full.log 2>&1 || true
tail -3 $OUT/ncu_full.log
ncu -i $OUT/full_steps.ncu-rep --page raw --csv > $OUT/full_steps_raw.csv 2>/dev/null
ls -la $OUT
