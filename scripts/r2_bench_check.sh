#!/bin/bash
set -u
OUT=gpurun_out/r2; mkdir -p $OUT
( time python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $OUT/bench_ref_driver.json 2> $OUT/bench_ref_driver.err ) 2>&1 | grep real
( time python bench.py --gpus 1 --steps 20 --warmup 5 > $OUT/bench_driver.json 2> $OUT/bench_driver.err ) 2>&1 | grep real
tail -3 $OUT/bench_driver.err
( time python bench.py --workload cfg3 --steps 2000 --warmup 100 > $OUT/bench_cfg3.json 2> $OUT/bench_cfg3.err ) 2>&1 | grep real
tail -3 $OUT/bench_cfg3.err
( time python bench.py > $OUT/bench_default.json 2> $OUT/bench_default.err ) 2>&1 | grep real
python - <<'PY'
import json
for f in ("bench_ref_driver", "bench_driver", "bench_cfg3", "bench_default"):
    try:
        d = json.load(open(f"gpurun_out/r2/{f}.json"))
        print(f, "value %.3e e2e %.3e ms/step %.5f" % (d["value"], d["e2e"]["value"], d["ms_per_step"]), "roof", d.get("roofline", {}).get("frac"), "cpu", d.get("cpu_baseline", {}).get("value"),
              "sweep", [(x["batch"], round(x["us_per_step"], 1)) for x in d.get("roofline_batch_sweep", [])])
    except Exception as e:
        print(f, "ERR", e)
PY
