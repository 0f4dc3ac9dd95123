#!/bin/bash
# weak scaling of the wide step: B = 65536 per GPU, 1 M-transition shard per GPU ($1 = list of rank counts, default "1 2")
mkdir -p gpurun_out/r2w
for N in ${1:-1 2}; do
  if [ "$N" = "1" ]; then
    timeout 300 python bench.py --batch 65536 --steps 100 --warmup 5 --no-cpu-baseline --no-batch-sweep > gpurun_out/r2w/bench_wide_weak_${N}gpu.json 2> gpurun_out/r2w/bench_wide_weak_${N}gpu.err
  else
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29700 + N)) bench.py --gpus $N \
      --batch 65536 --steps 100 --warmup 5 --no-cpu-baseline --no-batch-sweep > gpurun_out/r2w/bench_wide_weak_${N}gpu.json 2> gpurun_out/r2w/bench_wide_weak_${N}gpu.err
  fi
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2w/bench_wide_weak_${N}gpu.json").read().strip().splitlines()[-1])
    print("N=$N value %.4g e2e %.4g ms/step %.4f launches %s ranks_identical %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["gpu_launches"], d.get("ranks_identical")))
except Exception as e:
    print("N=$N failed", e); print(open("gpurun_out/r2w/bench_wide_weak_${N}gpu.err").read()[-1500:])
PY
done
