#!/bin/bash
# per-kernel durations of the wide step (ncu launch list; serialised, cold-cache: shares, not absolutes)
mkdir -p gpurun_out/r2w
for B in 2048 65536; do
  SWEEP_STEPS=3 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name regex:k_wide -c 200 --csv --log-file gpurun_out/r2w/launches_wide_B$B.csv \
    python scripts/batch_sweep.py $B > gpurun_out/r2w/ncu_B$B.log 2>&1
  python - <<PY
import csv, collections
rows = list(csv.reader(open("gpurun_out/r2w/launches_wide_B$B.csv")))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
agg = collections.OrderedDict()
for r in rows[hdr + 1:]:
    d = dict(zip(rows[hdr], r))
    if d.get("Metric Name") != "gpu__time_duration.sum": continue
    k = d["Kernel Name"][:60]
    v = float(d["Metric Value"].replace(",", ""))
    u = d["Metric Unit"]
    v = v / 1e3 if u in ("ns", "nsecond") else v
    agg.setdefault(k, []).append(v)
print("B=$B")
for k, v in agg.items():
    print(f"  {k:60s} n={len(v):3d} last={v[-1]:9.2f} us  median={sorted(v)[len(v)//2]:9.2f} us")
PY
done
