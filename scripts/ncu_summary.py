"""Turn the outputs of scripts/ncu_capture.sh (gpurun_out/) into the summaries kept under profiles/<round>/:
  launches_bench_summary.csv   per-kernel launches / total / share of the default bench command
  ncu_full_summary.txt         selected --set full metrics per captured kernel
  ncu_traffic.json             dram bytes per launch of the captured kernels (bench.py reports roofline.traffic from it)
usage: python scripts/ncu_summary.py gpurun_out profiles/r1"""
import csv
import json
import sys
from collections import OrderedDict

src, dst = sys.argv[1], sys.argv[2]
import os
os.makedirs(dst, exist_ok=True)

# ---- launch list ----
rows = list(csv.reader(l for l in open(f"{src}/launches_bench.csv") if not l.startswith("==")))
hdr = rows[0]
iK, iV, iU = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = OrderedDict()
for r in rows[1:]:
    if len(r) <= iV:
        continue
    v = float(r[iV].replace(",", ""))
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[iU], 1.0)   # -> us
    a = agg.setdefault(r[iK], [0, 0.0])
    a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
with open(f"{dst}/launches_bench_summary.csv", "w") as f:
    f.write("# ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 python bench.py --steps 2000 --warmup 100 --no-cpu-baseline\n")
    f.write("# whole process (incl. buffer construction: 1000x k_init_episode + 1000x single-episode k_sweep); times are cold-cache and serialised\n")
    f.write("kernel,launches,total_us,share_pct,mean_us\n")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"\"{k}\",{n},{t:.1f},{100 * t / tot:.2f},{t / n:.2f}\n")

# ---- full capture ----
rows = list(csv.reader(open(f"{src}/full_steps_raw.csv")))
hdr, units = rows[0], rows[1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum", "sm__inst_executed_pipe_fma.sum", "smsp__cycles_active.avg"]
iK = hdr.index("Kernel Name")
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
traffic = {}
with open(f"{dst}/ncu_full_summary.txt", "w") as f:
    f.write("ncu --set full --clock-control none --import-source on --profile-from-start off  (B200; scripts/ncu_capture.sh)\n")
    f.write("persistent kernel: 256 learner steps in one launch (scripts/profile_steps.py, cfg2 workload); sweeps: 1M-transition buffer\n")
    for r in rows[2:]:
        name = r[iK]
        f.write(f"\n== {name}\n")
        rd = wr = 0.0
        for m in want:
            if m in hdr:
                i = hdr.index(m)
                f.write(f"   {m:70s} {r[i]} {units[i]}\n")
                if m == "dram__bytes_read.sum":
                    rd = float(r[i].replace(",", "")) * scale.get(units[i], 1.0)
                if m == "dram__bytes_write.sum":
                    wr = float(r[i].replace(",", "")) * scale.get(units[i], 1.0)
        short = name.split("(")[0].split("<")[0].split("::")[-1].strip()
        traffic[short] = {"dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes": rd + wr}
traffic["_note"] = "per launch; k_steps_persistent launch = 256 learner steps (PROF_STEPS=256)"
traffic["_steps_per_persistent_launch"] = 256
json.dump(traffic, open(f"{dst}/ncu_traffic.json", "w"), indent=1)
print(json.dumps(traffic, indent=1))
