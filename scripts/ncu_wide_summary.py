"""Summarise the ncu --set full captures of the wide step kernels (gpurun_out/r2w/wide_{fwd,bwd,wgrad}.ncu-rep, one launch each at
B = 65536, scripts/r2_wide_profiles.sh) into profiles/r2/ncu_wide.json (read by bench.py -> roofline_wide_kernels) and
profiles/r2/ncu_wide_summary.txt."""
import csv, json, os, subprocess, sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
src = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "r2w")
dst = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "profiles", "r2")
B = int(sys.argv[3]) if len(sys.argv) > 3 else 65536
KEYS = {
    "gpu__time_duration.sum": "duration_us",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed": "tensor_pipe_pct_of_peak_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_pct_of_peak_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "dram__bytes_read.sum": "dram_read", "dram__bytes_write.sum": "dram_write",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
    "launch__grid_size": "grid", "launch__block_size": "block", "launch__registers_per_thread": "registers", "smsp__inst_executed.sum": "warp_instructions",
    "launch__shared_mem_per_block_dynamic": "dynamic_smem_bytes",
}
out = {"_batch": B, "_source": "ncu --set full --clock-control none, one launch per kernel (scripts/r2_wide_profiles.sh)"}
lines = []
for k in ("fwd", "loss", "bwd", "wgrad"):
    rep = os.path.join(src, f"wide_{k}.ncu-rep")
    if not os.path.exists(rep):
        continue
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = dict(zip(hdr, vals)); u = dict(zip(hdr, units))
    e = {"kernel": d.get("Kernel Name")}
    for key, name in KEYS.items():
        if key in d:
            v = float(d[key].replace(",", ""))
            if name.startswith("dram_r") or name.startswith("dram_w"):
                scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(u[key], 1.0)
                v *= scale
            if name == "duration_us":
                v *= {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3}.get(u[key], 1.0)
            e[name] = v
    st = {h.replace("smsp__pcsamp_warps_issue_stalled_", ""): float(d[h]) for h in hdr
          if h.startswith("smsp__pcsamp_warps_issue_stalled_") and "not_issued" not in h}
    tot = sum(st.values()) or 1.0
    e["stall_share"] = {a: round(b / tot, 3) for a, b in sorted(st.items(), key=lambda kv: -kv[1])[:6]}
    e["dram_bytes"] = e.get("dram_read", 0.0) + e.get("dram_write", 0.0)
    out["k_wide_" + k] = e
    lines.append(f"k_wide_{k}: {e.get('duration_us', 0):.1f} us, grid {int(e.get('grid', 0))} x {int(e.get('block', 0))} threads, {int(e.get('registers', 0))} registers, "
                 f"tensor pipe {e.get('tensor_pipe_pct_of_peak_elapsed', 0):.1f} % of peak (elapsed), issue active {e.get('issue_active_pct', 0):.1f} %, "
                 f"DRAM {e['dram_bytes'] / 1e6:.1f} MB ({e.get('dram_throughput_pct', 0):.1f} % of peak), stalls {e['stall_share']}")
os.makedirs(dst, exist_ok=True)
with open(os.path.join(dst, "ncu_wide.json"), "w") as f:
    json.dump(out, f, indent=1)
with open(os.path.join(dst, "ncu_wide_summary.txt"), "w") as f:
    f.write(f"wide step kernels at B = {B} (cfg2 network, 1 M-transition buffer), ncu --set full --clock-control none, one launch each\n")
    f.write("\n".join(lines) + "\n")
print("\n".join(lines))
