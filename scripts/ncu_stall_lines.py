"""Per-source-line view of an `ncu --set full --import-source on` capture of the persistent step kernel: where the warp-stall
samples sit (by CUDA source line, with the dominant stall reasons) and how many shared-memory wavefronts are bank-conflict
replays (loads / stores).  usage: python scripts/ncu_stall_lines.py gpurun_out/full_steps.ncu-rep > profiles/r2/ncu_step_kernel_stall_lines.txt"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/full_steps.ncu-rep"
kern = sys.argv[2] if len(sys.argv) > 2 else "k_steps_persistent"


def page(view):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", view, "--kernel-name", "regex:" + kern],
                         capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


rows = page("cuda,sass")
cur, hdr, lines = None, None, []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]; continue
    if r[0] == "Line No":
        hdr = r; continue
    if r[0].isdigit() and hdr:
        d = dict(zip(hdr, r))
        try:
            n = int(d["# Samples"])
        except (KeyError, ValueError):
            continue
        st = {k[6:]: int(v) for k, v in d.items() if k.startswith("stall_") and "Not Issued" not in k and v.isdigit() and int(v) > 0}
        lines.append((n, cur, int(r[0]), r[1].strip()[:90], st))
tot = sum(l[0] for l in lines)
by_reason = {}
for l in lines:
    for k, v in l[4].items():
        by_reason[k] = by_reason.get(k, 0) + v
print(f"{rep}: kernel {kern}, {tot} warp-stall samples")
print("stall reasons (share of all samples): " + ", ".join(f"{k} {100 * v / tot:.1f}%" for k, v in sorted(by_reason.items(), key=lambda kv: -kv[1])[:9]))
print("\ntop source lines (samples, share, file:line, dominant reasons, source):")
for n, f, ln, src, st in sorted(lines, reverse=True)[:28]:
    top = ", ".join(f"{k} {v}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:3])
    print(f"{n:8d} {100 * n / tot:5.1f}%  {f}:{ln:<5d} [{top}]  {src}")

rows = page("sass")
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
ld = [0, 0]; stv = [0, 0]
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    try:
        e, t = int(r[ix["L1 Wavefronts Shared Excessive"]]), int(r[ix["L1 Wavefronts Shared"]])
    except ValueError:
        continue
    op = r[ix["Source"]].strip()
    tgt = stv if (op.startswith("STS") or "LDGSTS" in op or op.startswith("ATOMS") or op.startswith("@") and " STS" in op) else ld
    tgt[0] += t; tgt[1] += e
print(f"\nshared-memory wavefronts: loads {ld[0]} ({100 * ld[1] / max(ld[0], 1):.1f}% bank-conflict replays), "
      f"stores {stv[0]} ({100 * stv[1] / max(stv[0], 1):.1f}% bank-conflict replays)")
