"""Stall samples of an ncu report aggregated over line ranges of one source file: ncu_phase_cuda.py rep file.cuh name:lo-hi ..."""
import csv, subprocess, sys
rep, fname = sys.argv[1], sys.argv[2]
phases = []
for a in sys.argv[3:]:
    n, r = a.split(":"); lo, hi = r.split("-"); phases.append((n, int(lo), int(hi)))
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "sass,cuda", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
cur = None; hdr = None; agg = {}; other = 0; tot = 0
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1].split("/")[-1]; hdr = None; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; continue
    if hdr and len(r) == len(hdr) and r[0].isdigit():      # CUDA source rows only (SASS rows have no line number)
        i = hdr.index("# Samples")
        if not r[i].isdigit(): continue
        n = int(r[i]); tot += n
        d = dict(zip(hdr[2:], r[2:]))
        st = {k[6:]: int(v) for k, v in d.items() if k.startswith("stall_") and "Not Issued" not in k and v.isdigit() and int(v) > 0}
        key = "other:" + cur
        if cur == fname:
            for nme, lo, hi in phases:
                if lo <= int(r[0]) <= hi: key = nme; break
        a = agg.setdefault(key, [0, {}]); a[0] += n
        for k, v in st.items(): a[1][k] = a[1].get(k, 0) + v
print("total", tot)
for k, (n, st) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"{n:7d} {100*n/tot:5.1f}%  {k:28s} {dict(sorted(st.items(), key=lambda kv: -kv[1])[:4])}")
