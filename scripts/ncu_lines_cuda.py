"""Per-CUDA-source-line stall samples of an ncu report (--import-source on): ncu -i rep --page source --print-source sass,cuda --csv"""
import csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "sass,cuda", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
cur = None; hdr = None; out = []
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1].split("/")[-1]; hdr = None; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; continue
    if hdr and len(r) == len(hdr):
        i = hdr.index("# Samples")
        if r[i].isdigit() and int(r[i]) > 0:
            d = dict(zip(hdr[2:], r[2:]))
            st = {k[6:]: int(v) for k, v in d.items() if k.startswith("stall_") and "Not Issued" not in k and v.isdigit() and int(v) > 0}
            out.append((int(r[i]), cur, r[0], r[1].strip()[:100], dict(sorted(st.items(), key=lambda kv: -kv[1])[:3])))
tot = sum(o[0] for o in out)
print("total samples", tot)
for o in sorted(out, key=lambda o: -o[0])[:top]:
    print(f"{o[0]:6d} {100*o[0]/tot:5.1f}%  {o[1]}:{o[2]:>5}  {o[3]}   {o[4]}")
