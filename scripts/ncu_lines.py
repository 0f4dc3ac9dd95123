"""Summarise an ncu `--page source --print-source cuda,sass --csv` dump: stall samples per CUDA
source line.  usage: ncu_lines.py file.csv [topN]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = None
lines = []
fname = ""
for r in rows:
    if r and r[0] == "File Path":
        fname = r[1].split("/")[-1]
        continue
    if r and r[0] == "Line No":
        hdr = r
        continue
    if hdr and r and r[0].isdigit():
        try:
            lines.append((fname, int(r[0]), r[1], float(r[hdr.index("# Samples")] or 0), float(r[hdr.index("Instructions Executed")] or 0)))
        except ValueError:
            pass
tot = sum(l[3] for l in lines) or 1
print("total samples", tot)
for f, ln, src, smp, ins in sorted(lines, key=lambda x: -x[3])[:top]:
    print(f"{smp:7.0f} {100 * smp / tot:5.1f}%  inst {ins:9.0f}  {f}:{ln:<4d} {src.strip()[:110]}")
