"""Instruction-class counts per kernel of the in-tree library (cuobjdump -sass): what proves the Blackwell-native paths
(UTC*MMA = tcgen05.mma, LDTM = tcgen05.ld, UBLKCP = cp.async.bulk, LDGSTS = cp.async, cluster / DSMEM instructions) and what the
latency-bound kernels spend their instructions on.   usage: python scripts/sass_summary.py [lib.so] > profiles/r2/sass_summary.txt"""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "smarties_b200/libsmarties_b200.so"
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
CLASSES = [("tcgen05.mma (UTC*MMA)", r"^UTC.*MMA"), ("tcgen05.ld/st (LDTM/STTM)", r"^(LDTM|STTM)"), ("tcgen05 alloc/commit (UTC*)", r"^UTC(?!.*MMA)"),
           ("cp.async.bulk (UBLKCP)", r"^UBLKCP"), ("TMA tensor (UTMALDG/UTMASTG)", r"^UTMA"), ("cp.async (LDGSTS)", r"^LDGSTS"),
           ("mbarrier (SYNCS)", r"^SYNCS"), ("cluster barrier (UCGABAR/CGABAR)", r"CGABAR|^BAR.*CLUSTER"), ("DSMEM store (ST.* cluster / STAS)", r"^STAS|^ST\.E.*\.CLUSTER"),
           ("FFMA", r"^FFMA"), ("FADD/FMUL", r"^(FADD|FMUL)"), ("DFMA/DADD/DMUL (f64)", r"^(DFMA|DADD|DMUL)"), ("MUFU", r"^MUFU"),
           ("IMAD/IADD3/LEA/LOP3/SHF (integer)", r"^(IMAD|IADD3|LEA|LOP3|SHF|VIADD|IABS)"), ("ISETP/FSETP/DSETP", r"^[IFD]SETP"),
           ("LDS", r"^LDS"), ("STS", r"^STS"), ("LDG", r"^LDG(?!STS)"), ("STG", r"^STG"), ("LDL/STL (local)", r"^(LDL|STL)"),
           ("SHFL", r"^SHFL"), ("BAR.SYNC", r"^BAR"), ("ATOM/RED", r"^(ATOM|RED|ATOMG)"), ("BRA/BSSY/BSYNC", r"^(BRA|BSSY|BSYNC)"), ("HMMA (legacy mma.sync)", r"^HMMA")]
cur, counts = None, collections.OrderedDict()
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1); counts[cur] = collections.Counter(); continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        op = m.group(1)
        counts[cur]["total"] += 1
        for name, pat in CLASSES:
            if re.search(pat, op):
                counts[cur][name] += 1
                break
print(f"cuobjdump -sass {lib}: instruction classes per kernel (static counts)\n")
for fn, c in counts.items():
    dem = subprocess.run(["cu++filt", fn], capture_output=True, text=True).stdout.strip() or fn
    print(f"== {dem[:110]}   [{c['total']} instructions]")
    print("   " + ", ".join(f"{k} {v}" for k, v in c.items() if k != "total"))
