"""Build the CUDA library IN-TREE for sm_100a: smarties_b200/libsmarties_b200.so.

nvcc cross-compiles without a GPU.  The shared object is git-ignored but travels with the
gpurun snapshot.  -fmad=false: fused multiply-adds are written fmaf() explicitly in the
kernels; everything else must round like the reference's x86-64 (no-FMA) build.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libsmarties_b200.so")
OUT_PROF = os.path.join(HERE, "libsmarties_b200_prof.so")     # same sources + phase timestamps (scripts/phase_report.py)
SOURCES = ["step_kernels.cu", "sweep_kernels.cu", "learner.cu"]
HEADERS = ["common.cuh", "step_kernels.cuh", "cluster_step.cuh", "wide_step.cuh", os.path.join("..", "..", "include", "smarties_b200.h")]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-fmad=false",
         "-Xcompiler", "-fPIC", "-Xptxas", "-v", "--threads", "3"]


def _stale() -> bool:
    if not os.path.exists(OUT) or not os.path.exists(OUT_PROF):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def _compile(out: str, suffix: str, defines, logname: str, verbose: bool) -> None:
    objs = []
    procs = []
    for s in SOURCES:
        o = os.path.join(CSRC, s.replace(".cu", suffix + ".o"))
        cmd = [NVCC, *FLAGS, *defines, "-c", os.path.join(CSRC, s), "-o", o]
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(o)
    log = []
    for s, p in procs:
        o, _ = p.communicate()
        log.append(o)
        if p.returncode != 0:
            sys.stderr.write(o)
            raise RuntimeError(f"nvcc failed on {s}")
    subprocess.run([NVCC, "-shared", "-o", out, *objs, "-cudart", "static"], check=True)
    # the tracked log keeps registers / spills / shared memory per kernel; compile times would change it on every build
    kept = [l for l in "\n".join(log).splitlines() if "Compile time" not in l]
    with open(os.path.join(CSRC, logname), "w") as f:
        f.write("\n".join(kept) + "\n")
    if verbose:
        print("\n".join(log))


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not _stale():
        return OUT
    extra = os.environ.get("SMB200_PROF_DEFINES", "").split()
    _compile(OUT_PROF, "_prof", ["-DSMB200_MARKERS", *extra], "ptxas_prof.log", False)
    _compile(OUT, "", [], "ptxas.log", verbose)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
