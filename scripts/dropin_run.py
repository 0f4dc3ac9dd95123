"""Drop-in runs of whole smarties applications, once with the reference CPU learner and once with the
learner steps on the GPU through libsmarties_b200.so (integration/RACER_B200.cpp, selected by
SMARTIES_B200=1).  Same binary interface, same settings file, same command line.

  --app cart_pole   BASELINE.json configs[0]: the reference's apps/cart_pole_cpp + settings/VRACER.json
  --app py_env      a Python environment (integration/py_env.py) through the reference's pybind11 module
  --app c_env       a C environment on the C interface Fortran apps bind (integration/c_env.c, smarties_extern.h)
  --app synth_env   configs[3] shape: integration/synth_env.cpp (17 states, 6 bounded actions, 1000-step
                    truncated episodes) with --envs 64 forked environment processes feeding one learner

Needs the prebuilt files under oracle/_ref/ (integration/Makefile).  Reports wall-clock of the whole run and
the environment time steps / gradient steps per second after the initial data collection.

usage: python scripts/dropin_run.py [--app cart_pole] [--envs 1] [--steps 20000] [--threads 8] [--arms ref,b200]"""
import argparse
import json
import os
import shutil
import subprocess
import sys
import tempfile
import time

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
SETTINGS = {"learner": "VRACER", "dataSamplingAlgo": "uniform", "returnsEstimator": "retrace", "ERoldSeqFilter": "oldest",
            "nnLayerSizes": [128, 128]}     # == settings/VRACER.json of the reference


def run_arm(arm, steps, threads, seed, settings=None, timeout=1500, app="cart_pole", envs=1, extra_env=None,
            keep_dir=None, restart=None):
    """keep_dir: run there and keep the files (checkpoints); restart: directory of an earlier run (--restart)."""
    exe = os.path.join(ROOT, "oracle", "_ref", "b200" if arm == "b200" else "", app)
    cmd0 = [exe]
    pyenv = {}
    if app == "py_env":      # a Python app through the reference's pybind11 module
        pydir = os.path.join(ROOT, "oracle", "_ref", "b200" if arm == "b200" else "", "py")
        exe = pydir
        cmd0 = [sys.executable, os.path.join(ROOT, "integration", "py_env.py")]
        pyenv = {"PYTHONPATH": pydir}
    if not os.path.exists(exe):
        return {"arm": arm, "error": f"{exe} missing (make -C integration)"}
    tmp = keep_dir or tempfile.mkdtemp(prefix=f"cartpole_{arm}_")
    os.makedirs(tmp, exist_ok=True)
    with open(os.path.join(tmp, "settings.json"), "w") as f:
        json.dump(settings or SETTINGS, f)
    env = dict(os.environ)
    env.pop("SMARTIES_B200", None)
    if arm == "b200":
        env["SMARTIES_B200"] = "1"
    env.update(pyenv)
    env.update(extra_env or {})
    t0 = time.perf_counter()
    p = subprocess.run(cmd0 + ["--nTrainSteps", str(steps), "--nThreads", str(threads), "--randSeed", str(seed),
                        "--nEnvironments", str(envs)] + (["--restart", restart] if restart else []),
                       cwd=tmp, env=env, capture_output=True, text=True, timeout=timeout)
    wall = time.perf_counter() - t0
    out = {"arm": arm, "app": app, "envs": envs, "host_threads": threads, "rc": p.returncode, "wall_s": round(wall, 2), "steps": steps}
    rows = []
    sp = os.path.join(tmp, "agent_00_stats.txt")
    if os.path.exists(sp):
        for l in open(sp):
            f = l.split()
            if len(f) > 3 and f[0].isdigit():
                rows.append(f)
    if rows:   # columns: ID #/T avgR avgr stdr DKL RMSE ... nFarP beta net
        out["stat_rows"] = len(rows)
        out["grad_steps_logged"] = 1000 * int(rows[-1][1])      # one row per 1000 gradient steps
        out["avgR_first"], out["avgR_last"] = float(rows[0][2]), float(rows[-1][2])
        out["avgR_max"] = max(float(r[2]) for r in rows)
        out["beta_last"] = float(rows[-1][-2])
    out["b200_lines"] = [l for l in p.stdout.splitlines() if l.startswith("smarties_b200")]
    out["restart_lines"] = sorted(set(l for l in p.stdout.splitlines() if l.startswith("Restarting from file")))
    if p.returncode != 0:
        out["tail"] = (p.stdout[-1500:] + p.stderr[-1500:])
    if not keep_dir:
        shutil.rmtree(tmp, ignore_errors=True)
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=20000)
    ap.add_argument("--threads", type=int, default=min(8, os.cpu_count() or 1))
    ap.add_argument("--seed", type=int, default=7)
    ap.add_argument("--arms", default="ref,b200")
    ap.add_argument("--app", default="cart_pole", choices=["cart_pole", "synth_env", "py_env", "c_env"])
    ap.add_argument("--envs", type=int, default=1)
    ap.add_argument("--max-steps-per-call", type=int, default=0, help="SMARTIES_B200_MAXSTEPS (0 = binding default)")
    a = ap.parse_args()
    xe = {"SMARTIES_B200_MAXSTEPS": str(a.max_steps_per_call)} if a.max_steps_per_call else None
    for arm in a.arms.split(","):
        print(json.dumps(run_arm(arm, a.steps, a.threads, a.seed, app=a.app, envs=a.envs, extra_env=xe)), flush=True)
