"""Small driver for ncu: build the cfg2 workload (or a reduced one), run a few learner steps."""
import os
import sys
import time

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
from smarties_b200 import Learner, synth  # noqa: E402

n_ep = int(os.environ.get("PROF_NEP", "1000"))
steps = int(os.environ.get("PROF_STEPS", "20"))
d = synth.make_replay(123, n_ep, 1000, 32, 8)
L = Learner(32, 8, {"maxTotObsNum": 1048576, "minTotObsNum": 1000 * n_ep})
L.load_replay(d)
L.initialize_learner()
L.seed_sampler(7)
L.train_steps(1, want_stats=False)
start = int(os.environ.get("PROF_START", "0"))
if start:
    L.set_grad_step(start)
L.presample(steps)
L.train_presampled(0, min(steps, 8)); L.sync()      # warm-up launch (not captured)
L.presample(steps)
if os.environ.get("PROF_RANGE"):                    # ncu --profile-from-start off: capture only what follows
    import torch
    torch.cuda.profiler.start()
t0 = time.perf_counter()
L.train_presampled(0, steps)
L.sync()
print("steps", steps, "device ms", L.last_timing(), "wall", time.perf_counter() - t0)
if os.environ.get("PROF_SWEEPS"):
    L.retrace_sweep(); print("retrace ms", L.last_timing())
    L.reward_state_moments(); print("moments ms", L.last_timing())
    L.fused_sweep(); print("fused sweep ms", L.last_timing())
if os.environ.get("PROF_RANGE"):
    torch.cuda.profiler.stop()
L.close()
