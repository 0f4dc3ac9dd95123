"""Batch-size sweep of the cfg2 MLP learner step (B = 256 / 4096 / 65 536 to show the approach to the HBM
bound): device time per step and achieved algorithmic GB/s (268 B per sampled transition + 7*4*nParams of Adam traffic).

NOT YET RUN ON A GPU — written when the round's GPU budget was spent.  Nothing above B = 256 has run on the device so far:
start with a short watchdog, one batch size per process, e.g.

    for B in 256 1024 4096; do timeout 60 python scripts/batch_sweep.py $B || break; done
"""
import os
import sys
import time

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
from smarties_b200 import Learner, synth  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
steps = int(os.environ.get("SWEEP_STEPS", "200"))
n_ep = int(os.environ.get("SWEEP_NEP", "1000"))
S = {"learner": "VRACER", "nnLayerSizes": [128, 128], "batchSize": B, "maxTotObsNum": 1048576, "minTotObsNum": 1000 * n_ep}
d = synth.make_replay(123, n_ep, 1000, 32, 8)
L = Learner(32, 8, S)
L.load_replay(d)
L.initialize_learner()
L.seed_sampler(7)
L.train_steps(1, want_stats=False)
L.presample(steps + 20)
L.train_presampled(0, 20); L.sync()
t0 = time.perf_counter()
L.train_presampled(20, steps)
L.sync()
ms, nl = L.last_timing()
us = 1e3 * ms / steps
alg = 268 * B + 7 * 4 * L.n_params
print(f"B={B}: {steps} steps, {us:.2f} us/step, {B * steps / (ms * 1e-3):.3e} transitions/s, algorithmic {alg / 1e3:.0f} KB/step "
      f"-> {alg / (us * 1e-6) / 1e9:.1f} GB/s, launches {nl}, wall {time.perf_counter() - t0:.3f}s")
print("stats", L.get_stats())
L.close()
