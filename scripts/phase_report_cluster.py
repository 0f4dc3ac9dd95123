"""Phase timing of the CLUSTER step kernel (cluster_step.cuh) on the cfg2 workload: clock64 markers of thread 0 of every CTA
(profiling flavour of the library), SM cycles -> us at SM_MHZ."""
import os
import sys

os.environ.setdefault("SMB200_PROFILE", "1")
import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
from smarties_b200 import Learner, synth  # noqa: E402

n = int(os.environ.get("PROF_STEPS", "200"))
B = int(os.environ.get("PROF_BATCH", "256"))
mhz = float(os.environ.get("SM_MHZ", "1965"))
d = synth.make_replay(123, 1000, 1000, 32, 8)
L = Learner(32, 8, {"maxTotObsNum": 1048576, "minTotObsNum": 1000000, "batchSize": B})
L.load_replay(d)
L.initialize_learner()
L.seed_sampler(7)
L.train_steps(1, want_stats=False)
L.presample(n)
L.train_presampled(0, 50); L.sync()
T, ms = L.profile_phases(n)
print(f"{n} steps in {ms:.3f} ms -> {1e3 * ms / n:.2f} us/step, grid {T.shape[1]}")
T = T[20:]
nP1 = min(T.shape[1] - 4, 4 * ((B + 7) // 8))
p1 = T[:, :nP1, :].astype(np.float64)
us = lambda c: c / mhz
seg = [("step start -> image in smem", 0, 1), ("gather + standardise", 1, 2), ("fwd L1 compute", 2, 25), ("fwd L1 cluster sync", 25, 9),
       ("fwd L2 compute", 9, 26), ("fwd L2 cluster sync", 26, 10), ("fwd out layer", 10, 12), ("loss stage 1", 12, 8),
       ("loss stage 2", 8, 16), ("loss stage 3", 16, 3), ("bwd out + broadcast", 3, 19), ("cluster sync", 19, 20),
       ("delta top slice", 20, 21), ("dX partial", 21, 28), ("cluster sync", 28, 22), ("delta lower slice", 22, 23),
       ("weight gradient", 23, 4), ("partial -> global", 4, 5), ("barrier 1 wait", 5, 6), ("prefetch issue", 6, 24),
       ("P2 (sum, Adam, images)", 24, 7)]
tot = 0.0
for name, a, b in seg:
    dlt = p1[:, :, b] - p1[:, :, a]
    tot += us(dlt.mean())
    print(f"  {name:30s} mean {us(dlt.mean()):6.2f} us   max-over-CTAs {us(dlt.max(axis=1).mean()):6.2f}")
per = (T[1:, 0, 0] - T[:-1, 0, 0]) / mhz
print(f"  sum of the means {tot:.2f} us; barrier 2 wait (P2 end -> next step start) {us((p1[1:, :, 0] - p1[:-1, :, 7]).mean()):.2f} us; step period {per.mean():.2f} us")
h = T[:, nP1:nP1 + 3, :].astype(np.float64)
print(f"  helper CTAs: start -> barrier 1 passed {us((h[:, :, 6] - h[:, :, 0]).mean()):.2f} us, P2 {us((h[:, :, 7] - h[:, :, 24]).mean()):.2f} us")
L.close()
