"""cfg3 probe (BASELINE.json configs[2]): RACER + LSTM(64), nnBPTTseq 32, batch 128 on the 1M-transition
synthetic buffer: device time per learner step."""
import os
import sys
import time

if os.environ.get("PROF_PHASES"):
    os.environ.setdefault("SMB200_PROFILE", "1")

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
from smarties_b200 import Learner, synth  # noqa: E402

n_ep = int(os.environ.get("PROF_NEP", "1000"))
steps = int(os.environ.get("PROF_STEPS", "200"))
S = {"learner": "RACER", "nnType": "LSTM", "nnLayerSizes": [64], "nnBPTTseq": 32, "batchSize": 128, "clipImpWeight": 4,
     "explNoise": 0.1, "gamma": 0.99, "epsAnneal": 0, "nnLambda": 1e-6, "maxTotObsNum": 1048576, "minTotObsNum": 1000 * n_ep}
d = synth.make_replay(123, n_ep, 1000, 32, 8)
L = Learner(32, 8, S)
L.load_replay(d)
L.initialize_learner()
L.seed_sampler(7)
L.train_steps(1, want_stats=False)
L.presample(steps + 20)
L.train_presampled(0, 20); L.sync()
if os.environ.get("PROF_RANGE"):                    # ncu --profile-from-start off: capture only the timed launch
    import torch
    torch.cuda.profiler.start()
t0 = time.perf_counter()
L.train_presampled(20, steps)
L.sync()
if os.environ.get("PROF_RANGE"):
    torch.cuda.profiler.stop()
ms, nl = L.last_timing()
print(f"cfg3: {steps} steps, device {ms:.3f} ms -> {1e3 * ms / steps:.1f} us/step, {128 * steps / (ms * 1e-3):.3e} transitions/s, launches {nl}, wall {time.perf_counter() - t0:.3f}s")
print("stats", L.get_stats())
if os.environ.get("PROF_PHASES"):
    import numpy as np
    n = 50
    L.presample(n)
    T, ms = L.profile_phases(n)
    T = T[10:]
    mhz = 1965.0
    w = T[:, :128, :]
    for name, a, b in (("gather", 0, 1), ("forward", 1, 2), ("loss", 2, 3), ("backward+store", 3, 4), ("P1 total", 0, 5), ("barrier1 wait", 5, 6), ("P2", 6, 7)):
        dlt = (w[:, :, b] - w[:, :, a]) / mhz
        print(f"  {name:16s} mean {dlt.mean():7.2f} us  max-over-CTAs {dlt.max(axis=1).mean():7.2f}")
    for name, a, b in (("fwd: wait for the weight image", 1, 9), ("fwd: LSTM layer 1 (input product + recurrence)", 9, 10), ("fwd: rest (output layer)", 10, 2)):
        dlt = (w[:, :, b] - w[:, :, a]) / mhz
        print(f"  {name:48s} median {np.median(dlt):7.2f} us")
    for name, a, b in (("TC staging", 42, 43), ("TC mma issue+wait", 43, 44), ("TC epilogue", 44, 45), ("TC barrier", 45, 46), ("tiles (reduce, exchange, Adam)", 46, 7)):
        dlt = (w[:, :100, b] - w[:, :100, a]) / mhz
        print(f"  {name:32s} median {np.median(dlt):7.2f} us")
    per = (T[1:, 0, 0] - T[:-1, 0, 0]) / mhz
    print(f"  step period {per.mean():.2f} us")
L.close()
if os.environ.get("CFG3_REF_STEPS"):     # the reference's CPU learner on the same workload, all host cores
    import bench
    n = int(os.environ["CFG3_REF_STEPS"])
    r = bench.run_reference(n, os.cpu_count() or 1, data=d, settings={k: v for k, v in S.items()})
    print(f"cfg3 reference CPU: {r['value']:.3e} transitions/s ({r['cores']} threads, {n} steps, {1e3 * r['seconds'] / n:.2f} ms/step)")
