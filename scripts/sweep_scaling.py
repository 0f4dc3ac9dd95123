"""The every-1000-steps sweeps at growing buffer sizes (1M / 8M / 32M transitions, dS = 32): the fused one-pass kernel
(k_sweep_fused: 156 algorithmic bytes per transition) next to the separate Retrace and moments kernels; best of 5 with a
flushed L2, fraction of the measured HBM peak (MEASURED_PEAKS.json)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import bench  # noqa: E402
from smarties_b200 import Learner, synth  # noqa: E402

peaks, which = bench.measured_peaks()
dev = torch.device("cuda", 0)
out = []
for n_ep in [int(x) for x in os.environ.get("SWEEP_EPISODES", "1000,8000,32000").split(",")]:
    d = synth.make_replay(123, n_ep, 1000, 32, 8)
    rows = int(d["N"].sum())
    L = Learner(32, 8, {"maxTotObsNum": 1 << (rows - 1).bit_length(), "minTotObsNum": rows - n_ep, "nnLayerSizes": [128, 128]})
    L.load_replay(d)
    L.initialize_learner()
    n_tr = L.n_transitions
    rec = {"transitions": n_tr}
    for name, fn, bpt in (("fused", L.fused_sweep, 156), ("retrace", L.retrace_sweep, 24), ("moments", L.reward_state_moments, 132)):
        best = None
        for _ in range(5):
            bench.flush_l2(torch, dev)
            fn()
            ms, _ = L.last_timing()
            best = ms if best is None else min(best, ms)
        gbs = bpt * n_tr / (best * 1e-3) / 1e9
        rec[name] = {"ms": best, "algorithmic_bytes": bpt * n_tr, "achieved_gbs": gbs, "frac": gbs / peaks["hbm_gbs"]}
    rec["separate_total_ms"] = rec["retrace"]["ms"] + rec["moments"]["ms"]
    out.append(rec)
    print(json.dumps(rec), flush=True)
    L.close()
    del d
