"""Phase timing of the persistent step kernel on the cfg2 workload (SM clock cycles -> us at the
reported SM clock).  Markers: 0 step start, 1 gather done, 2 forward done, 3 loss done,
4 backward done, 5 P1 done (scratch written), 6 after barrier 1, 7 P2/P3 done."""
import os
import sys

os.environ.setdefault("SMB200_PROFILE", "1")   # the library flavour with phase timestamps

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
from smarties_b200 import Learner, synth  # noqa: E402

n_ep = int(os.environ.get("PROF_NEP", "1000"))
n = int(os.environ.get("PROF_STEPS", "200"))
mhz = float(os.environ.get("SM_MHZ", "1965"))
world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
d = synth.make_replay(123 + rank, n_ep, 1000, 32, 8)
if world > 1:      # launched by torch.distributed.run: the fused gradient exchange is part of P2
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    L = Learner(32, 8, {"maxTotObsNum": 1048576 * world, "minTotObsNum": 1000 * n_ep * world, "batchSize": 256 * world},
                device=local, seed=42 + rank, world_rank=rank, world_size=world)
    L.attach_process_group(dist)
else:
    cap = 1 << (1000 * n_ep - 1).bit_length()          # PROF_NEP=8000: the 8 M-transition buffer of the strong-scaling arm
    L = Learner(32, 8, {"maxTotObsNum": max(1048576, cap), "minTotObsNum": 1000 * n_ep})
L.load_replay(d)
L.initialize_learner()
L.seed_sampler(7 + rank)
L.train_steps(1, want_stats=False)
L.presample(n)
L.train_presampled(0, 50); L.sync()
if world > 1:
    dist.barrier()
T, ms = L.profile_phases(n)
if world > 1:
    dist.barrier()
    import time
    time.sleep(0.3 * rank)     # keep the ranks' reports apart
    print(f"--- rank {rank} of {world}")
print(f"{n} steps in {ms:.3f} ms -> {1e3 * ms / n:.2f} us/step, grid {T.shape[1]}")
T = T[20:]                       # skip the first steps
nP1 = 64
p1 = T[:, :nP1, :]
us = lambda c: c / mhz
seg = [("start->gather", 0, 1), ("forward", 1, 2), ("loss", 2, 3), ("backward", 3, 4), ("scratch write", 4, 5),
       ("barrier1 wait", 5, 6), ("P2 (tile+adam)", 6, 7)]
for name, a, b in seg:
    dlt = p1[:, :, b] - p1[:, :, a]
    print(f"  P1 CTAs {name:16s} mean {us(dlt.mean()):6.2f} us  max-over-CTAs mean {us(dlt.max(axis=1).mean()):6.2f}")
fine = [("start: fence.proxy", 0, 41), ("start: TMA issue", 41, 37), ("start: next idx", 37, 38), ("start: gather", 38, 1), ("fwd L1 wait W", 9, 27), ("fwd L1 compute", 27, 10), ("fwd L2 wait W", 10, 28), ("fwd L2+res", 28, 12), ("fwd L4 wait W", 12, 30), ("fwd L4 out", 30, 31 if os.environ.get("ICACHE_PROBE") else 13), ("fwd L4 again", 31, 32 if os.environ.get("ICACHE_PROBE") else 31), ("fwd L5 param", 13, 2),
        ("loss stage1", 2, 8), ("loss stage2", 8, 16), ("loss stage3", 16, 3),
        ("bwd L4 out", 20, 18), ("bwd L2+res", 18, 17), ("bwd L1", 17, 4),
        ("P2 pf states", 6, 39), ("P2 pf old vals", 39, 40), ("P2 pf pairs", 40, 35), ("P2 ctrl+tile", 35, 36), ("P2 param loads", 36, 24), ("P2 tile load", 24, 25), ("P2 contraction", 25, 33 if world > 1 else 26), ("P2 peer stores", 33, 34 if world > 1 else 33),
        ("P2 peer wait", 34, 26 if world > 1 else 34), ("P2 adam+store", 26, 7)]
for name, a, b in fine:
    dlt = p1[:, :, b] - p1[:, :, a]
    print(f"     {name:18s} {us(np.median(dlt)):6.2f} us")      # median: some markers are skipped on a launch's last step
oth = T[:, nP1:-1, :]
if oth.shape[1]:
    dlt = oth[:, :, 7] - oth[:, :, 6]
    print(f"  tile-only workers P2: mean {us(dlt.mean()):.2f} us")
st = T[:, -1, :]
print(f"  statistics CTA (async P3): {us((st[:, 7] - st[:, 6]).mean()):.2f} us per step")
nxt = T[1:, :-1, 0] - T[:-1, :-1, 7]
print(f"  barrier2 wait (P2 end -> next step start): mean {us(nxt.mean()):.2f} us")
per = T[1:, 0, 0] - T[:-1, 0, 0]
print(f"  step period (CTA 0): {us(per.mean()):.2f} us")
sys.stdout.flush()
if world > 1:
    dist.barrier()
L.close()
if world > 1:
    dist.destroy_process_group()
