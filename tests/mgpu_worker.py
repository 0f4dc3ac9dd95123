"""Worker of tests/test_gpu_multirank.py: launched by torch.distributed.run, one rank per GPU."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
from smarties_b200 import Learner, synth  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dS, dA, Bl = 12, 4, 32
    net = ({"learner": "RACER", "nnType": "LSTM", "nnLayerSizes": [24], "nnBPTTseq": 6} if os.environ.get("MGPU_NET") == "lstm"
           else {"nnLayerSizes": [64, 64]})       # lstm: the tensor-core weight gradient feeds the exchange
    settings = dict(net, batchSize=Bl * world, maxTotObsNum=4096 * world, minTotObsNum=1000 * world)
    d = synth.make_replay(100 + rank, 30, (40, 70), dS, dA)
    L = Learner(dS, dA, settings, device=local, seed=5 + rank, world_rank=rank, world_size=world)
    L.attach_process_group(dist)
    w0 = L.get_weights()
    L.load_replay(d)
    L.initialize_learner()
    L.seed_sampler(11 + rank)
    # a single-rank clone of this shard: same weights, same samples, same global scaling
    S = Learner(dS, dA, dict(net, batchSize=Bl, maxTotObsNum=4096, minTotObsNum=1000), device=local, seed=5)
    S.set_weights(w0)
    S.load_replay(d)
    S.initialize_learner()
    S.set_scaling(*L.get_scaling())         # the 2-rank run normalises with moments summed over ranks
    S.retrace_sweep()
    pos, t = L.sample_minibatch()
    L.seed_sampler(11 + rank)
    beta0, beta0_alone = L.get_stats()["beta"], S.get_stats()["beta"]
    st = L.train_steps(1)[0]
    S.train_step_on(pos, t)
    g_alone = torch.from_numpy(S.get_grad()).cuda()
    gs = [torch.empty_like(g_alone) for _ in range(world)]
    dist.all_gather(gs, g_alone)
    g_sum = gs[0].clone()
    for q in range(1, world):
        g_sum = g_sum + gs[q]
    g_fused = L.get_grad()
    res = dict(rank=rank, grad_equal=bool(np.array_equal(g_fused, g_sum.cpu().numpy())),
               grad_maxabs=float(np.abs(g_fused).max()), beta0=beta0, beta0_alone=beta0_alone)
    stats = L.train_steps(40)
    L.comm_check()
    w = torch.from_numpy(L.get_weights()).cuda()
    ws = [torch.empty_like(w) for _ in range(world)]
    dist.all_gather(ws, w)
    res["weights_identical"] = all(bool(torch.equal(ws[0], x)) for x in ws)
    res["weights_moved"] = float((w.cpu() - torch.from_numpy(w0)).abs().max())
    res["beta"] = stats[-1]["beta"]
    res["grad_step"] = stats[-1]["grad_step"]
    res["finite"] = bool(np.isfinite(L.get_weights()).all())
    mean, scale, std, rew = L.get_scaling()
    res["scaling_sum"] = float(mean.sum() + scale.sum() + rew.sum())
    out = [None] * world
    dist.all_gather_object(out, res)
    if rank == 0:
        print("MGPU_RESULT " + json.dumps(out))
    L.close(); S.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
