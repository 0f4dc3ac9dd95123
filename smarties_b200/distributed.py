"""Plumbing between learner ranks (one process per GPU, torch.distributed): only what the path needs —
all-gather of small opaque byte strings (CUDA IPC handles) and a broadcast of the initial weights.
Works with the gloo backend too (CPU tests)."""
from __future__ import annotations

import numpy as np


def exchange_bytes(dist, payload: bytes) -> list[bytes]:
    """All-gather one byte string per rank, returned in rank order."""
    out = [None] * dist.get_world_size()
    dist.all_gather_object(out, payload)
    return [bytes(x) for x in out]


def broadcast_array(dist, arr: np.ndarray, src: int = 0) -> np.ndarray:
    """Broadcast a host array from `src` (MPI_Bcast of the initial weights, Parameters.h:43-46)."""
    import torch
    backend = dist.get_backend()
    t = torch.from_numpy(np.ascontiguousarray(arr).copy())
    if backend == "nccl":
        t = t.cuda()
    dist.broadcast(t, src=src)
    return t.cpu().numpy()


def shard_settings(settings: dict, world: int) -> dict:
    """What every rank derives from the global settings (HyperParameters::defineDistributedLearning,
    Settings/HyperParameters.cpp:178-205): local batch and local replay capacity."""
    from .settings import HyperParameters
    hp = HyperParameters(1, 1, {k: v for k, v in settings.items()})
    hp.define_distributed_learning(world)
    return dict(batchSize=hp.batchSize, batchSize_local=hp.batchSize_local, maxTotObsNum=hp.maxTotObsNum,
                maxTotObsNum_local=hp.maxTotObsNum_local, minTotObsNum_local=hp.minTotObsNum_local)
