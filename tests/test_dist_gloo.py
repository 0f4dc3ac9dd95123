"""CPU, world_size 2, gloo: the host-side plumbing of the multi-rank learner (handle exchange,
weight broadcast, per-rank shard sizes).  The device-side exchange itself is covered by the -m gpu
test that needs two GPUs."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from smarties_b200.distributed import broadcast_array, exchange_bytes, shard_settings
    handles = exchange_bytes(dist, bytes([rank]) * 64)
    w = np.full(1000, float(rank + 1), np.float32)
    w = broadcast_array(dist, w, src=0)
    sh = shard_settings({"batchSize": 256, "maxTotObsNum": 2097152}, world)
    # every rank samples its own shard with its own generator (randSeed += world_rank, ExecutionInfo.cpp:387): the
    # library's host sampler (smb200_host_replay_trace, no device) on this rank's episodes, exchanged for the cross-check
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from test_host_replay import _trace
    from smarties_b200 import load_library
    rows = np.full(40, 26, np.int32)                           # 40 episodes x 25 transitions per rank
    ids = np.arange(40) * world + rank                         # episodes dealt round-robin (DataCoordinator.cpp:91-112)
    rc, ep, t, n_after, _ = _trace(load_library(), sh["batchSize_local"], sh["maxTotObsNum_local"], ids, rows,
                                   np.zeros(40, np.int32), 42 + rank, 3)
    assert rc == 0
    picks = exchange_bytes(dist, np.stack([ep, t]).tobytes())
    q.put((rank, [h[0] for h in handles], [len(h) for h in handles], float(w[0]), float(w[-1]), sh, picks))
    dist.destroy_process_group()


def test_two_rank_plumbing():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 1000
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, firsts, lens, w0, w1, sh, picks in res:
        assert firsts == [0, 1] and lens == [64, 64]          # handles arrive in rank order
        assert w0 == 1.0 and w1 == 1.0                          # rank 0's weights everywhere
        assert sh["batchSize_local"] == 128 and sh["maxTotObsNum_local"] == 1048576
        # the local mini-batches of the two ranks: 128 unique (episode, t) each, from disjoint episode sets, different
        # draws (own generator per rank) — together the global batch of 256
        per_rank = [np.frombuffer(b, np.int64).reshape(2, 3, 128) for b in picks]
        for r, (ep, t) in enumerate(per_rank):
            assert np.all(ep % world == r) and np.all((t >= 0) & (t < 25))
            for k in range(3):
                assert len(set(zip(ep[k].tolist(), t[k].tolist()))) == 128
        assert not np.array_equal(per_rank[0][1], per_rank[1][1])
    # both ranks saw the same exchanged picks
    assert res[0][6] == res[1][6]


def test_per_rank_shards_cover_global_batch():
    from smarties_b200.distributed import shard_settings
    for world in (1, 2, 4, 8):
        sh = shard_settings({"batchSize": 256 * world, "maxTotObsNum": 1048576 * world}, world)
        assert sh["batchSize_local"] == 256 and sh["batchSize"] == 256 * world
        assert sh["maxTotObsNum_local"] == 1048576
