"""Synthetic replay buffers shared by tests, bench.py and the oracle harness.

The same arrays are fed to the GPU learner, to the numpy oracle and (through a binary file) to
the compiled reference harness, so all three see bit-identical inputs.

Layout of one buffer: `n_ep` episodes; episode e has `N[e]` rows (N-1 data steps + the final
terminal/truncated row whose action and behaviour policy are zero, as in the reference:
ReplayMemory/MemoryBuffer.cpp:118-131).  Rows of all episodes are concatenated.
"""
from __future__ import annotations

import numpy as np

MAGIC = 0x31424D53  # "SMB1"


def make_replay(seed: int, n_ep: int, ep_len, dS: int, dA: int, term_every: int = 10,
                mu_std: float = 0.4472135955, mu_mean_scale: float = 0.3):
    """Seeded synthetic buffer.  `ep_len` = number of data steps per episode (int) or a
    (lo, hi) range for ragged episodes.  Every `term_every`-th episode ends in a terminal
    state, the others are truncated.  All arrays are rounded to f32 before anyone sees them."""
    rng = np.random.default_rng(seed)
    if isinstance(ep_len, (tuple, list)):
        nd = rng.integers(ep_len[0], ep_len[1] + 1, size=n_ep)
    else:
        nd = np.full(n_ep, int(ep_len), dtype=np.int64)
    N = (nd + 1).astype(np.int64)
    tot = int(N.sum())
    start = np.concatenate([[0], np.cumsum(N)[:-1]]).astype(np.int64)
    S = rng.standard_normal((tot, dS), dtype=np.float32)
    mu_mean = (mu_mean_scale * rng.standard_normal((tot, dA), dtype=np.float32)).astype(np.float32)
    mu_sd = np.full((tot, dA), mu_std, dtype=np.float32)
    A = (mu_mean + mu_sd * rng.standard_normal((tot, dA), dtype=np.float32)).astype(np.float32)
    R = rng.standard_normal(tot, dtype=np.float32)
    MU = np.concatenate([mu_mean, mu_sd], axis=1).astype(np.float32)
    last = start + N - 1
    A[last] = 0
    MU[last] = 0
    R[start] = 0  # reward of the initial state is 0 (Episode.cpp:239 asserts it)
    term = (np.arange(n_ep) % term_every == 0).astype(np.int64) if term_every > 0 else np.zeros(n_ep, np.int64)
    return dict(dS=dS, dA=dA, N=N, term=term, start=start, S=S, A=A, MU=MU, R=R)


def make_replay_discrete(seed: int, n_ep: int, ep_len, dS: int, n_options: int, term_every: int = 10):
    """Discrete action space with `n_options` labels (one action component): the stored action is the reference's
    action message `label + 0.1` (Core/StateAction.h:320-341), the behaviour policy the vector of option probabilities
    (Math/Discrete_policy.h:198) the label was drawn from.  Same container as make_replay (dA = 1, MU has n_options
    columns)."""
    d = make_replay(seed, n_ep, ep_len, dS, 1, term_every)
    rng = np.random.default_rng(seed + 1000003)
    tot = d["S"].shape[0]
    logits = rng.standard_normal((tot, n_options)).astype(np.float32)
    p = np.exp(logits - logits.max(axis=1, keepdims=True))
    MU = (p / p.sum(axis=1, keepdims=True)).astype(np.float32)
    u = rng.random(tot)
    label = np.minimum((np.cumsum(MU.astype(np.float64), axis=1) < u[:, None]).sum(axis=1), n_options - 1)
    A = (label.astype(np.float32) + np.float32(0.1)).reshape(tot, 1).astype(np.float32)
    last = d["start"] + d["N"] - 1
    A[last] = 0
    MU[last] = 0
    d.update(A=A, MU=MU, n_options=n_options)
    return d


def write_replay_file(path: str, d) -> None:
    """Binary file read by oracle/ref_harness.cpp (SynthData::load)."""
    with open(path, "wb") as f:
        np.array([MAGIC, d["dS"], d["dA"], len(d["N"])], dtype=np.int64).tofile(f)
        np.stack([d["N"], d["term"]], axis=1).astype(np.int64).tofile(f)
        for k in ("S", "A", "MU", "R"):
            np.ascontiguousarray(d[k], dtype=np.float32).tofile(f)


def read_dump(path: str):
    """Reader for the harness' record stream (oracle/ref_harness.cpp struct Dump)."""
    out = {}
    buf = open(path, "rb").read()
    o = 0
    dt = {0: np.float32, 1: np.float64, 2: np.int64}
    while o < len(buf):
        nl = int(np.frombuffer(buf, np.uint32, 1, o)[0]); o += 4
        name = buf[o:o + nl].decode(); o += nl
        code, nd = (int(x) for x in np.frombuffer(buf, np.uint32, 2, o)); o += 8
        dims = tuple(int(x) for x in np.frombuffer(buf, np.uint64, nd, o)); o += 8 * nd
        n = int(np.prod(dims)) if dims else 1
        arr = np.frombuffer(buf, dt[code], n, o).reshape(dims).copy(); o += arr.nbytes
        out[name] = arr
    return out
