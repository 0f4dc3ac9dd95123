"""CPU: the library's own HOST code against the reference binary's golden runs, through host-only C-ABI diagnostics
(no device work, no GPU; the same functions the learner runs):
  smb200_host_replay_trace     sampler (Sample_uniform::sample + Sampling::IDtoSeqStep, ReplayMemory/Sampling.cpp:26-47,82-93),
                               FIFO episode removal (MemoryProcessing.cpp:327-351), ring allocator, the Adam update's draw from
                               the sampler's generator (Optimizer.cpp:139) — sampled (episode, t) and episode order bit-exact
  smb200_host_init_weights     RACER::setupNet + Builder::build: parameter blob layout and initial values bit-exact
  smb200_host_strip_weights    Network::save order: byte-identical to the reference's checkpoint weight files
  smb200_host_repack_episodes  MemoryBuffer::save/restart episode format: the reference's file parsed and re-packed, byte-identical
  smb200_host_write_grad_stats StatsTracker file: header, append rule, values
  smb200_host_adam / smb200_host_value_scaling / smb200_host_return_estimator
                               host builds of device source lines (Adam epilogue, scaleNet2V, Retrace / GAE / retraceExplore
                               recursion): bit-exact with the oracle"""
import ctypes as C
import os

import numpy as np
import pytest

from parity_utils import CASES, ORACLE_ONLY_CASES, RECURRENT_CASES, Golden

UNIFORM_FIFO = [c for c in CASES + RECURRENT_CASES + ORACLE_ONLY_CASES
                if c not in ("vracer_pererr", "vracer_perseq", "vracer_perrank", "vracer_farpolfrac", "vracer_maxkldiv", "vracer_minerror")]


def _trace(lib, batch, max_tot_obs, ids, rows, term, seed, steps, capacity=0, push_before_step=None, want_starts=False):
    P = C.POINTER
    lib.smb200_host_replay_trace.restype = C.c_int
    lib.smb200_host_replay_trace.argtypes = [C.c_int32, C.c_int64, C.c_int64, C.c_int32, P(C.c_int64), P(C.c_int32), P(C.c_int32),
                                             C.c_uint64, C.c_int32, P(C.c_int64), P(C.c_int64), P(C.c_int32), P(C.c_int64),
                                             P(C.c_int32), P(C.c_int64)]
    ids, rows, term = (np.ascontiguousarray(ids, np.int64), np.ascontiguousarray(rows, np.int32), np.ascontiguousarray(term, np.int32))
    n = len(ids)
    ep = np.zeros((steps, batch), np.int64); t = np.zeros((steps, batch), np.int64)
    n_after = np.zeros(steps, np.int32); order = np.zeros((steps, n), np.int64)
    sched = None if push_before_step is None else np.ascontiguousarray(push_before_step, np.int32)
    starts = np.zeros((steps, n), np.int64) if want_starts else None
    rc = lib.smb200_host_replay_trace(batch, max_tot_obs, capacity, n, ids.ctypes.data_as(P(C.c_int64)), rows.ctypes.data_as(P(C.c_int32)),
                                      term.ctypes.data_as(P(C.c_int32)), seed, steps, ep.ctypes.data_as(P(C.c_int64)),
                                      t.ctypes.data_as(P(C.c_int64)), n_after.ctypes.data_as(P(C.c_int32)), order.ctypes.data_as(P(C.c_int64)),
                                      None if sched is None else sched.ctypes.data_as(P(C.c_int32)),
                                      None if starts is None else starts.ctypes.data_as(P(C.c_int64)))
    if want_starts:
        return rc, ep, t, n_after, order, starts
    return rc, ep, t, n_after, order


@pytest.mark.parametrize("case", UNIFORM_FIFO)
def test_host_sampler_and_fifo_match_the_reference_run(built_library, case):
    from smarties_b200 import HyperParameters, load_library
    g = Golden(case)
    s = g.settings
    assert s.get("dataSamplingAlgo", "uniform") == "uniform" and s.get("ERoldSeqFilter", "oldest") in ("oldest", "default")
    max_obs = s.get("maxTotObsNum", HyperParameters(g.dS, g.dA, {}).maxTotObsNum)
    rows = np.asarray(g.replay["N"], np.int32)
    rc, ep, t, n_after, order = _trace(load_library(), g.B, max_obs, np.arange(len(rows)), rows, g.replay["term"], g.sample_seed, g.steps)
    assert rc == 0
    for k in range(g.steps):
        assert np.array_equal(ep[k], g.ref[f"s{k}/sampledEpID"]), f"step {k}: sampled episodes"
        assert np.array_equal(t[k], g.ref[f"s{k}/sampledT"]), f"step {k}: sampled time steps"
        if f"s{k}/post/epID" in g.ref:       # full dumps only at the steps make_golden.py lists
            kept = g.ref[f"s{k}/post/epID"]
            assert n_after[k] == len(kept) and np.array_equal(order[k, :len(kept)], kept), f"step {k}: episode vector"
            assert np.all(order[k, len(kept):] == -1)


def test_host_sampler_is_unique_ascending_and_full_range(built_library):
    """Sample_uniform::sample (Sampling.cpp:82-93): B unique ascending ids; here at the extremes — a batch as large as the
    buffer must return every transition exactly once, ragged episodes map to (episode, t) with t < ndata."""
    from smarties_b200 import load_library
    rows = np.array([2, 5, 3, 2, 9, 4], np.int32)                   # ndata = rows - 1 -> 19 transitions
    nd = int((rows - 1).sum())
    rc, ep, t, n_after, _ = _trace(load_library(), nd, 1 << 20, np.arange(6) + 10, rows, np.zeros(6, np.int32), 11, 4)
    assert rc == 0 and np.all(n_after == 6)
    want = sorted((10 + e, k) for e in range(6) for k in range(rows[e] - 1))
    for k in range(4):
        # step 0 samples in push order (ascending ids); from step 1 on the vector is sorted by id descending
        got = sorted(zip(ep[k].tolist(), t[k].tolist()))
        assert got == want
    assert ep[0].tolist() == sorted(ep[0].tolist()) and ep[1].tolist() == sorted(ep[1].tolist(), reverse=True)


def test_host_trace_rejects_bad_input(built_library):
    from smarties_b200 import load_library
    lib = load_library()
    rc, *_ = _trace(lib, 64, 1 << 20, [0, 1], [5, 5], [0, 0], 1, 1)         # 8 transitions < batch 64
    assert rc != 0 and b"not enough transitions" in lib.smb200_last_error()
    rc, *_ = _trace(lib, 2, 1 << 20, [0, 1], [5, 1], [0, 0], 1, 1)          # an episode needs s0 and sT
    assert rc != 0
    rc, *_ = _trace(lib, 2, 1 << 20, [0, 1, 2], [40, 40, 40], [0, 0, 0], 1, 1, capacity=64)   # ring of 64 rows
    assert rc != 0 and b"ring full" in lib.smb200_last_error()


def test_fifo_pruning_frees_ring_rows_for_reuse(built_library):
    """maxTotObsNum below the stored transitions: the oldest episodes leave from the back of the id-descending vector
    while `nTransitions - back.nsteps > maxTotObsNum` (MemoryBuffer.cpp:469-477 / MemoryProcessing.cpp:340-349)."""
    from smarties_b200 import load_library
    rows = np.full(10, 11, np.int32)                                  # 10 episodes x 10 transitions
    rc, ep, t, n_after, order = _trace(load_library(), 4, 55, np.arange(10), rows, np.ones(10, np.int32), 5, 3)
    assert rc == 0
    # pops while nTransitions - 11 > 55: 100 -> 90 -> 80 -> 70 -> 60 (60 - 11 = 49: stop), 6 episodes stay
    assert n_after.tolist() == [6, 6, 6] and order[0, :6].tolist() == [9, 8, 7, 6, 5, 4]
    assert set(ep[1].tolist()) <= {4, 5, 6, 7, 8, 9} and set(ep[2].tolist()) <= {4, 5, 6, 7, 8, 9}


DEVICE_NET_CASES = CASES + RECURRENT_CASES


@pytest.mark.parametrize("case", DEVICE_NET_CASES)
def test_network_construction_is_bit_exact_with_the_reference_builder(built_library, monkeypatch, case):
    """RACER::setupNet + Builder::build (Learners/RACER_common.cpp:70-115, Network/Builder.cpp:48-99,133-137): the padded
    parameter blob of Parameters.h:159-176 (Appendix B of SURVEY.md) with every layer initialised from generators[0] =
    mt19937(randSeed) — dense Xavier draws in (i, o) order, LSTM gate biases, residual ones, the ParamLayer's inverse SoftPlus
    of explNoise, RACER's advantage biases.  The library's own build_net + init_weights (what smb200_create uploads), run on
    the host, against the weights the reference binary started from in every golden run (randSeed 42): identical bits."""
    from smarties_b200 import load_library
    from smarties_b200.learner import make_config
    g = Golden(case)
    lib = load_library()
    lib.smb200_host_init_weights.restype = C.c_int64
    cfg, _ = make_config(g.dS, g.dA, dict(g.settings), bounded=g.bounded, seed=42, discrete_options=g.spec["replay"].get("n_options", 0))
    n = lib.smb200_host_init_weights(C.byref(cfg), None, 0)
    ref = g.ref["init/weights"]
    assert n == ref.size
    w = np.zeros(n, np.float32)
    assert lib.smb200_host_init_weights(C.byref(cfg), w.ctypes.data_as(C.POINTER(C.c_float)), n) == n
    assert np.array_equal(w.view(np.uint32), ref.view(np.uint32))


def test_network_construction_rejects_what_the_device_path_does_not_build(built_library):
    from smarties_b200 import load_library
    from smarties_b200.learner import make_config
    lib = load_library()
    lib.smb200_host_init_weights.restype = C.c_int64
    cfg, _ = make_config(6, 3, {"nnLayerSizes": [32, 32]})
    assert lib.smb200_host_init_weights(C.byref(cfg), None, 0) == 1616
    w = np.zeros(10, np.float32)
    assert lib.smb200_host_init_weights(C.byref(cfg), w.ctypes.data_as(C.POINTER(C.c_float)), 10) < 0   # wrong blob size
    cfg.n_hidden = 0
    assert lib.smb200_host_init_weights(C.byref(cfg), None, 0) < 0
    assert lib.smb200_host_init_weights(None, None, 0) < 0


@pytest.mark.parametrize("case", ["vracer_ckpt", "racer_lstm_ckpt"])
def test_checkpoint_weight_order_matches_the_reference_files(built_library, case):
    """Network::save (Network.cpp:22-67) strips the padding of the parameter blob layer by layer.  The library's strip_copy
    (used by smb200_save / smb200_restart), run on the host: the reference's final weights in blob form -> exactly the bytes of
    the <name>_net_weights.raw the reference wrote; the initial weights -> its _tgt_weights.raw (targetDelay 0: the target
    network keeps the weights of construction); and back into a blob."""
    from smarties_b200 import load_library
    from smarties_b200.learner import make_config
    g = Golden(case)
    lib = load_library()
    lib.smb200_host_strip_weights.restype = C.c_int64
    fp = C.POINTER(C.c_float)
    cfg, _ = make_config(g.dS, g.dA, dict(g.settings), bounded=g.bounded)
    ns = lib.smb200_host_strip_weights(C.byref(cfg), None, 0, None, 0, 1)
    for blob_key, fname in (("final/weights", "agent_00_net_weights.raw"), ("init/weights", "agent_00_net_tgt_weights.raw")):
        file_bytes = bytes(g.ckpt[fname])
        assert 4 * ns == len(file_bytes)
        blob = np.ascontiguousarray(g.ref[blob_key], np.float32).copy()
        flat = np.zeros(ns, np.float32)
        assert lib.smb200_host_strip_weights(C.byref(cfg), blob.ctypes.data_as(fp), blob.size, flat.ctypes.data_as(fp), ns, 1) == ns
        assert flat.tobytes() == file_bytes, fname
        back = np.zeros_like(blob)
        assert lib.smb200_host_strip_weights(C.byref(cfg), back.ctypes.data_as(fp), back.size, flat.ctypes.data_as(fp), ns, -1) == ns
        assert np.array_equal(back.view(np.uint32), blob.view(np.uint32))       # the reference keeps its padding at zero
    bad = np.zeros(3, np.float32)
    assert lib.smb200_host_strip_weights(C.byref(cfg), bad.ctypes.data_as(fp), 3, bad.ctypes.data_as(fp), 3, 1) < 0


@pytest.mark.parametrize("case", CASES + RECURRENT_CASES)
def test_grad_stats_writer_matches_the_reference_file(built_library, tmp_path, case):
    """StatsTracker (Utils/StatsTracker.cpp:28-107): the library's own file writer (row a26), run on the host with the
    reference's per-sample output gradients of the tracker steps, against the `<learner>_<net>_outGrad_stats.raw` the
    reference wrote in the same golden run (tests/golden/outgrad_stats.npz): same size, same header word, same
    overwrite-then-append rule, values to f32 round-off of the dumped gradients."""
    from smarties_b200 import load_library
    g = Golden(case)
    lib = load_library()
    want = np.load(g.path("outgrad_stats.npz"))[case]
    n_out = g.ref["s0/g"].shape[1]
    base = str(tmp_path / "agent_00_net")
    printed = [s for s in range(g.steps) if (g.start_step + s) % 1000 == 0]
    for s in printed:
        gs = np.ascontiguousarray(g.ref[f"s{s}/g"], np.float32)
        assert lib.smb200_host_write_grad_stats(base.encode(), gs.shape[0], n_out, gs.ctypes.data_as(C.POINTER(C.c_float)), int(s == 0)) == 0
    if not printed:
        assert want.size == 0
        return
    got = np.fromfile(base + "_outGrad_stats.raw", np.float32)
    header = 1 if printed[0] == 0 else 0
    assert got.size == want.size == header + 2 * n_out * len(printed)
    if header:
        assert got[0] == want[0] == np.float32(n_out + .1)
    assert np.allclose(got, want, rtol=2e-6, atol=1e-6 * np.abs(want[header:]).max())


@pytest.mark.parametrize("case", ["vracer_ckpt", "racer_lstm_ckpt"])
def test_episode_file_parser_and_packer_round_trip_the_reference_file(built_library, case):
    """MemoryBuffer::save / restart (MemoryBuffer.cpp:172-324), Episode::packEpisode / unpackEpisode (Episode.cpp:24-130):
    the reference's `agent_00_rank_000_learner_data.raw` read by the parser of smb200_restart and written again by the packer
    of smb200_save, both on the host: byte-identical file, episodes and lengths as the reference held them in memory."""
    from smarties_b200 import load_library
    g = Golden(case)
    lib = load_library()
    lib.smb200_host_repack_episodes.restype = C.c_int64
    raw = bytes(g.ckpt["agent_00_rank_000_learner_data.raw"])
    src = np.frombuffer(raw, np.uint8).copy()
    out = np.zeros(src.size + 64, np.uint8)
    n_ep = C.c_int64(0)
    cap = 1024
    ids = np.zeros(cap, np.int64); rows = np.zeros(cap, np.int32); term = np.zeros(cap, np.int32)
    u8, i64, i32 = C.POINTER(C.c_uint8), C.POINTER(C.c_int64), C.POINTER(C.c_int32)
    n = lib.smb200_host_repack_episodes(C.c_int32(g.dS), C.c_int32(g.dA), src.ctypes.data_as(u8), C.c_int64(src.size), out.ctypes.data_as(u8),
                                        C.c_int64(out.size), C.c_int64(cap), C.byref(n_ep), ids.ctypes.data_as(i64), rows.ctypes.data_as(i32),
                                        term.ctypes.data_as(i32))
    assert n == src.size and out[:n].tobytes() == raw
    k = n_ep.value
    assert ids[:k].tolist() == list(g.ref["final/epID"]) and rows[:k].tolist() == list(g.ref["final/epLen"])
    want_term = {int(i): int(t) for i, t in enumerate(g.replay["term"])}
    assert [want_term[int(i)] for i in ids[:k]] == term[:k].tolist()
    # a truncated image is an error, not a partial read
    bad = lib.smb200_host_repack_episodes(C.c_int32(g.dS), C.c_int32(g.dA), src.ctypes.data_as(u8), C.c_int64(src.size - 12), out.ctypes.data_as(u8),
                                          C.c_int64(out.size), C.c_int64(cap), None, None, None, None)
    assert bad < 0 and b"learner_data.raw" in lib.smb200_last_error()
    # wrong dimensions do not parse to the end of the file either
    bad = lib.smb200_host_repack_episodes(C.c_int32(g.dS + 1), C.c_int32(g.dA), src.ctypes.data_as(u8), C.c_int64(src.size), out.ctypes.data_as(u8),
                                          C.c_int64(out.size), C.c_int64(cap), None, None, None, None)
    assert bad < 0


def test_adam_epilogue_source_is_bit_exact_with_the_reference_optimizer(built_library):
    """struct Adam + AdamOptimizer::apply_update (Network/Optimizer.cpp:61-108,122-161: Nesterov, "safe", AdamW, annealed and
    bias-corrected learning rate, running beta powers): the source lines of the weight-gradient epilogue (`adam_step`,
    `adam_eta_for`), built for the host, against the oracle's f32 restatement (pinned to the reference's weights after every
    golden step) over 40 consecutive updates incl. late steps where beta_1^t underflows to 0: identical bits in W, M1, M2."""
    import vracer_oracle as vo
    from smarties_b200 import load_library
    lib = load_library()
    fp = C.POINTER(C.c_float)
    rng = np.random.default_rng(3)
    for start, B in ((0, 256), (130, 16), (20000, 256)):
        o = vo.VracerOracle(6, 3, hidden=[32, 32], batch=B)
        n = o.W.size
        o.W[:] = (rng.standard_normal(n) * 0.1).astype(np.float32)
        o.adam_step = start
        o.beta_t_1, o.beta_t_2 = 0.9, 0.999
        for _ in range(start):                                   # the running powers as Optimizer.cpp:155-158 leaves them
            o.beta_t_1 *= 0.9
            if o.beta_t_1 < vo.FLT_EPS: o.beta_t_1 = 0
            o.beta_t_2 *= 0.999
            if o.beta_t_2 < vo.FLT_EPS: o.beta_t_2 = 0
        W, M1, M2 = o.W.copy(), o.M1.copy(), o.M2.copy()
        for k in range(40):
            G = (rng.standard_normal(n) * (10.0 if k % 7 == 0 else 0.01)).astype(np.float32)
            done, bt1, bt2 = o.adam_step, o.beta_t_1, o.beta_t_2
            assert lib.smb200_host_adam(C.c_int64(n), G.ctypes.data_as(fp), W.ctypes.data_as(fp), M1.ctypes.data_as(fp), M2.ctypes.data_as(fp),
                                        C.c_double(o.eta), C.c_double(o.eps_anneal), C.c_int64(done), C.c_double(bt1), C.c_double(bt2),
                                        C.c_double(o.nn_lambda), C.c_int32(B)) == 0
            o.adam_step += 1                                     # prepare_update: nStep++ (Optimizer.cpp:119)
            o.apply_adam(G)
            for mine, ref, name in ((W, o.W, "W"), (M1, o.M1, "M1"), (M2, o.M2, "M2")):
                assert np.array_equal(mine.view(np.uint32), ref.view(np.uint32)), (start, k, name)


def test_value_scaling_source_matches_the_oracle(built_library):
    """scaleNet2V / scaleVdiff (Learners/RACER_common.cpp:23-32), the device source built for the host, f64."""
    import vracer_oracle as vo
    from smarties_b200 import load_library
    lib = load_library()
    dp = C.POINTER(C.c_double)
    x = np.concatenate([np.linspace(-30, 30, 2001), [0.0, -0.0, 1e-300, -1e-300, 1e6, -1e6]])
    v, d = np.zeros_like(x), np.zeros_like(x)
    assert lib.smb200_host_value_scaling(C.c_int64(x.size), x.ctypes.data_as(dp), v.ctypes.data_as(dp), d.ctypes.data_as(dp)) == 0
    assert np.array_equal(v, vo.scale_net2v(x)) and np.array_equal(d, vo.scale_vdiff(x))


@pytest.mark.parametrize("case,estimator", [("vracer_small", 0), ("vracer_gae", 1), ("vracer_explore", 2)])
def test_return_estimator_source_is_bit_exact_with_the_oracle(built_library, case, estimator):
    """computeRetrace / computeGAE / computeRetraceExplBonus (MemoryProcessing.cpp:391-417) and updateReturnEstimator (:23-44):
    the scalar functions the sweep kernels call (reward scaling, clipped importance weight, expression order of the recursion),
    built for the host and run sequentially over every episode of a golden buffer after some learner steps (non-trivial V, A,
    rho, non-zero maxAbsError), against the oracle's restatement — itself pinned to the reference's Q_ret: identical bits."""
    from parity_utils import make_oracle
    from smarties_b200 import load_library
    lib = load_library()
    lib.smb200_host_return_estimator.restype = C.c_double
    fp = C.POINTER(C.c_float)
    g = Golden(case)
    o = make_oracle(g)
    for _ in range(3):
        o.train_step()
    assert o.stats["maxAbsErr"] > 0
    for ep in o.episodes:
        N = ep.nsteps
        R, V, A, W = (np.ascontiguousarray(x, np.float32) for x in (ep.R, ep.V, ep.ADV, ep.rho))
        Q = np.ascontiguousarray(ep.Q, np.float32).copy()
        err = lib.smb200_host_return_estimator(C.c_int32(N), C.c_int32(int(ep.terminated)), C.c_int32(estimator), R.ctypes.data_as(fp),
                                               V.ctypes.data_as(fp), A.ctypes.data_as(fp), W.ctypes.data_as(fp), Q.ctypes.data_as(fp),
                                               C.c_double(o.gamma), C.c_double(o.lam), C.c_float(o.rew_mean), C.c_float(o.rew_scale),
                                               C.c_double(o.stats["maxAbsErr"]))
        want_err = o.retrace_episode(ep)
        assert np.array_equal(Q.view(np.uint32), np.asarray(ep.Q, np.float32).view(np.uint32)), ep.ID
        assert err == pytest.approx(want_err, rel=1e-5)        # the oracle squares with numpy's scalar power (powf), not d * d


def test_discrete_action_network_construction_matches_the_reference(built_library):
    """RACER<Discrete_advantage, Discrete_policy, Uint>::setupNet (Learners/RACER_common.cpp:70-135): outputs
    [V | advantages(K) | policy(K)] from one linear layer, no ParamLayer, no initial biases.  Only the construction of this
    network exists on the device side so far (smb200_create rejects discrete_options != 0): layout and
    initial weights identical to the reference's in the golden run with 5 options."""
    from smarties_b200 import load_library
    from smarties_b200.learner import make_config
    g = Golden("racer_discrete")
    K = g.spec["replay"]["n_options"]
    lib = load_library()
    lib.smb200_host_init_weights.restype = C.c_int64
    cfg, _ = make_config(g.dS, g.dA, dict(g.settings), seed=42, discrete_options=K)
    ref = g.ref["init/weights"]
    n = lib.smb200_host_init_weights(C.byref(cfg), None, 0)
    assert n == ref.size
    w = np.zeros(n, np.float32)
    assert lib.smb200_host_init_weights(C.byref(cfg), w.ctypes.data_as(C.POINTER(C.c_float)), n) == n
    assert np.array_equal(w.view(np.uint32), ref.view(np.uint32))
    # one action component, at least two options, learner RACER
    for bad in (dict(discrete_options=1), dict(discrete_options=65)):
        cfg, _ = make_config(g.dS, g.dA, dict(g.settings), **bad)
        assert lib.smb200_host_init_weights(C.byref(cfg), None, 0) < 0
    cfg, _ = make_config(g.dS, 2, dict(g.settings), discrete_options=K)
    assert lib.smb200_host_init_weights(C.byref(cfg), None, 0) < 0
    cfg, _ = make_config(g.dS, 1, {"learner": "VRACER"}, discrete_options=K)
    assert lib.smb200_host_init_weights(C.byref(cfg), None, 0) < 0


def _discrete_loss(lib, O, act, mu, qret, beta, cmax, cinv):
    fp, dp = C.POINTER(C.c_float), C.POINTER(C.c_double)
    O, act, mu, qret = (np.ascontiguousarray(x, np.float32) for x in (O, act, mu, qret))
    B, K = mu.shape
    g = np.zeros((B, 1 + 2 * K)); out = np.zeros((B, 6))
    rc = lib.smb200_host_discrete_loss(C.c_int32(B), C.c_int32(K), O.ctypes.data_as(fp), act.ctypes.data_as(fp), mu.ctypes.data_as(fp),
                                       qret.ctypes.data_as(fp), C.c_double(beta), C.c_double(cmax), C.c_double(cinv),
                                       g.ctypes.data_as(dp), out.ctypes.data_as(dp))
    assert rc == 0
    return g, out


@pytest.mark.parametrize("K", [2, 5, 17])
def test_discrete_loss_source_matches_the_oracle(built_library, K):
    """Groundwork for row f4: the per-sample loss of discrete-action RACER as the __host__ __device__ function the device loss
    stage will call, built for the host, against the oracle's Discrete_policy / Discrete_advantage restatement (pinned to the
    reference golden racer_discrete) on random samples, near- and far-policy, both signs of the value head."""
    import vracer_oracle as vo
    from smarties_b200 import load_library
    rng = np.random.default_rng(K)
    B = 96
    O = (rng.standard_normal((B, 1 + 2 * K)) * np.r_[3.0, np.ones(2 * K)]).astype(np.float32)
    p = np.exp(rng.standard_normal((B, K))); mu = (p / p.sum(1, keepdims=True)).astype(np.float32)
    label = rng.integers(0, K, B)
    act = (label + 0.1).astype(np.float32)
    qret = rng.standard_normal(B).astype(np.float32)
    for beta, cmax in ((0.3, 2.5), (1e-4, 1.2), (0.9, 1.0)):
        cinv = 1.0 / cmax
        g, out = _discrete_loss(load_library(), O, act, mu, qret, beta, cmax, cinv)
        r = vo.discrete_sample_math(O, act.reshape(B, 1), mu, qret, beta, cmax, cinv)
        assert np.array_equal(out[:, 2] != 0, r["is_far"])
        if cmax > 1.1:
            assert r["is_far"].any() and not r["is_far"].all()
        for k, name in ((0, "rho"), (1, "dkl"), (3, "V"), (4, "A"), (5, "dq")):
            assert np.allclose(out[:, k], r[name], rtol=1e-13, atol=1e-15), name
        assert np.allclose(g, r["g"], rtol=1e-12, atol=1e-15)


def test_discrete_loss_source_matches_the_reference_golden(built_library):
    """The same function on the reference's own step: network outputs, sampled actions / behaviour policies and ReF-ER scalars
    of step 0 of the golden racer_discrete -> the output gradient the reference back-propagated (f32 dump)."""
    from parity_utils import make_oracle, relerr
    from smarties_b200 import load_library
    g = Golden("racer_discrete")
    o = make_oracle(g)
    seq, obs = o.sample()
    assert np.array_equal([o.episodes[int(k)].ID for k in seq], g.ref["s0/sampledEpID"]) and np.array_equal(obs, g.ref["s0/sampledT"])
    eps = [o.episodes[int(k)] for k in seq]
    act = np.array([e.A[int(t)][0] for e, t in zip(eps, obs)], np.float32)
    mu = np.stack([e.MU[int(t)] for e, t in zip(eps, obs)])
    qret = np.array([e.Q[int(t)] for e, t in zip(eps, obs)], np.float32)
    O = np.asarray(g.ref["s0/O"], np.float32)
    grad, out = _discrete_loss(load_library(), O, act, mu, qret, o.beta, o.cmax, o.cinv)
    assert relerr(grad, g.ref["s0/g"]) < 1e-6


def test_host_sampler_at_full_buffer_size_matches_the_oracle(built_library):
    """BASELINE.json configs[1] size: 1000 ragged episodes, ~1 M transitions, batch 256 — the library's host sampler (64 Ki-bucket
    id -> (episode, t) lookup with a non-zero bucket shift, radix sort of the ids) against the oracle's restatement of
    libstdc++'s uniform_int_distribution + Sampling::IDtoSeqStep (pinned to the reference on the small goldens), 30 steps incl.
    the change of the episode order after the first step's FIFO sort.  Integer results: identical."""
    import vracer_oracle as vo
    from smarties_b200 import load_library
    rng = np.random.default_rng(11)
    n_ep, B, steps, seed = 1000, 256, 30, 42
    rows = rng.integers(900, 1101, n_ep).astype(np.int32)
    rc, ep, t, n_after, order = _trace(load_library(), B, 1 << 21, np.arange(n_ep), rows, np.zeros(n_ep, np.int32), seed, steps)
    assert rc == 0 and np.all(n_after == n_ep)
    gen = vo.Mt19937(seed)
    ids_order = np.arange(n_ep)                                   # push order at step 0
    n_tr = int((rows - 1).sum())
    assert n_tr > 900_000
    for k in range(steps):
        nd = (rows[ids_order] - 1).astype(np.int64)
        seq, obs = vo.id_to_seq_step(vo.sample_uniform(gen, n_tr, B), nd)
        assert np.array_equal(ep[k], ids_order[np.asarray(seq, np.int64)]), k
        assert np.array_equal(t[k], obs), k
        ids_order = np.arange(n_ep)[::-1]                         # applyEpisodesRemovalAlgo: ID descending from the first step on
        gen()                                                     # the Adam update's draw (Optimizer.cpp:139)
    assert order[0].tolist() == list(range(n_ep - 1, -1, -1))


@pytest.mark.parametrize("seed0,max_obs,cap", [(23, 600, 1024), (5, 300, 512), (77, 900, 1088), (101, 150, 512)])
def test_replay_under_churn_matches_the_reference_sequence_and_keeps_the_ring_consistent(built_library, seed0, max_obs, cap):
    """configs[3]'s situation on the host: actors keep pushing episodes while the learner steps, maxTotObsNum forces FIFO
    removal, the HBM ring (capacity barely above the live rows) wraps and reuses freed ranges.  Against a direct model of the
    reference's sequence — pushBackEpisode appends (MemoryBuffer.cpp:479-520); every step: sample on the current vector
    (Sampling.cpp:26-47,82-93), sort by ID descending, removeBackEpisode while nStoredSteps - back.nsteps > maxTotObsNum
    (MemoryProcessing.cpp:327-351), one generator draw for Adam (Optimizer.cpp:139) — with the oracle's sampler:
    sampled (episode, t) and the episode vector identical at every step; live ring ranges never overlap nor leave the ring."""
    import vracer_oracle as vo
    from smarties_b200 import load_library
    rng = np.random.default_rng(seed0)
    n_ep, B, steps, seed = 120, 16, 200, 9 + seed0
    rows = rng.integers(8, 41, n_ep).astype(np.int32)
    sched = np.sort(np.r_[np.zeros(20, np.int64), rng.integers(1, steps, n_ep - 20)]).astype(np.int32)   # 20 up front, the rest while training
    rc, ep, t, n_after, order, starts = _trace(load_library(), B, max_obs, np.arange(n_ep), rows, rng.integers(0, 2, n_ep), seed, steps,
                                               capacity=cap, push_before_step=sched, want_starts=True)
    assert rc == 0, load_library().smb200_last_error()
    gen = vo.Mt19937(seed)
    vec, nxt, pruned = [], 0, 0
    for k in range(steps):
        while nxt < n_ep and sched[nxt] <= k:
            vec.append(nxt); nxt += 1
        nd = (rows[vec] - 1).astype(np.int64)
        seq, obs = vo.id_to_seq_step(vo.sample_uniform(gen, int(nd.sum()), B), nd)
        assert np.array_equal(ep[k], np.asarray(vec)[np.asarray(seq, np.int64)]), k
        assert np.array_equal(t[k], obs), k
        vec.sort(reverse=True)
        while int((rows[vec] - 1).sum()) - int(rows[vec[-1]]) > max_obs:
            vec.pop(); pruned += 1
        gen()
        assert n_after[k] == len(vec) and order[k, :len(vec)].tolist() == vec, k
        # ring consistency: every live episode inside the ring, no two live ranges overlap
        s0 = starts[k, :len(vec)]; s1 = s0 + rows[vec]
        assert s0.min() >= 0 and s1.max() <= cap
        o = np.argsort(s0)
        assert np.all(s1[o][:-1] <= s0[o][1:]), k
    assert pruned > 40 and nxt == n_ep                     # the ring held several times its capacity over the run: ranges were reused
    assert int(rows.sum()) > 2 * cap


@pytest.mark.parametrize("fixture,batch", [("cfg2_full_props.npz", 256), ("cfg3_full_props.npz", 128)])
def test_host_sampler_at_full_buffer_size_matches_the_reference_binary(built_library, fixture, batch):
    """The workload of bench.py (BASELINE.json configs[1]: 1000 x 1000-step episodes = 1 M transitions, batch 256) and configs[2]
    on the same buffer (RACER + LSTM(64), batch 128): the sampled (episode, t) of the reference binary's first three learner
    steps (tests/golden/cfg{2,3}_full_props.npz, generator make_full_size_props.py) against the library's host sampler: identical."""
    import json
    from smarties_b200 import load_library
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", fixture))
    spec = json.loads(bytes(z["spec"]).decode())
    rows = np.asarray(z["init/epLen"], np.int32)
    assert rows.size == 1000 and int((rows - 1).sum()) == 1_000_000
    assert spec["settings"].get("batchSize", 256) == batch
    rc, ep, t, n_after, _ = _trace(load_library(), batch, spec["settings"]["maxTotObsNum"], np.arange(rows.size), rows,
                                   np.zeros(rows.size, np.int32), spec["sample_seed"], spec["steps"])
    assert rc == 0 and np.all(n_after == rows.size)
    for k in range(spec["steps"]):
        assert np.array_equal(ep[k], z[f"s{k}/sampledEpID"]) and np.array_equal(t[k], z[f"s{k}/sampledT"]), k


def test_retrace_at_full_buffer_size_matches_the_reference_binary(built_library):
    """rescaleAllReturnEstimator after initializeLearner on the 1 M-transition buffer of bench.py: the scalar functions of the
    sweep kernels (reward scaling with the reference's own normalisers, clipped importance weight, recursion order), run on the
    host over all 1000 episodes, against the reference binary's Q_ret at that size (subsample and checksums of
    tests/golden/cfg2_full_props.npz)."""
    import json
    import sys
    sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")))
    import bench
    from smarties_b200 import load_library
    lib = load_library()
    lib.smb200_host_return_estimator.restype = C.c_double
    fp = C.POINTER(C.c_float)
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "cfg2_full_props.npz"))
    spec = json.loads(bytes(z["spec"]).decode())
    d = bench.make_workload()
    rmean, rscale = float(z["init/rewards"][0]), float(z["init/rewards"][1])
    out = []
    for e in range(len(d["N"])):
        o, N = int(d["start"][e]), int(d["N"][e])
        R = np.ascontiguousarray(d["R"][o:o + N], np.float32).copy(); R[0] = 0.0
        V = np.zeros(N, np.float32); A = np.zeros(N, np.float32)
        W = np.ones(N, np.float32); W[-1] = 0.0                      # Episode::finalize: impW 1 except the last row
        Q = np.zeros(N, np.float32)
        err = lib.smb200_host_return_estimator(C.c_int32(N), C.c_int32(int(d["term"][e])), C.c_int32(0), R.ctypes.data_as(fp),
                                               V.ctypes.data_as(fp), A.ctypes.data_as(fp), W.ctypes.data_as(fp), Q.ctypes.data_as(fp),
                                               C.c_double(0.995), C.c_double(1.0), C.c_float(rmean), C.c_float(rscale), C.c_double(0.0))
        assert err >= 0
        out.append(Q)
    q = np.concatenate(out)
    assert q.size == 1001000
    sub = q[::spec["stride"]]
    assert np.array_equal(sub, z["init/Qret_sub"])                 # bit-identical to the reference binary (1005 of 1005 values)
    q64 = q.astype(np.float64)
    assert abs(q64.sum() - z["init/Qret_sum"][0]) < 1e-6 * np.abs(q64).sum()
    assert abs((q64 * q64).sum() - z["init/Qret_sum"][1]) < 1e-6 * z["init/Qret_sum"][1]
    assert abs(np.abs(q64).max() - z["init/Qret_sum"][2]) < 1e-5


# ------------------------------------------------------------------------------------------
# wide step (csrc/wide_step.cuh): plan, index maps and pre-split operand images, built on the host exactly as smb200_create does
# ------------------------------------------------------------------------------------------
def _wide_plan(cfg, blob=None):
    from smarties_b200 import load_library
    lib = load_library()
    P = C.POINTER
    lib.smb200_host_wide_plan.restype = C.c_int
    lib.smb200_host_wide_plan.argtypes = [C.c_void_p, P(C.c_int32), P(C.c_int32), P(C.c_float), C.c_int64, P(C.c_int32), P(C.c_float),
                                          P(C.c_float), P(C.c_float)]
    info = np.zeros(16, np.int32); dense = np.zeros(8 * 4, np.int32)
    ip = lambda a: a.ctypes.data_as(P(C.c_int32))
    fp = lambda a: a.ctypes.data_as(P(C.c_float))
    rc = lib.smb200_host_wide_plan(C.byref(cfg), ip(info), ip(dense), None, 0, None, None, None, None)
    out = dict(rc=rc, info=info, dense=dense.reshape(4, 8))
    if rc == 1 and blob is not None:
        n = blob.size
        idx = np.zeros(5 * n, np.int32)
        f, b, v = np.zeros(info[1], np.float32), np.zeros(info[2], np.float32), np.zeros(info[3], np.float32)
        assert lib.smb200_host_wide_plan(C.byref(cfg), ip(info), ip(dense), fp(blob), n, ip(idx), fp(f), fp(b), fp(v)) == 1
        out.update(idx=idx.reshape(5, n), f=f, b=b, v=v)
    return out


def _tf32_hi(w):
    """cvt.rna.tf32.f32: round to nearest, ties away from zero, on the 13 dropped mantissa bits."""
    u = w.view(np.uint32).astype(np.uint64)
    return ((u + 0x1000) & 0xFFFFE000).astype(np.uint32).view(np.float32)


@pytest.mark.parametrize("case", ["vracer_cfg2mini", "vracer_b4096", "vracer_hardsign", "vracer_small"])
def test_wide_step_plan_and_operand_images(built_library, case):
    """Every parameter of the network has exactly one position in a weight-gradient record; the forward image holds W[k][n] at
    float4 [k / 4][n] component k % 4 (K-major UMMA operand, no swizzle) as hi = TF32(w) and lo = w - hi with hi + lo == w
    exactly; the transposed image holds the same values at float4 [n / 4][k] component n % 4; biases, residual vectors and the
    ParamLayer sit in the vector block; the three tensor-core kernels fit the 227 KB of shared memory."""
    from smarties_b200.learner import make_config
    g = Golden(case)
    cfg, _ = make_config(g.dS, g.dA, dict(g.settings), bounded=g.bounded, seed=42)
    blob = g.ref["init/weights"].astype(np.float32).copy()
    p = _wide_plan(cfg, blob)
    assert p["rc"] == 1
    nD, fF, bF, vF, rec, cols, sf, sb, sg, stages, NpG = [int(x) for x in p["info"][:11]]
    hidden = [h for h in g.settings.get("nnLayerSizes", [128, 128])]
    assert nD == len(hidden) + 1 and max(sf, sb, sg) <= 227 * 1024 and cols <= 512 and stages in (1, 2)
    idx, f, b, v = p["idx"], p["f"], p["b"], p["v"]
    real = idx[0] >= 0
    # the record positions of the real parameters are distinct, inside the record, and every non-zero weight is a real parameter
    assert len(np.unique(idx[0][real])) == real.sum() and idx[0][real].max() < rec
    assert np.all(blob[~real] == 0)
    hi, lo = f[:fF // 2], f[fF // 2:]
    assert np.array_equal((hi.astype(np.float64) + lo).astype(np.float32)[idx[2][idx[2] >= 0]], blob[idx[2] >= 0])
    assert np.array_equal(hi[idx[2][idx[2] >= 0]], _tf32_hi(blob[idx[2] >= 0]))
    bh, bl = b[:bF // 2], b[bF // 2:]
    sel = idx[3] >= 0
    assert np.array_equal(bh[idx[3][sel]], _tf32_hi(blob[sel])) and np.array_equal(bl[idx[3][sel]], blob[sel] - _tf32_hi(blob[sel]))
    assert np.array_equal(v[idx[4][idx[4] >= 0]], blob[idx[4] >= 0])
    # layout formulas of the operand images, layer by layer (parameter blob: W[k][roundUp8(N)] then bias, Parameters.h:159-176)
    off = 0
    n_in = g.dS
    sizes = hidden + [1 + g.dA]
    for d, n_out in enumerate(sizes):
        K, Kp, N, Np, fImg, bImg, gN, gPart = [int(x) for x in p["dense"][d]]
        assert (K, N) == (n_in, n_out) and Kp % 16 == 0 and Np in (16, 32, 64, 128) and Kp >= K and Np >= N
        ld = (n_out + 7) // 8 * 8
        for k, n in ((0, 0), (K - 1, N - 1), (K // 2, N // 3)):
            w = blob[off + k * ld + n]
            assert hi[fImg + ((k >> 2) * Np + n) * 4 + (k & 3)] == _tf32_hi(np.array([w], np.float32))[0]
            if d >= 1:
                assert bh[bImg + ((n >> 2) * Kp + k) * 4 + (n & 3)] == _tf32_hi(np.array([w], np.float32))[0]
            assert idx[0][off + k * ld + n] == (gPart + n * 128 + k if d == nD - 1 else gPart + k * 128 + n)
        off += (ld * n_in + 7) // 8 * 8 + (n_out + 7) // 8 * 8
        if 0 < d < nD - 1:
            off += 2 * ((n_out + 7) // 8 * 8)        # ParametricResidual w, b after every hidden layer but the first
        n_in = n_out


def test_wide_step_plan_refuses_what_it_does_not_cover(built_library):
    from smarties_b200.learner import make_config
    for dS, dA, settings in ((8, 2, {"nnType": "LSTM", "nnLayerSizes": [32]}),                     # recurrent cells
                             (8, 2, {"nnLayerSizes": [256, 256]}),                                 # wider than one MMA tile
                             (8, 12, {"nnLayerSizes": [64, 64]}),                                  # more than 8 action components
                             (8, 2, {"nnLayerSizes": [32, 32, 32, 32]})):                          # more than 4 dense layers
        cfg, _ = make_config(dS, dA, settings)
        assert _wide_plan(cfg)["rc"] == 0, settings
    cfg, _ = make_config(32, 8, {"nnLayerSizes": [128, 128]})
    p = _wide_plan(cfg)
    assert p["rc"] == 1 and int(p["info"][0]) == 3 and int(p["info"][9]) == 2       # cfg2: three dense layers, two image stages
    cfg, _ = make_config(32, 8, {"learner": "RACER", "nnLayerSizes": [128, 128]})    # RACER: 26 dense outputs + 8 ParamLayer outputs
    p = _wide_plan(cfg)
    assert p["rc"] == 1 and int(p["info"][10]) == 48 and int(p["dense"][2][3]) == 32 and max(int(x) for x in p["info"][6:9]) <= 227 * 1024
