"""GPU: the CUDA learner (through the C-ABI) against the reference's golden vectors and the
numpy oracle on identical replay buffers, weights and sampler seed.

Bars: sampled indices, far-policy counts, Cmax bit-exact; beta to 1e-12 (same f64 formula
driven by the integer count); floats within the stated f32 tolerances (the GPU sums the
mini-batch gradient in batch order with FMAs, the reference without)."""
import os

import numpy as np
import pytest

from parity_utils import CASES, RECURRENT_CASES, SLOW_CASES, THREADED_CASES, Golden, make_learner, make_oracle, relerr

pytestmark = pytest.mark.gpu

TOL_O = 5e-6       # network outputs, absolute (values are O(1e-3..1))
TOL_G = 5e-5       # output gradient, relative to max |g|
TOL_GRAD = 1e-4    # summed parameter gradient, relative to max |G|
TOL_W = 5e-6       # weights after Adam, absolute


def _check_step(L, g, R, pre, stats):
    O, gg, X = L.get_last_batch()
    # identical standardized inputs <=> identical sampled rows and identical scaling
    assert np.array_equal(X, R[pre + "/S"]), "sampled transitions differ from the reference's"
    assert np.abs(O - R[pre + "/O"]).max() < TOL_O
    assert relerr(gg, R[pre + "/g"]) < TOL_G
    if pre + "/gradSum" in R:
        assert relerr(L.get_grad(), R[pre + "/gradSum"]) < TOL_GRAD
        assert np.abs(L.get_weights() - R[pre + "/weights"]).max() < TOL_W
    ref = g.refer(pre + "/post")
    assert stats["cmax"] == ref[1] and stats["cinv"] == ref[2]
    assert stats["n_far_policy"] == int(ref[3])
    assert stats["beta"] == pytest.approx(ref[0], rel=1e-12)
    assert stats["avg_kl"] == pytest.approx(ref[4], rel=2e-4, abs=1e-9)
    assert stats["avg_sq_err"] == pytest.approx(ref[5], rel=2e-4)
    ids, rows, _ = L.read_episodes()
    assert list(ids) == list(R[pre + "/post/epID"]) and list(rows) == list(R[pre + "/post/epLen"])


def _check_final(L, R):
    assert np.allclose(L.read_field("QRET"), R["final/Qret"], rtol=2e-5, atol=2e-5)
    assert np.allclose(L.read_field("V"), R["final/V"], rtol=2e-5, atol=2e-6)
    assert np.allclose(L.read_field("RHO"), R["final/rho"], rtol=2e-5)
    assert np.allclose(L.read_field("KL"), R["final/KL"], rtol=2e-4, atol=1e-7)
    assert np.allclose(L.read_field("DELTA"), R["final/delta"], rtol=2e-5, atol=2e-5)
    mean, scale, std, rew = L.get_scaling()
    assert np.allclose(mean, R["final/stateMean"], atol=1e-7)
    assert np.allclose(scale, R["final/stateScale"], rtol=1e-6)
    assert np.allclose(rew[:2], R["final/rewards"][:2], rtol=1e-6, atol=1e-8)
    _, _, agg = L.read_episodes()
    assert np.allclose(agg[:, :8], R["final/epAgg"][:, :8], rtol=2e-4, atol=2e-5)


@pytest.mark.parametrize("case", CASES + RECURRENT_CASES)
def test_initialize_learner_matches_reference(case):
    g = Golden(case)
    L = make_learner(g)
    R = g.ref
    mean, scale, std, rew = L.get_scaling()
    assert np.allclose(mean, R["init/stateMean"], atol=1e-7)
    assert np.allclose(scale, R["init/stateScale"], rtol=1e-6)
    assert np.allclose(rew, R["init/rewards"], rtol=1e-6, atol=1e-8)
    assert np.allclose(L.read_field("QRET"), R["init/Qret"], rtol=2e-5, atol=2e-5)
    assert np.array_equal(L.read_field("DELTA"), R["init/delta"])
    assert np.array_equal(L.read_field("RHO"), R["init/rho"])
    st = L.get_stats()
    assert st["beta"] == R["init/refer"][0] and st["cmax"] == R["init/refer"][1] and st["cinv"] == R["init/refer"][2]
    L.close()


@pytest.mark.parametrize("case", CASES + RECURRENT_CASES + THREADED_CASES + SLOW_CASES)
def test_learner_steps_match_reference(case):
    """Every step of the golden run.  Among CASES: vracer_da1 — one action component, clipImpWeight < 1: the reference's
    `Uint += float` far-policy count wraps through x86's cvttss2si (uint_plus_float_x86, csrc/common.cuh); vracer_explore —
    "returnsEstimator": "retraceExplore" (k_sweep_explore); vracer_b1024 — several P1 tiles per CTA.  THREADED_CASES: goldens
    of a reference that ran 8 / 16 OpenMP threads; the far-policy count is reduced over per-thread `Uint += float` partials
    (MemoryProcessing.cpp:202-227) and the device learner is built with refer_reduce_threads = that thread count.
    SLOW_CASES: prioritized samplers (PERrank / PERerr / PERseq: std::discrete_distribution over TD errors the device wrote the
    step before) and the farpolfrac / maxkldiv / minerror episode filters (std::sort of the episode vector by device-resident
    aggregates every step, pruning): sampled (episode, t) and episode order identical to the reference run."""
    g = Golden(case)
    L = make_learner(g)
    for s in range(g.steps):
        st = L.train_steps(1)[0]
        _check_step(L, g, g.ref, f"s{s}", st)
    _check_final(L, g.ref)
    L.close()


@pytest.mark.parametrize("case", ["vracer_small", "vracer_prune"])
def test_sampler_indices_bit_exact(case):
    """smb200_sample == Sample_uniform::sample + IDtoSeqStep of the reference."""
    g = Golden(case)
    L = make_learner(g)
    ids, _, _ = L.read_episodes()
    pos, t = L.sample_minibatch()
    assert np.array_equal(ids[pos], g.ref["s0/sampledEpID"])
    assert np.array_equal(t, g.ref["s0/sampledT"])
    L.close()


@pytest.mark.parametrize("case", ["vracer_small", "vracer_bounded", "racer_lstm"])
def test_multi_step_call_equals_single_steps(case):
    """One C-ABI call for the whole run (persistent kernel over many steps) gives bit-identical
    results to step-by-step calls."""
    g = Golden(case)
    A, Bm = make_learner(g), make_learner(g)
    sa = [A.train_steps(1)[0] for _ in range(g.steps)]
    sb = Bm.train_steps(g.steps)
    assert sa == sb
    assert np.array_equal(A.get_weights(), Bm.get_weights())
    assert np.array_equal(A.read_field("QRET"), Bm.read_field("QRET"))
    assert np.array_equal(A.read_field("RHO"), Bm.read_field("RHO"))
    _check_final(Bm, g.ref)
    A.close(); Bm.close()


@pytest.mark.parametrize("case", ["vracer_small", "vracer_lstm2", "racer_mgu", "racer_discrete"])
def test_two_kernel_mode_equals_persistent(monkeypatch, case):
    g = Golden(case)
    A = make_learner(g)
    monkeypatch.setenv("SMB200_MODE", "two")
    Bm = make_learner(g)
    monkeypatch.delenv("SMB200_MODE")
    sa, sb = A.train_steps(g.steps), Bm.train_steps(g.steps)
    if g.settings.get("nnType", "FFNN") == "LSTM":
        # the persistent kernel contracts the LSTM weight gradient on the tensor cores (tcgen05, 3xTF32 split), the
        # two-kernel mode on the SIMT tiles: same mathematics, different rounding of the 4-byte sums
        for x, y in zip(sa, sb):
            assert x["n_far_policy"] == y["n_far_policy"] and x["grad_step"] == y["grad_step"]
            assert x["beta"] == pytest.approx(y["beta"], rel=1e-12) and x["avg_sq_err"] == pytest.approx(y["avg_sq_err"], rel=1e-5)
        assert np.abs(A.get_weights() - Bm.get_weights()).max() < 1e-6
    else:
        assert sa == sb
        assert np.array_equal(A.get_weights(), Bm.get_weights())
    A.close(); Bm.close()


# discrete actions and hidden-layer functions other than Tanh: tile kernel only
FEED_FORWARD_CASES = [c for c in CASES + THREADED_CASES if c not in ("racer_discrete", "vracer_softsign", "vracer_hardsign", "racer_sigm", "vracer_relu", "vracer_lrelu",
                                                                      "vracer_expplus", "racer_softplus", "vracer_exp", "vracer_linear",
                                                                      "vracer_widen")]      # a residual over a narrower layer: tile kernels only


@pytest.mark.parametrize("case", FEED_FORWARD_CASES)
def test_cluster_kernel_matches_reference(monkeypatch, case):
    """The cluster step kernel (cluster_step.cuh, SMB200_CLUSTER=1: clusters of 4 CTAs with column slices of the network in
    shared memory, layer outputs over distributed shared memory, per-cluster weight-gradient sums) against the same goldens
    and bars as the persistent tile kernel."""
    monkeypatch.setenv("SMB200_CLUSTER", "1")
    g = Golden(case)
    L = make_learner(g)
    for s in range(g.steps):
        st = L.train_steps(1)[0]
        _check_step(L, g, g.ref, f"s{s}", st)
    _check_final(L, g.ref)
    L.close()


@pytest.mark.parametrize("case", ["vracer_small", "vracer_cfg2mini", "racer_bounded", "vracer_b1024"])
def test_cluster_kernel_equals_tile_kernel(monkeypatch, case):
    """SMB200_CLUSTER=1 selects the cluster step kernel for feed-forward nets, the default is the persistent tile kernel.
    Same samples, same integer far-policy counts, floats equal to f32 round-off of the different summation orders."""
    g = Golden(case)
    Bm = make_learner(g)
    monkeypatch.setenv("SMB200_CLUSTER", "1")
    A = make_learner(g)
    monkeypatch.delenv("SMB200_CLUSTER")
    sa, sb = A.train_steps(g.steps), Bm.train_steps(g.steps)
    for x, y in zip(sa, sb):
        assert x["n_far_policy"] == y["n_far_policy"] and x["grad_step"] == y["grad_step"]
        assert x["beta"] == pytest.approx(y["beta"], rel=1e-12) and x["avg_sq_err"] == pytest.approx(y["avg_sq_err"], rel=1e-5)
    assert relerr(A.get_grad(), Bm.get_grad()) < 5e-6
    assert np.abs(A.get_weights() - Bm.get_weights()).max() < 1e-6
    assert np.allclose(A.read_field("QRET"), Bm.read_field("QRET"), rtol=1e-5, atol=1e-6)
    A.close(); Bm.close()


WIDE_CASES = [c for c in CASES + THREADED_CASES if c not in ("racer_discrete", "vracer_widen")]      # feed-forward V-RACER / RACER with continuous actions: what the wide step takes


@pytest.mark.parametrize("case", WIDE_CASES)
def test_wide_step_matches_reference(monkeypatch, case):
    """The large-batch step (wide_step.cuh: tiles of 128 sampled transitions, every dense product as 3xTF32 tcgen05.mma with the
    activations as the A operand in tensor memory, split-K weight gradient on the tensor cores, parallel per-episode records)
    against the same goldens and bars as the tile kernel.  SMB200_WIDE=1 forces it for every batch size (partial tiles);
    vracer_b1024: eight full tiles (the default switches to the wide step at batch 2048)."""
    monkeypatch.setenv("SMB200_WIDE", "1")
    g = Golden(case)
    L = make_learner(g)
    assert L.wide_step_active()
    for s in range(g.steps):
        st = L.train_steps(1)[0]
        _check_step(L, g, g.ref, f"s{s}", st)
    _check_final(L, g.ref)
    L.close()


@pytest.mark.parametrize("case", ["vracer_small", "vracer_cfg2mini", "vracer_bounded", "vracer_b1024", "racer_bounded"])
def test_wide_step_equals_tile_kernel(monkeypatch, case):
    """Same samples, same integer far-policy counts, floats equal to f32 round-off (3xTF32 products, different summation order)."""
    g = Golden(case)
    monkeypatch.setenv("SMB200_WIDE", "0")
    Bm = make_learner(g)
    assert not Bm.wide_step_active()
    monkeypatch.setenv("SMB200_WIDE", "1")
    A = make_learner(g)
    monkeypatch.delenv("SMB200_WIDE")
    sa, sb = A.train_steps(g.steps), Bm.train_steps(g.steps)
    for x, y in zip(sa, sb):
        assert x["n_far_policy"] == y["n_far_policy"] and x["grad_step"] == y["grad_step"]
        assert x["beta"] == pytest.approx(y["beta"], rel=1e-12) and x["avg_sq_err"] == pytest.approx(y["avg_sq_err"], rel=1e-5)
    assert relerr(A.get_grad(), Bm.get_grad()) < 5e-6
    assert np.abs(A.get_weights() - Bm.get_weights()).max() < 1e-6
    assert np.allclose(A.read_field("QRET"), Bm.read_field("QRET"), rtol=1e-5, atol=1e-6)
    assert np.array_equal(A.read_field("RHO"), Bm.read_field("RHO")) or np.allclose(A.read_field("RHO"), Bm.read_field("RHO"), rtol=1e-5)
    A.close(); Bm.close()


@pytest.mark.parametrize("case", ["vracer_small", "vracer_bounded"])
def test_injected_samples_and_oracle_flags(case):
    """Feed the oracle's samples to the GPU step by step; with the SAME network outputs the
    far-policy flags and rho must agree: compare the oracle evaluated on the GPU's own outputs."""
    import vracer_oracle as vo
    g = Golden(case)
    L, o = make_learner(g), make_oracle(g)
    for s in range(g.steps):
        seq, obs = o.sample()
        beta, cmax, cinv = o.beta, o.cmax, o.cinv
        eps = [o.episodes[int(k)] for k in seq]
        act = np.stack([e.A[int(t)] for e, t in zip(eps, obs)])
        mu = np.stack([e.MU[int(t)] for e, t in zip(eps, obs)])
        qret = np.array([e.Q[int(t)] for e, t in zip(eps, obs)], np.float32)
        st = L.train_step_on(seq, obs)
        O, gg, _ = L.get_last_batch()
        r = vo.vracer_sample_math(O, act, mu, qret, beta, cmax, cinv, o.bounded)
        # output gradient of the GPU == oracle math on the GPU's outputs (f64 -> f32)
        assert relerr(gg, r["g"].astype(np.float32)) < 1e-6
        # importance weights written to the replay rows: bit-exact f32 of the f64 result
        ids, rows, _ = L.read_episodes()          # the GPU's episode order AFTER the step's FIFO sort
        off = dict(zip(ids.tolist(), np.concatenate([[0], np.cumsum(rows)[:-1]]).tolist()))
        idx = np.array([off[e.ID] + int(t) for e, t in zip(eps, obs)])
        rho_gpu = L.read_field("RHO")[idx]
        assert np.array_equal(rho_gpu, r["rho"].astype(np.float32))
        o.train_step(seq, obs)
        assert st["n_far_policy"] == o.n_far_policy
        assert st["beta"] == pytest.approx(o.beta, rel=1e-12)
    L.close()


def test_forward_matches_oracle():
    import vracer_oracle as vo
    g = Golden("vracer_cfg2mini")
    L, o = make_learner(g), make_oracle(g)
    rng = np.random.default_rng(0)
    S = rng.standard_normal((37, g.dS)).astype(np.float32)
    X = ((S - o.state_mean) * o.state_scale).astype(np.float32)
    O_ref, _ = o.net.forward(o.W, X)
    assert np.abs(L.forward(S) - O_ref).max() < TOL_O
    L.close()


@pytest.mark.parametrize("case", ["vracer_cfg2mini", "vracer_prune", "racer_small_t8", "racer_lstm", "vracer_cfg2mini_t16"])
def test_launch_resident_statistics_equal_the_full_scan(monkeypatch, case):
    """Persistent kernel: from the second step of a launch on the statistics CTA updates its sums by what the step's samples
    changed, takes the maxima as monotone and walks the far-policy terms kept in shared memory (stats_incremental) instead of
    scanning every episode's aggregates (SMB200_STATS_FULL=1).  Far-policy count, beta and the maxima must be identical, the
    f64 sums equal to round-off; weights and replay values identical."""
    g = Golden(case)
    monkeypatch.setenv("SMB200_STATS_FULL", "1")
    A = make_learner(g)
    monkeypatch.delenv("SMB200_STATS_FULL")
    Bm = make_learner(g)
    for k in (1, 7, 50, 300, 2, 990 - 360 + 25):       # the last call crosses the every-1000-steps sweep
        sa, sb = A.train_steps(k), Bm.train_steps(k)
        for i, (x, y) in enumerate(zip(sa, sb)):
            assert x["n_far_policy"] == y["n_far_policy"] and x["n_far_exact"] == y["n_far_exact"], (k, i)
            assert x["beta"] == y["beta"] and x["grad_step"] == y["grad_step"], (k, i)
            assert x["max_q"] == y["max_q"] and x["min_q"] == y["min_q"] and x["max_abs_err"] == pytest.approx(y["max_abs_err"], rel=1e-12), (k, i)
            for key in ("avg_kl", "avg_sq_err", "avg_q", "stdev_q", "avg_return"):
                assert x[key] == pytest.approx(y[key], rel=1e-9, abs=1e-12), (k, i, key)
    assert np.array_equal(A.get_weights(), Bm.get_weights())
    assert np.array_equal(A.read_field("RHO"), Bm.read_field("RHO"))
    assert np.array_equal(A.read_episodes()[2], Bm.read_episodes()[2])
    A.close(); Bm.close()


@pytest.mark.parametrize("case", ["vracer_cfg2mini", "vracer_prune", "racer_lstm"])
def test_sample_ahead_queue_is_invisible(monkeypatch, case):
    """smb200_train_steps draws the next call's mini-batches while it waits for the device (sample-ahead queue).  The sampled
    stream, the step counters and the weights must be those of a learner that never draws ahead: calls of many sizes, the
    every-1000-steps boundary, an episode pushed between two calls (the queue is dropped), a re-seeded sampler, the
    benchmark's presample path."""
    g = Golden(case)
    monkeypatch.setenv("SMB200_NO_SAMPLE_AHEAD", "1")
    A = make_learner(g)
    monkeypatch.delenv("SMB200_NO_SAMPLE_AHEAD")
    Bm = make_learner(g)
    rng = np.random.default_rng(5)
    ep_len = 12
    S = rng.standard_normal((ep_len, g.dS)).astype(np.float32)
    Aa = rng.standard_normal((ep_len, g.dA)).astype(np.float32) * 0.1
    MU = np.concatenate([Aa * 0.9, np.full((ep_len, g.dA), 0.3, np.float32)], axis=1)
    R = rng.standard_normal(ep_len).astype(np.float32)

    def both(f):
        ra, rb = f(A), f(Bm)
        return ra, rb

    sizes = [1, 1, 3, 20, 1, 64, 5, 200, 2, 990 - 297, 30, 1, 7]      # crosses grad step 1000 inside the 30-step call
    for i, k in enumerate(sizes):
        sa, sb = both(lambda L: L.train_steps(k))
        for x, y in zip(sa, sb):
            assert x["n_far_policy"] == y["n_far_policy"] and x["beta"] == y["beta"], (i, k)
        (Oa, ga, Xa), (Ob, gb, Xb) = both(lambda L: L.get_last_batch())
        assert np.array_equal(Xa, Xb) and np.array_equal(Oa, Ob), (i, k)
        if i == 4:
            both(lambda L: L.push_episode(10_000, S, Aa, MU, R, True))
        if i == 6:
            both(lambda L: L.seed_sampler(99))
        if i == 8:
            both(lambda L: (L.presample(6), L.train_presampled(0, 6), L.sync()))
    assert np.array_equal(A.get_weights(), Bm.get_weights())
    assert np.array_equal(A.read_field("RHO"), Bm.read_field("RHO"))
    A.close(); Bm.close()


@pytest.mark.parametrize("case", ["racer_lstm", "vracer_lstm2", "racer_mgu", "vracer_gru2", "racer_cfg3mini", "vracer_cfg2mini"])
def test_forward_seq_matches_oracle(case):
    """Actor-side policy evaluation (RACER::selectAction, RACER.cpp:30-47) on the window MemoryBuffer::agentToMinibatch
    builds (MemoryBuffer.cpp:440-467): ragged windows of 1 .. nnBPTTseq + 3 raw states per agent, zero initial recurrent
    state, outputs at the newest state; windows longer than nnBPTTseq + 1 are cut to their newest states.  Feed-forward
    nets see the newest state only."""
    g = Golden(case)
    L, o = make_learner(g), make_oracle(g)
    rng = np.random.default_rng(3)
    recurrent = g.settings.get("nnType", "FFNN") != "FFNN"
    bptt = g.settings.get("nnBPTTseq", 16)
    n, max_len = 19, bptt + 3
    W = rng.standard_normal((n, max_len, g.dS)).astype(np.float32)
    lens = rng.integers(1, max_len + 1, n).astype(np.int32)
    lens[0], lens[1], lens[2] = 1, max_len, bptt + 1
    out = L.forward_seq(W, lens)
    assert out.shape == (n, L.n_out)
    for i in range(n):
        keep = min(int(lens[i]), bptt + 1) if recurrent else 1
        X = ((W[i, lens[i] - keep:lens[i]] - o.state_mean) * o.state_scale).astype(np.float32)
        O_ref = o.net.forward_seq(o.W, X)[0][-1] if recurrent else o.net.forward(o.W, X)[0][-1]
        assert np.abs(out[i] - O_ref).max() < TOL_O, (i, int(lens[i]))
    if not recurrent:
        assert np.array_equal(out, L.forward(W[np.arange(n), lens - 1]))
    # a second call reuses the staging buffers; a larger one grows them
    W2 = rng.standard_normal((300, max_len, g.dS)).astype(np.float32)
    l2 = np.full(300, max_len, np.int32)
    out2 = L.forward_seq(W2, l2)
    assert np.array_equal(out2[:5], L.forward_seq(W2[:5], l2[:5]))
    with pytest.raises(Exception):
        L.forward_seq(W, np.zeros(n, np.int32))
    L.close()


@pytest.mark.parametrize("case", ["racer_lstm", "vracer_lstm2"])
def test_tensor_core_weight_gradient_matches_simt_tiles(monkeypatch, case):
    """Recurrent nets: the tcgen05 contraction of the LSTM weight gradient (3xTF32, accumulator in tensor memory, K-slices
    added in fixed order) against the SIMT tiles of the same persistent kernel (SMB200_TC=0): the summed parameter
    gradient agrees to f32 round-off, the run stays inside every reference tolerance either way."""
    g = Golden(case)
    A = make_learner(g)
    monkeypatch.setenv("SMB200_TC", "0")
    Bm = make_learner(g)
    monkeypatch.delenv("SMB200_TC")
    A.train_steps(1); Bm.train_steps(1)
    ga, gb = A.get_grad(), Bm.get_grad()
    assert np.abs(ga).max() > 0 and relerr(ga, gb) < 2e-6
    A.close(); Bm.close()


def test_reference_settings_files_train_on_the_device():
    """Every V-RACER / RACER settings file of the reference (tests/reference_settings.py) constructs a device learner for a
    HalfCheetah-shaped MDP (17 states, 6 actions) and trains: finite statistics, beta in (0, 1], weights that move, and the
    step counter of the reference.  Buffer sizes are cut to a small synthetic replay; everything else is the file's."""
    from reference_settings import DEVICE
    from smarties_b200 import Learner, synth
    d = synth.make_replay(77, 60, (40, 80), 17, 6)
    n_data = int((d["N"] - 1).sum())
    for name, js in DEVICE.items():
        S = dict(js, maxTotObsNum=8192, minTotObsNum=min(n_data, 2048))
        S["batchSize"] = min(int(S.get("batchSize", 256)), 64)
        L = Learner(17, 6, S, bounded=True)
        L.load_replay(d)
        L.initialize_learner()
        w0 = L.get_weights().copy()
        st = L.train_steps(12)
        assert st[-1]["grad_step"] == 12, name
        assert 0.0 < st[-1]["beta"] <= 1.0 and all(np.isfinite([st[-1][k] for k in ("avg_kl", "avg_sq_err", "avg_q", "stdev_q")])), name
        w1 = L.get_weights()
        assert np.isfinite(w1).all() and np.abs(w1 - w0).max() > 0, name
        out = L.forward_seq(d["S"][:3][None].repeat(2, axis=0), [1, 3])
        assert out.shape == (2, L.n_out) and np.isfinite(out).all(), name
        L.close()
