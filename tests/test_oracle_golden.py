"""CPU: the numpy oracle (oracle/vracer_oracle.py) against golden vectors produced by the
reference itself (tests/golden/*.npz, generator tests/golden/make_golden.py).

Tolerances: the reference is built with -ffast-math (vectorised approximate sqrt/div in Adam,
libmvec exp in tanh) and sums in a different association than numpy, so floats agree to f32
round-off; sampled indices, far-policy counts and the ReF-ER coefficient are exact."""
import numpy as np
import pytest

from parity_utils import CASES, ORACLE_ONLY_CASES, RECURRENT_CASES, SLOW_CASES, TARGET_CASES, THREADED_CASES, Golden, make_oracle, relerr


@pytest.mark.parametrize("case", CASES + RECURRENT_CASES + ORACLE_ONLY_CASES + THREADED_CASES + SLOW_CASES)
def test_oracle_matches_reference(case):
    g = Golden(case)
    o = make_oracle(g)
    R = g.ref
    # initializeLearner: scaling, Retrace after rescale, beta
    assert np.array_equal(o.state_mean, R["init/stateMean"])
    assert np.allclose(o.state_scale, R["init/stateScale"], rtol=1e-6, atol=0)
    assert np.allclose([o.rew_mean, o.rew_scale], R["init/rewards"][:2], rtol=1e-6)
    assert np.allclose(o.concat("Q"), R["init/Qret"], rtol=1e-6, atol=1e-6)
    assert o.beta == R["init/refer"][0]
    for s in range(g.steps):
        order = np.array([e.ID for e in o.episodes])          # the episode vector the sampler walks (before this step's re-sort)
        r = o.train_step()
        pre = f"s{s}"
        # sampled (episode, time step): bit-exact
        assert np.array_equal(r["obs"], R[pre + "/sampledT"])
        assert np.array_equal(order[r["seq"]], R[pre + "/sampledEpID"]) and len(r["seq"]) == g.B
        assert np.abs(r["O"] - R[pre + "/O"]).max() < 2e-6
        assert relerr(r["g"], R[pre + "/g"]) < 2e-5
        if pre + "/gradSum" in R:
            assert relerr(r["gradSum"], R[pre + "/gradSum"]) < 2e-5
            assert np.abs(o.W - R[pre + "/weights"]).max() < 2e-6
        ref = g.refer(pre + "/post")
        assert o.beta == pytest.approx(ref[0], rel=1e-12)
        assert o.cmax == pytest.approx(ref[1], rel=1e-15)
        assert o.n_far_policy == int(ref[3])                       # integer far-policy count: exact
        assert [e.ID for e in o.episodes] == list(R[pre + "/post/epID"])
    assert np.allclose(o.concat("Q"), R["final/Qret"], rtol=1e-5, atol=1e-5)
    assert np.allclose(o.concat("V"), R["final/V"], rtol=1e-5, atol=1e-6)
    assert np.allclose(o.concat("rho"), R["final/rho"], rtol=1e-5)
    assert np.allclose(o.concat("KL"), R["final/KL"], rtol=1e-4, atol=1e-7)
    assert np.allclose(o.concat("delta"), R["final/delta"], rtol=1e-5, atol=1e-5)
    assert np.allclose(o.state_mean, R["final/stateMean"], atol=1e-7)
    assert np.allclose(o.state_scale, R["final/stateScale"], rtol=1e-6)
    agg = np.array([[e.avgKL, e.fracFar, e.avgSqErr, e.maxAbsErr, e.sumQ2, e.sumQ, e.maxQ, e.minQ] for e in o.episodes])
    assert np.allclose(agg, R["final/epAgg"][:, :8], rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("case", TARGET_CASES)
def test_oracle_target_weights_match_reference(case):
    """"targetDelay" > 0 (AdamOptimizer::apply_update, Optimizer.cpp:162-177): the target weights after every update — the
    exponential average (0.05) and the periodic copy (3: updates 1, 4, 7) — against the reference's own."""
    g = Golden(case)
    o = make_oracle(g)
    R = g.ref
    seen = []
    for s in range(g.steps):
        o.train_step()
        assert np.abs(o.W_tgt - R[f"s{s}/tgt"]).max() < 2e-6, s
        seen.append(float(np.abs(o.W_tgt - o.W).max()))
    assert np.abs(o.W - R[f"s{g.steps - 1}/weights"]).max() < 2e-6
    if g.settings["targetDelay"] >= 1:
        assert [x == 0.0 for x in seen] == [s % 3 == 0 for s in range(g.steps)]      # copies at updates 1, 4, 7
    else:
        assert min(seen) > 0.0
    # the checkpoint file holds them, padding stripped (Optimizer.cpp:180-197)
    tgt_file = np.frombuffer(bytes(g.ckpt["agent_00_net_tgt_weights.raw"]), np.float32)
    assert np.abs(o.layout.strip_padding(o.W_tgt) - tgt_file).max() < 2e-6


def test_sampler_is_libstdcxx_uniform_int():
    """std::mt19937(5489) first outputs are the published MT19937 known-answer values; Lemire's
    reduction is what libstdc++ 13 uses for 32-bit URNGs (bits/uniform_int_dist.h)."""
    import vracer_oracle as vo
    gen = vo.Mt19937(5489)
    assert [gen() for _ in range(3)] == [3499211612, 581869302, 3890346734]
    gen = vo.Mt19937(7)
    ids = vo.sample_uniform(gen, 925, 16)
    assert len(ids) == 16 and np.all(np.diff(ids) > 0) and ids.min() >= 0 and ids.max() < 925


@pytest.mark.parametrize("case", ["vracer_ckpt", "racer_lstm_ckpt"])
def test_checkpoint_golden_layout(case):
    """The reference's checkpoint files (written by Learner_approximator::save in the golden run) parse with the
    layout the device library implements (include/smarties_b200.h, smb200_save): per episode `size_t N`, then
    N x [state | reward | action | policy], six length-N arrays (Qret, A, V, delta, rho, KL) and a 10-float tail
    {bool terminated; ssize_t ID, just_sampled, agentID}; they hold exactly the reference's final in-memory state."""
    from parity_utils import Golden
    g = Golden(case)
    R, dS, dA = g.ref, g.dS, g.dA
    dP = 2 * dA
    raw = bytes(g.ckpt["agent_00_rank_000_learner_data.raw"])
    pos, out = 0, {k: [] for k in ("Qret", "A", "V", "delta", "rho", "KL", "id", "len", "S0")}
    while pos < len(raw):
        N = int(np.frombuffer(raw, np.uint64, 1, pos)[0]); pos += 8
        tot = (dS + dA + dP + 7) * N + 10
        buf = np.frombuffer(raw, np.float32, tot, pos); pos += 4 * tot
        tup = buf[:(dS + 1 + dA + dP) * N].reshape(N, dS + 1 + dA + dP)
        out["S0"].append(tup[0, :dS])
        rest = buf[(dS + 1 + dA + dP) * N:]
        for i, k in enumerate(("Qret", "A", "V", "delta", "rho", "KL")):
            out[k].append(rest[i * N:(i + 1) * N])
        tail = rest[6 * N:].tobytes()
        out["id"].append(int(np.frombuffer(tail, np.int64, 1, 1)[0])); out["len"].append(N)
        assert int(np.frombuffer(tail, np.int64, 1, 9)[0]) == -1 and int(np.frombuffer(tail, np.int64, 1, 17)[0]) == 0
    assert out["id"] == list(R["final/epID"]) and out["len"] == list(R["final/epLen"])
    for k in ("Qret", "A", "V", "delta", "rho", "KL"):
        assert np.array_equal(np.concatenate(out[k]), R["final/" + k]), k
    sc = np.frombuffer(bytes(g.ckpt["agent_00_scaling.raw"]), np.float64)
    assert np.array_equal(sc[:dS].astype(np.float32), R["final/stateMean"])
    assert np.array_equal(sc[dS:2 * dS].astype(np.float32), R["final/stateScale"])
    assert np.array_equal(sc[3 * dS:].astype(np.float32), R["final/rewards"][[2, 1, 0]])     # file: stdev, scale, mean
    status = bytes(g.ckpt["agent_00_rank_000_learner_status.raw"]).decode()
    assert f"nGradSteps: {g.start_step + g.steps + 1}\n" in status and f"nStoredEps: {len(out['id'])}\n" in status


@pytest.mark.parametrize("case", CASES + RECURRENT_CASES)
def test_grad_stats_file_matches_reference(case):
    """StatsTracker (Utils/StatsTracker.cpp): the `<learner>_<net>_outGrad_stats.raw` file the reference wrote during the
    golden run (tests/golden/outgrad_stats.npz, generator `make_golden.py outgrad_stats`) against (a) the tracker
    restatement fed with the reference's own per-sample output gradients — header and layout exact, values to f32
    round-off of the dumped gradients — and (b) the oracle's run of the same steps."""
    import vracer_oracle as vo
    g = Golden(case)
    want = np.load(g.path("outgrad_stats.npz"))[case]
    n_out = g.ref["s0/g"].shape[1]
    tr = vo.StatsTracker(n_out)
    for s in range(g.steps):
        for row in g.ref[f"s{s}/g"]:
            tr.track_vector(row)
        tr.reduce_stats(g.start_step + s)
    got = tr.file_words()
    printed = [s for s in range(g.steps) if (g.start_step + s) % 1000 == 0]
    header = 1 if printed and printed[0] == 0 else 0
    assert want.size == header + 2 * n_out * len(printed) and got.size == want.size
    if header:
        assert got[0] == want[0] == np.float32(n_out + .1)
    # means cancel over the mini-batch: absolute bar at f32 round-off of the largest entry
    assert np.allclose(got, want, rtol=2e-6, atol=1e-6 * np.abs(want[header:]).max())
    o = make_oracle(g)
    o.grad_stats = vo.StatsTracker(n_out)
    for s in range(g.steps):
        o.train_step()
    mine = o.grad_stats.file_words()
    assert mine.size == want.size and np.abs(mine - want).max() < 2e-5 * max(np.abs(want[header:]).max(), 1e-30)
    assert o.grad_stats.n_step == g.steps


def test_std_sort_restatement_matches_libstdcxx():
    """oracle std_sort against permutations produced by g++'s std::sort / std::partial_sort on keys full of ties
    (tests/golden/std_sort_vectors.npz, generator make_std_sort_vectors.py): the unstable order is reproduced exactly."""
    import vracer_oracle as vo
    z = np.load(Golden.path("std_sort_vectors.npz"))
    names = sorted({k.split("/")[0] for k in z.files})
    assert len(names) == 96
    for name in names:
        keys, perm = z[name + "/keys"], z[name + "/perm"]
        heap = name.startswith("heap")
        if heap and len(keys) <= 16:
            continue                                   # runs of <= 16 never reach the heap-sort branch inside std::sort
        v = [(float(k), i) for i, k in enumerate(keys)]
        vo.std_sort(v, lambda a, b: a[0] < b[0], depth_limit=0 if heap else None)
        assert [i for _, i in v] == list(perm), name
