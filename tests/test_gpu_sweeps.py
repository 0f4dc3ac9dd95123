"""GPU: the whole-buffer sweeps (Retrace return recursion, reward/state moments) against numpy
and through size-independent properties at larger sizes."""
import numpy as np
import pytest

from smarties_b200 import synth

pytestmark = pytest.mark.gpu


def _learner(d, **kw):
    from smarties_b200 import Learner
    L = Learner(d["dS"], d["dA"], {"nnLayerSizes": [32, 32], "batchSize": 64, "maxTotObsNum": int(d["N"].sum()) + 64}, **kw)
    L.load_replay(d)
    return L


def _retrace_numpy(d, V, ADV, RHO, rmean, rscale, gamma=np.float32(0.995), lam=np.float32(1.0)):
    out = np.zeros(int(d["N"].sum()), np.float32)
    for e in range(len(d["N"])):
        o, N = int(d["start"][e]), int(d["N"][e])
        q = np.zeros(N, np.float32)
        q[N - 1] = 0 if d["term"][e] else V[o + N - 1]
        R = d["R"][o:o + N].astype(np.float64).copy(); R[0] = 0
        rs = ((R - float(rmean)) * float(rscale)).astype(np.float32)
        for t in range(N - 2, -1, -1):
            w = min(RHO[o + t + 1], np.float32(1))
            q[t] = rs[t + 1] + gamma * (V[o + t + 1] + (lam * w) * ((q[t + 1] - ADV[o + t + 1]) - V[o + t + 1]))
        out[o:o + N] = q
    return out


@pytest.mark.parametrize("shape", [dict(n_ep=7, ep_len=(1, 5)), dict(n_ep=33, ep_len=(31, 34)), dict(n_ep=20, ep_len=(60, 400))])
def test_retrace_sweep_ragged_episodes(shape):
    """Episode lengths around the 32-step warp chunk, single-step episodes, long episodes."""
    d = synth.make_replay(5, dS=5, dA=2, **shape)
    L = _learner(d)
    L.initialize_learner()
    mean, scale, std, rew = L.get_scaling()
    V, ADV, RHO = L.read_field("V"), L.read_field("ADV"), L.read_field("RHO")
    q_ref = _retrace_numpy(d, V, ADV, RHO, rew[0], rew[1])
    # the warp-scan composes 32 steps at a time: a few ulp(|Q|max) per chunk, geometric decay
    assert np.allclose(L.read_field("QRET"), q_ref, rtol=1e-5, atol=5e-5)
    # idempotence: a second sweep over unchanged V/rho changes nothing
    q1 = L.read_field("QRET")
    err2 = L.retrace_sweep()
    assert np.array_equal(L.read_field("QRET"), q1) and err2 == 0.0
    L.close()


def test_moments_match_numpy_and_ignore_terminal_rows():
    d = synth.make_replay(11, n_ep=50, ep_len=(3, 90), dS=17, dA=3)
    L = _learner(d)
    m = L.reward_state_moments()
    dS = 17
    keep_s = np.ones(len(d["R"]), bool); keep_s[d["start"] + d["N"] - 1] = False
    keep_r = np.ones(len(d["R"]), bool); keep_r[d["start"]] = False
    S = d["S"][keep_s].astype(np.float64)
    R = d["R"][keep_r].astype(np.float64)
    assert m[2 * dS] == keep_s.sum()
    assert np.allclose(m[:dS], S.sum(0), rtol=1e-12, atol=1e-9)
    assert np.allclose(m[dS:2 * dS], (S * S).sum(0), rtol=1e-12)
    assert np.allclose(m[2 * dS + 1:], [R.sum(), (R * R).sum()], rtol=1e-12, atol=1e-9)
    L.close()


@pytest.mark.parametrize("dS", [1, 3, 32, 100, 300])
def test_moments_any_state_width(dS):
    d = synth.make_replay(2, n_ep=9, ep_len=(5, 40), dS=dS, dA=1)
    L = _learner(d)
    m = L.reward_state_moments()
    keep_s = np.ones(len(d["R"]), bool); keep_s[d["start"] + d["N"] - 1] = False
    S = d["S"][keep_s].astype(np.float64)
    assert np.allclose(m[:dS], S.sum(0), rtol=1e-12, atol=1e-9)
    assert np.allclose(m[dS:2 * dS], (S * S).sum(0), rtol=1e-12)
    L.close()


def test_full_size_buffer_properties():
    """BASELINE cfg2 size (1M transitions, dS 32, dA 8): linearity of Retrace in the rewards
    scale and idempotence; checksum of moments against numpy."""
    d = synth.make_replay(123, n_ep=1000, ep_len=1000, dS=32, dA=8)
    from smarties_b200 import Learner
    L = Learner(32, 8, {"maxTotObsNum": 1048576, "minTotObsNum": 1000000})
    L.load_replay(d)
    assert L.n_transitions == 1_000_000
    m = L.reward_state_moments()
    keep_s = np.ones(len(d["R"]), bool); keep_s[d["start"] + d["N"] - 1] = False
    assert m[64] == 1_000_000
    assert np.allclose(m[:32], d["S"][keep_s].astype(np.float64).sum(0), rtol=1e-10, atol=1e-6)
    L.initialize_learner()
    q1 = L.read_field("QRET").copy()
    assert L.retrace_sweep() == 0.0                       # idempotent
    # with V = A = 0 and rho = 1 Retrace is the discounted reward-to-go: check one episode exactly
    mean, scale, std, rew = L.get_scaling()
    o, N = int(d["start"][3]), int(d["N"][3])
    rs = ((d["R"][o:o + N].astype(np.float64) - float(rew[0])) * float(rew[1])).astype(np.float32)
    q = np.zeros(N, np.float32)
    for t in range(N - 2, -1, -1):
        q[t] = rs[t + 1] + np.float32(0.995) * q[t + 1]
    assert np.allclose(q1[o:o + N], q, rtol=2e-5, atol=5e-5)
    # a few learner steps on the full buffer run and keep the integer counters consistent
    st = L.train_steps(3)
    assert st[-1]["grad_step"] == 3 and 0 <= st[-1]["n_far_exact"] <= 3 * 256
    L.close()


@pytest.mark.parametrize("shape", [dict(n_ep=9, ep_len=(1, 6), dS=4), dict(n_ep=40, ep_len=(100, 300), dS=8),
                                   dict(n_ep=12, ep_len=(900, 2300), dS=32), dict(n_ep=300, ep_len=(20, 40), dS=32)])
def test_fused_sweep_equals_separate_kernels(shape):
    """k_sweep_fused (Retrace + exact aggregates + reward / state moments in one pass, state rows through the cp.async.bulk
    ring) against k_sweep(recompute) + k_moments on identical buffers: Retrace estimates bit-identical (same scan, same
    operation order), moments equal to f64 round-off of the different summation order, aggregates to f32 round-off, the
    integer far-policy count exact.  Shapes: one-step episodes, episodes longer than the 1024-step super-chunk and than one
    16 KB state tile, many short episodes (several per CTA)."""
    dS = shape.pop("dS")
    d = synth.make_replay(21, dA=2, dS=dS, **shape)
    A, Bm = _learner(d), _learner(d)
    rng = np.random.default_rng(3)
    rows = int(d["N"].sum())
    for L in (A, Bm):        # non-trivial values in every array the sweep reads
        L.initialize_learner()
    for name, gen in (("V", lambda: rng.standard_normal(rows)), ("ADV", lambda: 0.1 * rng.standard_normal(rows)),
                      ("RHO", lambda: np.exp(0.7 * rng.standard_normal(rows))), ("KL", lambda: rng.random(rows)),
                      ("DELTA", lambda: rng.standard_normal(rows))):
        v = gen().astype(np.float32)
        A.write_field(name, v); Bm.write_field(name, v)
    eA, mA = A.fused_sweep()
    eB = Bm.retrace_sweep()
    mB = Bm.reward_state_moments()
    assert np.array_equal(A.read_field("QRET"), Bm.read_field("QRET"))
    assert eA == pytest.approx(eB, rel=1e-5)
    assert mA[2 * dS] == mB[2 * dS]
    assert np.allclose(mA, mB, rtol=1e-11, atol=1e-9)
    # aggregates: the fused kernel recomputed them; a learner step's every-1000-steps path is covered by the step goldens
    _, _, aggA = A.read_episodes()
    V, ADV, RHO, KL, DL = (A.read_field(k) for k in ("V", "ADV", "RHO", "KL", "DELTA"))
    cm = 1 + np.sqrt(2 / 2.0)        # Cmax at gradient step 0 with dA = 2 (clipImpWeight = sqrt(dA / 2)), epsAnneal irrelevant at step 0
    o = 0
    for e, N in enumerate(d["N"]):
        N = int(N); nd = N - 1
        w, dl = RHO[o:o + nd], DL[o:o + nd]
        far = np.count_nonzero((w > np.float32(cm)) | (w < np.float32(1 / cm)))
        assert aggA[e, 1] == pytest.approx(far / nd, rel=1e-6)
        assert aggA[e, 0] == pytest.approx(KL[o:o + N].astype(np.float64).sum() / nd, rel=2e-5)
        assert aggA[e, 2] == pytest.approx((dl.astype(np.float64) ** 2).sum() / nd, rel=2e-5)
        q = (ADV[o:o + nd] + V[o:o + nd])
        assert aggA[e, 6] == q.max() and aggA[e, 7] == q.min()
        o += N
    A.close(); Bm.close()
