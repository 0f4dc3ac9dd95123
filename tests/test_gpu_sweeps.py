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
