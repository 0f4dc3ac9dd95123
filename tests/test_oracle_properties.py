"""CPU: finite-difference properties of the oracle, mirroring the reference's own unit tests —
`units/Math/Continuous_policy.cpp:19-60` (policy gradient and KL gradient against the log-probability and the KL
divergence) and `units/Network/Network.cpp:17-173` (Network::backProp against differences of Network::forward,
dense and LSTM through time, tolerance sqrt(FLT_EPSILON)) — plus the same check for the Gaussian advantage head and
the value squashing.  The golden vectors pin the oracle's VALUES to the reference binary; these pin its internal
consistency the way the reference pins its own."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "oracle"))
import vracer_oracle as vo  # noqa: E402

f32, f64 = np.float32, np.float64


def _case(rng, Bn, dA, racer):
    """Behaviour policy and action close enough to the current policy that log(rho) stays inside the [-7, 7] clip
    (Continuous_policy.h:648-653) — the finite differences below go through rho."""
    m0 = 2 + 2 * dA if racer else 1
    nOut = m0 + 2 * dA
    O = rng.normal(0, 0.4, (Bn, nOut))
    mean, stdev = O[:, m0:m0 + dA], vo.softplus(O[:, m0 + dA:])
    act = mean + 0.8 * stdev * rng.normal(0, 1, (Bn, dA))
    mu = np.concatenate([mean + 0.3 * stdev * rng.normal(0, 1, (Bn, dA)), stdev * rng.uniform(0.8, 1.25, (Bn, dA))], axis=1)
    qret = rng.normal(0, 1.0, Bn)
    return O, act, mu, qret


def _fd(fun, O, j, h=1e-6):
    Op, Om = O.copy(), O.copy()
    Op[:, j] += h; Om[:, j] -= h
    return (fun(Op) - fun(Om)) / (2 * h)


@pytest.mark.parametrize("racer", [False, True])
@pytest.mark.parametrize("bounded", [False, True])
def test_policy_and_kl_gradients_match_finite_differences(racer, bounded):
    rng = np.random.default_rng(5 + racer + 2 * bounded)
    Bn, dA = 6, 3
    O, act, mu, qret = _case(rng, Bn, dA, racer)
    bnd = np.full(dA, bounded)
    big = dict(cmax=1e9, cinv=1e-9, bounded=bnd, racer=racer)       # nothing is "far policy", nothing is clipped
    m0 = 2 + 2 * dA if racer else 1
    r1 = vo.vracer_sample_math(O, act, mu, qret, beta=1.0, **big)
    r0 = vo.vracer_sample_math(O, act, mu, qret, beta=0.0, **big)
    fac = (qret - r1["V"]) * np.minimum(1e9, r1["rho"])               # A_RET * min(Cmax, rho)
    for j in range(m0, m0 + 2 * dA):
        # beta = 1: g = fac * d log pi(a) / dO  (Continuous_policy.h:694-701); log rho = log pi - log mu
        dlogpi = _fd(lambda X: np.log(vo.vracer_sample_math(X, act, mu, qret, beta=1.0, **big)["rho"]), O, j)
        assert np.allclose(r1["g"][:, j], fac * dlogpi, rtol=2e-6, atol=1e-8), j
        # beta = 0: g = -d D_KL / dO  (KLDivGradient(MU, -1), :709-716)
        dkl = _fd(lambda X: vo.vracer_sample_math(X, act, mu, qret, beta=0.0, **big)["dkl"], O, j)
        assert np.allclose(r0["g"][:, j], -dkl, rtol=2e-6, atol=1e-8), j
    # value head: g[0] = min(1, rho) * dQ * beta * dV/dO0  (RACER_train.cpp:50, RACER_common.cpp:28-32)
    dV = _fd(lambda X: vo.vracer_sample_math(X, act, mu, qret, beta=1.0, **big)["V"], O, 0)
    assert np.allclose(r1["g"][:, 0], np.minimum(1.0, r1["rho"]) * r1["dq"] * dV, rtol=2e-6, atol=1e-9)
    if racer:   # ADV.grad: err * dA/dO over the advantage outputs (Gaus_advantage.h:88-114)
        err = np.minimum(1e9, r1["rho"]) * r1["dq"]
        for j in range(1, 2 + 2 * dA):
            dA_dO = _fd(lambda X: vo.vracer_sample_math(X, act, mu, qret, beta=1.0, **big)["A"], O, j)
            assert np.allclose(r1["g"][:, j], err * dA_dO, rtol=5e-6, atol=1e-8), j


def test_far_policy_samples_only_carry_the_penalty_gradient():
    """ReF-ER rule 1 (Episode.h:28-33, RACER_train.cpp:50-56): outside [1/C, C] the value and policy-gradient terms vanish,
    the KL penalty stays."""
    rng = np.random.default_rng(11)
    O, act, mu, qret = _case(rng, 8, 2, False)
    near = vo.vracer_sample_math(O, act, mu, qret, beta=0.3, cmax=1e9, cinv=1e-9)
    far = vo.vracer_sample_math(O, act, mu, qret, beta=0.3, cmax=1.0 + 1e-6, cinv=1 / (1.0 + 1e-6))
    pen = vo.vracer_sample_math(O, act, mu, qret, beta=0.0, cmax=1e9, cinv=1e-9)
    assert far["is_far"].all() and not near["is_far"].any()
    assert np.all(far["g"][:, 0] == 0)
    assert np.allclose(far["g"][:, 1:], (1 - 0.3) * pen["g"][:, 1:], rtol=1e-12)
    assert np.array_equal(far["rho"], near["rho"]) and np.array_equal(far["dkl"], near["dkl"])


def _num_grad(loss, blob, idx, h):
    g = np.zeros(len(idx))
    for n, i in enumerate(idx):
        bp, bm = blob.copy(), blob.copy()
        bp[i] += f32(h); bm[i] -= f32(h)
        g[n] = (loss(bp) - loss(bm)) / (float(bp[i]) - float(bm[i]))
    return g


def test_mlp_backprop_matches_finite_differences():
    rng = np.random.default_rng(3)
    lay = vo.MlpLayout(5, [12, 12], 4, 3)             # dense-tanh, dense-tanh + parametric residual, linear out, ParamLayer
    net = vo.MlpNet(lay)
    blob = np.zeros(lay.n_params, f32)
    used = []
    for L in lay.layers:                               # weights only where the layout has parameters (padding stays zero)
        if L["kind"].startswith("dense"):
            for i in range(L["nIn"]):
                used += list(range(L["w"] + i * L["ldw"], L["w"] + i * L["ldw"] + L["nOut"]))
            used += list(range(L["b"], L["b"] + L["nOut"]))
        elif L["kind"] == "residual":
            used += list(range(L["w"], L["w"] + L["n"])) + list(range(L["b"], L["b"] + L["n"]))
        else:
            used += list(range(L["b"], L["b"] + L["n"]))
    blob[used] = rng.normal(0, 0.5, len(used)).astype(f32)
    x = rng.normal(0, 1, (3, 5)).astype(f32)
    gout = rng.normal(0, 1, (3, lay.n_out)).astype(f32)
    O, Y = net.forward(blob, x)
    G = net.backward(blob, Y, gout)
    pick = rng.choice(used, 60, replace=False)
    num = _num_grad(lambda b: float((net.forward(b, x)[0].astype(f64) * gout).sum()), blob, pick, 2e-2)
    scale = np.abs(G[pick]).max()
    assert np.abs(G[pick] - num).max() < 40 * np.sqrt(np.finfo(f32).eps) * scale     # f32 forward passes, h = 2e-2
    assert np.all(G[np.setdiff1d(np.arange(lay.n_params), used)] == 0)               # padding never receives a gradient


@pytest.mark.parametrize("cell,cells", [("LSTM", [6]), ("MGU", [8]), ("MGU", [8, 8])])
def test_recurrent_bptt_matches_finite_differences(cell, cells):
    """units/Network/Network.cpp of the reference checks every layer type this way (LSTM and MGU/GRU among them)."""
    rng = np.random.default_rng(9)
    lay = vo.SeqLayout(4, cells, 3, 2, cell)
    net = vo.SeqNet(lay)
    blob = rng.normal(0, 0.4, lay.n_params).astype(f32)
    X = rng.normal(0, 1, (5, 4)).astype(f32)          # seq_len 5 like units/Network/Network.cpp
    gout = rng.normal(0, 1, lay.n_out).astype(f32)
    O, cache = net.forward_seq(blob, X)
    G = np.zeros(lay.n_params, f32)
    net.backward_seq(blob, cache, gout, G)
    nz = np.flatnonzero(G)
    pick = rng.choice(nz, min(60, nz.size), replace=False)
    num = _num_grad(lambda b: float((net.forward_seq(b, X)[0][-1].astype(f64) * gout).sum()), blob, pick, 2e-2)
    scale = np.abs(G[pick]).max()
    assert np.abs(G[pick] - num).max() < 40 * np.sqrt(np.finfo(f32).eps) * scale
