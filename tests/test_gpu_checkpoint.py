"""GPU: checkpoint compatibility with the reference.

Golden cases `vracer_ckpt`, `racer_lstm_ckpt`: the reference ran some learner steps, called
Learner_approximator::save(), and a fresh reference process restarted from those files and trained on
(tests/golden/make_golden.py).  Here the device learner (1) restarts from the reference's files and holds
bit-identical state, (2) continues training like the restarted reference, (3) writes files that are
byte-identical to the ones it read, (4) produces checkpoints of its own runs that round-trip."""
import os

import numpy as np
import pytest

from parity_utils import CKPT_CASES, TARGET_CASES, Golden, PhaseB, make_learner, write_checkpoint
from test_gpu_parity import TOL_W, _check_final, _check_step

pytestmark = pytest.mark.gpu
FILES = ["agent_00_net_weights.raw", "agent_00_net_tgt_weights.raw", "agent_00_net_1stMom.raw", "agent_00_net_2ndMom.raw",
         "agent_00_scaling.raw", "agent_00_rank_000_learner_status.raw", "agent_00_rank_000_learner_data.raw"]


def fresh_learner(g):
    from smarties_b200 import Learner
    return Learner(g.dS, g.dA, dict(g.settings), bounded=g.bounded, refer_reduce_threads=1)


@pytest.mark.parametrize("case", CKPT_CASES + TARGET_CASES)
def test_restart_from_reference_files_then_train(case, tmp_path):
    g = Golden(case)
    R = g.ref2
    L = fresh_learner(g)
    L.restart(write_checkpoint(g, str(tmp_path)))
    if case in TARGET_CASES:       # the target weights of the checkpoint are on the device, bit for bit
        import vracer_oracle as vo
        from parity_utils import make_oracle
        lay = make_oracle(g).layout
        assert np.array_equal(lay.strip_padding(L.get_target_weights()), np.frombuffer(bytes(g.ckpt["agent_00_net_tgt_weights.raw"]), np.float32))
    # (1) state after restart == state of the restarted reference, bit for bit
    assert np.array_equal(L.get_weights(), R["init/weights"])
    mean, scale, std, rew = L.get_scaling()
    assert np.array_equal(mean, R["init/stateMean"]) and np.array_equal(scale, R["init/stateScale"])
    assert np.array_equal(std, R["init/stateStdDev"]) and np.array_equal(rew, R["init/rewards"])
    for name, key in (("V", "V"), ("ADV", "A"), ("QRET", "Qret"), ("DELTA", "delta"), ("RHO", "rho"), ("KL", "KL")):
        assert np.array_equal(L.read_field(name), R["init/" + key]), name
    ids, rows, agg = L.read_episodes()
    assert list(ids) == list(R["init/epID"]) and list(rows) == list(R["init/epLen"])
    assert np.allclose(agg[:, :9], R["init/epAgg"][:, :9], rtol=1e-5, atol=1e-6)       # Episode::updateCumulative on restart
    st = L.get_stats()
    assert st["beta"] == R["init/refer"][0] and st["cmax"] == R["init/refer"][1]
    assert st["grad_step"] == g.start_step + g.steps + 1      # "nGradSteps" of the status file (the reference's off-by-one)
    # (2) training continues like the restarted reference
    L.seed_sampler(g.spec["sample_seed_after"])
    pb = PhaseB(g)
    for s in range(pb.steps):
        stats = L.train_steps(1)[0]
        _check_step(L, pb, R, f"s{s}", stats)
        if case in TARGET_CASES:   # cntUpdateDelay restarts at 0 in the new process: copy / average from its first update on
            assert np.abs(L.get_target_weights() - R[f"s{s}/tgt"]).max() < TOL_W, s
    _check_final(L, R)
    L.close()


@pytest.mark.parametrize("case", TARGET_CASES)
@pytest.mark.parametrize("per_call", [1, 8])
def test_target_weights_follow_the_reference(case, per_call, tmp_path):
    """"targetDelay" 0.05 / 3 (AdamOptimizer::apply_update, Optimizer.cpp:162-177) in the Adam epilogue of the weight-gradient
    tiles: the target weights after every update against the reference's (one step per launch and all eight in one
    persistent launch), and the checkpoint file against the one the reference wrote."""
    g = Golden(case)
    R = g.ref
    L = make_learner(g)
    assert L.step_kernel() in (0, 1)                    # the tile kernels maintain them
    if per_call == 1:
        for s in range(g.steps):
            L.train_steps(1)
            assert np.abs(L.get_target_weights() - R[f"s{s}/tgt"]).max() < TOL_W, s
    else:
        L.train_steps(g.steps)
        assert np.abs(L.get_target_weights() - R[f"s{g.steps - 1}/tgt"]).max() < TOL_W
    assert np.abs(L.get_weights() - R[f"s{g.steps - 1}/weights"]).max() < TOL_W
    assert np.abs(L.get_target_weights() - L.get_weights()).max() > 1e-6
    L.save(str(tmp_path / "agent_00"))
    own = np.fromfile(tmp_path / "agent_00_net_tgt_weights.raw", np.float32)
    ref = np.frombuffer(bytes(g.ckpt["agent_00_net_tgt_weights.raw"]), np.float32)
    assert own.shape == ref.shape and np.abs(own - ref).max() < TOL_W
    L.close()


@pytest.mark.parametrize("case", CKPT_CASES + TARGET_CASES)
def test_restart_then_save_is_byte_identical(case, tmp_path):
    g = Golden(case)
    src, dst = tmp_path / "in", tmp_path / "out"
    src.mkdir(); dst.mkdir()
    L = fresh_learner(g)
    L.restart(write_checkpoint(g, str(src)))
    L.set_grad_step(L.get_stats()["grad_step"] - 1)     # save() writes nGradSteps + 1 like Learner::save inside logStats
    L.save(str(dst / "agent_00"))
    for fn in FILES:
        a, b = (dst / fn).read_bytes(), bytes(g.ckpt[fn])
        assert a == b, f"{fn}: {len(a)} vs {len(b)} bytes"
        assert (dst / fn.replace(".raw", "_backup.raw")).read_bytes() == a
    L.close()


def test_own_checkpoint_round_trip(tmp_path):
    """train -> save -> restart in a new learner: identical replay, network, moments, scaling, counters; the
    files written after the first save and after the round trip are byte-identical."""
    g = Golden("vracer_ckpt")
    L = make_learner(g)
    L.train_steps(g.steps)
    a, b = tmp_path / "a", tmp_path / "b"
    a.mkdir(); b.mkdir()
    L.save(str(a / "agent_00"))
    # the device run tracks the reference's run: network file within the f32 tolerance of the reference's file
    w_ref = np.frombuffer(bytes(g.ckpt["agent_00_net_weights.raw"]), np.float32)
    w_own = np.fromfile(a / "agent_00_net_weights.raw", np.float32)
    assert w_own.shape == w_ref.shape and np.abs(w_own - w_ref).max() < 5e-6
    assert (a / "agent_00_rank_000_learner_status.raw").read_bytes() == bytes(g.ckpt["agent_00_rank_000_learner_status.raw"])
    assert os.path.getsize(a / "agent_00_rank_000_learner_data.raw") == g.ckpt["agent_00_rank_000_learner_data.raw"].size
    M = fresh_learner(g)
    M.restart(str(a / "agent_00"))
    assert np.array_equal(M.get_weights(), L.get_weights())
    for x, y in zip(M.get_adam(), L.get_adam()):
        assert np.array_equal(x, y)
    for name in ("V", "ADV", "QRET", "DELTA", "RHO", "KL", "REWARD"):
        assert np.array_equal(M.read_field(name), L.read_field(name)), name
    for x, y in zip(M.get_scaling(), L.get_scaling()):
        assert np.array_equal(x, y)
    M.set_grad_step(M.get_stats()["grad_step"] - 1)
    M.save(str(b / "agent_00"))
    for fn in FILES:
        assert (a / fn).read_bytes() == (b / fn).read_bytes(), fn
    L.close(); M.close()


def test_restart_errors_are_loud(tmp_path):
    g = Golden("vracer_ckpt")
    L = fresh_learner(g)
    from smarties_b200 import SmartiesB200Error
    with pytest.raises(SmartiesB200Error):
        L.restart(str(tmp_path / "agent_00"))                 # no weights file
    base = write_checkpoint(g, str(tmp_path))
    with open(base + "_net_weights.raw", "ab") as f:
        f.write(b"\0\0\0\0")                                   # "Mismatch in restarted file"
    with pytest.raises(SmartiesB200Error):
        L.restart(base)
    L.close()
