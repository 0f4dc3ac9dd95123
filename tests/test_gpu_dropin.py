"""GPU: the drop-in boundary end to end.  The reference's own cart-pole app (apps/cart_pole_cpp, Engine +
Communicator + Worker + forked environment, settings/VRACER.json) runs with its learner steps executed by
libsmarties_b200.so through the reference-side binding integration/RACER_B200.cpp.  The binaries are built in
the development container (integration/Makefile, outputs under oracle/_ref/, which travel to the GPU box)."""
import os
import sys

import numpy as np

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(ROOT, "scripts"))


def test_cart_pole_runs_on_the_device_learner():
    exe = os.path.join(ROOT, "oracle", "_ref", "b200", "cart_pole")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/b200/cart_pole not built (make -C integration needs /root/reference)")
    from dropin_run import run_arm
    r = run_arm("b200", steps=3000, threads=4, seed=7, timeout=600)
    assert r.get("rc") == 0, r
    assert any("run on the GPU" in l for l in r["b200_lines"]), r
    done = [l for l in r["b200_lines"] if "gradient steps" in l]
    assert done and int(done[0].split()[1]) >= 2999, r    # the run ends when nGradSteps reaches nTrainSteps - 1
    assert r["stat_rows"] >= 2 and r["avgR_last"] > 5.0 and 0.0 < r["beta_last"] <= 1.0, r


def test_synthetic_env_many_actors_on_the_device_learner():
    """configs[3] shape: 16 forked environment processes (17 states, 6 bounded actions) feed one device learner."""
    exe = os.path.join(ROOT, "oracle", "_ref", "b200", "synth_env")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/b200/synth_env not built (make -C integration needs /root/reference)")
    from dropin_run import run_arm
    r = run_arm("b200", steps=20000, threads=4, seed=3, timeout=600, app="synth_env", envs=16)
    assert r.get("rc") == 0, r
    done = [l for l in r["b200_lines"] if "gradient steps" in l]
    assert done and int(done[0].split()[1]) >= 1000, r
    assert r["stat_rows"] >= 1 and 0.0 < r["beta_last"] <= 1.0, r


def test_checkpoints_cross_between_reference_and_device_learner(tmp_path):
    """saveFreq checkpoints written through the binding (the reference's own writers, fed from the device) are
    restarted by the plain reference binary and by the device learner; a checkpoint of the reference restarts
    the device learner."""
    for exe in ("b200/cart_pole", "cart_pole"):
        if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", exe)):
            pytest.skip("oracle/_ref binaries not built")
    from dropin_run import SETTINGS, run_arm
    S = dict(SETTINGS, saveFreq=1000)
    a, b, c, d = (str(tmp_path / n) for n in "abcd")
    r = run_arm("b200", steps=2500, threads=4, seed=7, settings=S, keep_dir=a)
    assert r["rc"] == 0, r
    files = ["agent_00_net_weights.raw", "agent_00_net_1stMom.raw", "agent_00_net_2ndMom.raw", "agent_00_scaling.raw",
             "agent_00_rank_000_learner_status.raw", "agent_00_rank_000_learner_data.raw"]
    for fn in files:
        assert os.path.getsize(os.path.join(a, fn)) > 0, fn
    m1 = np.fromfile(os.path.join(a, "agent_00_net_1stMom.raw"), np.float32)
    assert np.abs(m1).max() > 0, "Adam moments of the device learner did not reach the checkpoint"
    status = open(os.path.join(a, "agent_00_rank_000_learner_status.raw")).read()
    assert "nGradSteps: 2000" in status, status
    # the reference's output-gradient statistics file (StatsTracker) is written by the device learner as well:
    # header nOutputs + 0.1 (V, mean, stdev of one action), then mean | rms rows of steps 0, 1000, 2000
    gs = np.fromfile(os.path.join(a, "agent_00_net_outGrad_stats.raw"), np.float32)
    assert gs.size >= 1 + 2 * 6 and (gs.size - 1) % 6 == 0 and gs[0] == np.float32(3.1) and np.isfinite(gs).all(), gs
    # the reference restarts from the device learner's checkpoint ...
    r2 = run_arm("ref", steps=3600, threads=4, seed=8, settings=S, keep_dir=b, restart=a)
    assert r2["rc"] == 0 and any("agent_00_net_weights" in l for l in r2["restart_lines"]), r2
    assert r2["grad_steps_logged"] >= 3000, r2                 # continued from step 2000, not from 0
    # ... and so does the device learner, from its own and from the reference's checkpoint
    for src, dst in ((a, c), (b, d)):
        r3 = run_arm("b200", steps=3600 if src == a else 4700, threads=4, seed=9, settings=S, keep_dir=dst, restart=src)
        assert r3["rc"] == 0, r3
        assert any("restarted the device learner at gradient step" in l for l in r3["b200_lines"]), r3
        assert r3["grad_steps_logged"] >= (3000 if src == a else 4000) and 0.0 < r3["beta_last"] <= 1.0, r3


def test_python_app_through_pybind11_module_on_the_device_learner():
    """A Python environment (`import smarties`, the reference's pybind11 module linked against the library with the
    binding) trains on the device learner: same script, same settings file as with the reference learner."""
    if not os.path.isdir(os.path.join(ROOT, "oracle", "_ref", "b200", "py")):
        pytest.skip("oracle/_ref/b200/py not built (make -C integration needs /root/reference)")
    from dropin_run import run_arm
    r = run_arm("b200", steps=3000, threads=4, seed=5, timeout=600, app="py_env")
    assert any("run on the GPU" in l for l in r["b200_lines"]), r
    done = [l for l in r["b200_lines"] if "gradient steps" in l]
    assert done and int(done[0].split()[1]) >= 2999, r
    assert r["stat_rows"] >= 2 and 0.0 < r["beta_last"] <= 1.0, r


def test_c_app_through_the_fortran_interface_on_the_device_learner():
    """An environment written in C against include/smarties_extern.h (the ABI Fortran apps bind) on the device learner."""
    if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "b200", "c_env")):
        pytest.skip("oracle/_ref/b200/c_env not built (make -C integration needs /root/reference)")
    from dropin_run import run_arm
    r = run_arm("b200", steps=3000, threads=4, seed=5, timeout=600, app="c_env")
    assert r.get("rc") == 0, r
    done = [l for l in r["b200_lines"] if "gradient steps" in l]
    assert done and int(done[0].split()[1]) >= 2999, r
    assert r["stat_rows"] >= 2 and 0.0 < r["beta_last"] <= 1.0, r


def test_cart_pole_recurrent_racer_on_the_device_learner():
    """settings/RACER_RNN.json of the reference (RACER, LSTM layers, BPTT window): the actors run the reference's recurrent
    `Approximator` on the host, the learner steps — including the tcgen05 weight-gradient contraction — run on the GPU."""
    exe = os.path.join(ROOT, "oracle", "_ref", "b200", "cart_pole")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/b200/cart_pole not built (make -C integration needs /root/reference)")
    from dropin_run import run_arm
    S = {"learner": "RACER", "nnType": "LSTM", "nnLayerSizes": [32, 32], "clipImpWeight": 4, "explNoise": 0.1, "gamma": 0.99,
         "epsAnneal": 0, "nnLambda": 1e-6, "maxTotObsNum": 16384, "minTotObsNum": 4096}
    r = run_arm("b200", steps=2000, threads=4, seed=7, settings=S, timeout=600)
    assert r.get("rc") == 0, r
    done = [l for l in r["b200_lines"] if "gradient steps" in l]
    assert done and int(done[0].split()[1]) >= 1999, r
    assert r["stat_rows"] >= 1 and 0.0 < r["beta_last"] <= 1.0 and np.isfinite(r["avgR_last"]), r


def _policy_evaluations(r):
    l = [x for x in r["b200_lines"] if "policy evaluations in" in x]
    assert l, r
    f = l[0].split()
    return int(f[1]), int(f[5]), int(f[-2])        # agents answered, device calls, largest call


def test_cart_pole_with_the_actors_on_the_device():
    """SMARTIES_B200_ACTORS=1: RACER::selectAction / processTerminal (RACER.cpp:30-59) read the network outputs from
    smb200_forward_seq; everything else of the actor (policy sampling, advantage, episode storage) stays the reference's.
    The run has to learn like the host-actor run does."""
    exe = os.path.join(ROOT, "oracle", "_ref", "b200", "cart_pole")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/b200/cart_pole not built (make -C integration needs /root/reference)")
    from dropin_run import run_arm
    r = run_arm("b200", steps=3000, threads=4, seed=7, timeout=600, extra_env={"SMARTIES_B200_ACTORS": "1"})
    assert r.get("rc") == 0, r
    assert any("policy evaluations run on the GPU" in l for l in r["b200_lines"]), r
    agents, calls, _ = _policy_evaluations(r)
    assert agents >= 3000 and calls >= 1, r
    done = [l for l in r["b200_lines"] if "gradient steps" in l]
    assert done and int(done[0].split()[1]) >= 2999, r
    assert r["stat_rows"] >= 2 and r["avgR_last"] > 5.0 and 0.0 < r["beta_last"] <= 1.0, r


def test_many_actors_on_the_device_are_answered_in_batches():
    """16 environment processes, 4 worker threads: requests that arrive while a device call is in flight share the next
    call (the binding's combiner), so the number of device calls is below the number of policy evaluations."""
    exe = os.path.join(ROOT, "oracle", "_ref", "b200", "synth_env")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/b200/synth_env not built (make -C integration needs /root/reference)")
    from dropin_run import run_arm
    r = run_arm("b200", steps=20000, threads=4, seed=3, timeout=600, app="synth_env", envs=16, extra_env={"SMARTIES_B200_ACTORS": "1"})
    assert r.get("rc") == 0, r
    agents, calls, largest = _policy_evaluations(r)
    assert agents >= 20000 and calls <= agents and largest >= 1, r
    done = [l for l in r["b200_lines"] if "gradient steps" in l]
    assert done and int(done[0].split()[1]) >= 1000, r
    assert r["stat_rows"] >= 1 and 0.0 < r["beta_last"] <= 1.0, r


def test_recurrent_actors_on_the_device():
    """RACER with LSTM layers: every action of the reference's host actor is a forward pass over the window of up to
    nnBPTTseq + 1 states; here the window goes to the device (k_forward_seq, one CTA per agent)."""
    exe = os.path.join(ROOT, "oracle", "_ref", "b200", "cart_pole")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/b200/cart_pole not built (make -C integration needs /root/reference)")
    from dropin_run import run_arm
    S = {"learner": "RACER", "nnType": "LSTM", "nnLayerSizes": [32, 32], "clipImpWeight": 4, "explNoise": 0.1, "gamma": 0.99,
         "epsAnneal": 0, "nnLambda": 1e-6, "maxTotObsNum": 16384, "minTotObsNum": 4096}
    r = run_arm("b200", steps=2000, threads=4, seed=7, settings=S, timeout=600, extra_env={"SMARTIES_B200_ACTORS": "1"})
    assert r.get("rc") == 0, r
    agents, calls, _ = _policy_evaluations(r)
    assert agents >= 2000 and calls >= 1, r
    done = [l for l in r["b200_lines"] if "gradient steps" in l]
    assert done and int(done[0].split()[1]) >= 1999, r
    assert r["stat_rows"] >= 1 and 0.0 < r["beta_last"] <= 1.0 and np.isfinite(r["avgR_last"]), r


def test_cart_pole_with_encoder_layers_on_the_device_learner():
    """"encoderLayerSizes" (Learner_approximator::createEncoder, Learner_approximator.cpp:148-166): for RACER the encoder layers are
    the first layers of the one network (RACER::setupNet, RACER_common.cpp:82-91); the binding hands the stacked list to the device."""
    exe = os.path.join(ROOT, "oracle", "_ref", "b200", "cart_pole")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/b200/cart_pole not built (make -C integration needs /root/reference)")
    from dropin_run import SETTINGS, run_arm
    S = dict(SETTINGS, encoderLayerSizes=[64], nnLayerSizes=[64, 32])
    r = run_arm("b200", steps=3000, threads=4, seed=7, settings=S, timeout=600)
    assert r.get("rc") == 0, r
    assert any("run on the GPU" in l for l in r["b200_lines"]), r      # not the fall-back to the reference learner
    done = [l for l in r["b200_lines"] if "gradient steps" in l]
    assert done and int(done[0].split()[1]) >= 2999, r
    assert r["stat_rows"] >= 2 and r["avgR_last"] > 5.0 and 0.0 < r["beta_last"] <= 1.0, r


def test_target_delay_reaches_the_checkpoint_through_the_binding(tmp_path):
    """"targetDelay": 0.01 — the device maintains AdamOptimizer::target_weights (Optimizer.cpp:162-177), the binding copies them
    into the host optimiser before the reference's own writers save <name>_net_tgt_weights.raw; the reference restarts from it."""
    for exe in ("b200/cart_pole", "cart_pole"):
        if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", exe)):
            pytest.skip("oracle/_ref binaries not built")
    from dropin_run import SETTINGS, run_arm
    S = dict(SETTINGS, saveFreq=1000, targetDelay=0.01)
    a, b = str(tmp_path / "a"), str(tmp_path / "b")
    r = run_arm("b200", steps=2500, threads=4, seed=7, settings=S, keep_dir=a)
    assert r["rc"] == 0 and any("run on the GPU" in l for l in r["b200_lines"]), r
    w = np.fromfile(os.path.join(a, "agent_00_net_weights.raw"), np.float32)
    t = np.fromfile(os.path.join(a, "agent_00_net_tgt_weights.raw"), np.float32)
    assert w.shape == t.shape and np.isfinite(t).all()
    d = np.abs(w - t).max()
    assert 0 < d < 0.5, d                      # an exponential average that trails the weights, not a copy and not the initial weights
    r2 = run_arm("ref", steps=3600, threads=4, seed=8, settings=S, keep_dir=b, restart=a)
    assert r2["rc"] == 0 and r2["grad_steps_logged"] >= 3000, r2
