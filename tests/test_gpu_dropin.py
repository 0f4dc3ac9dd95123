"""GPU: the drop-in boundary end to end.  The reference's own cart-pole app (apps/cart_pole_cpp, Engine +
Communicator + Worker + forked environment, settings/VRACER.json) runs with its learner steps executed by
libsmarties_b200.so through the reference-side binding integration/RACER_B200.cpp.  The binaries are built in
the development container (integration/Makefile, outputs under oracle/_ref/, which travel to the GPU box)."""
import os
import sys

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
