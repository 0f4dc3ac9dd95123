"""GPU: the reference's output-gradient statistics file (StatsTracker, Utils/StatsTracker.cpp:28-107)
written by the device learner through smb200_set_grad_stats, against the file the reference wrote during the golden runs
(tests/golden/outgrad_stats.npz).  The library reads the per-sample output gradients of every step that starts at
nGradSteps % 1000 == 0 back from the step kernel's diagnostics buffer (such a step ends its launch) and reduces them on
the host in long double like the reference."""
import os

import numpy as np
import pytest

from parity_utils import CASES, RECURRENT_CASES, Golden, make_learner

pytestmark = pytest.mark.gpu

TOL = 5e-5      # relative to the largest entry: the bar of the output gradient itself (test_gpu_parity.TOL_G)


def _file(base):
    fn = base + "_outGrad_stats.raw"
    return np.fromfile(fn, dtype=np.float32) if os.path.exists(fn) else np.zeros(0, np.float32)


@pytest.mark.parametrize("case", CASES + RECURRENT_CASES)
def test_grad_stats_file_matches_reference(case, tmp_path):
    g = Golden(case)
    want = np.load(g.path("outgrad_stats.npz"))[case]
    base = str(tmp_path / "agent_00_net")
    L, Plain = make_learner(g), make_learner(g)
    L.set_grad_stats(base)
    sa = L.train_steps(g.steps)                 # one call: the tracked step splits the pipeline's segment
    got = _file(base)
    header = 1 if g.start_step % 1000 == 0 else 0
    assert got.size == want.size
    if header:
        assert got[0] == want[0] == np.float32(L.n_out + .1)
    assert np.abs(got - want).max() < TOL * np.abs(want[header:]).max()
    # observability only: the run itself is bit-identical with and without the file
    sb = Plain.train_steps(g.steps)
    assert sa == sb and np.array_equal(L.get_weights(), Plain.get_weights())
    L.close(); Plain.close()


@pytest.mark.parametrize("case", ["vracer_small", "racer_lstm"])
def test_grad_stats_step_by_step_and_injected(case, tmp_path):
    """Same file from step-by-step calls and from smb200_train_step_on with the sampler's own draws."""
    g = Golden(case)
    a, b, c = (str(tmp_path / n) for n in "abc")
    A, Bm, Cm = make_learner(g), make_learner(g), make_learner(g)
    A.set_grad_stats(a); Bm.set_grad_stats(b); Cm.set_grad_stats(c)
    A.train_steps(g.steps)
    for _ in range(g.steps):
        Bm.train_steps(1)
        pos, t = Cm.sample_minibatch()
        Cm.train_step_on(pos, t)
    fa, fb, fc = _file(a), _file(b), _file(c)
    assert fa.size > 0 and np.array_equal(fa, fb) and np.array_equal(fa, fc)
    A.close(); Bm.close(); Cm.close()
    # switched off again before training: no file
    D = make_learner(g)
    d = str(tmp_path / "d")
    D.set_grad_stats(d); D.set_grad_stats(None)
    D.train_steps(g.steps)
    assert not os.path.exists(d + "_outGrad_stats.raw")
    D.close()


def test_grad_stats_rows_every_1000_steps(tmp_path):
    """1001 steps in one call from step 0: header + the rows of steps 0 and 1000; the second row is the mean / rms of
    the last step's output gradients (step 1000 is the call's last step, smb200_get_last_batch)."""
    g = Golden("vracer_prune")
    want = np.load(g.path("outgrad_stats.npz"))["vracer_prune"]
    base = str(tmp_path / "agent_00_net")
    L = make_learner(g)
    L.set_grad_stats(base)
    st = L.train_steps(1001)
    assert st[-1]["grad_step"] == 1001
    got = _file(base)
    n = L.n_out
    assert got.size == 1 + 4 * n and got[0] == np.float32(n + .1)
    assert np.abs(got[:1 + 2 * n] - want).max() < TOL * np.abs(want[1:]).max()
    _, gg, _ = L.get_last_batch()
    gg = gg.astype(np.float64)
    row = np.concatenate([gg.mean(0), np.sqrt((gg * gg).mean(0))]).astype(np.float32)
    assert np.allclose(got[1 + 2 * n:], row, rtol=1e-6, atol=1e-7 * np.abs(row).max())
    L.close()
