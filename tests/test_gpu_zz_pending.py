"""GPU: device cases written AFTER the round's GPU budget was spent — compiled for sm_100a and pinned on the CPU side, but
never run on a B200 yet.  Each case runs in a child process (a CUDA fault cannot poison the context of other tests; the
file sorts last) and is a non-strict xfail until its first green run on the box, when it moves to test_gpu_parity.py.

vracer_da1 — one action component (cart-pole's shape): clipImpWeight = sqrt(1/2) < 1, so the reference starts with
CinvRet = 1/C > 1 (MemoryBuffer.h:41-44), every stored importance weight 1 counts as "was far", per-episode far-policy
fractions go negative and `Uint nOffPol += float` (MemoryProcessing.cpp:202-227) wraps through x86's cvttss2si.  The
statistics phase now converts with the same semantics (`uint_plus_float_x86`, csrc/common.cuh; host build pinned to the
oracle by tests/test_host_logic.py::test_uint_plus_float_x86_matches_oracle).

vracer_explore — "returnsEstimator": "retraceExplore" (computeRetraceExplBonus, MemoryProcessing.cpp:402-409): the bonus
(1 - gamma) * (|Q' - A - V| - stats.maxAbsError) makes the recursion non-affine; `k_sweep_explore` (csrc/sweep_kernels.cu) stages
an episode in shared memory and runs the recursion sequentially in the reference's operation order.  The golden covers the
initial sweep (baseline 0), insertion and the step-1000 recompute (baseline = the device-resident statistic).  The setting is
accepted only with SMB200_UNVERIFIED=1 until this test has been green."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))

CHILD = r"""
import sys
sys.path[:0] = [{root!r}, {oracle!r}, {tests!r}]
import numpy as np
from parity_utils import Golden, make_learner
from test_gpu_parity import _check_step, _check_final
g = Golden({case!r})
L = make_learner(g)
R = g.ref
st = L.get_stats()
assert st["beta"] == R["init/refer"][0] and st["cmax"] == R["init/refer"][1] and st["cinv"] == R["init/refer"][2]
assert np.allclose(L.read_field("QRET"), R["init/Qret"], rtol=2e-5, atol=2e-5)
for s in range(g.steps):
    _check_step(L, g, R, f"s{{s}}", L.train_steps(1)[0])
_check_final(L, R)
L.close()
print("ok")
"""


CHILD_FULL = r"""
import json, os, sys
sys.path[:0] = [{root!r}, {oracle!r}, {tests!r}]
import numpy as np
import bench
from smarties_b200 import Learner
z = np.load(os.path.join({tests!r}, "golden", {fixture!r}))
spec = json.loads(bytes(z["spec"]).decode())
# the library's own network initialisation at randSeed 42 IS the reference's (tests/test_host_replay.py), the harness ran 1 thread
L = Learner(32, 8, dict(spec["settings"]), seed=spec["seed"], refer_reduce_threads=1)
L.load_replay(bench.make_workload())
L.initialize_learner()
L.seed_sampler(spec["sample_seed"])
mean, scale, std, rew = L.get_scaling()
assert np.allclose(mean, z["init/stateMean"], atol=1e-7) and np.allclose(scale, z["init/stateScale"], rtol=1e-6)
assert np.allclose(rew, z["init/rewards"], rtol=1e-6, atol=1e-8)
q = L.read_field("QRET")
assert q.size == 1001000
assert np.allclose(q[::spec["stride"]], z["init/Qret_sub"], rtol=2e-5, atol=5e-5)
q64 = q.astype(np.float64)
assert abs(q64.sum() - z["init/Qret_sum"][0]) < 1e-5 * np.abs(q64).sum()
assert abs((q64 * q64).sum() - z["init/Qret_sum"][1]) < 1e-4 * z["init/Qret_sum"][1]
st = L.get_stats()
assert st["beta"] == z["init/refer"][0] and st["cmax"] == z["init/refer"][1]
for k in range(spec["steps"]):
    st = L.train_steps(1)[0]
    ref = z[f"s{{k}}/post/refer"]
    assert st["cmax"] == ref[1] and st["cinv"] == ref[2] and st["n_far_policy"] == int(ref[3]), (k, st, ref[:4])
    assert abs(st["beta"] - ref[0]) <= 1e-12 * ref[0]
    O, g, X = L.get_last_batch()
    assert np.abs(O[:, 0] - z[f"s{{k}}/O_V"]).max() < 5e-6, k
L.close()
print("ok")
"""


def _run_child(case):
    code = CHILD.format(root=os.path.dirname(HERE), oracle=os.path.join(os.path.dirname(HERE), "oracle"), tests=HERE, case=case)
    env = dict(os.environ, SMB200_UNVERIFIED="1")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=180, env=env)
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), (r.stdout[-3000:] + "\n" + r.stderr[-3000:])


@pytest.mark.xfail(strict=False, reason="first run on a B200 pending (written after the round's GPU budget was spent)")
def test_one_action_component_far_policy_count_wraps_like_the_reference():
    _run_child("vracer_da1")


@pytest.mark.xfail(strict=False, reason="first run on a B200 pending (written after the round's GPU budget was spent)")
def test_retrace_explore_estimator_matches_reference():
    _run_child("vracer_explore")


@pytest.mark.xfail(strict=False, reason="first run on a B200 pending (written after the round's GPU budget was spent)")
@pytest.mark.parametrize("fixture", ["cfg2_full_props.npz", "cfg3_full_props.npz"])
def test_full_size_run_matches_the_reference_binary(fixture):
    """BASELINE.json configs[1] (the bench workload) and configs[2] (RACER + LSTM(64), BPTT 32, batch 128 on the same buffer) at
    their full size against values of the reference binary at that size (tests/golden/cfg{2,3}_full_props.npz): normalisers and
    Retrace estimates after initializeLearner (subsample + checksums), then three learner steps — ReF-ER scalars, far-policy
    counts, value outputs of the sampled transitions.  Every ingredient is a verified path; only this combination has not run
    on a GPU yet."""
    root = os.path.dirname(HERE)
    code = CHILD_FULL.format(root=root, oracle=os.path.join(root, "oracle"), tests=HERE, fixture=fixture)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), (r.stdout[-3000:] + "\n" + r.stderr[-3000:])


@pytest.mark.xfail(strict=False, reason="first run on a B200 pending (written after the round's GPU budget was spent)")
def test_mini_batch_of_1024_matches_reference():
    """batchSize 1024 = 256 four-sample tiles, more than the worker CTAs of the persistent grid: every CTA loops over several
    tiles, the next-mini-batch staging and the helper CTAs are off, the weight gradient contracts four 256-column chunks —
    the regime of the batch-size sweep, which has not run on a GPU yet.  Golden vracer_b1024 (three steps
    across the every-1000-steps sweep)."""
    _run_child("vracer_b1024")
