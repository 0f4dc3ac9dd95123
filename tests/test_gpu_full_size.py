"""GPU: runs at the FULL size of BASELINE.json configs[1] (the bench workload: 1000 episodes x 1000 steps, 32 states, 8 actions,
MLP(128,128), batch 256) and configs[2] (RACER + LSTM(64), nnBPTTseq 32, batch 128 on the same buffer) against values the
reference binary produced at that size (tests/golden/cfg{2,3}_full_props*.npz, generator make_full_size_props.py):
normalisers and Retrace estimates after initializeLearner (strided subsample + checksums), then three learner steps — ReF-ER
scalars, integer far-policy counts, value outputs of the sampled transitions, and after the first step the summed parameter
gradient and the weights.  `_t16`: the reference ran 16 OpenMP threads (the thread count of bench.py's reference arm); the
far-policy count, and with it beta, depend on the thread count (MemoryProcessing.cpp:202-227) and the device reproduces it
with refer_reduce_threads = 16."""
import json
import os

import numpy as np
import pytest

from parity_utils import GOLDEN_DIR, relerr

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("fixture", ["cfg2_full_props.npz", "cfg2_full_props_t16.npz", "cfg3_full_props.npz"])
def test_full_size_run_matches_the_reference_binary(fixture):
    import bench
    from smarties_b200 import Learner
    z = np.load(os.path.join(GOLDEN_DIR, fixture))
    spec = json.loads(bytes(z["spec"]).decode())
    # the library's own network initialisation at randSeed 42 IS the reference's (tests/test_host_replay.py)
    L = Learner(32, 8, dict(spec["settings"]), seed=spec["seed"], refer_reduce_threads=spec.get("threads", 1))
    if spec.get("threads", 1) == 1:     # the library's own initialisation at randSeed 42 IS the single-thread reference's
        assert np.array_equal(L.get_weights(), z["init/weights"])
    # a T-thread reference draws T - 1 seeds from generators[0] before it builds the network (ExecutionInfo.cpp:389-393)
    L.set_weights(z["init/weights"])
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
        ref = z[f"s{k}/post/refer"]
        assert st["cmax"] == ref[1] and st["cinv"] == ref[2] and st["n_far_policy"] == int(ref[3]), (k, st, ref[:4])
        assert abs(st["beta"] - ref[0]) <= 1e-12 * ref[0]
        O, g, X = L.get_last_batch()
        assert np.abs(O[:, 0] - z[f"s{k}/O_V"]).max() < 5e-6, k
        if k == 0:
            assert relerr(L.get_grad(), z["s0/gradSum"]) < 1e-4          # same bars as test_gpu_parity.py
            assert np.abs(L.get_weights() - z["s0/weights"]).max() < 5e-6
    L.close()
