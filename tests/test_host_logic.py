"""CPU: settings surface, C-ABI library loads and exports every declared symbol (no compute
calls without a GPU), struct layout agreement between ctypes and the C header."""
import ctypes as C
import math
import os
import re

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def test_hyperparameters_defaults_match_reference_code_defaults():
    from smarties_b200 import HyperParameters
    hp = HyperParameters(32, 8, {})
    assert hp.learner == "VRACER" and hp.batchSize == 256 and hp.nnLayerSizes == [128, 128]
    assert hp.gamma == 0.995 and hp.lambda_ == 1 and hp.penalTol == 0.1 and hp.epsAnneal == 5e-7
    assert hp.clipImpWeight == math.sqrt(8 / 2.0)                      # HyperParameters.h:46
    assert hp.maxTotObsNum == int(2 ** 14 * math.sqrt(40))             # HyperParameters.h:53
    assert hp.minTotObsNum == hp.maxTotObsNum                           # HyperParameters.cpp:191
    assert hp.outWeightsPrefac == 1e-3 and hp.nnFunc == "Tanh"
    assert hp.returnsEstimator == "retrace"                             # AlgoFactory.cpp:134-136


def test_hyperparameters_distributed_split():
    from smarties_b200 import HyperParameters
    hp = HyperParameters(32, 8, {"maxTotObsNum": 8388608})
    hp.define_distributed_learning(8)
    assert hp.batchSize_local == 32 and hp.maxTotObsNum_local == 1048576


def test_settings_file_of_reference_parses():
    from smarties_b200 import HyperParameters
    hp = HyperParameters(4, 1, {"learner": "VRACER", "dataSamplingAlgo": "uniform", "returnsEstimator": "retrace",
                                "ERoldSeqFilter": "oldest", "nnLayerSizes": [128, 128]})   # settings/VRACER.json
    assert hp.nnLayerSizes == [128, 128]
    assert HyperParameters(4, 1, {"learner": "RACER"}).learner == "RACER"
    with pytest.raises(NotImplementedError):
        HyperParameters(4, 1, {"learner": "PPO"})
    with pytest.raises(KeyError):
        HyperParameters(4, 1, {"notAKey": 1})


def test_racer_family_settings_files_of_the_reference_are_covered():
    """Every V-RACER / RACER settings file the reference ships (settings/*.json, restated in tests/reference_settings.py) passes the
    host-side coverage check and yields a device configuration; the CMA variant is refused (it stays with the reference learner)."""
    from reference_settings import DEVICE, REFERENCE_ONLY
    from smarties_b200 import HyperParameters
    from smarties_b200.learner import make_config
    for name, js in DEVICE.items():
        hp = HyperParameters(17, 6, dict(js))
        cfg, _ = make_config(17, 6, dict(js))
        assert cfg.n_hidden == len([h for h in hp.nnLayerSizes if h > 0]), name
        assert cfg.batch_size == hp.batchSize_local and cfg.algo == (1 if hp.learner == "RACER" else 0), name
    for name, js in REFERENCE_ONLY.items():
        with pytest.raises(NotImplementedError):
            HyperParameters(17, 6, dict(js))


def test_settings_outside_the_device_path_are_rejected_loudly():
    """What the device library covers of createReturnEstimator / prepareSampler / getERfilterAlgo / Builder::addLayer
    (the oracle covers more: tests/parity_utils.ORACLE_ONLY_CASES); everything else raises instead of running something else."""
    from smarties_b200 import HyperParameters
    assert HyperParameters(4, 1, {"returnsEstimator": "GAE"}).returnsEstimator == "GAE"
    assert HyperParameters(4, 1, {"nnType": "LSTM", "nnLayerSizes": [32]}).nnType == "LSTM"
    assert HyperParameters(4, 1, {"nnType": "GRU", "nnLayerSizes": [32]}).bRecurrent
    assert HyperParameters(4, 1, {"dataSamplingAlgo": "PERrank", "ERoldSeqFilter": "minerror"}).dataSamplingAlgo == "PERrank"
    assert HyperParameters(4, 1, {"encoderLayerSizes": [32], "nnLayerSizes": [32, 16]}).encoderLayerSizes == [32]
    for bad in ({"returnsEstimator": "nonsense"}, {"dataSamplingAlgo": "PERx"}, {"ERoldSeqFilter": "youngest"},
                {"nnType": "RNN"}, {"nnFunc": "HardSigmoid"}, {"nnFunc": "Relu", "nnType": "LSTM", "nnLayerSizes": [16]}):
        with pytest.raises(NotImplementedError):
            HyperParameters(4, 1, bad)


def test_library_exports_every_declared_symbol(built_library):
    from smarties_b200 import EXPORTS, load_library
    header = open(os.path.join(ROOT, "include", "smarties_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(smb200_[a-z_0-9]+)\s*\(", header)))
    assert declared == sorted(EXPORTS)
    lib = load_library()
    for name in declared:
        assert hasattr(lib, name), name


def test_config_struct_layout_and_defaults(built_library):
    from smarties_b200 import load_library
    from smarties_b200.learner import Config
    lib = load_library()
    cfg = Config()
    assert lib.smb200_default_config(C.byref(cfg), 32, 8) == 0
    assert (cfg.dim_state, cfg.dim_action, cfg.n_hidden, cfg.hidden[0], cfg.hidden[1]) == (32, 8, 2, 128, 128)
    assert cfg.batch_size == 256 and cfg.max_tot_obs == int(2 ** 14 * math.sqrt(40))
    assert cfg.gamma == 0.995 and cfg.clip_imp_weight == 2.0 and cfg.seed == 42 and cfg.world_size == 1
    assert abs(cfg.nn_lambda - 1.1920928955078125e-07) < 1e-20
    assert cfg.returns_estimator == 0 and cfg.nn_type == 0 and cfg.min_tot_obs == 0
    # sizeof(smb200_config) as the C compiler sees it: a drifted ctypes mirror would shift every later field
    import subprocess
    import tempfile
    with tempfile.TemporaryDirectory() as tmp:
        src = os.path.join(tmp, "s.c")
        with open(src, "w") as f:
            f.write('#include <stdio.h>\n#include "smarties_b200.h"\nint main(void) { printf("%zu %zu", sizeof(smb200_config), '
                    'sizeof(smb200_step_stats)); return 0; }\n')
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), src, "-o", os.path.join(tmp, "s")], check=True)
        sizes = subprocess.run([os.path.join(tmp, "s")], capture_output=True, text=True, check=True).stdout.split()
    from smarties_b200.learner import StepStats
    assert [int(x) for x in sizes] == [C.sizeof(Config), C.sizeof(StepStats)]


def test_no_gpu_fails_loudly(built_library):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from smarties_b200 import Learner, SmartiesB200Error
    with pytest.raises(SmartiesB200Error, match="no CPU fallback"):
        Learner(6, 3, {"nnLayerSizes": [32, 32], "batchSize": 16, "maxTotObsNum": 2048})


def test_uint_plus_float_x86_matches_oracle(built_library):
    """The statistics phase counts far-policy steps with the reference's `Uint nOffPol += float`
    (MemoryProcessing.cpp:202-227).  The inline function the device code calls, compiled for the host, against the
    oracle's restatement (pinned to the reference goldens vracer_da1 / racer_discrete, where the sum goes negative):
    wrap-around of negative sums, the 2^63 'integer indefinite', zero from 2^64 on, float rounding of large counts."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import numpy as np
    import vracer_oracle as vo
    from smarties_b200 import load_library
    lib = load_library()
    lib.smb200_uint_plus_float.restype = C.c_uint64
    lib.smb200_uint_plus_float.argtypes = [C.c_uint64, C.c_float]
    M = (1 << 64) - 1
    ns = [0, 1, 2, 7, 255, (1 << 24) - 1, 1 << 24, (1 << 24) + 1, (1 << 31) + 5, (1 << 53) + 1, (1 << 62) + 12345,
          (1 << 63) - 1, 1 << 63, (1 << 63) + (1 << 40), M - 1000, M - 1, M]
    xs = [0.0, 0.4, 0.5, 0.99, 1.0, 1.5, 3.75, 199.0, -0.25, -1.0, -1.5, -7.0, -200.5, 1e10, -1e10, 9.3e18, -9.3e18, 1.9e19,
          -1.9e19, 3e19, float(2 ** 63), float(-2 ** 63), float(2 ** 64)]
    rng = np.random.default_rng(5)
    ns += [int(v) for v in rng.integers(0, 1 << 40, 200)]
    xs += [float(np.float32(v)) for v in rng.standard_normal(50) * 300.0]
    bad = []
    for n in ns:
        for x in xs:
            got, want = lib.smb200_uint_plus_float(n, x), vo.uint_plus_float(n, np.float32(x))
            if got != want:
                bad.append((n, x, got, want))
    assert not bad, bad[:5]
    # the sequence of the goldens: a negative sum wraps, the next addition gives 0
    n = lib.smb200_uint_plus_float(0, -40.0)
    assert n == (1 << 64) - 40 and lib.smb200_uint_plus_float(n, 3.0) == 0


def test_far_policy_chain_fast_path_is_exact(built_library):
    """The statistics CTA walks each virtual OpenMP thread's `Uint += float` chain (MemoryProcessing.cpp:202-227) as fadd + trunc
    on a float while the count stays below 2^24 and the terms are non-negative, and falls back to the term-by-term x86 emulation
    otherwise (far_chain).  Host build of that function against the oracle's integer round trip per term: terms a hair below
    and above integers (Ns * (k / Ns) in f32), strides like the T virtual threads, counts that cross 2^24, negative and NaN terms."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import numpy as np
    import vracer_oracle as vo
    from smarties_b200 import load_library
    lib = load_library()
    lib.smb200_host_far_chain.restype = C.c_uint64
    lib.smb200_host_far_chain.argtypes = [C.c_uint64, C.POINTER(C.c_float), C.c_int32, C.c_int32, C.c_int32]
    rng = np.random.default_rng(11)

    lib.smb200_uint_plus_float.restype = C.c_uint64
    lib.smb200_uint_plus_float.argtypes = [C.c_uint64, C.c_float]

    def want(n0, xs, first, stride):
        n = n0
        for x in xs[first::stride]:
            # a NaN term (never produced by the learner) is outside the oracle's restatement: there the reference is the
            # library's own term-by-term x86 emulation (cvttss2si of NaN = the 2^63 "integer indefinite")
            n = lib.smb200_uint_plus_float(n, float(x)) if np.isnan(x) else vo.uint_plus_float(n, np.float32(x))
        return n

    def got(n0, xs, first, stride):
        a = np.ascontiguousarray(xs, np.float32)
        return lib.smb200_host_far_chain(n0, a.ctypes.data_as(C.POINTER(C.c_float)), first, len(a), stride)

    cases = []
    for _ in range(40):                      # what the learner produces: Ns * fracFar with fracFar = f32(k / Ns)
        Ns = rng.integers(2, 1200, 600).astype(np.float32)
        k = (rng.random(600) * Ns).astype(np.int64).astype(np.float32)
        cases.append((Ns * (k / Ns).astype(np.float32)).astype(np.float32))
    cases.append(np.nextafter(np.arange(1, 400, dtype=np.float32), np.float32(0)))          # a hair below every integer
    cases.append(np.nextafter(np.arange(1, 400, dtype=np.float32), np.float32(1e9)))
    cases.append((rng.random(3000) * 9000).astype(np.float32))                              # crosses 2^24 on the way
    cases.append(np.concatenate([np.full(50, 3.5, np.float32), [-2.0], np.full(50, 1.25, np.float32)]).astype(np.float32))
    cases.append(np.concatenate([np.full(10, 2.0, np.float32), [np.nan], np.full(10, 2.0, np.float32)]).astype(np.float32))
    cases.append(np.zeros(0, np.float32) + 0)
    for xs in cases:
        for n0 in (0, 5, (1 << 24) - 3, (1 << 24) + 7, (1 << 40) + 1):
            for first, stride in ((0, 1), (3, 8), (31, 32), (0, 16)):
                assert got(n0, xs, first, stride) == want(n0, xs, first, stride), (n0, first, stride, xs[:5])


def test_every_settings_file_of_the_reference_is_parsed_or_rejected_with_a_reason():
    """The settings/*.json surface (Settings/HyperParameters.h:37-73, HyperParameters::initializeOpts): every file the reference
    ships (tests/golden/reference_settings.json, generator make_settings_fixture.py) either configures the device path or is
    refused with the setting that is outside it — never an unknown key, never a silent fallback."""
    import json
    from smarties_b200 import HyperParameters
    with open(os.path.join(ROOT, "tests", "golden", "reference_settings.json")) as f:
        files = json.load(f)
    assert len(files) >= 17
    covered = {"VRACER.json": ("VRACER", "FFNN"), "RACER.json": ("RACER", "FFNN"), "RACER_RNN.json": ("RACER", "LSTM"),
               "VRACER_LES.json": ("VRACER", "FFNN"), "RACER_glider.json": ("RACER", "FFNN"), "RACER_atari.json": ("RACER", "FFNN"),
               "VRACER_expensiveData.json": ("VRACER", "GRU"), "default.json": ("VRACER", "FFNN")}      # default.json: SoftSign hidden layers
    refused = {"ACER.json": "learner=ACER", "CMA.json": "learner=CMA", "DPG.json": "learner=DPG", "DPG_light.json": "learner=DPG",
               "DPG_orig.json": "learner=DPG", "DQN.json": "learner=DQN", "NAF.json": "learner=NAF", "PPO.json": "learner=PPO",
               "VRACER_CMA.json": "ESpopSize"}
    for name, settings in files.items():
        if name in covered:
            hp = HyperParameters(8, 2, settings)
            assert (hp.learner, hp.nnType) == covered[name], name
        else:
            with pytest.raises(NotImplementedError, match=refused[name]):
                HyperParameters(8, 2, settings)
    assert set(files) == set(covered) | set(refused)


def test_binding_loads_the_library_and_dies_loudly_without_a_gpu(tmp_path):
    """integration/RACER_B200.cpp compiled into the reference (oracle/_ref/b200/, integration/Makefile): the reference's own
    cart-pole app starts, the wrapped factory picks the device learner, fills smb200_config from the reference's settings and
    calls smb200_create — which, on a box without a GPU, must end the run with the library's message (never a CPU path)."""
    import sys
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: tests/test_gpu_dropin.py runs the real thing")
    if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "b200", "cart_pole")):
        pytest.skip("oracle/_ref binaries not built (python -c 'import __graft_entry__ as g; g.build()')")
    sys.path.insert(0, os.path.join(ROOT, "scripts"))
    from dropin_run import SETTINGS, run_arm
    r = run_arm("b200", steps=500, threads=2, seed=7, settings=dict(SETTINGS), keep_dir=str(tmp_path))
    assert r["rc"] != 0
    assert any("learner steps of this agent run on the GPU" in l for l in r["b200_lines"]), r
    assert "smarties_b200 create" in r["tail"] and "no CPU fallback" in r["tail"], r["tail"][-600:]


def test_binding_falls_back_to_the_reference_learner_outside_the_device_path(tmp_path):
    """SMARTIES_B200=1 with a setting the device path does not cover (here a non-FIFO episode filter, which this binding leaves to the reference learner): the wrapped factory
    hands the agent to the reference's own CPU learner (createLearner_reference) — the app trains, no device learner is created.
    Runs without a GPU: nothing of the library is called on this path."""
    import sys
    if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "b200", "cart_pole")):
        pytest.skip("oracle/_ref binaries not built (python -c 'import __graft_entry__ as g; g.build()')")
    sys.path.insert(0, os.path.join(ROOT, "scripts"))
    from dropin_run import SETTINGS, run_arm
    r = run_arm("b200", steps=1200, threads=2, seed=7, settings=dict(SETTINGS, ERoldSeqFilter="farpolfrac"), keep_dir=str(tmp_path))
    assert r["rc"] == 0, r
    assert r["b200_lines"] == [] and r["grad_steps_logged"] >= 1000, r
