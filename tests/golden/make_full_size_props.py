"""Generates tests/golden/cfg2_full_props.npz: size-independent evidence from the REFERENCE BINARY at the full size of
BASELINE.json configs[1] (the workload of bench.py: 1000 episodes x 1000 steps = 1 M transitions, 32 states, 8 actions,
MLP(128,128), batch 256, randSeed 42, sampler seed 7): the sampled (episode, t) of three learner steps, the ReF-ER scalars
around them, the reward / state normalisers and the Retrace estimates after initializeLearner as checksums and a strided
subsample (the full arrays are 4 MB each), plus the summed parameter gradient and the weights after the first step.
`tN` runs the reference with N OpenMP threads (cfg2_full_props_t16.npz: the thread count of bench.py's reference arm; the
far-policy count and beta depend on it, MemoryProcessing.cpp:202-227).
Run in the build container:  python tests/golden/make_full_size_props.py [cfg3] [t16]"""
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from smarties_b200 import synth  # noqa: E402

HARNESS = os.path.join(ROOT, "oracle", "_ref", "ref_harness")
STEPS, SAMPLE_SEED, STRIDE = 3, 7, 997

# configs[1] (cfg2_full_props.npz) and configs[2] (cfg3_full_props.npz: RACER + LSTM(64), nnBPTTseq 32, batch 128, same buffer)
CFG3 = {"learner": "RACER", "nnType": "LSTM", "nnLayerSizes": [64], "nnBPTTseq": 32, "batchSize": 128, "clipImpWeight": 4,
        "explNoise": 0.1, "gamma": 0.99, "epsAnneal": 0, "nnLambda": 1e-6, "maxTotObsNum": 1048576, "minTotObsNum": 1000000}
THREADS = max([int(a[1:]) for a in sys.argv[1:] if a[0] == "t" and a[1:].isdigit()] or [1])
NAME, SETTINGS = ("cfg3_full_props", CFG3) if "cfg3" in sys.argv[1:] else ("cfg2_full_props", bench.SETTINGS)
NAME += (f"_t{THREADS}" if THREADS > 1 else "") + ".npz"

d = bench.make_workload()
with tempfile.TemporaryDirectory() as tmp:
    synth.write_replay_file(os.path.join(tmp, "data.bin"), d)
    with open(os.path.join(tmp, "settings.json"), "w") as f:
        json.dump(SETTINGS, f)
    subprocess.run([HARNESS, "--data", "data.bin", "--settings", "settings.json", "--steps", str(STEPS), "--threads", str(THREADS),
                    "--sampleSeed", str(SAMPLE_SEED), "--dump", "out.bin", "--dumpSteps", "0,1,2", "--quiet"],
                   cwd=tmp, check=True, stdout=subprocess.DEVNULL, env=dict(os.environ, OMP_NUM_THREADS=str(THREADS)))
    D = synth.read_dump(os.path.join(tmp, "out.bin"))
keep = {"spec": np.frombuffer(json.dumps(dict(workload=bench.WORKLOAD, settings=SETTINGS, steps=STEPS, sample_seed=SAMPLE_SEED,
                                              stride=STRIDE, seed=42, threads=THREADS)).encode(), dtype=np.uint8)}
for k in ("init/refer", "init/stateMean", "init/stateScale", "init/stateStdDev", "init/rewards", "init/epLen"):
    keep[k] = D[k]
q = np.asarray(D["init/Qret"], np.float64)
keep["init/Qret_sum"] = np.array([q.sum(), (q * q).sum(), np.abs(q).max()])
keep["init/Qret_sub"] = np.asarray(D["init/Qret"][::STRIDE], np.float32)
for s in range(STEPS):
    for k in ("sampledEpID", "sampledT", "pre/refer", "post/refer"):
        keep[f"s{s}/{k}"] = D[f"s{s}/{k}"]
    keep[f"s{s}/O_V"] = np.asarray(D[f"s{s}/O"][:, 0], np.float64)          # value outputs of the sampled transitions
# the network the run started from: generators[0] = mt19937(randSeed) has already seeded the other OpenMP threads' generators
# (one draw each, Settings/ExecutionInfo.cpp:389-393), so the initial weights depend on the thread count
keep["init/weights"] = np.asarray(D["init/weights"], np.float32)
keep["s0/gradSum"] = np.asarray(D["s0/gradSum"], np.float32)               # summed parameter gradient of the first step
keep["s0/weights"] = np.asarray(D["s0/weights"], np.float32)               # weights after its Adam update
path = os.path.join(HERE, NAME)
np.savez_compressed(path, **keep)
print(path, os.path.getsize(path) // 1024, "KiB", sorted(keep))
