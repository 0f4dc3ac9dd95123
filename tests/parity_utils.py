"""Shared helpers of the parity tests: golden fixtures (reference outputs), the numpy oracle and
the GPU learner are all constructed from the same spec."""
from __future__ import annotations

import json
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CKPT_CASES = ["vracer_ckpt", "racer_lstm_ckpt"]
TARGET_CASES = ["vracer_tgt_ema", "vracer_tgt_copy"]     # "targetDelay" 0.05 (exponential average) and 3 (copy every 3 updates): checkpoint goldens
CASES = ["vracer_small", "vracer_cfg2mini", "vracer_bounded", "vracer_prune", "racer_small", "racer_bounded", "vracer_gae",
         "vracer_da1", "vracer_explore", "vracer_b1024", "vracer_b4096", "racer_discrete",
         "vracer_softsign", "vracer_hardsign", "racer_sigm", "vracer_relu", "vracer_lrelu",
         "vracer_expplus", "racer_softplus", "vracer_exp", "vracer_linear",
         "vracer_encoder", "racer_encoder2", "vracer_widen"]      # "encoderLayerSizes": stacked under nnLayerSizes in the one network (RACER::setupNet)      # "nnFunc" other than Tanh (settings/default.json: SoftSign)
# golden runs of a reference with 8 / 16 OpenMP threads: the far-policy count (and beta) depend on the thread count
# (MemoryProcessing.cpp:202-227); the device reproduces it with refer_reduce_threads = T
THREADED_CASES = ["vracer_small_t8", "vracer_small_t16", "vracer_cfg2mini_t8", "vracer_cfg2mini_t16", "racer_small_t8"]
# prioritized samplers and non-FIFO episode filters: one launch per step with a host round trip
SLOW_CASES = ["vracer_pererr", "vracer_perseq", "vracer_perrank", "vracer_farpolfrac", "vracer_maxkldiv", "vracer_minerror"]
ORACLE_ONLY_CASES = []
RECURRENT_CASES = ["racer_lstm", "vracer_lstm2", "racer_lstm64", "racer_cfg3mini", "racer_lstm_encoder",      # nnType LSTM + BPTT window (configs[2] family)
                   "racer_mgu", "vracer_gru2"]     # MGU cells (Layer_GRU.h): "MGU" / "GRU", the default of partially observable MDPs


class Golden:
    def __init__(self, name):
        z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
        self.spec = json.loads(bytes(z["spec"]).decode())
        self.ref = {k[4:]: z[k] for k in z.files if k.startswith("ref:")}
        self.replay = {k[7:]: z[k] for k in z.files if k.startswith("replay:")}
        # checkpoint cases (tests/golden/make_golden.py CKPT_CASES): the files Learner_approximator::save() wrote after
        # `steps` steps, and the dumps of a fresh reference process that restarted from them ("phase B")
        self.ckpt = {k[5:]: z[k] for k in z.files if k.startswith("ckpt:")}
        self.ref2 = {k[5:]: z[k] for k in z.files if k.startswith("ref2:")}
        r = self.spec["replay"]
        self.replay["dS"], self.replay["dA"] = r["dS"], r["dA"]
        self.dS, self.dA = r["dS"], r["dA"]
        self.settings = self.spec["settings"]
        self.steps, self.start_step = self.spec["steps"], self.spec["start_step"]
        self.sample_seed, self.bounded = self.spec["sample_seed"], bool(self.spec["bounded"])
        self.B = int(self.ref["meta/dims"][3])
        self.threads = int(self.spec.get("threads", 1))      # OpenMP threads of the reference run that produced the golden

    @staticmethod
    def path(name):
        return os.path.join(GOLDEN_DIR, name)

    def refer(self, key):
        """{beta, cmax, cinv, nFar, avgKL, avgSqErr, maxAbsErr, avgReturn, stdevQ, avgQ, maxQ, minQ, cntRet, sumRetErr}"""
        return self.ref[key + "/refer"]


class PhaseB:
    """View of a checkpoint golden that looks like a Golden of the restarted run."""
    def __init__(self, g: Golden):
        self.ref = g.ref2
        self.steps = g.spec["steps_after"]

    def refer(self, key):
        return self.ref[key + "/refer"]


def write_checkpoint(g: Golden, directory):
    for fn, b in g.ckpt.items():
        with open(os.path.join(directory, fn), "wb") as f:
            f.write(bytes(b))
    return os.path.join(directory, "agent_00")


def make_oracle(g: Golden):
    import vracer_oracle as vo
    s = g.settings
    # every hyper-parameter of settings.json the path reads (Settings/HyperParameters.h:37-73); absent keys = reference defaults
    kw = dict(batch=s.get("batchSize", 256), max_tot_obs=s.get("maxTotObsNum"), bounded=g.bounded, sample_seed=g.sample_seed,
              learner=s.get("learner", "VRACER"), refer_threads=g.threads)
    for key, arg in (("gamma", "gamma"), ("lambda", "lam"), ("clipImpWeight", "clip_imp_weight"), ("penalTol", "penal_tol"),
                     ("epsAnneal", "eps_anneal"), ("learnrate", "learnrate"), ("nnLambda", "nn_lambda"),
                     ("returnsEstimator", "returns_estimator"), ("dataSamplingAlgo", "sampling"), ("ERoldSeqFilter", "er_filter"),
                     ("nnFunc", "nn_func"), ("targetDelay", "target_delay")):
        if key in s:
            kw[arg] = s[key]
    if "n_options" in g.spec["replay"]:
        kw["discrete"] = g.spec["replay"]["n_options"]
    if s.get("nnType", "FFNN") in ("LSTM", "MGU", "GRU"):
        o = vo.RecurrentOracle(g.dS, g.dA, cells=[h for h in s.get("encoderLayerSizes", []) if h > 0] + s["nnLayerSizes"], bptt=s.get("nnBPTTseq", 16), cell_type=s["nnType"], **kw)
    else:
        # encoder layers are the first layers of the one network (RACER::setupNet, RACER_common.cpp:82-91)
        o = vo.VracerOracle(g.dS, g.dA, hidden=[h for h in s.get("encoderLayerSizes", []) if h > 0] + s.get("nnLayerSizes", [128, 128]), **kw)
    o.W[:] = g.ref["init/weights"]
    o.load_replay(g.replay)
    o.initialize_learner()
    o.n_grad_steps = g.start_step
    o.adam_step = g.start_step
    return o


def make_learner(g: Golden, refer_reduce_threads=None):
    from smarties_b200 import Learner
    L = Learner(g.dS, g.dA, dict(g.settings), bounded=g.bounded,
                refer_reduce_threads=g.threads if refer_reduce_threads is None else refer_reduce_threads,
                discrete_options=g.spec["replay"].get("n_options", 0))
    L.set_weights(g.ref["init/weights"])
    L.load_replay(g.replay)
    L.initialize_learner()
    L.set_grad_step(g.start_step)
    L.seed_sampler(g.sample_seed)
    return L


def relerr(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))
