"""Generate the golden vectors under tests/golden/ from the reference itself.

Runs oracle/_ref/ref_harness (the UNMODIFIED reference compiled from /root/reference by
oracle/Makefile, single OpenMP thread so the result is deterministic) on small seeded
synthetic replay buffers and stores inputs + reference outputs as compressed .npz files.
Only runs where /root/reference exists (this container); the .npz fixtures are committed.

    python tests/golden/make_golden.py
"""
from __future__ import annotations

import json
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, ROOT)
from smarties_b200 import synth  # noqa: E402

HARNESS = os.path.join(ROOT, "oracle", "_ref", "ref_harness")

CASES = {
    # name: (replay kwargs, settings.json, harness args, keys filter)
    "vracer_small": dict(
        replay=dict(seed=123, n_ep=24, ep_len=(20, 60), dS=6, dA=3),
        settings={"learner": "VRACER", "returnsEstimator": "retrace", "nnLayerSizes": [32, 32], "batchSize": 16,
                  "maxTotObsNum": 2048, "minTotObsNum": 500},
        steps=12, start_step=994, sample_seed=7, bounded=0, full_steps=list(range(12))),
    "vracer_cfg2mini": dict(
        replay=dict(seed=321, n_ep=40, ep_len=(50, 80), dS=32, dA=8),
        settings={"learner": "VRACER", "dataSamplingAlgo": "uniform", "returnsEstimator": "retrace",
                  "ERoldSeqFilter": "oldest", "nnLayerSizes": [128, 128], "maxTotObsNum": 4096, "minTotObsNum": 2000},
        steps=4, start_step=998, sample_seed=11, bounded=0, full_steps=[0, 1, 3]),
    "vracer_bounded": dict(
        replay=dict(seed=99, n_ep=12, ep_len=(30, 50), dS=17, dA=6),
        settings={"learner": "VRACER", "nnLayerSizes": [64, 64], "batchSize": 32, "maxTotObsNum": 1024,
                  "minTotObsNum": 300},
        steps=6, start_step=0, sample_seed=5, bounded=1, full_steps=list(range(6))),
    # a mini-batch beyond one wave of 4-sample tiles (1024 samples = 256 tiles > the persistent grid's worker CTAs): every CTA
    # loops over several tiles, no next-mini-batch staging, no helper CTAs — the regime of the batch-size sweep.
    # Oracle pinned; device case pending (tests/test_gpu_zz_pending.py): nothing above B = 256 has run on a GPU yet.
    "vracer_b1024": dict(
        replay=dict(seed=71, n_ep=60, ep_len=(40, 70), dS=12, dA=4),
        settings={"learner": "VRACER", "nnLayerSizes": [64, 64], "batchSize": 1024, "maxTotObsNum": 8192, "minTotObsNum": 2500},
        steps=3, start_step=998, sample_seed=17, bounded=0, full_steps=[0, 2]),
    # the batch size at which the device learner switches to the wide step by default (tcgen05 tiles of 128 sampled transitions,
    # wide_step.cuh): configs[1]'s own hidden layers, 32 tiles, three steps across the step-1000 sweep
    "vracer_b4096": dict(
        replay=dict(seed=83, n_ep=64, ep_len=(90, 130), dS=16, dA=4),
        settings={"learner": "VRACER", "nnLayerSizes": [128, 128], "batchSize": 4096, "maxTotObsNum": 16384, "minTotObsNum": 5000},
        steps=3, start_step=998, sample_seed=29, bounded=0, full_steps=[0, 2]),
    # hidden-layer functions other than Tanh (makeFunction, Network/Layers/Functions.h:643-668): settings/default.json asks for
    # SoftSign; HardSign and Sigm share its weight-initialisation factor
    "vracer_softsign": dict(
        replay=dict(seed=91, n_ep=24, ep_len=(30, 60), dS=6, dA=2),
        settings={"learner": "VRACER", "nnFunc": "SoftSign", "nnLayerSizes": [32, 32], "batchSize": 32, "maxTotObsNum": 4096,
                  "minTotObsNum": 600, "clipImpWeight": 4, "epsAnneal": 0, "nnLambda": 0},
        steps=4, start_step=997, sample_seed=13, bounded=0, full_steps=[0, 3]),
    "vracer_hardsign": dict(
        replay=dict(seed=92, n_ep=20, ep_len=(25, 50), dS=5, dA=3),
        settings={"learner": "VRACER", "nnFunc": "HardSign", "nnLayerSizes": [24, 24, 24], "batchSize": 16, "maxTotObsNum": 2048, "minTotObsNum": 400},
        steps=3, start_step=0, sample_seed=5, bounded=1, full_steps=[0, 2]),
    "racer_sigm": dict(
        replay=dict(seed=93, n_ep=20, ep_len=(25, 50), dS=7, dA=2),
        settings={"learner": "RACER", "nnFunc": "Sigm", "nnLayerSizes": [32, 32], "batchSize": 16, "maxTotObsNum": 2048, "minTotObsNum": 400},
        steps=3, start_step=0, sample_seed=9, bounded=0, full_steps=[0, 2]),
    "vracer_relu": dict(       # Relu: initFactor sqrt(2 / inputs) (Functions.h:404-413)
        replay=dict(seed=94, n_ep=20, ep_len=(25, 50), dS=6, dA=2),
        settings={"learner": "VRACER", "nnFunc": "Relu", "nnLayerSizes": [32, 32], "batchSize": 16, "maxTotObsNum": 2048, "minTotObsNum": 400},
        steps=3, start_step=0, sample_seed=21, bounded=0, full_steps=[0, 2]),
    "vracer_lrelu": dict(      # LRelu: initFactor sqrt(1 / inputs), slope PRELU_FAC = 0.1 below zero (Functions.h:16-18,452-468)
        replay=dict(seed=95, n_ep=20, ep_len=(25, 50), dS=6, dA=2),
        settings={"learner": "VRACER", "nnFunc": "LRelu", "nnLayerSizes": [32, 32], "batchSize": 16, "maxTotObsNum": 2048, "minTotObsNum": 400},
        steps=3, start_step=0, sample_seed=22, bounded=1, full_steps=[0, 2]),
    # the remaining names of makeFunction (Functions.h:643-668): initFactor sqrt(2 / inputs) for ExpPlus, SoftPlus, Exp and
    # sqrt(1 / inputs) for Linear; their evalDiff reads the pre-activation (or, Exp, the output)
    "vracer_expplus": dict(
        replay=dict(seed=96, n_ep=20, ep_len=(25, 50), dS=6, dA=2),
        settings={"learner": "VRACER", "nnFunc": "ExpPlus", "nnLayerSizes": [32, 32], "batchSize": 16, "maxTotObsNum": 2048, "minTotObsNum": 400},
        steps=3, start_step=0, sample_seed=23, bounded=0, full_steps=[0, 2]),
    "racer_softplus": dict(
        replay=dict(seed=97, n_ep=20, ep_len=(25, 50), dS=7, dA=2),
        settings={"learner": "RACER", "nnFunc": "SoftPlus", "nnLayerSizes": [32, 32], "batchSize": 16, "maxTotObsNum": 2048, "minTotObsNum": 400},
        steps=3, start_step=0, sample_seed=24, bounded=0, full_steps=[0, 2]),
    "vracer_exp": dict(
        replay=dict(seed=98, n_ep=20, ep_len=(25, 50), dS=6, dA=2),
        settings={"learner": "VRACER", "nnFunc": "Exp", "nnLayerSizes": [24, 24], "batchSize": 16, "maxTotObsNum": 2048, "minTotObsNum": 400},
        steps=3, start_step=0, sample_seed=25, bounded=1, full_steps=[0, 2]),
    "vracer_linear": dict(
        replay=dict(seed=99, n_ep=20, ep_len=(25, 50), dS=6, dA=2),
        settings={"learner": "VRACER", "nnFunc": "Linear", "nnLayerSizes": [32, 32], "batchSize": 16, "maxTotObsNum": 2048, "minTotObsNum": 400},
        steps=3, start_step=0, sample_seed=26, bounded=0, full_steps=[0, 2]),
    # "encoderLayerSizes": Learner_approximator::createEncoder (Learner_approximator.cpp:148-166) + RACER::setupNet
    # (RACER_common.cpp:82-91): the encoder layers and nnLayerSizes are stacked in the one network
    "vracer_encoder": dict(
        replay=dict(seed=101, n_ep=20, ep_len=(25, 50), dS=6, dA=2),
        settings={"learner": "VRACER", "encoderLayerSizes": [32], "nnLayerSizes": [32, 24], "batchSize": 16, "maxTotObsNum": 2048, "minTotObsNum": 400},
        steps=3, start_step=0, sample_seed=31, bounded=0, full_steps=[0, 2]),
    "racer_encoder2": dict(
        replay=dict(seed=102, n_ep=20, ep_len=(25, 50), dS=7, dA=2),
        settings={"learner": "RACER", "nnFunc": "SoftSign", "encoderLayerSizes": [32, 24], "nnLayerSizes": [24], "batchSize": 16, "maxTotObsNum": 2048, "minTotObsNum": 400},
        steps=3, start_step=0, sample_seed=32, bounded=1, full_steps=[0, 2]),
    # hidden sizes that GROW: the ParametricResidual after a layer links only the first min(size below, size) units
    # (ParametricResidualLayer::forward / backward, Layers.h:347-393); the other units' skip parameters get no gradient
    "vracer_widen": dict(
        replay=dict(seed=103, n_ep=20, ep_len=(25, 50), dS=6, dA=2),
        settings={"learner": "VRACER", "nnLayerSizes": [24, 32, 40], "batchSize": 16, "maxTotObsNum": 2048, "minTotObsNum": 400},
        steps=3, start_step=0, sample_seed=33, bounded=0, full_steps=[0, 2]),
    # FIFO pruning: capacity below the stored data, so applyEpisodesRemovalAlgo evicts on step 1
    "vracer_prune": dict(
        replay=dict(seed=17, n_ep=16, ep_len=(20, 30), dS=4, dA=2),
        settings={"learner": "VRACER", "nnLayerSizes": [16], "batchSize": 8, "maxTotObsNum": 256, "minTotObsNum": 100},
        steps=5, start_step=0, sample_seed=3, bounded=0, full_steps=list(range(5))),
    # RACER: Gaussian-shaped advantage head (Math/Gaus_advantage.h), unbounded and bounded actions
    "racer_small": dict(
        replay=dict(seed=41, n_ep=20, ep_len=(25, 55), dS=7, dA=3),
        settings={"learner": "RACER", "nnLayerSizes": [32, 32], "batchSize": 16, "maxTotObsNum": 2048, "minTotObsNum": 400},
        steps=10, start_step=995, sample_seed=9, bounded=0, full_steps=list(range(10))),
    "racer_bounded": dict(
        replay=dict(seed=43, n_ep=10, ep_len=(30, 40), dS=5, dA=2),
        settings={"learner": "RACER", "nnLayerSizes": [24], "batchSize": 8, "maxTotObsNum": 1024, "minTotObsNum": 200},
        steps=5, start_step=0, sample_seed=13, bounded=1, full_steps=list(range(5))),
    # recurrent: LSTM layer, BPTT window nnBPTTseq (Layer_LSTM.h, Network.h:155-193)
    "racer_lstm": dict(
        replay=dict(seed=51, n_ep=14, ep_len=(12, 40), dS=6, dA=2),
        settings={"learner": "RACER", "nnType": "LSTM", "nnLayerSizes": [16], "nnBPTTseq": 8, "batchSize": 8,
                  "maxTotObsNum": 1024, "minTotObsNum": 200},
        steps=6, start_step=997, sample_seed=21, bounded=0, full_steps=list(range(6))),
    "racer_lstm_encoder": dict(      # LSTM encoder layer under an LSTM layer (nnType applies to both, Approximator.cpp:265-270)
        replay=dict(seed=55, n_ep=14, ep_len=(12, 40), dS=6, dA=2),
        settings={"learner": "RACER", "nnType": "LSTM", "encoderLayerSizes": [16], "nnLayerSizes": [12], "nnBPTTseq": 8, "batchSize": 8,
                  "maxTotObsNum": 1024, "minTotObsNum": 200},
        steps=3, start_step=0, sample_seed=23, bounded=0, full_steps=list(range(3))),
    "vracer_lstm2": dict(
        replay=dict(seed=53, n_ep=10, ep_len=(6, 30), dS=5, dA=2),
        settings={"learner": "VRACER", "nnType": "LSTM", "nnLayerSizes": [12, 12], "nnBPTTseq": 5, "batchSize": 8,
                  "maxTotObsNum": 1024, "minTotObsNum": 100},
        steps=4, start_step=0, sample_seed=23, bounded=0, full_steps=list(range(4))),
    # returnsEstimator GAE (MemoryProcessing.cpp:411-417): Q[t] = r + gamma (V + lambda (Q[t+1] - V)), no importance weights
    "vracer_gae": dict(
        replay=dict(seed=29, n_ep=20, ep_len=(20, 60), dS=6, dA=3),
        settings={"learner": "VRACER", "returnsEstimator": "GAE", "lambda": 0.9, "nnLayerSizes": [32, 32], "batchSize": 16,
                  "maxTotObsNum": 2048, "minTotObsNum": 500},
        steps=10, start_step=995, sample_seed=17, bounded=0, full_steps=[0, 9]),
    # prioritized samplers (Sampling.cpp:173-296): which transitions are drawn depends on the stored TD errors; the PER
    # weights themselves are unused by RACER (Approximator.h:196).  Oracle-only so far.
    "vracer_pererr": dict(
        replay=dict(seed=81, n_ep=18, ep_len=(20, 50), dS=5, dA=2),
        settings={"learner": "VRACER", "dataSamplingAlgo": "PERerr", "nnLayerSizes": [24, 24], "batchSize": 16,
                  "maxTotObsNum": 2048, "minTotObsNum": 300},
        steps=8, start_step=0, sample_seed=51, bounded=0, full_steps=[0, 7]),
    "vracer_perseq": dict(
        replay=dict(seed=83, n_ep=18, ep_len=(20, 50), dS=5, dA=2),
        settings={"learner": "VRACER", "dataSamplingAlgo": "PERseq", "nnLayerSizes": [24, 24], "batchSize": 16,
                  "maxTotObsNum": 2048, "minTotObsNum": 300},
        steps=8, start_step=0, sample_seed=53, bounded=0, full_steps=[0, 7]),
    "vracer_perrank": dict(
        replay=dict(seed=85, n_ep=18, ep_len=(20, 50), dS=5, dA=2),
        settings={"learner": "VRACER", "dataSamplingAlgo": "PERrank", "nnLayerSizes": [24, 24], "batchSize": 16,
                  "maxTotObsNum": 2048, "minTotObsNum": 300},
        steps=8, start_step=0, sample_seed=55, bounded=0, full_steps=[0, 7]),
    # retraceExplore (MemoryProcessing.cpp:402-409): Retrace plus an exploration bonus on |Q - A - V|
    "vracer_explore": dict(
        replay=dict(seed=87, n_ep=20, ep_len=(20, 60), dS=6, dA=3),
        settings={"learner": "VRACER", "returnsEstimator": "retraceExplore", "lambda": 0.9, "nnLayerSizes": [32, 32],
                  "batchSize": 16, "maxTotObsNum": 2048, "minTotObsNum": 500},
        steps=10, start_step=995, sample_seed=57, bounded=0, full_steps=[0, 9]),
    # episode filters other than FIFO (getERfilterAlgo, MemoryProcessing.cpp:261-298): the whole episode vector is re-sorted
    # by a per-episode aggregate every step (unstable std::sort, many ties), more than 16 episodes so that the introsort
    # partitions, capacity below the stored data so that episodes are pruned.  Oracle-only so far.
    "vracer_farpolfrac": dict(
        replay=dict(seed=91, n_ep=40, ep_len=(10, 30), dS=4, dA=2),
        settings={"learner": "VRACER", "ERoldSeqFilter": "farpolfrac", "nnLayerSizes": [16], "batchSize": 16,
                  "maxTotObsNum": 700, "minTotObsNum": 300},
        steps=8, start_step=0, sample_seed=61, bounded=0, full_steps=[0, 7]),
    "vracer_maxkldiv": dict(
        replay=dict(seed=93, n_ep=40, ep_len=(10, 30), dS=4, dA=2),
        settings={"learner": "VRACER", "ERoldSeqFilter": "maxkldiv", "nnLayerSizes": [16], "batchSize": 16,
                  "maxTotObsNum": 700, "minTotObsNum": 300},
        steps=8, start_step=0, sample_seed=63, bounded=0, full_steps=[0, 7]),
    "vracer_minerror": dict(
        replay=dict(seed=95, n_ep=40, ep_len=(10, 30), dS=4, dA=2),
        settings={"learner": "VRACER", "ERoldSeqFilter": "minerror", "nnLayerSizes": [16], "batchSize": 16,
                  "maxTotObsNum": 700, "minTotObsNum": 300},
        steps=8, start_step=0, sample_seed=65, bounded=0, full_steps=[0, 7]),
    # one action component: clipImpWeight = sqrt(1/2) < 1, so the initial CinvRet = 1/C > 1 and every stored importance weight
    # counts as "was far" until the first recompute -> negative per-episode far-policy fractions and the wrap-around of
    # `Uint += float` in updateTrainingStatistics (oracle: uint_plus_float).  Oracle-only: the device count saturates instead.
    "vracer_da1": dict(
        replay=dict(seed=99, n_ep=16, ep_len=(20, 40), dS=4, dA=1),
        settings={"learner": "VRACER", "nnLayerSizes": [16, 16], "batchSize": 8, "maxTotObsNum": 1024, "minTotObsNum": 200},
        steps=8, start_step=0, sample_seed=69, bounded=0, full_steps=[0, 7]),
    # discrete action space (RACER<Discrete_advantage, Discrete_policy, Uint>: Math/Discrete_policy.h, Discrete_advantage.h):
    # 5 options, network outputs [V | advantages | policy].  Oracle-only so far.
    "racer_discrete": dict(
        replay=dict(seed=97, n_ep=20, ep_len=(20, 50), dS=6, dA=1, n_options=5),
        settings={"learner": "RACER", "nnLayerSizes": [32, 32], "batchSize": 16, "maxTotObsNum": 2048, "minTotObsNum": 300},
        steps=10, start_step=995, sample_seed=67, bounded=0, full_steps=[0, 9]),
    # MGU cells (Layer_GRU.h; "MGU" and "GRU" build the same layer, and it is what partially observable MDPs get by
    # default, Approximator.cpp:219-223): oracle-only so far — the device path does not cover them yet
    "racer_mgu": dict(
        replay=dict(seed=71, n_ep=14, ep_len=(12, 40), dS=6, dA=2),
        settings={"learner": "RACER", "nnType": "MGU", "nnLayerSizes": [16], "nnBPTTseq": 8, "batchSize": 8,
                  "maxTotObsNum": 1024, "minTotObsNum": 200},
        steps=6, start_step=997, sample_seed=41, bounded=0, full_steps=list(range(6))),
    "vracer_gru2": dict(
        replay=dict(seed=73, n_ep=10, ep_len=(6, 30), dS=5, dA=2),
        settings={"learner": "VRACER", "nnType": "GRU", "nnLayerSizes": [16, 16], "nnBPTTseq": 5, "batchSize": 8,
                  "maxTotObsNum": 1024, "minTotObsNum": 100},
        steps=4, start_step=0, sample_seed=43, bounded=1, full_steps=list(range(4))),
    # 64 LSTM cells: the shape at which the device recurrence keeps the recurrent weights in registers
    # (lstm_forward / lstm_backward, smarties_b200/csrc/step_kernels.cu) and P2 contracts 128 gate columns per tensor-core item
    "racer_lstm64": dict(
        replay=dict(seed=61, n_ep=12, ep_len=(20, 45), dS=6, dA=2),
        settings={"learner": "RACER", "nnType": "LSTM", "nnLayerSizes": [64], "nnBPTTseq": 12, "batchSize": 8,
                  "maxTotObsNum": 1024, "minTotObsNum": 200},
        steps=4, start_step=998, sample_seed=31, bounded=0, full_steps=[0, 3]),
    # BASELINE.json configs[2] itself on a small buffer: RACER + LSTM(64), nnBPTTseq 32, batch 128, state 32, action 8
    "racer_cfg3mini": dict(
        replay=dict(seed=77, n_ep=40, ep_len=(50, 80), dS=32, dA=8),
        settings={"learner": "RACER", "nnType": "LSTM", "nnLayerSizes": [64], "nnBPTTseq": 32, "batchSize": 128,
                  "clipImpWeight": 4, "explNoise": 0.1, "gamma": 0.99, "epsAnneal": 0, "nnLambda": 1e-6,
                  "maxTotObsNum": 4096, "minTotObsNum": 2000},
        steps=3, start_step=998, sample_seed=33, bounded=0, full_steps=[0, 2]),
}

# The same runs by a MULTI-THREADED reference (`threads` OpenMP threads: what bench.py's reference arm and the binding use).
# Sampling is thread-independent (generators[0]); what changes with the thread count is the far-policy count — the
# `Uint += float` partials of updateTrainingStatistics are per OpenMP thread with schedule(static, 1)
# (MemoryProcessing.cpp:202-227), deterministic for a given T — and, through it, beta; floats move by summation order only.
# (Checked when generating: four consecutive harness runs per case gave identical counts, beta and weights.)
for _base, _T in (("vracer_small", 8), ("vracer_small", 16), ("vracer_cfg2mini", 8), ("vracer_cfg2mini", 16), ("racer_small", 8)):
    CASES[f"{_base}_t{_T}"] = dict(CASES[_base], threads=_T)

# checkpoint cases: phase A runs `steps` steps and calls Learner_approximator::save(); phase B is a fresh process
# that calls restart() on those files and runs `steps_after` more steps.  Stored: the checkpoint files byte for
# byte ("ckpt:<file>") and the dumps of phase B ("ref2:<key>").
CKPT_CASES = {
    "vracer_ckpt": dict(
        replay=dict(seed=123, n_ep=24, ep_len=(20, 60), dS=6, dA=3),
        settings={"learner": "VRACER", "returnsEstimator": "retrace", "nnLayerSizes": [32, 32], "batchSize": 16,
                  "maxTotObsNum": 2048, "minTotObsNum": 500},
        steps=8, start_step=995, sample_seed=7, bounded=0, full_steps=[7], steps_after=4, sample_seed_after=19),
    "racer_lstm_ckpt": dict(
        replay=dict(seed=51, n_ep=14, ep_len=(12, 40), dS=6, dA=2),
        settings={"learner": "RACER", "nnType": "LSTM", "nnLayerSizes": [16], "nnBPTTseq": 8, "batchSize": 8,
                  "maxTotObsNum": 1024, "minTotObsNum": 200},
        steps=5, start_step=0, sample_seed=21, bounded=1, full_steps=[4], steps_after=3, sample_seed_after=29),
    # "targetDelay" (AdamOptimizer::tgtUpdateAlpha, Optimizer.cpp:162-177): RACER never evaluates the target weights, they only
    # reach <name>_net_tgt_weights.raw.  0.05: exponential average after every update; 3: copy at updates 1, 4, 7 (the checkpoint
    # after 8 updates holds the weights after update 7; a restarted process copies at its first update again)
    "vracer_tgt_ema": dict(
        replay=dict(seed=124, n_ep=24, ep_len=(20, 60), dS=6, dA=3),
        settings={"learner": "VRACER", "returnsEstimator": "retrace", "nnLayerSizes": [32, 32], "batchSize": 16, "targetDelay": 0.05,
                  "maxTotObsNum": 2048, "minTotObsNum": 500},
        steps=8, start_step=0, sample_seed=8, bounded=0, full_steps=[7], steps_after=4, sample_seed_after=20),
    "vracer_tgt_copy": dict(
        replay=dict(seed=125, n_ep=24, ep_len=(20, 60), dS=6, dA=3),
        settings={"learner": "VRACER", "returnsEstimator": "retrace", "nnLayerSizes": [32, 32], "batchSize": 16, "targetDelay": 3,
                  "maxTotObsNum": 2048, "minTotObsNum": 500},
        steps=8, start_step=0, sample_seed=9, bounded=0, full_steps=[7], steps_after=4, sample_seed_after=21),
}
CKPT_FILES = ["agent_00_net_weights.raw", "agent_00_net_tgt_weights.raw", "agent_00_net_1stMom.raw", "agent_00_net_2ndMom.raw",
              "agent_00_scaling.raw", "agent_00_rank_000_learner_status.raw", "agent_00_rank_000_learner_data.raw"]

BIG = ("/weights", "/m1", "/m2", "/gradSum")


def make_replay(spec):
    r = dict(spec["replay"])
    if "n_options" in r:       # discrete action space: dA = 1, the behaviour policy is a probability vector
        r.pop("dA")
        return synth.make_replay_discrete(**r)
    return synth.make_replay(**r)


def harness_flags(spec):
    return ["--discrete", str(spec["replay"]["n_options"])] if "n_options" in spec["replay"] else []


def run_case(name, spec, outdir):
    d = make_replay(spec)
    ckpt, D2 = {}, {}
    with tempfile.TemporaryDirectory() as tmp:
        synth.write_replay_file(os.path.join(tmp, "data.bin"), d)
        with open(os.path.join(tmp, "settings.json"), "w") as f:
            json.dump(spec["settings"], f)
        T = str(spec.get("threads", 1))
        cmd = [HARNESS, "--data", "data.bin", "--settings", "settings.json", "--steps", str(spec["steps"]),
               "--threads", T, "--startStep", str(spec["start_step"]), "--sampleSeed", str(spec["sample_seed"]),
               "--bounded", str(spec["bounded"]), "--dump", "out.bin", "--dumpAll", "--quiet"] + harness_flags(spec)
        env = dict(os.environ, OMP_NUM_THREADS=T)
        if "steps_after" in spec:
            cmd.append("--save")
        subprocess.run(cmd, cwd=tmp, check=True, stdout=subprocess.DEVNULL, env=env)
        D = synth.read_dump(os.path.join(tmp, "out.bin"))
        if "steps_after" in spec:
            for fn in CKPT_FILES:
                ckpt[fn] = np.fromfile(os.path.join(tmp, fn), dtype=np.uint8)
            tmp2 = os.path.join(tmp, "phaseB")
            os.makedirs(tmp2)
            cmd = [HARNESS, "--data", "../data.bin", "--settings", "../settings.json", "--steps", str(spec["steps_after"]),
                   "--threads", "1", "--sampleSeed", str(spec["sample_seed_after"]), "--bounded", str(spec["bounded"]),
                   "--restart", tmp, "--dump", "out2.bin", "--dumpAll", "--quiet"]
            subprocess.run(cmd, cwd=tmp2, check=True, stdout=subprocess.DEVNULL, env=env)
            D2 = synth.read_dump(os.path.join(tmp2, "out2.bin"))
    keep = {}
    for fn, b in ckpt.items():
        keep["ckpt:" + fn] = b
    for k, v in D2.items():
        keep["ref2:" + k] = v
    for k, v in D.items():
        if k.startswith("s") and k[1].isdigit():
            s = int(k[1:k.index("/")])
            if any(k.endswith(b) for b in BIG) and s not in spec["full_steps"]:
                continue
            if k.endswith("/m1") or k.endswith("/m2"):
                if s != spec["steps"] - 1:
                    continue
        keep["ref:" + k] = v
    for k in ("N", "term", "start", "S", "A", "MU", "R"):
        keep["replay:" + k] = d[k]
    keep["spec"] = np.frombuffer(json.dumps(dict(spec, name=name)).encode(), dtype=np.uint8)
    path = os.path.join(outdir, name + ".npz")
    np.savez_compressed(path, **keep)
    print(name, os.path.getsize(path) // 1024, "KiB", len(keep), "arrays")


def run_grad_stats(outdir):
    """outgrad_stats.npz: for every case the file `agent_00_net_outGrad_stats.raw` the reference's StatsTracker
    (Utils/StatsTracker.cpp:66-89) wrote during the same run as the case's .npz (as float32 words; empty when the run
    crossed no step with nGradSteps % 1000 == 0)."""
    keep = {}
    for name, spec in CASES.items():
        d = make_replay(spec)
        with tempfile.TemporaryDirectory() as tmp:
            synth.write_replay_file(os.path.join(tmp, "data.bin"), d)
            with open(os.path.join(tmp, "settings.json"), "w") as f:
                json.dump(spec["settings"], f)
            T = str(spec.get("threads", 1))
            cmd = [HARNESS, "--data", "data.bin", "--settings", "settings.json", "--steps", str(spec["steps"]),
                   "--threads", T, "--startStep", str(spec["start_step"]), "--sampleSeed", str(spec["sample_seed"]),
                   "--bounded", str(spec["bounded"]), "--quiet"] + harness_flags(spec)
            subprocess.run(cmd, cwd=tmp, check=True, stdout=subprocess.DEVNULL, env=dict(os.environ, OMP_NUM_THREADS=T))
            fn = os.path.join(tmp, "agent_00_net_outGrad_stats.raw")
            keep[name] = np.fromfile(fn, dtype=np.float32) if os.path.exists(fn) else np.zeros(0, np.float32)
        print(name, "outGrad_stats:", keep[name].size, "floats")
    np.savez_compressed(os.path.join(outdir, "outgrad_stats.npz"), **keep)


if __name__ == "__main__":
    if not os.path.exists(HARNESS):
        sys.exit("build oracle/_ref first: make -C oracle")
    if sys.argv[1:] == ["outgrad_stats"]:
        run_grad_stats(os.path.dirname(os.path.abspath(__file__)))
        sys.exit(0)
    only = sys.argv[1:]
    for n, s in {**CASES, **CKPT_CASES}.items():
        if (not only and n in CASES) or n in only:
            run_case(n, s, os.path.dirname(os.path.abspath(__file__)))
