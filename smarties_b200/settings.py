"""settings/*.json surface of the reference (Settings/HyperParameters.{h,cpp}).

Same keys, same defaults (the CODE defaults of HyperParameters.h:37-73, which differ from
the README), same derived quantities (defineDistributedLearning, HyperParameters.cpp:178-205)
and the same validation (check(), :207-224).
"""
from __future__ import annotations

import json
import math
import os

FLT_EPS = 1.1920928955078125e-07

_KEYS = ["learner", "ERoldSeqFilter", "dataSamplingAlgo", "returnsEstimator", "explNoise", "gamma", "lambda",
         "obsPerStep", "clipImpWeight", "penalTol", "klDivConstraint", "targetDelay", "epsAnneal", "minTotObsNum",
         "maxTotObsNum", "saveFreq", "encoderLayerSizes", "nnLayerSizes", "batchSize", "ESpopSize", "nnBPTTseq",
         "nnLambda", "learnrate", "outWeightsPrefac", "nnOutputFunc", "nnFunc", "nnType"]


class HyperParameters:
    def __init__(self, dimS: int, dimA: int, overrides: dict | str | None = None):
        # HyperParameters.h:37-73
        self.learner = "VRACER"
        self.ERoldSeqFilter = "oldest"
        self.dataSamplingAlgo = "uniform"
        self.returnsEstimator = "default"
        self.explNoise = math.sqrt(0.2)
        self.gamma = 0.995
        self.lambda_ = 1.0
        self.obsPerStep = 1.0
        self.clipImpWeight = math.sqrt(dimA / 2.0)
        self.penalTol = 0.1
        self.klDivConstraint = 0.01
        self.targetDelay = 0.0
        self.epsAnneal = 5e-7
        self.minTotObsNum = 0
        self.maxTotObsNum = int(2 ** 14 * math.sqrt(dimA + dimS))
        self.saveFreq = 50000
        self.encoderLayerSizes = [0]
        self.nnLayerSizes = [128, 128]
        self.batchSize = 256
        self.ESpopSize = 1
        self.nnBPTTseq = 16
        self.nnLambda = FLT_EPS
        self.learnrate = 1e-4
        self.outWeightsPrefac = 1e-3
        self.nnOutputFunc = "Linear"
        self.nnFunc = "Tanh"
        self.nnType = "FFNN"
        self.batchSize_local = 0
        self.minTotObsNum_local = 0
        self.maxTotObsNum_local = 0
        self.bRecurrent = False
        if isinstance(overrides, str):
            with open(overrides) as f:
                overrides = json.load(f)
        for k, v in (overrides or {}).items():   # initializeOpts, HyperParameters.cpp:123-176
            if k not in _KEYS:
                raise KeyError(f"unknown settings key {k!r}")
            setattr(self, "lambda_" if k == "lambda" else k, v)
        if self.returnsEstimator == "default":   # AlgoFactory.cpp:134-136 for RACER / VRACER
            self.returnsEstimator = "retrace"
        self.define_distributed_learning(1)
        self.check()

    def define_distributed_learning(self, n_learners: int):
        """HyperParameters::defineDistributedLearning (HyperParameters.cpp:178-205)."""
        nL = float(n_learners)
        if self.batchSize > 1:
            self.batchSize = int(math.ceil(self.batchSize / nL) * nL)
            self.batchSize_local = self.batchSize // n_learners
        else:
            self.batchSize_local = self.batchSize
        if self.minTotObsNum <= 0:
            self.minTotObsNum = self.maxTotObsNum
        self.minTotObsNum = min(self.minTotObsNum, self.maxTotObsNum)
        self.minTotObsNum = int(math.ceil(self.minTotObsNum / nL) * nL)
        self.minTotObsNum_local = self.minTotObsNum // n_learners
        self.maxTotObsNum = int(math.ceil(self.maxTotObsNum / nL) * nL)
        self.maxTotObsNum_local = self.maxTotObsNum // n_learners

    def check(self):
        """HyperParameters::check (HyperParameters.cpp:207-224) + what the device path supports."""
        self.bRecurrent = self.nnType in ("LSTM", "RNN", "MGU", "GRU")
        for name, bad in (("targetDelay<0", self.targetDelay < 0), ("obsPerStep<0", self.obsPerStep < 0),
                          ("learnrate>1", self.learnrate > 1), ("learnrate<0", self.learnrate < 0),
                          ("explNoise<0", self.explNoise < 0), ("epsAnneal<0", self.epsAnneal < 0),
                          ("batchSize<0", self.batchSize <= 0), ("nnLambda<0", self.nnLambda < 0),
                          ("gamma<0", self.gamma < 0), ("gamma>1", self.gamma > 1)):
            if bad:
                raise ValueError(name)
        if self.epsAnneal > 0.0001:
            self.epsAnneal = 5e-7   # "epsAnneal should be tiny. It will be set to 5e-7 for this run."
        unsupported = []
        if self.learner not in ("VRACER", "RACER"):
            unsupported.append(f"learner={self.learner}")
        if self.dataSamplingAlgo not in ("uniform", "PERrank", "PERerr", "PERseq"):
            unsupported.append(f"dataSamplingAlgo={self.dataSamplingAlgo}")
        # "retraceExplore" is not an affine recursion: it runs on the sequential sweep kernel k_sweep_explore
        if self.returnsEstimator not in ("retrace", "GAE", "retraceExplore"):
            unsupported.append(f"returnsEstimator={self.returnsEstimator}")
        if self.ERoldSeqFilter not in ("oldest", "default", "farpolfrac", "maxkldiv", "minerror"):
            unsupported.append(f"ERoldSeqFilter={self.ERoldSeqFilter}")
        # hidden-layer functions of makeFunction (Functions.h:643-668) the device evaluates; recurrent cells keep Tanh
        func_ok = self.nnFunc == "Tanh" or (self.nnType == "FFNN" and self.nnFunc in ("SoftSign", "HardSign", "Sigm", "Relu", "LRelu", "ExpPlus", "SoftPlus", "Exp", "Linear"))
        if self.nnType not in ("FFNN", "LSTM", "MGU", "GRU") or not func_ok or self.nnOutputFunc != "Linear":
            unsupported.append(f"nnType/nnFunc/nnOutputFunc={self.nnType}/{self.nnFunc}/{self.nnOutputFunc}")
        if self.ESpopSize != 1:
            unsupported.append("ESpopSize")
        if unsupported:
            raise NotImplementedError("smarties_b200 device path does not cover: " + ", ".join(unsupported))
