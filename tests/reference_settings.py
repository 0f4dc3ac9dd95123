"""The RACER-family settings files the reference ships (settings/*.json, values restated here: the tests must not read
/root/reference).  DEVICE: what the device path covers; REFERENCE_ONLY: what stays with the reference's CPU learner."""
DEVICE = {
    "RACER.json": {"learner": "RACER", "epsAnneal": 5e-7},
    "RACER_RNN.json": {"learner": "RACER", "nnLayerSizes": [32, 32], "gamma": 0.99, "epsAnneal": 0, "nnLambda": 1e-6, "penalTol": 0.1,
                       "clipImpWeight": 4, "maxTotObsNum": 262144, "nnType": "LSTM", "explNoise": 0.1, "batchSize": 128, "learnrate": 0.0001},
    "RACER_glider.json": {"learner": "RACER", "nnLayerSizes": [128, 128, 128], "gamma": 1, "epsAnneal": 2e-7, "nnLambda": 1e-6,
                          "penalTol": 0.05, "clipImpWeight": 1, "maxTotObsNum": 524288},
    "VRACER.json": {"learner": "VRACER", "dataSamplingAlgo": "uniform", "returnsEstimator": "retrace", "ERoldSeqFilter": "oldest",
                    "nnLayerSizes": [128, 128]},
    "VRACER_LES.json": {"learner": "VRACER", "batchSize": 256, "clipImpWeight": 1, "epsAnneal": 0, "penalTol": 0.05, "explNoise": 0.5,
                        "gamma": 0.99, "learnrate": 0.00001, "minTotObsNum": 1048576, "maxTotObsNum": 1048576, "nnLayerSizes": [32, 32],
                        "obsPerStep": 64, "ERoldSeqFilter": "oldest", "outWeightsPrefac": 0.00001},
    "VRACER_expensiveData.json": {"learner": "VRACER", "batchSize": 128, "clipImpWeight": 1, "penalTol": 0.1, "epsAnneal": 0, "explNoise": 0.2,
                                  "gamma": 0.99, "learnrate": 0.0001, "minTotObsNum": 4096, "maxTotObsNum": 32768, "nnLayerSizes": [32, 32],
                                  "nnType": "GRU", "saveFreq": 10000, "obsPerStep": 1, "outWeightsPrefac": 0.01},
    "default.json": {"learner": "VRACER", "ERoldSeqFilter": "oldest", "ESpopSize": 1, "batchSize": 256, "clipImpWeight": 4,
                     "dataSamplingAlgo": "uniform", "encoderLayerSizes": [0], "epsAnneal": 0, "explNoise": 0.4472135955, "gamma": 0.995,
                     "klDivConstraint": 0.01, "lambda": 0.95, "learnrate": 0.0001, "maxTotObsNum": 262144, "minTotObsNum": 131072,
                     "nnBPTTseq": 16, "nnFunc": "SoftSign", "nnLambda": 0, "nnLayerSizes": [128, 128], "nnOutputFunc": "Linear",
                     "nnType": "FFNN", "obsPerStep": 1, "outWeightsPrefac": 0.1, "penalTol": 0.1, "saveFreq": 200000, "targetDelay": 0},
    # RACER_atari.json without the convolutional preprocessing its app asks for (the hidden layer and the hyper-parameters)
    "RACER_atari.json (dense part)": {"learner": "RACER", "batchSize": 128, "clipImpWeight": 4, "epsAnneal": 0, "explNoise": 0.05, "gamma": 0.99,
                                      "learnrate": 0.0001, "maxTotObsNum": 262144, "minTotObsNum": 131072, "nnLayerSizes": [512], "obsPerStep": 1},
}
REFERENCE_ONLY = {
    "VRACER_CMA.json": {"learner": "VRACER", "batchSize": 60, "ESpopSize": 60, "clipImpWeight": 4, "epsAnneal": 0, "explNoise": 0.447214,
                        "gamma": 0.995, "learnrate": 0.001, "maxTotObsNum": 262144, "nnLayerSizes": [64, 64], "obsPerStep": 1, "outWeightsPrefac": 0.01},
}
