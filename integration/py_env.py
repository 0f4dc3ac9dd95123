#!/usr/bin/env python3
"""integration/py_env.py — a small Python environment written against the reference's PYTHON app API
(`import smarties`: the pybind11 module of /root/reference/source/smarties/smarties_pybind11.cpp, built by
integration/Makefile), in the shape of apps/cart_pole_py: `Engine(sys.argv).run(app_main)` with
`comm.setStateActionDims / setActionScales / sendInitState / recvAction / sendState / sendTermState / sendLastState`.
Ours, not a copy of a reference app: a damped point mass that has to be steered to the origin with a bounded force.
Used by the drop-in tests to show that a Python app runs unchanged on either learner.

    PYTHONPATH=oracle/_ref/b200/py SMARTIES_B200=1 python integration/py_env.py --nTrainSteps 3000 --nThreads 4
"""
import sys

import numpy as np
import smarties as rl


def app_main(comm):
    comm.setStateActionDims(4, 2)
    comm.setActionScales([1.0, 1.0], [-1.0, -1.0], areBounds=True)
    rng = np.random.default_rng(12345)
    dt = 0.1
    while True:
        pos, vel = rng.uniform(-1, 1, 2), rng.uniform(-0.2, 0.2, 2)
        comm.sendInitState(np.concatenate([pos, vel]))
        for step in range(1, 201):
            a = np.asarray(comm.recvAction())
            vel = 0.95 * vel + dt * a
            pos = pos + dt * vel
            reward = -float(pos @ pos) - 0.01 * float(a @ a)
            state = np.concatenate([pos, vel])
            if np.abs(pos).max() > 3.0:
                comm.sendTermState(state, reward - 10.0)      # left the arena: terminal
                break
            if step == 200:
                comm.sendLastState(state, reward)              # time limit: truncated, not terminal
                break
            comm.sendState(state, reward)


if __name__ == "__main__":
    e = rl.Engine(sys.argv)
    if e.parse():
        sys.exit()
    e.run(app_main)
