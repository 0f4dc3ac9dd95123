"""Print the measured GPU-vs-reference error of every compared quantity for each golden case
(run on the GPU box; output is quoted in DESIGN.md)."""
import os
import sys

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
from parity_utils import CASES, RECURRENT_CASES, Golden, make_learner, relerr  # noqa: E402

for case in (sys.argv[1:] or CASES + RECURRENT_CASES):
    g = Golden(case)
    L = make_learner(g)
    R = g.ref
    print(f"== {case}: init Qret abs {np.abs(L.read_field('QRET') - R['init/Qret']).max():.2e} "
          f"(max |Q| {np.abs(R['init/Qret']).max():.2f})")
    for s in range(g.steps):
        st = L.train_steps(1)[0]
        pre = f"s{s}"
        O, gg, X = L.get_last_batch()
        line = [f"s{s}", f"X_exact={np.array_equal(X, R[pre + '/S'])}", f"O {np.abs(O - R[pre + '/O']).max():.2e}",
                f"g {relerr(gg, R[pre + '/g']):.2e}"]
        if pre + "/gradSum" in R:
            line += [f"G {relerr(L.get_grad(), R[pre + '/gradSum']):.2e}", f"W {np.abs(L.get_weights() - R[pre + '/weights']).max():.2e}"]
        ref = g.refer(pre + "/post")
        line += [f"nFar {st['n_far_policy']}/{int(ref[3])} exact {st['n_far_exact']}", f"dbeta {st['beta'] - ref[0]:.1e}",
                 f"cmax_eq {st['cmax'] == ref[1]}"]
        print("  ", " ".join(line))
    for f, k in (("QRET", "Qret"), ("V", "V"), ("RHO", "rho"), ("KL", "KL"), ("DELTA", "delta")):
        a, b = L.read_field(f), R["final/" + k]
        print(f"   final {f}: abs {np.abs(a - b).max():.2e} rel {relerr(a, b):.2e}")
    _, _, agg = L.read_episodes()
    print("   final epAgg abs", np.abs(agg[:, :8] - R["final/epAgg"][:, :8]).max(axis=0))
    mean, scale, std, rew = L.get_scaling()
    print("   scaling", np.abs(mean - R["final/stateMean"]).max(), np.abs(scale - R["final/stateScale"]).max(), rew, R["final/rewards"])
    L.close()
