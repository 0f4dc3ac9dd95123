import os, sys, time
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from smarties_b200 import Learner, synth
d = synth.make_replay(123, 1000, 1000, 32, 8)
L = Learner(32, 8, {"maxTotObsNum": 1048576, "minTotObsNum": 1000000})
L.load_replay(d); L.initialize_learner(); L.seed_sampler(7)
L.train_steps(1000, want_stats=True)
for K in (5000, 5000):
    t0 = time.perf_counter(); L.train_steps(K, want_stats=True); dt = time.perf_counter() - t0
    print(f"e2e {K} steps: {1e6*dt/K:.2f} us/step -> {256*K/dt:.3e} tr/s; device ms {L.last_timing()[0]:.2f}")
L.presample(3000); L.train_presampled(0, 1000); L.sync()
L.train_presampled(1000, 2000); L.sync(); print("device-resident:", 1e3 * L.last_timing()[0] / 2000, "us/step")
L.close()
