"""End-to-end (host sampler + H2D ids + D2H statistics through smb200_train_steps) against device-resident steps at a given batch."""
import os, sys, time
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from smarties_b200 import Learner, synth
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
K = max(20, min(2000, 2_000_000 // B))
d = synth.make_replay(123, 1000, 1000, 32, 8)
L = Learner(32, 8, {"maxTotObsNum": 1048576, "minTotObsNum": 1000000, "batchSize": B})
L.load_replay(d); L.initialize_learner(); L.seed_sampler(7)
L.train_steps(10, want_stats=True)
for _ in range(2):
    t0 = time.perf_counter(); L.train_steps(K, want_stats=True); dt = time.perf_counter() - t0
    print(f"B={B} e2e {K} steps: {1e6*dt/K:.1f} us/step -> {B*K/dt:.3e} tr/s; device ms {L.last_timing()[0]:.2f} kernel {L.step_kernel()}")
L.presample(K); L.train_presampled(0, 10); L.sync()
L.train_presampled(10, K - 10); L.sync(); print("device-resident:", 1e3 * L.last_timing()[0] / (K - 10), "us/step")
L.close()
