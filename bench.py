#!/usr/bin/env python
"""bench.py — V-RACER / RACER learner throughput (transitions/s updated) on the BASELINE.json workloads.

    python bench.py --gpus N --steps K --warmup W                      # our arm (CUDA, through the C-ABI)
    python bench.py --impl reference --gpus N --steps K --warmup W     # the reference's CPU path (oracle/_ref)
    python bench.py --workload cfg3                                    # BASELINE.json configs[2]: RACER + LSTM(64), BPTT 32, batch 128
    python bench.py --gpus N --scaling strong                          # fixed global batch 256 split over N ranks

Workload cfg2 (BASELINE.json configs[1], the one the metric is quoted on): synthetic MemoryBuffer of 1 000 000 transitions
(1000 episodes x 1000 steps, every 10th terminal), state_dim 32, act_dim 8, MLP(128,128), settings/VRACER.json defaults
(batchSize 256, ...).  One "step" = one learner step {sample -> gather -> forward -> ReF-ER/Retrace loss -> backward -> Adam ->
replay statistics}, INCLUDING the every-1000-steps full-buffer Retrace + reward/state-moment sweeps when the timed region
crosses one.

value  = batch x K / device time (CUDA events on the library's stream, sampled transition ids already resident in HBM,
         max over ranks);
e2e    = the same through smb200_train_steps with HOST buffers: the sampled ids are produced on host cores by the bit-exact
         std::mt19937 sampler and copied H2D, per-step statistics are copied D2H, all inside the timed region (wall clock,
         barrier on both sides).
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "V-RACER learner transitions/sec updated"
REPLAY = dict(n_ep=1000, ep_len=1000, dS=32, dA=8, seed=123)
WORKLOADS = {
    "cfg2": dict(
        name="cfg2: 1M-transition synthetic replay, state_dim=32 act_dim=8, MLP(128,128), VRACER.json, batch 256",
        settings={"learner": "VRACER", "dataSamplingAlgo": "uniform", "returnsEstimator": "retrace", "ERoldSeqFilter": "oldest",
                  "nnLayerSizes": [128, 128], "maxTotObsNum": 1048576, "minTotObsNum": 1000000},
        batch=256, n_params=23064, window=1),
    # SURVEY.md §8d cfg3; the shipped settings/RACER_RNN.json says [32, 32] LSTM cells and nnBPTTseq 16 — BASELINE.json's
    # "LSTM(64) sequences len=32" is this override
    "cfg3": dict(
        name="cfg3: 1M-transition synthetic replay, state_dim=32 act_dim=8, RACER + LSTM(64), nnBPTTseq 32, batch 128",
        settings={"learner": "RACER", "nnType": "LSTM", "nnLayerSizes": [64], "nnBPTTseq": 32, "batchSize": 128,
                  "clipImpWeight": 4, "explNoise": 0.1, "gamma": 0.99, "epsAnneal": 0, "nnLambda": 1e-6,
                  "maxTotObsNum": 1048576, "minTotObsNum": 1000000},
        batch=128, n_params=None, window=33),
}
# kept for the scripts and tests that import them
WORKLOAD = dict(REPLAY, name=WORKLOADS["cfg2"]["name"])
SETTINGS = WORKLOADS["cfg2"]["settings"]
BATCH = 256
N_PARAMS = 23064
# algorithmic bytes
BYTES_PER_TRANSITION = 268                      # replay read 248 B + write-back 20 B
BYTES_RETRACE_PER_TRANSITION = 24               # r,V,A,rho,Q read + Q written
BYTES_MOMENTS_PER_TRANSITION = (32 + 1) * 4
BYTES_FUSED_SWEEP_PER_TRANSITION = BYTES_RETRACE_PER_TRANSITION + BYTES_MOMENTS_PER_TRANSITION   # 156 B


def step_bytes(batch, n_params, window=1):
    """Algorithmic bytes of one learner step: 268 B per sampled transition (+ 128 B of state per extra window step of a
    recurrent net) + 7 x 4 B per parameter of Adam traffic (w, m, v, g read; w, m, v written)."""
    return (BYTES_PER_TRANSITION + 128 * (window - 1)) * batch + 7 * 4 * n_params


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


def ncu_traffic(kernel, steps=None):
    """DRAM bytes per launch from the committed ncu --set full capture (scripts/ncu_capture.sh), or None."""
    for rnd in ("r2", "r1"):
        p = os.path.join(ROOT, "profiles", rnd, "ncu_traffic.json")
        if not os.path.exists(p):
            continue
        with open(p) as f:
            t = json.load(f)
        for k, v in t.items():
            if isinstance(v, dict) and kernel in k:
                b = v["dram_bytes"]
                return (b * steps / t["_steps_per_persistent_launch"] if steps else b), f"profiles/{rnd}/ncu_traffic.json"
    return None, None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.idx, self.lines, self.p = gpu_index, [], None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                       "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=lambda: [self.lines.append(l) for l in self.p.stdout], daemon=True).start()
        except OSError:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        sm, mx, reasons = [], [], set()
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def make_workload(rank=0, n_ep=None):
    from smarties_b200 import synth
    w = REPLAY
    return synth.make_replay(w["seed"] + rank, n_ep or w["n_ep"], w["ep_len"], w["dS"], w["dA"])


def bench_config(wl, world, scaling, batch_local, replay_per_gpu):
    """The `config` object — identical keys and values in both arms (ours and --impl reference)."""
    return {"workload": WORKLOADS[wl]["name"], "batch_per_gpu": batch_local, "global_batch": batch_local * world,
            "replay_transitions_per_gpu": replay_per_gpu, "scaling": scaling,
            "l2_policy": "replay buffer (252 MB per 1M transitions) larger than L2; sampled rows are random",
            "parallelism": f"dp{world}"}


def plan(args):
    """(batch per GPU, episodes per GPU) of the run.  weak: every rank keeps a 1M-transition shard and the workload's batch;
    strong (`batchSize / nLearners` of Settings/HyperParameters.cpp:178-205): the workload's batch is the
    GLOBAL batch and an 8M-transition buffer is sharded over the ranks."""
    w = WORKLOADS[args.workload]
    world = max(1, args.gpus)
    if args.scaling == "strong":
        if w["batch"] % world:
            sys.exit("bench.py: --scaling strong needs a batch divisible by the number of GPUs")
        return (args.batch or w["batch"]) // world, 8 * REPLAY["n_ep"] // world
    return args.batch or w["batch"], REPLAY["n_ep"]


# ------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the reference's own CPU implementation (oracle/_ref)
# ------------------------------------------------------------------------------------------
def run_reference(steps, threads, data=None, reps=1, settings=None, learner_flag=None, warmup=0):
    """Times `steps` learner steps of the UNMODIFIED reference (oracle/_ref/ref_harness, built from
    /root/reference by oracle/Makefile) on this box's host cores, after `warmup` untimed steps of the same loop.
    Falls back to the numpy oracle port if the harness binary is absent."""
    harness = os.path.join(ROOT, "oracle", "_ref", "ref_harness")
    from smarties_b200 import synth
    if data is None:
        data = make_workload()
    settings = settings or SETTINGS
    batch = settings.get("batchSize", 256)
    if os.path.exists(harness):
        with tempfile.TemporaryDirectory() as tmp:
            synth.write_replay_file(os.path.join(tmp, "data.bin"), data)
            with open(os.path.join(tmp, "settings.json"), "w") as f:
                json.dump(settings, f)
            env = dict(os.environ, OMP_NUM_THREADS=str(threads), OMP_PROC_BIND="close")
            out = subprocess.run([harness, "--data", "data.bin", "--settings", "settings.json", "--steps", str(steps),
                                  "--threads", str(threads), "--sampleSeed", "7", "--quiet", "--reps", str(reps),
                                  "--warmup", str(warmup)],
                                 cwd=tmp, env=env, check=True, capture_output=True, text=True).stdout
        line = [l for l in out.splitlines() if l.startswith('{"harness"')][-1]
        r = json.loads(line)
        med = reps >= 3 and "transitions_per_s_median" in r          # >= 3 repetitions, median
        n_tr = int(np.sum(np.asarray(data["N"]) - 1))
        return dict(value=r["transitions_per_s_median"] if med else r["transitions_per_s"],
                    seconds=r["seconds_median"] if med else r["seconds_mean"], steps=steps, kind="reference", cores=threads,
                    sample=(f"median of {reps} consecutive runs of " if med else "") +
                           f"{steps} learner steps (after {warmup} untimed warm-up steps) of the same workload ({n_tr}-transition buffer, "
                           f"batch {batch}, sweeps every 1000 steps) by the unmodified reference built -O3 -ffast-math "
                           f"-DSINGLE_PREC, {threads} OpenMP threads, OMP_PROC_BIND=close")
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import vracer_oracle as vo
    o = vo.VracerOracle(32, 8, batch=BATCH, max_tot_obs=1048576)
    rng = np.random.default_rng(0)
    o.W[:] = (0.05 * rng.standard_normal(o.layout.n_params)).astype(np.float32)
    small = {k: (v[:50] if k in ("N", "term", "start") else v) for k, v in data.items()}
    o.load_replay(small)
    o.initialize_learner()
    n = max(1, min(steps, 5))
    t0 = time.perf_counter()
    for _ in range(n):
        o.train_step()
    dt = time.perf_counter() - t0
    return dict(value=BATCH * n / dt, seconds=dt, steps=n, kind="port", cores=1,
                sample=f"{n} learner steps of the numpy oracle port on a 50-episode slice (reference harness binary absent)")


def cpu_model():
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def cpu_baseline_block(r, data=None, single_thread_steps=400, settings=None):
    """The `cpu_baseline` object: the all-cores run `r` plus the CPU model and the same harness on ONE thread
    over a shorter sample."""
    blk = {"value": r["value"], "unit": "transitions/s", "cores": r["cores"], "kind": r["kind"], "sample": r["sample"],
           "cpu_model": cpu_model()}
    if r["kind"] == "reference" and r["cores"] > 1 and single_thread_steps > 0:
        r1 = run_reference(single_thread_steps, 1, data=data, reps=3, settings=settings, warmup=5)
        blk["single_thread"] = {"value": r1["value"], "unit": "transitions/s", "sample": r1["sample"]}
    return blk


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w = WORKLOADS[args.workload]
    world = max(1, int(os.environ.get("WORLD_SIZE", str(args.gpus))))
    batch_local, n_ep = plan(args)
    threads = os.cpu_count() or 1
    steps = min(args.steps, 20000)
    # the single-process reference has no learner ranks here (no MPI in the image): it runs the GLOBAL batch of the
    # configuration on one shard-sized buffer — a bounded sample of the N-GPU workload
    settings = dict(w["settings"], batchSize=batch_local * world)
    data = make_workload(0, min(n_ep, REPLAY["n_ep"]))
    # warm-up steps inside the harness process, before its timed loop; the short timed loop is repeated and the median taken
    # so that the driver's 20-step runs are not a single cold sample
    reps = 5 if steps <= 200 else (3 if steps <= 2000 else 1)
    r = run_reference(steps, threads, data=data, settings=settings, warmup=max(args.warmup, 3), reps=reps)
    out = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": "transitions/s",
           "n_gpus": args.gpus, "steps": steps, "warmup": args.warmup, "ms_per_step": 1e3 * r["seconds"] / r["steps"],
           "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": bench_config(args.workload, world, args.scaling, batch_local, n_ep * REPLAY["ep_len"]),
           "cpu_baseline": cpu_baseline_block(r, single_thread_steps=0),
           "e2e": {"value": r["value"], "unit": "transitions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out))


# ------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------
def flush_l2(torch, dev):
    buf = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
    buf.fill_(1.0)
    torch.cuda.synchronize(dev)
    del buf


def kernel_roofline(L, torch, dev, batch, n_params, window, gs, kernel="k_steps_persistent"):
    """Roofline of the dominant kernel (the persistent step kernel): ONE launch, no sweep inside, L2 flushed before it."""
    n_roof = int(min(512, 999 - (gs % 1000))) if (gs % 1000) < 900 else 64
    n_roof = max(n_roof, 8)
    L.presample(n_roof)
    flush_l2(torch, dev)
    L.train_presampled(0, n_roof)
    L.sync()
    ms_k, _ = L.last_timing()
    peaks, which = measured_peaks()
    sb = step_bytes(batch, n_params, window)
    achieved = sb * n_roof / (ms_k * 1e-3) / 1e9
    traffic, src = ncu_traffic(kernel, n_roof)
    return {"bound": "hbm", "kernel": kernel, "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
            "frac": achieved / peaks["hbm_gbs"], "traffic": traffic, "peak_source": which,
            "traffic_source": (src + " (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum of one persistent launch, scaled to "
                               "this launch's step count)") if src else None,
            "algorithmic_bytes_per_launch": sb * n_roof, "algorithmic_bytes_per_step": sb, "launch_ms": ms_k,
            "steps_per_launch": n_roof, "us_per_step": 1e3 * ms_k / n_roof,
            "note": "the learner step is a chain of dependent phases (latency bound), not bandwidth bound; see roofline_sweeps "
                    "for the HBM-streaming kernels and roofline_batch_sweep for the approach to the bound with the batch size"}


# algorithmic multiply-adds per updated transition of the cfg2 network (~122 kFLOP): forward 32*128 + 128*128 +
# 128*9, input gradient 128*128 + 128*9 (none below the first layer), weight gradient = forward
MACS_PER_TRANSITION_CFG2 = 2 * (32 * 128 + 128 * 128 + 128 * 9) + (128 * 128 + 128 * 9)


def ncu_wide():
    """Per-kernel ncu --set full summary of the wide step at B = 65536 (scripts/r2_wide_profiles.sh -> profiles/r2/ncu_wide.json)."""
    p = os.path.join(ROOT, "profiles", "r2", "ncu_wide.json")
    if not os.path.exists(p):
        return None
    with open(p) as f:
        return json.load(f)


def batch_sweep(torch, dev, data, wl, batches, steps=60):
    """the learner step at growing mini-batches (per GPU), same buffer and network.  B = 256 runs the persistent tile
    kernel (a latency chain) and so does B = 1024, from B = 2048 on the wide step runs (wide_step.cuh: 128-sample tiles, every dense product as 3xTF32
    tcgen05.mma).  Per batch size: us/step, transitions/s, the fraction of the HBM roofline of the ALGORITHMIC bytes, and the
    tensor-side numbers: algorithmic TFLOP/s (2 x 60.8 k multiply-adds per transition), the TF32 TFLOP/s the tensor cores
    execute for it (x3: hi*hi + hi*lo + lo*hi keeps f32 accuracy) and its fraction of the TF32 peak (= half the measured bf16 peak)."""
    from smarties_b200 import Learner
    w = WORKLOADS[wl]
    peaks, _ = measured_peaks()
    tf32_peak = peaks["bf16_tflops"] / 2.0
    names = {0: "two kernels per step", 1: "persistent tile kernel", 2: "cluster kernel", 3: "wide step (tcgen05, 3xTF32)"}
    out = []
    for B in batches:
        L = Learner(32, 8, dict(w["settings"], batchSize=B), seed=42)
        L.load_replay(data)
        L.initialize_learner()
        L.seed_sampler(7)
        L.train_steps(1, want_stats=False)
        n = max(8, min(steps, 4_000_000 // B))
        L.presample(n + 4)
        L.train_presampled(0, 4)
        L.sync()
        flush_l2(torch, dev)
        L.train_presampled(4, n)
        L.sync()
        ms, _ = L.last_timing()
        sb = step_bytes(B, L.n_params, w["window"])
        a = sb * n / (ms * 1e-3) / 1e9
        tps = B * n / (ms * 1e-3)
        e = {"batch": B, "steps": n, "kernel": names.get(L.step_kernel(), "?"), "us_per_step": 1e3 * ms / n, "transitions_per_s": tps,
             "algorithmic_bytes_per_step": sb, "achieved_gbs": a, "frac": a / peaks["hbm_gbs"]}
        if wl == "cfg2":
            alg = 2.0 * MACS_PER_TRANSITION_CFG2 * tps / 1e12
            e["tensor"] = {"algorithmic_tflops": alg, "executed_tf32_tflops": 3.0 * alg if L.step_kernel() == 3 else 0.0,
                           "peak_tf32_tflops": tf32_peak, "frac_of_tf32_peak": (3.0 * alg if L.step_kernel() == 3 else 0.0) / tf32_peak,
                           "peak_source": "MEASURED_PEAKS.json bf16_tflops / 2"}
        out.append(e)
        L.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10000)
    ap.add_argument("--warmup", type=int, default=1000)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--batch", type=int, default=0, help="per-GPU mini-batch override (batch-size studies)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-batch-sweep", action="store_true")
    ap.add_argument("--cpu-steps", type=int, default=0)
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_arm(args)
    if args.warmup < 3:
        args.warmup = 3

    # keep stdout clean for the single JSON line (NCCL prints its version banner to stdout)
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    import torch
    import torch.distributed as dist
    from smarties_b200 import Learner

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        sys.exit("bench.py: no CUDA device — smarties_b200 has no CPU fallback")
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    args.gpus = world

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    w = WORKLOADS[args.workload]
    batch_local, n_ep = plan(args)
    data = make_workload(rank, n_ep)
    n_local = n_ep * REPLAY["ep_len"]
    settings = dict(w["settings"])
    # the reference's multi-learner settings are GLOBAL (HyperParameters.cpp:178-205): batchSize / nLearners samples and
    # maxTotObsNum / nLearners transitions per rank
    settings["batchSize"] = batch_local * world
    cap = 1 << (n_local - 1).bit_length()
    settings["maxTotObsNum"] = cap * world
    settings["minTotObsNum"] = n_local * world
    L = Learner(32, 8, settings, device=local, seed=42 + rank, world_rank=rank, world_size=world)
    if world > 1:
        L.attach_process_group(dist)
    L.load_replay(data)
    L.initialize_learner()
    L.seed_sampler(7 + rank)
    K, W = args.steps, args.warmup
    # the first learner step also applies the FIFO ordering of the episode table
    # (applyEpisodesRemovalAlgo sorts by ID); the resident-ids path needs that steady state
    L.train_steps(1, want_stats=False)

    # ---- value: device time, sampled ids resident in HBM ----
    L.presample(W + K)
    L.train_presampled(0, W)
    L.sync()
    barrier()
    clocks = ClockSampler(local); clocks.start()
    barrier()
    L.train_presampled(W, K)
    L.sync()
    barrier()
    ms, launches = L.last_timing()
    ck = clocks.stop()
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = batch_local * world * K / (ms_max * 1e-3)

    # ---- e2e: host sampler + H2D ids + D2H stats inside the timed region ----
    L.train_steps(W, want_stats=True)
    barrier()
    t0 = time.perf_counter()
    stats = L.train_steps(K, want_stats=True)
    barrier()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e = batch_local * world * K / float(t.item())

    # ---- learner ranks must hold bit-identical weights (rank-ordered sums in the fused exchange) and have seen no peer time-out ----
    ranks_identical = None
    if world > 1:
        L.comm_check()
        digest = np.frombuffer(hashlib.sha256(L.get_weights().tobytes()).digest()[:8], dtype=np.int64).copy()
        mine = torch.from_numpy(digest).to(dev)
        alld = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(alld, mine)
        ranks_identical = all(bool((x == alld[0]).all().item()) for x in alld)
        if not ranks_identical:
            sys.exit("bench.py: learner ranks diverged (weights differ across ranks after the timed region)")

    gs = stats[-1]["grad_step"]
    roof = kernel_roofline(L, torch, dev, batch_local, L.n_params, w["window"], gs)
    if args.workload == "cfg3":        # the LSTM weight gradient is the one tcgen05 contraction of the path: tensor-pipe share from ncu
        p3 = os.path.join(ROOT, "profiles", "r2", "ncu_cfg3.json")
        if os.path.exists(p3):
            with open(p3) as f:
                n3 = json.load(f)
            roof["tensor_pipe_pct_of_peak_ncu"] = n3["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
            roof["traffic"] = n3["dram_bytes"] * roof["steps_per_launch"] / n3["steps_per_launch"]
            roof["traffic_source"] = n3["source"]
    # the HBM-streaming sweeps, timed alone with a flushed L2
    peaks, which = measured_peaks()
    sweeps = {}
    n_tr = L.n_transitions
    # k_sweep_fused is what a learner step runs every 1000 steps (Retrace + aggregates + moments in one pass); the two separate
    # kernels remain for state widths it does not cover and as stand-alone entry points
    for name, fn, bpt in (("k_sweep_fused(retrace+moments)", L.fused_sweep, BYTES_FUSED_SWEEP_PER_TRANSITION),
                          ("k_sweep(retrace)", L.retrace_sweep, BYTES_RETRACE_PER_TRANSITION),
                          ("k_moments", L.reward_state_moments, BYTES_MOMENTS_PER_TRANSITION)):
        best = None
        for _ in range(5):
            flush_l2(torch, dev)
            fn()
            m, _ = L.last_timing()
            best = m if best is None else min(best, m)
        a = bpt * n_tr / (best * 1e-3) / 1e9
        sweeps[name] = {"bound": "hbm", "achieved": a, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": a / peaks["hbm_gbs"],
                        "ms": best, "algorithmic_bytes": bpt * n_tr, "traffic": ncu_traffic(name.split("(")[0])[0]}
    final_stats = {k: stats[-1][k] for k in ("beta", "cmax", "n_far_policy", "grad_step")}
    L.close()

    # ---- the same learner at a LARGE mini-batch (cfg2 only, default runs): 65 536 sampled transitions per GPU and step run the
    #      wide step (tensor cores, wide_step.cuh); device time, ids resident in HBM, max over ranks; weights compared across ranks ----
    large = None
    if args.workload == "cfg2" and not args.batch and args.scaling == "weak" and not args.no_batch_sweep:
        BL, KL, WL = 65536, 30, 3
        s2 = dict(settings, batchSize=BL * world)
        L2 = Learner(32, 8, s2, device=local, seed=42 + rank, world_rank=rank, world_size=world)
        if world > 1:
            L2.attach_process_group(dist)
        L2.load_replay(data)
        L2.initialize_learner()
        L2.seed_sampler(7 + rank)
        L2.train_steps(1, want_stats=False)
        L2.presample(WL + KL)
        L2.train_presampled(0, WL)
        L2.sync()
        barrier()
        L2.train_presampled(WL, KL)
        L2.sync()
        barrier()
        ms2, _ = L2.last_timing()
        t = torch.tensor([ms2], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms2 = float(t.item())
        same = None
        if world > 1:
            L2.comm_check()
            dg = torch.from_numpy(np.frombuffer(hashlib.sha256(L2.get_weights().tobytes()).digest()[:8], dtype=np.int64).copy()).to(dev)
            alld = [torch.zeros_like(dg) for _ in range(world)]
            dist.all_gather(alld, dg)
            same = all(bool((x == alld[0]).all().item()) for x in alld)
        names = {0: "two kernels per step", 1: "persistent tile kernel", 2: "cluster kernel", 3: "wide step (tcgen05, 3xTF32)"}
        tps = BL * world * KL / (ms2 * 1e-3)
        large = {"batch_per_gpu": BL, "global_batch": BL * world, "steps": KL, "warmup": WL, "kernel": names.get(L2.step_kernel(), "?"),
                 "us_per_step": 1e3 * ms2 / KL, "transitions_per_s": tps, "ranks_identical": same,
                 "executed_tf32_tflops": 3.0 * 2.0 * MACS_PER_TRANSITION_CFG2 * tps / 1e12,
                 "note": "device time with sampled ids resident in HBM (the host sampler of the e2e path costs ~30 ns per index)"}
        L2.close()

    if rank == 0:
        out = {"metric": METRIC, "value": value, "unit": "transitions/s", "n_gpus": world,
               "steps": K, "warmup": W, "ms_per_step": ms_max / K, "higher_is_better": True, "scaling": args.scaling,
               "vs_baseline": None, "dtype": "f32", "data": "synthetic",
               "config": bench_config(args.workload, world, args.scaling, batch_local, n_local),
               "mode": os.environ.get("SMB200_MODE", "persistent"),
               "clocks": ck,
               "e2e": {"value": e2e, "unit": "transitions/s", "h2d_bytes_per_step": 2 * 4 * batch_local, "d2h_bytes_per_step": 128},
               "gpu_launches": int(launches),
               "roofline": roof, "roofline_sweeps": sweeps, "final_stats": final_stats}
        if ranks_identical is not None:
            out["ranks_identical"] = ranks_identical
        if large is not None:
            out["large_batch"] = large
        if world == 1 and args.workload == "cfg2" and not args.batch and not args.no_batch_sweep:
            out["roofline_batch_sweep"] = batch_sweep(torch, dev, data, args.workload, (256, 1024, 2048, 4096, 16384, 65536))
            nw = ncu_wide()
            if nw:      # per-kernel ncu evidence of the wide step (tensor-pipe share, DRAM bytes), captured at B = 65536
                out["roofline_wide_kernels"] = nw
        if world == 1 and not args.no_cpu_baseline:
            cpu_steps = args.cpu_steps or (6000 if args.workload == "cfg2" else 150)
            rs = dict(w["settings"], batchSize=batch_local)
            r = run_reference(cpu_steps, os.cpu_count() or 1, data=data, reps=3, settings=rs, warmup=20)
            out["cpu_baseline"] = cpu_baseline_block(r, data=data, settings=rs, single_thread_steps=400 if args.workload == "cfg2" else 20)
        os.write(real_stdout, (json.dumps(out) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
