"""ctypes binding of the C-ABI (include/smarties_b200.h) + a thin host-side mirror of the
reference's learner interface for this path.

Method names follow the reference: `push_episode` (MemoryBuffer::pushBackEpisode),
`initialize_learner` (Learner::initializeLearner), `sample_minibatch`
(MemoryBuffer::sampleMinibatch), `train_steps` (= n x {spawnTrainTasks; processMemoryBuffer;
applyGradient; globalGradCounterUpdate}), `get_metrics` (MemoryBuffer::getMetrics).

There is NO CPU fallback: if the CUDA library is missing or no GPU is present, construction
raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# SMB200_PROFILE=1 selects the flavour with phase timestamps compiled in (scripts/phase_report.py)
LIB_PATH = os.environ.get("SMB200_LIB") or os.path.join(   # SMB200_LIB: another build of the same library (A/B experiments)
    _HERE, "libsmarties_b200_prof.so" if os.environ.get("SMB200_PROFILE") else "libsmarties_b200.so")

MAX_HIDDEN = 8
MAX_ACTION = 64

# every symbol include/smarties_b200.h declares
EXPORTS = [
    "smb200_default_config", "smb200_create", "smb200_destroy", "smb200_last_error", "smb200_n_params",
    "smb200_n_outputs", "smb200_set_weights", "smb200_get_weights", "smb200_get_target_weights", "smb200_set_target_weights", "smb200_set_adam", "smb200_get_adam",
    "smb200_get_grad", "smb200_set_scaling", "smb200_get_scaling", "smb200_push_episode", "smb200_n_transitions",
    "smb200_n_episodes", "smb200_initialize_learner", "smb200_set_grad_step", "smb200_seed_sampler", "smb200_sample",
    "smb200_train_steps", "smb200_train_steps_weights", "smb200_pin_host_buffer", "smb200_train_step_on", "smb200_get_last_batch", "smb200_retrace_sweep", "smb200_fused_sweep",
    "smb200_reward_state_moments", "smb200_read_field", "smb200_read_episodes", "smb200_n_rows", "smb200_get_stats",
    "smb200_forward", "smb200_forward_seq", "smb200_last_timing", "smb200_step_kernel", "smb200_presample", "smb200_train_presampled", "smb200_sync", "smb200_profile_phases",
    "smb200_comm_init", "smb200_comm_attach", "smb200_comm_error", "smb200_write_field", "smb200_save", "smb200_restart",
    "smb200_push_episode_restored", "smb200_set_refer", "smb200_set_grad_stats", "smb200_uint_plus_float", "smb200_host_far_chain", "smb200_host_replay_trace", "smb200_host_init_weights",
    "smb200_host_strip_weights", "smb200_host_write_grad_stats", "smb200_host_repack_episodes",
    "smb200_host_adam", "smb200_host_value_scaling", "smb200_host_return_estimator",
    "smb200_host_discrete_loss", "smb200_host_wide_plan",
]

FIELDS = dict(V=0, ADV=1, QRET=2, DELTA=3, RHO=4, KL=5, REWARD=6)


class Config(C.Structure):
    _fields_ = [
        ("device", C.c_int32), ("algo", C.c_int32), ("dim_state", C.c_int32), ("dim_action", C.c_int32),
        ("action_bounded", C.c_uint8 * MAX_ACTION), ("n_hidden", C.c_int32), ("hidden", C.c_int32 * MAX_HIDDEN),
        ("batch_size", C.c_int32), ("batch_size_global", C.c_int32), ("max_tot_obs", C.c_int64),
        ("max_tot_obs_global", C.c_int64), ("capacity_rows", C.c_int64), ("max_episodes", C.c_int32),
        ("gamma", C.c_double), ("lambda_", C.c_double), ("clip_imp_weight", C.c_double), ("penal_tol", C.c_double),
        ("eps_anneal", C.c_double), ("learnrate", C.c_double), ("nn_lambda", C.c_double), ("expl_noise", C.c_double),
        ("out_weights_prefac", C.c_double), ("refer_reduce_threads", C.c_int32), ("world_rank", C.c_int32),
        ("world_size", C.c_int32), ("seed", C.c_uint64), ("nn_type", C.c_int32), ("nn_bptt_seq", C.c_int32),
        ("min_tot_obs", C.c_int64), ("returns_estimator", C.c_int32), ("discrete_options", C.c_int32),
        ("data_sampling", C.c_int32), ("er_filter", C.c_int32), ("nn_func", C.c_int32), ("target_delay", C.c_double),
    ]


class StepStats(C.Structure):
    _fields_ = [
        ("beta", C.c_double), ("cmax", C.c_double), ("cinv", C.c_double), ("n_far_policy", C.c_int64),
        ("n_far_exact", C.c_int64), ("avg_kl", C.c_double), ("avg_sq_err", C.c_double), ("max_abs_err", C.c_double),
        ("avg_return", C.c_double), ("stdev_q", C.c_double), ("avg_q", C.c_double), ("max_q", C.c_double),
        ("min_q", C.c_double), ("sum_ret_err", C.c_double), ("cnt_ret", C.c_int64), ("grad_step", C.c_int64),
    ]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class StepStatsList:
    """Per-step statistics of one smb200_train_steps call: behaves like a list of dicts, converted on access (the C array
    is what the call filled; building 10^4 Python dicts is not part of a learner step)."""

    def __init__(self, arr):
        self._a = arr

    def __len__(self):
        return len(self._a)

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self._a[j].as_dict() for j in range(*i.indices(len(self._a)))]
        return self._a[i].as_dict()

    def __iter__(self):
        return (s.as_dict() for s in self._a)

    def __eq__(self, other):
        return list(self) == list(other)


_lib = None


def load_library(path: str = LIB_PATH):
    """dlopen the C-ABI library and declare prototypes.  Raises if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(path):
        raise RuntimeError(f"{path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(smarties_b200 has no CPU fallback)")
    lib = C.CDLL(path)
    P = C.POINTER
    H = C.c_void_p
    fp, ip, dp = P(C.c_float), P(C.c_int64), P(C.c_double)
    sig = {
        "smb200_default_config": (C.c_int, [P(Config), C.c_int32, C.c_int32]),
        "smb200_create": (C.c_int, [P(Config), P(H)]),
        "smb200_destroy": (None, [H]),
        "smb200_last_error": (C.c_char_p, []),
        "smb200_n_params": (C.c_int64, [H]), "smb200_n_outputs": (C.c_int32, [H]),
        "smb200_set_weights": (C.c_int, [H, fp, C.c_int64]), "smb200_get_weights": (C.c_int, [H, fp, C.c_int64]),
        "smb200_set_adam": (C.c_int, [H, fp, fp, C.c_int64, C.c_int64]), "smb200_get_adam": (C.c_int, [H, fp, fp, C.c_int64]),
        "smb200_get_grad": (C.c_int, [H, fp, C.c_int64]),
        "smb200_get_target_weights": (C.c_int, [H, fp, C.c_int64]), "smb200_set_target_weights": (C.c_int, [H, fp, C.c_int64]),
        "smb200_set_scaling": (C.c_int, [H, fp, fp, fp, fp]), "smb200_get_scaling": (C.c_int, [H, fp, fp, fp, fp]),
        "smb200_push_episode": (C.c_int, [H, C.c_int64, C.c_int32, C.c_int32, fp, fp, fp, fp, fp, fp]),
        "smb200_n_transitions": (C.c_int64, [H]), "smb200_n_episodes": (C.c_int64, [H]), "smb200_n_rows": (C.c_int64, [H]),
        "smb200_initialize_learner": (C.c_int, [H]), "smb200_set_grad_step": (C.c_int, [H, C.c_int64]),
        "smb200_seed_sampler": (C.c_int, [H, C.c_uint64]), "smb200_sample": (C.c_int, [H, ip, ip]),
        "smb200_train_steps": (C.c_int, [H, C.c_int32, P(StepStats)]),
        "smb200_train_steps_weights": (C.c_int, [H, C.c_int32, P(StepStats), fp, C.c_int64]),
        "smb200_pin_host_buffer": (C.c_int, [C.c_void_p, C.c_int64]),
        "smb200_train_step_on": (C.c_int, [H, ip, ip, C.c_int32, P(StepStats)]),
        "smb200_get_last_batch": (C.c_int, [H, fp, fp, fp]),
        "smb200_retrace_sweep": (C.c_int, [H, dp]), "smb200_reward_state_moments": (C.c_int, [H, dp]),
        "smb200_fused_sweep": (C.c_int, [H, dp, dp]),
        "smb200_read_field": (C.c_int, [H, C.c_int32, fp, C.c_int64]),
        "smb200_read_episodes": (C.c_int, [H, ip, ip, fp, C.c_int64]),
        "smb200_write_field": (C.c_int, [H, C.c_int32, fp, C.c_int64]),
        "smb200_save": (C.c_int, [H, C.c_char_p]), "smb200_restart": (C.c_int, [H, C.c_char_p]),
        "smb200_push_episode_restored": (C.c_int, [H, C.c_int64, C.c_int32, C.c_int32] + [fp] * 10 + [C.c_double]),
        "smb200_set_refer": (C.c_int, [H, C.c_double, C.c_double]),
        "smb200_set_grad_stats": (C.c_int, [H, C.c_char_p]),
        "smb200_get_stats": (C.c_int, [H, P(StepStats)]),
        "smb200_forward": (C.c_int, [H, fp, C.c_int32, fp]),
        "smb200_forward_seq": (C.c_int, [H, fp, C.POINTER(C.c_int32), C.c_int32, C.c_int32, fp]),
        "smb200_last_timing": (C.c_int, [H, dp, ip]), "smb200_step_kernel": (C.c_int, [H]),
        "smb200_presample": (C.c_int, [H, C.c_int32]), "smb200_train_presampled": (C.c_int, [H, C.c_int32, C.c_int32]),
        "smb200_sync": (C.c_int, [H]),
        "smb200_profile_phases": (C.c_int, [H, C.c_int32, ip, C.c_int64, P(C.c_int32)]),
        "smb200_comm_init": (C.c_int, [H, C.c_int32, C.c_int32, P(C.c_uint8), C.c_int32]),
        "smb200_comm_attach": (C.c_int, [H, P(C.c_uint8), C.c_int32]),
        "smb200_comm_error": (C.c_int, [H]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int64))


class SmartiesB200Error(RuntimeError):
    pass


def make_config(dim_state: int, dim_action: int, settings=None, *, device: int = 0, bounded=None, seed: int = 42,
                capacity_rows: int = 0, max_episodes: int = 0, refer_reduce_threads: int = 32, world_rank: int = 0,
                world_size: int = 1, discrete_options: int = 0):
    """settings.json surface -> smb200_config (what integration/RACER_B200.cpp fills from the reference's HyperParameters).
    Returns (Config, HyperParameters).  Host only: usable without a GPU."""
    from .settings import HyperParameters

    lib = load_library()
    hp = settings if isinstance(settings, HyperParameters) else HyperParameters(dim_state, dim_action, settings or {})
    hp.define_distributed_learning(world_size)
    cfg = Config()
    if lib.smb200_default_config(C.byref(cfg), dim_state, dim_action) != 0:
        raise SmartiesB200Error("smb200_default_config: " + lib.smb200_last_error().decode())
    cfg.device = device
    cfg.algo = {"VRACER": 0, "RACER": 1}[hp.learner]
    # RACER::setupNet (Learners/RACER_common.cpp:82-91): the "encoder" Approximator IS the one network — createEncoder builds the
    # encoderLayerSizes layers (Approximator::buildPreprocessing), buildFromSettings continues with nnLayerSizes in the same Builder
    hidden = [int(h) for h in hp.encoderLayerSizes if int(h) > 0] + [int(h) for h in hp.nnLayerSizes if int(h) > 0]
    cfg.n_hidden = len(hidden)
    for i, h in enumerate(hidden):
        cfg.hidden[i] = h
    cfg.batch_size, cfg.batch_size_global = hp.batchSize_local, hp.batchSize
    cfg.max_tot_obs, cfg.max_tot_obs_global = hp.maxTotObsNum_local, hp.maxTotObsNum
    cfg.capacity_rows, cfg.max_episodes = capacity_rows, max_episodes
    cfg.gamma, cfg.lambda_, cfg.clip_imp_weight, cfg.penal_tol = hp.gamma, hp.lambda_, hp.clipImpWeight, hp.penalTol
    cfg.eps_anneal, cfg.learnrate, cfg.nn_lambda = hp.epsAnneal, hp.learnrate, hp.nnLambda
    cfg.expl_noise, cfg.out_weights_prefac = hp.explNoise, hp.outWeightsPrefac
    cfg.refer_reduce_threads = refer_reduce_threads
    cfg.world_rank, cfg.world_size, cfg.seed = world_rank, world_size, seed
    cfg.nn_type, cfg.nn_bptt_seq = {"FFNN": 0, "LSTM": 1, "MGU": 2, "GRU": 2}[hp.nnType], int(hp.nnBPTTseq)
    cfg.min_tot_obs = hp.minTotObsNum_local
    cfg.returns_estimator = {"retrace": 0, "GAE": 1, "retraceExplore": 2}[hp.returnsEstimator]
    cfg.discrete_options = int(discrete_options)     # from the MDP (Communicator::setNumberOfOptions), not from settings.json
    cfg.data_sampling = {"uniform": 0, "PERrank": 1, "PERerr": 2, "PERseq": 3}[hp.dataSamplingAlgo]
    cfg.er_filter = {"oldest": 0, "default": 0, "farpolfrac": 1, "maxkldiv": 2, "minerror": 3}[hp.ERoldSeqFilter]
    cfg.nn_func = {"Tanh": 0, "SoftSign": 1, "HardSign": 2, "Sigm": 3, "Relu": 4, "LRelu": 5, "ExpPlus": 6, "SoftPlus": 7, "Exp": 8,
                   "Linear": 9}[hp.nnFunc]
    cfg.target_delay = float(hp.targetDelay)
    if bounded is not None:
        b = np.broadcast_to(np.asarray(bounded, dtype=bool), (dim_action,))
        for i in range(dim_action):
            cfg.action_bounded[i] = int(b[i])
    return cfg, hp


class Learner:
    """V-RACER learner on one B200 (the hot path of smarties::RACER<Zero_advantage,...>)."""

    def __init__(self, dim_state: int, dim_action: int, settings: dict | None = None, *, device: int = 0,
                 bounded=None, seed: int = 42, capacity_rows: int = 0, max_episodes: int = 0,
                 refer_reduce_threads: int = 32, world_rank: int = 0, world_size: int = 1, discrete_options: int = 0):
        self.lib = load_library()
        cfg, hp = make_config(dim_state, dim_action, settings, device=device, bounded=bounded, seed=seed,
                              capacity_rows=capacity_rows, max_episodes=max_episodes,
                              refer_reduce_threads=refer_reduce_threads, world_rank=world_rank, world_size=world_size,
                              discrete_options=discrete_options)
        self.hp = hp
        self.cfg = cfg
        self.dS, self.dA, self.B = dim_state, dim_action, hp.batchSize_local
        h = C.c_void_p()
        self._check(self.lib.smb200_create(C.byref(cfg), C.byref(h)))
        self.h = h
        self.n_params = int(self.lib.smb200_n_params(h))
        self.n_out = int(self.lib.smb200_n_outputs(h))

    # -- plumbing --
    def _check(self, rc):
        if rc != 0:
            raise SmartiesB200Error(f"smarties_b200 error {rc}: {self.lib.smb200_last_error().decode()}")

    def close(self):
        if getattr(self, "h", None):
            self.lib.smb200_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- learner ranks (one process per GPU) --
    def attach_process_group(self, dist, broadcast_weights=True):
        """Join the learner ranks of `dist` (torch.distributed, already initialised): exchange the CUDA
        IPC handles of the peer-memory blocks, and copy rank 0's initial weights to every rank
        (Parameters::broadcast, Network/Layers/Parameters.h:43-46)."""
        from .distributed import exchange_bytes, broadcast_array
        world, rank = dist.get_world_size(), dist.get_rank()
        HB = 64
        mine = (C.c_uint8 * HB)()
        self._check(self.lib.smb200_comm_init(self.h, world, rank, mine, HB))
        handles = exchange_bytes(dist, bytes(mine))
        blob = (C.c_uint8 * (HB * world)).from_buffer_copy(b"".join(handles))
        self._check(self.lib.smb200_comm_attach(self.h, blob, HB))
        if broadcast_weights:
            self.set_weights(broadcast_array(dist, self.get_weights(), src=0))
        dist.barrier()

    def comm_check(self):
        self._check(self.lib.smb200_comm_error(self.h))

    # -- replay memory --
    def push_episode(self, eid, states, actions, policies, rewards, terminated, value=None, advantage=None):
        S, A, MU, R = _f32(states), _f32(actions), _f32(policies), _f32(rewards)
        N = S.shape[0]
        V = _f32(value) if value is not None else None
        ADV = _f32(advantage) if advantage is not None else None
        self._check(self.lib.smb200_push_episode(self.h, int(eid), N, int(bool(terminated)), _fp(S), _fp(A), _fp(MU), _fp(R),
                                                 _fp(V) if V is not None else None, _fp(ADV) if ADV is not None else None))

    def load_replay(self, d):
        """Push a smarties_b200.synth buffer episode by episode (IDs = insertion order)."""
        for e in range(len(d["N"])):
            o, N = int(d["start"][e]), int(d["N"][e])
            self.push_episode(e, d["S"][o:o + N], d["A"][o:o + N], d["MU"][o:o + N], d["R"][o:o + N], d["term"][e])

    @property
    def n_transitions(self):
        return int(self.lib.smb200_n_transitions(self.h))

    @property
    def n_episodes(self):
        return int(self.lib.smb200_n_episodes(self.h))

    def read_field(self, name):
        n = int(self.lib.smb200_n_rows(self.h))
        out = np.empty(n, np.float32)
        self._check(self.lib.smb200_read_field(self.h, FIELDS[name], _fp(out), n))
        return out

    def write_field(self, name, values):
        v = _f32(values)
        self._check(self.lib.smb200_write_field(self.h, FIELDS[name], _fp(v), v.size))

    def save(self, base):
        """Learner_approximator::save(): the reference's checkpoint files `<base>_net_weights.raw`, ... (base = dir/agent_00)."""
        self._check(self.lib.smb200_save(self.h, os.fsencode(base)))

    def restart(self, base):
        """Learner_approximator::restart() from files written by the reference or by save()."""
        self._check(self.lib.smb200_restart(self.h, os.fsencode(base)))

    def read_episodes(self):
        n = self.n_episodes
        ids, rows, agg = np.empty(n, np.int64), np.empty(n, np.int64), np.empty((n, 9), np.float32)
        self._check(self.lib.smb200_read_episodes(self.h, _ip(ids), _ip(rows), _fp(agg), n))
        return ids, rows, agg

    # -- learner --
    def initialize_learner(self):
        self._check(self.lib.smb200_initialize_learner(self.h))

    def set_grad_step(self, n):
        self._check(self.lib.smb200_set_grad_step(self.h, int(n)))

    def seed_sampler(self, seed):
        self._check(self.lib.smb200_seed_sampler(self.h, int(seed)))

    def set_grad_stats(self, base):
        """StatsTracker file of the reference (Utils/StatsTracker.cpp:66-89): `<base>_outGrad_stats.raw`, one row of
        per-output gradient mean / rms for every step that starts at nGradSteps % 1000 == 0.  None switches it off."""
        self._check(self.lib.smb200_set_grad_stats(self.h, os.fsencode(base) if base else None))

    def sample_minibatch(self):
        pos, t = np.empty(self.B, np.int64), np.empty(self.B, np.int64)
        self._check(self.lib.smb200_sample(self.h, _ip(pos), _ip(t)))
        return pos, t

    def train_steps(self, n, want_stats=True):
        st = (StepStats * n)() if want_stats else None
        self._check(self.lib.smb200_train_steps(self.h, int(n), st))
        return StepStatsList(st) if want_stats else None

    def train_steps_weights(self, n):
        """train_steps + the weights after the last step from the same call (what the binding hands to the host actors)."""
        st = (StepStats * n)()
        w = np.empty(self.n_params, np.float32)
        self._check(self.lib.smb200_train_steps_weights(self.h, int(n), st, _fp(w), self.n_params))
        return StepStatsList(st), w

    def train_step_on(self, pos, t):
        pos, t = np.ascontiguousarray(pos, np.int64), np.ascontiguousarray(t, np.int64)
        st = StepStats()
        self._check(self.lib.smb200_train_step_on(self.h, _ip(pos), _ip(t), len(pos), C.byref(st)))
        return st.as_dict()

    def presample(self, n):
        self._check(self.lib.smb200_presample(self.h, int(n)))

    def train_presampled(self, first, n):
        self._check(self.lib.smb200_train_presampled(self.h, int(first), int(n)))

    def profile_phases(self, n):
        """(cycles[n][grid][8], device ms) of one persistent launch over n presampled steps."""
        cap = n * 148 * 48
        out = np.zeros(cap, np.int64)
        grid = C.c_int32()
        self._check(self.lib.smb200_profile_phases(self.h, int(n), _ip(out), cap, C.byref(grid)))
        g = grid.value
        return out[:n * g * 48].reshape(n, g, 48), self.last_timing()[0]

    def step_kernel(self):
        """0 two kernels per step, 1 persistent tile kernel, 2 cluster kernel, 3 wide step (tensor cores)"""
        return int(self.lib.smb200_step_kernel(self.h))

    def wide_step_active(self):
        return self.step_kernel() == 3

    def sync(self):
        self._check(self.lib.smb200_sync(self.h))

    def last_timing(self):
        ms, nl = C.c_double(), C.c_int64()
        self._check(self.lib.smb200_last_timing(self.h, C.byref(ms), C.byref(nl)))
        return ms.value, nl.value

    def get_stats(self):
        st = StepStats()
        self._check(self.lib.smb200_get_stats(self.h, C.byref(st)))
        return st.as_dict()

    def get_last_batch(self):
        O = np.empty((self.B, self.n_out), np.float32)
        g = np.empty((self.B, self.n_out), np.float32)
        X = np.empty((self.B, self.dS), np.float32)
        self._check(self.lib.smb200_get_last_batch(self.h, _fp(O), _fp(g), _fp(X)))
        return O, g, X

    # -- network --
    def get_target_weights(self):
        """AdamOptimizer::target_weights ("targetDelay"; RACER never evaluates them, they only reach the checkpoint)."""
        w = np.empty(self.n_params, np.float32)
        self._check(self.lib.smb200_get_target_weights(self.h, _fp(w), self.n_params))
        return w

    def get_weights(self):
        w = np.empty(self.n_params, np.float32)
        self._check(self.lib.smb200_get_weights(self.h, _fp(w), self.n_params))
        return w

    def set_weights(self, w):
        w = _f32(w)
        self._check(self.lib.smb200_set_weights(self.h, _fp(w), w.size))

    def get_grad(self):
        g = np.empty(self.n_params, np.float32)
        self._check(self.lib.smb200_get_grad(self.h, _fp(g), self.n_params))
        return g

    def get_adam(self):
        m1, m2 = np.empty(self.n_params, np.float32), np.empty(self.n_params, np.float32)
        self._check(self.lib.smb200_get_adam(self.h, _fp(m1), _fp(m2), self.n_params))
        return m1, m2

    def get_scaling(self):
        mean, scale, std = (np.empty(self.dS, np.float32) for _ in range(3))
        rew = np.empty(3, np.float32)
        self._check(self.lib.smb200_get_scaling(self.h, _fp(mean), _fp(scale), _fp(std), _fp(rew)))
        return mean, scale, std, rew

    def set_scaling(self, mean, scale, std, rew):
        mean, scale, std, rew = _f32(mean), _f32(scale), _f32(std), _f32(rew)
        self._check(self.lib.smb200_set_scaling(self.h, _fp(mean), _fp(scale), _fp(std), _fp(rew)))

    def forward(self, states):
        s = _f32(states).reshape(-1, self.dS)
        out = np.empty((s.shape[0], self.n_out), np.float32)
        self._check(self.lib.smb200_forward(self.h, _fp(s), s.shape[0], _fp(out)))
        return out

    def forward_seq(self, windows, lengths):
        """Policy evaluation of n agents on their windows: windows[n][max_len][dS] raw states (oldest first), the first
        lengths[i] rows of agent i are valid; outputs at the newest row (recurrent nets: zero initial state)."""
        w = _f32(windows)
        n, max_len = w.shape[0], w.shape[1]
        w = w.reshape(n, max_len, self.dS)
        ln = np.ascontiguousarray(np.asarray(lengths, np.int32).reshape(n))
        out = np.empty((n, self.n_out), np.float32)
        self._check(self.lib.smb200_forward_seq(self.h, _fp(w), ln.ctypes.data_as(C.POINTER(C.c_int32)), n, max_len, _fp(out)))
        return out

    # -- stand-alone sweeps --
    def retrace_sweep(self):
        e = C.c_double()
        self._check(self.lib.smb200_retrace_sweep(self.h, C.byref(e)))
        return e.value

    def fused_sweep(self):
        """(sum of squared Retrace changes, moments[2*dS+3]) of the one-pass sweep kernel (Retrace + aggregates + moments)."""
        e = C.c_double()
        out = np.empty(2 * self.dS + 3, np.float64)
        self._check(self.lib.smb200_fused_sweep(self.h, C.byref(e), out.ctypes.data_as(C.POINTER(C.c_double))))
        return e.value, out

    def reward_state_moments(self):
        out = np.empty(2 * self.dS + 3, np.float64)
        self._check(self.lib.smb200_reward_state_moments(self.h, out.ctypes.data_as(C.POINTER(C.c_double))))
        return out
