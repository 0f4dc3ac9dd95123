"""CPU restatement (numpy) of the reference's V-RACER / RACER learner hot path.

TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may
import this module; the product (smarties_b200/) never does.  (scripts/parity_report.py — the
measured-error table of DESIGN.md — and scripts/dropin_run.py — imported by tests/test_gpu_dropin.py —
are test tools of the same kind: they use it, and the binaries under oracle/_ref/, as the checker.)

Parity status: PINNED against outputs of the reference itself — golden vectors produced by
oracle/_ref/ref_harness (the unmodified reference compiled from /root/reference by
oracle/Makefile) and committed under tests/golden/ (generator: tests/golden/make_golden.py).
The reference ships no golden vectors of its own for this path.

The reference is compiled with -O3 -ffast-math (CMakeLists.txt:86-89), its per-sample weight
gradient accumulation is thread-partitioned, and libm/libmvec exp differ from numpy's, so
float results agree with the reference to f32 round-off (tests state the tolerances), while
sampled indices and — for identical network outputs — far-policy flags/counts are bit-exact.

Every function cites the reference file:line it follows (paths relative to
/root/reference/source/smarties/).
"""
from __future__ import annotations

import numpy as np

f32 = np.float32
f64 = np.float64
FLT_EPS = float(np.finfo(np.float32).eps)
FLT_MIN = float(np.finfo(np.float32).tiny)
SQUASH_MAX = 8.31776613503286  # Math/Continuous_policy.h:218


def round_up8(n: int) -> int:
    """Utils/FunctionUtilities.h:74-83 (VEC_WIDTH 32 B / 4 B floats = 8)."""
    return int(-(-n // 8) * 8)


# ------------------------------------------------------------------------------------------
# libstdc++ std::mt19937 + std::uniform_int_distribution<size_t>  (ReplayMemory/Sampling.cpp:82-96)
# ------------------------------------------------------------------------------------------
class Mt19937:
    """std::mt19937(seed): numpy's MT19937 with legacy (init_genrand) seeding is the same
    generator; raw 32-bit draws are consumed one at a time like libstdc++ does."""

    def __init__(self, seed: int):
        self.bg = np.random.MT19937()
        self.bg._legacy_seeding(int(seed))
        self._buf = np.empty(0, dtype=np.uint64)
        self._pos = 0

    def __call__(self) -> int:
        if self._pos >= len(self._buf):
            self._buf = self.bg.random_raw(4096)
            self._pos = 0
        v = int(self._buf[self._pos])
        self._pos += 1
        return v


def uniform_int(gen: Mt19937, n: int) -> int:
    """uniform_int_distribution<size_t>(0, n-1)(gen) as implemented by libstdc++ 13
    (bits/uniform_int_dist.h: 32-bit URNG, range < 2^32 -> Lemire's method _S_nd<uint64>)."""
    assert 0 < n <= 0xFFFFFFFF
    product = gen() * n
    low = product & 0xFFFFFFFF
    if low < n:
        threshold = ((1 << 32) - n) % n
        while low < threshold:
            product = gen() * n
            low = product & 0xFFFFFFFF
    return product >> 32


def sample_uniform(gen: Mt19937, n_transitions: int, batch: int) -> np.ndarray:
    """Sample_uniform::sample, non-episodic branch (ReplayMemory/Sampling.cpp:82-93):
    draw, sort, unique, redraw the tail until `batch` unique ascending ids remain."""
    ret: list[int] = []
    while len(ret) < batch:
        ret += [uniform_int(gen, n_transitions) for _ in range(batch - len(ret))]
        ret = sorted(set(ret))
    return np.asarray(ret, dtype=np.int64)


def id_to_seq_step(ids: np.ndarray, ndata: np.ndarray):
    """Sampling::IDtoSeqStep (ReplayMemory/Sampling.cpp:26-47): prefix walk over episodes in
    the buffer's CURRENT vector order; returns (episode position, time step)."""
    prefix = np.concatenate([[0], np.cumsum(ndata)])
    seq = np.searchsorted(prefix, ids, side="right") - 1
    obs = ids - prefix[seq]
    return seq.astype(np.int64), obs.astype(np.int64)


def std_sort(v: list, comp, depth_limit=None) -> None:
    """libstdc++ std::sort (bits/stl_algo.h: __introsort_loop with median-of-three pivots down to runs of 16, heap sort
    past depth 2*floor(log2 n), then __final_insertion_sort), in place.  It is not a stable sort: where the comparator
    ties (episodes with equal far-policy fraction, say) the resulting order is a property of this exact algorithm, and the
    order of the episode vector decides which episode a sampled transition id belongs to (Sampling.cpp:26-47) and which
    episode is pruned (MemoryProcessing.cpp:327-351).  `depth_limit` is for the tests only (0 = heap-sort branch at once)."""
    def unguarded_linear_insert(last):
        val = v[last]; nxt = last - 1
        while comp(val, v[nxt]):
            v[last] = v[nxt]; last = nxt; nxt -= 1
        v[last] = val

    def insertion_sort(first, last):
        for i in range(first + 1, last):
            if comp(v[i], v[first]):
                val = v[i]; v[first + 1:i + 1] = v[first:i]; v[first] = val
            else:
                unguarded_linear_insert(i)

    def heap_sort(first, last):                      # std::__partial_sort(first, last, last): make_heap + sort_heap
        def adjust(hole, length, val):               # std::__adjust_heap + __push_heap
            top = hole; child = hole
            while child < (length - 1) // 2:
                child = 2 * (child + 1)
                if comp(v[first + child], v[first + child - 1]):
                    child -= 1
                v[first + hole] = v[first + child]; hole = child
            if (length & 1) == 0 and child == (length - 2) // 2:
                child = 2 * (child + 1)
                v[first + hole] = v[first + child - 1]; hole = child - 1
            parent = (hole - 1) // 2
            while hole > top and comp(v[first + parent], val):
                v[first + hole] = v[first + parent]; hole = parent; parent = (hole - 1) // 2
            v[first + hole] = val
        n = last - first
        if n >= 2:
            parent = (n - 2) // 2
            while True:
                adjust(parent, n, v[first + parent])
                if parent == 0:
                    break
                parent -= 1
        while last - first > 1:
            last -= 1
            val = v[last]; v[last] = v[first]
            adjust(0, last - first, val)

    def introsort_loop(first, last, depth):
        while last - first > 16:
            if depth == 0:
                heap_sort(first, last)
                return
            depth -= 1
            mid = first + (last - first) // 2
            a, b, c = first + 1, mid, last - 1       # __move_median_to_first(first, first + 1, mid, last - 1)
            if comp(v[a], v[b]):
                m = b if comp(v[b], v[c]) else (c if comp(v[a], v[c]) else a)
            else:
                m = a if comp(v[a], v[c]) else (c if comp(v[b], v[c]) else b)
            v[first], v[m] = v[m], v[first]
            lo, hi = first + 1, last                  # __unguarded_partition(first + 1, last, first)
            while True:
                while comp(v[lo], v[first]):
                    lo += 1
                hi -= 1
                while comp(v[first], v[hi]):
                    hi -= 1
                if not lo < hi:
                    break
                v[lo], v[hi] = v[hi], v[lo]
                lo += 1
            introsort_loop(lo, last, depth)
            last = lo

    n = len(v)
    if n < 2:
        return
    introsort_loop(0, n, 2 * (n.bit_length() - 1) if depth_limit is None else depth_limit)   # std::__lg(n) * 2
    if n > 16:                                        # __final_insertion_sort
        insertion_sort(0, 16)
        for i in range(16, n):
            unguarded_linear_insert(i)
    else:
        insertion_sort(0, n)


def uint_plus_float(n: int, x) -> int:
    """`Uint n; n += x;` with float x as gcc compiles it for x86-64 (MemoryProcessing.cpp:202-227, nOffPol): n is rounded to
    float, added, and converted back with cvttss2si — directly below 2^63, so a negative sum wraps to 2^64 - |sum|, and as
    cvttss2si(f - 2^63) ^ 2^63 from there on, so anything >= 2^64 (such as the float nearest to a wrapped value) becomes 0.
    A negative sum does occur in the reference: with clipImpWeight < 1 (one action component) the initial CinvRet = 1/C
    exceeds 1 (MemoryBuffer.h:41-44), every stored importance weight 1 counts as "was far", and the per-episode far-policy
    fractions go negative until the first every-1000-steps recompute."""
    two63 = f32(9223372036854775808.0)
    f = f32(f32(n) + f32(x))
    indefinite = 1 << 63
    if f < two63:
        t = int(np.trunc(f)) if f > -two63 else -indefinite
        return t & 0xFFFFFFFFFFFFFFFF
    gq = f32(f - two63)
    t = int(np.trunc(gq)) if gq < two63 else indefinite
    return (t ^ indefinite) & 0xFFFFFFFFFFFFFFFF


class DiscreteDistribution:
    """libstdc++ std::discrete_distribution<Uint> built from float weights (bits/random.tcc, param_type::_M_initialize and
    operator()): weights widened to double, divided by their sequential sum, cumulated sequentially, last entry forced to
    1; a draw is generate_canonical<double, 53> — two 32-bit outputs of the generator, low word first — located by
    std::lower_bound.  Used by the prioritized samplers (ReplayMemory/Sampling.cpp:145,206,231)."""

    def __init__(self, weights):
        p = np.asarray(weights, f32).astype(f64)
        if p.size < 2:
            self.cp = np.zeros(0, f64)
            return
        total = np.add.accumulate(p)[-1]                      # std::accumulate(..., 0.0): left to right
        self.cp = np.add.accumulate(p / total)                # std::partial_sum
        self.cp[-1] = 1.0

    def __call__(self, gen: Mt19937) -> int:
        if self.cp.size == 0:
            return 0
        lo, hi = gen(), gen()
        u = (float(lo) + float(hi) * 4294967296.0) / 18446744073709551616.0
        if u >= 1.0:
            u = float(np.nextafter(1.0, 0.0))
        return int(np.searchsorted(self.cp, u, side="left"))


def canonical_float(gen: Mt19937) -> np.float32:
    """std::uniform_real_distribution<float>(0, 1): generate_canonical<float, 24> = one 32-bit draw / 2^32 in float,
    clamped below 1 (bits/random.tcc)."""
    u = f32(f32(gen()) / f32(4294967296.0))
    return f32(np.nextafter(f32(1), f32(0))) if u >= f32(1) else u


# ------------------------------------------------------------------------------------------
# scalar helpers
# ------------------------------------------------------------------------------------------
def scale_net2v(x):
    """Learners/RACER_common.cpp:23-27 (f64)."""
    x = np.asarray(x, f64)
    pos = 100 * (x + 51) - 100 * np.sqrt(2601 + 100 * np.maximum(x, 0))
    neg = 100 * (x - 51) + 100 * np.sqrt(2601 - 100 * np.minimum(x, 0))
    return np.where(x > 0, pos, neg)


def scale_vdiff(x):
    """Learners/RACER_common.cpp:28-32 (f64)."""
    x = np.asarray(x, f64)
    return np.where(x > 0, 100 - 5000 / np.sqrt(2601 + 100 * np.maximum(x, 0)),
                    100 - 5000 / np.sqrt(2601 - 100 * np.minimum(x, 0)))


def softplus(x):
    """Network/Layers/Functions.h:552-555 (SMARTIES_CHEAP_SOFTPLUS)."""
    return (x + np.sqrt(1 + x * x)) / 2


def softplus_diff(x):
    """Network/Layers/Functions.h:556-563."""
    return (1 + x / np.sqrt(1 + x * x)) / 2


def tanh_f32(x):
    """Tanh::_eval, Network/Layers/Functions.h:103-112, evaluated in f32."""
    x = np.asarray(x, f32)
    e = np.exp((f32(-2) * np.abs(x)).astype(f32)).astype(f32)
    y = ((f32(1) - e) / (f32(1) + e)).astype(f32)
    return np.where(x > 0, y, -y).astype(f32)


def sigm_act_f32(x):
    """Sigm::_eval, Functions.h:158-165, f32."""
    x = np.asarray(x, f32)
    e = np.exp(-np.abs(x)).astype(f32)
    return np.where(x > 0, f32(1) / (f32(1) + e), e / (f32(1) + e)).astype(f32)


# hidden-layer functions "nnFunc" (makeFunction, Functions.h:643-668) with initFactor sqrt(6 / (in + out)):
# name -> (eval(in), evalDiff(in, out)), f32 like the reference's SINGLE_PREC build
ACTIVATIONS = {
    "Tanh": (tanh_f32, lambda x, y: (f32(1) - y * y).astype(f32)),                                          # :103-116
    "SoftSign": (lambda x: (np.asarray(x, f32) / (f32(1) + np.abs(np.asarray(x, f32)))).astype(f32),          # :328-337
                 lambda x, y: (f32(1) / ((f32(1) + np.abs(x)) * (f32(1) + np.abs(x)))).astype(f32)),
    "HardSign": (lambda x: (np.asarray(x, f32) / np.sqrt(f32(1) + np.asarray(x, f32) ** 2)).astype(f32),       # :220-229
                 lambda x, y: (f32(1) / (np.sqrt(f32(1) + x * x) ** 3)).astype(f32)),
    "Sigm": (sigm_act_f32, lambda x, y: (y * (f32(1) - y)).astype(f32)),                                       # :158-182
    "Relu": (lambda x: np.where(np.asarray(x, f32) > 0, x, f32(0)).astype(f32),                                # :415-423
             lambda x, y: np.where(x > 0, f32(1), f32(0)).astype(f32)),
    "LRelu": (lambda x: np.where(np.asarray(x, f32) > 0, x, (f32(0.1) * np.asarray(x, f32)).astype(f32)).astype(f32),   # PRELU_FAC 0.1 (:16-18,461-468)
              lambda x, y: np.where(x > 0, f32(1), f32(0.1)).astype(f32)),
    # Utilities::safeExp clips its argument at +-SMARTIES_EXP_CUT = 8 (FunctionUtilities.h:50-54)
    "ExpPlus": (lambda x: np.log(f32(1) + np.exp(np.clip(np.asarray(x, f32), f32(-8), f32(8))).astype(f32)).astype(f32),      # :507-518
                lambda x, y: (f32(1) / (f32(1) + np.exp(np.clip(-x, f32(-8), f32(8))).astype(f32))).astype(f32)),
    "SoftPlus": (lambda x: ((np.asarray(x, f32) + np.sqrt(f32(1) + np.asarray(x, f32) ** 2)) / f32(2)).astype(f32),          # :552-563
                 lambda x, y: ((f32(1) + x / np.sqrt(f32(1) + x * x)) / f32(2)).astype(f32)),
    "Exp": (lambda x: np.exp(np.clip(np.asarray(x, f32), f32(-8), f32(8))).astype(f32),                                       # :604-617
            lambda x, y: np.asarray(y, f32)),
    "Linear": (lambda x: np.asarray(x, f32), lambda x, y: np.ones_like(np.asarray(x, f32))),                                  # :66-74
}


def anneal_rate(eta, t, eps):
    """Utils/FunctionUtilities.h:69-72."""
    return eta / (1 + t * eps)


# ------------------------------------------------------------------------------------------
# network description (MLP as built by RACER::setupNet, Learners/RACER_common.cpp:70-115,
# Network/Approximator.cpp:179-229, Network/Builder.cpp:48-99)
# ------------------------------------------------------------------------------------------
class MlpLayout:
    """Parameter blob layout (Network/Layers/Parameters.h:159-176): per layer W then b, each
    rounded up to 8 floats.  Dense W is [nIn][roundUp8(nOut)] (Layer_Base.h:46,64-95)."""

    def __init__(self, dS: int, hidden, n_dense_out: int, n_param_out: int):
        self.dS, self.hidden = int(dS), [int(h) for h in hidden if h > 0]
        self.n_dense_out, self.n_param_out = int(n_dense_out), int(n_param_out)
        self.layers = []  # dicts: kind, nIn, nOut, ldw, w, b
        off = 0
        n_in = self.dS
        lid = 1
        for li, h in enumerate(self.hidden):
            ld = round_up8(h)
            L = dict(kind="dense_tanh", nIn=n_in, nOut=h, ldw=ld, w=off, id=lid)
            off += round_up8(ld * n_in)
            L["b"] = off
            off += round_up8(h)
            self.layers.append(L)
            lid += 1
            if li > 0:  # ParametricResidual after every hidden layer except the first (Builder.cpp:92-95)
                R = dict(kind="residual", n=h, w=off, id=lid)
                off += round_up8(h)
                R["b"] = off
                off += round_up8(h)
                self.layers.append(R)
                lid += 1
            n_in = h
        ld = round_up8(n_dense_out)
        L = dict(kind="dense_linear", nIn=n_in, nOut=n_dense_out, ldw=ld, w=off, id=lid)
        off += round_up8(ld * n_in)
        L["b"] = off
        off += round_up8(n_dense_out)
        self.layers.append(L)
        if n_param_out > 0:
            P = dict(kind="param", n=n_param_out, b=off, id=lid + 1)
            off += round_up8(n_param_out)
            self.layers.append(P)
        self.n_params = off
        self.n_out = n_dense_out + n_param_out

    def strip_padding(self, blob):
        """Order of Network::save (Layer_Base.h:143-153, Layers.h:401-410,554-560)."""
        out = []
        for L in self.layers:
            if L["kind"].startswith("dense"):
                W = blob[L["w"]:L["w"] + L["nIn"] * L["ldw"]].reshape(L["nIn"], L["ldw"])[:, :L["nOut"]]
                out += [W.ravel(), blob[L["b"]:L["b"] + L["nOut"]]]
            elif L["kind"] == "residual":
                out += [blob[L["w"]:L["w"] + L["n"]], blob[L["b"]:L["b"] + L["n"]]]
            else:
                out += [blob[L["b"]:L["b"] + L["n"]]]
        return np.concatenate(out).astype(f32)


def _dense_fwd(x, W, b):
    """BaseLayer::forward (Layer_Base.h:64-80): X = b; for i: X += x_i * W[i][:] — f32,
    sequential in i, separate multiply and add (reference build has no FMA)."""
    acc = np.broadcast_to(b, (x.shape[0], b.shape[0])).astype(f32).copy()
    for i in range(W.shape[0]):
        acc += (x[:, i:i + 1] * W[i][None, :]).astype(f32)
    return acc


class _Acts(list):
    """Layer outputs of one forward pass + the pre-activations of the hidden dense layers."""
    def __init__(self, it=()):
        super().__init__(it)
        self.pre = {}


class MlpNet:
    def __init__(self, layout: MlpLayout, func: str = "Tanh"):
        self.L = layout
        self.act, self.act_diff = ACTIVATIONS[func]

    def views(self, blob):
        v = []
        for L in self.L.layers:
            if L["kind"].startswith("dense"):
                W = blob[L["w"]:L["w"] + L["nIn"] * L["ldw"]].reshape(L["nIn"], L["ldw"])[:, :L["nOut"]]
                v.append((W, blob[L["b"]:L["b"] + L["nOut"]]))
            elif L["kind"] == "residual":
                v.append((blob[L["w"]:L["w"] + L["n"]], blob[L["b"]:L["b"] + L["n"]]))
            else:
                v.append((None, blob[L["b"]:L["b"] + L["n"]]))
        return v

    def forward(self, blob, x):
        """Network::forward (Network.h:101-113).  x: [B, dS] f32.  Returns (O f32 [B, nOut], cache)."""
        x = np.asarray(x, f32)
        Bn = x.shape[0]
        Y = _Acts([x])  # Y[k] = output of layer k (Y[0] = input layer); Y.pre[k] = pre-activation of hidden dense layer k
        views = self.views(blob)
        outs = []
        for L, (W, b) in zip(self.L.layers, views):
            k = L["kind"]
            if k == "dense_tanh":
                pre = _dense_fwd(Y[-1], W, b)
                Y.pre[len(Y)] = pre
                Y.append(self.act(pre))
            elif k == "residual":  # ParametricResidualLayer::forward (Layers.h:347-361): the first min(size(ID-2), size) units
                m = min(Y[-2].shape[1], L["n"])
                y = Y[-1][:, :L["n"]].astype(f32).copy()
                y[:, :m] = (y[:, :m] + (Y[-2][:, :m] * W[None, :m] + b[None, :m]).astype(f32)).astype(f32)
                Y.append(y)
            elif k == "dense_linear":
                Y.append(_dense_fwd(Y[-1], W, b))
                outs.append(Y[-1])
            else:  # ParamLayer::forward (Layers.h:510-521), Linear
                Y.append(np.broadcast_to(b, (Bn, b.shape[0])).astype(f32))
                outs.append(Y[-1])
        return np.concatenate(outs, axis=1).astype(f32), Y

    def backward(self, blob, Y, gout, sequential=True):
        """Network::backProp for one time step (Network.h:216-226) over a batch; returns the
        summed parameter gradient (thread-private `partialGradient` + reduceThreadsGrad,
        Parameters.h:66-103, for ONE thread: samples accumulate in batch order, in f32).
        gout: [B, nOut] f32 output deltas (Activation::addOutputDelta, Activation.h:112-120)."""
        G = np.zeros(self.L.n_params, f32)
        views = self.views(blob)
        Bn = gout.shape[0]
        E = [np.zeros_like(y) for y in Y]  # errvals per layer (Y index = layer id)
        # place output deltas
        k0 = 0
        for li, L in enumerate(self.L.layers):
            if L["kind"] == "dense_linear":
                E[li + 1] = gout[:, k0:k0 + L["nOut"]].astype(f32).copy(); k0 += L["nOut"]
            elif L["kind"] == "param":
                E[li + 1] = gout[:, k0:k0 + L["n"]].astype(f32).copy(); k0 += L["n"]

        def acc_rows(dst, contrib):  # dst[...] += sum over batch, sequential in b, f32
            if sequential:
                for bb in range(Bn):
                    dst += contrib[bb]
            else:
                dst += contrib.sum(axis=0, dtype=f32)

        for li in range(len(self.L.layers) - 1, -1, -1):
            L = self.L.layers[li]
            W, b = views[li]
            yi = li + 1
            kind = L["kind"]
            if kind == "param":  # ParamLayer::backward (Layers.h:523-546)
                acc_rows(G[L["b"]:L["b"] + L["n"]], E[yi])
            elif kind.startswith("dense"):  # BaseLayer::backward (Layer_Base.h:97-113) + Layer::backward (Layers.h:123-188)
                if kind == "dense_tanh":
                    E[yi] = (E[yi] * self.act_diff(Y.pre[yi], Y[yi])).astype(f32)      # Function::evalDiff(in, out), Layer_Base.h:103-109
                d = E[yi]
                first = (li == 0)  # input gradient skipped for layer 1 (Approximator.cpp:145-169)
                if not first:
                    E[yi - 1] = (E[yi - 1] + (d @ W.T).astype(f32)).astype(f32)
                acc_rows(G[L["b"]:L["b"] + L["nOut"]], d)
                Gw = G[L["w"]:L["w"] + L["nIn"] * L["ldw"]].reshape(L["nIn"], L["ldw"])[:, :L["nOut"]]
                xin = Y[yi - 1]
                if sequential:
                    for bb in range(Bn):
                        Gw += (xin[bb][:, None] * d[bb][None, :]).astype(f32)
                else:
                    Gw += (xin.T @ d).astype(f32)
            else:  # ParametricResidualLayer::backward (Layers.h:363-393)
                d = E[yi]
                m = min(Y[yi - 2].shape[1], L["n"])
                E[yi - 1] = d.copy()  # memcpy into E(ID-1)
                E[yi - 2][:, :m] = (E[yi - 2][:, :m] + (d[:, :m] * W[None, :m]).astype(f32)).astype(f32)
                acc_rows(G[L["w"]:L["w"] + m], (d[:, :m] * Y[yi - 2][:, :m]).astype(f32))
                acc_rows(G[L["b"]:L["b"] + m], d[:, :m])
        return G


# ------------------------------------------------------------------------------------------
# per-sample V-RACER loss / gradient   (Learners/RACER_train.cpp:12-67)
# ------------------------------------------------------------------------------------------
def discrete_sample_math(O, act, mu, qret, beta, cmax, cinv):
    """RACER<Discrete_advantage, Discrete_policy, Uint>::Train (Learners/RACER_train.cpp:12-67) for K action options:
    O [B, 1 + 2K] = [V | advantages(K) | policy pre-activations(K)] (RACER_common.cpp:109-135), act [B, 1] the stored
    action message (label + 0.1, Core/StateAction.h:320-341), mu [B, K] the behaviour probabilities.  Policy: SoftPlus of
    the pre-activations, normalised (Math/Discrete_policy.h:64-85); advantage centred with the policy's expectation
    (Math/Discrete_advantage.h:44-75).  Same return dict as vracer_sample_math."""
    O = np.asarray(O, f64)
    mu = np.asarray(mu, f64)
    Bn = O.shape[0]
    K = (O.shape[1] - 1) // 2
    opt = np.floor(np.asarray(act, f64)[:, 0]).astype(np.int64)           # actionMessage2label (StateAction.h:304-319)
    adv, raw = O[:, 1:1 + K], O[:, 1 + K:1 + 2 * K]
    unnorm = softplus(raw)
    norm = np.zeros(Bn, f64)
    for j in range(K):
        norm = norm + unnorm[:, j]
    norm = np.maximum(norm, np.finfo(f64).eps)
    probs = unnorm / norm[:, None]
    rows = np.arange(Bn)
    rho = probs[rows, opt] / mu[rows, opt]                                # importanceWeight (:87-94): no clipping here
    dkl = np.zeros(Bn, f64)
    for j in range(K):                                                    # KLDivergence (:129-133)
        dkl = dkl + probs[:, j] * np.log(probs[:, j] / mu[:, j])
    W32, C32, I32 = rho.astype(f32), f32(cmax), f32(cinv)
    is_far = (C32 > f32(1)) & ((W32 > C32) | (W32 < I32))
    expA = np.zeros(Bn, f64)
    for j in range(K):
        expA = expA + probs[:, j] * adv[:, j]
    Aval = adv[rows, opt] - expA
    V = scale_net2v(O[:, 0])
    a_ret = np.asarray(qret, f64) - V
    dq = a_ret - Aval
    g = np.zeros_like(O)
    g[:, 0] = np.where(is_far, 0.0, np.minimum(1.0, rho) * dq * beta * scale_vdiff(O[:, 0]))
    dpos = softplus_diff(raw)
    onehot = np.zeros((Bn, K), f64); onehot[rows, opt] = 1.0
    penal = np.zeros((Bn, K), f64)                                        # KLDivGradient(MU, -1) (:158-167)
    for j in range(K):
        tmp = -1.0 * (1 + np.log(probs[:, j] / mu[:, j])) / norm
        ej = np.zeros((Bn, K), f64); ej[:, j] = 1.0
        penal = penal + tmp[:, None] * (ej - probs[:, j:j + 1])
    penal = penal * dpos
    fac = a_ret * np.minimum(cmax, rho)                                   # policyGradient(ACT, fac) (:139-147)
    pol = onehot * (fac / unnorm[rows, opt])[:, None]
    pol = (pol - (fac / norm)[:, None]) * dpos
    pol = np.where(is_far[:, None], 0.0, pol)
    g[:, 1 + K:1 + 2 * K] = beta * pol + (1 - beta) * penal               # penalizeReFER
    err = np.where(is_far, 0.0, beta * np.minimum(cmax, rho) * dq)        # ADV.grad(act, isFar ? 0 : beta*Aer) (:53-61)
    g[:, 1:1 + K] = err[:, None] * (onehot - probs)
    return dict(rho=rho, dkl=dkl, is_far=is_far, V=V, A=Aval, dq=dq, g=g)


def vracer_sample_math(O, act, mu, qret, beta, cmax, cinv, bounded=None, racer=False):
    """O: [B, nOut] network outputs (f32 values widened to f64, Approximator.h:117-173);
    V-RACER: [V | mean(dA) | stdev-param(dA)], RACER: [V | adv coef, p1(dA), p2(dA) | mean(dA) | stdev-param(dA)]
    (RACER_common.cpp:174-193,232-247).  act [B,dA], mu [B,2dA] (stored f32 -> f64 Rvec, Episode.h:66),
    qret [B] f32.  Returns rho, dkl, is_far, V, A (advantage), dq (f64) and g [B, nOut] (f64 output gradient)."""
    O = np.asarray(O, f64)
    act = np.asarray(act, f64)
    mu = np.asarray(mu, f64)
    Bn, dA = act.shape
    if bounded is None:
        bounded = np.zeros(dA, bool)
    bounded = np.asarray(bounded, bool)
    m0 = 2 + 2 * dA if racer else 1          # start of the policy means
    mean = O[:, m0:m0 + dA]
    sraw = O[:, m0 + dA:m0 + 2 * dA]
    stdev = softplus(sraw)                      # Continuous_policy.h:78-81
    inv = 1 / stdev
    mu_m, mu_s = mu[:, :dA], mu[:, dA:]
    cmean = np.where(bounded[None, :], np.clip(mean, -SQUASH_MAX, SQUASH_MAX), mean)  # :217-222

    def logp(a, m, invs):                       # Continuous_policy.h:91-97 (Jacobian term of :240-249 cancels in rho)
        return -((a - m) * invs) ** 2 / 2 + np.log(invs) - 9.1893853320467266954096885456237942e-01

    if bounded.any():
        squash = np.tanh(act)
        J = np.maximum(1 - squash * squash, FLT_MIN)
        lp_pi = -((act - cmean) * inv) ** 2 / 2 + np.log(inv / np.where(bounded, J, 1.0)) - 9.1893853320467266954096885456237942e-01
        lp_mu = -((act - mu_m) * (1 / mu_s)) ** 2 / 2 + np.log((1 / mu_s) / np.where(bounded, J, 1.0)) - 9.1893853320467266954096885456237942e-01
    else:
        lp_pi = logp(act, mean, inv)
        lp_mu = logp(act, mu_m, 1 / mu_s)
    logw = np.zeros(Bn, f64)
    for i in range(dA):                         # importanceWeight, Continuous_policy.h:648-653
        logw = logw + (lp_pi[:, i] - lp_mu[:, i])
    rho = np.exp(np.clip(logw, -7, 7))
    c = (stdev / mu_s) ** 2                     # KLdivergence, OPPOSITE_KL branch :138-142
    dm = ((mean - mu_m) / mu_s) ** 2
    klc = (c - 1 + dm - np.log(c)) / 2
    dkl = np.zeros(Bn, f64)
    for i in range(dA):
        dkl = dkl + klc[:, i]
    # isFarPolicy takes Fval arguments (ReplayMemory/Episode.h:28-33): compare in f32
    W32, C32, I32 = rho.astype(f32), f32(cmax), f32(cinv)
    is_far = (C32 > f32(1)) & ((W32 > C32) | (W32 < I32))
    V = scale_net2v(O[:, 0])
    Aval = np.zeros(Bn, f64)                    # Zero_advantage.h:39-42
    if racer:                                   # Gaussian_advantage::computeAdvantage (Gaus_advantage.h:73-86,116-126)
        coef_raw = O[:, 1]
        p1r, p2r = O[:, 2:2 + dA], O[:, 2 + dA:2 + 2 * dA]
        coef, p1, p2 = softplus(coef_raw), softplus(p1r), softplus(p2r)
        S = stdev * stdev                       # policy->getVariance (Continuous_policy.h:775-777)
        dm_adv = act - cmean                    # policy->getMean() is the clamped mean for bounded dims
        upper = act > cmean
        terms = dm_adv ** 2 / np.where(upper, p1, p2)
        shape_sum = np.zeros(Bn, f64)
        for i in range(dA):
            shape_sum = shape_sum + terms[:, i]
        orig = np.exp(-shape_sum / 2)
        rfac = np.sqrt(p1 / (p1 + S)) / 2 + np.sqrt(p2 / (p2 + S)) / 2
        ratio = np.ones(Bn, f64)
        for i in range(dA):
            ratio = ratio * rfac[:, i]
        Aval = coef * (orig - ratio)
    a_ret = np.asarray(qret, f64) - V
    dq = a_ret - Aval
    ver = np.minimum(1.0, rho) * dq
    g = np.zeros_like(O)
    g[:, 0] = np.where(is_far, 0.0, ver * beta * scale_vdiff(O[:, 0]))
    # penalG = KLDivGradient(MU, -1)  (Continuous_policy.h:709-716 -> gradKLdiv :154-170)
    dpos = softplus_diff(sraw)
    inv_var_mu = 1 / mu_s ** 2
    kg_mean = -1.0 * ((mean - mu_m) * inv_var_mu)
    kg_std = dpos * -1.0 * ((inv_var_mu - inv ** 2) * stdev)
    # polG = policyGradient(ACT, A_RET*min(Cmax,rho))  (:694-701 -> gradLogP :145-152 / :300-316)
    fac = (a_ret * np.minimum(cmax, rho))[:, None]
    u = (act - cmean) * inv
    dlp_mean = np.where(bounded[None, :], (act - mean) * inv * inv, u * inv)
    dlp_std = (u * u - 1) * inv
    pg_mean = fac * dlp_mean
    blocked = bounded[None, :] & (((mean >= SQUASH_MAX) & (pg_mean > 0)) | ((mean <= -SQUASH_MAX) & (pg_mean < 0)))
    pg_mean = np.where(blocked, 0.0, pg_mean)
    pg_std = dpos * fac * dlp_std
    far = is_far[:, None]
    pg_mean = np.where(far, 0.0, pg_mean)
    pg_std = np.where(far, 0.0, pg_std)
    g[:, m0:m0 + dA] = beta * pg_mean + (1 - beta) * kg_mean        # penalizeReFER, FunctionUtilities.h:221-228
    g[:, m0 + dA:m0 + 2 * dA] = beta * pg_std + (1 - beta) * kg_std
    if racer:                                   # ADV.grad(act, isFar ? 0 : beta*Aer, gradient)  (Gaus_advantage.h:88-114,64-69)
        aer = np.minimum(cmax, rho) * dq
        err = np.where(is_far, 0.0, beta * aer)
        expect = -ratio
        g[:, 1] = (orig + expect) * (err * softplus_diff(coef_raw))
        oc = (orig * coef)[:, None]
        x1 = np.where(upper, oc * (dm_adv / p1) ** 2 / 2, 0.0)
        x2 = np.where(act < cmean, oc * (dm_adv / p2) ** 2 / 2, 0.0)
        F = 2 / (np.sqrt(p1 / (p1 + S)) + np.sqrt(p2 / (p2 + S)))
        diff1 = S / np.sqrt(p1 * (p1 + S) ** 3) / 4
        diff2 = S / np.sqrt(p2 * (p2 + S) ** 3) / 4
        ec = (expect[:, None] * 1.0)
        x1 = x1 + F * ec * coef[:, None] * diff1
        x2 = x2 + F * ec * coef[:, None] * diff2
        g[:, 2:2 + dA] = x1 * (err[:, None] * softplus_diff(p1r))
        g[:, 2 + dA:2 + 2 * dA] = x2 * (err[:, None] * softplus_diff(p2r))
    return dict(rho=rho, dkl=dkl, is_far=is_far, V=V, A=Aval, dq=dq, g=g)



# ------------------------------------------------------------------------------------------
# Output-gradient statistics (Utils/StatsTracker.cpp:28-107)
# ------------------------------------------------------------------------------------------
CLIP_LEARNR = 1e-3   # Settings/Bund.h:97


class StatsTracker:
    """Per net output: sum and sum of squares of the gradient handed to Approximator::setGradient
    (Network/Approximator.h:197 -> track_vector, StatsTracker.cpp:28-36), reduced once per learner step by
    reduce_stats (:99-106, called from Learner_approximator.cpp:89 with iter = nGradSteps()).  `words` is what
    printToFile (:66-86) has written to `<base>_outGrad_stats.raw` so far, as float32 words: the header n_stats + 0.1
    if the very first reduce_stats call printed, then per printed step the means followed by the root mean squares.
    np.longdouble is the x86-64 80-bit long double of the reference's accumulators."""

    def __init__(self, n_stats: int):
        self.n = n_stats
        self.cnt = 0
        self.sum = np.zeros(n_stats, np.longdouble)
        self.sq = np.zeros(n_stats, np.longdouble)
        self.avg = np.zeros(n_stats, np.longdouble)     # exponential averages kept after finalize (:88-97)
        self.std = np.zeros(n_stats, np.longdouble)
        self.inst_mean = np.zeros(n_stats, np.longdouble)
        self.inst_stdv = np.zeros(n_stats, np.longdouble)
        self.n_step = 0
        self.words: list = []

    def track_vector(self, grad):
        g = np.asarray(grad, f64).astype(np.longdouble)
        assert g.shape == (self.n,)
        self.cnt += 1
        self.sum += g
        self.sq += g * g

    def reduce_stats(self, it: int):
        old_m, old_s = self.avg.copy(), self.std.copy()
        cnt = max(np.longdouble(2.2e-16), np.longdouble(self.cnt))          # advance + update (:38-64)
        mean = (self.sum / cnt).astype(f64)                                  # `const Real mean`
        rms = np.sqrt((self.sq / cnt).astype(f64))
        self.cnt = 0; self.sum[:] = 0; self.sq[:] = 0
        if it % 1000 == 0:                                                   # printToFile
            if self.n_step == 0:
                self.words.append(f32(self.n + .1))
            self.words.extend(mean.astype(f32)); self.words.extend(rms.astype(f32))
        self.inst_mean, self.inst_stdv = mean.astype(np.longdouble), rms.astype(np.longdouble)   # finalize
        self.n_step += 1
        self.avg = (1 - CLIP_LEARNR) * old_m + CLIP_LEARNR * self.inst_mean
        self.std = (1 - CLIP_LEARNR) * old_s + CLIP_LEARNR * self.inst_stdv

    def file_words(self):
        return np.asarray(self.words, f32)


# ------------------------------------------------------------------------------------------
# replay memory + learner
# ------------------------------------------------------------------------------------------
class Episode:
    """ReplayMemory/Episode.h:40-231 (fields needed by the path)."""

    def __init__(self, eid, S, A, MU, R, terminated):
        N = S.shape[0]
        self.ID = int(eid)
        self.S = S.astype(f32)
        self.A = A.astype(f64)         # actions / policies are Rvec = f64 in the reference
        self.MU = MU.astype(f64)
        self.R = R.astype(f64)
        self.R[0] = 0.0
        self.terminated = bool(terminated)
        # Episode::finalize (Episode.cpp:244-266)
        self.V = np.zeros(N, f32); self.ADV = np.zeros(N, f32); self.Q = np.zeros(N, f32)
        self.delta = np.zeros(N, f32); self.KL = np.zeros(N, f32)
        self.rho = np.ones(N, f32); self.rho[-1] = 0
        self.avgKL = f32(0); self.fracFar = f32(0); self.avgSqErr = f32(0); self.maxAbsErr = f32(0)
        self.sumQ2 = f32(0); self.sumQ = f32(0); self.maxQ = f32(-1e9); self.minQ = f32(1e9)
        self.totR = f32(np.sum(self.R[1:].astype(f32), dtype=f32))
        self.just_sampled = -1

    @property
    def nsteps(self):
        return self.S.shape[0]

    @property
    def ndata(self):
        return self.S.shape[0] - 1

    def update_cumulative(self, C, invC):
        """Episode::updateCumulative (Episode.cpp:213-242), f32."""
        N = self.ndata
        invN = f32(1) / f32(N)
        C, invC = f32(C), f32(invC)
        far = (self.rho[:N] > C) | (self.rho[:N] < invC)
        nfar = int(far.sum())
        sumE2 = f32(0); sumQ2 = f32(0); sumQ1 = f32(0)
        Qv = (self.ADV[:N] + self.V[:N]).astype(f32)
        d2 = (self.delta[:N] * self.delta[:N]).astype(f32)
        q2 = (Qv * Qv).astype(f32)
        for t in range(N):
            sumE2 = f32(sumE2 + d2[t]); sumQ2 = f32(sumQ2 + q2[t]); sumQ1 = f32(sumQ1 + Qv[t])
        self.fracFar = f32(invN * f32(nfar))
        self.avgSqErr = f32(invN * sumE2)
        self.maxAbsErr = f32(max(-1e9, float(np.max(np.abs(self.delta[:N])))))
        self.sumQ2, self.sumQ = sumQ2, sumQ1
        self.maxQ = f32(max(-1e9, float(Qv.max()))); self.minQ = f32(min(1e9, float(Qv.min())))
        self.totR = f32(np.sum(self.R, dtype=f64))          # Utilities::sum over Real rewards
        klsum = f32(0)
        for t in range(self.nsteps):                          # Utilities::sum(KullbLeibDiv) over all rows
            klsum = f32(klsum + self.KL[t])
        self.avgKL = f32(invN * klsum)


class VracerOracle:
    """Learner-only V-RACER: initializeLearner + {spawnTrainTasks, processMemoryBuffer,
    applyGradient, globalGradCounterUpdate} (Learners/RACER.cpp:61-110)."""

    def __init__(self, dS, dA, hidden=(128, 128), gamma=0.995, lam=1.0, clip_imp_weight=None,
                 penal_tol=0.1, eps_anneal=5e-7, learnrate=1e-4, nn_lambda=FLT_EPS,
                 batch=256, max_tot_obs=None, bounded=False, sample_seed=42, learner="VRACER",
                 returns_estimator="retrace", sampling="uniform", er_filter="oldest", discrete=0, refer_threads=1, nn_func="Tanh",
                 target_delay=0.0):
        self.dS, self.dA = dS, dA
        # "targetDelay" (AdamOptimizer::tgtUpdateAlpha, Optimizer.cpp:162-177): W_tgt is created on the first update (a copy of
        # the weights the learner was constructed with, Optimizer.h / Approximator::initializeNetwork)
        self.target_delay, self.cnt_update_delay, self.W_tgt = float(target_delay), 0, None
        # OpenMP threads of the reference run: only the far-policy count depends on it (`Uint += float` partials per thread with
        # schedule(static, 1), MemoryProcessing.cpp:202-227)
        self.refer_threads = max(1, int(refer_threads))
        self.discrete = int(discrete)        # K options of a discrete action space (then dA == 1 and learner == "RACER")
        # getERfilterAlgo (MemoryProcessing.cpp:261-298): "a goes before b"; the episodes to delete end up at the back
        self.er_before = {"oldest": lambda a, b: a.ID > b.ID, "default": lambda a, b: a.ID > b.ID,
                          "farpolfrac": lambda a, b: a.fracFar < b.fracFar, "maxkldiv": lambda a, b: a.avgKL < b.avgKL,
                          "minerror": lambda a, b: a.avgSqErr > b.avgSqErr}[er_filter]
        if sampling not in ("uniform", "PERerr", "PERseq", "PERrank"):    # Sampling::prepareSampler (Sampling.cpp:298-335)
            raise NotImplementedError(sampling)
        self.sampling, self.dist = sampling, None
        if returns_estimator not in ("retrace", "GAE", "retraceExplore"):   # createReturnEstimator (MemoryProcessing.cpp:419-450)
            raise NotImplementedError(returns_estimator)
        self.gae = returns_estimator == "GAE"
        self.explore = returns_estimator == "retraceExplore"
        self.racer = learner == "RACER"
        if self.discrete:                    # [V | adv(K) | policy(K)] from one linear layer, no ParamLayer (RACER_common.cpp:71-107)
            self.layout = MlpLayout(dS, hidden, 1 + 2 * self.discrete, 0)
        else:
            self.layout = MlpLayout(dS, hidden, (2 + 3 * dA) if self.racer else (1 + dA), dA)
        self.net = MlpNet(self.layout, nn_func)
        self.gamma, self.lam = gamma, lam
        self.C = float(np.sqrt(dA / 2.0)) if clip_imp_weight is None else float(clip_imp_weight)  # HyperParameters.h:46
        self.penal_tol, self.eps_anneal, self.eta, self.nn_lambda = penal_tol, eps_anneal, learnrate, nn_lambda
        self.B = batch
        self.max_tot_obs = int(2 ** 14 * np.sqrt(dA + dS)) if max_tot_obs is None else int(max_tot_obs)
        self.max_tot_obs_local = self.max_tot_obs
        self.bounded = np.full(dA, bool(bounded))
        # MemoryBuffer.h:41-44
        self.beta = 1.0 if self.C <= 0 else 1e-4
        self.cmax = 1 + self.C
        self.cinv = 1 / self.C if self.C > 0 else np.inf
        self.state_mean = np.zeros(dS, f32); self.state_std = np.ones(dS, f32); self.state_scale = np.ones(dS, f32)
        self.rew_mean = f32(0); self.rew_std = f32(1); self.rew_scale = f32(1)
        self.episodes: list[Episode] = []
        self.n_grad_steps = 0
        self.adam_step = 0
        self.beta_t_1, self.beta_t_2 = 0.9, 0.999            # Optimizer.h:93-94 (Real)
        self.W = np.zeros(self.layout.n_params, f32)
        self.M1 = np.zeros_like(self.W); self.M2 = np.zeros_like(self.W)
        self.gen = Mt19937(sample_seed)
        self.n_far_policy = 0
        self.stats = dict(avgKL=0.0, avgSqErr=0.0, maxAbsErr=0.0, avgReturn=0.0, stdevQ=0.0, avgQ=0.0, maxQ=0.0, minQ=0.0,
                          cntRet=0, sumRetErr=0.0)
        self.last = {}

    # ---- data ----
    def load_replay(self, d):
        for e in range(len(d["N"])):
            o, N = int(d["start"][e]), int(d["N"][e])
            ep = Episode(e, d["S"][o:o + N], d["A"][o:o + N], d["MU"][o:o + N], d["R"][o:o + N], d["term"][e])
            self.retrace_episode(ep)                          # MemoryBuffer.cpp:143 computeReturnEstimator at insertion
            self.push_back_episode(ep)

    def push_back_episode(self, ep: Episode):
        """MemoryBuffer::pushBackEpisode (MemoryBuffer.cpp:479-520): pre-training TD-error
        placeholder sqrt(max(FLT_EPS, stats.avgSquaredErr)) (Episode.cpp:268-273)."""
        err = f32(np.sqrt(max(FLT_EPS, self.stats["avgSqErr"])))
        ep.delta[:] = err
        ep.avgSqErr = f32(err * err)
        ep.maxAbsErr = err
        self.episodes.append(ep)

    @property
    def n_transitions(self):
        return int(sum(ep.ndata for ep in self.episodes))

    # ---- Retrace (ReplayMemory/MemoryProcessing.cpp:23-44, 391-400) ----
    def retrace_episode(self, ep: Episode) -> float:
        N = ep.nsteps
        if not ep.terminated:
            ep.Q[N - 1] = ep.V[N - 1]
        g, l = f32(self.gamma), f32(self.lam)
        rs = ((ep.R - f64(self.rew_mean)) * f64(self.rew_scale)).astype(f32)   # scaledReward<Fval>, Episode.h:184-189
        w = np.where(ep.rho < 1, ep.rho, f32(1)).astype(f32)                   # clippedOffPolW, Episode.h:190-194
        err2 = f32(0)
        for t in range(N - 2, -1, -1):
            old = ep.Q[t]
            Qn, Vn, An = ep.Q[t + 1], ep.V[t + 1], ep.ADV[t + 1]
            if self.gae:     # computeGAE (MemoryProcessing.cpp:411-417)
                new = f32(rs[t + 1] + g * f32(Vn + f32(l * f32(Qn - Vn))))
            else:            # computeRetrace (:391-400)
                new = f32(rs[t + 1] + g * f32(Vn + f32(f32(l * w[t + 1]) * f32(f32(Qn - An) - Vn))))
                if self.explore:   # computeRetraceExplBonus (:402-409): (1 - gamma) * (|Q' - A - V| - stats.maxAbsError) on top
                    E = f32(abs(f32(f32(Qn - An) - Vn)) - f32(self.stats["maxAbsErr"]))
                    new = f32(f32(f32(f32(1) - g) * E) + new)
            ep.Q[t] = new
            err2 = f32(err2 + f32(old - new) ** 2)
        return float(err2)

    # ---- reward / state moments (MemoryProcessing.cpp:94-185) ----
    def update_rewards_stats(self, b_init: bool, rate_fac: float = 1.0):
        learn_r = anneal_rate(self.eta, self.n_grad_steps, self.eps_anneal)
        w = 1.0 if b_init else min(1.0, rate_fac * learn_r)
        ld = np.longdouble
        cnt = ld(0); rs = ld(0); rs2 = ld(0)
        ss = np.zeros(self.dS, ld); ss2 = np.zeros(self.dS, ld)
        for ep in self.episodes:
            N = ep.ndata
            cnt += N
            dr = ep.R[1:N + 1].astype(ld) - ld(self.rew_mean)
            rs += dr.sum(dtype=ld); rs2 += (dr * dr).sum(dtype=ld)
            ds = ep.S[:N].astype(ld) - self.state_mean.astype(ld)[None, :]
            ss += ds.sum(axis=0, dtype=ld); ss2 += (ds * ds).sum(axis=0, dtype=ld)

        def upd(mean, std, lr, ev, ev2):
            mean = f32(ld(mean) + ld(lr) * ev)
            var = ev2 - ev * ev * ld(2 * lr - lr * lr)
            var = max(var, ld(FLT_EPS))
            std = f32(ld(std) + ld(lr) * (np.sqrt(var) - ld(std)))
            return mean, std, f32(f32(1) / std)

        self.rew_mean, self.rew_std, self.rew_scale = upd(self.rew_mean, self.rew_std, w, rs / cnt, rs2 / cnt)
        for k in range(self.dS):
            self.state_mean[k], self.state_std[k], self.state_scale[k] = upd(
                self.state_mean[k], self.state_std[k], w, ss[k] / cnt, ss2[k] / cnt)

    # ---- beta (MemoryProcessing.cpp:46-92) ----
    def update_counters(self):
        n_stored = self.n_transitions
        frac = self.n_far_policy / float(max(n_stored, 1))
        lr = 0.1 * self.B / max(float(self.max_tot_obs), float(n_stored))
        b = self.beta
        if frac > self.penal_tol:
            self.beta = (1 - min(lr, b)) * b
        else:
            self.beta = (1 - min(lr, b)) * b + min(lr, 1 - b)

    def initialize_learner(self):
        """Learner::initializeLearner (Learners/Learner.cpp:47-72)."""
        self.update_counters()
        self.update_rewards_stats(True)
        self.update_sampler()                                 # Learner.cpp:63
        for ep in self.episodes:                              # rescaleAllReturnEstimator :460-481
            self.retrace_episode(ep)

    # ---- samplers ----
    def update_sampler(self):
        """MemoryBuffer::updateSampler -> Sampling::prepare: TSample_impErr (Sampling.cpp:173-206) weighs every transition
        by sqrt(sqrt(delta^2 + eps)), Sample_impSeq (:230-255) every episode by sqrt(sqrt(avgSquaredErr + eps)) * ndata;
        float arithmetic, the distribution itself in double.  Sample_uniform::prepare does nothing."""
        eps32 = f32(FLT_EPS)
        if self.sampling == "PERerr":
            w = np.concatenate([np.sqrt(np.sqrt((ep.delta[:ep.ndata] * ep.delta[:ep.ndata]).astype(f32) + eps32)).astype(f32)
                                for ep in self.episodes])
            self.dist = DiscreteDistribution(w)
        elif self.sampling == "PERseq":
            w = np.array([f32(np.sqrt(np.sqrt(f32(ep.avgSqErr + eps32)))) * f32(ep.ndata) for ep in self.episodes], f32)
            self.dist = DiscreteDistribution(w)
        elif self.sampling == "PERrank":
            # TSample_impRank::prepare (Sampling.cpp:102-146): all squared errors ranked by std::sort (descending, unstable:
            # never-sampled transitions of an episode share one placeholder error), weight (rank+1)^-1/4 computed in double
            # (std::sqrt of an unsigned) and stored as float; zero errors weigh 1
            errs, prefix = [], []
            for i, ep in enumerate(self.episodes):
                prefix.append(len(errs))
                d2 = (ep.delta[:ep.ndata] * ep.delta[:ep.ndata]).astype(f32)
                errs += [(float(d2[j]), i, j) for j in range(ep.ndata)]
            std_sort(errs, lambda a, b: a[0] > b[0])
            w = np.ones(len(errs), f32)
            for r, (e, i, j) in enumerate(errs):
                w[prefix[i] + j] = f32(1.0 / np.sqrt(np.sqrt(float(r + 1)))) if e > 0 else f32(1)
            self.dist = DiscreteDistribution(w)

    # ---- one gradient step ----
    def sample(self):
        nd = np.array([ep.ndata for ep in self.episodes], np.int64)
        if self.sampling == "PERseq":      # Sample_impSeq::sample, transition branch (Sampling.cpp:276-294)
            S: list = []
            while len(S) < self.B:
                for _ in range(self.B - len(S)):
                    k = self.dist(self.gen)
                    S.append((k, int(canonical_float(self.gen) * f32(nd[k]))))     # float * Uint -> float -> Uint
                S = sorted(set(S))
            return np.array([a for a, _ in S], np.int64), np.array([b for _, b in S], np.int64)
        if self.sampling in ("PERerr", "PERrank"):   # TSample_impErr / TSample_impRank::sample (:148-166, :208-225): the
            # same draw / sort / unique loop as the uniform sampler
            ret: list = []
            while len(ret) < self.B:
                ret += [self.dist(self.gen) for _ in range(self.B - len(ret))]
                ret = sorted(set(ret))
            ids = np.asarray(ret, np.int64)
        else:
            ids = sample_uniform(self.gen, self.n_transitions, self.B)
        return id_to_seq_step(ids, nd)

    def standardized(self, ep: Episode, t: int):
        return ((ep.S[t] - self.state_mean) * self.state_scale).astype(f32)   # Episode.h:171-183

    def train_step(self, seq=None, obs=None, inject_O=None):
        """spawnTrainTasks + processMemoryBuffer + applyGradient + globalGradCounterUpdate."""
        if seq is None:
            seq, obs = self.sample()
        B = len(seq)
        eps = [self.episodes[int(s)] for s in seq]
        X = np.stack([self.standardized(ep, int(t)) for ep, t in zip(eps, obs)])
        O32, Y = self.net.forward(self.W, X)
        # V(s_{t+1}) for truncated episodes (RACER_train.cpp:23-27)
        trunc = [b for b in range(B) if (int(obs[b]) + 2 == eps[b].nsteps and not eps[b].terminated)]
        Vnext = {}
        if trunc:
            Xn = np.stack([self.standardized(eps[b], int(obs[b]) + 1) for b in trunc])
            On, _ = self.net.forward(self.W, Xn)
            for j, b in enumerate(trunc):
                Vnext[b] = f32(scale_net2v(f64(On[j, 0])))
        Ouse = O32 if inject_O is None else np.asarray(inject_O)
        act = np.stack([ep.A[int(t)] for ep, t in zip(eps, obs)])
        mu = np.stack([ep.MU[int(t)] for ep, t in zip(eps, obs)])
        qret = np.array([ep.Q[int(t)] for ep, t in zip(eps, obs)], f32)
        if self.discrete:
            r = discrete_sample_math(Ouse, act, mu, qret, self.beta, self.cmax, self.cinv)
        else:
            r = vracer_sample_math(Ouse, act, mu, qret, self.beta, self.cmax, self.cinv, self.bounded, self.racer)
        g32 = r["g"].astype(f32)
        self._track_grad_stats(r["g"])
        # write-back + per-episode aggregates, in batch order (Episode.h:112-145)
        C32, I32 = f32(self.cmax), f32(self.cinv)
        for b in range(B):
            ep, t = eps[b], int(obs[b])
            invN = f32(1) / f32(ep.nsteps)
            if b in Vnext:
                self._update_values(ep, t + 1, Vnext[b], Vnext[b])
            E, D, Wt = f32(r["dq"][b]), f32(r["dkl"][b]), f32(r["rho"][b])
            was = f32((ep.rho[t] > C32) or (ep.rho[t] < I32))
            isf = f32((Wt > C32) or (Wt < I32))
            ep.avgKL = f32(ep.avgKL + f32(invN * f32(D - ep.KL[t])))
            ep.fracFar = f32(ep.fracFar + f32(invN * f32(isf - was)))
            ep.avgSqErr = f32(ep.avgSqErr + f32(invN * f32(f32(E * E) - f32(ep.delta[t] * ep.delta[t]))))
            ep.maxAbsErr = f32(max(ep.maxAbsErr, abs(E)))
            ep.delta[t], ep.KL[t], ep.rho[t] = E, D, Wt
            Vv = f32(r["V"][b])
            self._update_values(ep, t, Vv, f32(r["A"][b] + r["V"][b]))   # Qval = Aval + Vval
        G = self.net.backward(self.W, Y, g32)
        self.adam_step += 1                                    # prepare_update: nStep++ (Optimizer.cpp:119)
        self.last = dict(r, seq=np.asarray(seq), obs=np.asarray(obs), X=X, O=O32, g=g32, g64=r['g'], gradSum=G.copy())
        self.process_memory_buffer()
        self.apply_adam(G)
        self.n_grad_steps += 1
        return self.last

    def _track_grad_stats(self, g64):
        """Approximator::setGradient -> track_vector per sample, then updateGradStats once per step
        (Network/Approximator.h:65-68,197; Learner_approximator.cpp:89).  Active once `self.grad_stats` is set."""
        tr = getattr(self, "grad_stats", None)
        if tr is None:
            return
        for row in g64:
            tr.track_vector(row)
        tr.reduce_stats(self.n_grad_steps)

    @staticmethod
    def _update_values(ep: Episode, t: int, V, Q):
        """Episode::updateValues_atomic (Episode.h:131-145)."""
        V, Q = f32(V), f32(Q)
        oldQ = f32(ep.ADV[t] + ep.V[t])
        ep.sumQ2 = f32(ep.sumQ2 + f32(f32(Q * Q) - f32(oldQ * oldQ)))
        ep.sumQ = f32(ep.sumQ + f32(Q - oldQ))
        ep.maxQ = f32(max(ep.maxQ, Q)); ep.minQ = f32(min(ep.minQ, Q))
        ep.V[t] = V; ep.ADV[t] = f32(Q - V)

    def process_memory_buffer(self):
        """Learner::processMemoryBuffer (Learners/Learner.cpp:74-100)."""
        step = self.n_grad_steps + 1
        recompute = step % 1000 == 0
        # updateTrainingStatistics (MemoryProcessing.cpp:187-259)
        self.cmax = 1 + anneal_rate(self.C, step, self.eps_anneal)
        self.cinv = 1 / self.cmax
        T = self.refer_threads
        n_off_thr = [0] * T                                   # reduction(+ : nOffPol): one private Uint per OpenMP thread
        sumDKL = sumE2 = sumQ2 = sumQ1 = sumR = sumERet = 0.0
        maxAbsE, maxQ, minQ = f32(-1e9), f32(-1e9), f32(1e9)
        n_ret = 0
        for pos, ep in enumerate(self.episodes):              # schedule(static, 1): iteration i runs on thread i % T
            if recompute:
                ep.update_cumulative(self.cmax, self.cinv)
                sumERet += self.retrace_episode(ep)
                n_ret += ep.nsteps - 1
            Ns = f32(ep.nsteps)
            maxAbsE = max(maxAbsE, ep.maxAbsErr); maxQ = max(maxQ, ep.maxQ); minQ = min(minQ, ep.minQ)
            sumDKL += float(f32(Ns * ep.avgKL))
            n_off_thr[pos % T] = uint_plus_float(n_off_thr[pos % T], f32(Ns * ep.fracFar))  # `Uint += float` (SURVEY §7 hard part 2)
            sumE2 += float(f32(Ns * ep.avgSqErr))
            sumQ2 += float(ep.sumQ2); sumQ1 += float(ep.sumQ); sumR += float(ep.totR)
            ep.just_sampled = -1
        n_off = sum(n_off_thr) & ((1 << 64) - 1)              # the partial counts are combined as integers
        if self.cmax <= 1:
            n_off = 0
        n_data = self.n_transitions
        self.n_far_policy = n_off
        lr = 0.1 * self.B / max(float(self.max_tot_obs), float(n_data))
        st = self.stats
        st["maxAbsErr"] += lr * (float(maxAbsE) - st["maxAbsErr"])
        st["avgKL"] = sumDKL / n_data; st["avgSqErr"] = sumE2 / n_data
        st["avgReturn"] = sumR / len(self.episodes); st["avgQ"] = sumQ1 / n_data
        st["maxQ"], st["minQ"] = float(maxQ), float(minQ)
        st["stdevQ"] = float(np.sqrt(max(sumQ2 / n_data - st["avgQ"] ** 2, 1e-16)))
        st["cntRet"] = max(st["cntRet"], 0) + n_ret; st["sumRetErr"] += sumERet
        if recompute:
            self.update_rewards_stats(False, 10)
        # applyEpisodesRemovalAlgo (MemoryProcessing.cpp:327-351): "oldest" = sort by ID descending
        std_sort(self.episodes, self.er_before)               # std::sort: ties fall as libstdc++'s introsort leaves them
        while self.n_transitions - self.episodes[-1].nsteps > self.max_tot_obs_local:
            self.episodes.pop()                               # removeBackEpisode (MemoryBuffer.cpp:469-477)
        self.update_sampler()                                 # RM.updateSampler() (MemoryProcessing.cpp:350)
        self.update_counters()

    def apply_adam(self, G):
        """AdamOptimizer::apply_update + struct Adam (Network/Optimizer.cpp:61-108,122-161), f32."""
        # `Saru gen(nStep, thrID, generators[thrID]())` (Optimizer.cpp:139): thread 0 draws one
        # 32-bit value from generators[0] — the sampler's generator — every update.
        self.gen()
        eta0 = f32(f64(f32(self.eta)) / (1 + f64(f32(self.adam_step)) * self.eps_anneal))  # annealRate<nnReal>
        bt1, bt2 = f32(self.beta_t_1), f32(self.beta_t_2)
        eta = f32(f32(eta0 * f32(np.sqrt(f32(f32(1) - bt2)))) / f32(f32(1) - bt1))
        B1, B2, lam, fac = f32(0.9), f32(0.999), f32(self.nn_lambda), f32(1.0 / self.B)
        W, M1, M2 = self.W, self.M1, self.M2
        penal = (-W * lam).astype(f32)
        DW = (fac * G).astype(f32)
        M1[:] = (B1 * M1 + (f32(1) - B1) * DW).astype(f32)
        M2[:] = (B2 * M2 + ((f32(1) - B2) * DW).astype(f32) * DW).astype(f32)
        numer = (B1 * M1 + (f32(1) - B1) * DW).astype(f32)
        M2[:] = np.where(M2 < M1 * M1, (M1 * M1).astype(f32), M2)
        ret = (numer / (f32(FLT_EPS) + np.sqrt(M2).astype(f32)).astype(f32)).astype(f32)
        if self.target_delay > 0 and self.W_tgt is None:
            self.W_tgt = W.copy()                                    # target_weights->copy(weights) at construction
        W += (eta * (ret + penal).astype(f32)).astype(f32)
        if self.target_delay > 0:                                    # "update frozen weights" (Optimizer.cpp:162-177)
            if self.cnt_update_delay == 0:
                self.cnt_update_delay = int(self.target_delay)       # Uint cntUpdateDelay = tgtUpdateAlpha
                if self.target_delay >= 1:
                    self.W_tgt[:] = W
                else:                                                # Real alpha times the f32 difference, added to the f32 target
                    self.W_tgt[:] = (self.W_tgt.astype(f64) + self.target_delay * (W - self.W_tgt).astype(f32).astype(f64)).astype(f32)
            if self.cnt_update_delay > 0:
                self.cnt_update_delay -= 1
        self.beta_t_1 *= 0.9
        if self.beta_t_1 < FLT_EPS: self.beta_t_1 = 0
        self.beta_t_2 *= 0.999
        if self.beta_t_2 < FLT_EPS: self.beta_t_2 = 0

    # ---- flat views used by tests ----
    def concat(self, field):
        return np.concatenate([getattr(ep, field) for ep in self.episodes])


# ------------------------------------------------------------------------------------------
# recurrent networks (nnType "LSTM"): LSTMLayer (Network/Layers/Layer_LSTM.h:77-166) + BPTT in the
# order of Network::backProp (Network/Network.h:155-193: layer-major, time-minor)
# ------------------------------------------------------------------------------------------
def sigm_f32(x):
    """Sigm::_eval with safeExp clipped at +-SMARTIES_EXP_CUT = 8 (Functions.h:158-165, Definitions.h:43)."""
    x = np.asarray(x, f32)
    e = np.exp(np.clip(-np.abs(x), -8, 8).astype(f32)).astype(f32)     # exp(-|x|), |x| clipped at 8
    pos = (f32(1) / (f32(1) + e)).astype(f32)
    neg = (e / (f32(1) + e)).astype(f32)
    return np.where(x > 0, pos, neg).astype(f32)


class SeqLayout:
    """Parameter blob for Input -> LSTM x H (ParametricResidual after every hidden layer but the first)
    -> Linear out -> ParamLayer; same padding rules as MlpLayout (Parameters.h:159-176)."""

    def __init__(self, dS, cells, n_dense_out, n_param_out, cell_type="LSTM"):
        self.dS, self.cells = int(dS), [int(c) for c in cells if c > 0]
        self.layers = []
        off, n_in = 0, self.dS
        # gates per cell: LSTMLayer 4 (Layer_LSTM.h:30-37), MGULayer 2 — "MGU" and "GRU" both build it (Builder.cpp:67-72,
        # Layer_GRU.h:27-33)
        kind, ng = ("lstm", 4) if cell_type == "LSTM" else ("mgu", 2)
        if cell_type not in ("LSTM", "MGU", "GRU"):
            raise NotImplementedError(cell_type)
        for li, c in enumerate(self.cells):
            L = dict(kind=kind, nIn=n_in, nC=c, ld=ng * c, w=off)
            off += round_up8(ng * c * (n_in + c)); L["b"] = off; off += round_up8(ng * c)
            self.layers.append(L)
            if li > 0:
                R = dict(kind="residual", n=c, w=off); off += round_up8(c); R["b"] = off; off += round_up8(c)
                self.layers.append(R)
            n_in = c
        ld = round_up8(n_dense_out)
        L = dict(kind="dense_linear", nIn=n_in, nOut=n_dense_out, ldw=ld, w=off)
        off += round_up8(ld * n_in); L["b"] = off; off += round_up8(n_dense_out)
        self.layers.append(L)
        P = dict(kind="param", n=n_param_out, b=off); off += round_up8(n_param_out)
        self.layers.append(P)
        self.n_params, self.n_out = off, n_dense_out + n_param_out


class SeqNet:
    def __init__(self, layout: SeqLayout):
        self.L = layout

    def forward_seq(self, blob, X):
        """X: [T+1, dS] standardized states of the window; the recurrent state starts at zero at the first
        window step (Approximator.h:129-139: no `prev` activation).  Returns (O of every step [T+1, nOut], cache)."""
        T1 = X.shape[0]
        cache = []          # per step: list over layers of dicts
        outs = []
        prev = None
        for k in range(T1):
            acts = [dict(y=X[k].astype(f32))]
            o_parts = []
            for li, L in enumerate(self.L.layers):
                kind = L["kind"]
                if kind == "lstm":
                    nI, nC = L["nIn"], L["nC"]
                    W = blob[L["w"]:L["w"] + 4 * nC * (nI + nC)].reshape(nI + nC, 4 * nC)
                    s = blob[L["b"]:L["b"] + 4 * nC].astype(f32).copy()
                    xin = acts[-1]["y"][:nI]
                    for i in range(nI):
                        s += (xin[i] * W[i]).astype(f32)
                    hp = prev[li + 1]["y"] if prev is not None else None
                    if hp is not None:
                        for i in range(nC):
                            s += (hp[i] * W[nI + i]).astype(f32)
                    cin = s[:nC].copy()
                    g = sigm_f32(s[nC:])
                    ig, fg, og = g[:nC], g[nC:2 * nC], g[2 * nC:]
                    stp = prev[li + 1]["st"] if prev is not None else None
                    old = (stp * fg).astype(f32) if stp is not None else np.zeros(nC, f32)
                    st = ((cin * ig).astype(f32) + old).astype(f32)
                    cop = tanh_f32(st)
                    y = (og * cop).astype(f32)
                    acts.append(dict(y=y, st=st, cop=cop, cin=cin, ig=ig, fg=fg, og=og, xin=xin, hp=hp, stp=stp))
                elif kind == "mgu":     # MGULayer::forward (Layer_GRU.h:65-120): W rows [input | recurrent], columns [forget | state]
                    nI, nC = L["nIn"], L["nC"]
                    W = blob[L["w"]:L["w"] + 2 * nC * (nI + nC)].reshape(nI + nC, 2 * nC)
                    s = blob[L["b"]:L["b"] + 2 * nC].astype(f32).copy()
                    xin = acts[-1]["y"][:nI]
                    for i in range(nI):
                        s += (xin[i] * W[i]).astype(f32)
                    hp = prev[li + 1]["y"] if prev is not None else None
                    fpre, spre = s[:nC].copy(), s[nC:].copy()
                    if hp is not None:
                        for i in range(nC):
                            fpre += (W[nI + i, :nC] * hp[i]).astype(f32)
                        fg = sigm_f32(fpre)
                        for i in range(nC):
                            spre += ((W[nI + i, nC:] * hp[i]).astype(f32) * fg[i]).astype(f32)
                        st = tanh_f32(spre)
                        y = ((fg * st).astype(f32) + ((f32(1) - fg).astype(f32) * hp).astype(f32)).astype(f32)
                    else:
                        fg, st = sigm_f32(fpre), tanh_f32(spre)
                        y = (fg * st).astype(f32)
                    acts.append(dict(y=y, fg=fg, st=st, xin=xin, hp=hp))
                elif kind == "residual":
                    n = L["n"]
                    w, b = blob[L["w"]:L["w"] + n], blob[L["b"]:L["b"] + n]
                    y = (acts[-1]["y"][:n] + (acts[-2]["y"][:n] * w + b).astype(f32)).astype(f32)
                    acts.append(dict(y=y))
                elif kind == "dense_linear":
                    W = blob[L["w"]:L["w"] + L["nIn"] * L["ldw"]].reshape(L["nIn"], L["ldw"])[:, :L["nOut"]]
                    y = _dense_fwd(acts[-1]["y"][None, :L["nIn"]], W, blob[L["b"]:L["b"] + L["nOut"]])[0]
                    acts.append(dict(y=y)); o_parts.append(y)
                else:
                    y = blob[L["b"]:L["b"] + L["n"]].astype(f32)
                    acts.append(dict(y=y)); o_parts.append(y)
            cache.append(acts)
            outs.append(np.concatenate(o_parts))
            prev = acts
        return np.stack(outs).astype(f32), cache

    def backward_seq(self, blob, cache, gout, G):
        """BPTT with the output delta `gout` placed at the LAST cached step only; accumulates into G
        in the reference's order (layers top to bottom, for each layer time T..0)."""
        T = len(cache) - 1
        nl = len(self.L.layers)
        E = [[np.zeros_like(cache[k][li + 1]["y"]) if self.L.layers[li]["kind"] not in ("lstm", "mgu")
              else np.zeros(self.L.layers[li]["ld"], f32) for li in range(nl)] for k in range(T + 1)]
        Ein = [np.zeros(self.L.dS, f32) for _ in range(T + 1)]
        k0 = 0
        for li, L in enumerate(self.L.layers):
            if L["kind"] == "dense_linear":
                E[T][li] = gout[k0:k0 + L["nOut"]].astype(f32).copy(); k0 += L["nOut"]
            elif L["kind"] == "param":
                E[T][li] = gout[k0:k0 + L["n"]].astype(f32).copy(); k0 += L["n"]
        sdelta = {}
        for li in range(nl - 1, -1, -1):
            L = self.L.layers[li]
            kind = L["kind"]
            for k in range(T, -1, -1):
                a = cache[k][li + 1]
                below = E[k][li - 1] if li > 0 else Ein[k]
                if kind == "param":
                    G[L["b"]:L["b"] + L["n"]] += E[k][li]
                elif kind == "dense_linear":
                    d = E[k][li]
                    W = blob[L["w"]:L["w"] + L["nIn"] * L["ldw"]].reshape(L["nIn"], L["ldw"])[:, :L["nOut"]]
                    below[:L["nIn"]] = (below[:L["nIn"]] + (W @ d).astype(f32)).astype(f32)
                    G[L["b"]:L["b"] + L["nOut"]] += d
                    Gw = G[L["w"]:L["w"] + L["nIn"] * L["ldw"]].reshape(L["nIn"], L["ldw"])[:, :L["nOut"]]
                    Gw += (cache[k][li]["y"][:L["nIn"], None] * d[None, :]).astype(f32)
                elif kind == "residual":
                    n = L["n"]
                    d = E[k][li]
                    w = blob[L["w"]:L["w"] + n]
                    E[k][li - 1][:n] = d                      # memcpy into E(ID-1) (first n entries)
                    E[k][li - 2][:n] = (E[k][li - 2][:n] + (d * w).astype(f32)).astype(f32)
                    G[L["w"]:L["w"] + n] += (d * cache[k][li - 1]["y"][:n]).astype(f32)
                    G[L["b"]:L["b"] + n] += d
                elif kind == "mgu":  # MGULayer::backward (Layer_GRU.h:122-214)
                    nI, nC = L["nIn"], L["nC"]
                    W = blob[L["w"]:L["w"] + 2 * nC * (nI + nC)].reshape(nI + nC, 2 * nC)
                    fg, st, hp = a["fg"], a["st"], a["hp"]
                    dO = E[k][li][:nC].copy()
                    pO = hp if hp is not None else np.zeros(nC, f32)
                    dS_ = ((dO * fg).astype(f32) * (f32(1) - (st * st).astype(f32)).astype(f32)).astype(f32)             # 1)
                    dFp = (W[nI:, nC:] @ dS_).astype(f32) if hp is not None else np.zeros(nC, f32)                        # 2)
                    dF = ((((st - pO).astype(f32) * dO).astype(f32) + (dFp * pO).astype(f32)).astype(f32)
                          * fg).astype(f32) * (f32(1) - fg).astype(f32)                                                  # 3)
                    dF = dF.astype(f32)
                    if hp is not None:                                                                                   # 4)
                        Ep = E[k - 1][li]
                        Ep[:nC] = (Ep[:nC] + (((f32(1) - fg).astype(f32) * dO).astype(f32) + (fg * dFp).astype(f32)).astype(f32)).astype(f32)
                        Ep[:nC] = (Ep[:nC] + (W[nI:, :nC] @ dF).astype(f32)).astype(f32)
                    E[k][li][nC:] = dF
                    # input gradient: spanCompInpGrads = nInputs for every MGU layer (Layer_GRU.h:47), the first one included
                    below[:nI] = (below[:nI] + (W[:nI, :nC] @ dF).astype(f32)).astype(f32)
                    below[:nI] = (below[:nI] + (W[:nI, nC:] @ dS_).astype(f32)).astype(f32)
                    dl = np.concatenate([dF, dS_]).astype(f32)
                    G[L["b"]:L["b"] + 2 * nC] += dl
                    Gw = G[L["w"]:L["w"] + 2 * nC * (nI + nC)].reshape(nI + nC, 2 * nC)
                    Gw[:nI] += (a["xin"][:, None] * dl[None, :]).astype(f32)
                    if hp is not None:
                        Gw[nI:, :nC] += (hp[:, None] * dF[None, :]).astype(f32)
                        Gw[nI:, nC:] += ((hp[:, None] * dS_[None, :]).astype(f32) * fg[:, None]).astype(f32)
                else:  # LSTMLayer::backward (Layer_LSTM.h:127-166)
                    nI, nC = L["nIn"], L["nC"]
                    W = blob[L["w"]:L["w"] + 4 * nC * (nI + nC)].reshape(nI + nC, 4 * nC)
                    D = E[k][li][:nC].copy()
                    diff = ((f32(1) - a["cop"] * a["cop"]).astype(f32) * D).astype(f32)
                    sd = (diff * a["og"]).astype(f32)
                    if k < T:
                        nxt = cache[k + 1][li + 1]
                        sd = (sd + (sdelta[(li, k + 1)] * nxt["fg"]).astype(f32)).astype(f32)
                    sdelta[(li, k)] = sd
                    dl = np.zeros(4 * nC, f32)
                    dl[:nC] = a["ig"] * sd
                    dl[nC:2 * nC] = ((a["ig"] * (f32(1) - a["ig"])).astype(f32) * a["cin"]).astype(f32) * sd
                    if a["stp"] is not None:
                        dl[2 * nC:3 * nC] = ((a["fg"] * (f32(1) - a["fg"])).astype(f32) * a["stp"]).astype(f32) * sd
                    dl[3 * nC:] = ((a["og"] * (f32(1) - a["og"])).astype(f32) * D).astype(f32) * a["cop"]
                    E[k][li] = dl
                    if li > 0:                               # input gradient (skipped for the first layer)
                        below[:nI] = (below[:nI] + (W[:nI] @ dl).astype(f32)).astype(f32)
                    if a["hp"] is not None:                  # recurrent error into the previous step's E(ID)[0:nC]
                        E[k - 1][li][:nC] = (E[k - 1][li][:nC] + (W[nI:] @ dl).astype(f32)).astype(f32)
                    G[L["b"]:L["b"] + 4 * nC] += dl
                    Gw = G[L["w"]:L["w"] + 4 * nC * (nI + nC)].reshape(nI + nC, 4 * nC)
                    Gw[:nI] += (a["xin"][:, None] * dl[None, :]).astype(f32)
                    if a["hp"] is not None:
                        Gw[nI:] += (a["hp"][:, None] * dl[None, :]).astype(f32)
        return G


class RecurrentOracle(VracerOracle):
    """RACER / V-RACER with nnType LSTM: sampled transition t is evaluated on the window
    [t - min(nnBPTTseq, t), t] (MemoryBuffer.cpp:393-402), loss at t only, BPTT over the window."""

    def __init__(self, dS, dA, cells=(64,), bptt=16, cell_type="LSTM", **kw):
        learner = kw.get("learner", "VRACER")
        super().__init__(dS, dA, hidden=(8,), **kw)
        racer = learner == "RACER"
        self.layout = SeqLayout(dS, cells, (2 + 3 * dA) if racer else (1 + dA), dA, cell_type)
        self.net = SeqNet(self.layout)
        self.bptt = int(bptt)
        self.W = np.zeros(self.layout.n_params, f32)
        self.M1 = np.zeros_like(self.W); self.M2 = np.zeros_like(self.W)

    def train_step(self, seq=None, obs=None):
        if seq is None:
            seq, obs = self.sample()
        B = len(seq)
        eps = [self.episodes[int(s)] for s in seq]
        G = np.zeros(self.layout.n_params, f32)
        Os, caches, Vnext = [], [], {}
        for b in range(B):
            ep, t = eps[b], int(obs[b])
            n_rec = min(self.bptt, t)
            X = np.stack([self.standardized(ep, k) for k in range(t - n_rec, t + 1)])
            O, cache = self.net.forward_seq(self.W, X)
            if t + 2 == ep.nsteps and not ep.terminated:     # V(s_{t+1}) with the recurrent state of step t
                X2 = np.concatenate([X, self.standardized(ep, t + 1)[None]])
                O2, _ = self.net.forward_seq(self.W, X2)
                Vnext[b] = f32(scale_net2v(f64(O2[-1, 0])))
            Os.append(O[-1]); caches.append(cache)
        O32 = np.stack(Os)
        act = np.stack([ep.A[int(t)] for ep, t in zip(eps, obs)])
        mu = np.stack([ep.MU[int(t)] for ep, t in zip(eps, obs)])
        qret = np.array([ep.Q[int(t)] for ep, t in zip(eps, obs)], f32)
        r = vracer_sample_math(O32, act, mu, qret, self.beta, self.cmax, self.cinv, self.bounded, self.racer)
        g32 = r["g"].astype(f32)
        self._track_grad_stats(r["g"])
        C32, I32 = f32(self.cmax), f32(self.cinv)
        for b in range(B):
            ep, t = eps[b], int(obs[b])
            invN = f32(1) / f32(ep.nsteps)
            if b in Vnext:
                self._update_values(ep, t + 1, Vnext[b], Vnext[b])
            E, D, Wt = f32(r["dq"][b]), f32(r["dkl"][b]), f32(r["rho"][b])
            was = f32((ep.rho[t] > C32) or (ep.rho[t] < I32)); isf = f32((Wt > C32) or (Wt < I32))
            ep.avgKL = f32(ep.avgKL + f32(invN * f32(D - ep.KL[t])))
            ep.fracFar = f32(ep.fracFar + f32(invN * f32(isf - was)))
            ep.avgSqErr = f32(ep.avgSqErr + f32(invN * f32(f32(E * E) - f32(ep.delta[t] * ep.delta[t]))))
            ep.maxAbsErr = f32(max(ep.maxAbsErr, abs(E)))
            ep.delta[t], ep.KL[t], ep.rho[t] = E, D, Wt
            self._update_values(ep, t, f32(r["V"][b]), f32(r["A"][b] + r["V"][b]))
            self.net.backward_seq(self.W, caches[b], g32[b], G)
        self.adam_step += 1
        X_last = np.stack([c[-1][0]["y"] for c in caches])
        self.last = dict(r, seq=np.asarray(seq), obs=np.asarray(obs), X=X_last, O=O32, g=g32, g64=r["g"], gradSum=G.copy())
        self.process_memory_buffer()
        self.apply_adam(G)
        self.n_grad_steps += 1
        return self.last
