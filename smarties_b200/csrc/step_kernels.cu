// step_kernels.cu — the V-RACER learner step on one B200 (sm_100a).
//
// One learner step of the reference,
//     spawnTrainTasks(); processMemoryBuffer(); applyGradient();      (Learners/RACER.cpp:81-109)
// is three device phases separated by grid-wide dependencies:
//   P1  per tile of TB sampled transitions: gather + standardise states from the HBM replay
//       buffer, MLP forward, ReF-ER / Retrace loss and output gradient in f64, write-back of
//       V/Q/delta/KL/rho to the replay rows, input-gradient backward; activations and deltas
//       are left feature-major in a scratch that P2 reads.             (RACER_train.cpp:12-67)
//   P2  per 16x16 tile of every weight matrix: dW = A^T * Delta contracted over the whole
//       mini-batch in batch order, fused with the reference's Adam variant in the epilogue
//       (the per-thread gradient buffers and their reduction, Parameters.h:66-103, vanish).
//   P3  one CTA, concurrent with P2: per-episode aggregate updates in sample order
//       (Episode.h:112-145), the per-step replay statistics, Cmax annealing and the ReF-ER
//       beta fixed-point update (MemoryProcessing.cpp:46-92,187-259).
// The phases run either as two kernels per step or inside one persistent cooperative kernel
// that loops over many steps with two grid barriers per step.
//
// Compiled with -fmad=false: every multiply-add that must be fused is written fmaf()
// explicitly; everything else rounds like the reference's x86-64 (no-FMA) build.
#include "step_kernels.cuh"

#include <cfloat>
#include <cmath>

namespace smb200 {

// ------------------------------------------------------------------------------------------
// small helpers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float ld_cg(const float* p) { return __ldcg(p); }
__device__ __forceinline__ float4 ld_cg4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }

__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// Grid-wide barrier for the persistent kernel (all CTAs co-resident: cooperative launch).
// `counter` only grows; it is zeroed by the host before each launch.
__device__ __forceinline__ void grid_barrier(unsigned* counter, unsigned nblocks) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned old = atomicAdd(counter, 1u);
    const unsigned target = (old / nblocks + 1u) * nblocks;
    while (ld_acquire(counter) < target) { }
    __threadfence();
  }
  __syncthreads();
}

// Tanh::_eval (Network/Layers/Functions.h:103-112), f32
__device__ __forceinline__ float tanh_ref(float x) {
  const float e = expf(-2.0f * fabsf(x));
  const float y = (1.0f - e) / (1.0f + e);
  return x > 0.0f ? y : -y;
}

// scaleNet2V / scaleVdiff (Learners/RACER_common.cpp:23-32), f64
__device__ __forceinline__ double net2v(double x) {
  return x > 0 ? 100.0 * (x + 51.0) - 100.0 * sqrt(2601.0 + 100.0 * x)
               : 100.0 * (x - 51.0) + 100.0 * sqrt(2601.0 - 100.0 * x);
}
__device__ __forceinline__ double vdiff(double x) {
  return x > 0 ? 100.0 - 5000.0 / sqrt(2601.0 + 100.0 * x) : 100.0 - 5000.0 / sqrt(2601.0 - 100.0 * x);
}

// ------------------------------------------------------------------------------------------
// batched GEMV for a tile of TB samples.  x, y live in shared memory feature-major:
// x[k*TB + s].  Weights stream from L2 with coalesced loads over the output index; the K
// range is split over thread groups when the layer is narrower than the CTA.
//   mode 0: y = b + W^T x         mode 1: y = tanh(b + W^T x)        mode 2: y += W^T x
// ------------------------------------------------------------------------------------------
template <int TB>
__device__ __forceinline__ void dense_apply(const float* __restrict__ W, int ldw, int K, int N,
                                            const float* __restrict__ bias, const float* x, float* y,
                                            float* red, int mode) {
  const int tid = threadIdx.x;
  for (int n0 = 0; n0 < N; n0 += kThreads) {
    const int nc = min(kThreads, N - n0);
    int NR = 8;
    while (NR < nc) NR <<= 1;
    const int G = kThreads / NR;
    const int g = tid / NR, nl = tid - g * NR;
    const int Kc = (K + G - 1) / G;
    const int kb = g * Kc, ke = min(K, kb + Kc);
    float acc[TB];
#pragma unroll
    for (int s = 0; s < TB; ++s) acc[s] = 0.f;
    if (nl < nc) {
      const float* w = W + (size_t)kb * ldw + n0 + nl;
#pragma unroll 8
      for (int k = kb; k < ke; ++k, w += ldw) {
        const float wv = ld_cg(w);
#pragma unroll
        for (int s = 0; s < TB; ++s) acc[s] = fmaf(x[k * TB + s], wv, acc[s]);
      }
    }
    if (G > 1) {
      __syncthreads();
#pragma unroll
      for (int s = 0; s < TB; ++s) red[(g * NR + nl) * TB + s] = acc[s];
      __syncthreads();
      for (int idx = tid; idx < nc * TB; idx += kThreads) {
        const int n2 = idx / TB, s = idx - n2 * TB;
        float v = 0.f;
        for (int gg = 0; gg < G; ++gg) v += red[(gg * NR + n2) * TB + s];
        const int n = n0 + n2;
        if (mode == 2) { y[n * TB + s] += v; }
        else { v += ld_cg(bias + n); y[n * TB + s] = mode == 1 ? tanh_ref(v) : v; }
      }
    } else if (nl < nc) {
      const int n = n0 + nl;
      const float bv = mode == 2 ? 0.f : ld_cg(bias + n);
#pragma unroll
      for (int s = 0; s < TB; ++s) {
        if (mode == 2) y[n * TB + s] += acc[s];
        else { const float v = acc[s] + bv; y[n * TB + s] = mode == 1 ? tanh_ref(v) : v; }
      }
    }
  }
  __syncthreads();
}

// Network::forward (Network/Network.h:101-113) for the tile; act[L.actOff*TB ...] per layer.
template <int TB>
__device__ void net_forward(const StepArgs& a, const NetDesc& net, float* act, float* red) {
  for (int l = 1; l < net.nLayers; ++l) {
    const LayerDesc& L = net.L[l];
    float* y = act + L.actOff * TB;
    if (L.kind == kDenseTanh || L.kind == kDenseLinear) {
      dense_apply<TB>(a.W + L.wOff, L.ld, L.nIn, L.size, a.W + L.bOff, act + net.L[L.in].actOff * TB, y, red,
                      L.kind == kDenseTanh ? 1 : 0);
    } else if (L.kind == kResidual) {   // ParametricResidualLayer::forward (Layers.h:347-361)
      const float* y1 = act + net.L[l - 1].actOff * TB;
      const float* y2 = act + net.L[l - 2].actOff * TB;
      for (int idx = threadIdx.x; idx < L.size * TB; idx += kThreads) {
        const int n = idx / TB;
        y[idx] = y1[idx] + (y2[idx] * ld_cg(a.W + L.wOff + n) + ld_cg(a.W + L.bOff + n));
      }
      __syncthreads();
    } else {                            // ParamLayer::forward (Layers.h:510-521), Linear
      for (int idx = threadIdx.x; idx < L.size * TB; idx += kThreads) y[idx] = ld_cg(a.W + L.bOff + idx / TB);
      __syncthreads();
    }
  }
}

__device__ __forceinline__ float net_out(const NetDesc& net, const float* act, int TB, int j, int s) {
  const LayerDesc& Lo = net.L[net.nLayers - 2];   // linear output layer
  const LayerDesc& Lp = net.L[net.nLayers - 1];   // param layer (stdev)
  return j < net.nOutDense ? act[(Lo.actOff + j) * TB + s] : act[(Lp.actOff + j - net.nOutDense) * TB + s];
}

// ------------------------------------------------------------------------------------------
// P1
// ------------------------------------------------------------------------------------------
template <int TB>
__device__ void p1_tile(const StepArgs& a, const NetDesc& net, const Hyper& hp, const StepCtrl& c, int step, int tile,
                        float* smem) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float* act = smem;                             // [actPerSample][TB]
  float* err = act + net.actPerSample * TB;      // [actPerSample][TB]
  float* red = err + net.actPerSample * TB;      // [kThreads*TB]
  int* info = reinterpret_cast<int*>(red + kThreads * TB);   // row[TB], slot[TB], hasNext[TB], valid[TB]
  float* vnext = reinterpret_cast<float*>(info + 4 * TB);
  const int b0 = tile * TB;
  const ReplayView& rp = a.rp;
  const int dS = net.dS, dA = net.dA;

  if (tid < TB) {
    const int b = b0 + tid;
    int row = 0, slot = 0, hn = 0, valid = 0;
    if (b < a.B) {
      slot = a.sampSlot[(size_t)(step - a.stepBase) * a.B + b];
      const int t = a.sampT[(size_t)(step - a.stepBase) * a.B + b];
      row = __ldcg(rp.epStart + slot) + t;
      hn = (t + 2 == __ldcg(rp.epLen + slot)) && !__ldcg(rp.epTerm + slot);   // Episode::isTruncated(t+1)
      valid = 1;
    }
    info[tid] = row; info[TB + tid] = slot; info[2 * TB + tid] = hn; info[3 * TB + tid] = valid;
  }
  __syncthreads();
  int anyNext = 0;
#pragma unroll
  for (int s = 0; s < TB; ++s) anyNext |= info[2 * TB + s];

  // V(s_{t+1}) of truncated episodes (RACER_train.cpp:23-27): rare, extra forward pass
  if (anyNext) {
    for (int idx = tid; idx < dS * TB; idx += kThreads) {
      const int k = idx / TB, s = idx - k * TB;
      const int row = info[s] + 1;
      act[idx] = info[2 * TB + s] ? (ld_cg(rp.S + (size_t)row * dS + k) - ld_cg(rp.stateMean + k)) * ld_cg(rp.stateScale + k) : 0.f;
    }
    __syncthreads();
    net_forward<TB>(a, net, act, red);
    if (tid < TB) vnext[tid] = (float)net2v((double)net_out(net, act, TB, 0, tid));
    __syncthreads();
  }

  // gather + standardise: (s - mean) * scale   (Episode.h:171-183)
  for (int idx = tid; idx < dS * TB; idx += kThreads) {
    const int k = idx / TB, s = idx - k * TB;
    const float x = info[3 * TB + s] ? (ld_cg(rp.S + (size_t)info[s] * dS + k) - ld_cg(rp.stateMean + k)) * ld_cg(rp.stateScale + k) : 0.f;
    act[idx] = x;
    if (info[3 * TB + s]) a.lastX[(size_t)(b0 + s) * dS + k] = x;
  }
  for (int idx = tid; idx < net.actPerSample * TB; idx += kThreads) err[idx] = 0.f;   // clearErrors
  __syncthreads();
  net_forward<TB>(a, net, act, red);

  // ---- loss: one warp per sample, lanes over action components, f64 (RACER_train.cpp:31-60) ----
  if (warp < TB && info[3 * TB + warp]) {
    const int s = warp, b = b0 + s;
    const size_t row = info[s];
    const double beta = c.beta, cmax = c.cmax, cinv = c.cinv;
    const LayerDesc& Lo = net.L[net.nLayers - 2];
    const LayerDesc& Lp = net.L[net.nLayers - 1];
    double logw = 0.0, dkl = 0.0;
    // importanceWeight / KLDivergence: sums run sequentially over components like the reference
    for (int i0 = 0; i0 < dA; i0 += 32) {
      const int i = i0 + lane;
      double lw_i = 0.0, kl_i = 0.0;
      if (i < dA) {
        const double m = (double)net_out(net, act, TB, 1 + i, s);
        const double sraw = (double)net_out(net, act, TB, 1 + dA + i, s);
        const double av = (double)ld_cg(rp.A + row * dA + i);
        const double mm = (double)ld_cg(rp.MU + row * 2 * dA + i);
        const double ms = (double)ld_cg(rp.MU + row * 2 * dA + dA + i);
        const double stdev = (sraw + sqrt(1.0 + sraw * sraw)) / 2.0;      // SoftPlus::_eval, Functions.h:552-555
        const double inv = 1.0 / stdev, invmu = 1.0 / ms;
        const bool bnd = hp.bounded[i] != 0;
        const double MAXM = 8.31776613503286;
        const double cm = bnd ? (m > MAXM ? MAXM : (m < -MAXM ? -MAXM : m)) : m;   // Continuous_policy.h:217-222
        const double fac0 = 9.1893853320467266954096885456237942e-01;
        double J = 1.0;
        if (bnd) { const double sq = tanh(av); J = fmax(1.0 - sq * sq, (double)FLT_MIN); }
        const double z1 = (av - cm) * inv, z2 = (av - mm) * invmu;
        const double lp_pi = -(z1 * z1) / 2.0 + log(bnd ? inv / J : inv) - fac0;     // :91-97 / :240-249
        const double lp_mu = -(z2 * z2) / 2.0 + log(bnd ? invmu / J : invmu) - fac0;
        lw_i = lp_pi - lp_mu;
        const double r1 = stdev / ms, r2 = (m - mm) / ms;
        const double cc = r1 * r1, dd = r2 * r2;                                        // OPPOSITE_KL, :138-142
        kl_i = (cc - 1.0 + dd - log(cc)) / 2.0;
      }
      const int cnt = min(32, dA - i0);
      for (int j = 0; j < cnt; ++j) {
        logw += __shfl_sync(0xffffffffu, lw_i, j);
        dkl += __shfl_sync(0xffffffffu, kl_i, j);
      }
    }
    const double rho = exp(logw > 7.0 ? 7.0 : (logw < -7.0 ? -7.0 : logw));             // :648-653
    // isFarPolicy takes Fval arguments (Episode.h:28-33)
    const float W32 = (float)rho, C32 = (float)cmax, I32 = (float)cinv;
    const bool offW = (W32 > C32) || (W32 < I32);
    const bool isFar = (C32 > 1.0f) && offW;
    const double O0 = (double)net_out(net, act, TB, 0, s);
    const double Vval = net2v(O0);
    const double Aval = 0.0;                                                           // Zero_advantage.h:39-42
    const float qret = ld_cg(rp.Q + row);
    const double A_RET = (double)qret - Vval, deltaQ = A_RET - Aval;
    const double Ver = fmin(1.0, rho) * deltaQ;
    const double pgfac = A_RET * fmin(cmax, rho);
    if (lane == 0) {
      const double g0 = isFar ? 0.0 : Ver * beta * vdiff(O0);
      err[(Lo.actOff + 0) * TB + s] = (float)g0;
      a.lastG[(size_t)b * net.nOut + 0] = (float)g0;
      a.lastO[(size_t)b * net.nOut + 0] = (float)O0;
    }
    for (int i = lane; i < dA; i += 32) {
      const double m = (double)net_out(net, act, TB, 1 + i, s);
      const double sraw = (double)net_out(net, act, TB, 1 + dA + i, s);
      const double av = (double)ld_cg(rp.A + row * dA + i);
      const double mm = (double)ld_cg(rp.MU + row * 2 * dA + i);
      const double ms = (double)ld_cg(rp.MU + row * 2 * dA + dA + i);
      const double root = sqrt(1.0 + sraw * sraw);
      const double stdev = (sraw + root) / 2.0, inv = 1.0 / stdev;
      const double dpos = (1.0 + sraw / root) / 2.0;                                    // SoftPlus::_evalDiff
      const bool bnd = hp.bounded[i] != 0;
      const double MAXM = 8.31776613503286;
      const double cm = bnd ? (m > MAXM ? MAXM : (m < -MAXM ? -MAXM : m)) : m;
      // penalG = KLDivGradient(MU, -1): gradKLdiv, OPPOSITE_KL branch (Continuous_policy.h:154-170)
      const double invVarMu = 1.0 / (ms * ms);
      const double kg_mean = -1.0 * ((m - mm) * invVarMu);
      const double kg_std = (dpos * -1.0) * ((invVarMu - inv * inv) * stdev);
      // polG = policyGradient(ACT, A_RET*min(Cmax,rho)): gradLogP (:145-152 / :300-316)
      const double u = (av - cm) * inv;
      const double dLogPdMean = bnd ? (av - m) * inv * inv : u * inv;
      const double dLogPdStdv = (u * u - 1.0) * inv;
      double pg_mean = pgfac * dLogPdMean;
      if (bnd && ((m >= MAXM && pg_mean > 0.0) || (m <= -MAXM && pg_mean < 0.0))) pg_mean = 0.0;
      double pg_std = (dpos * pgfac) * dLogPdStdv;
      if (isFar) { pg_mean = 0.0; pg_std = 0.0; }
      const double g_mean = beta * pg_mean + (1.0 - beta) * kg_mean;                     // penalizeReFER
      const double g_std = beta * pg_std + (1.0 - beta) * kg_std;
      err[(Lo.actOff + 1 + i) * TB + s] = (float)g_mean;
      err[(Lp.actOff + i) * TB + s] = (float)g_std;
      a.lastG[(size_t)b * net.nOut + 1 + i] = (float)g_mean;
      a.lastG[(size_t)b * net.nOut + 1 + dA + i] = (float)g_std;
      a.lastO[(size_t)b * net.nOut + 1 + i] = (float)m;
      a.lastO[(size_t)b * net.nOut + 1 + dA + i] = (float)sraw;
    }
    if (lane == 0) {
      // write-back (RACER_train.cpp:59-60) + record for the aggregate updates (Episode.h:112-145)
      SampleRec r;
      r.slot = info[TB + s]; r.hasNext = info[2 * TB + s];
      r.qNextOld = 0.f; r.qNextNew = 0.f;
      if (r.hasNext) {
        const float vn = vnext[s];
        r.qNextOld = ld_cg(rp.ADV + row + 1) + ld_cg(rp.V + row + 1);
        r.qNextNew = vn;
        rp.V[row + 1] = vn; rp.ADV[row + 1] = vn - vn;
      }
      const float E = (float)deltaQ, D = (float)dkl;
      const float oldRho = ld_cg(rp.RHO + row), oldKL = ld_cg(rp.KL + row), oldE = ld_cg(rp.DELTA + row);
      const bool wasOff = (oldRho > C32) || (oldRho < I32);
      r.dKL = D - oldKL;
      r.dFar = (float)offW - (float)wasOff;
      r.farDelta = (C32 > 1.0f) ? ((int)offW - (int)wasOff) : 0;
      r.dE2 = E * E - oldE * oldE;
      r.absE = fabsf(E);
      const float Vf = (float)Vval, Qf = (float)(Aval + Vval);
      r.qOld = ld_cg(rp.ADV + row) + ld_cg(rp.V + row);
      r.qNew = Qf;
      r.pad = 0;
      rp.DELTA[row] = E; rp.KL[row] = D; rp.RHO[row] = W32;
      rp.V[row] = Vf; rp.ADV[row] = Qf - Vf;
      a.rec[b] = r;
    }
  }
  __syncthreads();

  // ---- backward: Network::backProp, layers last to first (Network.h:216-226) ----
  for (int l = net.nLayers - 1; l >= 1; --l) {
    const LayerDesc& L = net.L[l];
    float* e = err + L.actOff * TB;
    if (L.kind == kDenseTanh || L.kind == kDenseLinear) {
      if (L.kind == kDenseTanh) {   // deltas *= 1 - Y^2 (Layer_Base.h:103-109)
        const float* y = act + L.actOff * TB;
        for (int idx = tid; idx < L.size * TB; idx += kThreads) e[idx] = e[idx] * (1.0f - y[idx] * y[idx]);
        __syncthreads();
      }
      if (L.wtOff >= 0)             // E_in += W * delta (Layers.h:131-145); skipped for the first layer
        dense_apply<TB>(a.WT + L.wtOff, L.ldt, L.size, L.nIn, nullptr, e, err + net.L[L.in].actOff * TB, red, 2);
    } else if (L.kind == kResidual) {   // ParametricResidualLayer::backward (Layers.h:363-393)
      float* e1 = err + net.L[l - 1].actOff * TB;
      float* e2 = err + net.L[l - 2].actOff * TB;
      for (int idx = tid; idx < L.size * TB; idx += kThreads) {
        e1[idx] = e[idx];
        e2[idx] += e[idx] * ld_cg(a.W + L.wOff + idx / TB);
      }
      __syncthreads();
    }
  }

  // ---- activations and deltas to the feature-major scratch read by P2 ----
  const int per = net.actPerSample;
  if (TB == 4 && b0 + TB <= a.B) {
    for (int f = tid; f < per; f += kThreads) {
      *reinterpret_cast<float4*>(a.actG + (size_t)f * a.Bpad + b0) = *reinterpret_cast<const float4*>(act + f * 4);
      *reinterpret_cast<float4*>(a.errG + (size_t)f * a.Bpad + b0) = *reinterpret_cast<const float4*>(err + f * 4);
    }
  } else {
    for (int idx = tid; idx < per * TB; idx += kThreads) {
      const int f = idx / TB, s = idx - f * TB;
      if (b0 + s < a.B) { a.actG[(size_t)f * a.Bpad + b0 + s] = act[idx]; a.errG[(size_t)f * a.Bpad + b0 + s] = err[idx]; }
    }
  }
}

// ------------------------------------------------------------------------------------------
// P2: weight gradient tile + Adam  (Layers.h:160-187, Optimizer.cpp:61-108,122-161)
// ------------------------------------------------------------------------------------------
constexpr int kBC = 256;          // batch chunk staged in shared memory
constexpr int kBCP = kBC + 4;     // padded row stride

struct AdamCoef { float eta, B1, B2, lambda, fac; };

__device__ __forceinline__ AdamCoef adam_coef(const Hyper& hp, const StepCtrl& c) {
  AdamCoef k;
  const long long nStep = c.adam_step + 1;                                  // prepare_update: nStep++ (Optimizer.cpp:119)
  const float etaf = (float)hp.learnrate;
  const float eta0 = (float)((double)etaf / (1.0 + (double)(float)(double)nStep * hp.epsAnneal));   // annealRate<nnReal>
  const float bt1 = (float)c.adam_bt1, bt2 = (float)c.adam_bt2;
  k.eta = eta0 * sqrtf(1.0f - bt2) / (1.0f - bt1);                          // struct Adam ctor (Optimizer.cpp:64-67)
  k.B1 = 0.9f; k.B2 = 0.999f;
  k.lambda = (float)hp.nnLambda;
  k.fac = (float)(1.0 / (double)hp.batchGlobal);
  return k;
}

__device__ __forceinline__ float adam_step(const AdamCoef& k, float G, float* w, float* m1, float* m2) {
  const float W = *w;
  const float penal = -W * k.lambda;                       // SMARTIES_ADAMW
  const float DW = k.fac * G;
  float M1 = k.B1 * (*m1) + (1.0f - k.B1) * DW;
  float M2 = k.B2 * (*m2) + (1.0f - k.B2) * DW * DW;
  const float numer = k.B1 * M1 + (1.0f - k.B1) * DW;      // SMARTIES_NESTEROV_ADAM
  M2 = M2 < M1 * M1 ? M1 * M1 : M2;                        // SMARTIES_SAFE_ADAM
  const float ret = numer / (FLT_EPSILON + sqrtf(M2));
  const float Wn = W + k.eta * (ret + penal);
  *m1 = M1; *m2 = M2; *w = Wn;
  return Wn;
}

__device__ void p2_tile(const StepArgs& a, const NetDesc& net, const Hyper& hp, const StepCtrl& c, const GradTile& t,
                        float* smem) {
  const int tid = threadIdx.x;
  float* As = smem;                 // [16][kBCP]
  float* Ds = smem + kTileK * kBCP; // [16][kBCP]
  const LayerDesc& L = net.L[t.layer];
  // operand rows: A = activation of the input layer (dense) / of layer ID-2 (residual); D = this layer's deltas
  const int K = t.kind == 0 ? L.nIn : L.size;
  const int aOff = t.kind == 0 ? net.L[L.in].actOff : (t.kind == 1 ? net.L[t.layer - 2].actOff : 0);
  const int dOff = L.actOff;
  const int N = L.size;
  float acc = 0.f, acc2 = 0.f;
  const int kk = tid >> 4, nn = tid & 15;
  const int warp = tid >> 5, lane = tid & 31;
  for (int bc = 0; bc < a.Bpad; bc += kBC) {
    if (bc) __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int q = tid + i * kThreads;
      const int r = q >> 6, c4 = (q & 63) * 4;
      float4 av = make_float4(0.f, 0.f, 0.f, 0.f), dv = av;
      if (t.kind == 0) {
        const int k = t.k0 + r;
        if (k < K) av = ld_cg4(a.actG + (size_t)(aOff + k) * a.Bpad + bc + c4);
        else if (k == K) av = make_float4(1.f, 1.f, 1.f, 1.f);            // bias row: db += delta
        const int n = t.n0 + r;
        if (n < N) dv = ld_cg4(a.errG + (size_t)(dOff + n) * a.Bpad + bc + c4);
      } else {
        const int n = t.n0 + r;
        if (n < N) {
          dv = ld_cg4(a.errG + (size_t)(dOff + n) * a.Bpad + bc + c4);
          if (t.kind == 1) av = ld_cg4(a.actG + (size_t)(aOff + n) * a.Bpad + bc + c4);
        }
      }
      *reinterpret_cast<float4*>(As + r * kBCP + c4) = av;
      *reinterpret_cast<float4*>(Ds + r * kBCP + c4) = dv;
    }
    __syncthreads();
    if (t.kind == 0) {
      const float* ap = As + kk * kBCP;
      const float* dp = Ds + nn * kBCP;
#pragma unroll 8
      for (int b4 = 0; b4 < kBC; b4 += 4) {
        const float4 x = *reinterpret_cast<const float4*>(ap + b4);
        const float4 d = *reinterpret_cast<const float4*>(dp + b4);
        acc = fmaf(x.x, d.x, acc); acc = fmaf(x.y, d.y, acc); acc = fmaf(x.z, d.z, acc); acc = fmaf(x.w, d.w, acc);
      }
    } else {
      // vector tiles: warp w reduces rows 2w, 2w+1 over the batch chunk
#pragma unroll
      for (int rr = 0; rr < 2; ++rr) {
        const int r = warp * 2 + rr;
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const float4 d = *reinterpret_cast<const float4*>(Ds + r * kBCP + lane * 8 + j * 4);
          const float4 x = *reinterpret_cast<const float4*>(As + r * kBCP + lane * 8 + j * 4);
          s1 += d.x; s1 += d.y; s1 += d.z; s1 += d.w;
          s2 = fmaf(d.x, x.x, s2); s2 = fmaf(d.y, x.y, s2); s2 = fmaf(d.z, x.z, s2); s2 = fmaf(d.w, x.w, s2);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { s1 += __shfl_xor_sync(0xffffffffu, s1, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o); }
        if (rr == 0) { if (lane == 0) { acc += s1; acc2 += s2; } }
        else if (lane == 1) { acc += s1; acc2 += s2; }
      }
    }
  }
  const AdamCoef ac = adam_coef(hp, c);
  if (t.kind == 0) {
    const int k = t.k0 + kk, n = t.n0 + nn;
    if (n < N && k <= K) {
      const int p = k < K ? L.wOff + k * L.ld + n : L.bOff + n;
      a.G[p] = acc;
      const float wn = adam_step(ac, acc, a.W + p, a.M1 + p, a.M2 + p);
      if (k < K && L.wtOff >= 0) a.WT[L.wtOff + n * L.ldt + k] = wn;
    }
  } else if (lane < 2) {
    const int n = t.n0 + warp * 2 + lane;
    if (n < N) {
      if (t.kind == 1) {
        const int pw = L.wOff + n, pb = L.bOff + n;
        a.G[pw] = acc2; a.G[pb] = acc;
        adam_step(ac, acc2, a.W + pw, a.M1 + pw, a.M2 + pw);
        adam_step(ac, acc, a.W + pb, a.M1 + pb, a.M2 + pb);
      } else {
        const int pb = L.bOff + n;
        a.G[pb] = acc;
        adam_step(ac, acc, a.W + pb, a.M1 + pb, a.M2 + pb);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// P3: replay statistics, Cmax annealing, ReF-ER beta update; writes ctrl[(step+1)&1]
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ double block_sum(double v, double* sh) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  double r = 0.0;
  for (int w = 0; w < kThreads / 32; ++w) r += sh[w];
  return r;
}
__device__ __forceinline__ float block_max(float v, float* sh) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  __syncthreads();
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  float r = sh[0];
  for (int w = 1; w < kThreads / 32; ++w) r = fmaxf(r, sh[w]);
  return r;
}

// Episode::updateCumulative_atomic / updateValues_atomic applied in sample order, one thread
// per run of samples that share an episode (samples are sorted, so runs are contiguous).
__device__ void apply_sample_records(const StepArgs& a) {
  const ReplayView& rp = a.rp;
  const int ME = rp.maxEpisodes;
  for (int b = threadIdx.x; b < a.B; b += kThreads) {
    const int slot = a.rec[b].slot;
    if (b > 0 && a.rec[b - 1].slot == slot) continue;
    float avgKL = rp.epAgg[AGG_KL * ME + slot], frac = rp.epAgg[AGG_FAR * ME + slot];
    float avgE2 = rp.epAgg[AGG_E2 * ME + slot], maxE = rp.epAgg[AGG_MAXE * ME + slot];
    float sQ2 = rp.epAgg[AGG_Q2 * ME + slot], sQ = rp.epAgg[AGG_Q1 * ME + slot];
    float maxQ = rp.epAgg[AGG_MAXQ * ME + slot], minQ = rp.epAgg[AGG_MINQ * ME + slot];
    const float invN = 1.0f / (float)rp.epLen[slot];
    for (int j = b; j < a.B; ++j) {
      const SampleRec r = a.rec[j];
      if (r.slot != slot) break;
      if (r.hasNext) {
        sQ2 += r.qNextNew * r.qNextNew - r.qNextOld * r.qNextOld; sQ += r.qNextNew - r.qNextOld;
        maxQ = fmaxf(maxQ, r.qNextNew); minQ = fminf(minQ, r.qNextNew);
      }
      avgKL += invN * r.dKL; frac += invN * r.dFar; avgE2 += invN * r.dE2; maxE = fmaxf(maxE, r.absE);
      sQ2 += r.qNew * r.qNew - r.qOld * r.qOld; sQ += r.qNew - r.qOld;
      maxQ = fmaxf(maxQ, r.qNew); minQ = fminf(minQ, r.qNew);
    }
    rp.epAgg[AGG_KL * ME + slot] = avgKL; rp.epAgg[AGG_FAR * ME + slot] = frac;
    rp.epAgg[AGG_E2 * ME + slot] = avgE2; rp.epAgg[AGG_MAXE * ME + slot] = maxE;
    rp.epAgg[AGG_Q2 * ME + slot] = sQ2; rp.epAgg[AGG_Q1 * ME + slot] = sQ;
    rp.epAgg[AGG_MAXQ * ME + slot] = maxQ; rp.epAgg[AGG_MINQ * ME + slot] = minQ;
  }
}

// updateTrainingStatistics reductions + updateCounters (MemoryProcessing.cpp:46-92,187-259).
// `sweep` != nullptr on the every-1000-steps recompute: Retrace error sums come from the sweep.
__device__ void stats_and_refer(const StepArgs& a, const Hyper& hp, const StepCtrl& c, StepCtrl& nx, int step,
                                const SweepSums* sweep, long long farExactOverride) {
  __shared__ double shd[kThreads / 32];
  __shared__ float shf[kThreads / 32];
  __shared__ unsigned long long shn[kThreads];
  const ReplayView& rp = a.rp;
  const int ME = rp.maxEpisodes, nEp = a.nEpisodes, tid = threadIdx.x;
  double sumDKL = 0, sumE2 = 0, sumQ2 = 0, sumQ1 = 0, sumR = 0;
  float maxAbsE = -1e9f, maxQ = -1e9f, negMinQ = -1e9f;
  for (int pos = tid; pos < nEp; pos += kThreads) {
    const int slot = rp.epOrder[pos];
    const float Ns = (float)rp.epLen[slot];
    sumDKL += (double)(Ns * rp.epAgg[AGG_KL * ME + slot]);
    sumE2 += (double)(Ns * rp.epAgg[AGG_E2 * ME + slot]);
    sumQ2 += (double)rp.epAgg[AGG_Q2 * ME + slot];
    sumQ1 += (double)rp.epAgg[AGG_Q1 * ME + slot];
    sumR += (double)rp.epAgg[AGG_TOTR * ME + slot];
    maxAbsE = fmaxf(maxAbsE, rp.epAgg[AGG_MAXE * ME + slot]);
    maxQ = fmaxf(maxQ, rp.epAgg[AGG_MAXQ * ME + slot]);
    negMinQ = fmaxf(negMinQ, -rp.epAgg[AGG_MINQ * ME + slot]);
  }
  // `Uint nOffPol += float` inside an OpenMP reduction with schedule(static,1): thread j of T
  // accumulates episodes j, j+T, ... with a float round trip per addition, partial counts are
  // then added as integers (MemoryProcessing.cpp:202-227).
  const int T = hp.referThreads;
  unsigned long long nOff = 0;
  if (tid < T) {
    for (int pos = tid; pos < nEp; pos += T) {
      const int slot = rp.epOrder[pos];
      const float x = (float)rp.epLen[slot] * rp.epAgg[AGG_FAR * ME + slot];
      nOff = (unsigned long long)(__ull2float_rn(nOff) + x);
    }
  }
  shn[tid] = nOff;
  sumDKL = block_sum(sumDKL, shd); sumE2 = block_sum(sumE2, shd); sumQ2 = block_sum(sumQ2, shd);
  sumQ1 = block_sum(sumQ1, shd); sumR = block_sum(sumR, shd);
  maxAbsE = block_max(maxAbsE, shf); maxQ = block_max(maxQ, shf); negMinQ = block_max(negMinQ, shf);
  __syncthreads();
  if (tid == 0) {
    unsigned long long tot = 0;
    for (int j = 0; j < min(T, kThreads); ++j) tot += shn[j];
    const long long gstep = c.grad_step + 1;                                   // nGradSteps()+1
    const double C = hp.clipImpWeight, E = hp.epsAnneal;
    const double cmax = 1.0 + C / (1.0 + (double)gstep * E);                  // annealRate
    const double cinv = 1.0 / cmax;
    if (cmax <= 1.0) tot = 0;
    const double nData = (double)a.nTransitions;
    const double maxN = (double)hp.maxTotObsGlobal, BS = (double)hp.batchGlobal;
    const double learnRefer = 0.1 * BS / fmax(maxN, nData);
    nx = c;
    nx.cmax = cmax; nx.cinv = cinv;
    nx.n_far_ref = (long long)tot;
    nx.max_abs_err = c.max_abs_err + learnRefer * ((double)maxAbsE - c.max_abs_err);
    nx.avg_kl = sumDKL / nData; nx.avg_sq_err = sumE2 / nData;
    nx.avg_return = sumR / (double)nEp; nx.avg_q = sumQ1 / nData;
    nx.max_q = (double)maxQ; nx.min_q = (double)(-negMinQ);
    const double var = sumQ2 / nData - nx.avg_q * nx.avg_q;
    nx.stdev_q = sqrt(fmax(var, 1e-16));
    if (sweep) { nx.cnt_ret = (c.cnt_ret < 0 ? 0 : c.cnt_ret) + sweep->nRet; nx.sum_ret_err = c.sum_ret_err + sweep->sumErr2; }
    long long exact = c.n_far_exact;
    if (farExactOverride >= 0) exact = farExactOverride;
    else for (int b = 0; b < a.B; ++b) exact += a.rec[b].farDelta;
    nx.n_far_exact = exact;
    // updateCounters: beta fixed-point iteration (MemoryProcessing.cpp:73-85); it runs after
    // applyEpisodesRemovalAlgo, so nStoredSteps() is the post-pruning count
    const double nPost = (double)(step == a.lastStep ? a.nTransitionsPost : a.nTransitions);
    const double fracOff = (double)(long long)tot / fmax(nPost, 1.0);
    const double lrB = 0.1 * BS / fmax(maxN, nPost);
    const double b0 = c.beta;
    const double mn = fmin(lrB, b0);
    nx.beta = fracOff > hp.penalTol ? (1.0 - mn) * b0 : (1.0 - mn) * b0 + fmin(lrB, 1.0 - b0);
    // Adam bookkeeping for the next step (Optimizer.cpp:155-158)
    nx.adam_step = c.adam_step + 1;
    double t1 = c.adam_bt1 * 0.9; if (t1 < (double)FLT_EPSILON) t1 = 0; nx.adam_bt1 = t1;
    double t2 = c.adam_bt2 * 0.999; if (t2 < (double)FLT_EPSILON) t2 = 0; nx.adam_bt2 = t2;
    nx.grad_step = c.grad_step + 1;
    if (a.statsOut) {
      smb200_step_stats& o = a.statsOut[step - a.stepBase];
      o.beta = nx.beta; o.cmax = nx.cmax; o.cinv = nx.cinv; o.n_far_policy = nx.n_far_ref; o.n_far_exact = nx.n_far_exact;
      o.avg_kl = nx.avg_kl; o.avg_sq_err = nx.avg_sq_err; o.max_abs_err = nx.max_abs_err; o.avg_return = nx.avg_return;
      o.stdev_q = nx.stdev_q; o.avg_q = nx.avg_q; o.max_q = nx.max_q; o.min_q = nx.min_q;
      o.sum_ret_err = nx.sum_ret_err; o.cnt_ret = nx.cnt_ret; o.grad_step = nx.grad_step;
    }
  }
}

__device__ void p3_stats(const StepArgs& a, const Hyper& hp, const StepCtrl& c, StepCtrl& nx, int step) {
  apply_sample_records(a);
  __threadfence_block();
  __syncthreads();
  stats_and_refer(a, hp, c, nx, step, nullptr, -1);
}

// On a sweep step the statistics phase is replaced by the sweep kernels + k_finalize_sweep; Adam
// bookkeeping must still advance, which stats_and_refer does there.

// ------------------------------------------------------------------------------------------
// kernels
// ------------------------------------------------------------------------------------------
using SmemHdr = DevDescs;

__device__ __forceinline__ float* load_descs(const StepArgs& a, unsigned char* raw, const NetDesc*& net, const Hyper*& hp) {
  SmemHdr* h = reinterpret_cast<SmemHdr*>(raw);
  const int nw = (int)(sizeof(SmemHdr) / 4);
  const int* src = reinterpret_cast<const int*>(a.descs);
  int* dst = reinterpret_cast<int*>(h);
  for (int i = threadIdx.x; i < nw; i += kThreads) dst[i] = src[i];
  __syncthreads();
  net = &h->net; hp = &h->hp;
  return reinterpret_cast<float*>(raw + ((sizeof(SmemHdr) + 15) / 16) * 16);
}

// StepCtrl is rewritten by other CTAs between steps: read it through L2
__device__ __forceinline__ void load_ctrl(StepCtrl& dst, const StepCtrl* src) {
  static_assert(sizeof(StepCtrl) % 8 == 0, "StepCtrl must be a multiple of 8 bytes");
  const long long* s = reinterpret_cast<const long long*>(src);
  long long* d = reinterpret_cast<long long*>(&dst);
  for (int i = 0; i < (int)(sizeof(StepCtrl) / 8); ++i) d[i] = __ldcg(s + i);
}

template <int TB>
__global__ void __launch_bounds__(kThreads) k_p1(StepArgs a, int step) {
  extern __shared__ __align__(16) unsigned char smraw[];
  const NetDesc* net; const Hyper* hp;
  float* smem = load_descs(a, smraw, net, hp);
  __shared__ StepCtrl c;
  if (threadIdx.x == 0) load_ctrl(c, &a.ctrl[step & 1]);
  __syncthreads();
  p1_tile<TB>(a, *net, *hp, c, step, blockIdx.x, smem);
}

__global__ void __launch_bounds__(kThreads) k_p2p3(StepArgs a, int step, int skipStats) {
  extern __shared__ __align__(16) unsigned char smraw[];
  const NetDesc* net; const Hyper* hp;
  float* smem = load_descs(a, smraw, net, hp);
  __shared__ StepCtrl c;
  if (threadIdx.x == 0) load_ctrl(c, &a.ctrl[step & 1]);
  __syncthreads();
  if ((int)blockIdx.x < a.nTiles) p2_tile(a, *net, *hp, c, a.tiles[blockIdx.x], smem);
  else if (!skipStats) p3_stats(a, *hp, c, a.ctrl[(step + 1) & 1], step);
}

template <int TB>
__global__ void __launch_bounds__(kThreads, 1) k_steps_persistent(StepArgs a, int step0, int nSteps, int skipStatsLast) {
  extern __shared__ __align__(16) unsigned char smraw[];
  const NetDesc* net; const Hyper* hp;
  float* smem = load_descs(a, smraw, net, hp);
  __shared__ StepCtrl c;
  const int nb = gridDim.x;
  const int nP1 = (a.B + TB - 1) / TB;
  for (int s = 0; s < nSteps; ++s) {
    const int step = step0 + s;
    if (threadIdx.x == 0) load_ctrl(c, &a.ctrl[step & 1]);
    __syncthreads();
    for (int t = blockIdx.x; t < nP1; t += nb) { p1_tile<TB>(a, *net, *hp, c, step, t, smem); __syncthreads(); }
    grid_barrier(a.barrier, nb);
    const bool skip = skipStatsLast && s == nSteps - 1;
    for (int t = blockIdx.x; t < a.nTiles + 1; t += nb) {
      if (t < a.nTiles) { p2_tile(a, *net, *hp, c, a.tiles[t], smem); __syncthreads(); }
      else if (!skip) p3_stats(a, *hp, c, a.ctrl[(step + 1) & 1], step);
    }
    grid_barrier(a.barrier, nb);
  }
}

// statistics after the every-1000-steps sweep (replaces P3 on that step)
__global__ void __launch_bounds__(kThreads) k_finalize_sweep(StepArgs a, int step, const SweepSums* sweep) {
  __shared__ Hyper hp; __shared__ StepCtrl c;
  if (threadIdx.x == 0) { hp = a.descs->hp; load_ctrl(c, &a.ctrl[step & 1]); }
  __syncthreads();
  stats_and_refer(a, hp, c, a.ctrl[(step + 1) & 1], step, sweep, sweep->nFarExact);
}

// actor-side / diagnostic forward on caller-provided raw states [n][dS] -> out[n][nOut]
template <int TB>
__global__ void __launch_bounds__(kThreads) k_forward(StepArgs a, const float* states, int n, float* out) {
  extern __shared__ __align__(16) unsigned char smraw[];
  const NetDesc* net; const Hyper* hp;
  float* smem = load_descs(a, smraw, net, hp);
  float* act = smem;
  float* red = act + 2 * net->actPerSample * TB;
  const int b0 = blockIdx.x * TB, dS = net->dS;
  for (int idx = threadIdx.x; idx < dS * TB; idx += kThreads) {
    const int k = idx / TB, s = idx - k * TB;
    act[idx] = b0 + s < n ? (states[(size_t)(b0 + s) * dS + k] - a.rp.stateMean[k]) * a.rp.stateScale[k] : 0.f;
  }
  __syncthreads();
  net_forward<TB>(a, *net, act, red);
  for (int idx = threadIdx.x; idx < net->nOut * TB; idx += kThreads) {
    const int j = idx / TB, s = idx - j * TB;
    if (b0 + s < n) out[(size_t)(b0 + s) * net->nOut + j] = net_out(*net, act, TB, j, s);
  }
}

// ------------------------------------------------------------------------------------------
// host launchers
// ------------------------------------------------------------------------------------------
static size_t p1_smem_bytes(const NetDesc& net, int TB) {
  size_t hdr = ((sizeof(SmemHdr) + 15) / 16) * 16;
  return hdr + sizeof(float) * (2 * (size_t)net.actPerSample * TB + (size_t)kThreads * TB) + sizeof(int) * 4 * TB + sizeof(float) * TB + 64;
}
static size_t p2_smem_bytes() {
  size_t hdr = ((sizeof(SmemHdr) + 15) / 16) * 16;
  return hdr + sizeof(float) * 2 * kTileK * kBCP;
}
size_t step_smem_bytes(const NetDesc& net, int TB) { return p1_smem_bytes(net, TB) > p2_smem_bytes() ? p1_smem_bytes(net, TB) : p2_smem_bytes(); }

static bool g_attr_set = false;
int step_kernels_prepare(const NetDesc& net) {
  const size_t s4 = step_smem_bytes(net, 4);
  if (s4 > 227 * 1024) { set_error_msg("network too wide for the shared-memory tile of the step kernel"); return -1; }
  SMB200_CUDA_CHECK(cudaFuncSetAttribute(k_p1<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s4));
  SMB200_CUDA_CHECK(cudaFuncSetAttribute(k_p2p3, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p2_smem_bytes()));
  SMB200_CUDA_CHECK(cudaFuncSetAttribute(k_steps_persistent<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s4));
  SMB200_CUDA_CHECK(cudaFuncSetAttribute(k_forward<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s4));
  g_attr_set = true;
  return 0;
}

int launch_step_two_kernels(const StepArgs& a, const NetDesc& net, int step, int skipStats, cudaStream_t st) {
  const int nP1 = (a.B + 3) / 4;
  k_p1<4><<<nP1, kThreads, p1_smem_bytes(net, 4), st>>>(a, step);
  k_p2p3<<<a.nTiles + 1, kThreads, p2_smem_bytes(), st>>>(a, step, skipStats);
  SMB200_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int persistent_grid(const StepArgs& a, const NetDesc& net, int numSMs) {
  int perSM = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, k_steps_persistent<4>, kThreads, step_smem_bytes(net, 4)) != cudaSuccess || perSM < 1) return 0;
  const int want = max((a.B + 3) / 4, a.nTiles + 1);
  return min(want, numSMs * perSM);
}

int launch_steps_persistent(const StepArgs& a, const NetDesc& net, int grid, int step0, int nSteps, int skipStatsLast, cudaStream_t st) {
  SMB200_CUDA_CHECK(cudaMemsetAsync(a.barrier, 0, sizeof(unsigned), st));
  StepArgs aa = a;
  void* args[] = {&aa, &step0, &nSteps, &skipStatsLast};
  SMB200_CUDA_CHECK(cudaLaunchCooperativeKernel((void*)k_steps_persistent<4>, dim3(grid), dim3(kThreads), args, step_smem_bytes(net, 4), st));
  return 0;
}

int launch_finalize_sweep(const StepArgs& a, int step, const SweepSums* sweep, cudaStream_t st) {
  k_finalize_sweep<<<1, kThreads, 0, st>>>(a, step, sweep);
  SMB200_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int launch_forward(const StepArgs& a, const NetDesc& net, const float* states, int n, float* out, cudaStream_t st) {
  k_forward<4><<<(n + 3) / 4, kThreads, p1_smem_bytes(net, 4), st>>>(a, states, n, out);
  SMB200_CUDA_CHECK(cudaGetLastError());
  return 0;
}

}  // namespace smb200
