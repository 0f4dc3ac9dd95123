// step_kernels.cu — the V-RACER learner step on one B200 (sm_100a).
//
// One learner step of the reference,
//     spawnTrainTasks(); processMemoryBuffer(); applyGradient();      (Learners/RACER.cpp:81-109)
// is three device phases separated by grid-wide dependencies:
//   P1  per tile of TB sampled transitions: the whole network (weight IMAGE, ~100 KB) is pulled
//       into shared memory with cp.async.bulk (TMA) right after the grid barrier; gather +
//       standardise states from the HBM replay rows, MLP forward, ReF-ER / Retrace loss and
//       output gradient in f64, write-back of V/Q/delta/KL/rho to the replay rows, input-
//       gradient backward; activations and deltas are left in a scratch for P2 (tile-major for
//       feed-forward nets, feature-major [feature][sample*Tc+k] for recurrent ones).  The next
//       step's inputs arrive by cp.async during P2; V(s_t+1) of truncated episodes is evaluated by
//       otherwise idle worker CTAs (next_state_helper).                 (RACER_train.cpp:12-67)
//   P2  per 16x16 tile of every weight matrix: dW = A^T * Delta contracted over the whole
//       mini-batch in batch order, fused with the gradient exchange between learner ranks (peer
//       memory over NVLink) and the reference's Adam variant in the epilogue (the per-thread
//       gradient buffers and their reduction, Parameters.h:66-103, vanish).  LSTM layers contract
//       on the tensor cores first (tcgen05.mma kind::tf32, 3xTF32, accumulator in TMEM:
//       tc_wgrad_item), the tiles then add the K-slices.
//   P3  one CTA, concurrent with P2: per-episode aggregate updates in sample order
//       (Episode.h:112-145), the per-step replay statistics, Cmax annealing and the ReF-ER
//       beta fixed-point update (MemoryProcessing.cpp:46-92,187-259).
// The phases run either as two kernels per step or inside one persistent cooperative kernel
// that loops over many steps with two grid barriers per step.
//
// Compiled with -fmad=false: every multiply-add that must be fused is written fmaf()
// explicitly; everything else rounds like the reference's x86-64 (no-FMA) build.
#include "step_kernels.cuh"

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <type_traits>
#include <vector>

namespace smb200 {

// threads per CTA of the step kernels.  The per-tile work is a chain of short dependent phases, so
// it is bound by instruction latency: more resident warps per scheduler hide it.
#ifndef SMB200_STEP_THREADS
#define SMB200_STEP_THREADS 512
#endif
constexpr int kST = SMB200_STEP_THREADS;
constexpr int kSTlog2 = kST == 256 ? 8 : (kST == 512 ? 9 : 10);
static_assert(kST == 256 || kST == 512 || kST == 1024, "step kernels need 256, 512 or 1024 threads");
int step_threads() { return kST; }

// ------------------------------------------------------------------------------------------
// small helpers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float ld_cg(const float* p) { return __ldcg(p); }
__device__ __forceinline__ float4 ld_cg4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// Grid-wide barrier of the persistent kernel (all CTAs co-resident: cooperative launch).
// `counter` only grows and is zeroed by the host before each launch; every CTA keeps the
// running target locally so the arrival is a fire-and-forget reduction.
__device__ __forceinline__ void grid_barrier(unsigned* counter, unsigned& target, unsigned nblocks) {
  __syncthreads();
  if (threadIdx.x == 0) {
    target += nblocks;
    // release: orders the CTA's earlier writes (made visible to this thread by bar.sync) before the
    // arrival; acquire: orders the other CTAs' writes before everything after the second bar.sync
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
    while (ld_acquire(counter) < target) { }
  }
  __syncthreads();
}

// ---- system-scope flags in peer memory (NVLink): release store / acquire load ----
__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// bounded spin: a lost peer must not hang the GPU; the error flag is reported by the host
__device__ __forceinline__ void wait_stamp_sys(const unsigned* p, unsigned stamp, const CommView& cm) {
  const long long t0 = clock64();
  while ((int)(ld_acquire_sys(p) - stamp) < 0) {
    if (clock64() - t0 > cm.timeoutCycles) { *cm.error = 1; break; }
  }
}

__device__ __forceinline__ void st_volatile_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// Poll the stamped elements `p[q * stride]` (q = 0..N-1, q != me) in LOCAL memory until every
// peer's store of this step has landed.  All N-1 loads of a round are issued back to back, so a
// round costs one L2 latency, not N-1.  out[q] = low word of slot q.
__device__ __forceinline__ void wait_values_ll(const unsigned long long* p, size_t stride, int N, int me, unsigned stamp,
                                               const CommView& cm, unsigned (&out)[kMaxWorld]) {
  unsigned long long v[kMaxWorld];
  const long long t0 = clock64();
  while (true) {
#pragma unroll
    for (int q = 0; q < kMaxWorld; ++q)
      if (q < N && q != me) asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v[q]) : "l"(p + (size_t)q * stride) : "memory");
    bool ok = true;
#pragma unroll
    for (int q = 0; q < kMaxWorld; ++q)
      if (q < N && q != me) ok = ok && ((unsigned)(v[q] >> 32) == stamp);
    if (ok) break;
    if (clock64() - t0 > cm.timeoutCycles) { *cm.error = 1; break; }
  }
#pragma unroll
  for (int q = 0; q < kMaxWorld; ++q) out[q] = (q < N && q != me) ? (unsigned)(v[q] & 0xffffffffull) : 0u;
}
// ---- gradient exchange elements: plain 4-byte values, "not arrived yet" = the poison pattern ----
// Every slot is written once per use by exactly one peer and reset to the poison by its reader right after it was
// read; slots rotate over four steps, so the next writer of a slot is causally three exchanges behind the reset.
// 0xFFFFFFFF is a NaN no arithmetic produces (the default NaN is 0x7FFFFFFF); a gradient that really carried it would
// end in the bounded wait's error flag, not in a wrong sum.
constexpr unsigned kPoison = 0xFFFFFFFFu;
__device__ __forceinline__ void st_volatile_u32(unsigned* p, unsigned v) {
  asm volatile("st.volatile.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// Returns false if a peer's value did not arrive within the time-out (error flag set): the caller must not consume `out`.
__device__ __forceinline__ bool wait_values_poison(unsigned* p, size_t stride, int N, int me, const CommView& cm, unsigned (&out)[kMaxWorld]) {
  unsigned v[kMaxWorld];
  const long long t0 = clock64();
  bool arrived = true;
  while (true) {
#pragma unroll
    for (int q = 0; q < kMaxWorld; ++q)
      if (q < N && q != me) asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v[q]) : "l"(p + (size_t)q * stride) : "memory");
    bool ok = true;
#pragma unroll
    for (int q = 0; q < kMaxWorld; ++q)
      if (q < N && q != me) ok = ok && (v[q] != kPoison);
    if (ok) break;
    if (clock64() - t0 > cm.timeoutCycles) { *cm.error = 1; arrived = false; break; }
  }
#pragma unroll
  for (int q = 0; q < kMaxWorld; ++q) {
    out[q] = (q < N && q != me) ? v[q] : 0u;
    if (q < N && q != me) st_volatile_u32(p + (size_t)q * stride, kPoison);      // free the slot for its next use
  }
  return arrived;
}

__device__ __forceinline__ unsigned long long ll_pack(unsigned v, unsigned stamp) {
  return ((unsigned long long)stamp << 32) | (unsigned long long)v;
}

// ---- mbarrier + bulk async copy (TMA, cp.async.bulk -> SASS UBLKCP) ----
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, unsigned parity) {
  unsigned ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
  while (!mbar_try_wait(bar, parity)) { }
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// ---- cp.async (LDGSTS): global -> shared without a register round trip ----
__device__ __forceinline__ void cp_async16_cg(void* dst, const void* src) {     // bypasses L1: safe for data other CTAs rewrite
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async16_ca(void* dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async4_ca(void* dst, const void* src) {      // only for replay data that is immutable during a launch
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// Tanh::_eval (Network/Layers/Functions.h:103-112), f32
// (hardware exp2 / reciprocal: 2-ulp error on e and on the quotient, far inside the f32 tolerance)
__device__ __forceinline__ float tanh_ref(float x) {
  const float e = __expf(-2.0f * fabsf(x));
  const float y = __fdividef(1.0f - e, 1.0f + e);
  return x > 0.0f ? y : -y;
}

__device__ __forceinline__ float sigm_ref(float x);
// Hidden-layer function "nnFunc" (makeFunction, Network/Layers/Functions.h:643-668): value from the pre-activation, derivative
// from the OUTPUT (the reference's evalDiff(in, out) of Tanh and Sigm only reads `out`; SoftSign and HardSign read `in`, whose
// terms are functions of the output: 1 + |x| = 1 / (1 - |y|),  1 + x^2 = 1 / (1 - y^2);  ExpPlus y = log(1 + e^x): 1 / (1 + e^-x)
// = 1 - e^-y, also where safeExp clips at +-8;  SoftPlus y = (x + sqrt(1 + x^2)) / 2: x = y - 1 / (4y), (1 + x / sqrt(1 + x^2)) / 2
// = 4y^2 / (4y^2 + 1);  Exp: evalDiff(in, out) = out).
__device__ __forceinline__ float safe_exp_ref(float x) { return expf(fminf(8.0f, fmaxf(-8.0f, x))); }   // Utilities::safeExp (FunctionUtilities.h:50-54)
template <int F> __device__ __forceinline__ float act_eval_t(float x) {
  if (F == 0) return tanh_ref(x);                                      // Tanh::_eval (:103-112)
  if (F == 1) return __fdividef(x, 1.0f + fabsf(x));                   // SoftSign::_eval (:328-331)
  if (F == 2) return x * rsqrtf(1.0f + x * x);                         // HardSign::_eval (:220-223)
  if (F == 3) return sigm_ref(x);                                      // Sigm::_eval (:158-165)
  if (F == 4) return x > 0.0f ? x : 0.0f;                              // Relu::_eval (:415-418)
  if (F == 5) return x > 0.0f ? x : 0.1f * x;                          // LRelu::_eval, PRELU_FAC 0.1 (:16-18,461-464)
  if (F == 6) return logf(1.0f + safe_exp_ref(x));                     // ExpPlus::_eval (:507-510)
  if (F == 7) return (x + sqrtf(1.0f + x * x)) * 0.5f;                 // SoftPlus::_eval (:552-555)
  if (F == 8) return safe_exp_ref(x);                                  // Exp::_eval (:604-607)
  return x;                                                            // Linear as the hidden-layer function (:66-69)
}
template <int F> __device__ __forceinline__ float act_diff_t(float y) {
  if (F == 0) return 1.0f - y * y;                                     // Tanh::_evalDiff
  if (F == 1) { const float t = 1.0f - fabsf(y); return t * t; }       // SoftSign: 1 / (1 + |x|)^2
  if (F == 2) { const float t = 1.0f - y * y; return t * sqrtf(t); }   // HardSign: 1 / (1 + x^2)^(3/2)
  if (F == 3) return y * (1.0f - y);                                   // Sigm::_evalDiff(in, out)
  if (F == 4) return y > 0.0f ? 1.0f : 0.0f;                           // Relu: in > 0 <=> out > 0
  if (F == 5) return y > 0.0f ? 1.0f : 0.1f;                           // LRelu
  if (F == 6) return -expm1f(-y);                                      // ExpPlus: 1 / (1 + safeExp(-x)) (:515-518)
  if (F == 7) { const float q = 4.0f * y * y; return __fdividef(q, q + 1.0f); }   // SoftPlus (:560-563)
  if (F == 8) return y;                                                // Exp::_evalDiff(in, out) = out (:614-617)
  return 1.0f;                                                         // Linear
}
// run `body(std::integral_constant<int, F>)` with F = the runtime function id: the selection happens once per call site, the
// element loops inside `body` are branch-free
template <class Body> __device__ __forceinline__ void act_dispatch(int f, Body&& body) {
  switch (f) {
    case 0: body(std::integral_constant<int, 0>{}); break;
    case 1: body(std::integral_constant<int, 1>{}); break;
    case 2: body(std::integral_constant<int, 2>{}); break;
    case 3: body(std::integral_constant<int, 3>{}); break;
    case 4: body(std::integral_constant<int, 4>{}); break;
    case 5: body(std::integral_constant<int, 5>{}); break;
    case 6: body(std::integral_constant<int, 6>{}); break;
    case 7: body(std::integral_constant<int, 7>{}); break;
    case 8: body(std::integral_constant<int, 8>{}); break;
    default: body(std::integral_constant<int, 9>{}); break;
  }
}
__device__ __forceinline__ float act_eval(int f, float x) {
  float y = 0.f;
  act_dispatch(f, [&](auto F) { y = act_eval_t<decltype(F)::value>(x); });
  return y;
}
__device__ __forceinline__ float act_diff(int f, float y) {
  float d = 0.f;
  act_dispatch(f, [&](auto F) { d = act_diff_t<decltype(F)::value>(y); });
  return d;
}

// scaleNet2V / scaleVdiff (Learners/RACER_common.cpp:23-32), f64
__host__ __device__ __forceinline__ double net2v(double x) {
  return x > 0 ? 100.0 * (x + 51.0) - 100.0 * sqrt(2601.0 + 100.0 * x)
               : 100.0 * (x - 51.0) + 100.0 * sqrt(2601.0 - 100.0 * x);
}
__host__ __device__ __forceinline__ double vdiff(double x) {
  return x > 0 ? 100.0 - 5000.0 / sqrt(2601.0 + 100.0 * x) : 100.0 - 5000.0 / sqrt(2601.0 - 100.0 * x);
}

// RACER<Discrete_advantage, Discrete_policy, Uint>::Train for ONE sample with K action options (Learners/RACER_train.cpp:12-67;
// Math/Discrete_policy.h:64-167: SoftPlus of the pre-activations, normalised; importance weight without clipping; KL
// divergence and its gradient; policy gradient; Math/Discrete_advantage.h:44-75: advantage centred with the policy's
// expectation), f64.  O = [V | advantages(K) | policy pre-activations(K)], act = stored action message (label + 0.1,
// Core/StateAction.h:304-341), mu = behaviour probabilities.  g receives the 1 + 2K output gradients; out = {rho, dkl,
// isFar, V, A, deltaQ}.  GROUNDWORK for SURVEY.md §8 row f4: no kernel calls this yet; its host build is pinned to the
// oracle (itself pinned to the reference golden racer_discrete) by tests/test_host_replay.py.
constexpr int kMaxOptions = 64;
__host__ __device__ inline void discrete_sample_loss(int K, const float* O, float act, const float* mu, float qret, double beta,
                                                     double cmax, double cinv, double* g, double* out) {
  double unnorm[kMaxOptions], probs[kMaxOptions], dpos[kMaxOptions];
  const int opt = (int)floor((double)act);                                  // actionMessage2label
  const float* adv = O + 1; const float* raw = O + 1 + K;
  double norm = 0.0;
  for (int j = 0; j < K; ++j) {
    const double x = (double)raw[j], root = sqrt(1.0 + x * x);
    unnorm[j] = (x + root) / 2.0;                                           // SoftPlus::_eval (Functions.h:552-555)
    dpos[j] = (1.0 + x / root) / 2.0;                                       // SoftPlus::_evalDiff
    norm = norm + unnorm[j];
  }
  norm = norm > 2.220446049250313e-16 ? norm : 2.220446049250313e-16;       // max(norm, eps<Real>)
  double dkl = 0.0, expA = 0.0;
  for (int j = 0; j < K; ++j) probs[j] = unnorm[j] / norm;
  for (int j = 0; j < K; ++j) dkl = dkl + probs[j] * log(probs[j] / (double)mu[j]);          // KLDivergence (:129-133)
  for (int j = 0; j < K; ++j) expA = expA + probs[j] * (double)adv[j];
  const double rho = probs[opt] / (double)mu[opt];                          // importanceWeight (:87-94)
  const float W32 = (float)rho, C32 = (float)cmax, I32 = (float)cinv;       // isFarPolicy takes Fval arguments (Episode.h:28-33)
  const bool isFar = (C32 > 1.0f) && ((W32 > C32) || (W32 < I32));
  const double Aval = (double)adv[opt] - expA;
  const double O0 = (double)O[0];
  const double V = net2v(O0);
  const double a_ret = (double)qret - V, dq = a_ret - Aval;
  const double rmin1 = rho < 1.0 ? rho : 1.0, rminC = rho < cmax ? rho : cmax;
  g[0] = isFar ? 0.0 : rmin1 * dq * beta * vdiff(O0);
  // penalG = KLDivGradient(MU, -1) (:158-167); polG = policyGradient(ACT, A_RET * min(Cmax, rho)) (:139-147)
  for (int i = 0; i < K; ++i) g[1 + K + i] = 0.0;
  for (int j = 0; j < K; ++j) {
    const double tmp = -1.0 * (1.0 + log(probs[j] / (double)mu[j])) / norm;
    for (int i = 0; i < K; ++i) g[1 + K + i] = g[1 + K + i] + tmp * ((i == j ? 1.0 : 0.0) - probs[j]);
  }
  const double fac = a_ret * rminC;
  const double err = isFar ? 0.0 : beta * rminC * dq;                       // ADV.grad(act, isFar ? 0 : beta * Aer) (:53-61)
  for (int i = 0; i < K; ++i) {
    const double penal = g[1 + K + i] * dpos[i];
    double pol = ((i == opt ? fac / unnorm[opt] : 0.0) - fac / norm) * dpos[i];
    if (isFar) pol = 0.0;
    g[1 + K + i] = beta * pol + (1.0 - beta) * penal;                       // penalizeReFER (FunctionUtilities.h:221-228)
    g[1 + i] = err * ((i == opt ? 1.0 : 0.0) - probs[i]);
  }
  out[0] = rho; out[1] = dkl; out[2] = isFar ? 1.0 : 0.0; out[3] = V; out[4] = Aval; out[5] = dq;
}

// StepCtrl is rewritten by other CTAs between steps: read it through L2
__device__ __forceinline__ void load_ctrl(StepCtrl& dst, const StepCtrl* src) {
  static_assert(sizeof(StepCtrl) % 8 == 0, "StepCtrl must be a multiple of 8 bytes");
  const long long* s = reinterpret_cast<const long long*>(src);
  long long* d = reinterpret_cast<long long*>(&dst);
  for (int i = 0; i < (int)(sizeof(StepCtrl) / 8); ++i) d[i] = __ldcg(s + i);
}

// phase timestamps exist only in the profiling flavour of the library (-DSMB200_MARKERS, libsmarties_b200_prof.so)
#ifdef SMB200_MARKERS
#define DBG_T(a, step, m) do { if ((a).dbgT && threadIdx.x == 0) \
  (a).dbgT[((size_t)((step) - (a).stepBase) * gridDim.x + blockIdx.x) * 48 + (m)] = clock64(); } while (0)
#define DBG_TW(a, step, m, w) do { if ((a) && (a)->dbgT && threadIdx.x == (w) * 32) \
  (a)->dbgT[((size_t)((step) - (a)->stepBase) * gridDim.x + blockIdx.x) * 48 + (m)] = clock64(); } while (0)
#else
#define DBG_T(a, step, m) do { } while (0)
#define DBG_TW(a, step, m, w) do { } while (0)
#endif

// ------------------------------------------------------------------------------------------
// Batched GEMV over a tile of TB samples.  Activations live in shared memory feature-major,
// x[k*TB + s].  The weight image `Wp` (shared memory; global memory only for networks that do
// not fit) stores W[k][n] with row stride ldp = roundUp4(N)+4, which makes both the forward
// access (lanes over n, one k) and the backward access (lanes over k, float4 along n) free of
// bank conflicts.
// ------------------------------------------------------------------------------------------
template <bool SM> __device__ __forceinline__ float ldw(const float* p) { return SM ? *p : __ldcg(p); }
template <bool SM> __device__ __forceinline__ float4 ldw4(const float* p) {
  return SM ? *reinterpret_cast<const float4*>(p) : __ldcg(reinterpret_cast<const float4*>(p));
}

//   mode 0: y = b + x W        mode 1 + f: y = act_f(b + x W), f = NetDesc::func       (BaseLayer::forward, Layer_Base.h:64-95)
// If yres != nullptr the ParametricResidualLayer that follows this layer is evaluated in the same
// epilogue: yres = y + (x * resW + resB)   (Layers.h:347-361; x = output of layer ID-2 = this layer's input).
template <int TB, bool SM>
__device__ __forceinline__ void dense_fwd(const float* Wp, int ldp, int K, int N, const float* bias, const float* x, float* y,
                                          float* red, int mode, int shift, const float* resW = nullptr, const float* resB = nullptr,
                                          float* yres = nullptr) {
  static_assert(TB == 4, "activations are read as float4 over the 4 samples of the tile");
  const int tid = threadIdx.x;
  for (int n0 = 0; n0 < N; n0 += kST) {
    const int nc = min(kST, N - n0);
    const int NR = 1 << shift;                       // thread-group width (host: power of two >= min(N, threads), >= 8)
    const int G = kST >> shift;
    const int g = tid >> shift, nl = tid & (NR - 1);
    const int Kc = (K + G - 1) >> (kSTlog2 - shift);
    const int kb = g * Kc, ke = min(K, kb + Kc);
    float acc[TB];
#pragma unroll
    for (int s = 0; s < TB; ++s) acc[s] = 0.f;
    if (nl < nc) {
      const float* w = Wp + (size_t)kb * ldp + n0 + nl;
      const float4* x4 = reinterpret_cast<const float4*>(x) + kb;   // TB == 4: one float4 per input feature
      int k = kb;
      for (; k + 8 <= ke; k += 8, w += 8 * ldp, x4 += 8) {
        float wv[8]; float4 xv[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) { wv[j] = ldw<SM>(w + j * ldp); xv[j] = x4[j]; }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          acc[0] = fmaf(xv[j].x, wv[j], acc[0]); acc[1] = fmaf(xv[j].y, wv[j], acc[1]);
          acc[2] = fmaf(xv[j].z, wv[j], acc[2]); acc[3] = fmaf(xv[j].w, wv[j], acc[3]);
        }
      }
      for (; k < ke; ++k, w += ldp, ++x4) {
        const float wv = ldw<SM>(w); const float4 xv = *x4;
        acc[0] = fmaf(xv.x, wv, acc[0]); acc[1] = fmaf(xv.y, wv, acc[1]);
        acc[2] = fmaf(xv.z, wv, acc[2]); acc[3] = fmaf(xv.w, wv, acc[3]);
      }
    }
    if (G > 1) {
      __syncthreads();
      *reinterpret_cast<float4*>(red + (g * NR + nl) * TB) = make_float4(acc[0], acc[1], acc[2], acc[3]);
      __syncthreads();
      for (int idx = tid; idx < nc * TB; idx += kST) {
        const int n2 = idx / TB, s = idx - n2 * TB;
        float v = 0.f;
#pragma unroll 4
        for (int gg = 0; gg < G; ++gg) v += red[(gg * NR + n2) * TB + s];
        const int n = n0 + n2;
        v += ldw<SM>(bias + n);
        v = mode == 1 ? tanh_ref(v) : (mode > 1 ? act_eval(mode - 1, v) : v);
        y[n * TB + s] = v;
        if (yres) yres[n * TB + s] = n < K ? v + (x[n * TB + s] * ldw<SM>(resW + n) + ldw<SM>(resB + n)) : v;
      }
    } else if (nl < nc) {
      const int n = n0 + nl;
      const float bv = ldw<SM>(bias + n);
      float rw = 0.f, rb = 0.f;
      if (yres && n < K) { rw = ldw<SM>(resW + n); rb = ldw<SM>(resB + n); }
#pragma unroll
      for (int s = 0; s < TB; ++s) {
        float v = acc[s] + bv;
        v = mode == 1 ? tanh_ref(v) : (mode > 1 ? act_eval(mode - 1, v) : v);
        y[n * TB + s] = v;
        if (yres) yres[n * TB + s] = n < K ? v + (x[n * TB + s] * rw + rb) : v;
      }
    }
  }
  __syncthreads();
}

// E_in[k] += sum_n W[k][n] * delta[n]   (Layer::backward input gradient, Layers.h:131-145)
template <int TB, bool SM>
__device__ __forceinline__ void dense_bwd_dx(const float* Wp, int ldp, int K, int N, const float* e, float* ein, float* red, int shift) {
  static_assert(TB == 4, "delta rows are read as float4 over the 4 samples of the tile");
  const int tid = threadIdx.x;
  const int N4 = (N + 3) >> 2;   // image rows and delta buffers are zero-padded to a multiple of 4
  for (int k0 = 0; k0 < K; k0 += kST) {
    const int kc = min(kST, K - k0);
    const int KR = 1 << shift;
    const int G = kST >> shift;
    const int g = tid >> shift, kl = tid & (KR - 1);
    const int Nc = (N4 + G - 1) >> (kSTlog2 - shift);
    const int nb = g * Nc, ne = min(N4, nb + Nc);
    float acc[TB];
#pragma unroll
    for (int s = 0; s < TB; ++s) acc[s] = 0.f;
    if (kl < kc) {
      const float* w = Wp + (size_t)(k0 + kl) * ldp;
      const float4* e4 = reinterpret_cast<const float4*>(e);
      int n4 = nb;
      for (; n4 + 2 <= ne; n4 += 2) {
        const float4 wa = ldw4<SM>(w + n4 * 4), wb = ldw4<SM>(w + n4 * 4 + 4);
        float4 d[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) d[j] = e4[n4 * 4 + j];
        const float wv[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          acc[0] = fmaf(wv[j], d[j].x, acc[0]); acc[1] = fmaf(wv[j], d[j].y, acc[1]);
          acc[2] = fmaf(wv[j], d[j].z, acc[2]); acc[3] = fmaf(wv[j], d[j].w, acc[3]);
        }
      }
      for (; n4 < ne; ++n4) {
        const float4 wa = ldw4<SM>(w + n4 * 4);
        const float4 d0 = e4[n4 * 4 + 0], d1 = e4[n4 * 4 + 1], d2 = e4[n4 * 4 + 2], d3 = e4[n4 * 4 + 3];
        acc[0] = fmaf(wa.x, d0.x, acc[0]); acc[1] = fmaf(wa.x, d0.y, acc[1]); acc[2] = fmaf(wa.x, d0.z, acc[2]); acc[3] = fmaf(wa.x, d0.w, acc[3]);
        acc[0] = fmaf(wa.y, d1.x, acc[0]); acc[1] = fmaf(wa.y, d1.y, acc[1]); acc[2] = fmaf(wa.y, d1.z, acc[2]); acc[3] = fmaf(wa.y, d1.w, acc[3]);
        acc[0] = fmaf(wa.z, d2.x, acc[0]); acc[1] = fmaf(wa.z, d2.y, acc[1]); acc[2] = fmaf(wa.z, d2.z, acc[2]); acc[3] = fmaf(wa.z, d2.w, acc[3]);
        acc[0] = fmaf(wa.w, d3.x, acc[0]); acc[1] = fmaf(wa.w, d3.y, acc[1]); acc[2] = fmaf(wa.w, d3.z, acc[2]); acc[3] = fmaf(wa.w, d3.w, acc[3]);
      }
    }
    if (G > 1) {
      __syncthreads();
      *reinterpret_cast<float4*>(red + (g * KR + kl) * TB) = make_float4(acc[0], acc[1], acc[2], acc[3]);
      __syncthreads();
      for (int idx = tid; idx < kc * TB; idx += kST) {
        const int k2 = idx / TB, s = idx - k2 * TB;
        float v = 0.f;
#pragma unroll 4
        for (int gg = 0; gg < G; ++gg) v += red[(gg * KR + k2) * TB + s];
        ein[(k0 + k2) * TB + s] += v;
      }
    } else if (kl < kc) {
#pragma unroll
      for (int s = 0; s < TB; ++s) ein[(k0 + kl) * TB + s] += acc[s];
    }
  }
  __syncthreads();
}

// Network::forward (Network/Network.h:101-113) for the tile; act[L.actOff*TB ...] per layer.
// A ParametricResidualLayer is evaluated in the epilogue of the dense layer below it; the
// ParamLayer (state-independent stdev parameters, Layers.h:510-521) is read straight from the
// weight image by its consumers.
template <int TB, bool SM>
__device__ void net_forward(const NetDesc& net, const float* Wp, float* act, float* red, uint64_t* bars, unsigned parity,
                            const StepArgs* dbg = nullptr, int step = 0) {
  for (int l = 1; l < net.nLayers; ++l) {
    if (dbg && l < 8) DBG_T(*dbg, step, 8 + l);
    const LayerDesc& L = net.L[l];
    float* y = act + L.actOff * TB;
    if (L.kind == kDenseTanh || L.kind == kDenseLinear) {
      const bool fuse = l + 1 < net.nLayers && net.L[l + 1].kind == kResidual;
      if (bars) { mbar_wait(&bars[l], parity); if (fuse) mbar_wait(&bars[l + 1], parity); }
      if (dbg && l < 5) DBG_T(*dbg, step, 26 + l);
      const LayerDesc& R = net.L[fuse ? l + 1 : l];
      dense_fwd<TB, SM>(Wp + L.imgW, L.ldp, L.nIn, L.size, Wp + L.imgB, act + net.L[L.in].actOff * TB, y, red,
                        L.kind == kDenseTanh ? 1 + net.func : 0, L.fwdShift, fuse ? Wp + R.imgW : nullptr, fuse ? Wp + R.imgB : nullptr,
                        fuse ? act + R.actOff * TB : nullptr);
#ifdef SMB200_ICACHE_PROBE   // does a warm instruction cache change the cost of a layer?  (idempotent repeat of the output layer)
      if (dbg && L.kind == kDenseLinear) {
        DBG_T(*dbg, step, 31);
        dense_fwd<TB, SM>(Wp + L.imgW, L.ldp, L.nIn, L.size, Wp + L.imgB, act + net.L[L.in].actOff * TB, y, red, 0, L.fwdShift);
        DBG_T(*dbg, step, 32);
      }
#endif
      if (fuse) ++l;
    } else if (L.kind == kResidual) {   // not preceded by a dense layer: cannot happen with Builder::addLayer
      if (bars) mbar_wait(&bars[l], parity);
      const float* y1 = act + net.L[l - 1].actOff * TB;
      const float* y2 = act + net.L[l - 2].actOff * TB;
      for (int idx = threadIdx.x; idx < L.size * TB; idx += kST) {
        const int n = idx / TB;
        y[idx] = y1[idx] + (y2[idx] * ldw<SM>(Wp + L.imgW + n) + ldw<SM>(Wp + L.imgB + n));
      }
      __syncthreads();
    } else if (bars) {
      mbar_wait(&bars[l], parity);      // ParamLayer: just make sure its slice has landed
    }
  }
}

// network output j of sample s: dense outputs from the activations, stdev parameters from the image
template <bool SM>
__device__ __forceinline__ float net_out(const NetDesc& net, const float* Wp, const float* act, int TB, int j, int s) {
  const LayerDesc& Lo = net.L[net.nLayers - 2];   // linear output layer
  const LayerDesc& Lp = net.L[net.nLayers - 1];   // param layer (stdev)
  return j < net.nOutDense ? act[(Lo.actOff + j) * TB + s] : ldw<SM>(Wp + Lp.imgB + j - net.nOutDense);
}

// ------------------------------------------------------------------------------------------
// ReF-ER / Retrace loss and output gradient of a tile of TB samples (RACER::Train,
// Learners/RACER_train.cpp:31-60), f64.  `act` / `err` hold the network outputs / receive the
// output gradient at [(layer.actOff + j) * TB + s].
// ------------------------------------------------------------------------------------------
struct LossIO {
  float* act; float* err; const int* info; const float* old; double* pair; double* samp; const float* vnext;
  int b0; double pa, pmm, pms; int p0s, p0i;   // behaviour policy / action of the thread's first (sample, component) pair
};

// LATE: the kernel may reach the loss before the statistics CTA of the previous step has published beta (cluster kernel):
// wait for it during stage 2 instead of stage 1 (see below).
// DISC: discrete action space (RACER<Discrete_advantage, Discrete_policy, Uint>): one thread per sample evaluates
// discrete_sample_loss, the write-back and the record are those of the continuous case.
template <int TB, bool SM, bool LATE = false, bool DISC = false>
__device__ __forceinline__ void loss_stages(const StepArgs& a, const NetDesc& net, const Hyper& hp, StepCtrl& c, int step, const float* Wp,
                                            const LossIO& io, bool fetchCtrl, const unsigned* readyFlag, unsigned readyTarget) {
  const int tid = threadIdx.x;
  float* act = io.act; float* err = io.err; const int* info = io.info; const float* old = io.old;
  double* pair = io.pair; double* samp = io.samp; const float* vnext = io.vnext;
  if (DISC) {
    const int K = net.discrete, b0d = io.b0;
    const LayerDesc& Lod = net.L[net.nLayers - 2];
    const bool keepd = step == a.lastStep || a.lastStep < 0;
    if (fetchCtrl && tid == kST - 1) {
      if (readyFlag) { while (ld_acquire(readyFlag) < readyTarget) { } }
      load_ctrl(c, &a.ctrl[step & 1]);
    }
    __syncthreads();
    if (tid < TB && info[3 * TB + tid]) {
      const int s = tid, b = b0d + s;
      const size_t row = info[s];
      const ReplayView& rpd = a.rp;
      float O[2 * kMaxOptions + 1], mu[kMaxOptions];
      double g[2 * kMaxOptions + 1], out[6];
      for (int j = 0; j < 1 + 2 * K; ++j) O[j] = act[(Lod.actOff + j) * TB + s];
      for (int j = 0; j < K; ++j) mu[j] = ld_cg(rpd.MU + row * K + j);
      const float actMsg = ld_cg(rpd.A + row);
      discrete_sample_loss(K, O, actMsg, mu, old[7 * TB + s], c.beta, c.cmax, c.cinv, g, out);
      for (int j = 0; j < 1 + 2 * K; ++j) {
        err[(Lod.actOff + j) * TB + s] = (float)g[j];
        if (keepd) { a.lastG[(size_t)b * net.nOut + j] = (float)g[j]; a.lastO[(size_t)b * net.nOut + j] = O[j]; }
      }
      const double rho = out[0], dkl = out[1], Vval = out[3], Aval = out[4], deltaQ = out[5];
      const float W32 = (float)rho, C32 = (float)c.cmax, I32 = (float)c.cinv;
      const bool offW = (W32 > C32) || (W32 < I32);
      SampleRec r;
      r.slot = info[TB + s]; r.hasNext = info[2 * TB + s];
      r.qNextOld = 0.f; r.qNextNew = 0.f;
      if (r.hasNext && vnext) {
        const float vn = vnext[s];
        r.qNextOld = old[6 * TB + s] + old[5 * TB + s];
        r.qNextNew = vn;
        rpd.V[row + 1] = vn; rpd.ADV[row + 1] = vn - vn;
      }
      const float E = (float)deltaQ, D = (float)dkl;
      const float oldRho = old[2 * TB + s], oldKL = old[3 * TB + s], oldE = old[4 * TB + s];
      const bool wasOff = (oldRho > C32) || (oldRho < I32);
      r.dKL = D - oldKL;
      r.dFar = (float)offW - (float)wasOff;
      r.farDelta = (C32 > 1.0f) ? ((int)offW - (int)wasOff) : 0;
      r.dE2 = E * E - oldE * oldE;
      r.absE = fabsf(E);
      const float Vf = (float)Vval, Qf = (float)(Aval + Vval);
      r.qOld = old[1 * TB + s] + old[0 * TB + s];
      r.qNew = Qf;
      r.pad = 0;
      rpd.DELTA[row] = E; rpd.KL[row] = D; rpd.RHO[row] = W32;
      rpd.V[row] = Vf; rpd.ADV[row] = Qf - Vf;
      if (vnext) a.rec[b] = r;
      else {
        *reinterpret_cast<int4*>(&a.rec[b].slot) = make_int4(r.slot, r.hasNext, r.farDelta, 0);
        *reinterpret_cast<float4*>(&a.rec[b].dKL) = make_float4(r.dKL, r.dFar, r.dE2, r.absE);
        *reinterpret_cast<float2*>(&a.rec[b].qOld) = make_float2(r.qOld, r.qNew);
      }
    }
    __syncthreads();
    return;
  }
  const int b0 = io.b0, p0 = tid, p0s = io.p0s, p0i = io.p0i;
  const bool keep = step == a.lastStep || a.lastStep < 0;      // smb200_get_last_batch only ever sees a launch's last step
  const double pa = io.pa, pmm = io.pmm, pms = io.pms;
  const ReplayView& rp = a.rp;
  const int dA = net.dA, nPair = TB * dA;
  const bool racer = hp.algo == 1;
  const int m0 = racer ? 2 + 2 * net.dA : 1;
  // ---- loss (RACER_train.cpp:31-60), f64.  Stage 1: one thread per (sample, action component);
  //      meanwhile warp 7 evaluates the value terms and fetches this step's ReF-ER scalars ----
  const LayerDesc& Lo = net.L[net.nLayers - 2];
  const LayerDesc& Lp = net.L[net.nLayers - 1];
  for (int p = tid; p < nPair; p += kST) {
    const int s = p == p0 ? p0s : p / dA, i = p == p0 ? p0i : p - s * dA;
    if (!info[3 * TB + s]) continue;
    double av = pa, mm = pmm, ms = pms;
    if (p != p0) {
      const size_t row = info[s];
      av = (double)ld_cg(rp.A + row * dA + i); mm = (double)ld_cg(rp.MU + row * 2 * dA + i); ms = (double)ld_cg(rp.MU + row * 2 * dA + dA + i);
    }
    const double m = (double)act[(Lo.actOff + m0 + i) * TB + s];
    const double sraw = (double)ldw<SM>(Wp + Lp.imgB + i);
    const double root = sqrt(1.0 + sraw * sraw);
    const double stdev = (sraw + root) / 2.0;                          // SoftPlus::_eval, Functions.h:552-555
    const double dpos = (1.0 + sraw / root) / 2.0;                     // SoftPlus::_evalDiff
    const double inv = 1.0 / stdev, invmu = 1.0 / ms;
    const bool bnd = hp.bounded[i] != 0;
    const double MAXM = 8.31776613503286;
    const double cm = bnd ? (m > MAXM ? MAXM : (m < -MAXM ? -MAXM : m)) : m;   // Continuous_policy.h:217-222
    const double fac0 = 9.1893853320467266954096885456237942e-01;
    double J = 1.0;
    if (bnd) { const double sq = tanh(av); J = fmax(1.0 - sq * sq, (double)FLT_MIN); }
    const double z1 = (av - cm) * inv, z2 = (av - mm) * invmu;
    const double lp_pi = -(z1 * z1) / 2.0 + log(bnd ? inv / J : inv) - fac0;       // :91-97 / :240-249
    const double lp_mu = -(z2 * z2) / 2.0 + log(bnd ? invmu / J : invmu) - fac0;
    const double r1 = stdev / ms, r2 = (m - mm) / ms;
    const double cc = r1 * r1, dd = r2 * r2;                                          // OPPOSITE_KL, :138-142
    // penalG = KLDivGradient(MU, -1): gradKLdiv, OPPOSITE_KL branch (:154-170)
    const double invVarMu = 1.0 / (ms * ms);
    // polG = policyGradient(ACT, f): gradLogP (:145-152 / :300-316), f applied in stage 3
    const double u = z1;
    pair[0 * nPair + p] = lp_pi - lp_mu;
    pair[1 * nPair + p] = (cc - 1.0 + dd - log(cc)) / 2.0;
    pair[2 * nPair + p] = -1.0 * ((m - mm) * invVarMu);                               // kg_mean
    pair[3 * nPair + p] = (dpos * -1.0) * ((invVarMu - inv * inv) * stdev);           // kg_std
    pair[4 * nPair + p] = bnd ? (av - m) * inv * inv : u * inv;                       // dLogPdMean
    pair[5 * nPair + p] = (u * u - 1.0) * inv;                                        // dLogPdStdv
    pair[6 * nPair + p] = dpos;
    if (racer) {   // Gaussian_advantage terms of this component (Gaus_advantage.h:73-126)
      const double p1r = (double)act[(Lo.actOff + 2 + i) * TB + s], p2r = (double)act[(Lo.actOff + 2 + dA + i) * TB + s];
      const double rt1 = sqrt(1.0 + p1r * p1r), rt2 = sqrt(1.0 + p2r * p2r);
      const double p1 = (p1r + rt1) / 2.0, p2 = (p2r + rt2) / 2.0;                 // PosDefFunction::_eval
      const double S = stdev * stdev;                                             // policy->getVariance(i)
      const double dmA = av - cm;                                                 // policy->getMean(i): clamped for bounded dims
      const double sq1 = sqrt(p1 / (p1 + S)), sq2 = sqrt(p2 / (p2 + S));
      const double x1 = dmA / p1, x2 = dmA / p2;
      pair[7 * nPair + p] = (dmA * dmA) / (av > cm ? p1 : p2);                    // diagInvMul term
      pair[8 * nPair + p] = sq1 / 2.0 + sq2 / 2.0;                                // coefMixRatio factor
      pair[9 * nPair + p] = av > cm ? x1 * x1 : -1.0;                             // ((a-m)/p1)^2 or "not on this side"
      pair[10 * nPair + p] = av < cm ? x2 * x2 : -1.0;
      pair[11 * nPair + p] = 2.0 / (sq1 + sq2);                                   // F
      pair[12 * nPair + p] = S / sqrt(p1 * ((p1 + S) * (p1 + S) * (p1 + S))) / 4.0;   // diff1
      pair[13 * nPair + p] = S / sqrt(p2 * ((p2 + S) * (p2 + S) * (p2 + S))) / 4.0;   // diff2
      pair[14 * nPair + p] = (1.0 + p1r / rt1) / 2.0;                             // PosDefFunction::_evalDiff
      pair[15 * nPair + p] = (1.0 + p2r / rt2) / 2.0;
    }
  }
  if (tid >= kST - 32 && tid < kST - 32 + TB) {     // value head: V = scaleNet2V(O[0]), dV/dO (RACER_common.cpp:23-32)
    const int s = tid - (kST - 32);
    const double O0 = (double)act[(Lo.actOff + 0) * TB + s];
    samp[2 * TB + s] = net2v(O0);
    samp[3 * TB + s] = vdiff(O0);
  }
  // This step's ReF-ER scalars.  First step of a launch (nothing to wait for): load them now.  Later steps: the statistics CTA of
  // the previous step may still be running — Cmax / Cinv only depend on the step number (annealing, same f64 expression as
  // stats_and_refer) and are evaluated locally, beta is waited for during stage 2 and first used in stage 3.
  const bool lateCtrl = LATE && fetchCtrl && readyFlag != nullptr && readyTarget > 0;
  if (fetchCtrl && !lateCtrl && tid == kST - 1) {
    if (readyFlag) { while (ld_acquire(readyFlag) < readyTarget) { } }
    load_ctrl(c, &a.ctrl[step & 1]);
  }
  __syncthreads();
  DBG_T(a, step, 8);
  if (lateCtrl && tid == kST - 1) {
    while (ld_acquire(readyFlag) < readyTarget) { }
    load_ctrl(c, &a.ctrl[step & 1]);
  }
  // Stage 2: one thread per sample — sums in component order like the reference, flags, value terms,
  // replay write-back (RACER_train.cpp:59-60) and the record for the aggregate updates.
  if (tid < TB && info[3 * TB + tid]) {
    const int s = tid, b = b0 + s;
    const size_t row = info[s];
    double cmax, cinv;
    if (lateCtrl) {                 // stats_and_refer of the previous step: gstep = c.grad_step + 1 = step
      cmax = 1.0 + hp.clipImpWeight / (1.0 + (double)(long long)step * hp.epsAnneal);
      cinv = 1.0 / cmax;
    } else { cmax = c.cmax; cinv = c.cinv; }
    double logw = 0.0, dkl = 0.0;
    for (int i = 0; i < dA; ++i) { logw += pair[0 * nPair + s * dA + i]; dkl += pair[1 * nPair + s * dA + i]; }
    const double rho = exp(logw > 7.0 ? 7.0 : (logw < -7.0 ? -7.0 : logw));           // :648-653
    // isFarPolicy takes Fval arguments (Episode.h:28-33)
    const float W32 = (float)rho, C32 = (float)cmax, I32 = (float)cinv;
    const bool offW = (W32 > C32) || (W32 < I32);
    const bool isFar = (C32 > 1.0f) && offW;
    const float O0f = act[(Lo.actOff + 0) * TB + s];
    const double Vval = samp[2 * TB + s];
    double Aval = 0.0;                                                                // Zero_advantage.h:39-42
    double orig = 0.0, ratio = 1.0, coef = 0.0, coefRaw = 0.0, rtc = 1.0;
    if (racer) {                                                                      // computeAdvantage, Gaus_advantage.h:73-78
      double shape = 0.0;
      for (int i = 0; i < dA; ++i) { shape += pair[7 * nPair + s * dA + i]; ratio *= pair[8 * nPair + s * dA + i]; }
      orig = exp(-shape / 2.0);
      coefRaw = (double)act[(Lo.actOff + 1) * TB + s];
      rtc = sqrt(1.0 + coefRaw * coefRaw);
      coef = (coefRaw + rtc) / 2.0;
      Aval = coef * (orig - ratio);
    }
    const double A_RET = (double)old[7 * TB + s] - Vval, deltaQ = A_RET - Aval;
    const double Ver = fmin(1.0, rho) * deltaQ;
    samp[8 * TB + s] = Ver;                         // g[0] = isFar ? 0 : Ver * beta * dV/dO is formed in stage 3 (beta)
    samp[0 * TB + s] = A_RET * fmin(cmax, rho);     // pgfac
    samp[1 * TB + s] = isFar ? 1.0 : 0.0;
    if (racer) {                                    // ADV.grad(act, isFar ? 0 : beta*Aer, gradient), Gaus_advantage.h:88-114
      const double Aer = fmin(cmax, rho) * deltaQ;
      const double expect = -ratio;
      samp[4 * TB + s] = orig * coef; samp[5 * TB + s] = expect; samp[6 * TB + s] = coef; samp[7 * TB + s] = Aer;
      samp[9 * TB + s] = orig; samp[10 * TB + s] = (1.0 + coefRaw / rtc) / 2.0;
      if (keep) a.lastO[(size_t)b * net.nOut + 1] = (float)coefRaw;
    }
    if (keep) a.lastO[(size_t)b * net.nOut + 0] = O0f;
    SampleRec r;
    r.slot = info[TB + s]; r.hasNext = info[2 * TB + s];
    r.qNextOld = 0.f; r.qNextNew = 0.f;
    if (r.hasNext && vnext) {
      const float vn = vnext[s];
      r.qNextOld = old[6 * TB + s] + old[5 * TB + s];
      r.qNextNew = vn;
      rp.V[row + 1] = vn; rp.ADV[row + 1] = vn - vn;
    }
    const float E = (float)deltaQ, D = (float)dkl;
    const float oldRho = old[2 * TB + s], oldKL = old[3 * TB + s], oldE = old[4 * TB + s];
    const bool wasOff = (oldRho > C32) || (oldRho < I32);
    r.dKL = D - oldKL;
    r.dFar = (float)offW - (float)wasOff;
    r.farDelta = (C32 > 1.0f) ? ((int)offW - (int)wasOff) : 0;
    r.dE2 = E * E - oldE * oldE;
    r.absE = fabsf(E);
    const float Vf = (float)Vval, Qf = (float)(Aval + Vval);
    r.qOld = old[1 * TB + s] + old[0 * TB + s];
    r.qNew = Qf;
    r.pad = 0;
    rp.DELTA[row] = E; rp.KL[row] = D; rp.RHO[row] = W32;
    rp.V[row] = Vf; rp.ADV[row] = Qf - Vf;
    if (vnext) a.rec[b] = r;
    else {      // qNextOld / qNextNew of this record belong to the helper CTA
      *reinterpret_cast<int4*>(&a.rec[b].slot) = make_int4(r.slot, r.hasNext, r.farDelta, 0);
      *reinterpret_cast<float4*>(&a.rec[b].dKL) = make_float4(r.dKL, r.dFar, r.dE2, r.absE);
      *reinterpret_cast<float2*>(&a.rec[b].qOld) = make_float2(r.qOld, r.qNew);
    }
  }
  __syncthreads();
  DBG_T(a, step, 16);
  // Stage 3: policy / penalty gradient of every pair (penalizeReFER, FunctionUtilities.h:221-228)
  {
    const double beta = c.beta;
    const double MAXM = 8.31776613503286;
    for (int p = tid; p < nPair; p += kST) {
      const int s = p == p0 ? p0s : p / dA, i = p == p0 ? p0i : p - s * dA;
      if (!info[3 * TB + s]) continue;
      const int b = b0 + s;
      const double pgfac = samp[0 * TB + s];
      const bool isFar = samp[1 * TB + s] != 0.0;
      if (i == 0) {                 // per-sample gradients that need beta: value head (RACER_train.cpp:46) and advantage coefficient
        const double g0 = isFar ? 0.0 : samp[8 * TB + s] * beta * samp[3 * TB + s];
        err[(Lo.actOff + 0) * TB + s] = (float)g0;
        if (keep) a.lastG[(size_t)b * net.nOut + 0] = (float)g0;
        if (racer) {
          const double errA0 = isFar ? 0.0 : beta * samp[7 * TB + s];
          const double gc = (samp[9 * TB + s] + samp[5 * TB + s]) * (errA0 * samp[10 * TB + s]);
          err[(Lo.actOff + 1) * TB + s] = (float)gc;
          if (keep) a.lastG[(size_t)b * net.nOut + 1] = (float)gc;
        }
      }
      const float mf = act[(Lo.actOff + m0 + i) * TB + s];
      const double m = (double)mf;
      double pg_mean = pgfac * pair[4 * nPair + p];
      if (hp.bounded[i] && ((m >= MAXM && pg_mean > 0.0) || (m <= -MAXM && pg_mean < 0.0))) pg_mean = 0.0;
      double pg_std = (pair[6 * nPair + p] * pgfac) * pair[5 * nPair + p];
      if (isFar) { pg_mean = 0.0; pg_std = 0.0; }
      const double g_mean = beta * pg_mean + (1.0 - beta) * pair[2 * nPair + p];
      const double g_std = beta * pg_std + (1.0 - beta) * pair[3 * nPair + p];
      err[(Lo.actOff + m0 + i) * TB + s] = (float)g_mean;
      err[(Lp.actOff + i) * TB + s] = (float)g_std;
      if (keep) a.lastG[(size_t)b * net.nOut + m0 + i] = (float)g_mean;
      if (keep) a.lastG[(size_t)b * net.nOut + m0 + dA + i] = (float)g_std;
      if (keep) a.lastO[(size_t)b * net.nOut + m0 + i] = mf;
      if (keep) a.lastO[(size_t)b * net.nOut + m0 + dA + i] = ldw<SM>(Wp + Lp.imgB + i);
      if (racer) {
        const double oc = samp[4 * TB + s], expect = samp[5 * TB + s], coef = samp[6 * TB + s];
        const double errA = isFar ? 0.0 : beta * samp[7 * TB + s];
        const double F = pair[11 * nPair + p];
        double g1 = pair[9 * nPair + p] >= 0.0 ? oc * pair[9 * nPair + p] / 2.0 : 0.0;
        double g2 = pair[10 * nPair + p] >= 0.0 ? oc * pair[10 * nPair + p] / 2.0 : 0.0;
        g1 += F * expect * coef * pair[12 * nPair + p];
        g2 += F * expect * coef * pair[13 * nPair + p];
        g1 *= errA * pair[14 * nPair + p];
        g2 *= errA * pair[15 * nPair + p];
        err[(Lo.actOff + 2 + i) * TB + s] = (float)g1;
        err[(Lo.actOff + 2 + dA + i) * TB + s] = (float)g2;
        if (keep) a.lastG[(size_t)b * net.nOut + 2 + i] = (float)g1;
        if (keep) a.lastG[(size_t)b * net.nOut + 2 + dA + i] = (float)g2;
        if (keep) a.lastO[(size_t)b * net.nOut + 2 + i] = act[(Lo.actOff + 2 + i) * TB + s];
        if (keep) a.lastO[(size_t)b * net.nOut + 2 + dA + i] = act[(Lo.actOff + 2 + dA + i) * TB + s];
      }
    }
  }
  __syncthreads();
  DBG_T(a, step, 3);
}

// ------------------------------------------------------------------------------------------
// shared-memory carve-up of the step kernels
// ------------------------------------------------------------------------------------------
struct SmemPlan {
  size_t img, act, err, red, info, old, pair, samp, tiles, bars, stage, chunks, total;
};
__host__ __device__ inline SmemPlan smem_plan(const NetDesc& net, int TB, bool imgInSmem) {
  SmemPlan p;
  size_t o = ((sizeof(DevDescs) + 15) / 16) * 16;
  p.img = o;   o += imgInSmem ? sizeof(float) * (size_t)net.imgFloats : 0;
  p.act = o;   o += sizeof(float) * (size_t)net.actPerSample * TB;
  p.err = o;   o += sizeof(float) * (size_t)net.actPerSample * TB;
  p.red = o;   o += sizeof(float) * (size_t)kST * TB;
  p.info = o;  o += sizeof(int) * 4 * TB;
  p.old = o;   o += sizeof(float) * 8 * TB;                          // old V/ADV/rho/KL/delta (+ next row) per sample
  o = (o + 15) / 16 * 16;
  p.pair = o;  o += sizeof(double) * 16 * (size_t)TB * net.dA;       // per (sample, action component) terms
  p.samp = o;  o += sizeof(double) * 12 * TB;                        // per sample scalars
  p.tiles = o; o += sizeof(float) * 2 * kTileK * (256 + 4);          // P2 operand tiles
  p.bars = o;  o += sizeof(uint64_t) * kMaxLayers;
  // inputs of the NEXT step, prefetched while the weight-gradient phase runs:
  // raw states [TB][dS], old values [8][TB], (a, mu_mean, mu_std) [3][TB*dA], info [4][TB]; then mean/scale [2][dS]
  // then the 16-byte chunks [8][TB][4] the old values arrive in (cp.async.cg moves 16 bytes)
  o = (o + 15) / 16 * 16;
  p.stage = o; o += sizeof(float) * ((size_t)TB * net.dS + 8 * TB + 3 * (size_t)TB * net.dA + 4 * TB + 2 * (size_t)net.dS);
  o = (o + 15) / 16 * 16;
  p.chunks = o; o += sizeof(float) * 8 * TB * 4;
  o = (o + 15) / 16 * 16;
  p.total = o;
  return p;
}

// ------------------------------------------------------------------------------------------
// weight image -> shared memory
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ int layer_img_begin(const NetDesc& net, int l) {
  const LayerDesc& L = net.L[l];
  return (L.kind == kParam) ? L.imgB : L.imgW;
}
__device__ __forceinline__ int layer_img_end(const NetDesc& net, int l) {
  return l + 1 < net.nLayers ? layer_img_begin(net, l + 1) : net.imgFloats;
}

// Called by all threads after the data dependency (grid barrier / kernel start) is satisfied.
__device__ __forceinline__ void load_weight_image(const StepArgs& a, const NetDesc& net, float* img, uint64_t* bars, int step = 0) {
  if (a.useTma) {
    // One issuing thread per layer, in different warps: a fence.proxy.async costs ~0.4 us and every bulk copy
    // ~0.17 us of issue time (measured, profiles/r1/phase_report_mlp_v4.txt) — serialised in one thread the last
    // layer's copy left 1.3 us after the barrier.  The fence orders the CTA's earlier generic-proxy accesses (shared
    // reads of the old image, made visible to this thread by bar.sync; the acquire of the grid barrier) before
    // this thread's async-proxy copies.
    if ((threadIdx.x & 31) == 0) {
      const int w = threadIdx.x >> 5;
      if (w + 1 < net.nLayers) {
        // state-space specific proxy fences (SASS FENCE.VIEW.ASYNC.G + MEMBAR.ALL.CTA, FENCE.VIEW.ASYNC.S); the
        // unqualified fence.proxy.async adds a MEMBAR.ALL.GPU (0.4-0.7 us measured) that the acquire of the grid
        // barrier has already paid for
        asm volatile("fence.proxy.async.global;\n\tfence.proxy.async.shared::cta;" ::: "memory");
        if (w == 0) DBG_T(a, step, 41);
        for (int l = 1 + w; l < net.nLayers; l += kST / 32) {
          const int b = layer_img_begin(net, l), e = layer_img_end(net, l);
          const unsigned bytes = (unsigned)(e - b) * 4u;
          mbar_expect_tx(&bars[l], bytes);
          bulk_g2s(img + b, a.Wimg + b, bytes, &bars[l]);
        }
      }
    }
  } else {
    const int n4 = net.imgFloats >> 2;
    for (int i = threadIdx.x; i < n4; i += kST)
      reinterpret_cast<float4*>(img)[i] = __ldcg(reinterpret_cast<const float4*>(a.Wimg) + i);
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------
// P1
// ------------------------------------------------------------------------------------------
// `c` is this CTA's shared-memory copy of ctrl[step&1].  If `fetchCtrl` is set it is (re)loaded here,
// as late as possible (just before the loss needs beta/Cmax), after waiting for the statistics CTA
// to have published the previous step's result (`readyFlag`, `readyTarget`; nullptr = already valid).
// `staged`: this tile's inputs (sample info, raw states, action / behaviour policy, old replay
// values) were prefetched into the shared-memory staging area during the previous step's P2.
struct Staging {
  float* S; float* old; float* pair; int* info; float* mean; float* scale;
};
template <int TB>
__device__ __forceinline__ Staging staging_view(const NetDesc& net, unsigned char* smraw, const SmemPlan& sp) {
  Staging g;
  float* f = reinterpret_cast<float*>(smraw + sp.stage);
  g.S = f; f += TB * net.dS;
  g.old = f; f += 8 * TB;
  g.pair = f; f += 3 * TB * net.dA;
  g.info = reinterpret_cast<int*>(f); f += 4 * TB;
  g.mean = f; f += net.dS;
  g.scale = f;
  return g;
}

template <int TB, bool SM, bool DISC = false>
__device__ void p1_tile(const StepArgs& a, const NetDesc& net, const Hyper& hp, StepCtrl& c, int step, int tile,
                        unsigned char* smraw, const SmemPlan& sp, unsigned parity, bool fetchCtrl,
                        const unsigned* readyFlag, unsigned readyTarget, bool staged, bool helped = false) {
  const int tid = threadIdx.x;
  float* img = reinterpret_cast<float*>(smraw + sp.img);
  const float* Wp = SM ? img : a.Wimg;
  float* act = reinterpret_cast<float*>(smraw + sp.act);     // [actPerSample][TB]
  float* err = reinterpret_cast<float*>(smraw + sp.err);     // [actPerSample][TB]
  float* red = reinterpret_cast<float*>(smraw + sp.red);
  const Staging stg = staging_view<TB>(net, smraw, sp);
  int* info = staged ? stg.info : reinterpret_cast<int*>(smraw + sp.info);   // row[TB], slot[TB], hasNext[TB], valid[TB]
  float* old = staged ? stg.old : reinterpret_cast<float*>(smraw + sp.old);  // [8][TB]: V, ADV, RHO, KL, DELTA, Vnext, ADVnext, Qret
  double* pair = reinterpret_cast<double*>(smraw + sp.pair); // [16][TB*dA]
  double* samp = reinterpret_cast<double*>(smraw + sp.samp); // [12][TB]
  const bool racer = hp.algo == 1;                           // RACER: Gaussian advantage head (Math/Gaus_advantage.h)
  const int m0 = racer ? 2 + 2 * net.dA : 1;                 // first policy-mean output (RACER_common.cpp:174-193,232-247)
  uint64_t* bars = (SM && a.useTma) ? reinterpret_cast<uint64_t*>(smraw + sp.bars) : nullptr;
  const int b0 = tile * TB;
  const ReplayView& rp = a.rp;
  const int dS = net.dS, dA = net.dA;
  const int nPair = TB * dA;

  if (!staged) {
    if (tid < TB) {
      const int b = b0 + tid;
      int row = 0, slot = 0, hn = 0, valid = 0;
      if (b < a.B) {
        const size_t j = (size_t)(step - a.stepBase) * a.B + b;
        row = a.sampRow[j];
        const int sf = a.sampSlot[j];
        slot = sf & 0x7fffffff; hn = (sf >> 31) & 1;          // Episode::isTruncated(t+1), resolved on the host
        valid = 1;
        // old per-transition values needed by the write-back, fetched early
        old[0 * TB + tid] = ld_cg(rp.V + row); old[1 * TB + tid] = ld_cg(rp.ADV + row);
        old[2 * TB + tid] = ld_cg(rp.RHO + row); old[3 * TB + tid] = ld_cg(rp.KL + row);
        old[4 * TB + tid] = ld_cg(rp.DELTA + row); old[7 * TB + tid] = ld_cg(rp.Q + row);
        if (hn) { old[5 * TB + tid] = ld_cg(rp.V + row + 1); old[6 * TB + tid] = ld_cg(rp.ADV + row + 1); }
      }
      info[tid] = row; info[TB + tid] = slot; info[2 * TB + tid] = hn; info[3 * TB + tid] = valid;
    }
    __syncthreads();
  }
  // `helped`: V(s_{t+1}) of truncated episodes is evaluated by an otherwise idle CTA (next_state_helper), which
  // also writes it back; this CTA then only flags the sample in its record
  int anyNext = 0;
#pragma unroll
  for (int s = 0; s < TB; ++s) anyNext |= info[2 * TB + s];
  if (helped) anyNext = 0;

  // V(s_{t+1}) of truncated episodes (RACER_train.cpp:23-27): rare, extra forward pass
  float* vnext = reinterpret_cast<float*>(samp + 11 * TB);  // samp[11][*] is not used by the loss stages
  if (anyNext) {
    for (int idx = tid; idx < dS * TB; idx += kST) {
      const int k = idx / TB, s = idx - k * TB;
      const size_t row = (size_t)info[s] + 1;
      act[idx] = info[2 * TB + s] ? (ld_cg(rp.S + row * dS + k) - ld_cg(rp.stateMean + k)) * ld_cg(rp.stateScale + k) : 0.f;
    }
    __syncthreads();
    net_forward<TB, SM>(net, Wp, act, red, bars, parity);
    if (tid < TB) vnext[tid] = (float)net2v((double)net_out<SM>(net, Wp, act, TB, 0, tid));
    __syncthreads();
  }

  // gather + standardise: (s - mean) * scale   (Episode.h:171-183)
  for (int idx = tid; idx < dS * TB; idx += kST) {
    const int k = idx / TB, s = idx - k * TB;
    float x = 0.f;
    if (info[3 * TB + s])
      x = staged ? (stg.S[s * dS + k] - stg.mean[k]) * stg.scale[k]
                 : (ld_cg(rp.S + (size_t)info[s] * dS + k) - ld_cg(rp.stateMean + k)) * ld_cg(rp.stateScale + k);
    act[idx] = x;
    if (info[3 * TB + s] && (step == a.lastStep || a.lastStep < 0)) a.lastX[(size_t)(b0 + s) * dS + k] = x;
  }
  for (int idx = tid; idx < net.actPerSample * TB; idx += kST) err[idx] = 0.f;   // clearErrors
  // behaviour policy and action of this thread's first (sample, component) pair
  double pa = 0, pmm = 0, pms = 1;
  const int p0 = tid;
  const int p0s = p0 < nPair ? p0 / dA : 0, p0i = p0 - p0s * dA;   // one integer division per tile, reused below
  if (!DISC && p0 < nPair) {
    const int s = p0s, i = p0i;
    if (info[3 * TB + s]) {
      if (staged) { pa = (double)stg.pair[p0]; pmm = (double)stg.pair[nPair + p0]; pms = (double)stg.pair[2 * nPair + p0]; }
      else {
        const size_t row = info[s];
        pa = (double)ld_cg(rp.A + row * dA + i);
        pmm = (double)ld_cg(rp.MU + row * 2 * dA + i);
        pms = (double)ld_cg(rp.MU + row * 2 * dA + dA + i);
      }
    }
  }
  __syncthreads();
  DBG_T(a, step, 1);
  net_forward<TB, SM>(net, Wp, act, red, bars, parity, &a, step);
  DBG_T(a, step, 2);

  {
    LossIO io{act, err, info, old, pair, samp, helped ? nullptr : vnext, b0, pa, pmm, pms, p0s, p0i};
    loss_stages<TB, SM, false, DISC>(a, net, hp, c, step, Wp, io, fetchCtrl, readyFlag, readyTarget);
  }

  // ---- backward: Network::backProp, layers last to first (Network.h:216-226).  A residual layer
  //      is folded into the dense layer below it: E(l) = E(res) * f'(Y_l), E(l-1) += E(res) * w_res ----
  const int func = net.func;          // "nnFunc" of the hidden layers (0 = Tanh: the expression of round 1, unchanged)
  for (int l = net.nLayers - 1; l >= 1; --l) {
    const LayerDesc& L = net.L[l];
    if (L.kind != kDenseTanh && L.kind != kDenseLinear) continue;
    if (l < 8) DBG_T(a, step, 16 + l);
    float* e = err + L.actOff * TB;
    float* ein = err + net.L[L.in].actOff * TB;
    const bool fuse = l + 1 < net.nLayers && net.L[l + 1].kind == kResidual;
    if (fuse) {                       // ParametricResidualLayer::backward (Layers.h:363-393) + deltas *= f' (Layer_Base.h:103-109)
      const LayerDesc& R = net.L[l + 1];
      const float* e3 = err + R.actOff * TB;
      const float* y = act + L.actOff * TB;
      for (int idx = tid; idx < L.size * TB; idx += kST) {
        const float d3 = e3[idx];
        e[idx] = L.kind == kDenseTanh ? d3 * (func == 0 ? 1.0f - y[idx] * y[idx] : act_diff(func, y[idx])) : d3;
        if (idx < L.nIn * TB) ein[idx] += d3 * ldw<SM>(Wp + R.imgW + idx / TB);
      }
      __syncthreads();
    } else if (L.kind == kDenseTanh) {
      const float* y = act + L.actOff * TB;
      for (int idx = tid; idx < L.size * TB; idx += kST) e[idx] = e[idx] * (func == 0 ? 1.0f - y[idx] * y[idx] : act_diff(func, y[idx]));
      __syncthreads();
    }
    if (L.needDx)                   // E_in += W * delta; skipped for the first layer (Approximator.cpp:145-169)
      dense_bwd_dx<TB, SM>(Wp + L.imgW, L.ldp, L.nIn, L.size, e, ein, red, L.bwdShift);
  }
  DBG_T(a, step, 4);

  // ---- activations and deltas to the feature-major scratch read by P2 ----
  //      TILE-major for feed-forward nets: scratch[(tile * per + feature) * 4 + sample] — the tile's block is one
  //      contiguous, fully coalesced 16-byte-per-thread store (the feature-major layout cost 32 sectors per warp
  //      store); P2 reads it back as 256-byte runs of 16 features (p2_tile) ----
  const int per = net.actPerSample;
  static_assert(TB == 4, "tile-major scratch layout assumes 4 samples per tile");
  float* actT = a.actG + (size_t)tile * per * 4;
  float* errT = a.errG + (size_t)tile * per * 4;
  if (b0 + TB <= a.B) {
    for (int f = tid; f < per; f += kST) {
      *reinterpret_cast<float4*>(actT + f * 4) = *reinterpret_cast<const float4*>(act + f * 4);
      *reinterpret_cast<float4*>(errT + f * 4) = *reinterpret_cast<const float4*>(err + f * 4);
    }
  } else {
    for (int idx = tid; idx < per * TB; idx += kST) {
      const int s = idx & 3;
      if (b0 + s < a.B) { actT[idx] = act[idx]; errT[idx] = err[idx]; }
    }
  }
}

// ------------------------------------------------------------------------------------------
// Recurrent networks (nnType LSTM): P1 of ONE sampled transition per CTA.
//
// The reference evaluates the sampled step t on the window [t - min(nnBPTTseq, t), t] starting
// from a zero recurrent state (MemoryBuffer.cpp:393-402, Approximator.h:129-139), places the
// output gradient at step t only and back-propagates through time layer by layer
// (Network::backProp, Network.h:155-193).  Here the forward pass is layer-major as well: the
// input contribution b + x_k Wx of every window step is one batched product, only h_{k-1} Wh and
// the gates run sequentially; the backward recurrence carries the state delta and the forget
// gate of step k+1 in registers and overwrites the stored gates by the gate deltas, which then
// leave for the P2 scratch as [feature][sample*Tc + k] rows.
// ------------------------------------------------------------------------------------------
static inline int seq_stride(int n) {            // >= n, == 4 (mod 32): float4-aligned and conflict-free for seq_store
  int s = (n + 3) / 4 * 4;
  while ((s & 31) != 4) s += 4;
  return s;
}

void seq_plan(const NetDesc& net, SeqPlan& p) {
  memset(&p, 0, sizeof(p));
  const int T = net.Tc + 1;
  p.T = T;
  int o = 0;
  for (int l = 0; l < net.nLayers; ++l) {
    const LayerDesc& L = net.L[l];
    if (L.kind == kInput || L.kind == kResidual || is_cell_layer(L.kind)) {
      const int n4 = (L.size + 3) / 4 * 4;
      // LSTM: [y | cell state | tanh(state)]; MGU: [y | h_prev * forget]
      p.yStride[l] = seq_stride(L.kind == kLSTM ? 3 * n4 : (L.kind == kMGU ? 2 * n4 : n4)); p.yOff[l] = o; o += T * p.yStride[l];
      if (is_cell_layer(L.kind)) { p.gStride[l] = seq_stride(cell_gates(L.kind) * L.size); p.gOff[l] = o; o += T * p.gStride[l]; }
      if (L.kind != kInput) { p.eStride[l] = seq_stride(L.size); p.eOff[l] = o; o += T * p.eStride[l]; }
    }
  }
  const int aps = (net.actPerSample + 3) / 4 * 4;
  p.actTop = o; o += aps;
  p.errTop = o; o += aps;
  p.red = o; o += 2 * kST;
  p.total = o;
}
int seq_workspace_floats(const NetDesc& net) { SeqPlan p; seq_plan(net, p); return p.total; }

struct SeqSmem { size_t img, ws, info, old, pair, samp, bars, total; };
__host__ __device__ inline SeqSmem smem_plan_seq(const NetDesc& net, bool imgInSmem) {
  SeqSmem p;
  size_t o = ((sizeof(DevDescs) + 15) / 16) * 16;
  p.img = o;  o += imgInSmem ? sizeof(float) * (size_t)net.imgFloats : 0;
  const size_t wsB = sizeof(float) * (size_t)net.seqFloats, tilesB = sizeof(float) * 2 * kTileK * (256 + 4);
  p.ws = o;   o += wsB > tilesB ? wsB : tilesB;          // the P2 operand tiles alias the sequence workspace
  p.info = o; o += sizeof(int) * 8;
  p.old = o;  o += sizeof(float) * 8;
  o = (o + 15) / 16 * 16;
  p.pair = o; o += sizeof(double) * 16 * (size_t)net.dA;
  p.samp = o; o += sizeof(double) * 12;
  p.bars = o; o += sizeof(uint64_t) * kMaxLayers;
  o = (o + 15) / 16 * 16;
  p.total = o;
  return p;
}

// Sigm::_eval with safeExp clipped at +-SMARTIES_EXP_CUT = 8 (Functions.h:158-165, FunctionUtilities.h:51-54)
__device__ __forceinline__ float sigm_ref(float x) {
  const float e = __expf(-fminf(fabsf(x), 8.0f));
  return x > 0.0f ? __fdividef(1.0f, 1.0f + e) : __fdividef(e, 1.0f + e);
}

// rows [rowOff, rowOff+nFeat) x columns [col0, col0+Tc) of a feature-major P2 scratch
//   <- src[(k - kShift)*stride + f] for window steps k < T1 (zero outside the window).
// A warp covers 4 features x 8 steps: conflict-free shared reads (stride == 4 mod 32), 32-byte global runs.
__device__ __forceinline__ void seq_store(float* dst, int Bpad, int rowOff, int nFeat, int col0, int Tc, int T1,
                                          const float* src, int stride, int kShift) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int fB = (nFeat + 3) >> 2, kB = (Tc + 7) >> 3;
  for (int blk = warp; blk < fB * kB; blk += kST / 32) {
    const int fb = blk / kB;
    const int f = fb * 4 + (lane & 3), k = (blk - fb * kB) * 8 + (lane >> 2);
    if (f < nFeat && k < Tc) {
      const int ks = k - kShift;
      dst[(size_t)(rowOff + f) * Bpad + col0 + k] = (k < T1 && ks >= 0) ? src[ks * stride + f] : 0.0f;
    }
  }
}

// LSTMLayer::forward for window steps [0, Tn) (Layers/Layer_LSTM.h:77-125)
template <bool SM>
__device__ void lstm_forward(const LayerDesc& L, const float* Wp, const float* in, int is, float* Gt, int gs, float* Y, int ys,
                             float* red, int Tn) {
  const int tid = threadIdx.x;
  const int nC = L.size, nI = L.nIn, N4 = 4 * nC, ldp = L.ldp, nC4 = (nC + 3) / 4 * 4;
  const float* Wx = Wp + L.imgW;
  const float* Wh = Wx + (size_t)nI * ldp;
  const float* bias = Wp + L.imgB;
  // (a) suminp = b + x_k Wx for every window step, four steps per thread
  const int nq = (Tn + 3) >> 2;
  for (int idx = tid; idx < N4 * nq; idx += kST) {
    const int kq = idx / N4, n = idx - kq * N4, k0 = kq * 4;
    const float bv = ldw<SM>(bias + n);
    float acc[4] = {bv, bv, bv, bv};
    const float* x[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) x[j] = in + (size_t)min(k0 + j, Tn - 1) * is;
    const float* w = Wx + n;
    int i = 0;
    for (; i + 4 <= nI; i += 4) {
      const float w0 = ldw<SM>(w + (size_t)(i + 0) * ldp), w1 = ldw<SM>(w + (size_t)(i + 1) * ldp);
      const float w2 = ldw<SM>(w + (size_t)(i + 2) * ldp), w3 = ldw<SM>(w + (size_t)(i + 3) * ldp);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 xv = *reinterpret_cast<const float4*>(x[j] + i);
        acc[j] = fmaf(xv.x, w0, acc[j]); acc[j] = fmaf(xv.y, w1, acc[j]); acc[j] = fmaf(xv.z, w2, acc[j]); acc[j] = fmaf(xv.w, w3, acc[j]);
      }
    }
    for (; i < nI; ++i) {
      const float wv = ldw<SM>(w + (size_t)i * ldp);
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[j] = fmaf(x[j][i], wv, acc[j]);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) if (k0 + j < Tn) Gt[(size_t)(k0 + j) * gs + n] = acc[j];
  }
  __syncthreads();
  // (b) the recurrence: suminp += h_{k-1} Wh, gates, cell state, output
  const int sh = L.fwdShift, NR = 1 << sh, G = kST >> sh;
  const int g = tid >> sh, nl = tid & (NR - 1);
  const int Kc = (((nC + G - 1) / G) + 3) / 4 * 4;
  const int kb = min(nC, g * Kc), ke = min(nC, kb + Kc);
  // Register-resident recurrent weights: when one pass covers all 4*nCells gate columns and this thread's K-slice is
  // exactly 32 rows (cfg3: 64 cells, 512 threads), its 32 weights stay in registers for the whole window instead of
  // 32 scalar shared-memory loads per step (the recurrence is bound by the shared-memory pipe).  Same FMA order.
  const bool regW = SM && N4 <= NR && (ke - kb) == 32 && nl < N4;
  float wreg[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) wreg[i] = regW ? ldw<SM>(Wh + nl + (size_t)(kb + i) * ldp) : 0.0f;
  for (int k = 0; k < Tn; ++k) {
    if (k > 0 && regW) {
      const float* hprev = Y + (size_t)(k - 1) * ys + kb;
      float acc = 0.0f;
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        const float4 hv = *reinterpret_cast<const float4*>(hprev + i);
        acc = fmaf(hv.x, wreg[i + 0], acc); acc = fmaf(hv.y, wreg[i + 1], acc);
        acc = fmaf(hv.z, wreg[i + 2], acc); acc = fmaf(hv.w, wreg[i + 3], acc);
      }
      if (G == 1) Gt[(size_t)k * gs + nl] += acc; else red[g * NR + nl] = acc;
      __syncthreads();
    } else if (k > 0 && N4 <= NR && (ke - kb) == 32 && SM) {      // idle lanes of the register path still join the barrier
      __syncthreads();
    } else
    if (k > 0) {
      const float* hprev = Y + (size_t)(k - 1) * ys;
      for (int n0 = 0; n0 < N4; n0 += NR) {
        const int n = n0 + nl;
        float acc = 0.0f;
        if (n < N4) {
          const float* w = Wh + n;
          int i = kb;
          for (; i + 4 <= ke; i += 4) {
            const float4 hv = *reinterpret_cast<const float4*>(hprev + i);
            acc = fmaf(hv.x, ldw<SM>(w + (size_t)(i + 0) * ldp), acc); acc = fmaf(hv.y, ldw<SM>(w + (size_t)(i + 1) * ldp), acc);
            acc = fmaf(hv.z, ldw<SM>(w + (size_t)(i + 2) * ldp), acc); acc = fmaf(hv.w, ldw<SM>(w + (size_t)(i + 3) * ldp), acc);
          }
          for (; i < ke; ++i) acc = fmaf(hprev[i], ldw<SM>(w + (size_t)i * ldp), acc);
          if (G == 1) Gt[(size_t)k * gs + n] += acc;
        }
        if (G > 1) red[g * NR + nl] = acc;
      }
      __syncthreads();
    }
    if (tid < nC) {
      const int o = tid;
      float* gk = Gt + (size_t)k * gs;
      float s0 = gk[o], s1 = gk[nC + o], s2 = gk[2 * nC + o], s3 = gk[3 * nC + o];
      if (k > 0 && G > 1) {
        for (int gg = 0; gg < G; ++gg) {
          s0 += red[gg * NR + o]; s1 += red[gg * NR + nC + o]; s2 += red[gg * NR + 2 * nC + o]; s3 += red[gg * NR + 3 * nC + o];
        }
      }
      const float ig = sigm_ref(s1), fg = sigm_ref(s2), og = sigm_ref(s3);
      const float oldPass = k > 0 ? Y[(size_t)(k - 1) * ys + nC4 + o] * fg : 0.0f;
      const float st = s0 * ig + oldPass;
      const float cop = tanh_ref(st);
      gk[o] = s0; gk[nC + o] = ig; gk[2 * nC + o] = fg; gk[3 * nC + o] = og;
      float* yk = Y + (size_t)k * ys;
      yk[o] = og * cop; yk[nC4 + o] = st; yk[2 * nC4 + o] = cop;
    }
    __syncthreads();
  }
}

// LSTMLayer::backward + Layer::backward for window steps T1-1 .. 0 (Layer_LSTM.h:127-166, Layers.h:123-188).
// E: error on y per step (from the layers above); on return Gt holds the four gate deltas of every step
// and, if the layer is not the first one, Ein has received Wx * delta.
template <bool SM>
__device__ void lstm_backward(const LayerDesc& L, const float* Wp, float* Gt, int gs, const float* Y, int ys, const float* E, int es,
                              float* Ein, int eis, float* red, int T1) {
  const int tid = threadIdx.x;
  const int nC = L.size, nI = L.nIn, N4 = 4 * nC, ldp = L.ldp, nC4 = (nC + 3) / 4 * 4;
  const float* Wx = Wp + L.imgW;
  const float* Wh = Wx + (size_t)nI * ldp;
  const int sh = L.bwdShift, KR = 1 << sh, G = kST >> sh;
  const int g = tid >> sh, il = tid & (KR - 1);
  const int N4q = (N4 + 3) >> 2;
  const int Nc = (N4q + G - 1) / G;
  const int nb = min(N4q, g * Nc), ne = min(N4q, nb + Nc);
  float sdNext = 0.0f, fgNext = 0.0f;
  // register-resident recurrent weights (see lstm_forward): this thread's 8 float4 of row `il` of Wh
  const bool regW = SM && il < nC && (ne - nb) == 8;
  float4 wreg[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) wreg[j] = regW ? ldw4<SM>(Wh + (size_t)il * ldp + (nb + j) * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
  for (int k = T1 - 1; k >= 0; --k) {
    if (tid < nC) {
      const int o = tid;
      float D = E[(size_t)k * es + o];
      if (k < T1 - 1) for (int gg = 0; gg < G; ++gg) D += red[gg * KR + o];      // Wh * delta of step k+1
      float* gk = Gt + (size_t)k * gs;
      const float cin = gk[o], ig = gk[nC + o], fg = gk[2 * nC + o], og = gk[3 * nC + o];
      const float cop = Y[(size_t)k * ys + 2 * nC4 + o];
      const float diff = (1.0f - cop * cop) * D;
      const float sd = diff * og + (k < T1 - 1 ? sdNext * fgNext : 0.0f);
      gk[o] = ig * sd;
      gk[nC + o] = ig * (1.0f - ig) * cin * sd;
      gk[2 * nC + o] = k > 0 ? fg * (1.0f - fg) * Y[(size_t)(k - 1) * ys + nC4 + o] * sd : 0.0f;
      gk[3 * nC + o] = og * (1.0f - og) * D * cop;
      sdNext = sd; fgNext = fg;
    }
    __syncthreads();
    if (k > 0) {
      float acc = 0.0f;
      if (regW) {
        const float* dl = Gt + (size_t)k * gs + nb * 4;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 wv = wreg[j], dv = *reinterpret_cast<const float4*>(dl + j * 4);
          acc = fmaf(wv.x, dv.x, acc); acc = fmaf(wv.y, dv.y, acc); acc = fmaf(wv.z, dv.z, acc); acc = fmaf(wv.w, dv.w, acc);
        }
      } else
      if (il < nC) {
        const float* w = Wh + (size_t)il * ldp;
        const float* dl = Gt + (size_t)k * gs;
        for (int n4 = nb; n4 < ne; ++n4) {
          const float4 wv = ldw4<SM>(w + n4 * 4), dv = *reinterpret_cast<const float4*>(dl + n4 * 4);
          acc = fmaf(wv.x, dv.x, acc); acc = fmaf(wv.y, dv.y, acc); acc = fmaf(wv.z, dv.z, acc); acc = fmaf(wv.w, dv.w, acc);
        }
      }
      red[g * KR + il] = acc;
      __syncthreads();
    }
  }
  if (L.needDx) {   // error on the layer below, all window steps at once
    for (int idx = tid; idx < nI * T1; idx += kST) {
      const int k = idx / nI, i = idx - k * nI;
      const float* w = Wx + (size_t)i * ldp;
      const float* dl = Gt + (size_t)k * gs;
      float acc = 0.0f;
      for (int n4 = 0; n4 < N4q; ++n4) {
        const float4 wv = ldw4<SM>(w + n4 * 4), dv = *reinterpret_cast<const float4*>(dl + n4 * 4);
        acc = fmaf(wv.x, dv.x, acc); acc = fmaf(wv.y, dv.y, acc); acc = fmaf(wv.z, dv.z, acc); acc = fmaf(wv.w, dv.w, acc);
      }
      Ein[(size_t)k * eis + i] += acc;
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------
// MGULayer (Layers/Layer_GRU.h:17-275; "MGU" and "GRU" build it, and it is what a partially observable MDP gets by default):
//   forget(k) = sigm(b_f + x_k Wff + h_{k-1} Wfr),  state(k) = tanh(b_s + x_k Wsf + (forget(k) * h_{k-1}) Wsr),
//   h_k = forget * state + (1 - forget) * h_{k-1};   W rows [input | recurrent], columns [forget | state].
// Gt per step: [forget | state], overwritten by [dL/dforget-input | dL/dstate-input] in the backward pass;
// Y per step: [h_k | h_{k-1} * forget(k)] (the second block feeds the weight gradient of the Wsr rows).
// ------------------------------------------------------------------------------------------
template <bool SM>
__device__ void mgu_forward(const LayerDesc& L, const float* Wp, const float* in, int is, float* Gt, int gs, float* Y, int ys,
                            float* red, int Tn) {
  const int tid = threadIdx.x;
  const int nC = L.size, nI = L.nIn, N2 = 2 * nC, ldp = L.ldp, nC4 = (nC + 3) / 4 * 4;
  const float* Wx = Wp + L.imgW;
  const float* Wh = Wx + (size_t)nI * ldp;
  const float* bias = Wp + L.imgB;
  // (a) b + x_k Wx for every window step (as lstm_forward)
  const int nq = (Tn + 3) >> 2;
  for (int idx = tid; idx < N2 * nq; idx += kST) {
    const int kq = idx / N2, n = idx - kq * N2, k0 = kq * 4;
    const float bv = ldw<SM>(bias + n);
    float acc[4] = {bv, bv, bv, bv};
    const float* x[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) x[j] = in + (size_t)min(k0 + j, Tn - 1) * is;
    const float* w = Wx + n;
    for (int i = 0; i < nI; ++i) {
      const float wv = ldw<SM>(w + (size_t)i * ldp);
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[j] = fmaf(x[j][i], wv, acc[j]);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) if (k0 + j < Tn) Gt[(size_t)(k0 + j) * gs + n] = acc[j];
  }
  __syncthreads();
  // (b) the recurrence: two dependent products per step.  Thread group g of G takes a K-slice of the nCells recurrent rows.
  int sh = 3; while ((1 << sh) < min(nC, kST)) ++sh;
  const int NR = 1 << sh, G = kST >> sh;
  const int g = tid >> sh, nl = tid & (NR - 1);
  const int Kc = (nC + G - 1) / G;
  const int kb = min(nC, g * Kc), ke = min(nC, kb + Kc);
  for (int k = 0; k < Tn; ++k) {
    float* gk = Gt + (size_t)k * gs;
    float* yk = Y + (size_t)k * ys;
    const float* hprev = Y + (size_t)(k - 1) * ys;
    if (k > 0) {                                 // forget pre-activation += h_{k-1} Wfr
      float acc = 0.0f;
      if (nl < nC) for (int i = kb; i < ke; ++i) acc = fmaf(hprev[i], ldw<SM>(Wh + (size_t)i * ldp + nl), acc);
      red[g * NR + nl] = acc;
      __syncthreads();
    }
    float fg = 0.0f;
    if (tid < nC) {
      float s0 = gk[tid];
      if (k > 0) for (int gg = 0; gg < G; ++gg) s0 += red[gg * NR + tid];
      fg = sigm_ref(s0);
      gk[tid] = fg;
      yk[nC4 + tid] = k > 0 ? hprev[tid] * fg : 0.0f;      // forget(k) * h_{k-1}
    }
    __syncthreads();
    if (k > 0) {                                 // state pre-activation += (forget * h_{k-1}) Wsr
      float acc = 0.0f;
      if (nl < nC) for (int i = kb; i < ke; ++i) acc = fmaf(yk[nC4 + i], ldw<SM>(Wh + (size_t)i * ldp + nC + nl), acc);
      red[g * NR + nl] = acc;
      __syncthreads();
    }
    if (tid < nC) {
      float s1 = gk[nC + tid];
      if (k > 0) for (int gg = 0; gg < G; ++gg) s1 += red[gg * NR + tid];
      const float st = tanh_ref(s1);
      gk[nC + tid] = st;
      yk[tid] = k > 0 ? fg * st + (1.0f - fg) * hprev[tid] : fg * st;
    }
    __syncthreads();
  }
}

// MGULayer::backward for window steps T1-1 .. 0 (Layer_GRU.h:122-214).  E: error on h per step — the recurrent part of step k
// is added into E of step k-1 here; on return Gt holds [dLdF | dLdS] of every step and Ein has received Wx * delta.
template <bool SM>
__device__ void mgu_backward(const LayerDesc& L, const float* Wp, float* Gt, int gs, const float* Y, int ys, float* E, int es,
                             float* Ein, int eis, float* red, int T1) {
  const int tid = threadIdx.x;
  const int nC = L.size, nI = L.nIn, N2 = 2 * nC, ldp = L.ldp;
  const float* Wx = Wp + L.imgW;
  const float* Wh = Wx + (size_t)nI * ldp;
  int sh = 3; while ((1 << sh) < min(nC, kST)) ++sh;
  const int KR = 1 << sh, G = kST >> sh;
  const int g = tid >> sh, il = tid & (KR - 1);
  const int Nc = (nC + G - 1) / G;
  const int nb = min(nC, g * Nc), ne = min(nC, nb + Nc);
  for (int k = T1 - 1; k >= 0; --k) {
    float* gk = Gt + (size_t)k * gs;
    const float* hprev = Y + (size_t)(k - 1) * ys;
    float dO = 0.f, fg = 0.f, st = 0.f;
    if (tid < nC) {                                                   // 1) dLdS = dLdO * forget * tanh'
      dO = E[(size_t)k * es + tid]; fg = gk[tid]; st = gk[nC + tid];
      gk[nC + tid] = dO * fg * (1.0f - st * st);
    }
    __syncthreads();
    if (k > 0) {                                                      // 2) dLdFprevOut = Wsr dLdS
      float acc = 0.0f;
      if (il < nC) {
        const float* w = Wh + (size_t)il * ldp + nC;
        for (int n = nb; n < ne; ++n) acc = fmaf(ldw<SM>(w + n), gk[nC + n], acc);
      }
      red[g * KR + il] = acc;
      __syncthreads();
    }
    float dFp = 0.f;
    if (tid < nC) {                                                   // 3) dLdF, and the element-wise part of 4)
      const float pO = k > 0 ? hprev[tid] : 0.0f;
      if (k > 0) for (int gg = 0; gg < G; ++gg) dFp += red[gg * KR + tid];
      gk[tid] = ((st - pO) * dO + dFp * pO) * fg * (1.0f - fg);
      if (k > 0) E[(size_t)(k - 1) * es + tid] += (1.0f - fg) * dO + fg * dFp;
    }
    __syncthreads();
    if (k > 0) {                                                      // 4) dLdprevOut += Wfr dLdF
      float acc = 0.0f;
      if (il < nC) {
        const float* w = Wh + (size_t)il * ldp;
        for (int n = nb; n < ne; ++n) acc = fmaf(ldw<SM>(w + n), gk[n], acc);
      }
      red[g * KR + il] = acc;
      __syncthreads();
      if (tid < nC) {
        float v = 0.0f;
        for (int gg = 0; gg < G; ++gg) v += red[gg * KR + tid];
        E[(size_t)(k - 1) * es + tid] += v;
      }
      __syncthreads();
    }
  }
  if (L.needDx) {   // error on the layer below, all window steps at once: Wff dLdF + Wsf dLdS
    const int N2q = (N2 + 3) >> 2;
    for (int idx = tid; idx < nI * T1; idx += kST) {
      const int k = idx / nI, i = idx - k * nI;
      const float* w = Wx + (size_t)i * ldp;
      const float* dl = Gt + (size_t)k * gs;
      float acc = 0.0f;
      for (int n4 = 0; n4 < N2q; ++n4) {
        const float4 wv = ldw4<SM>(w + n4 * 4), dv = *reinterpret_cast<const float4*>(dl + n4 * 4);
        acc = fmaf(wv.x, dv.x, acc); acc = fmaf(wv.y, dv.y, acc); acc = fmaf(wv.z, dv.z, acc); acc = fmaf(wv.w, dv.w, acc);
      }
      Ein[(size_t)k * eis + i] += acc;
    }
    __syncthreads();
  }
}

// Network::forward over a window, layer-major (Network.h:101-113 evaluated step by step from a zero recurrent state,
// Approximator.h:129-139): window steps [0, Tn) of the standardised states already in ws + sq.yOff[0]; the linear output
// layer is evaluated at step T1-1 (-> actTop) and, if vnext != nullptr, at step Tn-1 (-> V(s_{t+1}), RACER_train.cpp:23-27).
// Shared by the learner's P1 (p1_seq) and by the actors' policy evaluation (k_forward_seq).
template <bool SM>
__device__ __forceinline__ void seq_forward(const NetDesc& net, const SeqPlan& sq, const float* Wp, float* ws, float* red, float* actTop,
                                            uint64_t* bars, unsigned parity, int T1, int Tn, float* vnext, const StepArgs* dbg, int step) {
  const int tid = threadIdx.x;
  for (int l = 1; l < net.nLayers; ++l) {
    const LayerDesc& L = net.L[l];
    if (bars) mbar_wait(&bars[l], parity);
    if (dbg && l == 1) DBG_T(*dbg, step, 9);
    if (L.kind == kLSTM) {
      lstm_forward<SM>(L, Wp, ws + sq.yOff[L.in], sq.yStride[L.in], ws + sq.gOff[l], sq.gStride[l], ws + sq.yOff[l], sq.yStride[l], red, Tn);
      if (dbg && l == 1) DBG_T(*dbg, step, 10);
    } else if (L.kind == kMGU) {
      mgu_forward<SM>(L, Wp, ws + sq.yOff[L.in], sq.yStride[L.in], ws + sq.gOff[l], sq.gStride[l], ws + sq.yOff[l], sq.yStride[l], red, Tn);
    } else if (L.kind == kResidual) {       // ParametricResidualLayer::forward (Layers.h:347-361)
      const float* y1 = ws + sq.yOff[l - 1]; const float* y2 = ws + sq.yOff[l - 2];
      const int s1 = sq.yStride[l - 1], s2 = sq.yStride[l - 2], ys = sq.yStride[l];
      float* y = ws + sq.yOff[l];
      for (int idx = tid; idx < Tn * L.size; idx += kST) {
        const int k = idx / L.size, j = idx - k * L.size;
        y[k * ys + j] = y1[k * s1 + j] + (y2[k * s2 + j] * ldw<SM>(Wp + L.imgW + j) + ldw<SM>(Wp + L.imgB + j));
      }
      __syncthreads();
    } else if (L.kind == kDenseLinear) {    // outputs at the sampled step and, if needed, at the step after it
      const float* in = ws + sq.yOff[L.in]; const int is = sq.yStride[L.in];
      const float* x0 = in + (size_t)(T1 - 1) * is;
      const float* x1 = in + (size_t)(Tn - 1) * is;
      const int sh = L.fwdShift, NR = 1 << sh, G = kST >> sh;
      const int g = tid >> sh, nl = tid & (NR - 1);
      const int K = L.nIn, N = L.size, Kc = (K + G - 1) / G;
      const int kb = min(K, g * Kc), ke = min(K, kb + Kc);
      float a0 = 0.0f, a1 = 0.0f;
      if (nl < N) {
        const float* w = Wp + L.imgW + nl;
        for (int i = kb; i < ke; ++i) { const float wv = ldw<SM>(w + (size_t)i * L.ldp); a0 = fmaf(x0[i], wv, a0); a1 = fmaf(x1[i], wv, a1); }
      }
      red[g * NR + nl] = a0; red[kST + g * NR + nl] = a1;
      __syncthreads();
      if (tid < N) {
        float v0 = 0.0f, v1 = 0.0f;
        for (int gg = 0; gg < G; ++gg) { v0 += red[gg * NR + tid]; v1 += red[kST + gg * NR + tid]; }
        const float bv = ldw<SM>(Wp + L.imgB + tid);
        actTop[L.actOff + tid] = v0 + bv;
        if (tid == 0 && vnext) vnext[0] = (float)net2v((double)(v1 + bv));
      }
      __syncthreads();
    }
  }
}

template <bool SM>
__device__ void p1_seq(const StepArgs& a, const DevDescs& dd, StepCtrl& c, int step, int b, unsigned char* smraw, const SeqSmem& sp,
                       unsigned parity, bool fetchCtrl, const unsigned* readyFlag, unsigned readyTarget) {
  const NetDesc& net = dd.net; const SeqPlan& sq = dd.seq; const Hyper& hp = dd.hp;
  const int tid = threadIdx.x;
  float* img = reinterpret_cast<float*>(smraw + sp.img);
  const float* Wp = SM ? img : a.Wimg;
  float* ws = reinterpret_cast<float*>(smraw + sp.ws);
  int* info = reinterpret_cast<int*>(smraw + sp.info);        // row, slot, hasNext, valid, nRecurr
  float* old = reinterpret_cast<float*>(smraw + sp.old);      // V, ADV, RHO, KL, DELTA, Vnext, ADVnext, Qret
  double* pair = reinterpret_cast<double*>(smraw + sp.pair);
  double* samp = reinterpret_cast<double*>(smraw + sp.samp);
  uint64_t* bars = (SM && a.useTma) ? reinterpret_cast<uint64_t*>(smraw + sp.bars) : nullptr;
  float* red = ws + sq.red;
  float* actTop = ws + sq.actTop;
  float* errTop = ws + sq.errTop;
  const ReplayView& rp = a.rp;
  const int dS = net.dS, dA = net.dA, Tc = net.Tc;

  if (tid == 0) {
    const size_t j = (size_t)(step - a.stepBase) * a.B + b;
    const int row = a.sampRow[j];
    const int sf = a.sampSlot[j];
    const int slot = sf & 0x7fffffff, hn = (sf >> 31) & 1;
    const int t = row - __ldcg(rp.epStart + slot);
    info[0] = row; info[1] = slot; info[2] = hn; info[3] = 1; info[4] = min(net.bptt, t);
    old[0] = ld_cg(rp.V + row); old[1] = ld_cg(rp.ADV + row); old[2] = ld_cg(rp.RHO + row); old[3] = ld_cg(rp.KL + row);
    old[4] = ld_cg(rp.DELTA + row); old[7] = ld_cg(rp.Q + row);
    old[5] = 0.f; old[6] = 0.f;
    if (hn) { old[5] = ld_cg(rp.V + row + 1); old[6] = ld_cg(rp.ADV + row + 1); }
  }
  for (int idx = tid; idx < sq.red; idx += kST) ws[idx] = 0.0f;     // pads, error buffers, top vectors
  __syncthreads();
  const int row = info[0], hn = info[2], nRec = info[4];
  const int T1 = nRec + 1, Tn = T1 + hn;            // window steps; + the state after the sampled step (RACER_train.cpp:23-27)
  // gather + standardise the window's states (Episode.h:171-183)
  float* X = ws + sq.yOff[0];
  const int xs = sq.yStride[0];
  for (int idx = tid; idx < Tn * dS; idx += kST) {
    const int k = idx / dS, i = idx - k * dS;
    X[k * xs + i] = (ld_cg(rp.S + (size_t)(row - nRec + k) * dS + i) - ld_cg(rp.stateMean + i)) * ld_cg(rp.stateScale + i);
  }
  double pa = 0, pmm = 0, pms = 1;
  if (tid < dA) {
    pa = (double)ld_cg(rp.A + (size_t)row * dA + tid);
    pmm = (double)ld_cg(rp.MU + (size_t)row * 2 * dA + tid);
    pms = (double)ld_cg(rp.MU + (size_t)row * 2 * dA + dA + tid);
  }
  __syncthreads();
  if (tid < dS && (step == a.lastStep || a.lastStep < 0)) a.lastX[(size_t)b * dS + tid] = X[nRec * xs + tid];
  DBG_T(a, step, 1);

  // ---- forward, layer-major ----
  const LayerDesc& Lo = net.L[net.nLayers - 2];
  const LayerDesc& Lp = net.L[net.nLayers - 1];
  float* vnext = reinterpret_cast<float*>(samp + 11);
  seq_forward<SM>(net, sq, Wp, ws, red, actTop, bars, parity, T1, Tn, hn ? vnext : nullptr, &a, step);
  DBG_T(a, step, 2);

  {
    LossIO io{actTop, errTop, info, old, pair, samp, vnext, b, pa, pmm, pms, 0, tid < dA ? tid : 0};
    loss_stages<1, SM>(a, net, hp, c, step, Wp, io, fetchCtrl, readyFlag, readyTarget);
  }

  // ---- backward through time, layers last to first ----
  const int col0 = b * Tc;
  {   // output layer: error on the top hidden layer at the sampled step (Layers.h:131-145)
    const int lt = Lo.in;
    float* e = ws + sq.eOff[lt] + (size_t)(T1 - 1) * sq.eStride[lt];
    const int sh = Lo.bwdShift, KR = 1 << sh, G = kST >> sh;
    const int g = tid >> sh, il = tid & (KR - 1);
    const int K = Lo.nIn, N = Lo.size, Nc = (N + G - 1) / G;
    const int nb = min(N, g * Nc), ne = min(N, nb + Nc);
    float acc = 0.0f;
    if (il < K) {
      const float* w = Wp + Lo.imgW + (size_t)il * Lo.ldp;
      for (int n = nb; n < ne; ++n) acc = fmaf(ldw<SM>(w + n), errTop[Lo.actOff + n], acc);
    }
    red[g * KR + il] = acc;
    __syncthreads();
    if (tid < K) { float v = 0.0f; for (int gg = 0; gg < G; ++gg) v += red[gg * KR + tid]; e[tid] += v; }
    __syncthreads();
  }
  for (int l = net.nLayers - 3; l >= 1; --l) {
    const LayerDesc& L = net.L[l];
    if (L.kind == kResidual) {               // ParametricResidualLayer::backward (Layers.h:363-393)
      const float* e = ws + sq.eOff[l]; const int es = sq.eStride[l];
      float* e1 = ws + sq.eOff[l - 1]; float* e2 = ws + sq.eOff[l - 2];
      const int es1 = sq.eStride[l - 1], es2 = sq.eStride[l - 2];
      for (int idx = tid; idx < T1 * L.size; idx += kST) {
        const int k = idx / L.size, j = idx - k * L.size;
        const float d = e[k * es + j];
        e1[k * es1 + j] = d;
        e2[k * es2 + j] += d * ldw<SM>(Wp + L.imgW + j);
      }
      __syncthreads();
      seq_store(a.errG, a.Bpad, L.actOff, L.size, col0, Tc, T1, e, es, 0);
      seq_store(a.actG, a.Bpad, L.actOff, L.size, col0, Tc, T1, ws + sq.yOff[l], sq.yStride[l], 0);
    } else if (L.kind == kLSTM) {
      float* Gt = ws + sq.gOff[l];
      const float* Y = ws + sq.yOff[l];
      lstm_backward<SM>(L, Wp, Gt, sq.gStride[l], Y, sq.yStride[l], ws + sq.eOff[l], sq.eStride[l],
                        L.in > 0 ? ws + sq.eOff[L.in] : nullptr, L.in > 0 ? sq.eStride[L.in] : 0, red, T1);
      seq_store(a.errG, a.Bpad, L.actOff, 4 * L.size, col0, Tc, T1, Gt, sq.gStride[l], 0);            // gate deltas
      seq_store(a.actG, a.Bpad, L.actOff, L.size, col0, Tc, T1, Y, sq.yStride[l], 0);                 // y_k
      seq_store(a.actG, a.Bpad, L.actOff + L.size, L.size, col0, Tc, T1, Y, sq.yStride[l], 1);        // h_{k-1}
    } else if (L.kind == kMGU) {
      float* Gt = ws + sq.gOff[l];
      const float* Y = ws + sq.yOff[l];
      const int nC4 = (L.size + 3) / 4 * 4;
      mgu_backward<SM>(L, Wp, Gt, sq.gStride[l], Y, sq.yStride[l], ws + sq.eOff[l], sq.eStride[l],
                       L.in > 0 ? ws + sq.eOff[L.in] : nullptr, L.in > 0 ? sq.eStride[L.in] : 0, red, T1);
      seq_store(a.errG, a.Bpad, L.actOff, 2 * L.size, col0, Tc, T1, Gt, sq.gStride[l], 0);            // [dLdF | dLdS]
      seq_store(a.actG, a.Bpad, L.actOff, L.size, col0, Tc, T1, Y, sq.yStride[l], 0);                 // h_k
      seq_store(a.actG, a.Bpad, L.actOff + L.size, L.size, col0, Tc, T1, Y, sq.yStride[l], 1);        // h_{k-1}
      seq_store(a.actG, a.Bpad, L.actOff + 2 * L.size, L.size, col0, Tc, T1, Y + nC4, sq.yStride[l], 0);   // h_{k-1} * forget(k)
    }
  }
  seq_store(a.actG, a.Bpad, net.L[0].actOff, dS, col0, Tc, T1, X, xs, 0);
  // compact columns (one per sample): output / stdev gradients and the top hidden output at the sampled step
  {
    const float* ytop = ws + sq.yOff[Lo.in] + (size_t)(T1 - 1) * sq.yStride[Lo.in];
    for (int i = tid; i < Lo.nIn; i += kST) a.actG[(size_t)(net.topInOff + i) * a.Bpad + b] = ytop[i];
    for (int n = tid; n < Lo.size; n += kST) a.errG[(size_t)(Lo.actOff + n) * a.Bpad + b] = errTop[Lo.actOff + n];
    for (int i = tid; i < Lp.size; i += kST) a.errG[(size_t)(Lp.actOff + i) * a.Bpad + b] = errTop[Lp.actOff + i];
  }
  DBG_T(a, step, 4);
}

// ------------------------------------------------------------------------------------------
// P2: weight gradient tile + Adam  (Layers.h:160-187, Optimizer.cpp:61-108,122-161)
// ------------------------------------------------------------------------------------------
constexpr int kBC = 256;          // batch chunk staged in shared memory
constexpr int kBCP = kBC + 4;     // padded row stride

struct AdamCoef { float eta, B1, B2, lambda, fac; };

__device__ __forceinline__ AdamCoef adam_coef(const Hyper& hp, const StepCtrl& c) {
  AdamCoef k;
  k.eta = c.adam_eta;   // precomputed with the rest of ctrl (adam_eta_for)
  k.B1 = 0.9f; k.B2 = 0.999f;
  k.lambda = (float)hp.nnLambda;
  k.fac = (float)(1.0 / (double)hp.batchGlobal);
  return k;
}

__host__ __device__ __forceinline__ float adam_step(const AdamCoef& k, float G, float W, float m1, float m2, float* w, float* pm1, float* pm2) {
  const float penal = -W * k.lambda;                       // SMARTIES_ADAMW
  const float DW = k.fac * G;
  float M1 = k.B1 * m1 + (1.0f - k.B1) * DW;
  float M2 = k.B2 * m2 + (1.0f - k.B2) * DW * DW;
  const float numer = k.B1 * M1 + (1.0f - k.B1) * DW;      // SMARTIES_NESTEROV_ADAM
  M2 = M2 < M1 * M1 ? M1 * M1 : M2;                        // SMARTIES_SAFE_ADAM
  const float ret = numer / (FLT_EPSILON + sqrtf(M2));
  const float Wn = W + k.eta * (ret + penal);
  *pm1 = M1; *pm2 = M2; *w = Wn;
  return Wn;
}

// "update frozen weights" of AdamOptimizer::apply_update (Network/Optimizer.cpp:162-177) for one parameter, right after its Adam
// step: targetDelay >= 1: `cntUpdateDelay` reaches 0 every floor(targetDelay) updates, counted from construction / restart, and the
// weights are copied; targetDelay < 1: `targetAry[j] += tgtUpdateAlpha * (paramAry[j] - targetAry[j])` (Real = double times the f32
// difference, added to the f32 target) after every update.
__device__ __forceinline__ void target_step(const StepArgs& a, const StepCtrl& c, int p, float wNew) {
  if (a.tgtAlpha >= 1.0) {
    const long long D = (long long)a.tgtAlpha;
    if ((c.adam_step - a.tgtPhase) % D == 0) a.Wtgt[p] = wNew;
  } else {
    const float t = a.Wtgt[p];
    a.Wtgt[p] = (float)((double)t + a.tgtAlpha * (double)(wNew - t));
  }
}

__device__ void p2_tile(const StepArgs& a, const NetDesc& net, const Hyper& hp, const StepCtrl& c, const GradTile t,
                        float* tiles, int step, int tileIdx, const TcPlan* tc = nullptr) {
  const int tid = threadIdx.x;
  float* As = tiles;                 // [16][kBCP]
  float* Ds = tiles + kTileK * kBCP; // [16][kBCP]
  const LayerDesc& L = net.L[t.layer];
  // operand rows: A = activation of the input layer (dense) / of layer ID-2 (residual); D = this layer's deltas
  // LSTM layers: rows [0, nIn) of W multiply the layer input, rows [nIn, nIn + nCells) the previous step's
  // output (stored next to y in the layer's activation rows), N = 4 nCells gate deltas (Layer_LSTM.h:24-29)
  const bool lstm = is_cell_layer(L.kind);           // LSTM or MGU cells: rows [input | recurrent]
  const int K = t.kind == 0 ? (lstm ? L.nIn + L.size : L.nIn) : L.size;
  const int aOff = t.kind == 0 ? ((net.recurrent && L.kind == kDenseLinear) ? net.topInOff : net.L[L.in].actOff)
                               : (t.kind == 1 ? net.L[t.layer - 2].actOff : 0);
  // recurrent rows k >= nIn read h_{k-1}; the state columns of an MGU layer read h_{k-1} * forget(k) (Layer_GRU.h:205-212)
  const int hOff = (L.kind == kMGU && t.n0 >= L.size ? L.actOff + 2 * L.size : L.actOff + L.size) - L.nIn;
  const int dOff = L.actOff;
  const int N = t.nLimit > 0 ? t.nLimit : cell_gates(L.kind) * L.size;
  const int kk = tid >> 4, nn = tid & 15;
  const int warp = tid >> 5, lane = tid & 31;
  // parameters this thread owns: fetch W, M1, M2 now, use them after the contraction
  int p0 = -1, p1 = -1, pimg0 = -1, pimg1 = -1;
  if (tid >= 256) {
    // only the first 256 threads own parameters; the rest help with loads and the contraction
  } else if (t.kind == 0) {
    const int k = t.k0 + kk, n = t.n0 + nn;
    if (n < N && k <= K) { p0 = k < K ? L.wOff + k * L.ld + n : L.bOff + n; pimg0 = k < K ? L.imgW + k * L.ldp + n : L.imgB + n; }
  } else if (lane < 2) {
    const int n = t.n0 + warp * 2 + lane;
    if (n < N) {
      p0 = L.bOff + n; pimg0 = L.imgB + n;
      if (t.kind == 1) { p1 = L.wOff + n; pimg1 = L.imgW + n; }
    }
  }
  float w0 = 0.f, m10 = 0.f, m20 = 0.f, w1 = 0.f, m11 = 0.f, m21 = 0.f;
  if (p0 >= 0) { w0 = ld_cg(a.W + p0); m10 = ld_cg(a.M1 + p0); m20 = ld_cg(a.M2 + p0); }
  if (p1 >= 0) { w1 = ld_cg(a.W + p1); m11 = ld_cg(a.M1 + p1); m21 = ld_cg(a.M2 + p1); }
  float acc = 0.f, acc2 = 0.f;
  float q4[4] = {0.f, 0.f, 0.f, 0.f};
  const AdamCoef ac = adam_coef(hp, c);
  DBG_T(a, step, 24);
  // Operand tiles of one batch chunk: global -> registers -> shared memory.  The loads of chunk c+1 are issued
  // before the contraction of chunk c (software pipelining: recurrent nets contract over B * (Tc+1) columns, 17 chunks
  // at cfg3, and paid one L2 round trip per chunk).
  constexpr int kLD = 1024 / kST;   // float4 per thread and operand (16 rows x 64 float4)
  float4 avs[kLD], dvs[kLD];
  const int nLink = t.kind == 1 ? net.L[t.layer - 2].size : 0x7fffffff;    // ParametricResidual: min(size of layer ID-2, size) linked units
  auto load_chunk = [&](int bc) {
#pragma unroll
    for (int i = 0; i < kLD; ++i) {
      const int q = tid + i * kST;
      // feed-forward nets: tile-major scratch, 16 consecutive lanes read 16 consecutive features of one P1 tile (256 B);
      // recurrent nets: feature-major scratch [feature][column], 64 consecutive lanes read one 1 KB row
      const bool tm = !net.recurrent;
      const int r = tm ? (q & 15) : (q >> 6), c4 = tm ? (q >> 4) * 4 : (q & 63) * 4;
      const size_t tb = (size_t)((bc + c4) >> 2) * net.actPerSample;       // first feature of that P1 tile (tile-major)
      float4 av = make_float4(0.f, 0.f, 0.f, 0.f), dv = av;
      if (t.kind == 0) {
        const int k = t.k0 + r;
        if (k < K) av = tm ? ld_cg4(a.actG + (tb + aOff + k) * 4)
                           : ld_cg4(a.actG + (size_t)((lstm && k >= L.nIn ? hOff : aOff) + k) * a.Bpad + bc + c4);
        else if (k == K) av = make_float4(1.f, 1.f, 1.f, 1.f);            // bias row: db += delta
        const int n = t.n0 + r;
        if (n < N) dv = tm ? ld_cg4(a.errG + (tb + dOff + n) * 4) : ld_cg4(a.errG + (size_t)(dOff + n) * a.Bpad + bc + c4);
      } else {
        const int n = t.n0 + r;
        if (n < N && n < nLink) {     // residual: units beyond the layer below have no skip link and no gradient (Layers.h:363-393)
          dv = tm ? ld_cg4(a.errG + (tb + dOff + n) * 4) : ld_cg4(a.errG + (size_t)(dOff + n) * a.Bpad + bc + c4);
          if (t.kind == 1) av = tm ? ld_cg4(a.actG + (tb + aOff + n) * 4) : ld_cg4(a.actG + (size_t)(aOff + n) * a.Bpad + bc + c4);
        }
      }
      avs[i] = av; dvs[i] = dv;
    }
  };
  // LSTM layers whose contraction ran on the tensor cores (tc_wgrad_item): add the K-slices of this tile's outputs in
  // slice order and skip the SIMT contraction
  const bool fromTc = tc && lstm && t.kind == 0 && tc->slices[t.layer] > 0;
  if (fromTc) {
    if (tid < 256) {
      const int k = t.k0 + kk, n = t.n0 + nn;
      if (n < N && k <= K) {
        const int nS = tc->slices[t.layer];
        const float* src = a.tcPartial + ((size_t)(tc->item0[t.layer] + (n >> 7) * nS) * 128 + k) * 128 + (n & 127);
        float v = 0.f;
        int sl = 0;
        for (; sl + 8 <= nS; sl += 8) {       // eight loads in flight, added in slice order
          float x[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) x[u] = ld_cg(src + (size_t)(sl + u) * 128 * 128);
#pragma unroll
          for (int u = 0; u < 8; ++u) v += x[u];
        }
        for (; sl < nS; ++sl) v += ld_cg(src + (size_t)sl * 128 * 128);
        acc = v;
      }
    }
  } else {
  load_chunk(0);
  for (int bc = 0; bc < t.cols; bc += kBC) {
    if (bc) __syncthreads();
#pragma unroll
    for (int i = 0; i < kLD; ++i) {
      const int q = tid + i * kST;
      const bool tm = !net.recurrent;
      const int r = tm ? (q & 15) : (q >> 6), c4 = tm ? (q >> 4) * 4 : (q & 63) * 4;
      *reinterpret_cast<float4*>(As + r * kBCP + c4) = avs[i];
      *reinterpret_cast<float4*>(Ds + r * kBCP + c4) = dvs[i];
    }
    if (bc + kBC < t.cols) load_chunk(bc + kBC);
    __syncthreads();
    DBG_T(a, step, 25);
    if (t.kind == 0) {
      // 2x2 register blocking: thread (g, k2, n2) accumulates outputs (k2|k2+8, n2|n2+8) over the
      // slice g of the batch chunk; partials are combined through shared memory below
      constexpr int kG = kST / 64, kBL = kBC / kG;
      const int g = tid >> 6, k2 = (tid >> 3) & 7, n2 = tid & 7;
      const float* a0 = As + k2 * kBCP + g * kBL;
      const float* a1 = a0 + 8 * kBCP;
      const float* d0 = Ds + n2 * kBCP + g * kBL;
      const float* d1 = d0 + 8 * kBCP;
#pragma unroll 4
      for (int b4 = 0; b4 < kBL; b4 += 4) {
        const float4 xa = *reinterpret_cast<const float4*>(a0 + b4), xb = *reinterpret_cast<const float4*>(a1 + b4);
        const float4 da = *reinterpret_cast<const float4*>(d0 + b4), db = *reinterpret_cast<const float4*>(d1 + b4);
        q4[0] = fmaf(xa.x, da.x, q4[0]); q4[0] = fmaf(xa.y, da.y, q4[0]); q4[0] = fmaf(xa.z, da.z, q4[0]); q4[0] = fmaf(xa.w, da.w, q4[0]);
        q4[1] = fmaf(xa.x, db.x, q4[1]); q4[1] = fmaf(xa.y, db.y, q4[1]); q4[1] = fmaf(xa.z, db.z, q4[1]); q4[1] = fmaf(xa.w, db.w, q4[1]);
        q4[2] = fmaf(xb.x, da.x, q4[2]); q4[2] = fmaf(xb.y, da.y, q4[2]); q4[2] = fmaf(xb.z, da.z, q4[2]); q4[2] = fmaf(xb.w, da.w, q4[2]);
        q4[3] = fmaf(xb.x, db.x, q4[3]); q4[3] = fmaf(xb.y, db.y, q4[3]); q4[3] = fmaf(xb.z, db.z, q4[3]); q4[3] = fmaf(xb.w, db.w, q4[3]);
      }
    } else if (tid < 256) {
      // vector tiles: warp w (of the first 8) reduces rows 2w, 2w+1 over the batch chunk
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
  }
  if (t.kind == 0 && !fromTc) {   // combine the batch slices: thread (kk, nn) of the first 256 owns output (kk, nn)
    __syncthreads();
    float* part = As;  // [kST/64][64][4]
    *reinterpret_cast<float4*>(part + tid * 4) = make_float4(q4[0], q4[1], q4[2], q4[3]);
    __syncthreads();
    if (tid < 256) {
      const int q = (kk & 7) * 8 + (nn & 7), j = (kk >> 3) * 2 + (nn >> 3);
      float v = 0.f;
#pragma unroll
      for (int g = 0; g < kST / 64; ++g) v += part[(g * 64 + q) * 4 + j];
      acc = v;
    }
  }
  // ---- gradient sum over learner ranks, fused into the tile (replaces the MPI_Iallreduce of
  //      AdamOptimizer::prepare_update, Optimizer.cpp:114-118): every rank pushes its partial tile
  //      into every peer's slot over NVLink as 4-byte elements (poison = not arrived yet), polls its own
  //      slots in LOCAL memory and adds them in rank order, so all ranks apply the identical update ----
  if (a.comm.world > 1) {
    DBG_T(a, step, 33);
    const CommView& cm = a.comm;
    const int N = cm.world, me = cm.rank, rot = step & 3;
    const size_t slotMe = ((size_t)rot * N + me) * cm.nParamsPad;
    if (p0 >= 0) {
      const unsigned pk = __float_as_uint(acc);
      for (int q = 0; q < N; ++q) if (q != me) st_volatile_u32(cm.grad(q) + slotMe + p0, pk);
    }
    if (p1 >= 0) {
      const unsigned pk = __float_as_uint(acc2);
      for (int q = 0; q < N; ++q) if (q != me) st_volatile_u32(cm.grad(q) + slotMe + p1, pk);
    }
    DBG_T(a, step, 34);
    unsigned* mine = cm.grad(me) + (size_t)rot * N * cm.nParamsPad;
    unsigned got[kMaxWorld];
    // a peer that timed out (error flag, reported by the host after the call): leave the parameter untouched instead of
    // summing the poison pattern into the gradient and Adam
    if (p0 >= 0) {
      if (!wait_values_poison(mine + p0, cm.nParamsPad, N, me, cm, got)) p0 = -1;
      float v = 0.f;
#pragma unroll
      for (int q = 0; q < kMaxWorld; ++q) if (q < N) v += q == me ? acc : __uint_as_float(got[q]);
      acc = v;
    }
    if (p1 >= 0) {
      if (!wait_values_poison(mine + p1, cm.nParamsPad, N, me, cm, got)) p1 = -1;
      float v = 0.f;
#pragma unroll
      for (int q = 0; q < kMaxWorld; ++q) if (q < N) v += q == me ? acc2 : __uint_as_float(got[q]);
      acc2 = v;
    }
  }
  DBG_T(a, step, 26);
  // the parameter / moment loads issued before the contraction are consumed only here: keep the
  // compiler from hoisting the first Adam multiplications (and with them the wait) above the tile loop
  asm volatile("" : "+f"(w0), "+f"(m10), "+f"(m20), "+f"(w1), "+f"(m11), "+f"(m21));
  if (p0 >= 0) {
    a.G[p0] = acc;
    const float wn = adam_step(ac, acc, w0, m10, m20, a.W + p0, a.M1 + p0, a.M2 + p0);
    a.Wimg[pimg0] = wn;
    if (a.Wtgt) target_step(a, c, p0, wn);
  }
  if (p1 >= 0) {
    a.G[p1] = acc2;
    const float wn = adam_step(ac, acc2, w1, m11, m21, a.W + p1, a.M1 + p1, a.M2 + p1);
    a.Wimg[pimg1] = wn;
    if (a.Wtgt) target_step(a, c, p1, wn);
  }
}


// ------------------------------------------------------------------------------------------
// Tensor-core weight gradient of LSTM layers (recurrent nets, persistent kernel only).
//
// dW[k][n] = sum over the B*(Tc+1) (sample, window step) columns of  A[k][col] * Delta[n][col],  A = [x | h_prev | 1]
// (K = nIn + nCells + 1 <= 128 rows), Delta = the 4*nCells gate deltas — the one large contraction of this path
// (97 x 256 outputs over 4224 columns at cfg3).  Work item = (layer, n-tile of 64 gate columns, K-slice of Wc columns):
//   1. all threads stage the slice of both operands from the feature-major scratch into shared memory, split into a
//      TF32 "hi" part and the f32 remainder "lo" (3xTF32: hi*hi + hi*lo + lo*hi keeps f32 accuracy — plain TF32 misses the
//      parity tolerances, profiles/r1/microbench_tcgen05_tf32.txt), in the UMMA no-swizzle K-major layout
//      smem[k-chunk][row] (one float4 = 4 columns; leading byte offset = (rows+1)*16 B so that the staging stores are
//      conflict-free, stride byte offset 128 B);
//   2. one thread issues Wc/8 x 3 tcgen05.mma.cta_group::1.kind::tf32 (M 128, N 64, accumulator in tensor memory) and
//      commits them to an mbarrier;
//   3. warps 0-3 read the accumulator with tcgen05.ld.32x32b and store the partial tile to global memory.
// After a grid barrier the 16x16 tile owners of p2_tile add the K-slices in slice order (deterministic) and continue with the
// exchange and Adam.  The staging area is the whole dynamic shared memory between the descriptors and the sample scalars
// (weight image + sequence workspace: both dead during P2, the image is reloaded after barrier 2).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t umma_desc(const void* base, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;                                               // cute/arch/mma_sm100_desc.hpp:98-123
  d |= (uint64_t)((smem_u32(base) >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;                                       // version 1 (Blackwell); no swizzle, base offset 0
  return d;
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
               ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ float tf32_hi(float x) { uint32_t u; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x)); return __uint_as_float(u); }
__device__ __forceinline__ void split4(const float4 v, float4& h, float4& l) {
  h = make_float4(tf32_hi(v.x), tf32_hi(v.y), tf32_hi(v.z), tf32_hi(v.w));
  l = make_float4(v.x - h.x, v.y - h.y, v.z - h.z, v.w - h.w);
}
constexpr int kTcLdA = 129, kTcLdB = 129;         // rows + 1 float4 per k-chunk
constexpr int kTcN = 128;                         // gate columns per item = N of the MMA = TMEM columns

__device__ void tc_wgrad_item(const StepArgs& a, const NetDesc& net, const TcPlan& plan, int item, unsigned char* staging,
                              uint32_t tmem_d, uint64_t* bar, unsigned& phase, int step) {
  DBG_T(a, step, 42);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  int l = 1;
  for (; l < net.nLayers; ++l) if (plan.slices[l] && item >= plan.item0[l] && item < plan.item0[l] + plan.slices[l] * plan.nT[l]) break;
  const LayerDesc& L = net.L[l];
  const int li = item - plan.item0[l], nt = li / plan.slices[l], sl = li - nt * plan.slices[l];
  const int Wc = plan.Wc, KC = Wc >> 2, col0 = sl * Wc;
  const int K = L.nIn + L.size, N = 4 * L.size, n0 = nt * kTcN;
  const int aOff = net.L[L.in].actOff, hOff = L.actOff + L.size - L.nIn, dOff = L.actOff;
  const int cols = a.Bpad;
  float4* Ah = reinterpret_cast<float4*>(staging);
  float4* Al = Ah + (size_t)KC * kTcLdA;
  float4* Bh = Al + (size_t)KC * kTcLdA;
  float4* Bl = Bh + (size_t)KC * kTcLdB;
  // ---- 1. stage + split: 8 consecutive lanes read 8 consecutive k-chunks (128 B) of one scratch row.  Wc <= 96 means at
  //         most 6 float4 per thread and operand: ALL loads of both operands are issued before the first is split
  //         (the staging is a latency chain of L2 round trips otherwise: 4.8 us with four loads in flight) ----
  constexpr int kU = 6;
  float4 va[kU], vb[kU]; int da[kU], db[kU];
#pragma unroll
  for (int u = 0; u < kU; ++u) {
    const int q = tid + u * kST;
    va[u] = make_float4(0.f, 0.f, 0.f, 0.f); da[u] = -1;
    if (q < 128 * KC) {
      const int r = q / KC, cc = q - r * KC, col = col0 + 4 * cc;
      da[u] = cc * kTcLdA + r;
      if (col < cols) {
        if (r < K) va[u] = ld_cg4(a.actG + (size_t)((r >= L.nIn ? hOff : aOff) + r) * a.Bpad + col);
        else if (r == K) va[u] = make_float4(1.f, 1.f, 1.f, 1.f);       // bias row: db += delta
      }
    }
  }
#pragma unroll
  for (int u = 0; u < kU; ++u) {
    const int q = tid + u * kST;
    vb[u] = make_float4(0.f, 0.f, 0.f, 0.f); db[u] = -1;
    if (q < kTcN * KC) {
      const int r = q / KC, cc = q - r * KC, col = col0 + 4 * cc, n = n0 + r;
      db[u] = cc * kTcLdB + r;
      if (col < cols && n < N) vb[u] = ld_cg4(a.errG + (size_t)(dOff + n) * a.Bpad + col);
    }
  }
#pragma unroll
  for (int u = 0; u < kU; ++u) if (da[u] >= 0) { float4 h, lo; split4(va[u], h, lo); Ah[da[u]] = h; Al[da[u]] = lo; }
#pragma unroll
  for (int u = 0; u < kU; ++u) if (db[u] >= 0) { float4 h, lo; split4(vb[u], h, lo); Bh[db[u]] = h; Bl[db[u]] = lo; }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic stores -> tensor-core (async proxy) reads
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  DBG_T(a, step, 43);
  // ---- 2. MMA issue by one thread ----
  if (tid == 0) {
    constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (((uint32_t)kTcN >> 3) << 17) | ((128u >> 4) << 24);   // f32 <- tf32 x tf32, K-major, N 128, M 128
    for (int kk = 0; kk < Wc / 8; ++kk) {
      const uint64_t dah = umma_desc(Ah + (size_t)(2 * kk) * kTcLdA, kTcLdA * 16, 128), dal = umma_desc(Al + (size_t)(2 * kk) * kTcLdA, kTcLdA * 16, 128);
      const uint64_t dbh = umma_desc(Bh + (size_t)(2 * kk) * kTcLdB, kTcLdB * 16, 128), dbl = umma_desc(Bl + (size_t)(2 * kk) * kTcLdB, kTcLdB * 16, 128);
      umma_tf32(tmem_d, dal, dbh, idesc, kk > 0);
      umma_tf32(tmem_d, dah, dbl, idesc, 1);
      umma_tf32(tmem_d, dah, dbh, idesc, 1);
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
  }
  // ---- 3. accumulator -> registers -> partial tile ----
  if (warp < 4) {
    uint32_t ok = 0;
    for (int spin = 0; spin < (1 << 24) && !ok; ++spin)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok) : "r"(smem_u32(bar)), "r"(phase) : "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    DBG_T(a, step, 44);
    if (warp * 32 <= K) {                     // rows beyond the bias row are padding
      uint32_t v[64];
#define SMB200_TMEM_LD32(off) asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 " \
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];" \
        : "=r"(v[off+0]),"=r"(v[off+1]),"=r"(v[off+2]),"=r"(v[off+3]),"=r"(v[off+4]),"=r"(v[off+5]),"=r"(v[off+6]),"=r"(v[off+7]), \
          "=r"(v[off+8]),"=r"(v[off+9]),"=r"(v[off+10]),"=r"(v[off+11]),"=r"(v[off+12]),"=r"(v[off+13]),"=r"(v[off+14]),"=r"(v[off+15]), \
          "=r"(v[off+16]),"=r"(v[off+17]),"=r"(v[off+18]),"=r"(v[off+19]),"=r"(v[off+20]),"=r"(v[off+21]),"=r"(v[off+22]),"=r"(v[off+23]), \
          "=r"(v[off+24]),"=r"(v[off+25]),"=r"(v[off+26]),"=r"(v[off+27]),"=r"(v[off+28]),"=r"(v[off+29]),"=r"(v[off+30]),"=r"(v[off+31]) \
        : "r"(taddr + off))
      const int row = warp * 32 + lane;
#pragma unroll
      for (int half = 0; half < kTcN / 64; ++half) {
        const uint32_t taddr = tmem_d + ((uint32_t)(warp * 32) << 16) + half * 64;
        SMB200_TMEM_LD32(0); SMB200_TMEM_LD32(32);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (ok && row <= K) {
          float4* dst = reinterpret_cast<float4*>(a.tcPartial + ((size_t)item * 128 + row) * kTcN + half * 64);
#pragma unroll
          for (int j = 0; j < 16; ++j)
            dst[j] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
        }
      }
    }
    if (!ok && lane == 0 && a.comm.error) *a.comm.error = 2;       // reported by the host like a peer time-out
  }
  phase ^= 1u;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();                            // the next item overwrites the staging area and the accumulator
  DBG_T(a, step, 45);
}

// ------------------------------------------------------------------------------------------
// P3: replay statistics, Cmax annealing, ReF-ER beta update; writes ctrl[(step+1)&1]
// ------------------------------------------------------------------------------------------
// Episode::updateCumulative_atomic / updateValues_atomic applied in sample order, one thread
// per run of samples that share an episode (samples are sorted, so runs are contiguous).
// Records are staged through shared memory 256 at a time.  Returns this thread's share of the
// exact far-policy flag changes.
// Incremental statistics (persistent kernel, see stats_incremental): what the thread that owns a run of samples of one episode
// adds to the launch-resident sums / maxima, and the episode's new far-policy term at its position of the episode vector.
struct StatInc { float* xsAll; const int* epPos; double d[4]; float mx[3]; };

__device__ int apply_sample_records(const StepArgs& a, float* stage /* >= 256*12 floats */, StatInc* inc = nullptr) {
  const ReplayView& rp = a.rp;
  const int ME = rp.maxEpisodes;
  int farDelta = 0;
  constexpr int CH = 256;
  for (int c0 = 0; c0 < a.B; c0 += CH) {
    const int b = c0 + threadIdx.x;
    const int cend = min(a.B, c0 + CH);
    const bool mine = threadIdx.x < CH && b < a.B;
    int slot = -1, prevSlot = -2;
    if (mine) {
      const int4 h = __ldcg(reinterpret_cast<const int4*>(&a.rec[b]));
      const float4 d = __ldcg(reinterpret_cast<const float4*>(&a.rec[b].dKL));
      const float4 q = __ldcg(reinterpret_cast<const float4*>(&a.rec[b].qOld));
      if (b > 0) prevSlot = __ldcg(&a.rec[b - 1].slot);
      slot = h.x; farDelta += h.z;
      float4* st = reinterpret_cast<float4*>(stage + threadIdx.x * 12);
      st[0] = make_float4(__int_as_float(h.x), __int_as_float(h.y), 0.f, 0.f); st[1] = d; st[2] = q;
    }
    __syncthreads();
    if (mine && prevSlot != slot) {
      float avgKL = rp.epAgg[AGG_KL * ME + slot], frac = rp.epAgg[AGG_FAR * ME + slot];
      float avgE2 = rp.epAgg[AGG_E2 * ME + slot], maxE = rp.epAgg[AGG_MAXE * ME + slot];
      float sQ2 = rp.epAgg[AGG_Q2 * ME + slot], sQ = rp.epAgg[AGG_Q1 * ME + slot];
      float maxQ = rp.epAgg[AGG_MAXQ * ME + slot], minQ = rp.epAgg[AGG_MINQ * ME + slot];
      const float Nf = (float)rp.epLen[slot];
      const int pos = inc ? inc->epPos[slot] : 0;     // issued with the aggregate loads, used after the run
      const float invN = 1.0f / Nf;
      const float oKL = avgKL, oE2 = avgE2, oQ2 = sQ2, oQ1 = sQ;
      for (int j = b; j < a.B; ++j) {
        int sj, hn; float4 d, q;
        if (j < cend) {
          const float4* st = reinterpret_cast<const float4*>(stage + (j - c0) * 12);
          const float4 h = st[0]; sj = __float_as_int(h.x); hn = __float_as_int(h.y); d = st[1]; q = st[2];
        } else {   // run continues past the staged chunk
          const int4 h = __ldcg(reinterpret_cast<const int4*>(&a.rec[j])); sj = h.x; hn = h.y;
          d = __ldcg(reinterpret_cast<const float4*>(&a.rec[j].dKL)); q = __ldcg(reinterpret_cast<const float4*>(&a.rec[j].qOld));
        }
        if (sj != slot) break;
        if (hn) {
          sQ2 += q.w * q.w - q.z * q.z; sQ += q.w - q.z;
          maxQ = fmaxf(maxQ, q.w); minQ = fminf(minQ, q.w);
        }
        avgKL += invN * d.x; frac += invN * d.y; avgE2 += invN * d.z; maxE = fmaxf(maxE, d.w);
        sQ2 += q.y * q.y - q.x * q.x; sQ += q.y - q.x;
        maxQ = fmaxf(maxQ, q.y); minQ = fminf(minQ, q.y);
      }
      rp.epAgg[AGG_KL * ME + slot] = avgKL; rp.epAgg[AGG_FAR * ME + slot] = frac;
      rp.epAgg[AGG_E2 * ME + slot] = avgE2; rp.epAgg[AGG_MAXE * ME + slot] = maxE;
      rp.epAgg[AGG_Q2 * ME + slot] = sQ2; rp.epAgg[AGG_Q1 * ME + slot] = sQ;
      rp.epAgg[AGG_MAXQ * ME + slot] = maxQ; rp.epAgg[AGG_MINQ * ME + slot] = minQ;
      if (inc) {      // the same float products the full scan of stats_and_refer forms, old term out, new term in
        inc->d[0] += (double)(Nf * avgKL) - (double)(Nf * oKL); inc->d[1] += (double)(Nf * avgE2) - (double)(Nf * oE2);
        inc->d[2] += (double)sQ2 - (double)oQ2; inc->d[3] += (double)sQ - (double)oQ1;
        inc->mx[0] = fmaxf(inc->mx[0], maxE); inc->mx[1] = fmaxf(inc->mx[1], maxQ); inc->mx[2] = fmaxf(inc->mx[2], -minQ);
        inc->xsAll[pos] = Nf * frac;
      }
    }
    __syncthreads();
  }
  return farDelta;
}

// updateTrainingStatistics reductions + updateCounters (MemoryProcessing.cpp:46-92,187-259).
// `sweep` != nullptr on the every-1000-steps recompute: Retrace error sums come from the sweep.
constexpr int kStatChunk = 4096;   // episode positions staged per pass (floats of shared memory)

// `Uint += float` for the common case of a small count and a small non-negative addend: 32-bit conversions (one instruction
// each) give what uint_plus_float_x86 gives — (float)n is exact below 2^24 and truncation agrees on [0, 2^31)
__host__ __device__ __forceinline__ unsigned long long uint_plus_float_fast(unsigned long long n, float x) {
  if (n < (1ull << 24)) {
#ifdef __CUDA_ARCH__
    const float f = (float)(unsigned)n + x;
    if (f >= 0.0f && f < 2147483648.0f) return (unsigned long long)__float2uint_rz(f);
#else
    const volatile float fv = (float)(unsigned)n + x;      // volatile: a plain f32 sum on the host as well
    const float f = fv;
    if (f >= 0.0f && f < 2147483648.0f) return (unsigned long long)(unsigned)f;
#endif
  }
  return uint_plus_float_x86(n, x);
}

// One virtual OpenMP thread's chain `n += xs[p]` over positions p, p + T, ... < n_pos (MemoryProcessing.cpp:202-227).  While the
// count is below 2^24 and the terms are non-negative the integer round trip of every addition is a float truncation:
// (float)n is exact, cvttss2si(s) == trunc(s), and the next (float)n' is that same truncated value — so the chain runs as
// fadd + trunc on a float (a dozen cycles per term instead of two int<->float conversions); the count only grows, so checking
// the FINAL value proves every intermediate one was in range.  Anything else (a negative or NaN term, 2^24 reached) is redone
// with the exact x86 emulation from where it started.
__host__ __device__ __forceinline__ unsigned long long far_chain(unsigned long long n, const float* xs, int p, int n_pos, int T) {
  if (n < (1ull << 24)) {
    float f = (float)(unsigned)n;
    bool bad = false;
    int q = p;
#ifdef __CUDA_ARCH__
#pragma unroll 8
    for (; q < n_pos; q += T) { const float x = xs[q]; bad |= !(x >= 0.0f); f = truncf(f + x); }
#else
    for (; q < n_pos; q += T) { const float x = xs[q]; bad |= !(x >= 0.0f); const volatile float sv = f + x; f = truncf(sv); }
#endif
    if (!bad && f < 16777216.0f) return (unsigned long long)f;
  }
  for (; p < n_pos; p += T) n = uint_plus_float_fast(n, xs[p]);
  return n;
}

// epCache: optional shared-memory copy of {slot, episode length} per position of the episode vector (constant during a launch)
// Launch-resident state of the statistics CTA of the persistent kernel (the episode table is fixed during a launch, the
// per-episode maxima only grow between two sweeps): the sums and maxima of the last full scan, kept up to date by the deltas of
// the episodes each step touches.
struct StatKeep { double sum[5]; float mx[3]; int valid; };

__device__ void refer_tail(const StepArgs& a, const Hyper& hp, const StepCtrl& c, StepCtrl& nx, int step, const SweepSums* sweep,
                           long long farExactOverride, double sumDKL, double sumE2, double sumQ2, double sumQ1, double sumR, double farD,
                           float maxAbsE, float maxQ, float negMinQ, unsigned long long tot);

__device__ void stats_and_refer(const StepArgs& a, const Hyper& hp, const StepCtrl& c, StepCtrl& nx, int step,
                                const SweepSums* sweep, long long farExactOverride, int farDeltaMine, float* xs,
                                const int2* epCache = nullptr, StatKeep* keep = nullptr, float* xsAll = nullptr) {
  __shared__ double shd[kST / 32][6];
  __shared__ float shf[kST / 32][3];
  __shared__ unsigned long long shn[kST];
  const ReplayView& rp = a.rp;
  const int ME = rp.maxEpisodes, nEp = a.nEpisodes, tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  double sumDKL = 0, sumE2 = 0, sumQ2 = 0, sumQ1 = 0, sumR = 0, farD = (double)farDeltaMine;
  float maxAbsE = -1e9f, maxQ = -1e9f, negMinQ = -1e9f;
  // `Uint nOffPol += float` inside an OpenMP reduction with schedule(static,1): thread j of T
  // accumulates episodes j, j+T, ... with a float round trip per addition, partial counts are
  // then added as integers (MemoryProcessing.cpp:202-227).
  const int T = hp.referThreads;
  unsigned long long nOff = 0;
  for (int base = 0; base < nEp; base += kStatChunk) {
    const int n = min(kStatChunk, nEp - base);
    for (int p0 = tid; p0 < n; p0 += 4 * kST) {
      int sl[4]; float Ns[4], far[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int p = p0 + u * kST;
        sl[u] = -1; Ns[u] = 0.f;
        if (p < n) {
          if (epCache) { const int2 e = epCache[base + p]; sl[u] = e.x; Ns[u] = (float)e.y; }
          else sl[u] = rp.epOrder[base + p];
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (sl[u] < 0) continue;
        const int slot = sl[u];
        if (!epCache) Ns[u] = (float)rp.epLen[slot];
        far[u] = rp.epAgg[AGG_FAR * ME + slot];
        sumDKL += (double)(Ns[u] * rp.epAgg[AGG_KL * ME + slot]);
        sumE2 += (double)(Ns[u] * rp.epAgg[AGG_E2 * ME + slot]);
        sumQ2 += (double)rp.epAgg[AGG_Q2 * ME + slot];
        sumQ1 += (double)rp.epAgg[AGG_Q1 * ME + slot];
        sumR += (double)rp.epAgg[AGG_TOTR * ME + slot];
        maxAbsE = fmaxf(maxAbsE, rp.epAgg[AGG_MAXE * ME + slot]);
        maxQ = fmaxf(maxQ, rp.epAgg[AGG_MAXQ * ME + slot]);
        negMinQ = fmaxf(negMinQ, -rp.epAgg[AGG_MINQ * ME + slot]);
        xs[p0 + u * kST] = Ns[u] * far[u];
        if (xsAll) xsAll[base + p0 + u * kST] = Ns[u] * far[u];
      }
    }
    __syncthreads();
    if (tid < T) {
      const int p = (tid - base % T + T) % T;     // first position of this chunk owned by virtual thread `tid`
      nOff = far_chain(nOff, xs, p, n, T);        // the reference's `Uint += float`, x86 semantics
    }
    __syncthreads();
  }
  shn[tid] = nOff;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    sumDKL += __shfl_xor_sync(0xffffffffu, sumDKL, o); sumE2 += __shfl_xor_sync(0xffffffffu, sumE2, o);
    sumQ2 += __shfl_xor_sync(0xffffffffu, sumQ2, o); sumQ1 += __shfl_xor_sync(0xffffffffu, sumQ1, o);
    sumR += __shfl_xor_sync(0xffffffffu, sumR, o); farD += __shfl_xor_sync(0xffffffffu, farD, o);
    maxAbsE = fmaxf(maxAbsE, __shfl_xor_sync(0xffffffffu, maxAbsE, o));
    maxQ = fmaxf(maxQ, __shfl_xor_sync(0xffffffffu, maxQ, o));
    negMinQ = fmaxf(negMinQ, __shfl_xor_sync(0xffffffffu, negMinQ, o));
  }
  if (lane == 0) {
    shd[warp][0] = sumDKL; shd[warp][1] = sumE2; shd[warp][2] = sumQ2; shd[warp][3] = sumQ1; shd[warp][4] = sumR; shd[warp][5] = farD;
    shf[warp][0] = maxAbsE; shf[warp][1] = maxQ; shf[warp][2] = negMinQ;
  }
  __syncthreads();
  if (tid == 0) {
    sumDKL = sumE2 = sumQ2 = sumQ1 = sumR = farD = 0.0;
    for (int w = 0; w < kST / 32; ++w) {
      sumDKL += shd[w][0]; sumE2 += shd[w][1]; sumQ2 += shd[w][2]; sumQ1 += shd[w][3]; sumR += shd[w][4]; farD += shd[w][5];
      maxAbsE = fmaxf(maxAbsE, shf[w][0]); maxQ = fmaxf(maxQ, shf[w][1]); negMinQ = fmaxf(negMinQ, shf[w][2]);
    }
    unsigned long long tot = 0;
    for (int j = 0; j < min(T, kST); ++j) tot += shn[j];
    if (keep) {
      keep->sum[0] = sumDKL; keep->sum[1] = sumE2; keep->sum[2] = sumQ2; keep->sum[3] = sumQ1; keep->sum[4] = sumR;
      keep->mx[0] = maxAbsE; keep->mx[1] = maxQ; keep->mx[2] = negMinQ; keep->valid = 1;
    }
    refer_tail(a, hp, c, nx, step, sweep, farExactOverride, sumDKL, sumE2, sumQ2, sumQ1, sumR, farD, maxAbsE, maxQ, negMinQ, tot);
  }
}

// the scalar end of updateTrainingStatistics + updateCounters (MemoryProcessing.cpp:46-92,229-259), one thread
__device__ void refer_tail(const StepArgs& a, const Hyper& hp, const StepCtrl& c, StepCtrl& nx, int step, const SweepSums* sweep,
                           long long farExactOverride, double sumDKL, double sumE2, double sumQ2, double sumQ1, double sumR, double farD,
                           float maxAbsE, float maxQ, float negMinQ, unsigned long long tot) {
  const int nEp = a.nEpisodes;
  {
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
    nx.n_far_exact = farExactOverride >= 0 ? farExactOverride : c.n_far_exact + (long long)farD;
    // updateCounters: beta fixed-point iteration (MemoryProcessing.cpp:73-85); it runs after
    // applyEpisodesRemovalAlgo, so nStoredSteps() is the post-pruning count
    double nPost = (double)(step == a.lastStep ? a.nTransitionsPost : a.nTransitions);
    double farGlobal = (double)(long long)tot;
    nx.gl_far_prev = farGlobal; nx.gl_stored_prev = nPost; nx.cnt_seed_step = c.cnt_seed_step;
    if (a.comm.world > 1) {
      // DelayedReductor<long> (MemoryProcessing.cpp:48-58): the sums over learner ranks that reach
      // updateCounters are those of the PREVIOUS step.  Push this step's local counts, consume the
      // peers' counts of the previous step.
      const CommView& cm = a.comm;
      const int N = cm.world, me = cm.rank;
      const unsigned stamp = (unsigned)(step + 1);
      // stamped 64-bit words (32 value bits + the step stamp, one NVLink store each: no fence, no
      // separate flag): [far lo, far hi, stored lo, stored hi] per (step & 3, rank).  Four slots:
      // the statistics CTAs run asynchronously, a peer may be up to two steps ahead of this reader.
      const unsigned long long fw = (unsigned long long)(long long)farGlobal, sw = (unsigned long long)(long long)nPost;
      for (int q = 0; q < N; ++q) {
        if (q == me) continue;
        unsigned long long* d = reinterpret_cast<unsigned long long*>(cm.cnt(q)) + ((size_t)(step & 3) * N + me) * 4;
        st_volatile_u64(d + 0, ll_pack((unsigned)fw, stamp)); st_volatile_u64(d + 1, ll_pack((unsigned)(fw >> 32), stamp));
        st_volatile_u64(d + 2, ll_pack((unsigned)sw, stamp)); st_volatile_u64(d + 3, ll_pack((unsigned)(sw >> 32), stamp));
      }
      if ((long long)step == c.cnt_seed_step) { farGlobal = c.gl_far_prev; nPost = c.gl_stored_prev; }   // seeded by the host at (re)start
      else {
        // this rank's own counts of the previous step travel in ctrl (gl_*_prev); the peers' arrived
        // a whole step ago (stamp of the previous step == step)
        const unsigned long long* d = reinterpret_cast<const unsigned long long*>(cm.cnt(me)) + (size_t)((step + 3) & 3) * N * 4;
        unsigned w0[kMaxWorld], w1[kMaxWorld], w2[kMaxWorld], w3[kMaxWorld];
        wait_values_ll(d + 0, 4, N, me, (unsigned)step, cm, w0); wait_values_ll(d + 1, 4, N, me, (unsigned)step, cm, w1);
        wait_values_ll(d + 2, 4, N, me, (unsigned)step, cm, w2); wait_values_ll(d + 3, 4, N, me, (unsigned)step, cm, w3);
        farGlobal = 0.0; nPost = 0.0;
#pragma unroll
        for (int q = 0; q < kMaxWorld; ++q) {
          if (q >= N) continue;
          if (q == me) { farGlobal += c.gl_far_prev; nPost += c.gl_stored_prev; continue; }
          farGlobal += (double)(long long)(((unsigned long long)w1[q] << 32) | w0[q]);
          nPost += (double)(long long)(((unsigned long long)w3[q] << 32) | w2[q]);
        }
      }
    }
    const double fracOff = farGlobal / fmax(nPost, 1.0);
    const double lrB = 0.1 * BS / fmax(maxN, nPost);
    const double b0 = c.beta;
    const double mn = fmin(lrB, b0);
    nx.beta = fracOff > hp.penalTol ? (1.0 - mn) * b0 : (1.0 - mn) * b0 + fmin(lrB, 1.0 - b0);
    // Adam bookkeeping for the next step (Optimizer.cpp:155-158)
    nx.adam_step = c.adam_step + 1;
    double t1 = c.adam_bt1 * 0.9; if (t1 < (double)FLT_EPSILON) t1 = 0; nx.adam_bt1 = t1;
    double t2 = c.adam_bt2 * 0.999; if (t2 < (double)FLT_EPSILON) t2 = 0; nx.adam_bt2 = t2;
    nx.grad_step = c.grad_step + 1;
    nx.adam_eta = adam_eta_for(hp.learnrate, hp.epsAnneal, nx.adam_step, nx.adam_bt1, nx.adam_bt2);
    if (a.statsOut) {
      smb200_step_stats& o = a.statsOut[step - a.stepBase];
      o.beta = nx.beta; o.cmax = nx.cmax; o.cinv = nx.cinv; o.n_far_policy = nx.n_far_ref; o.n_far_exact = nx.n_far_exact;
      o.avg_kl = nx.avg_kl; o.avg_sq_err = nx.avg_sq_err; o.max_abs_err = nx.max_abs_err; o.avg_return = nx.avg_return;
      o.stdev_q = nx.stdev_q; o.avg_q = nx.avg_q; o.max_q = nx.max_q; o.min_q = nx.min_q;
      o.sum_ret_err = nx.sum_ret_err; o.cnt_ret = nx.cnt_ret; o.grad_step = nx.grad_step;
    }
  }
}

// Statistics of a step of the persistent kernel after the first one of a launch: the sums move by what apply_sample_records
// changed (StatInc), the maxima can only have grown, the far-policy count walks the launch-resident terms `xsAll` (shared memory)
// with the reference's `Uint += float` chains (MemoryProcessing.cpp:202-227) — no pass over the episode aggregates in HBM.
__device__ void stats_incremental(const StepArgs& a, const Hyper& hp, const StepCtrl& c, StepCtrl& nx, int step, int farDeltaMine,
                                  const StatInc& inc, StatKeep& keep, const float* xsAll) {
  __shared__ double shd[kST / 32][5];
  __shared__ float shf[kST / 32][3];
  __shared__ unsigned long long shn[kST];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nEp = a.nEpisodes, T = hp.referThreads;
  double d0 = inc.d[0], d1 = inc.d[1], d2 = inc.d[2], d3 = inc.d[3], farD = (double)farDeltaMine;
  float m0 = inc.mx[0], m1 = inc.mx[1], m2 = inc.mx[2];
  __syncthreads();                                   // xsAll entries written by the run owners
  unsigned long long nOff = 0;
  if (tid < T) nOff = far_chain(0, xsAll, tid, nEp, T);
  shn[tid] = nOff;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    d0 += __shfl_xor_sync(0xffffffffu, d0, o); d1 += __shfl_xor_sync(0xffffffffu, d1, o);
    d2 += __shfl_xor_sync(0xffffffffu, d2, o); d3 += __shfl_xor_sync(0xffffffffu, d3, o);
    farD += __shfl_xor_sync(0xffffffffu, farD, o);
    m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, o)); m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, o));
    m2 = fmaxf(m2, __shfl_xor_sync(0xffffffffu, m2, o));
  }
  if (lane == 0) { shd[warp][0] = d0; shd[warp][1] = d1; shd[warp][2] = d2; shd[warp][3] = d3; shd[warp][4] = farD;
                   shf[warp][0] = m0; shf[warp][1] = m1; shf[warp][2] = m2; }
  __syncthreads();
  if (tid == 0) {
    d0 = d1 = d2 = d3 = farD = 0.0;
    for (int w = 0; w < kST / 32; ++w) {
      d0 += shd[w][0]; d1 += shd[w][1]; d2 += shd[w][2]; d3 += shd[w][3]; farD += shd[w][4];
      m0 = fmaxf(m0, shf[w][0]); m1 = fmaxf(m1, shf[w][1]); m2 = fmaxf(m2, shf[w][2]);
    }
    keep.sum[0] += d0; keep.sum[1] += d1; keep.sum[2] += d2; keep.sum[3] += d3;
    keep.mx[0] = fmaxf(keep.mx[0], m0); keep.mx[1] = fmaxf(keep.mx[1], m1); keep.mx[2] = fmaxf(keep.mx[2], m2);
    unsigned long long tot = 0;
    for (int j = 0; j < min(T, kST); ++j) tot += shn[j];
    refer_tail(a, hp, c, nx, step, nullptr, -1, keep.sum[0], keep.sum[1], keep.sum[2], keep.sum[3], keep.sum[4], farD,
               keep.mx[0], keep.mx[1], keep.mx[2], tot);
  }
}

// keep / xsAll (persistent kernel only): launch-resident statistics state; xsAll holds a.nEpisodes floats of shared memory.
__device__ void p3_stats(const StepArgs& a, const Hyper& hp, const StepCtrl& c, StepCtrl& nx, int step, float* stage,
                         const int2* epCache = nullptr, StatKeep* keep = nullptr, float* xsAll = nullptr) {
  if (keep && xsAll && keep->valid) {
    StatInc inc; inc.xsAll = xsAll; inc.epPos = a.rp.epPos;
    inc.d[0] = inc.d[1] = inc.d[2] = inc.d[3] = 0.0; inc.mx[0] = inc.mx[1] = inc.mx[2] = -1e9f;
    const int fd = apply_sample_records(a, stage, &inc);
    stats_incremental(a, hp, c, nx, step, fd, inc, *keep, xsAll);
    return;
  }
  const int fd = apply_sample_records(a, stage);
  __threadfence_block();
  __syncthreads();
  stats_and_refer(a, hp, c, nx, step, nullptr, -1, fd, stage, epCache, keep, xsAll);
}
// The statistics CTA of the persistent kernel calls it as a real function: the statistics code keeps its own registers instead
// of competing with the worker path of the same kernel.  `a` must NOT be the kernel parameter itself (see the caller).
__device__ __noinline__ void p3_stats_call(const StepArgs& a, const Hyper& hp, const StepCtrl& c, StepCtrl& nx, int step, float* stage,
                                          StatKeep* keep, float* xsAll) {
  p3_stats(a, hp, c, nx, step, stage, nullptr, keep, xsAll);
}


// ------------------------------------------------------------------------------------------
// V(s_{t+1}) of sampled transitions whose successor is the last row of a TRUNCATED episode
// (RACER_train.cpp:22-27: `NET.forward(bID, t+1); MB.setValues(bID, t+1, V)`).  About one sample in
// 1100 at cfg2, i.e. one in four steps has one — and the extra forward pass (4.7 us) used to delay
// the grid barrier for every CTA.  The worker CTAs that own no P1 tile are idle during P1: helper j
// scans this step's samples (bit 31 of sampSlot), takes the flagged samples [TB*j, TB*j + TB), ...,
// pulls the weight image, evaluates the value head and performs the write-back of the reference
// (Episode::updateValues_atomic inputs: old and new Q of row t+1 go to the sample's record).
// Returns true if the image was loaded (the caller tracks the mbarrier parity).
// ------------------------------------------------------------------------------------------
template <int TB, bool SM>
__device__ bool next_state_helper(const StepArgs& a, const NetDesc& net, int step, int helper, int nHelpers, unsigned char* smraw,
                                  const SmemPlan& sp, unsigned parity) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  __shared__ int shCount[kST / 32];
  __shared__ int shList[TB];
  float* img = reinterpret_cast<float*>(smraw + sp.img);
  const float* Wp = SM ? img : a.Wimg;
  float* act = reinterpret_cast<float*>(smraw + sp.act);
  float* red = reinterpret_cast<float*>(smraw + sp.red);
  uint64_t* bars = (SM && a.useTma) ? reinterpret_cast<uint64_t*>(smraw + sp.bars) : nullptr;
  const ReplayView& rp = a.rp;
  const int dS = net.dS;
  const size_t j0 = (size_t)(step - a.stepBase) * a.B;
  bool loaded = false;
  for (int first = helper * TB; ; first += nHelpers * TB) {       // flagged samples [first, first + TB)
    // ---- ranks of the flagged samples: ballot + prefix over the batch ----
    int total = 0;
    if (tid < TB) shList[tid] = -1;
    for (int c0 = 0; c0 < a.B; c0 += kST) {
      const int b = c0 + tid;
      const int flagged = (b < a.B) ? (int)((unsigned)__ldcg(a.sampSlot + j0 + b) >> 31) : 0;
      const unsigned m = __ballot_sync(0xffffffffu, flagged);
      __syncthreads();
      if (lane == 0) shCount[warp] = __popc(m);
      __syncthreads();
      int before = total;
      for (int w = 0; w < warp; ++w) before += shCount[w];
      const int rank = before + __popc(m & ((1u << lane) - 1u));
      if (flagged && rank >= first && rank < first + TB) shList[rank - first] = b;
      for (int w = 0; w < kST / 32; ++w) total += shCount[w];
    }
    __syncthreads();
    if (total <= first) break;
    if (!loaded) { if (SM) load_weight_image(a, net, img, bars, step); loaded = true; }
    // ---- gather + standardise the successor states, forward, value head ----
    for (int idx = tid; idx < dS * TB; idx += kST) {
      const int k = idx / TB, s = idx - k * TB;
      const int b = shList[s];
      float x = 0.f;
      if (b >= 0) {
        const size_t row = (size_t)__ldcg(a.sampRow + j0 + b) + 1;
        x = (ld_cg(rp.S + row * dS + k) - ld_cg(rp.stateMean + k)) * ld_cg(rp.stateScale + k);
      }
      act[idx] = x;
    }
    __syncthreads();
    net_forward<TB, SM>(net, Wp, act, red, first == helper * TB ? bars : nullptr, parity);
    if (tid < TB && shList[tid] >= 0) {
      const int b = shList[tid];
      const size_t row = (size_t)__ldcg(a.sampRow + j0 + b) + 1;
      const float vn = (float)net2v((double)net_out<SM>(net, Wp, act, TB, 0, tid));
      const float qOld = ld_cg(rp.ADV + row) + ld_cg(rp.V + row);
      rp.V[row] = vn; rp.ADV[row] = vn - vn;
      *reinterpret_cast<float2*>(&a.rec[b].qNextOld) = make_float2(qOld, vn);
    }
    __syncthreads();
  }
  return loaded;
}

// ------------------------------------------------------------------------------------------
// kernels
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void load_descs(const StepArgs& a, unsigned char* raw, const NetDesc*& net, const Hyper*& hp) {
  DevDescs* h = reinterpret_cast<DevDescs*>(raw);
  const int nw = (int)(sizeof(DevDescs) / 4);
  const int* src = reinterpret_cast<const int*>(a.descs);
  int* dst = reinterpret_cast<int*>(h);
  for (int i = threadIdx.x; i < nw; i += kST) dst[i] = src[i];
  __syncthreads();
  net = &h->net; hp = &h->hp;
}

__device__ __forceinline__ void init_bars(const StepArgs& a, const NetDesc& net, unsigned char* smraw, size_t barsOff) {
  if (a.useTma && threadIdx.x == 0) {
    uint64_t* bars = reinterpret_cast<uint64_t*>(smraw + barsOff);
    for (int l = 0; l < net.nLayers; ++l) mbar_init(&bars[l], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
}

template <int TB, bool SM, bool DISC = false>
__global__ void __launch_bounds__(kST) k_p1(StepArgs a, int step) {
  extern __shared__ __align__(128) unsigned char smraw[];
  const NetDesc* net; const Hyper* hp;
  load_descs(a, smraw, net, hp);
  const SmemPlan sp = smem_plan(*net, TB, SM);
  __shared__ StepCtrl c;
  if (SM) {
    init_bars(a, *net, smraw, sp.bars);
    load_weight_image(a, *net, reinterpret_cast<float*>(smraw + sp.img), reinterpret_cast<uint64_t*>(smraw + sp.bars));
  }
  __syncthreads();
  p1_tile<TB, SM, DISC>(a, *net, *hp, c, step, blockIdx.x, smraw, sp, 0, true, nullptr, 0, false);
}

// recurrent networks: one sampled transition (its whole BPTT window) per CTA
template <bool SM>
__global__ void __launch_bounds__(kST) k_p1_seq(StepArgs a, int step) {
  extern __shared__ __align__(128) unsigned char smraw[];
  const NetDesc* net; const Hyper* hp;
  load_descs(a, smraw, net, hp);
  const SeqSmem sp = smem_plan_seq(*net, SM);
  __shared__ StepCtrl c;
  if (SM) {
    init_bars(a, *net, smraw, sp.bars);
    load_weight_image(a, *net, reinterpret_cast<float*>(smraw + sp.img), reinterpret_cast<uint64_t*>(smraw + sp.bars));
  }
  __syncthreads();
  p1_seq<SM>(a, *reinterpret_cast<const DevDescs*>(smraw), c, step, blockIdx.x, smraw, sp, 0, true, nullptr, 0);
}

__global__ void __launch_bounds__(kST) k_p2p3(StepArgs a, int step, int skipStats) {
  extern __shared__ __align__(128) unsigned char smraw[];
  const NetDesc* net; const Hyper* hp;
  load_descs(a, smraw, net, hp);
  __shared__ StepCtrl c;
  if (threadIdx.x == 0) load_ctrl(c, &a.ctrl[step & 1]);
  __syncthreads();
  float* tiles = reinterpret_cast<float*>(smraw + ((sizeof(DevDescs) + 15) / 16) * 16);
  if ((int)blockIdx.x < a.nTiles) p2_tile(a, *net, *hp, c, a.tiles[blockIdx.x], tiles, step, blockIdx.x);
  else if (!skipStats) p3_stats(a, *hp, c, a.ctrl[(step + 1) & 1], step, tiles);
}

// Persistent cooperative kernel: the last CTA is the statistics CTA (P3), all others are workers
// (P1 tiles, then P2 tiles, two grid barriers per step).  The statistics CTA never joins the
// barriers: it watches the barrier counter to learn that every worker finished P1 of a step and
// publishes ctrl[(step+1)&1] through `ready`, which the workers only need at their next loss stage
// — so the replay statistics are off the critical path.
template <int TB, bool SM, bool REC, bool DISC = false>
__global__ void __launch_bounds__(kST, 1) k_steps_persistent(StepArgs a, int step0, int nSteps, int skipStatsLast) {
  extern __shared__ __align__(128) unsigned char smraw[];
  const NetDesc* net; const Hyper* hp;
  load_descs(a, smraw, net, hp);
  SmemPlan sp;
  SeqSmem sps;
  if (REC) { sps = smem_plan_seq(*net, SM); sp.img = sps.img; sp.tiles = sps.ws; sp.bars = sps.bars; sp.stage = sps.ws; sp.chunks = sps.ws; }
  else sp = smem_plan(*net, TB, SM);
  __shared__ StepCtrl c;
  const int nw = gridDim.x - 1;                 // worker CTAs
  unsigned* ready = a.barrier + 1;              // local steps whose statistics are published
  float* tiles = reinterpret_cast<float*>(smraw + sp.tiles);
  // recurrent nets: LSTM weight gradients on the tensor cores (one more grid barrier per step, between the K-slice
  // items and the tiles that add them up)
  TcPlan tcp; tcp.nItems = 0;
  const size_t descBytes = ((sizeof(DevDescs) + 15) / 16) * 16;
  if (REC && a.useTc) tcp = tc_plan(*net, a.Bpad, sps.info - descBytes);
  const bool tcOn = REC && a.useTc && tcp.nItems > 0;
  const int barsPerStep = tcOn ? 3 : 2;
  __shared__ __align__(8) uint64_t tcBar;
  __shared__ uint32_t tcTmem;
  unsigned tcPhase = 0;
  if ((int)blockIdx.x == nw) {                  // ---- statistics CTA ----
    // launch-resident statistics (StatKeep): the far-policy terms of all episodes live in the shared memory the workers use for
    // the weight image; larger episode tables (or no image in shared memory) keep the full scan of every step
    __shared__ StatKeep keep;
    // p3_stats is not inlined: it gets a shared-memory copy of the launch arguments (taking the address of the kernel parameter
    // would move every access of the WORKER path from the constant bank to a per-thread stack copy)
    __shared__ StepArgs aS;
    float* xsAll = SM && a.nEpisodes <= net->imgFloats && a.statsIncremental ? reinterpret_cast<float*>(smraw + sp.img) : nullptr;
    if (threadIdx.x == 0) { aS = a; keep.valid = 0; }
    __syncthreads();
    for (int s = 0; s < nSteps; ++s) {
      if (skipStatsLast && s == nSteps - 1) break;
      const int step = step0 + s;
      if (threadIdx.x == 0) {
        const unsigned target = (unsigned)(barsPerStep * s + 1) * (unsigned)nw;      // every worker passed barrier 1 of step s
        while (ld_acquire(a.barrier) < target) { }
        __threadfence();
        load_ctrl(c, &a.ctrl[step & 1]);
      }
      __syncthreads();
      DBG_T(a, step, 6);
      p3_stats_call(aS, *hp, c, aS.ctrl[(step + 1) & 1], step, tiles, xsAll ? &keep : nullptr, xsAll);
      __syncthreads();
      if (threadIdx.x == 0) {
        __threadfence();
        asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(ready), "r"((unsigned)(s + 1)) : "memory");
      }
      DBG_T(a, step, 7);
    }
    return;
  }
  if (tcOn) {       // worker CTAs: 128 columns of tensor memory (the 128 x 128 f32 accumulator) for the whole launch
    if (threadIdx.x == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&tcBar)) : "memory");
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 32) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tcTmem)), "r"(128u) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }
  const int nP1 = REC ? a.B : (a.B + TB - 1) / TB;
  unsigned barTarget = 0;
  float* img = reinterpret_cast<float*>(smraw + sp.img);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smraw + sp.bars);
  const bool doP1 = (int)blockIdx.x < nP1;
  const int tid = threadIdx.x;
  const int dS = net->dS, dA = net->dA, nPair = TB * dA;
  const int nS4 = dS * TB / 4;                      // float4 chunks of this tile's raw states
  // inputs of step s+1 are copied into the shared-memory staging area (cp.async) while P2 of step s runs:
  // possible when every worker owns at most one P1 tile
  const bool pf = !REC && !DISC && doP1 && nP1 <= nw && (dS & 3) == 0 && nS4 <= 2 * kST && nPair <= kST;
  // idle worker CTAs evaluate V(s_{t+1}) of truncated episodes for the P1 CTAs (next_state_helper)
  const int nHelpers = (!REC && nP1 < nw) ? nw - nP1 : 0;
  unsigned helpParity = 0;
  const Staging stg = staging_view<TB>(*net, smraw, sp);
  const int b0 = blockIdx.x * TB;
  const ReplayView& rp = a.rp;
  if (pf) {   // state normalisers are constant during a launch (they change only at sweeps)
    for (int k = tid; k < dS; k += kST) { stg.mean[k] = ld_cg(rp.stateMean + k); stg.scale[k] = ld_cg(rp.stateScale + k); }
  }
  if (SM) init_bars(a, *net, smraw, sp.bars);
  // the first GradTile of this CTA never changes: keep it in registers
  GradTile myTile = {0, 0, 0, 0, 0, 0, {0, 0}};
  if ((int)blockIdx.x < a.nTiles) myTile = a.tiles[blockIdx.x];
  bool staged = false;
  // loop-invariant index arithmetic of the input prefetch (integer divisions)
  const int q4s = max(dS >> 2, 1);
  int pfSi[2], pfC4[2];
#pragma unroll
  for (int u = 0; u < 2; ++u) { const int q = tid + u * kST; pfSi[u] = q / q4s; pfC4[u] = (q - pfSi[u] * q4s) * 4; }
  const int pfPs = tid / dA, pfPi = tid - pfPs * dA;
  for (int s = 0; s < nSteps; ++s) {
    const int step = step0 + s;
    DBG_T(a, step, 0);
    if (SM && doP1) load_weight_image(a, *net, img, bars, step);
    DBG_T(a, step, 37);
    // ring rows of the NEXT step's samples this thread will prefetch (known long before they are needed)
    const bool pfNow = pf && s + 1 < nSteps;
    int nxRowS[2] = {-1, -1}, nxRowT = -1, nxSf = 0, nxRowP = -1;
    if (pfNow) {
      const size_t jb = (size_t)(step + 1 - a.stepBase) * a.B + b0;
      // the index loads below are only consumed after barrier 1 and the compiler is free to issue them there: pull
      // their lines into L2 now so that they cost an L2 hit, not a DRAM round trip, on the path to P2
      if (tid == 64) {
        asm volatile("prefetch.global.L2 [%0];" ::"l"(a.sampRow + jb));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(a.sampSlot + jb));
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int q = tid + u * kST;
        if (q < nS4) { const int si = pfSi[u]; if (b0 + si < a.B) nxRowS[u] = a.sampRow[jb + si]; }
      }
      if (tid < TB && b0 + tid < a.B) { nxRowT = a.sampRow[jb + tid]; nxSf = a.sampSlot[jb + tid]; }
      if (tid < nPair) { const int si = pfPs; if (b0 + si < a.B) nxRowP = a.sampRow[jb + si]; }
    }
    bool first = true;
    DBG_T(a, step, 38);
    for (int t = blockIdx.x; t < nP1; t += nw) {
      if (REC) p1_seq<SM>(a, *reinterpret_cast<const DevDescs*>(smraw), c, step, t, smraw, sps, (unsigned)(s & 1), first, ready, (unsigned)s);
      else p1_tile<TB, SM, DISC>(a, *net, *hp, c, step, t, smraw, sp, (unsigned)(s & 1), first, ready, (unsigned)s, staged, nHelpers > 0);
      first = false;
      __syncthreads();
    }
    if (!REC && nHelpers > 0 && !doP1) {
      if (next_state_helper<TB, SM>(a, *net, step, (int)blockIdx.x - nP1, nHelpers, smraw, sp, helpParity)) helpParity ^= 1u;
    }
    DBG_T(a, step, 5);
    grid_barrier(a.barrier, barTarget, nw);
    DBG_T(a, step, 6);
    // ---- copy the next step's inputs straight into the shared-memory staging area (cp.async: no registers are
    //      held across P2; the staging area was last read by this step's P1).  Raw states, actions and behaviour
    //      policies are immutable during a launch (.ca); the per-transition values other CTAs rewrite every step
    //      (V, A, rho, KL, delta, Q) bypass L1 (.cg moves 16 bytes: the aligned chunk around the element) ----
    float* chunks = reinterpret_cast<float*>(smraw + sp.chunks);
    if (pfNow) {
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int q = tid + u * kST;
        if (q < nS4) {
          float4* dst = reinterpret_cast<float4*>(stg.S) + q;
          if (nxRowS[u] >= 0) cp_async16_ca(dst, rp.S + (size_t)nxRowS[u] * dS + pfC4[u]);
          else *dst = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
      DBG_T(a, step, 39);
      if (tid < TB) {
        const int row = nxRowT, hn = (nxSf >> 31) & 1;
        stg.info[0 * TB + tid] = row < 0 ? 0 : row; stg.info[1 * TB + tid] = nxSf & 0x7fffffff;
        stg.info[2 * TB + tid] = row < 0 ? 0 : hn;  stg.info[3 * TB + tid] = row < 0 ? 0 : 1;
        if (row >= 0) {
          const int c0 = row & ~3, c1 = (row + 1) & ~3;
          float* ch = chunks + tid * 4;
          cp_async16_cg(ch + 0 * TB * 4, rp.V + c0);     cp_async16_cg(ch + 1 * TB * 4, rp.ADV + c0);
          cp_async16_cg(ch + 2 * TB * 4, rp.RHO + c0);   cp_async16_cg(ch + 3 * TB * 4, rp.KL + c0);
          cp_async16_cg(ch + 4 * TB * 4, rp.DELTA + c0); cp_async16_cg(ch + 7 * TB * 4, rp.Q + c0);
          if (hn) { cp_async16_cg(ch + 5 * TB * 4, rp.V + c1); cp_async16_cg(ch + 6 * TB * 4, rp.ADV + c1); }
        }
      }
      DBG_T(a, step, 40);
      if (tid < nPair) {
        if (nxRowP >= 0) {
          const size_t row = nxRowP;
          cp_async4_ca(stg.pair + tid, rp.A + row * dA + pfPi);
          cp_async4_ca(stg.pair + nPair + tid, rp.MU + row * 2 * dA + pfPi);
          cp_async4_ca(stg.pair + 2 * nPair + tid, rp.MU + row * 2 * dA + dA + pfPi);
        } else { stg.pair[tid] = 0.f; stg.pair[nPair + tid] = 0.f; stg.pair[2 * nPair + tid] = 1.f; }
      }
    }
    DBG_T(a, step, 35);
    if (!doP1) {          // tile-only workers still need this step's Adam scalars
      if (tid == 0) load_ctrl(c, &a.ctrl[step & 1]);
      __syncthreads();
    }
    if (tcOn) {
      for (int it = blockIdx.x; it < tcp.nItems; it += nw)
        tc_wgrad_item(a, *net, tcp, it, smraw + descBytes, tcTmem, &tcBar, tcPhase, step);
      grid_barrier(a.barrier, barTarget, nw);
      DBG_T(a, step, 46);
    }
    for (int t = blockIdx.x; t < a.nTiles; t += nw) {
      GradTile gt = myTile;
      if (t != (int)blockIdx.x) gt = a.tiles[t];
      DBG_T(a, step, 36);
      p2_tile(a, *net, *hp, c, gt, tiles, step, t, tcOn ? &tcp : nullptr);
      __syncthreads();
    }
    // ---- the copies have had the whole weight-gradient phase to land; pick each old value out of its chunk ----
    if (pfNow) {
      cp_async_wait_all();
      if (tid < TB) {
        const int row = nxRowT < 0 ? 0 : nxRowT, e0 = row & 3, e1 = (row + 1) & 3, hn = nxRowT < 0 ? 0 : (nxSf >> 31) & 1;
        const float* ch = chunks + tid * 4;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const bool nextRow = j == 5 || j == 6;
          float v = 0.f;
          if (nxRowT >= 0 && (!nextRow || hn)) v = ch[j * TB * 4 + (nextRow ? e1 : e0)];
          stg.old[j * TB + tid] = v;
        }
      }
    }
    staged = pfNow;
    DBG_T(a, step, 7);
    grid_barrier(a.barrier, barTarget, nw);
  }
  if (tcOn) {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tcTmem), "r"(128u) : "memory");
  }
}

// statistics after the every-1000-steps sweep (replaces P3 on that step)
__global__ void __launch_bounds__(kST) k_finalize_sweep(StepArgs a, int step, const SweepSums* sweep) {
  __shared__ Hyper hp; __shared__ StepCtrl c;
  __shared__ float xs[kStatChunk];
  if (threadIdx.x == 0) { hp = a.descs->hp; load_ctrl(c, &a.ctrl[step & 1]); }
  __syncthreads();
  stats_and_refer(a, hp, c, a.ctrl[(step + 1) & 1], step, sweep, sweep->nFarExact, 0, xs);
}

// actor-side / diagnostic forward on caller-provided raw states [n][dS] -> out[n][nOut]
template <int TB>
__global__ void __launch_bounds__(kST) k_forward(StepArgs a, const float* states, int n, float* out) {
  extern __shared__ __align__(128) unsigned char smraw[];
  const NetDesc* net; const Hyper* hp;
  load_descs(a, smraw, net, hp);
  const SmemPlan sp = smem_plan(*net, TB, false);
  float* act = reinterpret_cast<float*>(smraw + sp.act);
  float* red = reinterpret_cast<float*>(smraw + sp.red);
  const int b0 = blockIdx.x * TB, dS = net->dS;
  for (int idx = threadIdx.x; idx < dS * TB; idx += kST) {
    const int k = idx / TB, s = idx - k * TB;
    act[idx] = b0 + s < n ? (states[(size_t)(b0 + s) * dS + k] - a.rp.stateMean[k]) * a.rp.stateScale[k] : 0.f;
  }
  __syncthreads();
  net_forward<TB, false>(*net, a.Wimg, act, red, nullptr, 0);
  for (int idx = threadIdx.x; idx < net->nOut * TB; idx += kST) {
    const int j = idx / TB, s = idx - j * TB;
    if (b0 + s < n) out[(size_t)(b0 + s) * net->nOut + j] = net_out<false>(*net, a.Wimg, act, TB, j, s);
  }
}

// Actor-side policy evaluation of recurrent networks (RACER::selectAction, Learners/RACER.cpp:30-47, on the window
// MemoryBuffer::agentToMinibatch builds: the last min(nnBPTTseq, t) + 1 states of the episode in progress, zero initial
// recurrent state, MemoryBuffer.cpp:440-467): one CTA per agent, raw states[agent][maxLen][dS] (the first lengths[agent]
// rows are the window, oldest first), net outputs of the window's last step -> out[agent][nOut].
template <bool SM>
__global__ void __launch_bounds__(kST) k_forward_seq(StepArgs a, const float* states, const int* lengths, int maxLen, float* out) {
  extern __shared__ __align__(128) unsigned char smraw[];
  const NetDesc* net; const Hyper* hp;
  load_descs(a, smraw, net, hp);
  const SeqSmem sp = smem_plan_seq(*net, SM);
  if (SM) {
    init_bars(a, *net, smraw, sp.bars);
    load_weight_image(a, *net, reinterpret_cast<float*>(smraw + sp.img), reinterpret_cast<uint64_t*>(smraw + sp.bars));
  }
  __syncthreads();
  const SeqPlan& sq = reinterpret_cast<const DevDescs*>(smraw)->seq;
  const float* Wp = SM ? reinterpret_cast<const float*>(smraw + sp.img) : a.Wimg;
  float* ws = reinterpret_cast<float*>(smraw + sp.ws);
  uint64_t* bars = (SM && a.useTma) ? reinterpret_cast<uint64_t*>(smraw + sp.bars) : nullptr;
  const int tid = threadIdx.x, b = blockIdx.x, dS = net->dS;
  for (int idx = tid; idx < sq.red; idx += kST) ws[idx] = 0.0f;     // pads and top vectors
  __syncthreads();
  const int have = max(1, min(lengths[b], maxLen));
  const int len = min(have, net->Tc);                // at most nnBPTTseq + 1 steps: the newest ones
  const float* src = states + ((size_t)b * maxLen + (have - len)) * dS;
  float* X = ws + sq.yOff[0];
  const int xs = sq.yStride[0];
  for (int idx = tid; idx < len * dS; idx += kST) {
    const int k = idx / dS, i = idx - k * dS;
    X[k * xs + i] = (src[(size_t)k * dS + i] - ld_cg(a.rp.stateMean + i)) * ld_cg(a.rp.stateScale + i);
  }
  __syncthreads();
  float* actTop = ws + sq.actTop;
  seq_forward<SM>(*net, sq, Wp, ws, ws + sq.red, actTop, bars, 0, len, len, nullptr, nullptr, 0);
  for (int j = tid; j < net->nOut; j += kST) out[(size_t)b * net->nOut + j] = net_out<SM>(*net, Wp, actTop, 1, j, 0);
}

// ------------------------------------------------------------------------------------------
// host launchers
// ------------------------------------------------------------------------------------------
static size_t p2_smem_bytes() { return ((sizeof(DevDescs) + 15) / 16) * 16 + sizeof(float) * 2 * kTileK * kBCP; }

constexpr size_t kSmemBudget = 200 * 1024;
static size_t plan_total(const NetDesc& net, bool img) { return net.recurrent ? smem_plan_seq(net, img).total : smem_plan(net, 4, img).total; }
bool step_image_in_smem(const NetDesc& net) { return plan_total(net, true) <= kSmemBudget; }
size_t tc_staging_bytes(const NetDesc& net) {
  if (!net.recurrent) return 0;
  return smem_plan_seq(net, step_image_in_smem(net)).info - ((sizeof(DevDescs) + 15) / 16) * 16;
}

size_t step_smem_bytes(const NetDesc& net, int TB) { (void)TB; return plan_total(net, step_image_in_smem(net)); }

int step_kernels_prepare(const NetDesc& net) {
  const bool sm = step_image_in_smem(net);
  const int sOn = (int)plan_total(net, true), sOff = (int)plan_total(net, false);
  if (sOff > (int)kSmemBudget) { set_error_msg("network too wide for the shared-memory tile of the step kernel"); return -1; }
  if (net.recurrent) {
    if (sm) {
      SMB200_CUDA_CHECK(cudaFuncSetAttribute(k_p1_seq<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, sOn));
      SMB200_CUDA_CHECK(cudaFuncSetAttribute(k_forward_seq<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, sOn));
      SMB200_CUDA_CHECK(cudaFuncSetAttribute(k_steps_persistent<4, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, sOn));
    }
    SMB200_CUDA_CHECK(cudaFuncSetAttribute(k_p1_seq<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, sOff));
    SMB200_CUDA_CHECK(cudaFuncSetAttribute(k_forward_seq<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, sOff));
    SMB200_CUDA_CHECK(cudaFuncSetAttribute(k_steps_persistent<4, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, sOff));
  } else {
    if (sm) {
      SMB200_CUDA_CHECK(cudaFuncSetAttribute(k_p1<4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, sOn));
      SMB200_CUDA_CHECK(cudaFuncSetAttribute(k_steps_persistent<4, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, sOn));
    }
    SMB200_CUDA_CHECK(cudaFuncSetAttribute(k_p1<4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, sOff));
    SMB200_CUDA_CHECK(cudaFuncSetAttribute(k_steps_persistent<4, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, sOff));
    if (net.discrete) {
      if (!sm) { set_error_msg("discrete-action network too large for the shared-memory weight image"); return -1; }
      SMB200_CUDA_CHECK(cudaFuncSetAttribute(k_p1<4, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, sOn));
      SMB200_CUDA_CHECK(cudaFuncSetAttribute(k_steps_persistent<4, true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, sOn));
    }
    SMB200_CUDA_CHECK(cudaFuncSetAttribute(k_forward<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_plan(net, 4, false).total));
  }
  SMB200_CUDA_CHECK(cudaFuncSetAttribute(k_p2p3, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p2_smem_bytes()));
  return 0;
}

int launch_step_two_kernels(const StepArgs& a, const NetDesc& net, int step, int skipStats, cudaStream_t st) {
  const bool sm = step_image_in_smem(net);
  const size_t smem = step_smem_bytes(net, 4);
  if (net.recurrent) {
    if (sm) k_p1_seq<true><<<a.B, kST, smem, st>>>(a, step);
    else k_p1_seq<false><<<a.B, kST, smem, st>>>(a, step);
  } else {
    const int nP1 = (a.B + 3) / 4;
    if (net.discrete) k_p1<4, true, true><<<nP1, kST, smem, st>>>(a, step);
    else if (sm) k_p1<4, true><<<nP1, kST, smem, st>>>(a, step);
    else k_p1<4, false><<<nP1, kST, smem, st>>>(a, step);
  }
  k_p2p3<<<a.nTiles + 1, kST, p2_smem_bytes(), st>>>(a, step, skipStats);
  SMB200_CUDA_CHECK(cudaGetLastError());
  return 0;
}

static const void* persistent_fn(const NetDesc& net) {
  const bool sm = step_image_in_smem(net);
  if (net.recurrent) return sm ? (const void*)k_steps_persistent<4, true, true> : (const void*)k_steps_persistent<4, false, true>;
  if (net.discrete) return (const void*)k_steps_persistent<4, true, false, true>;
  return sm ? (const void*)k_steps_persistent<4, true, false> : (const void*)k_steps_persistent<4, false, false>;
}

int persistent_grid(const StepArgs& a, const NetDesc& net, int numSMs) {
  int perSM = 0;
  const cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, persistent_fn(net), kST, step_smem_bytes(net, 4));
  if (e != cudaSuccess || perSM < 1) return 0;
  const int nP1 = net.recurrent ? a.B : (a.B + 3) / 4;
  const int want = max(nP1, a.nTiles) + 1;   // workers + the statistics CTA
  const int g = min(want, numSMs * perSM);
  return g >= 2 ? g : 0;
}

int launch_steps_persistent(const StepArgs& a, const NetDesc& net, int grid, int step0, int nSteps, int skipStatsLast, cudaStream_t st) {
  SMB200_CUDA_CHECK(cudaMemsetAsync(a.barrier, 0, 2 * sizeof(unsigned), st));
  StepArgs aa = a;
  void* args[] = {&aa, &step0, &nSteps, &skipStatsLast};
  SMB200_CUDA_CHECK(cudaLaunchCooperativeKernel(persistent_fn(net), dim3(grid), dim3(kST), args, step_smem_bytes(net, 4), st));
  return 0;
}

int launch_finalize_sweep(const StepArgs& a, int step, const SweepSums* sweep, cudaStream_t st) {
  k_finalize_sweep<<<1, kST, 0, st>>>(a, step, sweep);
  SMB200_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int launch_forward(const StepArgs& a, const NetDesc& net, const float* states, int n, float* out, cudaStream_t st) {
  k_forward<4><<<(n + 3) / 4, kST, smem_plan(net, 4, false).total, st>>>(a, states, n, out);
  SMB200_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int launch_forward_seq(const StepArgs& a, const NetDesc& net, const float* states, const int* lengths, int n, int maxLen, float* out,
                       cudaStream_t st) {
  const bool sm = step_image_in_smem(net);
  const size_t smem = step_smem_bytes(net, 4);
  if (sm) k_forward_seq<true><<<n, kST, smem, st>>>(a, states, lengths, maxLen, out);
  else k_forward_seq<false><<<n, kST, smem, st>>>(a, states, lengths, maxLen, out);
  SMB200_CUDA_CHECK(cudaGetLastError());
  return 0;
}

#include "cluster_step.cuh"
#include "wide_step.cuh"

}  // namespace smb200

// ------------------------------------------------------------------------------------------
// Host builds of scalar device functions (diagnostics for the CPU test suite, tests/test_host_replay.py): the same source
// lines the kernels compile, evaluated by the host compiler (x86-64 has no contraction into FMAs, the device build uses
// -fmad=false: identical IEEE operations).
// ------------------------------------------------------------------------------------------
extern "C" {

// n elements of AdamOptimizer::apply_update (struct Adam, Network/Optimizer.cpp:61-108) exactly as the P2 epilogue applies it:
// eta from adam_eta_for after `adam_step` completed updates with the running beta powers bt1 / bt2, then adam_step per element.
// The far-policy chain of one virtual OpenMP thread as the statistics CTA runs it (far_chain: float truncation while the count
// is below 2^24 and the terms are non-negative, the exact x86 emulation otherwise), compiled for the host.
uint64_t smb200_host_far_chain(uint64_t n0, const float* xs, int32_t first, int32_t n_pos, int32_t stride) {
  if (!xs || first < 0 || stride < 1) return ~0ull;
  return (uint64_t)smb200::far_chain((unsigned long long)n0, xs, first, n_pos, stride);
}

int smb200_host_adam(int64_t n, const float* G, float* W, float* M1, float* M2, double learnrate, double eps_anneal, int64_t adam_step_done,
                     double bt1, double bt2, double nn_lambda, int32_t batch_global) {
  if (n < 0 || !G || !W || !M1 || !M2 || batch_global < 1) return -1;
  smb200::AdamCoef k;
  k.eta = smb200::adam_eta_for(learnrate, eps_anneal, adam_step_done, bt1, bt2);
  k.B1 = 0.9f; k.B2 = 0.999f; k.lambda = (float)nn_lambda; k.fac = (float)(1.0 / (double)batch_global);     // adam_coef
  for (int64_t i = 0; i < n; ++i) smb200::adam_step(k, G[i], W[i], M1[i], M2[i], &W[i], &M1[i], &M2[i]);
  return 0;
}

// scaleNet2V / scaleVdiff (Learners/RACER_common.cpp:23-32) of n network outputs, f64
int smb200_host_value_scaling(int64_t n, const double* x, double* v, double* dvdx) {
  if (n < 0 || !x || !v || !dvdx) return -1;
  for (int64_t i = 0; i < n; ++i) { v[i] = smb200::net2v(x[i]); dvdx[i] = smb200::vdiff(x[i]); }
  return 0;
}

// discrete_sample_loss over a batch: O [B][1 + 2K] f32, act [B], mu [B][K], qret [B]; g [B][1 + 2K], out [B][6]
int smb200_host_discrete_loss(int32_t B, int32_t K, const float* O, const float* act, const float* mu, const float* qret, double beta,
                              double cmax, double cinv, double* g, double* out) {
  if (B < 0 || K < 2 || K > smb200::kMaxOptions || !O || !act || !mu || !qret || !g || !out) return -1;
  for (int b = 0; b < B; ++b)
    smb200::discrete_sample_loss(K, O + (size_t)b * (1 + 2 * K), act[b], mu + (size_t)b * K, qret[b], beta, cmax, cinv,
                                 g + (size_t)b * (1 + 2 * K), out + (size_t)b * 6);
  return 0;
}

}  // extern "C"
