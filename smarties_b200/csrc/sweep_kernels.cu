// sweep_kernels.cu — whole-buffer passes over the HBM replay MemoryBuffer (sm_100a).
//
//   k_sweep    per episode (one warp each): exact recompute of the per-episode aggregates
//              (Episode::updateCumulative, ReplayMemory/Episode.cpp:213-242) fused with the
//              Retrace backward recursion  Q[t] = r~[t+1] + g*(V[t+1] + l*min(1,rho[t+1])*
//              (Q[t+1]-A[t+1]-V[t+1]))  (MemoryProcessing.cpp:23-44,391-400), evaluated as a
//              segmented affine scan: 32 time steps per warp iteration, warp-shuffle prefix
//              composition, coalesced 128-byte loads of r, V, A, rho.  24-32 B per transition.
//   k_moments  streaming sum / sum-of-squares of rewards and of every state component around the
//              current means (updateRewardsStats, MemoryProcessing.cpp:94-185), f64 accumulation,
//              (dS+1)*4 B per transition.
//   k_update_scaling / k_init_episode: the scalar tails of those functions and
//              Episode::finalize + initPreTrainErrorPlaceholder (Episode.cpp:244-273).
// Compiled with -fmad=false (see step_kernels.cu).
#include "step_kernels.cuh"

#include <cfloat>

namespace smb200 {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) k_init_episode(ReplayView rp, int slot, float deltaInit, int haveValues) {
  const int N = rp.epLen[slot];
  const size_t r0 = (size_t)rp.epStart[slot];
  const int ME = rp.maxEpisodes;
  float tot = 0.f;
  for (int t = threadIdx.x; t < N; t += kThreads) {
    const size_t r = r0 + t;
    if (!haveValues) { rp.V[r] = 0.f; rp.ADV[r] = 0.f; }
    rp.Q[r] = 0.f; rp.DELTA[r] = deltaInit; rp.KL[r] = 0.f;
    rp.RHO[r] = t + 1 == N ? 0.f : 1.f;
    rp.rowFlag[r] = (uint8_t)(4 | (t == 0 ? 1 : 0) | (t + 1 == N ? 2 : 0));
    if (t > 0) tot += rp.R[r];
  }
  __shared__ float sh[kThreads / 32];
  tot = warp_sum(tot);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = tot;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int w = 0; w < kThreads / 32; ++w) s += sh[w];
    rp.epAgg[AGG_KL * ME + slot] = 0.f; rp.epAgg[AGG_FAR * ME + slot] = 0.f;
    rp.epAgg[AGG_E2 * ME + slot] = deltaInit * deltaInit; rp.epAgg[AGG_MAXE * ME + slot] = deltaInit;
    rp.epAgg[AGG_Q2 * ME + slot] = 0.f; rp.epAgg[AGG_Q1 * ME + slot] = 0.f;
    rp.epAgg[AGG_MAXQ * ME + slot] = -1e9f; rp.epAgg[AGG_MINQ * ME + slot] = 1e9f;
    rp.epAgg[AGG_TOTR * ME + slot] = s;
  }
}

// ------------------------------------------------------------------------------------------
// One CTA (8 warps) per episode.  The episode is walked from its end in super-chunks of 1024 time
// steps: every thread holds 4 steps (coalesced: position p = j*256 + tid, p = 0 the latest step), so
// a 1000-step episode is ONE pass with all of its loads in flight at once.  The recursion
// Q[t] = a_t + b_t Q[t+1] is composed by a warp-shuffle scan inside each run of 32 steps, the 32 run
// composites are scanned by warp 0 through shared memory, and every element is then re-evaluated with
// the reference's operation order on the scanned Q[t+1].
constexpr int kSweepPer = 4;                         // time steps per thread and super-chunk
constexpr int kSweepChunk = kSweepPer * kThreads;    // 1024

__global__ void __launch_bounds__(kThreads) k_sweep(ReplayView rp, int nEpisodes, int oneSlot, float gamma, float lambda, int gae,
                                                    int recompute, float C, float invC, SweepSums* sums) {
  __shared__ float segA[32], segB[32], segQin[32];
  __shared__ float shCarry;
  __shared__ float shRed[kThreads / 32][8];
  __shared__ int shFar[kThreads / 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int ME = rp.maxEpisodes;
  const double rmean = (double)rp.rew[0], rscale = (double)rp.rew[1];
  const int count = nEpisodes > 0 ? nEpisodes : 1;
  float errAcc = 0.f; long long nRet = 0, nFar = 0;
  for (int pos = blockIdx.x; pos < count; pos += gridDim.x) {
    const int slot = nEpisodes > 0 ? rp.epOrder[pos] : oneSlot;
    const int N = rp.epLen[slot];
    const size_t r0 = (size_t)rp.epStart[slot];
    if (recompute) {   // Episode::updateCumulative (Episode.cpp:213-242)
      const int nd = N - 1;
      int far = 0; float sE2 = 0.f, mAE = -1e9f, mxQ = -1e9f, mnQ = 1e9f, sQ2 = 0.f, sQ1 = 0.f, sKL = 0.f, sR = 0.f;
      for (int t = tid; t < N; t += kThreads) {
        const size_t r = r0 + t;
        sKL += rp.KL[r]; sR += rp.R[r];
        if (t < nd) {
          const float w = rp.RHO[r], d = rp.DELTA[r];
          far += (w > C || w < invC) ? 1 : 0;
          sE2 += d * d; mAE = fmaxf(mAE, fabsf(d));
          const float q = rp.ADV[r] + rp.V[r];
          mxQ = fmaxf(mxQ, q); mnQ = fminf(mnQ, q); sQ2 += q * q; sQ1 += q;
        }
      }
      far = __reduce_add_sync(0xffffffffu, far);
      sE2 = warp_sum(sE2); sQ2 = warp_sum(sQ2); sQ1 = warp_sum(sQ1); sKL = warp_sum(sKL); sR = warp_sum(sR);
      mAE = warp_max(mAE); mxQ = warp_max(mxQ); mnQ = -warp_max(-mnQ);
      __syncthreads();
      if (lane == 0) {
        shFar[warp] = far;
        shRed[warp][0] = sE2; shRed[warp][1] = sQ2; shRed[warp][2] = sQ1; shRed[warp][3] = sKL; shRed[warp][4] = sR;
        shRed[warp][5] = mAE; shRed[warp][6] = mxQ; shRed[warp][7] = mnQ;
      }
      __syncthreads();
      if (tid == 0) {
        far = 0; sE2 = sQ2 = sQ1 = sKL = sR = 0.f; mAE = -1e9f; mxQ = -1e9f; mnQ = 1e9f;
        for (int w = 0; w < kThreads / 32; ++w) {
          far += shFar[w]; sE2 += shRed[w][0]; sQ2 += shRed[w][1]; sQ1 += shRed[w][2]; sKL += shRed[w][3]; sR += shRed[w][4];
          mAE = fmaxf(mAE, shRed[w][5]); mxQ = fmaxf(mxQ, shRed[w][6]); mnQ = fminf(mnQ, shRed[w][7]);
        }
        const float invN = 1.0f / (float)nd;
        rp.epAgg[AGG_FAR * ME + slot] = invN * (float)far;
        rp.epAgg[AGG_E2 * ME + slot] = invN * sE2; rp.epAgg[AGG_MAXE * ME + slot] = mAE;
        rp.epAgg[AGG_Q2 * ME + slot] = sQ2; rp.epAgg[AGG_Q1 * ME + slot] = sQ1;
        rp.epAgg[AGG_MAXQ * ME + slot] = mxQ; rp.epAgg[AGG_MINQ * ME + slot] = mnQ;
        rp.epAgg[AGG_TOTR * ME + slot] = sR; rp.epAgg[AGG_KL * ME + slot] = invN * sKL;
        if (C > 1.0f) nFar += far;
      }
    }
    if (recompute == 2) { __syncthreads(); continue; }     // aggregates only (restart: the stored return estimates are kept)
    // ---- Retrace: updateReturnEstimator(EP, N-2) (MemoryProcessing.cpp:23-44) ----
    __syncthreads();
    if (tid == 0) {
      float c0;   // Q of the last row
      if (rp.epTerm[slot]) c0 = rp.Q[r0 + N - 1];
      else { c0 = rp.V[r0 + N - 1]; rp.Q[r0 + N - 1] = c0; }
      shCarry = c0;
    }
    for (int top = N - 2; top >= 0; top -= kSweepChunk) {     // `top` = latest time step of this super-chunk
      float R[kSweepPer], Vn[kSweepPer], An[kSweepPer], cw[kSweepPer], oldQ[kSweepPer], A[kSweepPer], Bc[kSweepPer];
#pragma unroll
      for (int j = 0; j < kSweepPer; ++j) {
        const int t = top - (j * kThreads + tid);
        float rr = 0.f, w = 0.f;
        Vn[j] = 0.f; An[j] = 0.f; oldQ[j] = 0.f;
        if (t >= 0) {
          const size_t r = r0 + t + 1;
          rr = rp.R[r]; Vn[j] = rp.V[r]; An[j] = rp.ADV[r]; w = rp.RHO[r]; oldQ[j] = rp.Q[r - 1];
          // computeGAE (MemoryProcessing.cpp:411-417): R + gamma*(V + lambda*(Q' - V)) is the Retrace expression below
          // with the clipped importance weight 1 and no advantage term (Q' - 0 - V rounds like Q' - V)
          if (gae) { An[j] = 0.f; w = 1.f; }
        }
        R[j] = t >= 0 ? (float)(((double)rr - rmean) * rscale) : 0.f;      // scaledReward<Fval> (Episode.h:184-189)
        cw[j] = t >= 0 ? lambda * (w < 1.f ? w : 1.f) : 0.f;               // clippedOffPolW (Episode.h:190-194)
        // Q[t] = a + b*Q[t+1]; identity map outside the episode
        Bc[j] = t >= 0 ? gamma * cw[j] : 1.f;
        A[j] = t >= 0 ? R[j] + gamma * (Vn[j] - cw[j] * (An[j] + Vn[j])) : 0.f;
      }
      // inclusive composition inside each run of 32 steps (lane 0 = latest time step)
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
#pragma unroll
        for (int j = 0; j < kSweepPer; ++j) {
          const float Ap = __shfl_up_sync(0xffffffffu, A[j], d), Bp = __shfl_up_sync(0xffffffffu, Bc[j], d);
          if (lane >= d) { A[j] = fmaf(Bc[j], Ap, A[j]); Bc[j] = Bc[j] * Bp; }
        }
      }
      if (lane == 31) {
#pragma unroll
        for (int j = 0; j < kSweepPer; ++j) { segA[j * (kThreads / 32) + warp] = A[j]; segB[j * (kThreads / 32) + warp] = Bc[j]; }
      }
      __syncthreads();
      if (warp == 0) {      // the 32 run composites, in time order: exclusive scan seeded with the incoming carry
        float sa = segA[lane], sb = segB[lane];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          const float Ap = __shfl_up_sync(0xffffffffu, sa, d), Bp = __shfl_up_sync(0xffffffffu, sb, d);
          if (lane >= d) { sa = fmaf(sb, Ap, sa); sb = sb * Bp; }
        }
        const float qEnd = fmaf(sb, shCarry, sa);          // Q after run `lane`
        float qin = __shfl_up_sync(0xffffffffu, qEnd, 1);
        if (lane == 0) qin = shCarry;
        segQin[lane] = qin;
      }
      __syncthreads();
      float lastQ = 0.f; int lastT = -1;
#pragma unroll
      for (int j = 0; j < kSweepPer; ++j) {
        const int t = top - (j * kThreads + tid);
        const float qin = segQin[j * (kThreads / 32) + warp];
        const float Qscan = fmaf(Bc[j], qin, A[j]);
        // re-evaluate with the reference's operation order on the scanned Q[t+1]
        float Qn = __shfl_up_sync(0xffffffffu, Qscan, 1);
        if (lane == 0) Qn = qin;
        const float Qt = R[j] + gamma * (Vn[j] + cw[j] * (Qn - An[j] - Vn[j]));
        if (t >= 0) {
          rp.Q[r0 + t] = Qt;
          const float dq = oldQ[j] - Qt;
          errAcc += dq * dq;
          if (t == top - (kSweepChunk - 1) || t == 0) { lastQ = Qt; lastT = t; }
        }
      }
      __syncthreads();
      if (lastT >= 0 && lastT == max(0, top - (kSweepChunk - 1))) shCarry = lastQ;     // earliest step of the super-chunk
      __syncthreads();
    }
    nRet += N - 1;
  }
  if (sums) {
    const float e = warp_sum(errAcc);
    if (lane == 0) atomicAdd(&sums->sumErr2, (double)e);
    if (tid == 0) {
      atomicAdd(reinterpret_cast<unsigned long long*>(&sums->nRet), (unsigned long long)nRet);
      if (recompute) atomicAdd(reinterpret_cast<unsigned long long*>(&sums->nFarExact), (unsigned long long)nFar);
    }
  }
}

// ------------------------------------------------------------------------------------------
// Scalar pieces of the return estimators, host + device (the host build is pinned to the oracle by
// tests/test_host_replay.py through smb200_host_return_estimator; -fmad=false keeps the device build on the same IEEE operations).
__host__ __device__ __forceinline__ float scaled_reward(float r, double rmean, double rscale) {      // scaledReward<Fval>, Episode.h:184-189
  return (float)(((double)r - rmean) * rscale);
}
__host__ __device__ __forceinline__ float lambda_clipped_weight(float lambda, float w) {            // lambda * clippedOffPolW, Episode.h:190-194
  return lambda * (w < 1.f ? w : 1.f);
}
// computeRetrace (MemoryProcessing.cpp:391-400) on the row t+1: R, V, A of that row, cw = lambda * min(1, rho), Qn = Q[t+1]
__host__ __device__ __forceinline__ float retrace_step(float R, float Vn, float An, float cw, float Qn, float gamma) {
  return R + gamma * (Vn + cw * (Qn - An - Vn));
}
// computeRetraceExplBonus (:402-409): C * (|Q' - A - V| - B) + computeRetrace, C = 1 - gamma, B = stats.maxAbsError
__host__ __device__ __forceinline__ float retrace_explore_step(float R, float Vn, float An, float cw, float Qn, float gamma, float coef,
                                                               float baseline) {
  const float E = fabsf(Qn - An - Vn) - baseline;
  return coef * E + retrace_step(R, Vn, An, cw, Qn, gamma);
}

// ------------------------------------------------------------------------------------------
// "returnsEstimator": "retraceExplore" (computeRetraceExplBonus, MemoryProcessing.cpp:402-409):
//   Q[t] = (1 - g) * (|Q[t+1] - A[t+1] - V[t+1]| - baseline) + Retrace(t),  baseline = stats.maxAbsError when the estimator
// is created (createReturnEstimator, :427-435).  The absolute value makes the recursion non-affine, so there is no scan:
// one CTA per episode stages 1024 time steps (coalesced, all loads in flight at once) in shared memory, ONE thread runs the
// recursion over the staged chunk in the reference's operation order (~8 dependent f32 operations per step), all threads
// write Q back and accumulate the squared change.  Aggregates are recomputed by k_sweep(recompute = 2) beforehand.
// `ctrl` != nullptr: the baseline is read from the device-resident statistics of the step (the every-1000-steps recompute
// is enqueued behind a running launch, the host does not hold the value).
// Pinned on a B200 by the golden vracer_explore (tests/test_gpu_parity.py).
__global__ void __launch_bounds__(kThreads) k_sweep_explore(ReplayView rp, int nEpisodes, int oneSlot, float gamma, float lambda,
                                                            float baselineHost, const StepCtrl* ctrl, SweepSums* sums) {
  __shared__ float sR[kSweepChunk], sV[kSweepChunk], sA[kSweepChunk], sW[kSweepChunk], sQ[kSweepChunk];
  __shared__ float shCarry;
  const int tid = threadIdx.x, lane = tid & 31;
  const double rmean = (double)rp.rew[0], rscale = (double)rp.rew[1];
  const float baseline = ctrl ? (float)ctrl->max_abs_err : baselineHost;      // `const Fval baseline = RM.stats.maxAbsError`
  const float coef = 1.0f - gamma;                                            // `const Fval coef = (1-gamma)`
  const int count = nEpisodes > 0 ? nEpisodes : 1;
  float errAcc = 0.f; long long nRet = 0;
  for (int pos = blockIdx.x; pos < count; pos += gridDim.x) {
    const int slot = nEpisodes > 0 ? rp.epOrder[pos] : oneSlot;
    const int N = rp.epLen[slot];
    const size_t r0 = (size_t)rp.epStart[slot];
    __syncthreads();
    if (tid == 0) {      // updateReturnEstimator (MemoryProcessing.cpp:23-33): Q of the last row
      float c0;
      if (rp.epTerm[slot]) c0 = rp.Q[r0 + N - 1];
      else { c0 = rp.V[r0 + N - 1]; rp.Q[r0 + N - 1] = c0; }
      shCarry = c0;
    }
    for (int top = N - 2; top >= 0; top -= kSweepChunk) {     // `top` = latest time step of this chunk, position p = top - t
      const int n = min(kSweepChunk, top + 1);
      float oldQ[kSweepPer];
#pragma unroll
      for (int j = 0; j < kSweepPer; ++j) {
        const int p = j * kThreads + tid;
        oldQ[j] = 0.f;
        if (p < n) {
          const size_t r = r0 + (top - p) + 1;
          const float w = rp.RHO[r];
          sR[p] = scaled_reward(rp.R[r], rmean, rscale);
          sV[p] = rp.V[r]; sA[p] = rp.ADV[r];
          sW[p] = lambda_clipped_weight(lambda, w);
          oldQ[j] = rp.Q[r - 1];
        }
      }
      __syncthreads();
      if (tid == 0) {
        float Qn = shCarry;
        for (int p = 0; p < n; ++p) {
          const float Qt = retrace_explore_step(sR[p], sV[p], sA[p], sW[p], Qn, gamma, coef, baseline);
          sQ[p] = Qt; Qn = Qt;
        }
        shCarry = Qn;
      }
      __syncthreads();
#pragma unroll
      for (int j = 0; j < kSweepPer; ++j) {
        const int p = j * kThreads + tid;
        if (p < n) {
          const float Qt = sQ[p];
          rp.Q[r0 + (top - p)] = Qt;
          const float dq = oldQ[j] - Qt;
          errAcc += dq * dq;
        }
      }
      __syncthreads();
    }
    nRet += N - 1;
  }
  if (sums) {
    const float e = warp_sum(errAcc);
    if (lane == 0) atomicAdd(&sums->sumErr2, (double)e);
    if (tid == 0) atomicAdd(reinterpret_cast<unsigned long long*>(&sums->nRet), (unsigned long long)nRet);
  }
}

// ------------------------------------------------------------------------------------------
constexpr int kMomRows = 64;   // rows per tile

__global__ void __launch_bounds__(kThreads) k_moments(ReplayView rp, long long rowEnd, SweepSums* sums) {
  extern __shared__ __align__(16) float tile[];   // [kMomRows][dS]
  __shared__ uint8_t flags[kMomRows];
  __shared__ double shd[kThreads / 32];
  const int dS = rp.dS, tid = threadIdx.x;
  const int CW = dS < kThreads ? dS : kThreads;       // column threads per group
  const int G = kThreads / CW;                          // row groups
  const int g = tid / CW, c0 = tid - g * CW;
  const int nColPer = (dS + CW - 1) / CW;               // columns per thread when dS > 256
  double s1[2] = {0.0, 0.0}, s2[2] = {0.0, 0.0};        // supports dS <= 512
  double rs1 = 0.0, rs2 = 0.0, cnt = 0.0;
  const double rmean = (double)rp.rew[0];
  const long long nTiles = (rowEnd + kMomRows - 1) / kMomRows;
  for (long long tIdx = blockIdx.x; tIdx < nTiles; tIdx += gridDim.x) {
    const long long row0 = tIdx * kMomRows;
    const int nr = (int)min((long long)kMomRows, rowEnd - row0);
    __syncthreads();
    const float* src = rp.S + (size_t)row0 * dS;
    const int nf = nr * dS;
    if ((dS & 3) == 0) {
      for (int i = tid * 4; i < nf; i += kThreads * 4)
        *reinterpret_cast<float4*>(tile + i) = __ldcs(reinterpret_cast<const float4*>(src + i));
    } else {
      for (int i = tid; i < nf; i += kThreads) tile[i] = __ldcs(src + i);
    }
    if (tid < nr) {
      const uint8_t f = rp.rowFlag[row0 + tid];
      flags[tid] = f;
      if ((f & 4) && !(f & 1)) {            // rewards of rows 1..N-1
        const double dr = (double)__ldcs(rp.R + row0 + tid) - rmean;
        rs1 += dr; rs2 += dr * dr;
      }
      if ((f & 4) && !(f & 2)) cnt += 1.0;  // data rows 0..N-2
    }
    __syncthreads();
    if (g < G) {
      for (int r = g; r < nr; r += G) {
        const uint8_t f = flags[r];
        if (!(f & 4) || (f & 2)) continue;  // states of rows 0..N-2
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int c = c0 + j * CW;
          if (j < nColPer && c < dS) {
            const double x = (double)tile[r * dS + c] - (double)rp.stateMean[c];
            s1[j] += x; s2[j] += x * x;
          }
        }
      }
    }
  }
  // block reduction of column sums over the row groups, then one atomic per column
  __syncthreads();
  double* red = reinterpret_cast<double*>(tile);    // reuse: needs 2*kThreads doubles (checked by the launcher)
  for (int j = 0; j < nColPer && j < 2; ++j) {
    __syncthreads();
    red[tid] = s1[j]; red[kThreads + tid] = s2[j];
    __syncthreads();
    if (g == 0) {
      const int c = c0 + j * CW;
      if (c < dS) {
        double a = 0.0, b = 0.0;
        for (int gg = 0; gg < G; ++gg) { a += red[gg * CW + c0]; b += red[kThreads + gg * CW + c0]; }
        atomicAdd(&sums->moments[c], a); atomicAdd(&sums->moments[dS + c], b);
      }
    }
  }
  // rewards / count
  const int lane = tid & 31, warp = tid >> 5;
  double v[3] = {cnt, rs1, rs2};
  for (int q = 0; q < 3; ++q) {
    const double w = warp_sum_d(v[q]);
    __syncthreads();
    if (lane == 0) shd[warp] = w;
    __syncthreads();
    if (tid == 0) {
      double t = 0.0;
      for (int k = 0; k < kThreads / 32; ++k) t += shd[k];
      atomicAdd(&sums->moments[2 * dS + q], t);
    }
  }
}

// Streaming variant for state widths with dS/4 a power of two (4, 8, 16, 32, 64, ...): no shared-memory
// staging.  Thread (rl, cq) owns the column quad cq of rows rl, rl+RP, ...; a warp reads whole 128-byte
// row segments; U float4 loads per thread are in flight before the first is consumed; f64 accumulation.
constexpr int kMomU = 8;

__global__ void __launch_bounds__(kThreads, 2) k_moments_v4(ReplayView rp, long long rowEnd, SweepSums* sums) {
  __shared__ double red[kThreads][8];
  __shared__ double shd[kThreads / 32];
  const int dS = rp.dS, CQ = dS >> 2, RP = kThreads / CQ;
  const int tid = threadIdx.x, cq = tid % CQ, rl = tid / CQ;
  double m[4], s1[4] = {0.0, 0.0, 0.0, 0.0}, s2[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
  for (int c = 0; c < 4; ++c) m[c] = (double)rp.stateMean[cq * 4 + c];
  double rs1 = 0.0, rs2 = 0.0, cnt = 0.0;
  const double rmean = (double)rp.rew[0];
  const long long tileRows = (long long)RP * kMomU;
  const float4* S4 = reinterpret_cast<const float4*>(rp.S);
  for (long long row0 = (long long)blockIdx.x * tileRows; row0 < rowEnd; row0 += (long long)gridDim.x * tileRows) {
    float4 x[kMomU]; unsigned f[kMomU]; float rw[kMomU];
#pragma unroll
    for (int u = 0; u < kMomU; ++u) {
      const long long row = row0 + (long long)u * RP + rl;
      const bool ok = row < rowEnd;
      f[u] = ok ? (unsigned)rp.rowFlag[row] : 0u;
      x[u] = ok ? __ldcs(S4 + row * CQ + cq) : make_float4(0.f, 0.f, 0.f, 0.f);
      rw[u] = (ok && cq == 0) ? __ldcs(rp.R + row) : 0.f;
    }
#pragma unroll
    for (int u = 0; u < kMomU; ++u) {
      if ((f[u] & 4u) && !(f[u] & 2u)) {            // states of rows 0..N-2
        const double d0 = (double)x[u].x - m[0], d1 = (double)x[u].y - m[1], d2 = (double)x[u].z - m[2], d3 = (double)x[u].w - m[3];
        s1[0] += d0; s1[1] += d1; s1[2] += d2; s1[3] += d3;
        s2[0] = fma(d0, d0, s2[0]); s2[1] = fma(d1, d1, s2[1]); s2[2] = fma(d2, d2, s2[2]); s2[3] = fma(d3, d3, s2[3]);
        if (cq == 0) cnt += 1.0;
      }
      if (cq == 0 && (f[u] & 4u) && !(f[u] & 1u)) {  // rewards of rows 1..N-1
        const double dr = (double)rw[u] - rmean;
        rs1 += dr; rs2 = fma(dr, dr, rs2);
      }
    }
  }
#pragma unroll
  for (int c = 0; c < 4; ++c) { red[tid][c] = s1[c]; red[tid][4 + c] = s2[c]; }
  __syncthreads();
  for (int idx = tid; idx < CQ * 8; idx += kThreads) {
    const int q = idx >> 3, v = idx & 7;
    double a = 0.0;
    for (int r = 0; r < RP; ++r) a += red[r * CQ + q][v];
    atomicAdd(&sums->moments[(v < 4 ? 0 : dS) + q * 4 + (v & 3)], a);
  }
  const int lane = tid & 31, warp = tid >> 5;
  double v3[3] = {cnt, rs1, rs2};
  for (int q = 0; q < 3; ++q) {
    const double w = warp_sum_d(v3[q]);
    __syncthreads();
    if (lane == 0) shd[warp] = w;
    __syncthreads();
    if (tid == 0) {
      double t = 0.0;
      for (int k = 0; k < kThreads / 32; ++k) t += shd[k];
      atomicAdd(&sums->moments[2 * dS + q], t);
    }
  }
}


// ------------------------------------------------------------------------------------------
// The every-1000-steps pass over the whole buffer in ONE kernel (24 B + 132 B per transition): per episode
//   * Retrace / GAE backward recursion (same segmented affine scan and re-evaluation as k_sweep),
//   * exact recompute of the episode aggregates (Episode::updateCumulative) from the SAME loaded rows (+ KL, delta),
//   * reward moments from the rewards the recursion loads anyway,
//   * state moments (updateRewardsStats) from the episode's state rows, which stream into shared memory through a ring of
//     cp.async.bulk copies (TMA, 16 KB per stage, kFuseStages stages per CTA, mbarrier completion): the copies of the next
//     tiles — also of the CTA's NEXT episodes — are in flight while the scan of the current episode runs, so the DRAM pipe
//     never waits for a register-bound load loop.  Episode bounds replace the row flags (no flag reads, no dead ring rows).
// One CTA walks the episodes pos = blockIdx.x, blockIdx.x + gridDim.x, ...; f64 moment accumulators live in registers for
// the whole kernel, one block reduction + one atomic per column at the end.
// Requires dS % 4 == 0 and dS / 4 a power of two <= 256 (else: k_sweep + k_moments).
// ------------------------------------------------------------------------------------------
constexpr int kFuseStages = 6;
constexpr int kFuseTileBytes = 16384;
constexpr int kFuseMaxEp = 512;                 // episodes of one CTA whose (start, length) are cached in shared memory per round

__device__ __forceinline__ uint32_t sw_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void sw_mbar_init(uint64_t* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(sw_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void sw_mbar_expect_tx(uint64_t* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sw_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void sw_mbar_wait(uint64_t* bar, unsigned parity) {
  unsigned ok = 0;
  while (!ok)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(sw_smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void sw_bulk_g2s(void* dst, const void* src, unsigned bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(sw_smem_u32(dst)), "l"(src), "r"(bytes), "r"(sw_smem_u32(bar)) : "memory");
}

__global__ void __launch_bounds__(kThreads, 2) k_sweep_fused(ReplayView rp, int nEpisodes, float gamma, float lambda, int gae,
                                                             float C, float invC, SweepSums* sums) {
  extern __shared__ __align__(128) unsigned char fsm[];
  float* ring = reinterpret_cast<float*>(fsm);                                   // [kFuseStages][kFuseTileBytes]
  __shared__ __align__(8) uint64_t full[kFuseStages];
  __shared__ int epR0[kFuseMaxEp], epN[kFuseMaxEp], epSlot[kFuseMaxEp], epTermS[kFuseMaxEp];
  __shared__ float segA[32], segB[32], segQin[32];
  __shared__ float shCarry;
  __shared__ float shRed[kThreads / 32][8];
  __shared__ int shFar[kThreads / 32];
  __shared__ double redD[kThreads / 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int ME = rp.maxEpisodes, dS = rp.dS, CQ = dS >> 2, RP = kThreads / CQ;
  const int tileRows = kFuseTileBytes / (dS * 4);                                // = 4 * RP
  const int cq = tid % CQ, rl = tid / CQ;
  const double rmean = (double)rp.rew[0], rscale = (double)rp.rew[1];
  double m[4], s1[4] = {0.0, 0.0, 0.0, 0.0}, s2[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
  for (int c = 0; c < 4; ++c) m[c] = (double)rp.stateMean[cq * 4 + c];
  double rs1 = 0.0, rs2 = 0.0, cnt = 0.0;
  float errAcc = 0.f; long long nRet = 0, nFar = 0;
  if (tid == 0) {
    for (int s = 0; s < kFuseStages; ++s) sw_mbar_init(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  const int myCount = nEpisodes > (int)blockIdx.x ? (nEpisodes - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  unsigned long long consumed = 0;               // tiles consumed so far by this CTA (stage = consumed % stages, parity from the count)
  for (int e0 = 0; e0 < myCount; e0 += kFuseMaxEp) {
    const int ne = min(kFuseMaxEp, myCount - e0);
    __syncthreads();                             // the previous round's tables and ring are drained
    for (int i = tid; i < ne; i += kThreads) {
      const int slot = rp.epOrder[blockIdx.x + (size_t)(e0 + i) * gridDim.x];
      epSlot[i] = slot; epR0[i] = rp.epStart[slot]; epN[i] = rp.epLen[slot]; epTermS[i] = rp.epTerm[slot];
    }
    __syncthreads();
    // producer state (thread 0): next tile to issue = (pEp, pTile); the stream of this round's tiles is issued in order
    int pEp = 0, pTile = 0; unsigned long long issued = consumed;
    auto produce = [&]() {                       // thread 0 only
      while (pEp < ne && issued < consumed + kFuseStages) {
        const int nS = epN[pEp] - 1;
        const int nT = (nS + tileRows - 1) / tileRows;
        if (pTile >= nT) { ++pEp; pTile = 0; continue; }
        const int rows = min(tileRows, nS - pTile * tileRows);
        const int st = (int)(issued % kFuseStages);
        const unsigned bytes = (unsigned)rows * (unsigned)dS * 4u;
        sw_mbar_expect_tx(&full[st], bytes);
        sw_bulk_g2s(ring + (size_t)st * (kFuseTileBytes / 4), rp.S + ((size_t)epR0[pEp] + (size_t)pTile * tileRows) * dS, bytes, &full[st]);
        ++issued; ++pTile;
      }
    };
    if (tid == 0) { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); produce(); }
    for (int ei = 0; ei < ne; ++ei) {
      const int slot = epSlot[ei], N = epN[ei];
      const size_t r0 = (size_t)epR0[ei];
      const int nd = N - 1;
      // ---- Retrace + aggregates of the episode (rows in registers; see k_sweep for the scan) ----
      int far = 0; float sE2 = 0.f, mAE = -1e9f, mxQ = -1e9f, mnQ = 1e9f, sQ2 = 0.f, sQ1 = 0.f, sKL = 0.f, sR = 0.f;
      // lane 0 of the LAST warp: Q of the last row (the carry of the recursion) and row 0 of the aggregates (the loop below
      // covers rows 1 .. N-1).  Its loads are in flight together with the first chunk's; shCarry is first read after the
      // chunk's first block barrier.
      float c0 = 0.f, kl0 = 0.f, rr0 = 0.f, w0 = 1.f, d0 = 0.f, a0 = 0.f, v0 = 0.f;
      const bool edge = tid == kThreads - 32;
      if (edge) {
        c0 = epTermS[ei] ? rp.Q[r0 + N - 1] : rp.V[r0 + N - 1];
        kl0 = rp.KL[r0]; rr0 = rp.R[r0]; w0 = rp.RHO[r0]; d0 = rp.DELTA[r0]; a0 = rp.ADV[r0]; v0 = rp.V[r0];
      }
      bool firstChunk = true;
      for (int top = N - 2; top >= 0; top -= kSweepChunk) {
        float R[kSweepPer], Vn[kSweepPer], An[kSweepPer], cw[kSweepPer], oldQ[kSweepPer], A[kSweepPer], Bc[kSweepPer];
#pragma unroll
        for (int j = 0; j < kSweepPer; ++j) {
          const int t = top - (j * kThreads + tid);
          float rr = 0.f, w = 0.f;
          Vn[j] = 0.f; An[j] = 0.f; oldQ[j] = 0.f;
          if (t >= 0) {
            const size_t r = r0 + t + 1;
            rr = rp.R[r]; Vn[j] = rp.V[r]; An[j] = rp.ADV[r]; w = rp.RHO[r]; oldQ[j] = rp.Q[r - 1];
            const float kl = rp.KL[r];
            // aggregates of row t + 1 (Episode::updateCumulative): KL and reward of every row, the rest of data rows only
            sKL += kl; sR += rr;
            const double dr = (double)rr - rmean;           // reward moments: rows 1 .. N-1
            rs1 += dr; rs2 = fma(dr, dr, rs2);
            if (t + 1 < nd) {
              const float d = rp.DELTA[r];
              far += (w > C || w < invC) ? 1 : 0;
              sE2 += d * d; mAE = fmaxf(mAE, fabsf(d));
              const float q = An[j] + Vn[j];
              mxQ = fmaxf(mxQ, q); mnQ = fminf(mnQ, q); sQ2 += q * q; sQ1 += q;
            }
            if (gae) { An[j] = 0.f; w = 1.f; }
          }
          R[j] = t >= 0 ? (float)(((double)rr - rmean) * rscale) : 0.f;
          cw[j] = t >= 0 ? lambda * (w < 1.f ? w : 1.f) : 0.f;
          Bc[j] = t >= 0 ? gamma * cw[j] : 1.f;
          A[j] = t >= 0 ? R[j] + gamma * (Vn[j] - cw[j] * (An[j] + Vn[j])) : 0.f;
        }
        if (firstChunk && edge) {
          if (!epTermS[ei]) rp.Q[r0 + N - 1] = c0;
          shCarry = c0;
          sKL += kl0; sR += rr0;
          if (nd > 0) {
            far += (w0 > C || w0 < invC) ? 1 : 0;
            sE2 += d0 * d0; mAE = fmaxf(mAE, fabsf(d0));
            const float q = a0 + v0;
            mxQ = fmaxf(mxQ, q); mnQ = fminf(mnQ, q); sQ2 += q * q; sQ1 += q;
          }
        }
        firstChunk = false;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
#pragma unroll
          for (int j = 0; j < kSweepPer; ++j) {
            const float Ap = __shfl_up_sync(0xffffffffu, A[j], d), Bp = __shfl_up_sync(0xffffffffu, Bc[j], d);
            if (lane >= d) { A[j] = fmaf(Bc[j], Ap, A[j]); Bc[j] = Bc[j] * Bp; }
          }
        }
        if (lane == 31) {
#pragma unroll
          for (int j = 0; j < kSweepPer; ++j) { segA[j * (kThreads / 32) + warp] = A[j]; segB[j * (kThreads / 32) + warp] = Bc[j]; }
        }
        __syncthreads();
        if (warp == 0) {
          float sa = segA[lane], sb = segB[lane];
#pragma unroll
          for (int d = 1; d < 32; d <<= 1) {
            const float Ap = __shfl_up_sync(0xffffffffu, sa, d), Bp = __shfl_up_sync(0xffffffffu, sb, d);
            if (lane >= d) { sa = fmaf(sb, Ap, sa); sb = sb * Bp; }
          }
          const float qEnd = fmaf(sb, shCarry, sa);
          float qin = __shfl_up_sync(0xffffffffu, qEnd, 1);
          if (lane == 0) qin = shCarry;
          segQin[lane] = qin;
        }
        __syncthreads();
        float lastQ = 0.f; int lastT = -1;
#pragma unroll
        for (int j = 0; j < kSweepPer; ++j) {
          const int t = top - (j * kThreads + tid);
          const float qin = segQin[j * (kThreads / 32) + warp];
          const float Qscan = fmaf(Bc[j], qin, A[j]);
          float Qn = __shfl_up_sync(0xffffffffu, Qscan, 1);
          if (lane == 0) Qn = qin;
          const float Qt = R[j] + gamma * (Vn[j] + cw[j] * (Qn - An[j] - Vn[j]));
          if (t >= 0) {
            rp.Q[r0 + t] = Qt;
            const float dq = oldQ[j] - Qt;
            errAcc += dq * dq;
            if (t == top - (kSweepChunk - 1) || t == 0) { lastQ = Qt; lastT = t; }
          }
        }
        __syncthreads();
        if (lastT >= 0 && lastT == max(0, top - (kSweepChunk - 1))) shCarry = lastQ;
        __syncthreads();
      }
      nRet += N - 1;
      // ---- aggregates: block reduction, written by thread 0 ----
      far = __reduce_add_sync(0xffffffffu, far);
      sE2 = warp_sum(sE2); sQ2 = warp_sum(sQ2); sQ1 = warp_sum(sQ1); sKL = warp_sum(sKL); sR = warp_sum(sR);
      mAE = warp_max(mAE); mxQ = warp_max(mxQ); mnQ = -warp_max(-mnQ);
      if (lane == 0) {
        shFar[warp] = far;
        shRed[warp][0] = sE2; shRed[warp][1] = sQ2; shRed[warp][2] = sQ1; shRed[warp][3] = sKL; shRed[warp][4] = sR;
        shRed[warp][5] = mAE; shRed[warp][6] = mxQ; shRed[warp][7] = mnQ;
      }
      __syncthreads();
      if (tid == 0) {
        far = 0; sE2 = sQ2 = sQ1 = sKL = sR = 0.f; mAE = -1e9f; mxQ = -1e9f; mnQ = 1e9f;
        for (int w = 0; w < kThreads / 32; ++w) {
          far += shFar[w]; sE2 += shRed[w][0]; sQ2 += shRed[w][1]; sQ1 += shRed[w][2]; sKL += shRed[w][3]; sR += shRed[w][4];
          mAE = fmaxf(mAE, shRed[w][5]); mxQ = fmaxf(mxQ, shRed[w][6]); mnQ = fminf(mnQ, shRed[w][7]);
        }
        const float invN = 1.0f / (float)nd;
        rp.epAgg[AGG_FAR * ME + slot] = invN * (float)far;
        rp.epAgg[AGG_E2 * ME + slot] = invN * sE2; rp.epAgg[AGG_MAXE * ME + slot] = mAE;
        rp.epAgg[AGG_Q2 * ME + slot] = sQ2; rp.epAgg[AGG_Q1 * ME + slot] = sQ1;
        rp.epAgg[AGG_MAXQ * ME + slot] = mxQ; rp.epAgg[AGG_MINQ * ME + slot] = mnQ;
        rp.epAgg[AGG_TOTR * ME + slot] = sR; rp.epAgg[AGG_KL * ME + slot] = invN * sKL;
        if (C > 1.0f) nFar += far;
      }
      // ---- state moments of the data rows 0 .. N-2 from the shared-memory ring ----
      const int nS = nd, nT = (nS + tileRows - 1) / tileRows;
      for (int ti = 0; ti < nT; ++ti) {
        const int st = (int)(consumed % kFuseStages);
        sw_mbar_wait(&full[st], (unsigned)((consumed / kFuseStages) & 1));
        const int rows = min(tileRows, nS - ti * tileRows);
        const float4* tile = reinterpret_cast<const float4*>(ring + (size_t)st * (kFuseTileBytes / 4));
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int row = u * RP + rl;
          if (row < rows) {
            const float4 x = tile[row * CQ + cq];
            const double d0 = (double)x.x - m[0], d1 = (double)x.y - m[1], d2 = (double)x.z - m[2], d3 = (double)x.w - m[3];
            s1[0] += d0; s1[1] += d1; s1[2] += d2; s1[3] += d3;
            s2[0] = fma(d0, d0, s2[0]); s2[1] = fma(d1, d1, s2[1]); s2[2] = fma(d2, d2, s2[2]); s2[3] = fma(d3, d3, s2[3]);
          }
        }
        ++consumed;
        __syncthreads();                         // every thread is done with the stage: it can be refilled
        if (tid == 0) { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); produce(); }
      }
      cnt += tid == 0 ? (double)nS : 0.0;
    }
  }
  // ---- block reduction of the moment accumulators, one atomic per column ----
  __syncthreads();
  double* red = reinterpret_cast<double*>(fsm);          // [kThreads][8] doubles = 16 KB of the (drained) ring
#pragma unroll
  for (int c = 0; c < 4; ++c) { red[tid * 8 + c] = s1[c]; red[tid * 8 + 4 + c] = s2[c]; }
  __syncthreads();
  for (int idx = tid; idx < CQ * 8; idx += kThreads) {
    const int q = idx >> 3, v = idx & 7;
    double acc = 0.0;
    for (int r = 0; r < RP; ++r) acc += red[(r * CQ + q) * 8 + v];
    atomicAdd(&sums->moments[(v < 4 ? 0 : dS) + q * 4 + (v & 3)], acc);
  }
  double v3[3] = {cnt, rs1, rs2};
  for (int q = 0; q < 3; ++q) {
    const double w = warp_sum_d(v3[q]);
    __syncthreads();
    if (lane == 0) redD[warp] = w;
    __syncthreads();
    if (tid == 0) {
      double t = 0.0;
      for (int k = 0; k < kThreads / 32; ++k) t += redD[k];
      atomicAdd(&sums->moments[2 * dS + q], t);
    }
  }
  const float e = warp_sum(errAcc);
  if (lane == 0) atomicAdd(&sums->sumErr2, (double)e);
  if (tid == 0) {
    atomicAdd(reinterpret_cast<unsigned long long*>(&sums->nRet), (unsigned long long)nRet);
    atomicAdd(reinterpret_cast<unsigned long long*>(&sums->nFarExact), (unsigned long long)nFar);
  }
}

bool sweep_fused_supported(const ReplayView& rp, int estimator) {
  const int CQ = rp.dS >> 2;
  return estimator != 2 && (rp.dS & 3) == 0 && CQ >= 1 && (CQ & (CQ - 1)) == 0 && CQ <= kThreads && rp.dS * 4 * 4 * (kThreads / CQ) == kFuseTileBytes;
}

int launch_sweep_fused(const ReplayView& rp, int nEpisodes, float gamma, float lambda, int estimator, float cmax, float cinv,
                       SweepSums* sums, int numSMs, cudaStream_t st) {
  static bool attr = false;
  const int smem = kFuseStages * kFuseTileBytes;
  if (!attr) { SMB200_CUDA_CHECK(cudaFuncSetAttribute(k_sweep_fused, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); attr = true; }
  int blocks = nEpisodes < numSMs * 2 ? nEpisodes : numSMs * 2;
  if (blocks < 1) blocks = 1;
  k_sweep_fused<<<blocks, kThreads, smem, st>>>(rp, nEpisodes, gamma, lambda, estimator == 1, cmax, cinv, sums);
  SMB200_CUDA_CHECK(cudaGetLastError());
  return 0;
}

// updateStats lambda of updateRewardsStats (MemoryProcessing.cpp:153-184); the reference
// accumulates in long double, here f64.
__device__ __forceinline__ void update_stats(float& mean, float& stdev, float& invstd, double lr, double Evar, double Evar2) {
  mean = (float)((double)mean + lr * Evar);
  double variance = Evar2 - Evar * Evar * (2.0 * lr - lr * lr);
  variance = fmax(variance, (double)FLT_EPSILON);
  stdev = (float)((double)stdev + lr * (sqrt(variance) - (double)stdev));
  invstd = 1.0f / stdev;
}

__global__ void k_update_scaling(ReplayView rp, const StepCtrl* ctrlCur, const DevDescs* descs, const SweepSums* sums, int bInit) {
  const int dS = rp.dS;
  const double eta = descs->hp.learnrate, eps = descs->hp.epsAnneal;
  const double learnR = eta / (1.0 + (double)ctrlCur->grad_step * eps);       // annealRate(eta, nGradSteps, eps)
  const double w = bInit ? 1.0 : fmin(1.0, 10.0 * learnR);                     // Learner.cpp:83
  const double count = sums->moments[2 * dS];
  for (int k = threadIdx.x; k <= dS; k += blockDim.x) {
    if (k < dS) {
      float m = rp.stateMean[k], s = rp.stateStd[k], is = rp.stateScale[k];
      update_stats(m, s, is, w, sums->moments[k] / count, sums->moments[dS + k] / count);
      rp.stateMean[k] = m; rp.stateStd[k] = s; rp.stateScale[k] = is;
    } else {
      float m = rp.rew[0], is = rp.rew[1], s = rp.rew[2];
      update_stats(m, s, is, w, sums->moments[2 * dS + 1] / count, sums->moments[2 * dS + 2] / count);
      rp.rew[0] = m; rp.rew[1] = is; rp.rew[2] = s;
    }
  }
}

__global__ void k_clear_sums(SweepSums* s) {
  const int n = (int)(sizeof(SweepSums) / 8);
  long long* p = reinterpret_cast<long long*>(s);
  for (int i = threadIdx.x; i < n; i += blockDim.x) p[i] = 0;
}

// Sum of a small f64 vector over the learner ranks through peer memory (push + stamp + local wait,
// slots added in rank order: bit-identical on every rank).  Used for the reward/state moments and
// the replay counters at initialisation and on sweep steps (DelayedReductor, StateRewRdx).
__global__ void __launch_bounds__(kThreads) k_peer_allreduce(CommView cm, double* vec, int n, unsigned stamp) {
  const int N = cm.world, me = cm.rank, par = stamp & 1;
  for (int i = threadIdx.x; i < n; i += kThreads) {
    const double v = vec[i];
    for (int q = 0; q < N; ++q) cm.vec(q)[((size_t)par * N + me) * kCommVec + i] = v;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();
    for (int q = 0; q < N; ++q) asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(cm.vecFlag(q) + me), "r"(stamp) : "memory");
  }
  if (threadIdx.x < N) {
    const long long t0 = clock64();
    unsigned v;
    do {
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(cm.vecFlag(me) + threadIdx.x) : "memory");
      if (clock64() - t0 > cm.timeoutCycles) { *cm.error = 1; break; }
    } while ((int)(v - stamp) < 0);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += kThreads) {
    double s = 0.0;
    for (int q = 0; q < N; ++q) s += __ldcg(cm.vec(me) + ((size_t)par * N + q) * kCommVec + i);
    vec[i] = s;
  }
}

int launch_peer_allreduce(const CommView& comm, double* vec, int n, unsigned stamp, cudaStream_t st) {
  if (comm.world <= 1) return 0;
  if (n > kCommVec) { set_error_msg("peer all-reduce vector too long"); return -1; }
  k_peer_allreduce<<<1, kThreads, 0, st>>>(comm, vec, n, stamp);
  SMB200_CUDA_CHECK(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------
int launch_init_episode(const ReplayView& rp, int slot, float deltaInit, int haveValues, cudaStream_t st) {
  k_init_episode<<<1, kThreads, 0, st>>>(rp, slot, deltaInit, haveValues);
  SMB200_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int launch_sweep(const ReplayView& rp, int nEpisodes, int oneSlot, float gamma, float lambda, int estimator, int recompute, float cmax,
                 float cinv, SweepSums* sums, cudaStream_t st, float exploreBaseline, const StepCtrl* exploreCtrl) {
  const int count = nEpisodes > 0 ? nEpisodes : 1;
  int blocks = count;                           // one CTA per episode, at most one full wave of 8 CTAs per SM
  if (blocks > 148 * 8) blocks = 148 * 8;
  const int gae = estimator == 1;
  if (estimator == 2) {                         // retraceExplore: aggregates by k_sweep, the sequential recursion on its own
    if (recompute) {
      k_sweep<<<blocks, kThreads, 0, st>>>(rp, nEpisodes, oneSlot, gamma, lambda, 0, 2, cmax, cinv, sums);
      SMB200_CUDA_CHECK(cudaGetLastError());
    }
    if (recompute != 2) {
      if (blocks > 148 * 4) blocks = 148 * 4;   // 20 KB of static shared memory per CTA
      k_sweep_explore<<<blocks, kThreads, 0, st>>>(rp, nEpisodes, oneSlot, gamma, lambda, exploreBaseline, exploreCtrl, sums);
      SMB200_CUDA_CHECK(cudaGetLastError());
    }
    return 0;
  }
  k_sweep<<<blocks, kThreads, 0, st>>>(rp, nEpisodes, oneSlot, gamma, lambda, gae, recompute, cmax, cinv, sums);
  SMB200_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int launch_moments(const ReplayView& rp, long long rowEnd, SweepSums* sums, int numSMs, cudaStream_t st) {
  if (rp.dS > 512) { set_error_msg("moments kernel supports dim_state <= 512"); return -1; }
  const int CQ = rp.dS >> 2;
  if ((rp.dS & 3) == 0 && (CQ & (CQ - 1)) == 0 && CQ <= kThreads) {     // streaming variant
    const long long tileRows = (long long)(kThreads / CQ) * kMomU;
    const long long nT = (rowEnd + tileRows - 1) / tileRows;
    long long blocks = nT < (long long)numSMs * 2 ? nT : (long long)numSMs * 2;
    if (blocks < 1) blocks = 1;
    k_moments_v4<<<(int)blocks, kThreads, 0, st>>>(rp, rowEnd, sums);
    SMB200_CUDA_CHECK(cudaGetLastError());
    return 0;
  }
  size_t sm = sizeof(float) * kMomRows * rp.dS;
  if (sm < sizeof(double) * 2 * kThreads) sm = sizeof(double) * 2 * kThreads;
  const long long nTiles = (rowEnd + kMomRows - 1) / kMomRows;
  long long blocks = nTiles < (long long)numSMs * 8 ? nTiles : (long long)numSMs * 8;
  if (blocks < 1) blocks = 1;
  if (sm > 48 * 1024) SMB200_CUDA_CHECK(cudaFuncSetAttribute(k_moments, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
  k_moments<<<(int)blocks, kThreads, sm, st>>>(rp, rowEnd, sums);
  SMB200_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int launch_update_scaling(const ReplayView& rp, const StepCtrl* ctrlCur, const DevDescs* descs, const SweepSums* sums, int bInit, cudaStream_t st) {
  k_update_scaling<<<1, 128, 0, st>>>(rp, ctrlCur, descs, sums, bInit);
  SMB200_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int launch_clear_sums(SweepSums* sums, cudaStream_t st) {
  k_clear_sums<<<1, 256, 0, st>>>(sums);
  SMB200_CUDA_CHECK(cudaGetLastError());
  return 0;
}

}  // namespace smb200

// Host build of the estimator arithmetic above (diagnostics for the CPU test suite): updateReturnEstimator(EP, N-2)
// (MemoryProcessing.cpp:23-44) of ONE episode, sequentially, with the scalar functions the device code calls.
// estimator = smb200_returns_estimator (0 retrace, 1 GAE, 2 retraceExplore); arrays hold the N rows of the episode;
// Q is updated in place; returns the sum of squared changes (f32 accumulation like the reference's sumErr2 of one episode).
extern "C" double smb200_host_return_estimator(int32_t n_rows, int32_t terminated, int32_t estimator, const float* R, const float* V,
                                               const float* ADV, const float* RHO, float* Q, double gamma, double lambda,
                                               float reward_mean, float reward_scale, double max_abs_err) {
  using namespace smb200;
  if (n_rows < 2 || !R || !V || !ADV || !RHO || !Q) return -1.0;
  const float g = (float)gamma, l = (float)lambda, coef = 1.0f - g, baseline = (float)max_abs_err;
  const double rmean = (double)reward_mean, rscale = (double)reward_scale;
  const int N = n_rows;
  if (!terminated) Q[N - 1] = V[N - 1];
  float err2 = 0.f;
  for (int t = N - 2; t >= 0; --t) {
    const float r = scaled_reward(R[t + 1], rmean, rscale);
    const float An = estimator == 1 ? 0.f : ADV[t + 1];
    const float cw = lambda_clipped_weight(l, estimator == 1 ? 1.f : RHO[t + 1]);
    const float Qt = estimator == 2 ? retrace_explore_step(r, V[t + 1], An, cw, Q[t + 1], g, coef, baseline)
                                    : retrace_step(r, V[t + 1], An, cw, Q[t + 1], g);
    const float d = Q[t] - Qt;
    err2 += d * d;
    Q[t] = Qt;
  }
  return (double)err2;
}
