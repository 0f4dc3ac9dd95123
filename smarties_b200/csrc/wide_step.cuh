// wide_step.cuh — the LARGE-BATCH learner step of feed-forward V-RACER / RACER nets on the 5th-generation tensor cores.
// Included by step_kernels.cu (inside namespace smb200, after the tile kernel's device functions).
//
// For mini-batches of >= 1024 sampled transitions the dense products of the step ARE contractions (SURVEY.md §8d batch
// sweep): forward  Y[s][n] = sum_k X[s][k] W[k][n],  input gradient  E[s][k] = sum_n D[s][n] W[k][n],  weight gradient
// dW[k][n] = sum_s X[s][k] D[s][n]  (Network::forward / backProp, Network/Network.h:101-226; BaseLayer, Layers/Layer_Base.h:
// 64-113; Layers.h:123-188).  They run as tcgen05.mma.cta_group::1.kind::tf32 with M = 128 (one TILE of 128 sampled
// transitions, or 128 feature rows) and the f32 accumulator in tensor memory.  Every operand is split into its TF32 "hi" part
// and the f32 remainder "lo" and every product is three MMAs (lo*hi + hi*lo + hi*hi): f32 accuracy, the parity tolerances of
// the tile kernel hold (plain TF32 misses them, profiles/r1/microbench_tcgen05_tf32.txt).
//
// One learner step = stream-ordered kernels (the batch is large, launches are noise):
//   k_wide_fwd    persistent, one CTA per SM, tiles of 128 samples: the forward weights of the whole network stay in shared
//                 memory as pre-split UMMA operand images (K-major, no swizzle; written by the Adam kernel; fetched once per
//                 launch with cp.async.bulk); the tile's activations are the A operand IN TENSOR MEMORY (tcgen05.st by the
//                 epilogue that produced them — they never touch shared memory), accumulator in tensor memory; gather +
//                 standardise, all layers, then the ReF-ER / Retrace loss in f64 (the tile kernel's formulas, one thread per
//                 (sample, action component)), replay write-back and per-sample records; activations and the output gradient
//                 leave for the other kernels as a scratch [tile][feature][128 samples] (coalesced 128-byte warp stores, one
//                 contiguous block per tile).
//   k_wide_next   V(s_t+1) of the few samples whose successor ends a truncated episode (RACER_train.cpp:22-27).
//   k_wide_records / k_wide_stats   Episode::updateCumulative_atomic in sample order — one thread per run of samples of an
//                 episode, all runs in parallel — then the tile kernel's own statistics / ReF-ER beta update on one CTA.
//   k_wide_bwd    the same structure with the TRANSPOSED weight images: deltas of a tile as the TMEM A operand,
//                 E = D W^T per layer, tanh' and the ParametricResidual in the epilogue.
//   k_wide_wgrad  persistent split-K contraction over the samples: per 16-sample stage the operand rows come from the scratch
//                 (float4 = 4 samples of one feature), are split and stored as K-major operand images (double-buffered stages,
//                 freed by tcgen05.commit -> mbarrier), one MMA set per dense layer accumulates in tensor memory for the whole
//                 launch; bias / residual / ParamLayer gradients are summed by the staging threads from the registers they
//                 hold anyway.  One partial record per CTA.
//   k_wide_adam   adds the per-CTA partial records in CTA order (deterministic), the reference's Adam variant (adam_step), and
//                 rewrites every image of the weights: blob, tile-kernel image, split forward and transposed operand images.
#pragma once

// ------------------------------------------------------------------------------------------
// tcgen05 / tensor-memory primitives (forms checked on a B200 by scripts/micro/tc_ts_mn.cu)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n"
               ::"r"(tmem_d), "r"(tmem_a), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tm_ld8(uint32_t addr, uint32_t (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(addr) : "memory");
}
__device__ __forceinline__ void tm_st8(uint32_t addr, const uint32_t (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
               ::"r"(addr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
}
__device__ __forceinline__ void tm_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tm_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// bounded mbarrier wait: a lost MMA completion must not hang the box (reported like a peer time-out)
__device__ __forceinline__ bool mbar_wait_bounded(uint64_t* bar, unsigned parity) {
  for (int spin = 0; spin < (1 << 26); ++spin) if (mbar_try_wait(bar, parity)) return true;
  return false;
}
__host__ __device__ constexpr uint32_t wide_idesc(int n) {      // f32 <- tf32 x tf32, both operands K-major, M = 128 (mma_sm100_desc.hpp:412-439)
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24);
}
__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_free(uint32_t base, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(cols) : "memory");
}

// the plan (host-built, a few hundred bytes) -> shared memory
__device__ __forceinline__ const WidePlan* load_wide_plan(const StepArgs& a, unsigned char* dst) {
  const int nw = (int)(sizeof(WidePlan) / 4);
  const int* src = reinterpret_cast<const int*>(a.wplan);
  int* d = reinterpret_cast<int*>(dst);
  for (int i = threadIdx.x; i < nw; i += kST) d[i] = src[i];
  return reinterpret_cast<const WidePlan*>(dst);
}
constexpr int kWideDescBytes = (int)(((sizeof(DevDescs) + 15) / 16) * 16);
constexpr int kWidePlanBytes = (int)(((sizeof(WidePlan) + 15) / 16) * 16);

// TMEM columns of the forward / input-gradient kernels
constexpr uint32_t kWAcc = 0, kWAh = 128, kWAl = 256, kWCarry = 384;

// one dense product of a 128-sample tile: D[128][Np] = A[128][Kc] (TMEM, hi at kWAh, lo at kWAl) * image (shared memory,
// float4 [Kc/4][rows], rows = N of the MMA), three MMAs per 8 contraction columns.  Issued by ONE thread.
__device__ __forceinline__ void wide_issue(uint32_t tmem, const float* imgHi, const float* imgLo, int Kc, int rows, uint64_t* bar) {
  tc_fence_after();
  const uint32_t idesc = wide_idesc(rows);
  // one descriptor per image, advanced by two k-chunks (2 * rows * 16 bytes, in 16-byte units) per MMA set: the issuing thread
  // is the critical path of a layer, every instruction in this loop counts
  uint64_t dbh = umma_desc(imgHi, rows * 16, 128), dbl = umma_desc(imgLo, rows * 16, 128);
  const uint64_t inc = (uint64_t)(2 * rows);
  uint32_t ah = tmem + kWAh, al = tmem + kWAl;
  const uint32_t acc = tmem + kWAcc;
  const int n = Kc / 8;
  umma_tf32_ts(acc, al, dbh, idesc, 0u);
  umma_tf32_ts(acc, ah, dbl, idesc, 1u);
  umma_tf32_ts(acc, ah, dbh, idesc, 1u);
#pragma unroll 4
  for (int kk = 1; kk < n; ++kk) {
    dbh += inc; dbl += inc; ah += 8; al += 8;
    umma_tf32_ts(acc, al, dbh, idesc, 1u);
    umma_tf32_ts(acc, ah, dbl, idesc, 1u);
    umma_tf32_ts(acc, ah, dbh, idesc, 1u);
  }
  tc_commit(bar);
}

// ------------------------------------------------------------------------------------------
// k_wide_fwd
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kST, 1) k_wide_fwd(StepArgs a, int step, int sfBars, int sfImg, int fFloats) {
  extern __shared__ __align__(128) unsigned char smraw[];
  // the weight image first: its copy (<= 180 KB out of L2) runs under everything else the prologue does
  {
    uint64_t* bars0 = reinterpret_cast<uint64_t*>(smraw + sfBars);
    if (threadIdx.x == 0) {
      mbar_init(&bars0[0], 1); mbar_init(&bars0[1], 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      asm volatile("fence.proxy.async.global;\n\tfence.proxy.async.shared::cta;" ::: "memory");
      mbar_expect_tx(&bars0[0], (unsigned)fFloats * 4u);
      float* img0 = reinterpret_cast<float*>(smraw + sfImg);
      for (int off = 0; off < fFloats; off += 16384) {          // hi halves then lo halves, <= 64 KB per bulk copy
        const unsigned n = (unsigned)min(16384, fFloats - off) * 4u;
        bulk_g2s(img0 + off, a.wimgF + off, n, &bars0[0]);
      }
    }
  }
  const NetDesc* netp; const Hyper* hpp;
  load_descs(a, smraw, netp, hpp);
  const WidePlan& wp = *load_wide_plan(a, smraw + kWideDescBytes);
  __shared__ uint32_t tmemSlot;
  __syncthreads();
  const NetDesc& net = *netp;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q = warp & 3, cg = warp >> 2;          // TMEM lane quarter of this warp (hardware rule: warp % 4), column group
  float* vec = reinterpret_cast<float*>(smraw + wp.sfVec);
  float* img = reinterpret_cast<float*>(smraw + wp.sfImg);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smraw + wp.sfBars);   // [0] weight image, [1] MMA completion
  const ReplayView& rp = a.rp;
  const int dS = net.dS;
  constexpr int TBW = kWideM;
  const LayerDesc& Lo = net.L[net.nLayers - 2];
  const int fHalf = wp.fFloats >> 1;

  if (warp == 0) tmem_alloc(&tmemSlot, 512u);
  for (int i = tid; i < wp.vFloats; i += kST) vec[i] = ld_cg(a.wvec + i);
  for (int i = tid; i < wp.D[0].Kp; i += kST) {
    vec[wp.vMsc + i] = i < net.dS ? ld_cg(a.rp.stateMean + i) : 0.f;
    vec[wp.vMsc + wp.D[0].Kp + i] = i < net.dS ? ld_cg(a.rp.stateScale + i) : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmemSlot;
  const int nTiles = (a.B + TBW - 1) / TBW;
  const size_t j0 = (size_t)(step - a.stepBase) * a.B;
  const bool keep = step == a.lastStep || a.lastStep < 0;
  unsigned mmaPar = 0;
  bool imgReady = false, fault = false;
  const uint32_t laneBase = (uint32_t)(q * 32) << 16;

  // Inputs of a tile, fetched ONE TILE AHEAD into registers (the gather is a dependent chain sample index -> ring row -> state
  // row, ~2 DRAM round trips, that would otherwise open every tile): the thread's <= 2 chunks of 8 state components of its sample.
  const WDense& D0 = wp.D[0];
  const float* msc = vec + wp.vMsc;                // [2][Kp0]: state mean, state scale (Core/StateAction.h:56-58)
  float pfX[2][8];
  const bool vec4S = (dS & 3) == 0;
  auto prefetch = [&](int tile) {
    const int bb = tile * TBW + q * 32 + lane;
    const bool ok = tile < nTiles && bb < a.B;
    const int pfRow = ok ? __ldg(a.sampRow + j0 + bb) : 0;
#pragma unroll
    for (int c2 = 0; c2 < 2; ++c2) {
      const int j8 = cg + 4 * c2;
      if (vec4S) {        // state rows are 16-byte aligned: two 16-byte loads per chunk instead of eight scalar ones (LSU queue)
        const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
        const float* src = rp.S + (size_t)pfRow * dS + 8 * j8;
        const float4 u0 = (ok && 8 * j8 < dS) ? ld_cg4(src) : z4, u1 = (ok && 8 * j8 + 4 < dS) ? ld_cg4(src + 4) : z4;
        pfX[c2][0] = u0.x; pfX[c2][1] = u0.y; pfX[c2][2] = u0.z; pfX[c2][3] = u0.w;
        pfX[c2][4] = u1.x; pfX[c2][5] = u1.y; pfX[c2][6] = u1.z; pfX[c2][7] = u1.w;
      } else {
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
          const int k = 8 * j8 + jj;
          pfX[c2][jj] = (ok && k < dS) ? ld_cg(rp.S + (size_t)pfRow * dS + k) : 0.f;
        }
      }
    }
  };
  prefetch(blockIdx.x);
  for (int tile = blockIdx.x; tile < nTiles; tile += gridDim.x) {
    const int b0 = tile * TBW, s = q * 32 + lane, b = b0 + s;
    const bool valid = b < a.B;
    // scratch of the tile: [stage of 16 samples][feature][16], tile after tile: every 16-sample stage of the weight-gradient
    // kernel is ONE contiguous block (all features x 64 bytes) that it streams front to back, and the 8 features a thread
    // writes for its sample are 8 adjacent 64-byte rows
    float* actT = a.actG + (size_t)tile * net.actPerSample * TBW;
    float* errT = a.errG + (size_t)tile * net.actPerSample * TBW;
    const int sOff = (s >> 4) * net.actPerSample * 16 + (s & 15);
    // ---- standardise (Episode.h:171-183): the tile's states become the A operand of the first layer ----
#pragma unroll
    for (int c2 = 0; c2 < 2; ++c2) {
      const int j8 = cg + 4 * c2;
      if (j8 >= D0.Kp / 8) continue;
      uint32_t hi[8], lo[8];
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) {
        const int k = 8 * j8 + jj;
        const float x = (valid && k < dS) ? (pfX[c2][jj] - msc[k]) * msc[D0.Kp + k] : 0.f;
        const float h = tf32_hi(x);
        hi[jj] = __float_as_uint(h); lo[jj] = __float_as_uint(x - h);
        if (k < dS) {
          actT[(D0.inOff + k) * 16 + sOff] = x;
          if (keep && valid) a.lastX[(size_t)b * dS + k] = x;
        }
      }
      tm_st8(tmem + laneBase + kWAh + 8 * j8, hi);
      tm_st8(tmem + laneBase + kWAl + 8 * j8, lo);
    }
    tm_wait_st();
    prefetch(tile + gridDim.x);
    tc_fence_before();
    __syncthreads();
    if (!imgReady) { mbar_wait(&bars[0], 0); imgReady = true; }
    // ---- layers ----
    for (int d = 0; d < wp.nD; ++d) {
      const WDense& D = wp.D[d];
      if (tid == 0) wide_issue(tmem, img + D.fImg, img + fHalf + D.fImg, D.Kp, D.Np, &bars[1]);
      if (!mbar_wait_bounded(&bars[1], mmaPar)) fault = true;
      mmaPar ^= 1u;
      tc_fence_after();
      const float* bias = vec + D.vB;
      const int dN = D.N, dNp8 = D.Np / 8, dY = D.yOff, dZ = D.zOff, func = net.func;      // registers (see k_wide_bwd)
      if (D.isTanh) {
        const bool res = D.res >= 0;
        const float* rw = vec + (res ? D.vRW : 0); const float* rb = vec + (res ? D.vRB : 0);
        for (int j8 = cg; j8 < dNp8; j8 += 4) {
          uint32_t v[8], xh[8], xl[8];
          tm_ld8(tmem + laneBase + kWAcc + 8 * j8, v);
          if (res) { tm_ld8(tmem + laneBase + kWAh + 8 * j8, xh); tm_ld8(tmem + laneBase + kWAl + 8 * j8, xl); }
          const int n0 = 8 * j8;
          float bv[8], rwv[8], rbv[8];
          *reinterpret_cast<float4*>(bv) = *reinterpret_cast<const float4*>(bias + n0);
          *reinterpret_cast<float4*>(bv + 4) = *reinterpret_cast<const float4*>(bias + n0 + 4);
          if (res) {
            *reinterpret_cast<float4*>(rwv) = *reinterpret_cast<const float4*>(rw + n0);
            *reinterpret_cast<float4*>(rwv + 4) = *reinterpret_cast<const float4*>(rw + n0 + 4);
            *reinterpret_cast<float4*>(rbv) = *reinterpret_cast<const float4*>(rb + n0);
            *reinterpret_cast<float4*>(rbv + 4) = *reinterpret_cast<const float4*>(rb + n0 + 4);
          }
          float* yp = actT + (dY + n0) * 16 + sOff;
          float* zp = actT + (dZ + n0) * 16 + sOff;
          tm_wait_ld();
          uint32_t hi[8], lo[8];
          // no branches inside: the eight columns are independent instruction streams (the padding columns have zero weights
          // and biases: tanh(0) = 0, and are not stored)
          float yv8[8];
          act_dispatch(func, [&](auto F) {
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) yv8[jj] = act_eval_t<decltype(F)::value>(__uint_as_float(v[jj]) + bv[jj]);      // BaseLayer::forward (Layer_Base.h:64-95)
          });
#pragma unroll
          for (int jj = 0; jj < 8; ++jj) {
            const float y = yv8[jj];
            float z = y;
            if (res) {                                                         // ParametricResidualLayer::forward (Layers.h:347-361)
              const float xin = __uint_as_float(xh[jj]) + __uint_as_float(xl[jj]);     // hi + lo is the f32 value, exactly
              z = y + (xin * rwv[jj] + rbv[jj]);
            }
            const bool in = n0 + jj < dN;
            z = in ? z : 0.f;
            if (in) { yp[jj * 16] = y; if (res) zp[jj * 16] = z; }
            const float h = tf32_hi(z);
            hi[jj] = __float_as_uint(h); lo[jj] = __float_as_uint(z - h);
          }
          tm_st8(tmem + laneBase + kWAh + 8 * j8, hi);
          tm_st8(tmem + laneBase + kWAl + 8 * j8, lo);
        }
        tm_wait_st();
      } else {      // linear output layer: the outputs leave for the loss kernel in the scratch rows its gradient will overwrite
        const int oOff = Lo.actOff;
        for (int j8 = cg; j8 < dNp8; j8 += 4) {
          uint32_t v[8];
          tm_ld8(tmem + laneBase + kWAcc + 8 * j8, v);
          tm_wait_ld();
#pragma unroll
          for (int jj = 0; jj < 8; ++jj) {
            const int n = 8 * j8 + jj;
            if (n < dN) errT[(oOff + n) * 16 + sOff] = __uint_as_float(v[jj]) + bias[n];
          }
        }
      }
      tc_fence_before();
      __syncthreads();
    }
  }
  if (!imgReady) mbar_wait(&bars[0], 0);       // no tile: still consume the copy before the CTA (and its shared memory) goes away
  if (fault && lane == 0 && a.comm.error) *a.comm.error = 2;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_free(tmem, 512u);
}

// ------------------------------------------------------------------------------------------
// k_wide_loss: ReF-ER / Retrace loss and output gradient (RACER::Train, Learners/RACER_train.cpp:31-60; the formulas and their
// operation order are those of loss_stages above), f64, V-RACER: outputs [V | mean(dA)], stdev = ParamLayer.  Its own kernel:
// 65536 x dA (sample, component) pairs of f64 transcendental work run at full occupancy on every SM instead of on the 16 warps
// of a tile CTA (a third of k_wide_fwd's time when it lived there).  CTA = 256 / dA samples; one thread per pair, one per sample
// for the sums in component order, the flags, the replay write-back and the record.  Reads the outputs from the scratch rows
// of the output layer and overwrites them with the gradient.
// ------------------------------------------------------------------------------------------
template <bool RACER>
__global__ void __launch_bounds__(256) k_wide_loss(StepArgs a, int step, int vP) {
  __shared__ double pairS[RACER ? 4 : 2][256];
  __shared__ double comp[5][8];         // per action component: root, stdev, dpos, 1 / stdev, log(1 / stdev) of the ParamLayer's stdev
  __shared__ StepCtrl c;
  const DevDescs* dd = a.descs;
  const NetDesc& net = dd->net; const Hyper& hp = dd->hp;
  const ReplayView& rp = a.rp;
  const int tid = threadIdx.x, dA = net.dA, SPC = 256 / dA, nPair = SPC * dA;
  const int per = net.actPerSample;
  const LayerDesc& Lo = net.L[net.nLayers - 2];
  const LayerDesc& Lp = net.L[net.nLayers - 1];
  const int Bt = (a.B + kWideM - 1) / kWideM * kWideM;
  const size_t j0 = (size_t)(step - a.stepBase) * a.B;
  const bool keep = step == a.lastStep || a.lastStep < 0;
  // outputs: V-RACER [V | mean(dA)], RACER [V | coef, p1(dA), p2(dA) | mean(dA)] (RACER_common.cpp:174-193,232-247)
  const int m0 = RACER ? 2 + 2 * dA : 1;
  if (tid == 0) load_ctrl(c, &a.ctrl[step & 1]);
  // terms of the ParamLayer's stdev outputs: the same for every sample, evaluated once per CTA (same expressions, same bits)
  if (tid >= 256 - 8 && tid - (256 - 8) < dA) {
    const int k = tid - (256 - 8);
    const double sraw = (double)ld_cg(a.wvec + vP + k);
    const double root = sqrt(1.0 + sraw * sraw);
    const double stdev = (sraw + root) / 2.0;                          // SoftPlus::_eval, Functions.h:552-555
    comp[0][k] = root; comp[1][k] = stdev;
    comp[2][k] = (1.0 + sraw / root) / 2.0;                            // SoftPlus::_evalDiff
    comp[3][k] = 1.0 / stdev; comp[4][k] = log(1.0 / stdev);
  }
  // one thread per (sample, action component); the thread of component 0 also owns the sample's scalars (old replay values,
  // write-back, record).  The per-sample terms (rho, V, flags) are evaluated by ALL threads of the sample, redundantly: a serial
  // per-sample stage on 1 / dA of the threads between two more block barriers left the f64 pipe 80 % idle.
  const int sp = tid / dA, i = tid - sp * dA;
  const int b = blockIdx.x * SPC + sp;
  const bool pairOn = tid < nPair && b < Bt;
  const bool valid = pairOn && b < a.B;
  float* errT = a.errG + ((size_t)(b >> 7) * per) * kWideM + ((b >> 4) & 7) * per * 16 + (b & 15);       // + scratch row * 16
  int row = 0, slotf = 0;
  float oldv[5] = {0.f, 0.f, 0.f, 0.f, 0.f};      // V, ADV, RHO, KL, DELTA of the transition (component-0 thread)
  float qretf = 0.f, O0f = 0.f, mf = 0.f, srawf = 0.f, avf = 0.f, mmf = 1.f, msf = 1.f, p1rf = 0.f, p2rf = 0.f, coefRawf = 0.f;
  if (valid) {      // every load of the thread in flight at once
    row = __ldg(a.sampRow + j0 + b);
    avf = ld_cg(rp.A + (size_t)row * dA + i); mmf = ld_cg(rp.MU + (size_t)row * 2 * dA + i); msf = ld_cg(rp.MU + (size_t)row * 2 * dA + dA + i);
    mf = ld_cg(errT + (Lo.actOff + m0 + i) * 16);
    O0f = ld_cg(errT + (Lo.actOff + 0) * 16);
    if (RACER) {
      coefRawf = ld_cg(errT + (Lo.actOff + 1) * 16);
      p1rf = ld_cg(errT + (Lo.actOff + 2 + i) * 16); p2rf = ld_cg(errT + (Lo.actOff + 2 + dA + i) * 16);
    }
    qretf = ld_cg(rp.Q + row);
    srawf = ld_cg(a.wvec + vP + i);
    if (i == 0) {
      slotf = __ldg(a.sampSlot + j0 + b);
      oldv[0] = ld_cg(rp.V + row); oldv[1] = ld_cg(rp.ADV + row); oldv[2] = ld_cg(rp.RHO + row);
      oldv[3] = ld_cg(rp.KL + row); oldv[4] = ld_cg(rp.DELTA + row);
    }
  }
  __syncthreads();        // comp, c
  double r_kgm = 0.0, r_kgs = 0.0, r_dlm = 0.0, r_dls = 0.0, r_dpos = 0.0;
  double r_p9 = -1.0, r_p10 = -1.0, r_F = 0.0, r_d1 = 0.0, r_d2 = 0.0, r_e1 = 0.0, r_e2 = 0.0;      // RACER: Gaussian advantage terms of the pair
  if (valid) {
    const double av = (double)avf, mm = (double)mmf, ms = (double)msf;
    const double m = (double)mf;
    const double stdev = comp[1][i], dpos = comp[2][i];
    const double inv = comp[3][i], invmu = 1.0 / ms;
    const bool bnd = hp.bounded[i] != 0;
    const double MAXM = 8.31776613503286;
    const double cm = bnd ? (m > MAXM ? MAXM : (m < -MAXM ? -MAXM : m)) : m;   // Continuous_policy.h:217-222
    const double fac0 = 9.1893853320467266954096885456237942e-01;
    double J = 1.0;
    if (bnd) { const double sq = tanh(av); J = fmax(1.0 - sq * sq, (double)FLT_MIN); }
    const double z1 = (av - cm) * inv, z2 = (av - mm) * invmu;
    const double lp_pi = -(z1 * z1) / 2.0 + (bnd ? log(inv / J) : comp[4][i]) - fac0;       // :91-97 / :240-249
    const double lp_mu = -(z2 * z2) / 2.0 + log(bnd ? invmu / J : invmu) - fac0;
    const double r1 = stdev / ms, r2 = (m - mm) / ms;
    const double cc = r1 * r1, dd2 = r2 * r2;                                         // OPPOSITE_KL, :138-142
    const double invVarMu = 1.0 / (ms * ms);
    const double u = z1;
    pairS[0][tid] = lp_pi - lp_mu;
    pairS[1][tid] = (cc - 1.0 + dd2 - log(cc)) / 2.0;
    r_kgm = -1.0 * ((m - mm) * invVarMu);                               // kg_mean
    r_kgs = (dpos * -1.0) * ((invVarMu - inv * inv) * stdev);           // kg_std
    r_dlm = bnd ? (av - m) * inv * inv : u * inv;                       // dLogPdMean
    r_dls = (u * u - 1.0) * inv;                                        // dLogPdStdv
    r_dpos = dpos;
    if (RACER) {   // Gaussian_advantage terms of this component (Gaus_advantage.h:73-126)
      const double p1r = (double)p1rf, p2r = (double)p2rf;
      const double rt1 = sqrt(1.0 + p1r * p1r), rt2 = sqrt(1.0 + p2r * p2r);
      const double p1 = (p1r + rt1) / 2.0, p2 = (p2r + rt2) / 2.0;                 // PosDefFunction::_eval
      const double S = stdev * stdev;                                             // policy->getVariance(i)
      const double dmA = av - cm;                                                 // policy->getMean(i): clamped for bounded dims
      const double sq1 = sqrt(p1 / (p1 + S)), sq2 = sqrt(p2 / (p2 + S));
      const double x1 = dmA / p1, x2 = dmA / p2;
      pairS[2][tid] = (dmA * dmA) / (av > cm ? p1 : p2);                          // diagInvMul term
      pairS[3][tid] = sq1 / 2.0 + sq2 / 2.0;                                      // coefMixRatio factor
      r_p9 = av > cm ? x1 * x1 : -1.0;                                            // ((a-m)/p1)^2 or "not on this side"
      r_p10 = av < cm ? x2 * x2 : -1.0;
      r_F = 2.0 / (sq1 + sq2);
      r_d1 = S / sqrt(p1 * ((p1 + S) * (p1 + S) * (p1 + S))) / 4.0;               // diff1
      r_d2 = S / sqrt(p2 * ((p2 + S) * (p2 + S) * (p2 + S))) / 4.0;               // diff2
      r_e1 = (1.0 + p1r / rt1) / 2.0;                                             // PosDefFunction::_evalDiff
      r_e2 = (1.0 + p2r / rt2) / 2.0;
    }
  }
  __syncthreads();        // pairS
  if (!pairOn) return;
  float g_mean_f = 0.f, g_std_f = 0.f, g0_f = 0.f, g1_f = 0.f, g2_f = 0.f, gc_f = 0.f;
  if (valid) {
    // ---- per-sample terms: sums in component order like the reference, flags, value terms ----
    const double O0 = (double)O0f;
    const double Vval = net2v(O0);                                                     // scaleNet2V (RACER_common.cpp:23-32)
    const double cmax = c.cmax, cinv = c.cinv, beta = c.beta;
    double logw = 0.0, dkl = 0.0;
    for (int k = 0; k < dA; ++k) { logw += pairS[0][sp * dA + k]; dkl += pairS[1][sp * dA + k]; }
    const double rho = exp(logw > 7.0 ? 7.0 : (logw < -7.0 ? -7.0 : logw));           // :648-653
    const float W32 = (float)rho, C32 = (float)cmax, I32 = (float)cinv;               // isFarPolicy takes Fval arguments (Episode.h:28-33)
    const bool offW = (W32 > C32) || (W32 < I32);
    const bool isFar = (C32 > 1.0f) && offW;
    double Aval = 0.0;                                                                // Zero_advantage.h:39-42
    double orig = 0.0, ratio = 1.0, coef = 0.0, dcoef = 0.0;
    if (RACER) {                                                                      // computeAdvantage, Gaus_advantage.h:73-78
      double shape = 0.0;
      for (int k = 0; k < dA; ++k) { shape += pairS[2][sp * dA + k]; ratio *= pairS[3][sp * dA + k]; }
      orig = exp(-shape / 2.0);
      const double coefRaw = (double)coefRawf, rtc = sqrt(1.0 + coefRaw * coefRaw);
      coef = (coefRaw + rtc) / 2.0;
      dcoef = (1.0 + coefRaw / rtc) / 2.0;
      Aval = coef * (orig - ratio);
    }
    const double A_RET = (double)qretf - Vval, deltaQ = A_RET - Aval;
    const double pgfac = A_RET * fmin(cmax, rho);
    // ---- policy / penalty gradient of the pair (penalizeReFER, FunctionUtilities.h:221-228) ----
    const double MAXM = 8.31776613503286;
    const double m = (double)mf;
    double pg_mean = pgfac * r_dlm;
    if (hp.bounded[i] && ((m >= MAXM && pg_mean > 0.0) || (m <= -MAXM && pg_mean < 0.0))) pg_mean = 0.0;
    double pg_std = (r_dpos * pgfac) * r_dls;
    if (isFar) { pg_mean = 0.0; pg_std = 0.0; }
    g_mean_f = (float)(beta * pg_mean + (1.0 - beta) * r_kgm);
    g_std_f = (float)(beta * pg_std + (1.0 - beta) * r_kgs);
    if (RACER) {        // ADV.grad(act, isFar ? 0 : beta * Aer, gradient), Gaus_advantage.h:88-114
      const double Aer = fmin(cmax, rho) * deltaQ;
      const double errA = isFar ? 0.0 : beta * Aer;
      const double oc = orig * coef, expect = -ratio;
      double g1 = r_p9 >= 0.0 ? oc * r_p9 / 2.0 : 0.0;
      double g2 = r_p10 >= 0.0 ? oc * r_p10 / 2.0 : 0.0;
      g1 += r_F * expect * coef * r_d1;
      g2 += r_F * expect * coef * r_d2;
      g1 *= errA * r_e1;
      g2 *= errA * r_e2;
      g1_f = (float)g1; g2_f = (float)g2;
      if (i == 0) gc_f = (float)((orig + expect) * (errA * dcoef));
    }
    if (keep) {
      a.lastG[(size_t)b * net.nOut + m0 + i] = g_mean_f; a.lastG[(size_t)b * net.nOut + m0 + dA + i] = g_std_f;
      a.lastO[(size_t)b * net.nOut + m0 + i] = mf; a.lastO[(size_t)b * net.nOut + m0 + dA + i] = srawf;
      if (RACER) {
        a.lastG[(size_t)b * net.nOut + 2 + i] = g1_f; a.lastG[(size_t)b * net.nOut + 2 + dA + i] = g2_f;
        a.lastO[(size_t)b * net.nOut + 2 + i] = p1rf; a.lastO[(size_t)b * net.nOut + 2 + dA + i] = p2rf;
        if (i == 0) { a.lastG[(size_t)b * net.nOut + 1] = gc_f; a.lastO[(size_t)b * net.nOut + 1] = coefRawf; }
      }
    }
    if (i == 0) {     // the sample's thread: value-head gradient (RACER_train.cpp:46), replay write-back (:59-60), record
      const double Ver = fmin(1.0, rho) * deltaQ;
      g0_f = (float)(isFar ? 0.0 : Ver * beta * vdiff(O0));
      if (keep) { a.lastG[(size_t)b * net.nOut + 0] = g0_f; a.lastO[(size_t)b * net.nOut + 0] = O0f; }
      const int slot = slotf & 0x7fffffff, hn = (slotf >> 31) & 1;
      if (hn) { const int k = atomicAdd(a.wcnt, 1); a.wlist[k] = b; }      // V(s_t+1): k_wide_next
      const float E = (float)deltaQ, Dk = (float)dkl;
      const float oldRho = oldv[2], oldKL = oldv[3], oldE = oldv[4];
      const bool wasOff = (oldRho > C32) || (oldRho < I32);
      const float Vf = (float)Vval, Qf = (float)(Aval + Vval);
      rp.DELTA[row] = E; rp.KL[row] = Dk; rp.RHO[row] = W32;
      rp.V[row] = Vf; rp.ADV[row] = Qf - Vf;
      // qNextOld / qNextNew of the record belong to k_wide_next
      *reinterpret_cast<int4*>(&a.rec[b].slot) = make_int4(slot, hn, (C32 > 1.0f) ? ((int)offW - (int)wasOff) : 0, 0);
      *reinterpret_cast<float4*>(&a.rec[b].dKL) = make_float4(Dk - oldKL, (float)offW - (float)wasOff, E * E - oldE * oldE, fabsf(E));
      *reinterpret_cast<float2*>(&a.rec[b].qOld) = make_float2(oldv[1] + oldv[0], Qf);
    }
  }
  // the gradient replaces the outputs in the scratch (zero for the padding samples of the last tile)
  errT[(Lo.actOff + m0 + i) * 16] = g_mean_f;
  errT[(Lp.actOff + i) * 16] = g_std_f;
  if (i == 0) errT[(Lo.actOff + 0) * 16] = g0_f;
  if (RACER) {
    errT[(Lo.actOff + 2 + i) * 16] = g1_f; errT[(Lo.actOff + 2 + dA + i) * 16] = g2_f;
    if (i == 0) errT[(Lo.actOff + 1) * 16] = gc_f;
  }
}

// ------------------------------------------------------------------------------------------
// k_wide_next: V(s_t+1) of the flagged samples (list filled by k_wide_fwd), 4 per pass, weights from the tile image in L2
// ------------------------------------------------------------------------------------------
template <bool SM>
__global__ void __launch_bounds__(kST) k_wide_next(StepArgs a, int step) {
  extern __shared__ __align__(128) unsigned char smraw[];
  const NetDesc* netp; const Hyper* hpp;
  load_descs(a, smraw, netp, hpp);
  const NetDesc& net = *netp;
  constexpr int TB = 4;
  const SmemPlan sp = smem_plan(net, TB, SM);
  float* img = reinterpret_cast<float*>(smraw + sp.img);
  const float* Wp = SM ? img : a.Wimg;
  float* act = reinterpret_cast<float*>(smraw + sp.act);
  float* red = reinterpret_cast<float*>(smraw + sp.red);
  uint64_t* bars = (SM && a.useTma) ? reinterpret_cast<uint64_t*>(smraw + sp.bars) : nullptr;
  const ReplayView& rp = a.rp;
  const int tid = threadIdx.x, dS = net.dS;
  const int count = min(__ldcg(a.wcnt), a.B);
  if ((int)blockIdx.x * TB >= count) return;
  if (SM) {        // the tile kernel's weight image (one bulk copy per layer), as in its own V(s_t+1) helper
    init_bars(a, net, smraw, sp.bars);
    load_weight_image(a, net, img, reinterpret_cast<uint64_t*>(smraw + sp.bars), step);
  }
  const size_t j0 = (size_t)(step - a.stepBase) * a.B;
  for (int first = blockIdx.x * TB; first < count; first += gridDim.x * TB) {
    for (int idx = tid; idx < dS * TB; idx += kST) {
      const int k = idx / TB, s = idx - k * TB;
      float x = 0.f;
      if (first + s < count) {
        const int b = __ldcg(a.wlist + first + s);
        const size_t row = (size_t)__ldcg(a.sampRow + j0 + b) + 1;
        x = (ld_cg(rp.S + row * dS + k) - ld_cg(rp.stateMean + k)) * ld_cg(rp.stateScale + k);
      }
      act[idx] = x;
    }
    __syncthreads();
    net_forward<TB, SM>(net, Wp, act, red, first == (int)blockIdx.x * TB ? bars : nullptr, 0);
    if (tid < TB && first + tid < count) {
      const int b = __ldcg(a.wlist + first + tid);
      const size_t row = (size_t)__ldcg(a.sampRow + j0 + b) + 1;
      const float vn = (float)net2v((double)net_out<SM>(net, Wp, act, TB, 0, tid));
      const float qOld = ld_cg(rp.ADV + row) + ld_cg(rp.V + row);
      rp.V[row] = vn; rp.ADV[row] = vn - vn;
      *reinterpret_cast<float2*>(&a.rec[b].qNextOld) = make_float2(qOld, vn);
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------
// k_wide_records: Episode::updateCumulative_atomic / updateValues_atomic (Episode.h:112-145) in sample order.  Samples are
// sorted, so the samples of an episode are one contiguous run: the thread of a run's first sample applies the whole run
// (the arithmetic of apply_sample_records above), all runs in parallel.
// ------------------------------------------------------------------------------------------
constexpr int kRecWin = 512;          // records staged per CTA: its own 256 and the 256 that follow (runs that start here and end there)
__global__ void __launch_bounds__(256) k_wide_records(StepArgs a) {
  __shared__ __align__(16) float win[kRecWin * 12];
  const ReplayView& rp = a.rp;
  const int ME = rp.maxEpisodes;
  const int c0 = blockIdx.x * 256, b = c0 + threadIdx.x;
  const int nWin = min(kRecWin, a.B - c0);
  {
    const float4* src = reinterpret_cast<const float4*>(a.rec + c0);
    float4* dst = reinterpret_cast<float4*>(win);
    for (int i = threadIdx.x; i < nWin * 3; i += 256) dst[i] = __ldcg(src + i);      // one coalesced pass
  }
  __syncthreads();
  int far = 0;
  if (b < a.B) {
    const float4* mine = reinterpret_cast<const float4*>(win + threadIdx.x * 12);
    const float4 h0 = mine[0];
    far = __float_as_int(h0.z);
    const int slot = __float_as_int(h0.x);
    const int prevSlot = threadIdx.x > 0 ? __float_as_int(mine[-3].x) : (b > 0 ? __ldcg(&a.rec[b - 1].slot) : -2);
    if (prevSlot != slot) {
      float avgKL = rp.epAgg[AGG_KL * ME + slot], frac = rp.epAgg[AGG_FAR * ME + slot];
      float avgE2 = rp.epAgg[AGG_E2 * ME + slot], maxE = rp.epAgg[AGG_MAXE * ME + slot];
      float sQ2 = rp.epAgg[AGG_Q2 * ME + slot], sQ = rp.epAgg[AGG_Q1 * ME + slot];
      float maxQ = rp.epAgg[AGG_MAXQ * ME + slot], minQ = rp.epAgg[AGG_MINQ * ME + slot];
      const float invN = 1.0f / (float)rp.epLen[slot];
      auto apply = [&](int hn, const float4 d, const float4 qv) {
        if (hn) {
          sQ2 += qv.w * qv.w - qv.z * qv.z; sQ += qv.w - qv.z;
          maxQ = fmaxf(maxQ, qv.w); minQ = fminf(minQ, qv.w);
        }
        avgKL += invN * d.x; frac += invN * d.y; avgE2 += invN * d.z; maxE = fmaxf(maxE, d.w);
        sQ2 += qv.y * qv.y - qv.x * qv.x; sQ += qv.y - qv.x;
        maxQ = fmaxf(maxQ, qv.y); minQ = fminf(minQ, qv.y);
      };
      bool open = true;
      int j = b;
      for (; j < c0 + nWin; ++j) {                    // the staged part of the run, in order
        const float4* rj = reinterpret_cast<const float4*>(win + (j - c0) * 12);
        const float4 hj = rj[0];
        if (__float_as_int(hj.x) != slot) { open = false; break; }
        apply(__float_as_int(hj.y), rj[1], rj[2]);
      }
      for (; open && j < a.B; j += 4) {               // a run longer than the window: four records travel at a time
        int4 hj[4]; float4 dj[4], qj[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int jj = min(j + u, a.B - 1);
          hj[u] = __ldcg(reinterpret_cast<const int4*>(&a.rec[jj]));
          dj[u] = __ldcg(reinterpret_cast<const float4*>(&a.rec[jj].dKL));
          qj[u] = __ldcg(reinterpret_cast<const float4*>(&a.rec[jj].qOld));
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (!open || j + u >= a.B || hj[u].x != slot) { open = false; continue; }
          apply(hj[u].y, dj[u], qj[u]);
        }
      }
      rp.epAgg[AGG_KL * ME + slot] = avgKL; rp.epAgg[AGG_FAR * ME + slot] = frac;
      rp.epAgg[AGG_E2 * ME + slot] = avgE2; rp.epAgg[AGG_MAXE * ME + slot] = maxE;
      rp.epAgg[AGG_Q2 * ME + slot] = sQ2; rp.epAgg[AGG_Q1 * ME + slot] = sQ;
      rp.epAgg[AGG_MAXQ * ME + slot] = maxQ; rp.epAgg[AGG_MINQ * ME + slot] = minQ;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) far += __shfl_xor_sync(0xffffffffu, far, o);
  if ((threadIdx.x & 31) == 0 && far != 0) atomicAdd(a.wcnt + 1, far);      // exact far-policy flag changes (integers: any order)
}

__global__ void __launch_bounds__(kST) k_wide_stats(StepArgs a, int step) {
  extern __shared__ __align__(128) unsigned char smraw[];
  const NetDesc* netp; const Hyper* hpp;
  load_descs(a, smraw, netp, hpp);
  __shared__ StepCtrl c;
  int fd = 0;
  if (threadIdx.x == 0) { load_ctrl(c, &a.ctrl[step & 1]); fd = __ldcg(a.wcnt + 1); a.wcnt[1] = 0; }
  __syncthreads();
  stats_and_refer(a, *hpp, c, a.ctrl[(step + 1) & 1], step, nullptr, -1, fd, reinterpret_cast<float*>(smraw + kWideDescBytes));
}

// ------------------------------------------------------------------------------------------
// k_wide_bwd: Network::backProp (Network.h:216-226) of a tile on the tensor cores.  For the dense layers from the output
// layer down to the second hidden layer:  E = D W^T  (A = the layer's deltas in TMEM, B = the TRANSPOSED image), then in the
// epilogue, for the layer below:  + the ParametricResidual path of the layer above (Layers.h:363-393), the residual layer's
// own error to the scratch (its parameter gradients are formed by k_wide_wgrad), deltas *= tanh' (Layer_Base.h:103-109).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kST, 1) k_wide_bwd(StepArgs a, int step, int sbBars, int sbImg, int bFloats) {
  extern __shared__ __align__(128) unsigned char smraw[];
  {
    uint64_t* bars0 = reinterpret_cast<uint64_t*>(smraw + sbBars);
    if (threadIdx.x == 0) {
      mbar_init(&bars0[0], 1); mbar_init(&bars0[1], 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      asm volatile("fence.proxy.async.global;\n\tfence.proxy.async.shared::cta;" ::: "memory");
      mbar_expect_tx(&bars0[0], (unsigned)bFloats * 4u);
      float* img0 = reinterpret_cast<float*>(smraw + sbImg);
      for (int off = 0; off < bFloats; off += 16384) {
        const unsigned n = (unsigned)min(16384, bFloats - off) * 4u;
        bulk_g2s(img0 + off, a.wimgB + off, n, &bars0[0]);
      }
    }
  }
  const NetDesc* netp; const Hyper* hpp;
  load_descs(a, smraw, netp, hpp);
  const WidePlan& wp = *load_wide_plan(a, smraw + kWideDescBytes);
  __shared__ uint32_t tmemSlot;
  __syncthreads();
  const NetDesc& net = *netp;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q = warp & 3, cg = warp >> 2;
  float* vec = reinterpret_cast<float*>(smraw + wp.sbVec);
  float* img = reinterpret_cast<float*>(smraw + wp.sbImg);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smraw + wp.sbBars);
  const int bHalf = wp.bFloats >> 1;
  constexpr int TBW = kWideM;
  if (warp == 0) tmem_alloc(&tmemSlot, 512u);
  for (int i = tid; i < wp.vFloats; i += kST) vec[i] = ld_cg(a.wvec + i);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmemSlot;
  const int nTiles = (a.B + TBW - 1) / TBW;
  unsigned mmaPar = 0;
  bool imgReady = false, fault = false;
  const uint32_t laneBase = (uint32_t)(q * 32) << 16;
  const LayerDesc& Lo = net.L[net.nLayers - 2];

  // the output gradient of a tile: chunk cg of the output layer's Np / 8 column chunks (Np <= 32: at most one chunk per thread),
  // fetched one tile ahead
  const WDense& DO = wp.D[wp.nD - 1];
  const bool gMine = cg < DO.Np / 8;
  float gv[8];
  const int sOff = ((q * 32 + lane) >> 4) * net.actPerSample * 16 + (lane & 15);      // scratch [stage of 16 samples][feature][16] (k_wide_fwd)
  auto fetch_g = [&](int tile) {
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) {
      const int n = 8 * cg + jj;
      gv[jj] = (gMine && tile < nTiles && n < DO.N) ? ld_cg(a.errG + (size_t)tile * net.actPerSample * TBW + (Lo.actOff + n) * 16 + sOff) : 0.f;
    }
  };
  fetch_g(blockIdx.x);
  for (int tile = blockIdx.x; tile < nTiles; tile += gridDim.x) {
    const int s = q * 32 + lane;
    const float* actT = a.actG + (size_t)tile * net.actPerSample * TBW;
    float* errT = a.errG + (size_t)tile * net.actPerSample * TBW;
    // ---- the output gradient of the tile -> A operand ----
    if (gMine) {
      uint32_t hi[8], lo[8];
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) { const float h = tf32_hi(gv[jj]); hi[jj] = __float_as_uint(h); lo[jj] = __float_as_uint(gv[jj] - h); }
      tm_st8(tmem + laneBase + kWAh + 8 * cg, hi);
      tm_st8(tmem + laneBase + kWAl + 8 * cg, lo);
      tm_wait_st();
    }
    tc_fence_before();
    __syncthreads();
    fetch_g(tile + gridDim.x);
    if (!imgReady) { mbar_wait(&bars[0], 0); imgReady = true; }
    for (int d = wp.nD - 1; d >= 1; --d) {
      const WDense& D = wp.D[d];           // product through this layer's weights
      const WDense& H = wp.D[d - 1];       // the hidden layer that receives the error
      // contraction over this layer's outputs (Np columns of the A operand), N = its Kp input rows
      if (tid == 0) wide_issue(tmem, img + D.bImg, img + bHalf + D.bImg, D.Np, D.Kp, &bars[1]);
      // the layer's outputs (tanh') do not depend on the product: fetch them while the tensor core works
      // plan fields in registers: read through the shared-memory copy inside the unrolled loops they are re-read after
      // every global store (possible aliasing)
      const int hN = H.N, hNp8 = H.Np / 8, hY = H.yOff, hZ = H.zOff, dKp = D.Kp, func = net.func;
      float yv[32];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int j8 = cg + 4 * i;
        const float* yp = actT + (hY + 8 * j8) * 16 + sOff;
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) yv[8 * i + jj] = (j8 < hNp8 && 8 * j8 + jj < hN) ? ld_cg(yp + jj * 16) : 0.f;
      }
      if (!mbar_wait_bounded(&bars[1], mmaPar)) fault = true;
      mmaPar ^= 1u;
      tc_fence_after();
      const bool res = H.res >= 0;
      const bool more = d - 1 >= 1;        // the layer below propagates further
      const bool haveCarry = d < wp.nD - 1 && D.res >= 0;      // E_z(d) * w_res(d), left in tensor memory by the previous epilogue
      const float* rw = vec + (res ? H.vRW : 0);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int j8 = cg + 4 * i;
        if (j8 >= hNp8) continue;
        uint32_t v[8], cy[8];
        const bool have = 8 * j8 < dKp;
        if (have) tm_ld8(tmem + laneBase + kWAcc + 8 * j8, v);
        if (haveCarry) tm_ld8(tmem + laneBase + kWCarry + 8 * j8, cy);
        if (have || haveCarry) tm_wait_ld();
        uint32_t hi[8], lo[8], cn[8];
        const int n0 = 8 * j8;
        float rwv[8];
        if (res) {
          *reinterpret_cast<float4*>(rwv) = *reinterpret_cast<const float4*>(rw + n0);
          *reinterpret_cast<float4*>(rwv + 4) = *reinterpret_cast<const float4*>(rw + n0 + 4);
        }
        float* ep = errT + (hZ + n0) * 16 + sOff;
        float* dp = errT + (hY + n0) * 16 + sOff;
        float dv8[8];
        act_dispatch(func, [&](auto F) {
#pragma unroll
          for (int jj = 0; jj < 8; ++jj) dv8[jj] = act_diff_t<decltype(F)::value>(yv[8 * i + jj]);      // deltas *= f' (Layer_Base.h:103-109)
        });
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
          const bool in = n0 + jj < hN;
          const float ez = (have ? __uint_as_float(v[jj]) : 0.f) + (haveCarry ? __uint_as_float(cy[jj]) : 0.f);   // E_in = W * delta (+ residual path)
          const float delta = in ? ez * dv8[jj] : 0.f;
          float cnew = 0.f;
          if (res) { cnew = in ? ez * rwv[jj] : 0.f; if (in) ep[jj * 16] = ez; }
          if (in) dp[jj * 16] = delta;
          cn[jj] = __float_as_uint(cnew);
          const float h = tf32_hi(delta);
          hi[jj] = __float_as_uint(h); lo[jj] = __float_as_uint(delta - h);
        }
        if (more) {
          tm_st8(tmem + laneBase + kWAh + 8 * j8, hi); tm_st8(tmem + laneBase + kWAl + 8 * j8, lo);
          if (res) tm_st8(tmem + laneBase + kWCarry + 8 * j8, cn);
        }
      }
      if (more) tm_wait_st();
      tc_fence_before();
      __syncthreads();
    }
  }
  if (!imgReady) mbar_wait(&bars[0], 0);
  if (fault && lane == 0 && a.comm.error) *a.comm.error = 2;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_free(tmem, 512u);
}

// ------------------------------------------------------------------------------------------
// k_wide_wgrad: the weight gradients as split-K contractions over the samples (Layers.h:160-187).  CTA g owns the 16-sample
// stages [g * nStages / G, (g + 1) * nStages / G).  Per stage and dense layer two operands are staged from the feature-major
// scratch: thread (r = tid / 4, cc = tid % 4) loads the float4 of feature row r, samples 4 cc .. 4 cc + 3 (four lanes = one
// 64-byte run), splits it and stores hi / lo into the K-major images  float4 [4 k-chunks][LD rows]  (LD = 130 or rows + 2:
// conflict-free 16-byte stores).  Hidden layers: M side = the layer's deltas (D[n][k], N = Kp input rows), output layer: M side
// = its input rows, N side = the output-gradient rows [dense | ParamLayer].  Vector gradients (biases, residual w / b,
// ParamLayer) are the row sums of what the thread just loaded.
// ------------------------------------------------------------------------------------------
template <int ND>
__global__ void __launch_bounds__(kST, 1) k_wide_wgrad(StepArgs a, int step) {
  extern __shared__ __align__(128) unsigned char smraw[];
  const NetDesc* netp; const Hyper* hpp;
  load_descs(a, smraw, netp, hpp);
  const WidePlan& wp = *load_wide_plan(a, smraw + kWideDescBytes);
  __shared__ uint32_t tmemSlot;
  __syncthreads();
  const NetDesc& net = *netp;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int r = tid >> 2, cc = tid & 3;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smraw + wp.sgBars);      // [image stage]: the MMAs that read it have completed
  if (tid == 0) {
    mbar_init(&bars[0], 1); mbar_init(&bars[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc(&tmemSlot, (uint32_t)wp.gCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmemSlot;
  const int Bt = (a.B + kWideM - 1) / kWideM * kWideM;       // the scratch is defined (zero deltas) up to the tile boundary
  const int nSt = Bt / kWideKS;
  const int G = gridDim.x, g = blockIdx.x;
  const int st0 = (int)((long long)g * nSt / G), st1 = (int)((long long)(g + 1) * nSt / G);
  const LayerDesc& Lo = net.L[net.nLayers - 2];
  const LayerDesc& Lp = net.L[net.nLayers - 1];
  // this thread's operand rows: scratch row pointers (nullptr = zero row) and shared-memory slots, fixed for the whole launch
  const float* pA[ND]; const float* pB[ND]; const float* pE[ND];
  int oA[ND], oB[ND], ldB[ND];
#pragma unroll
  for (int d = 0; d < ND; ++d) {
    const WDense& D = wp.D[d];
    const bool out = d == ND - 1;
    pA[d] = pB[d] = pE[d] = nullptr;
    if (!out) {
      if (r < D.N) pA[d] = a.errG + (D.yOff + r) * 16 + 4 * cc;                        // deltas of the layer (stage-relative)
      if (r < D.K) pB[d] = a.actG + (D.inOff + r) * 16 + 4 * cc;                       // its input
      if (D.res >= 0 && r < D.N) pE[d] = a.errG + (D.zOff + r) * 16 + 4 * cc;          // error on the residual layer
    } else {
      if (r < D.K) pA[d] = a.actG + (D.inOff + r) * 16 + 4 * cc;
      if (r < net.nOut) pB[d] = a.errG + (r < net.nOutDense ? Lo.actOff + r : Lp.actOff + (r - net.nOutDense)) * 16 + 4 * cc;
    }
    ldB[d] = wp.sgRowsB[d] + 2;
    oA[d] = wp.sgOpA[d] + (cc * kWideLD + r) * 16;
    oB[d] = r < wp.sgRowsB[d] ? wp.sgOpB[d] + (cc * ldB[d] + r) * 16 : -1;
  }
  // operand descriptors of stage 0 (the issuing thread adds the stage offset and the k-chunk offset: the issue loop is the
  // serial part of a stage)
  __shared__ uint64_t gd[ND][5]; __shared__ uint32_t gi[ND][2];      // read by the issuing thread only: no registers of the other 511
  if (tid == 0) {
    for (int d = 0; d < ND; ++d) {
      const unsigned char* s0 = smraw + wp.sgStage;
      const int lb = wp.sgRowsB[d] + 2;
      gd[d][0] = umma_desc(s0 + wp.sgOpA[d], kWideLD * 16, 128); gd[d][1] = umma_desc(s0 + wp.sgOpA[d] + 4 * kWideLD * 16, kWideLD * 16, 128);
      gd[d][2] = umma_desc(s0 + wp.sgOpB[d], lb * 16, 128); gd[d][3] = umma_desc(s0 + wp.sgOpB[d] + 4 * lb * 16, lb * 16, 128);
      gd[d][4] = (uint64_t)lb; gi[d][0] = wide_idesc(wp.D[d].gN); gi[d][1] = tmem + (uint32_t)wp.D[d].gCol;
    }
  }
  float sA[ND], sE[ND], sEX[ND];
#pragma unroll
  for (int d = 0; d < ND; ++d) { sA[d] = 0.f; sE[d] = 0.f; sEX[d] = 0.f; }
  bool fault = false;
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  // Pipeline of a stage (16 samples): raw rows travel global -> shared memory by cp.async (LDGSTS, no registers) into ONE raw
  // stage [operand][thread] float4 — every thread reads back exactly the 16-byte slots it copied, so cp.async.wait_group is the
  // only synchronisation the raw stage needs —, are picked up into registers one stage ahead, split into the hi / lo operand
  // images of one of TWO image stages, and multiplied; the MMAs of stage i (tcgen05.commit -> mbarrier of its image stage) run
  // while stage i + 1 is split and stage i + 2 travels.
  float4* raw = reinterpret_cast<float4*>(smraw + wp.sgRaw) + tid;
  const int nImg = wp.sgStages;            // image stages: 2, or 1 when four dense layers leave no room for the second
  int qA[ND], qB[ND], qE[ND], nOps = 0;
#pragma unroll
  for (int d = 0; d < ND; ++d) { qA[d] = nOps++; qB[d] = nOps++; qE[d] = (d < ND - 1 && wp.D[d].res >= 0) ? nOps++ : -1; }
  auto issue = [&](int it) {
    if (it < st1) {
      // stage `it` (samples [16 (it % 8), +16) of tile it / 8) is one contiguous block of the scratch: [feature][16 samples]
      const size_t col = (size_t)it * net.actPerSample * kWideKS;
#pragma unroll
      for (int d = 0; d < ND; ++d) {
        if (pA[d]) cp_async16_cg(raw + qA[d] * kST, pA[d] + col);
        if (pB[d]) cp_async16_cg(raw + qB[d] * kST, pB[d] + col);
        if (pE[d]) cp_async16_cg(raw + qE[d] * kST, pE[d] + col);
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  float4 vA[ND], vB[ND], vE[ND];
  auto pickup = [&]() {
    asm volatile("cp.async.wait_group 0;" ::: "memory");
#pragma unroll
    for (int d = 0; d < ND; ++d) {
      vA[d] = pA[d] ? raw[qA[d] * kST] : zero4;
      vB[d] = pB[d] ? raw[qB[d] * kST] : zero4;
      vE[d] = pE[d] ? raw[qE[d] * kST] : zero4;
    }
  };
  issue(st0);
  pickup();
  issue(st0 + 1);

  for (int it = st0; it < st1; ++it) {
    const int slot = (it - st0) % nImg, use = (it - st0) / nImg;
    unsigned char* stg = smraw + wp.sgStage + (size_t)slot * wp.sgStageBytes;
    if (use > 0) { if (!mbar_wait_bounded(&bars[slot], (unsigned)((use - 1) & 1))) fault = true; }     // MMAs of the image stage's previous use
#pragma unroll
    for (int d = 0; d < ND; ++d) {
      const bool out = d == ND - 1;
      {
        float4 h, l; split4(vA[d], h, l);
        float4* Ah = reinterpret_cast<float4*>(stg + oA[d]);
        Ah[0] = h; Ah[4 * kWideLD] = l;
      }
      if (oB[d] >= 0) {
        float4 h, l; split4(vB[d], h, l);
        float4* Bh = reinterpret_cast<float4*>(stg + oB[d]);
        Bh[0] = h; Bh[4 * ldB[d]] = l;
      }
      if (!out) {
        sA[d] += vA[d].x; sA[d] += vA[d].y; sA[d] += vA[d].z; sA[d] += vA[d].w;                   // db += delta
        if (pE[d]) {                                                                              // ParametricResidualLayer::backward
          sE[d] += vE[d].x; sE[d] += vE[d].y; sE[d] += vE[d].z; sE[d] += vE[d].w;
          sEX[d] = fmaf(vE[d].x, vB[d].x, sEX[d]); sEX[d] = fmaf(vE[d].y, vB[d].y, sEX[d]);
          sEX[d] = fmaf(vE[d].z, vB[d].z, sEX[d]); sEX[d] = fmaf(vE[d].w, vB[d].w, sEX[d]);
        }
      } else { sA[d] += vB[d].x; sA[d] += vB[d].y; sA[d] += vB[d].z; sA[d] += vB[d].w; }         // output bias / ParamLayer
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      const uint64_t sOff = (uint64_t)((slot * wp.sgStageBytes) >> 4);
      const uint32_t first = it > st0 ? 1u : 0u;
#pragma unroll
      for (int d = 0; d < ND; ++d) {
#pragma unroll
        for (int kk = 0; kk < kWideKS / 8; ++kk) {
          const uint64_t dah = gd[d][0] + sOff + (uint64_t)(2 * kk * kWideLD), dal = gd[d][1] + sOff + (uint64_t)(2 * kk * kWideLD);
          const uint64_t dbh = gd[d][2] + sOff + (uint64_t)(2 * kk) * gd[d][4], dbl = gd[d][3] + sOff + (uint64_t)(2 * kk) * gd[d][4];
          umma_tf32(gi[d][1], dal, dbh, gi[d][0], kk > 0 ? 1u : first);
          umma_tf32(gi[d][1], dah, dbl, gi[d][0], 1u);
          umma_tf32(gi[d][1], dah, dbh, gi[d][0], 1u);
        }
      }
      tc_commit(&bars[slot]);
    }
    pickup();                  // stage it + 1 (landed while this one was split) -> registers
    issue(it + 2);             // refills the raw slots this thread has just read
  }
  asm volatile("cp.async.wait_all;" ::: "memory");
  // ---- all MMAs done: accumulators and vector sums -> this CTA's partial record ----
  float* rec = a.wpart + (size_t)g * wp.recFloats;
  const int nMine = st1 - st0;
  if (nMine > 0) {
    const int last = nMine - 1;
    if (!mbar_wait_bounded(&bars[last % nImg], (unsigned)((last / nImg) & 1))) fault = true;
  }
  tc_fence_after();
  {
    const int q = warp & 3, cg = warp >> 2;
    const uint32_t laneBase = (uint32_t)(q * 32) << 16;
    const int m = q * 32 + lane;
    for (int d = 0; d < ND; ++d) {
      const WDense& D = wp.D[d];
      for (int j8 = cg; j8 < D.gN / 8; j8 += 4) {
        uint32_t v[8];
        if (nMine > 0) { tm_ld8(tmem + laneBase + (uint32_t)D.gCol + 8 * j8, v); tm_wait_ld(); }
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) rec[D.gPart + (size_t)(8 * j8 + jj) * 128 + m] = nMine > 0 ? __uint_as_float(v[jj]) : 0.f;
      }
    }
  }
#pragma unroll
  for (int d = 0; d < ND; ++d) {
    const WDense& D = wp.D[d];
    float x = sA[d], y = sE[d], z = sEX[d];
    x += __shfl_xor_sync(0xffffffffu, x, 1); x += __shfl_xor_sync(0xffffffffu, x, 2);
    y += __shfl_xor_sync(0xffffffffu, y, 1); y += __shfl_xor_sync(0xffffffffu, y, 2);
    z += __shfl_xor_sync(0xffffffffu, z, 1); z += __shfl_xor_sync(0xffffffffu, z, 2);
    if (cc == 0) {
      const bool out = d == ND - 1;
      const int rows = out ? wp.NpG : D.Np;
      if (r < rows) {
        rec[D.vSum + r] = x;
        if (!out && D.res >= 0) { rec[D.vSum + D.Np + r] = y; rec[D.vSum + 2 * D.Np + r] = z; }
      }
    }
  }
  if (fault && lane == 0 && a.comm.error) *a.comm.error = 2;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_free(tmem, (uint32_t)wp.gCols);
}

// ------------------------------------------------------------------------------------------
// k_wide_adam: gradient = sum of the partial records in CTA order; AdamOptimizer::apply_update (adam_step); every image of
// the weights is rewritten here: blob, tile-kernel image, split forward / transposed operand images, vector block.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_wide_adam(StepArgs a, int step, int nParams, int recFloats, int fHalf, int bHalf) {
  __shared__ float part[8][33];
  const int px = threadIdx.x & 31, sy = threadIdx.x >> 5;       // 32 consecutive parameters (coalesced rows of a record) x 8 slices of the records
  const int p = blockIdx.x * 32 + px;
  if (blockIdx.x == 0 && threadIdx.x == 0) a.wcnt[0] = 0;       // the V(s_t+1) list of the next step starts empty
  const int rpos = p < nParams ? __ldg(a.widx + p) : -1;        // -1: padding of the parameter blob
  const int G = a.wGridG;
  float acc = 0.f;
  if (rpos >= 0) {
    const int g0 = sy * G / 8, g1 = (sy + 1) * G / 8;
    const float* src = a.wpart + rpos;
    int g = g0;
    for (; g + 8 <= g1; g += 8) {                    // eight loads in flight, added in CTA order
      float x[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) x[u] = ld_cg(src + (size_t)(g + u) * recFloats);
#pragma unroll
      for (int u = 0; u < 8; ++u) acc += x[u];
    }
    for (; g < g1; ++g) acc += ld_cg(src + (size_t)g * recFloats);
  }
  part[sy][px] = acc;
  __syncthreads();
  if (sy != 0 || rpos < 0) return;
  acc = 0.f;
#pragma unroll
  for (int u = 0; u < 8; ++u) acc += part[u][px];    // slices in order: the sum is a fixed function of the grid size
  // ---- gradient sum over learner ranks (replaces the MPI_Iallreduce of AdamOptimizer::prepare_update, Optimizer.cpp:114-118):
  //      the tile kernel's exchange — every rank stores its element straight into every peer's slot over NVLink (poison = not
  //      arrived yet), polls its own slots in LOCAL memory and adds the N values in rank order, so all ranks apply the identical
  //      update.  92 KB per rank and step against a step of >= 60 us. ----
  if (a.comm.world > 1) {
    const CommView& cm = a.comm;
    const int N = cm.world, me = cm.rank, rot = step & 3;
    const size_t slotMe = ((size_t)rot * N + me) * cm.nParamsPad;
    const unsigned pk = __float_as_uint(acc);
    for (int q = 0; q < N; ++q) if (q != me) st_volatile_u32(cm.grad(q) + slotMe + p, pk);
    unsigned* mine = cm.grad(me) + (size_t)rot * N * cm.nParamsPad;
    unsigned got[kMaxWorld];
    if (!wait_values_poison(mine + p, cm.nParamsPad, N, me, cm, got)) return;      // a lost peer (error flag set): leave the parameter untouched
    float v = 0.f;
#pragma unroll
    for (int q = 0; q < kMaxWorld; ++q) if (q < N) v += q == me ? acc : __uint_as_float(got[q]);
    acc = v;
  }
  const DevDescs* dd = a.descs;
  AdamCoef ac;
  ac.eta = __ldcg(&a.ctrl[step & 1].adam_eta);
  ac.B1 = 0.9f; ac.B2 = 0.999f;
  ac.lambda = (float)dd->hp.nnLambda;
  ac.fac = (float)(1.0 / (double)dd->hp.batchGlobal);
  a.G[p] = acc;
  const float Wn = adam_step(ac, acc, ld_cg(a.W + p), ld_cg(a.M1 + p), ld_cg(a.M2 + p), a.W + p, a.M1 + p, a.M2 + p);
  const int pi = __ldg(a.widx + nParams + p), pf = __ldg(a.widx + 2 * nParams + p), pb = __ldg(a.widx + 3 * nParams + p),
            pv = __ldg(a.widx + 4 * nParams + p);
  if (pi >= 0) a.Wimg[pi] = Wn;
  const float h = tf32_hi(Wn), l = Wn - h;
  if (pf >= 0) { a.wimgF[pf] = h; a.wimgF[fHalf + pf] = l; }
  if (pb >= 0) { a.wimgB[pb] = h; a.wimgB[bHalf + pb] = l; }
  if (pv >= 0) a.wvec[pv] = Wn;
}

// ------------------------------------------------------------------------------------------
// host side: plan, index maps, images, launchers
// ------------------------------------------------------------------------------------------
static inline int wide_np(int n) { return n <= 16 ? 16 : (n <= 32 ? 32 : (n <= 64 ? 64 : 128)); }
static inline int wide_up(int n, int m) { return (n + m - 1) / m * m; }

// idx: [5][nParams] — position of the parameter's gradient in a partial record; its positions in the tile-kernel image, the
// forward image (hi half), the transposed image (hi half) and the vector block; -1 = none.
void wide_plan_build(const NetDesc& net, const Hyper& hp, WidePlan& wp, std::vector<int>& idx) {
  memset(&wp, 0, sizeof(wp));
  const int nP = net.nParams;
  idx.assign((size_t)5 * nP, -1);
  if (net.recurrent || net.discrete || (hp.algo != 0 && hp.algo != 1) || net.dA > 8 || net.dS > 64) return;     // feed-forward V-RACER / RACER, continuous actions, dA <= 8
  int nD = 0;
  for (int l = 1; l < net.nLayers; ++l) {
    const LayerDesc& L = net.L[l];
    if (L.kind != kDenseTanh && L.kind != kDenseLinear) continue;
    if (nD == kWideMaxD || L.size > 128 || L.nIn > 128) return;
    WDense& d = wp.D[nD];
    d.layer = l; d.res = (l + 1 < net.nLayers && net.L[l + 1].kind == kResidual) ? l + 1 : -1;
    d.K = L.nIn; d.Kp = wide_up(L.nIn, 16); d.N = L.size; d.Np = wide_np(L.size); d.isTanh = L.kind == kDenseTanh ? 1 : 0;
    d.inOff = net.L[L.in].actOff; d.yOff = L.actOff; d.zOff = d.res >= 0 ? net.L[d.res].actOff : L.actOff;
    if (d.res >= 0 && d.N > d.K) return;
    ++nD;
  }
  if (nD < 2 || wp.D[nD - 1].isTanh || net.L[net.nLayers - 1].kind != kParam || wp.D[nD - 1].layer != net.nLayers - 2) return;
  for (int d = 0; d + 1 < nD; ++d) if (!wp.D[d].isTanh || wp.D[d + 1].Kp > wp.D[d].Np) return;
  wp.nD = nD;
  wp.NpG = wide_up(net.nOut, 16);
  if (wp.NpG > 128) return;
  // images: [hi of every layer | lo of every layer]
  int f = 0, b = 0, v = 0;
  for (int d = 0; d < nD; ++d) {
    WDense& D = wp.D[d];
    D.fImg = f; f += D.Kp * D.Np;
    D.bImg = -1;
    if (d >= 1) { D.bImg = b; b += D.Np * D.Kp; }
    D.vB = v; v += D.Np;
    D.vRW = D.vRB = -1;
    if (D.res >= 0) { D.vRW = v; v += D.Np; D.vRB = v; v += D.Np; }
  }
  const int dA4 = wide_up(net.dA, 4);
  const int vP = v; v += dA4;                       // ParamLayer values
  wp.vP = vP;
  wp.fFloats = 2 * f; wp.bFloats = 2 * b; wp.vFloats = v;
  wp.vMsc = v;                                      // forward kernel only: state mean / scale [2][Kp0] behind the vector block
  // weight-gradient kernel: accumulators, partial record, stage layout
  int col = 0, rec = 0, sg = 0;
  for (int d = 0; d < nD; ++d) {
    WDense& D = wp.D[d];
    const bool out = d == nD - 1;
    D.gN = out ? wp.NpG : D.Kp;
    D.gCol = col; col += D.gN;
    D.gPart = rec; rec += D.gN * 128;
    wp.sgRowsA[d] = 128; wp.sgRowsB[d] = D.gN;
    wp.sgOpA[d] = sg; sg += 2 * 4 * kWideLD * 16;
    wp.sgOpB[d] = sg; sg += 2 * 4 * (D.gN + 2) * 16;
  }
  for (int d = 0; d < nD; ++d) {
    WDense& D = wp.D[d];
    const bool out = d == nD - 1;
    D.vSum = rec; rec += out ? wp.NpG : (D.res >= 0 ? 3 * D.Np : D.Np);
  }
  if (col > 512) return;
  wp.gCols = col <= 32 ? 32 : (col <= 64 ? 64 : (col <= 128 ? 128 : (col <= 256 ? 256 : 512)));
  wp.recFloats = wide_up(rec, 32);
  wp.sgStageBytes = sg;
  // shared-memory layouts
  const int kMax = 227 * 1024 - 2048;               // static shared memory of the kernels (ctrl, slots, statistics scratch) stays below 2 KB
  int o = kWideDescBytes + kWidePlanBytes;
  auto take = [&](int bytes) { const int at = o; o += (bytes + 127) / 128 * 128; return at; };
  wp.sfBars = take(64);
  wp.sfVec = take(4 * (wp.vFloats + 2 * wp.D[0].Kp));
  wp.sfImg = take(4 * wp.fFloats);
  wp.sfTotal = o;
  o = kWideDescBytes + kWidePlanBytes;
  wp.sbBars = take(64);
  wp.sbVec = take(4 * wp.vFloats);
  wp.sbImg = take(4 * wp.bFloats);
  wp.sbTotal = o;
  o = kWideDescBytes + kWidePlanBytes;
  wp.sgBars = take(64);
  {
    int nOps = 0;
    for (int d = 0; d < nD; ++d) nOps += 2 + ((d < nD - 1 && wp.D[d].res >= 0) ? 1 : 0);
    const int rawBytes = nOps * kST * 16;
    wp.sgStages = (o + 2 * sg + rawBytes + 256 <= kMax) ? 2 : 1;
    wp.sgStage = take(wp.sgStages * sg);
    wp.sgRaw = take(rawBytes);
  }
  wp.sgTotal = o;
  if (wp.sfTotal > kMax || wp.sbTotal > kMax || wp.sgTotal > kMax) return;
  // index maps
  int* iRec = idx.data(); int* iImg = iRec + nP; int* iF = iImg + nP; int* iB = iF + nP; int* iV = iB + nP;
  for (int d = 0; d < nD; ++d) {
    const WDense& D = wp.D[d];
    const LayerDesc& L = net.L[D.layer];
    const bool out = d == nD - 1;
    for (int k = 0; k < D.K; ++k)
      for (int n = 0; n < D.N; ++n) {
        const int p = L.wOff + k * L.ld + n;
        iRec[p] = out ? D.gPart + n * 128 + k : D.gPart + k * 128 + n;
        iImg[p] = L.imgW + k * L.ldp + n;
        iF[p] = D.fImg + ((k >> 2) * D.Np + n) * 4 + (k & 3);
        if (D.bImg >= 0) iB[p] = D.bImg + ((n >> 2) * D.Kp + k) * 4 + (n & 3);
      }
    for (int n = 0; n < D.N; ++n) {
      const int p = L.bOff + n;
      iRec[p] = D.vSum + n; iImg[p] = L.imgB + n; iV[p] = D.vB + n;
    }
    if (D.res >= 0) {
      const LayerDesc& R = net.L[D.res];
      for (int n = 0; n < D.N; ++n) {
        iRec[R.bOff + n] = D.vSum + D.Np + n; iImg[R.bOff + n] = R.imgB + n; iV[R.bOff + n] = D.vRB + n;
        iRec[R.wOff + n] = D.vSum + 2 * D.Np + n; iImg[R.wOff + n] = R.imgW + n; iV[R.wOff + n] = D.vRW + n;
      }
    }
  }
  {
    const LayerDesc& P = net.L[net.nLayers - 1];
    const WDense& D = wp.D[nD - 1];
    for (int i = 0; i < P.size; ++i) {
      const int p = P.bOff + i;
      iRec[p] = D.vSum + net.nOutDense + i; iImg[p] = P.imgB + i; iV[p] = vP + i;
    }
  }
  wp.ok = 1;
}

// operand images and vector block of a parameter blob (host; the Adam kernel keeps them current afterwards)
void wide_fill_images(const NetDesc& net, const WidePlan& wp, const std::vector<int>& idx, const float* blob,
                      std::vector<float>& imgF, std::vector<float>& imgB, std::vector<float>& vecs) {
  const int nP = net.nParams;
  imgF.assign(wp.fFloats, 0.f); imgB.assign(wp.bFloats > 0 ? wp.bFloats : 1, 0.f); vecs.assign(wp.vFloats, 0.f);
  const int* iF = idx.data() + 2 * (size_t)nP; const int* iB = iF + nP; const int* iV = iB + nP;
  const int fHalf = wp.fFloats / 2, bHalf = wp.bFloats / 2;
  for (int p = 0; p < nP; ++p) {
    const float w = blob[p];
    uint32_t u; memcpy(&u, &w, 4);
    // cvt.rna.tf32.f32: round to nearest, ties away from zero, on the 13 dropped mantissa bits
    uint32_t hu = u;
    if ((u & 0x7f800000u) != 0x7f800000u) hu = (u + 0x1000u) & 0xffffe000u;
    float h; memcpy(&h, &hu, 4);
    const float l = w - h;
    if (iF[p] >= 0) { imgF[iF[p]] = h; imgF[fHalf + iF[p]] = l; }
    if (iB[p] >= 0) { imgB[iB[p]] = h; imgB[bHalf + iB[p]] = l; }
    if (iV[p] >= 0) vecs[iV[p]] = w;
  }
}

int wide_prepare(const WidePlan& wp, const NetDesc& net) {
  SMB200_CUDA_CHECK(cudaFuncSetAttribute(k_wide_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, wp.sfTotal));
  SMB200_CUDA_CHECK(cudaFuncSetAttribute(k_wide_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize, wp.sbTotal));
  SMB200_CUDA_CHECK(cudaFuncSetAttribute(k_wide_wgrad<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, wp.sgTotal));
  SMB200_CUDA_CHECK(cudaFuncSetAttribute(k_wide_wgrad<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, wp.sgTotal));
  SMB200_CUDA_CHECK(cudaFuncSetAttribute(k_wide_wgrad<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, wp.sgTotal));
  SMB200_CUDA_CHECK(cudaFuncSetAttribute(k_wide_next<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_plan(net, 4, false).total));
  if (step_image_in_smem(net))
    SMB200_CUDA_CHECK(cudaFuncSetAttribute(k_wide_next<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_plan(net, 4, true).total));
  SMB200_CUDA_CHECK(cudaFuncSetAttribute(k_wide_stats, cudaFuncAttributeMaxDynamicSharedMemorySize, kWideDescBytes + 4 * kStatChunk));
  return 0;
}

int wide_grid_g(const WidePlan& wp, int B, int numSMs) {
  const int nSt = wide_up(B, kWideM) / kWideKS;
  return nSt < numSMs ? nSt : numSMs;
}

// Stream plan of a step.  main: fwd -> bwd -> wgrad -> adam;  aux (forked after fwd): next -> records -> stats.  The tile
// kernels run on the SMALLEST grid that keeps their makespan (512 tiles on 148 SMs are four rounds on 128 CTAs just as well), so
// the auxiliary kernels find free SMs beside them.  adam waits for `next` (which reads the old weight image), the next
// step's fwd waits for the statistics (ReF-ER coefficients, record buffer).
int launch_steps_wide(const StepArgs& a, const NetDesc& net, const WidePlan& wp, int numSMs, int step0, int nSteps, int skipStatsLast,
                      cudaStream_t st, cudaStream_t aux, cudaEvent_t evF, cudaEvent_t evN, cudaEvent_t evS) {
  const int nTiles = (a.B + kWideM - 1) / kWideM;
  const int rounds = (nTiles + numSMs - 1) / numSMs;
  const int gridT = (nTiles + rounds - 1) / rounds;
  const bool sm = step_image_in_smem(net);
  const int spc = 256 / net.dA;
  const bool racer = net.nOutDense == 2 + 3 * net.dA;      // RACER's Gaussian advantage head (the plan admits V-RACER and RACER)
  for (int s = 0; s < nSteps; ++s) {
    const int step = step0 + s;
    const bool skipStats = skipStatsLast && s == nSteps - 1;
    k_wide_fwd<<<gridT, kST, wp.sfTotal, st>>>(a, step, wp.sfBars, wp.sfImg, wp.fFloats);
    if (racer) k_wide_loss<true><<<(nTiles * kWideM + spc - 1) / spc, 256, 0, st>>>(a, step, wp.vP);
    else k_wide_loss<false><<<(nTiles * kWideM + spc - 1) / spc, 256, 0, st>>>(a, step, wp.vP);
    SMB200_CUDA_CHECK(cudaEventRecord(evF, st));
    SMB200_CUDA_CHECK(cudaStreamWaitEvent(aux, evF, 0));
    if (sm) k_wide_next<true><<<16, kST, smem_plan(net, 4, true).total, aux>>>(a, step);
    else k_wide_next<false><<<16, kST, smem_plan(net, 4, false).total, aux>>>(a, step);
    SMB200_CUDA_CHECK(cudaEventRecord(evN, aux));
    if (!skipStats) {
      k_wide_records<<<(a.B + 255) / 256, 256, 0, aux>>>(a);
      k_wide_stats<<<1, kST, kWideDescBytes + 4 * kStatChunk, aux>>>(a, step);
    }
    SMB200_CUDA_CHECK(cudaEventRecord(evS, aux));
    k_wide_bwd<<<gridT, kST, wp.sbTotal, st>>>(a, step, wp.sbBars, wp.sbImg, wp.bFloats);
    if (wp.nD == 2) k_wide_wgrad<2><<<a.wGridG, kST, wp.sgTotal, st>>>(a, step);
    else if (wp.nD == 3) k_wide_wgrad<3><<<a.wGridG, kST, wp.sgTotal, st>>>(a, step);
    else k_wide_wgrad<4><<<a.wGridG, kST, wp.sgTotal, st>>>(a, step);
    SMB200_CUDA_CHECK(cudaStreamWaitEvent(st, evN, 0));
    k_wide_adam<<<(net.nParams + 31) / 32, 256, 0, st>>>(a, step, net.nParams, wp.recFloats, wp.fFloats / 2, wp.bFloats / 2);
    SMB200_CUDA_CHECK(cudaStreamWaitEvent(st, evS, 0));
  }
  SMB200_CUDA_CHECK(cudaGetLastError());
  return 0;
}
