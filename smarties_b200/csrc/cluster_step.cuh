// cluster_step.cuh — the learner step of feed-forward nets on thread-block clusters (included by step_kernels.cu, inside
// namespace smb200, after the phase functions it reuses: loss_stages, p3_stats, grid_barrier, ...).
//
// The persistent kernel of step_kernels.cu gives every CTA a tile of 4 sampled transitions and the WHOLE network: each
// layer moves the full weight matrix through one SM's shared-memory pipe for four samples, the weight gradient is contracted
// afterwards over the whole mini-batch by other CTAs (P2), and only 64 of the 148 SMs work during P1 at B = 256.  Here
//   * a CLUSTER of kCL = 4 CTAs owns a pass of kTS = 8 sampled transitions.  CTA r keeps 1/4 of the output columns of every
//     hidden layer — its slice of the weight image: forward layout [k][n], a transposed copy [n][k] for the input gradient —
//     and computes that slice for all 8 samples (4 x 2 register blocks, K split over 16 lanes, shuffle reduction: no
//     block barrier inside a layer); the layer output is broadcast to the other CTAs through distributed shared memory
//     (st.shared::cluster) and a cluster barrier ends the layer;
//   * the small linear output layer is replicated: every CTA evaluates it, the f64 ReF-ER / Retrace loss (loss_stages, the
//     same code as the tile kernel) and the output-layer input gradient for ITS OWN 2 samples;
//   * backward: the error on the top hidden layer is broadcast, each CTA forms the deltas of its slice, multiplies by its
//     transposed slice and sends the partial input gradient to the CTA that owns those rows;
//   * the weight gradient of the pass is formed IN the cluster from the activations and deltas that are still in shared
//     memory (each CTA: its column slice) and accumulated in a shared-memory accumulator; after the last pass of a step the
//     cluster writes ONE partial sum per parameter to global memory;
//   * P2 (all CTAs, one parameter per thread, weights and Adam moments resident in registers for the whole launch): adds
//     the per-cluster partial sums in cluster order, exchanges with the other learner ranks (same poison-slot protocol as
//     p2_tile), applies the reference's Adam variant and writes the parameter to the weight images.
// 32 clusters of 4 = 128 CTAs cover B = 256 in one pass; the 33rd cluster holds the asynchronous statistics CTA (P3) and
// three helper CTAs that evaluate V(s_t+1) of truncated episodes (single-CTA forward over all four slices).
// Two grid barriers per step, four cluster barriers per pass.
#pragma once

// ---- cluster primitives ----
__device__ __forceinline__ unsigned cluster_ctarank() { unsigned r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ unsigned cluster_id_x() { unsigned r; asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t map_to_rank(uint32_t addr, unsigned rank) {
  uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank)); return r;
}
__device__ __forceinline__ void st_cluster_f32(uint32_t addr, float v) { asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory"); }
__device__ __forceinline__ void st_cluster_v2(uint32_t addr, float x, float y) {
  asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" ::"r"(addr), "f"(x), "f"(y) : "memory");
}

// Bounded waits of the cluster kernel: a grid barrier or an image copy that does not complete within ~2 s ends the launch
// with an error code in the learner's device error flag (reported by smb200_comm_error) instead of hanging the GPU.
constexpr long long kWaitCycles = 4000000000LL;
__device__ __forceinline__ void cl_fail(const StepArgs& a, int code) {
  if (a.comm.error) { *a.comm.error = code; __threadfence_system(); }
  __trap();
}
__device__ __forceinline__ void cl_grid_barrier(const StepArgs& a, unsigned& target, unsigned nblocks) {
  __syncthreads();
  if (threadIdx.x == 0) {
    target += nblocks;
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(a.barrier) : "memory");
    const long long t0 = clock64();
    while (ld_acquire(a.barrier) < target) { if (clock64() - t0 > kWaitCycles) cl_fail(a, 16); }
  }
  __syncthreads();
}
__device__ __forceinline__ void cl_mbar_wait(const StepArgs& a, uint64_t* bar, unsigned parity) {
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) { if (clock64() - t0 > kWaitCycles) cl_fail(a, 17); }
}

// index load that stays where it is written (the compiler would otherwise sink it to its first use, after the grid barrier)
__device__ __forceinline__ int ld_index_now(const int* p) {
  int v; asm volatile("ld.global.nc.s32 %0, [%1];" : "=r"(v) : "l"(p)); return v;
}

struct ClusterCtx {
  unsigned rank;            // CTA rank in the cluster = slice index
  uint32_t peer[kCL];       // cluster-window address of every rank's activation area
  bool share;               // false: single-CTA evaluation (helper CTAs): local stores only
};

// store one value at float offset `off` of the activation area of ranks [r0, r0 + n)
__device__ __forceinline__ void bcast_store(const ClusterCtx& cx, float* act, int off, float v, int r0, int n) {
  if (!cx.share) { act[off] = v; return; }
  for (int r = r0; r < r0 + n; ++r) st_cluster_f32(cx.peer[r] + 4u * (unsigned)off, v);
}

// ------------------------------------------------------------------------------------------
// Hidden dense layer, forward, column slice `slice` for all kTS samples:
//   y[n][s] = f(b[n] + sum_k x[k][s] W[k][n]),  out = y (+ x[n] * resW[n] + resB[n] with a ParametricResidual), out -> every CTA.
// Thread mapping: half-warp = one 4-sample x 2-column block, its 16 lanes split K (k = lane, lane + 16, ...); the 8
// accumulators are reduced over the 16 lanes by a halving butterfly (8 shuffles), after which lane pair (2a, 2a + 1) holds
// output a = 2 * sample + column.  Row strides ldf = NS + 2 (ldf / 2 odd) and kXS = 12 make the 8-byte weight loads and the
// 16-byte activation loads conflict-free.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void cl_fwd_hidden(const CDense& L, const float* rk, const float* cm, float* act, const float* xin,
                                              int slice, const ClusterCtx& cx) {
  const int tid = threadIdx.x, hw = tid >> 4, l16 = tid & 15;
  const int nBlk = L.NS;                                  // (NS / 2 column pairs) x 2 sample groups
  const float* Wf = rk + L.iWf;
  float* ys = act + L.sYs;
  for (int blk = hw; blk < nBlk; blk += kST / 16) {
    const int sg = blk & 1, cp = blk >> 1;
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
    const float* xp = xin + l16 * kXS + sg * 4;
    const float* wp = Wf + l16 * L.ldf + cp * 2;
#pragma unroll 4
    for (int k = l16; k < L.Kp; k += 16, xp += 16 * kXS, wp += 16 * L.ldf) {
      const float4 x = *reinterpret_cast<const float4*>(xp);
      const float2 w = *reinterpret_cast<const float2*>(wp);
      acc[0] = fmaf(x.x, w.x, acc[0]); acc[1] = fmaf(x.x, w.y, acc[1]);
      acc[2] = fmaf(x.y, w.x, acc[2]); acc[3] = fmaf(x.y, w.y, acc[3]);
      acc[4] = fmaf(x.z, w.x, acc[4]); acc[5] = fmaf(x.z, w.y, acc[5]);
      acc[6] = fmaf(x.w, w.x, acc[6]); acc[7] = fmaf(x.w, w.y, acc[7]);
    }
    const bool h3 = (l16 & 8) != 0, h2 = (l16 & 4) != 0, h1 = (l16 & 2) != 0;
    float r4[4], r2[2];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float send = h3 ? acc[i] : acc[i + 4], keep = h3 ? acc[i + 4] : acc[i];
      r4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const float send = h2 ? r4[i] : r4[i + 2], keep = h2 ? r4[i + 2] : r4[i];
      r2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
    float v;
    { const float send = h1 ? r2[0] : r2[1], keep = h1 ? r2[1] : r2[0]; v = keep + __shfl_xor_sync(0xffffffffu, send, 2); }
    v = v + __shfl_xor_sync(0xffffffffu, v, 1);
    const int a = (l16 >> 1) & 7;
    const int s = sg * 4 + (a >> 1), nl = cp * 2 + (a & 1);
    const int n = slice * L.NS + nl;
    float y = v + rk[L.iB + nl];
    y = tanh_ref(y);
    float out = y;
    if (L.iRW >= 0) {
      if ((l16 & 1) == 0) ys[nl * kXS + s] = y;
      if (n < L.K) out = y + (xin[n * kXS + s] * cm[L.iRW + n] + cm[L.iRB + n]);   // ParametricResidualLayer::forward (Layers.h:347-361)
    }
    const int off = L.sXout + n * kXS + s;
    if (!cx.share) { if ((l16 & 1) == 0) act[off] = out; }
    else bcast_store(cx, act, off, out, (l16 & 1) * 2, 2);        // the two lanes of the pair serve two ranks each
  }
}

// ------------------------------------------------------------------------------------------
// Linear output layer for the samples [s0, s0 + 2) of the pass: out2[(actOff + n) * 2 + j] = b[n] + sum_k x[k][s0 + j] W[k][n]
// (2 x 2 register blocks, K split over 16 lanes).  nCols: only the first nCols outputs are needed (helper: the value head).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void cl_fwd_out(const ClusterPlan& cp, const float* cm, const float* xtop, int s0, float* out2, int actOff, int nCols) {
  const int tid = threadIdx.x, hw = tid >> 4, l16 = tid & 15;
  const int nBlk = ((min(nCols, cp.oN) + 3) >> 2) << 1;     // column pairs, rounded up to whole warps (two half-warps shuffle together)
  const float* Wo = cm + cp.iOW;
  for (int blk = hw; blk < nBlk; blk += kST / 16) {
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    const float* xp = xtop + l16 * kXS + s0;
    const float* wp = Wo + l16 * cp.ldo + blk * 2;
#pragma unroll 4
    for (int k = l16; k < cp.oKp; k += 16, xp += 16 * kXS, wp += 16 * cp.ldo) {
      const float2 x = *reinterpret_cast<const float2*>(xp);
      const float2 w = *reinterpret_cast<const float2*>(wp);
      acc[0] = fmaf(x.x, w.x, acc[0]); acc[1] = fmaf(x.x, w.y, acc[1]);
      acc[2] = fmaf(x.y, w.x, acc[2]); acc[3] = fmaf(x.y, w.y, acc[3]);
    }
    const bool h3 = (l16 & 8) != 0, h2 = (l16 & 4) != 0;
    float r2[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const float send = h3 ? acc[i] : acc[i + 2], keep = h3 ? acc[i + 2] : acc[i];
      r2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
    float v;
    { const float send = h2 ? r2[0] : r2[1], keep = h2 ? r2[1] : r2[0]; v = keep + __shfl_xor_sync(0xffffffffu, send, 4); }
    v = v + __shfl_xor_sync(0xffffffffu, v, 2);
    v = v + __shfl_xor_sync(0xffffffffu, v, 1);
    const int a = ((l16 >> 3) & 1) * 2 + ((l16 >> 2) & 1);
    const int j = a >> 1, n = blk * 2 + (a & 1);
    if ((l16 & 3) == 0 && n < cp.oN) out2[(actOff + n) * kOwn + j] = v + cm[cp.iOB + n];
  }
}

// ------------------------------------------------------------------------------------------
// Input-gradient partial of a hidden layer: Epart[k][s] = sum_{n in slice} W[k][n] delta[n][s] for ALL k, sent to the CTA
// that owns row k of the layer below (slot `rank` of its partial buffer).  4 k x 4 s register blocks, the slice's columns
// split over 8 lanes, halving butterfly (14 shuffles) -> every lane ends with 2 outputs (one 8-byte remote store).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void cl_bwd_dx(const CDense& L, const CDense& Lin, const float* rk, float* act, const ClusterCtx& cx) {
  const int tid = threadIdx.x, grp = tid >> 3, l8 = tid & 7;
  const float* Wt = rk + L.iWt;
  const float* ds = act + L.sDs;
  const int nBlk = (L.Kp >> 2) * 2;                  // (Kp / 4 row blocks) x 2 sample groups
  for (int blk = grp; blk < nBlk; blk += kST / 8) {
    const int sg = blk & 1, kb = blk >> 1;
    float acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = 0.f;
    for (int n = l8; n < L.NS; n += 8) {
      const float4 w = *reinterpret_cast<const float4*>(Wt + n * L.ldt + kb * 4);
      const float4 d = *reinterpret_cast<const float4*>(ds + n * kXS + sg * 4);
      acc[0] = fmaf(w.x, d.x, acc[0]); acc[1] = fmaf(w.x, d.y, acc[1]); acc[2] = fmaf(w.x, d.z, acc[2]); acc[3] = fmaf(w.x, d.w, acc[3]);
      acc[4] = fmaf(w.y, d.x, acc[4]); acc[5] = fmaf(w.y, d.y, acc[5]); acc[6] = fmaf(w.y, d.z, acc[6]); acc[7] = fmaf(w.y, d.w, acc[7]);
      acc[8] = fmaf(w.z, d.x, acc[8]); acc[9] = fmaf(w.z, d.y, acc[9]); acc[10] = fmaf(w.z, d.z, acc[10]); acc[11] = fmaf(w.z, d.w, acc[11]);
      acc[12] = fmaf(w.w, d.x, acc[12]); acc[13] = fmaf(w.w, d.y, acc[13]); acc[14] = fmaf(w.w, d.z, acc[14]); acc[15] = fmaf(w.w, d.w, acc[15]);
    }
    const bool h2 = (l8 & 4) != 0, h1 = (l8 & 2) != 0, h0 = (l8 & 1) != 0;
    float r8[8], r4[4], r2[2];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float send = h2 ? acc[i] : acc[i + 8], keep = h2 ? acc[i + 8] : acc[i];
      r8[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float send = h1 ? r8[i] : r8[i + 4], keep = h1 ? r8[i + 4] : r8[i];
      r4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const float send = h0 ? r4[i] : r4[i + 2], keep = h0 ? r4[i + 2] : r4[i];
      r2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 1);
    }
    const int a = (h2 ? 8 : 0) + (h1 ? 4 : 0) + (h0 ? 2 : 0);     // outputs a, a + 1: row a >> 2, samples (a & 3), (a & 3) + 1
    const int k = kb * 4 + (a >> 2), s = sg * 4 + (a & 3);
    if (k < L.K) {
      const int owner = k / Lin.NS, kl = k - owner * Lin.NS;
      const int off = Lin.sEpart + ((int)cx.rank * Lin.NS + kl) * kXS + s;
      if (cx.share) st_cluster_v2(cx.peer[owner] + 4u * (unsigned)off, r2[0], r2[1]);
      else { act[off] = r2[0]; act[off + 1] = r2[1]; }
    }
  }
}

// ------------------------------------------------------------------------------------------
// Weight gradient of one pass, formed in the cluster: the CTA's column slice of every hidden layer, its rows of the output
// layer, (rank 0) the output bias and the ParamLayer.  One work item per thread and pass: 4 x 4 register blocks over the 8
// samples with interleaved rows / columns (conflict-free: lanes read consecutive rows), vector items for biases and
// residual parameters.  Results go to the shared-memory accumulator `g` (`first`: store, else add).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void cl_dot8(const float* a, const float* b, float& acc) {   // sum over the 8 samples, sample order
  const float4 a0 = *reinterpret_cast<const float4*>(a), a1 = *reinterpret_cast<const float4*>(a + 4);
  const float4 b0 = *reinterpret_cast<const float4*>(b), b1 = *reinterpret_cast<const float4*>(b + 4);
  acc = fmaf(a0.x, b0.x, acc); acc = fmaf(a0.y, b0.y, acc); acc = fmaf(a0.z, b0.z, acc); acc = fmaf(a0.w, b0.w, acc);
  acc = fmaf(a1.x, b1.x, acc); acc = fmaf(a1.y, b1.y, acc); acc = fmaf(a1.z, b1.z, acc); acc = fmaf(a1.w, b1.w, acc);
}
__device__ __forceinline__ float cl_sum8(const float* a) {
  const float4 a0 = *reinterpret_cast<const float4*>(a), a1 = *reinterpret_cast<const float4*>(a + 4);
  float v = 0.f; v += a0.x; v += a0.y; v += a0.z; v += a0.w; v += a1.x; v += a1.y; v += a1.z; v += a1.w;
  return v;
}

// One work item per thread, fixed for the whole launch and decoded on the host (cluster_plan_build): 12 ints
//   [0] kind  1: 4 x 4 block  acc[i][j] += <a_i, b_j> over the 8 samples   (dense / output-layer weights)
//             2: column triple acc[0] += sum(a_0)  (bias), with a residual: acc[1] += <b_0, b_1> (dRW), acc[2] += sum(b_0) (dRB)
//             3: single        acc[0] += sum(a_0)  (output bias, ParamLayer)
//   [1] aoff [2] astr   rows a_i = act + aoff + i * astr        [3] boff [4] bstr   rows b_j = act + boff + j * bstr
//   [5] obase [6] ostrA [7] ostrB   parameter index of acc[i][j] = obase + i * ostrA + j * ostrB  (kind 2: obase, ostrA, ostrB = the three parameters)
//   [8] na [9] nb       valid rows / columns (kind 2: nb = 1 with a residual)
// Lanes run over consecutive columns (b rows at stride kXS: conflict-free, a rows broadcast), so the stores of a warp fall into
// whole 32-byte sectors of the partial-gradient row.
constexpr int kItemInts = 12;
// `early`: the items that only read the deltas of the top hidden layer and of the output layer (record [10] = 1) — they run
// while the partial input gradients of the top layer travel; !early: all the others.
__device__ __forceinline__ void cl_wgrad(const int* tab, const float* act, float (&acc)[16], bool early) {
  const int4 q0 = __ldg(reinterpret_cast<const int4*>(tab)), q1 = __ldg(reinterpret_cast<const int4*>(tab) + 1);
  const int4 q2 = __ldg(reinterpret_cast<const int4*>(tab) + 2);
  const int kind = (q2.z != 0) == early ? q0.x : 0;
  if (kind == 1) {
    const float* ar = act + q0.y; const float* br = act + q0.w;
    const int astr = q0.z, bstr = q1.x;
    float4 a[8], b[8];
#pragma unroll
    for (int i = 0; i < 4; ++i) { a[2 * i] = *reinterpret_cast<const float4*>(ar + i * astr); a[2 * i + 1] = *reinterpret_cast<const float4*>(ar + i * astr + 4); }
#pragma unroll
    for (int j = 0; j < 4; ++j) { b[2 * j] = *reinterpret_cast<const float4*>(br + j * bstr); b[2 * j + 1] = *reinterpret_cast<const float4*>(br + j * bstr + 4); }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float v = acc[i * 4 + j];
        v = fmaf(a[2 * i].x, b[2 * j].x, v); v = fmaf(a[2 * i].y, b[2 * j].y, v); v = fmaf(a[2 * i].z, b[2 * j].z, v); v = fmaf(a[2 * i].w, b[2 * j].w, v);
        v = fmaf(a[2 * i + 1].x, b[2 * j + 1].x, v); v = fmaf(a[2 * i + 1].y, b[2 * j + 1].y, v);
        v = fmaf(a[2 * i + 1].z, b[2 * j + 1].z, v); v = fmaf(a[2 * i + 1].w, b[2 * j + 1].w, v);
        acc[i * 4 + j] = v;
      }
  } else if (kind == 2) {
    acc[0] += cl_sum8(act + q0.y);
    if (q2.y) {
      const float* e = act + q0.w; const float* x = e + q1.x;
      float v = 0.f; cl_dot8(e, x, v);
      acc[1] += v; acc[2] += cl_sum8(e);
    }
  } else if (kind == 3) acc[0] += cl_sum8(act + q0.y);
}
__device__ __forceinline__ void cl_wgrad_store(const int* tab, float* part, const float (&acc)[16]) {
  const int4 q0 = __ldg(reinterpret_cast<const int4*>(tab)), q1 = __ldg(reinterpret_cast<const int4*>(tab) + 1);
  const int4 q2 = __ldg(reinterpret_cast<const int4*>(tab) + 2);
  const int kind = q0.x;
  if (kind == 1) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (i < q2.x && j < q2.y) part[q1.y + i * q1.z + j * q1.w] = acc[i * 4 + j];
  } else if (kind == 2) {
    part[q1.y] = acc[0];
    if (q2.y) { part[q1.z] = acc[1]; part[q1.w] = acc[2]; }
  } else if (kind == 3) part[q1.y] = acc[0];
}

// ------------------------------------------------------------------------------------------
// weight image -> shared memory (cp.async.bulk, one mbarrier per destination block)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void cl_load_image(const StepArgs& a, const ClusterPlan& cp, unsigned char* smraw, int rank, uint64_t* bars) {
  if (threadIdx.x == 0) {
    asm volatile("fence.proxy.async.global;\n\tfence.proxy.async.shared::cta;" ::: "memory");
    // forward parts (common block + the slice's forward layout) on bars[0], the transposed copies on bars[1]: the forward
    // pass starts as soon as the first 2/3 of the image has landed
    const unsigned cb = 4u * (unsigned)cp.commonFloats, fb = 4u * (unsigned)cp.rankFwdFloats, tb = 4u * (unsigned)(cp.rankFloats - cp.rankFwdFloats);
    const float* src = a.cimg + cp.commonFloats + (size_t)rank * cp.rankFloats;
    mbar_expect_tx(&bars[0], cb + fb);
    bulk_g2s(smraw + cp.bCommon, a.cimg, cb, &bars[0]);
    bulk_g2s(smraw + cp.bRank, src, fb, &bars[0]);
    if (tb) { mbar_expect_tx(&bars[1], tb); bulk_g2s(smraw + cp.bRank + fb, src + cp.rankFwdFloats, tb, &bars[1]); }
  }
}

// ------------------------------------------------------------------------------------------
// One pass of the cluster: kTS sampled transitions [b0, b0 + 8) of the step.  `staged`: the pass inputs were prefetched into
// the staging area during the previous step's P2 (see k_steps_cluster).
// Staging area (floats): S [8][dS] | old [8][kOwn] | pair [3][kOwn * dA] | info int [4][kOwn] | rows int [8] | mean [dS] | scale [dS] | chunks [8][kOwn][4]
// ------------------------------------------------------------------------------------------
struct CStage { float* S; float* old; float* pair; int* info; int* rows; float* mean; float* scale; float* chunks; };
__device__ __forceinline__ CStage cl_stage_view(const NetDesc& net, unsigned char* smraw, const ClusterPlan& cp) {
  CStage g;
  float* f = reinterpret_cast<float*>(smraw + cp.bStage);
  g.S = f; f += kTS * net.dS;
  g.old = f; f += 8 * kOwn;
  g.pair = f; f += 3 * kOwn * net.dA;
  g.info = reinterpret_cast<int*>(f); f += 4 * kOwn;
  g.rows = reinterpret_cast<int*>(f); f += kTS;
  g.mean = f; f += net.dS;
  g.scale = f; f += net.dS;
  f = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(f) + 15) & ~(uintptr_t)15);
  g.chunks = f;
  return g;
}

__device__ void cl_pass(const StepArgs& a, const DevDescs& dd, const ClusterPlan& cp, StepCtrl& c, int step, int b0,
                        unsigned char* smraw, const ClusterCtx& cx, bool fetchCtrl, const unsigned* readyFlag, unsigned readyTarget,
                        bool staged, bool helped, float (&wacc)[16], uint64_t* imgBars, unsigned imgParity) {
  const NetDesc& net = dd.net; const Hyper& hp = dd.hp;
  const int tid = threadIdx.x, rank = (int)cx.rank;
  const float* cm = reinterpret_cast<const float*>(smraw + cp.bCommon);
  const float* rk = reinterpret_cast<const float*>(smraw + cp.bRank);
  float* act = reinterpret_cast<float*>(smraw + cp.bAct);
  float* act2 = reinterpret_cast<float*>(smraw + cp.bAct2);     // [actPerSample][kOwn]: network outputs of the own samples
  float* err2 = reinterpret_cast<float*>(smraw + cp.bErr2);
  const CStage stg = cl_stage_view(net, smraw, cp);
  int* info = staged ? stg.info : reinterpret_cast<int*>(smraw + cp.bInfo);      // row, slot, hasNext, valid of the own samples
  float* old = staged ? stg.old : reinterpret_cast<float*>(smraw + cp.bOld);     // [8][kOwn]
  int* rows = stg.rows;                                                           // ring rows of the 8 samples (-1: beyond the batch)
  double* pair = reinterpret_cast<double*>(smraw + cp.bPair);
  double* samp = reinterpret_cast<double*>(smraw + cp.bSamp);
  const ReplayView& rp = a.rp;
  const int dS = net.dS, dA = net.dA, nPair = kOwn * dA;
  const int s0 = rank * kOwn;                       // own samples of the pass
  const size_t jb = (size_t)(step - a.stepBase) * a.B + b0;
  const LayerDesc& Lo = net.L[cp.oLayer];
  const LayerDesc& Lp = net.L[cp.pLayer];
  const int* witem = a.citems + ((size_t)rank * kST + tid) * kItemInts;

  if (!staged) {
    if (tid < kTS) rows[tid] = b0 + tid < a.B ? a.sampRow[jb + tid] : -1;
    if (tid >= 32 && tid < 32 + kOwn) {
      const int j = tid - 32, b = b0 + s0 + j;
      int row = 0, slot = 0, hn = 0, valid = 0;
      if (b < a.B) {
        row = a.sampRow[jb + s0 + j];
        const int sf = a.sampSlot[jb + s0 + j];
        slot = sf & 0x7fffffff; hn = (sf >> 31) & 1; valid = 1;
        old[0 * kOwn + j] = ld_cg(rp.V + row); old[1 * kOwn + j] = ld_cg(rp.ADV + row);
        old[2 * kOwn + j] = ld_cg(rp.RHO + row); old[3 * kOwn + j] = ld_cg(rp.KL + row);
        old[4 * kOwn + j] = ld_cg(rp.DELTA + row); old[7 * kOwn + j] = ld_cg(rp.Q + row);
        if (hn) { old[5 * kOwn + j] = ld_cg(rp.V + row + 1); old[6 * kOwn + j] = ld_cg(rp.ADV + row + 1); }
      }
      info[j] = row; info[kOwn + j] = slot; info[2 * kOwn + j] = hn; info[3 * kOwn + j] = valid;
    }
    __syncthreads();
  }
  // ---- gather + standardise all 8 states of the pass: x0[k][s] = (S - mean) * scale (Episode.h:171-183) ----
  float* x0 = act + cp.sX0;
  const bool keep = step == a.lastStep || a.lastStep < 0;
  for (int idx = tid; idx < dS * kTS; idx += kST) {
    const int s = idx / dS, k = idx - s * dS;
    float x = 0.f;
    if (rows[s] >= 0)
      x = staged ? (stg.S[s * dS + k] - stg.mean[k]) * stg.scale[k]
                 : (ld_cg(rp.S + (size_t)rows[s] * dS + k) - ld_cg(rp.stateMean + k)) * ld_cg(rp.stateScale + k);
    x0[k * kXS + s] = x;
    if (keep && rows[s] >= 0 && s >= s0 && s < s0 + kOwn) a.lastX[(size_t)(b0 + s) * dS + k] = x;
  }
  for (int idx = tid; idx < net.actPerSample * kOwn; idx += kST) err2[idx] = 0.f;   // clearErrors
  double pa = 0, pmm = 0, pms = 1;
  const int p0s = tid < nPair ? tid / dA : 0, p0i = tid - p0s * dA;
  if (tid < nPair && info[3 * kOwn + p0s]) {
    if (staged) { pa = (double)stg.pair[tid]; pmm = (double)stg.pair[nPair + tid]; pms = (double)stg.pair[2 * nPair + tid]; }
    else {
      const size_t row = info[p0s];
      pa = (double)ld_cg(rp.A + row * dA + p0i);
      pmm = (double)ld_cg(rp.MU + row * 2 * dA + p0i);
      pms = (double)ld_cg(rp.MU + row * 2 * dA + dA + p0i);
    }
  }
  if (imgBars) cl_mbar_wait(a, &imgBars[0], imgParity);      // first pass of the step: the forward part of the weight image
  __syncthreads();
  DBG_T(a, step, 2);

  // ---- forward: hidden layers slice by slice, outputs exchanged through distributed shared memory ----
  for (int li = 0; li < cp.nDense; ++li) {
    const CDense& L = cp.L[li];
    cl_fwd_hidden(L, rk, cm, act, li == 0 ? x0 : act + cp.L[li - 1].sXout, rank, cx);
    if (li < 3) DBG_T(a, step, 25 + li);
    cluster_sync_all();
    if (li < 3) DBG_T(a, step, 9 + li);
  }
  const CDense& Lt = cp.L[cp.nDense - 1];
  cl_fwd_out(cp, cm, act + Lt.sXout, s0, act2, Lo.actOff, cp.oN);
  __syncthreads();
  DBG_T(a, step, 12);

  // ---- ReF-ER / Retrace loss of the own samples (f64; the tile kernel's code) ----
  {
    float* vnext = reinterpret_cast<float*>(samp + 11 * kOwn);
    // the ParamLayer values are read at Wp + Lp.imgB: point Wp so that this lands on the common block's copy
    const float* Wp = cm + cp.iP - Lp.imgB;
    LossIO io{act2, err2, info, old, pair, samp, helped ? nullptr : vnext, b0 + s0, pa, pmm, pms, p0s, p0i};
    loss_stages<kOwn, true, true>(a, net, hp, c, step, Wp, io, fetchCtrl, readyFlag, readyTarget);
  }

  // ---- backward, output layer: E_top[k][own samples] = W_out delta_out (Layers.h:131-145), broadcast with delta_out ----
  {
    const float* Wo = cm + cp.iOW;
    for (int idx = tid; idx < cp.oK * kOwn; idx += kST) {
      const int k = idx / kOwn, j = idx - k * kOwn;
      float e = 0.f;
      for (int n = 0; n < cp.oN; ++n) e = fmaf(Wo[k * cp.ldo + n], err2[(Lo.actOff + n) * kOwn + j], e);
      bcast_store(cx, act, Lt.sE + k * kXS + s0 + j, e, 0, kCL);
    }
    for (int idx = tid; idx < cp.oN * kOwn; idx += kST) {
      const int n = idx / kOwn, j = idx - n * kOwn;
      bcast_store(cx, act, cp.sDout + n * kXS + s0 + j, err2[(Lo.actOff + n) * kOwn + j], 0, kCL);
    }
    for (int idx = tid; idx < cp.nP * kOwn; idx += kST) {
      const int i = idx / kOwn, j = idx - i * kOwn;
      bcast_store(cx, act, cp.sGstd + i * kXS + s0 + j, err2[(Lp.actOff + i) * kOwn + j], 0, 1);     // ParamLayer gradient: rank 0 sums it
    }
  }
  DBG_T(a, step, 19);
  cluster_sync_all();
  DBG_T(a, step, 20);

  // ---- backward, hidden layers top down ----
  for (int li = cp.nDense - 1; li >= 0; --li) {
    const CDense& L = cp.L[li];
    float* ds = act + L.sDs;
    if (li < cp.nDense - 1) {
      // error on this layer's output: partial sums from the four slices of the layer above (slot order) + the residual path
      const CDense& Lu = cp.L[li + 1];
      const float* ep = act + L.sEpart;
      const float* eu = act + Lu.sE;
      for (int idx = tid; idx < L.NS * kTS; idx += kST) {
        const int nl = idx >> 3, s = idx & 7, n = rank * L.NS + nl;
        float e = 0.f;
#pragma unroll
        for (int r = 0; r < kCL; ++r) e += ep[(r * L.NS + nl) * kXS + s];
        if (Lu.iRW >= 0 && n < Lu.N) e += eu[n * kXS + s] * cm[Lu.iRW + n];
        if (L.iRW >= 0) bcast_store(cx, act, L.sE + n * kXS + s, e, 0, kCL);     // a residual below needs the full error vector
        else {
          const float y = act[L.sXout + n * kXS + s];
          ds[nl * kXS + s] = e * (1.0f - y * y);
        }
      }
      if (L.iRW >= 0) cluster_sync_all(); else __syncthreads();
    }
    if (li == cp.nDense - 1 || L.iRW >= 0) {     // deltas of the slice from the full error vector
      const float* e = act + L.sE;
      for (int idx = tid; idx < L.NS * kTS; idx += kST) {
        const int nl = idx >> 3, s = idx & 7, n = rank * L.NS + nl;
        const float y = L.iRW >= 0 ? act[L.sYs + nl * kXS + s] : act[L.sXout + n * kXS + s];
        ds[nl * kXS + s] = e[n * kXS + s] * (1.0f - y * y);
      }
      __syncthreads();
    }
    if (li == cp.nDense - 1) DBG_T(a, step, 21);
    if (L.needDx) {
      if (imgBars && li == cp.nDense - 1) cl_mbar_wait(a, &imgBars[1], imgParity);     // the transposed slices
      cl_bwd_dx(L, cp.L[li - 1], rk, act, cx);
      if (li == cp.nDense - 1) DBG_T(a, step, 28);
      cluster_arrive();
      // while the partial input gradients travel: the weight-gradient items that only need this layer's (and the output
      // layer's) deltas — item flag [10] = lowest hidden layer whose deltas the item reads
      if (li == cp.nDense - 1) cl_wgrad(witem, act, wacc, true);
      cluster_wait();
      if (li == cp.nDense - 1) DBG_T(a, step, 22);
    } else if (li == cp.nDense - 1) cl_wgrad(witem, act, wacc, true);
  }
  DBG_T(a, step, 23);

  // ---- the remaining weight-gradient items (accumulated in registers over the passes of the step) ----
  cl_wgrad(witem, act, wacc, false);
  DBG_T(a, step, 4);
}

// ------------------------------------------------------------------------------------------
// Helper CTAs: V(s_t+1) of sampled transitions whose successor is the last row of a truncated episode
// (RACER_train.cpp:22-27), single-CTA forward over all four slices; see next_state_helper for the protocol.
// ------------------------------------------------------------------------------------------
__device__ bool cl_helper(const StepArgs& a, const DevDescs& dd, const ClusterPlan& cp, int step, int helper, int nHelpers,
                          unsigned char* smraw, uint64_t* bars, unsigned parity) {
  const NetDesc& net = dd.net;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  __shared__ int shCount[kST / 32];
  __shared__ int shList[kTS];
  float* img = reinterpret_cast<float*>(smraw + cp.bHelpImg);       // [common | kCL forward parts]
  float* act = reinterpret_cast<float*>(smraw + cp.bHelpAct);
  float* out2 = act + cp.actFloats;                                  // value heads [kOwn] per call of cl_fwd_out
  const ReplayView& rp = a.rp;
  const int dS = net.dS;
  const size_t j0 = (size_t)(step - a.stepBase) * a.B;
  bool loaded = false;
  ClusterCtx cx; cx.rank = 0; cx.share = false;
  for (int first = helper * kTS; ; first += nHelpers * kTS) {
    int total = 0;
    if (tid < kTS) shList[tid] = -1;
    for (int c0 = 0; c0 < a.B; c0 += kST) {
      const int b = c0 + tid;
      const int flagged = (b < a.B) ? (int)((unsigned)__ldcg(a.sampSlot + j0 + b) >> 31) : 0;
      const unsigned m = __ballot_sync(0xffffffffu, flagged);
      __syncthreads();
      if (lane == 0) shCount[warp] = __popc(m);
      __syncthreads();
      int before = total;
      for (int w = 0; w < warp; ++w) before += shCount[w];
      const int rnk = before + __popc(m & ((1u << lane) - 1u));
      if (flagged && rnk >= first && rnk < first + kTS) shList[rnk - first] = b;
      for (int w = 0; w < kST / 32; ++w) total += shCount[w];
    }
    __syncthreads();
    if (total <= first) break;
    if (!loaded) {
      if (tid == 0) {
        asm volatile("fence.proxy.async.global;\n\tfence.proxy.async.shared::cta;" ::: "memory");
        const unsigned cb = 4u * (unsigned)cp.commonFloats, fb = 4u * (unsigned)cp.rankFwdFloats;
        mbar_expect_tx(&bars[0], cb + kCL * fb);
        bulk_g2s(img, a.cimg, cb, &bars[0]);
        for (int r = 0; r < kCL; ++r)
          bulk_g2s(img + cp.commonFloats + (size_t)r * cp.rankFwdFloats, a.cimg + cp.commonFloats + (size_t)r * cp.rankFloats, fb, &bars[0]);
      }
      loaded = true;
    }
    float* x0 = act + cp.sX0;
    for (int idx = tid; idx < dS * kTS; idx += kST) {
      const int s = idx / dS, k = idx - s * dS;
      const int b = shList[s];
      float x = 0.f;
      if (b >= 0) {
        const size_t row = (size_t)__ldcg(a.sampRow + j0 + b) + 1;
        x = (ld_cg(rp.S + row * dS + k) - ld_cg(rp.stateMean + k)) * ld_cg(rp.stateScale + k);
      }
      x0[k * kXS + s] = x;
    }
    if (first == helper * kTS) cl_mbar_wait(a, &bars[0], parity);
    __syncthreads();
    const float* cm = img;
    for (int li = 0; li < cp.nDense; ++li) {
      const CDense& L = cp.L[li];
      for (int r = 0; r < kCL; ++r)
        cl_fwd_hidden(L, img + cp.commonFloats + (size_t)r * cp.rankFwdFloats, cm, act, li == 0 ? x0 : act + cp.L[li - 1].sXout, r, cx);
      __syncthreads();
    }
    const CDense& Lt = cp.L[cp.nDense - 1];
    for (int q = 0; q < kCL; ++q) {
      cl_fwd_out(cp, cm, act + Lt.sXout, q * kOwn, out2, 0, 1);
      __syncthreads();
      if (tid < kOwn && shList[q * kOwn + tid] >= 0) {
        const int b = shList[q * kOwn + tid];
        const size_t row = (size_t)__ldcg(a.sampRow + j0 + b) + 1;
        const float vn = (float)net2v((double)out2[tid]);
        const float qOld = ld_cg(rp.ADV + row) + ld_cg(rp.V + row);
        rp.V[row] = vn; rp.ADV[row] = vn - vn;
        *reinterpret_cast<float2*>(&a.rec[b].qNextOld) = make_float2(qOld, vn);
      }
      __syncthreads();
    }
  }
  return loaded;
}

// ------------------------------------------------------------------------------------------
// P2 of the cluster kernel: one parameter per thread.
// ------------------------------------------------------------------------------------------
struct P2Reg { int p; int iA, iB, iO; float w, m1, m2; };

__device__ __forceinline__ void cl_p2(const StepArgs& a, const ClusterPlan& cp, const NetDesc& net, const Hyper& hp, const StepCtrl& c,
                                      P2Reg& st, int part, int nActive, float* comb, int step) {
  // partial sums of this thread's group of clusters, then the groups in order (deterministic)
  const int per = (nActive + cp.parts - 1) / cp.parts;
  const int c0 = part * per, c1 = min(nActive, c0 + per);
  float v = 0.f;
  if (st.p >= 0) {
    const float* src = a.cpart + (size_t)c0 * net.nParams + st.p;
    int cc = c0;
    for (; cc + 16 <= c1; cc += 16, src += (size_t)16 * net.nParams) {
      float x[16];
#pragma unroll
      for (int u = 0; u < 16; ++u) x[u] = ld_cg(src + (size_t)u * net.nParams);
#pragma unroll
      for (int u = 0; u < 16; ++u) v += x[u];
    }
    for (; cc + 4 <= c1; cc += 4, src += (size_t)4 * net.nParams) {
      float x[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) x[u] = ld_cg(src + (size_t)u * net.nParams);
#pragma unroll
      for (int u = 0; u < 4; ++u) v += x[u];
    }
    for (; cc < c1; ++cc, src += net.nParams) v += ld_cg(src);
  }
  if (cp.parts > 1) {
    comb[threadIdx.x] = v;
    __syncthreads();
    if (part == 0) for (int q = 1; q < cp.parts; ++q) v += comb[q * cp.chunkPad + threadIdx.x];
  }
  if (part != 0 || st.p < 0) return;
  float acc = v;
  int p0 = st.p;
  if (a.comm.world > 1) {      // gradient sum over learner ranks: see p2_tile
    const CommView& cm = a.comm;
    const int N = cm.world, me = cm.rank, rot = step & 3;
    const size_t slotMe = ((size_t)rot * N + me) * cm.nParamsPad;
    const unsigned pk = __float_as_uint(acc);
    for (int q = 0; q < N; ++q) if (q != me) st_volatile_u32(cm.grad(q) + slotMe + p0, pk);
    unsigned* mine = cm.grad(me) + (size_t)rot * N * cm.nParamsPad;
    unsigned got[kMaxWorld];
    if (!wait_values_poison(mine + p0, cm.nParamsPad, N, me, cm, got)) return;      // peer time-out: leave the parameter untouched
    float s = 0.f;
#pragma unroll
    for (int q = 0; q < kMaxWorld; ++q) if (q < N) s += q == me ? acc : __uint_as_float(got[q]);
    acc = s;
  }
  const AdamCoef ac = adam_coef(hp, c);
  a.G[p0] = acc;
  float w, m1, m2;
  adam_step(ac, acc, st.w, st.m1, st.m2, &w, &m1, &m2);
  st.w = w; st.m1 = m1; st.m2 = m2;
  a.W[p0] = w; a.M1[p0] = m1; a.M2[p0] = m2;
  if (st.iA >= 0) a.cimg[st.iA] = w;
  if (st.iB >= 0) a.cimg[st.iB] = w;
  if (st.iO >= 0) a.Wimg[st.iO] = w;
}

// ------------------------------------------------------------------------------------------
// The kernel.  Grid = (P1 clusters + 1) x kCL CTAs; the last CTA of the last cluster is the asynchronous statistics CTA,
// the other CTAs of that cluster are helpers.  All CTAs but the statistics CTA are "workers": they join the two grid
// barriers of a step and own a chunk of parameters in P2.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kST, 1) k_steps_cluster(StepArgs a, int step0, int nSteps, int skipStatsLast) {
  extern __shared__ __align__(128) unsigned char smraw[];
  const NetDesc* netp; const Hyper* hpp;
  load_descs(a, smraw, netp, hpp);
  const DevDescs& dd = *reinterpret_cast<const DevDescs*>(smraw);
  const NetDesc& net = dd.net; const Hyper& hp = dd.hp;
  const size_t descBytes = ((sizeof(DevDescs) + 15) / 16) * 16;
  ClusterPlan& cp = *reinterpret_cast<ClusterPlan*>(smraw + descBytes);
  {
    const int* src = reinterpret_cast<const int*>(a.cplan); int* dst = reinterpret_cast<int*>(&cp);
    for (int i = threadIdx.x; i < (int)(sizeof(ClusterPlan) / 4); i += kST) dst[i] = src[i];
  }
  __syncthreads();
  __shared__ StepCtrl c;
  const int tid = threadIdx.x;
  const int nw = gridDim.x - 1;                       // worker CTAs
  const int nP1c = a.cClusters;                       // P1 clusters
  unsigned* ready = a.barrier + 1;
  const int clusterId = (int)blockIdx.x / kCL;
  if ((int)blockIdx.x == nw) {                        // ---- statistics CTA ----
    float* tiles = reinterpret_cast<float*>(smraw + cp.bHelpImg);      // scratch of the statistics phase (>= 4096 floats)
    // {slot, length} of every position of the episode vector: constant during a launch, kept in shared memory if it fits
    int2* epCache = reinterpret_cast<int2*>(smraw + cp.bHelpImg + 4 * (4096 + 256 * 12));
    if ((size_t)cp.bHelpImg + 4 * (4096 + 256 * 12) + 8 * (size_t)a.nEpisodes <= (size_t)cp.bTotal) {
      for (int p = tid; p < a.nEpisodes; p += kST) { const int sl = a.rp.epOrder[p]; epCache[p] = make_int2(sl, a.rp.epLen[sl]); }
      __syncthreads();
    } else epCache = nullptr;
    for (int s = 0; s < nSteps; ++s) {
      if (skipStatsLast && s == nSteps - 1) break;
      const int step = step0 + s;
      if (tid == 0) {
        const unsigned target = (unsigned)(2 * s + 1) * (unsigned)nw;      // every worker passed barrier 1 of step s
        const long long t0 = clock64();
        while (ld_acquire(a.barrier) < target) { if (clock64() - t0 > kWaitCycles) cl_fail(a, 18); }
        __threadfence();
        load_ctrl(c, &a.ctrl[step & 1]);
      }
      __syncthreads();
      p3_stats(a, hp, c, a.ctrl[(step + 1) & 1], step, tiles, epCache);
      __syncthreads();
      if (tid == 0) {
        __threadfence();
        asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(ready), "r"((unsigned)(s + 1)) : "memory");
      }
    }
    return;
  }
  const bool isP1 = clusterId < nP1c;
  ClusterCtx cx;
  cx.rank = cluster_ctarank(); cx.share = true;
  {
    const uint32_t base = smem_u32(smraw + cp.bAct);
    for (int r = 0; r < kCL; ++r) cx.peer[r] = map_to_rank(base, (unsigned)r);
  }
  __shared__ __align__(8) uint64_t bars[2];
  if (tid == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  // activations: zero once (pad rows / columns are never written afterwards)
  if (isP1) { float* act = reinterpret_cast<float*>(smraw + cp.bAct); for (int i = tid; i < cp.actFloats; i += kST) act[i] = 0.f; }
  else { float* act = reinterpret_cast<float*>(smraw + cp.bHelpAct); for (int i = tid; i < cp.actFloats + 16; i += kST) act[i] = 0.f; }
  __syncthreads();
  if (isP1) cluster_sync_all();                        // nobody writes into a peer before its area is zeroed

  const int nPass = (a.B + kTS - 1) / kTS;             // passes of a step, pass g on cluster g % nP1c
  const int nActive = min(nP1c, nPass);                // clusters that write a partial-gradient row
  const bool hasPass = isP1 && clusterId < nPass;
  const int nHelpers = kCL - 1;
  const int helper = (int)blockIdx.x - nP1c * kCL;     // 0..2 for the helper CTAs
  // P2 ownership
  const int wi = (int)blockIdx.x;                      // worker index
  const int part = tid / cp.chunkPad, pl = tid - part * cp.chunkPad;
  P2Reg st; st.p = -1; st.iA = st.iB = st.iO = -1; st.w = st.m1 = st.m2 = 0.f;
  if (part < cp.parts && pl < cp.chunk) {
    const int p = wi * cp.chunk + pl;
    if (p < net.nParams && __ldg(a.cidx + p) >= -1) {     // -2 in the first map: padding of the parameter blob
      st.p = p;
      if (part == 0) {
        st.iA = __ldg(a.cidx + p); st.iB = __ldg(a.cidx + net.nParams + p); st.iO = __ldg(a.cidx + 2 * net.nParams + p);
        st.w = ld_cg(a.W + p); st.m1 = ld_cg(a.M1 + p); st.m2 = ld_cg(a.M2 + p);
      }
    }
  }
  __shared__ float comb[kST];                                   // P2: partial sums of the cluster groups
  unsigned barTarget = 0;
  unsigned imgParity = 0;
  const CStage stg = cl_stage_view(net, smraw, cp);
  const int dS = net.dS, dA = net.dA, nPair = kOwn * dA;
  const bool pf = hasPass && nPass <= nP1c && (dS & 3) == 0 && kTS * dS / 4 <= kST && nPair <= kST;
  const int b0 = clusterId * kTS, s0 = (int)cx.rank * kOwn;
  const ReplayView& rp = a.rp;
  if (pf) for (int k = tid; k < dS; k += kST) { stg.mean[k] = ld_cg(rp.stateMean + k); stg.scale[k] = ld_cg(rp.stateScale + k); }
  bool staged = false;
  const int q4s = max(dS >> 2, 1);
  const int pfS = tid / q4s, pfC4 = (tid - pfS * q4s) * 4;
  const int pfPs = tid / dA, pfPi = tid - pfPs * dA;
  for (int s = 0; s < nSteps; ++s) {
    const int step = step0 + s;
    DBG_T(a, step, 0);
    if (hasPass) cl_load_image(a, cp, smraw, (int)cx.rank, bars);
    // rows of the next step's samples (prefetch)
    const bool pfNow = pf && s + 1 < nSteps;
    int nxRowS = -1, nxRowT = -1, nxSf = 0, nxRowP = -1, nxRow8 = -1;
    if (pfNow) {
      const size_t jn = (size_t)(step + 1 - a.stepBase) * a.B + b0;
      if (tid >= 64 && tid < 64 + kTS && b0 + tid - 64 < a.B) nxRow8 = ld_index_now(a.sampRow + jn + tid - 64);
      if (tid < kTS * dS / 4 && b0 + pfS < a.B) nxRowS = ld_index_now(a.sampRow + jn + pfS);
      if (tid < kOwn && b0 + s0 + tid < a.B) { nxRowT = ld_index_now(a.sampRow + jn + s0 + tid); nxSf = ld_index_now(a.sampSlot + jn + s0 + tid); }
      if (tid < nPair && b0 + s0 + pfPs < a.B) nxRowP = ld_index_now(a.sampRow + jn + s0 + pfPs);
    }
    if (hasPass) {
      DBG_T(a, step, 1);
      bool first = true;
      float wacc[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) wacc[i] = 0.f;
      for (int g = clusterId; g < nPass; g += nP1c) {
        cl_pass(a, dd, cp, c, step, g * kTS, smraw, cx, first, ready, (unsigned)s, staged, true, wacc, first ? bars : nullptr, imgParity);
        first = false;
        if (g + nP1c < nPass) cluster_sync_all();       // the next pass overwrites buffers the peers may still read
      }
      imgParity ^= 1u;
      cl_wgrad_store(a.citems + ((size_t)cx.rank * kST + tid) * kItemInts, a.cpart + (size_t)clusterId * net.nParams, wacc);
      DBG_T(a, step, 5);
    } else if (!isP1) {
      if (cl_helper(a, dd, cp, step, helper, nHelpers, smraw, bars, imgParity)) imgParity ^= 1u;
    }
    cl_grid_barrier(a, barTarget, (unsigned)nw);
    DBG_T(a, step, 6);
    // ---- next step's inputs -> staging area (cp.async), while P2 runs ----
    if (pfNow) {
      if (tid < kTS * dS / 4) {
        float4* dst = reinterpret_cast<float4*>(stg.S) + tid;
        if (nxRowS >= 0) cp_async16_ca(dst, rp.S + (size_t)nxRowS * dS + pfC4);
        else *dst = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      if (tid >= 64 && tid < 64 + kTS) stg.rows[tid - 64] = nxRow8;
      if (tid < kOwn) {
        const int row = nxRowT, hn = (nxSf >> 31) & 1;
        stg.info[0 * kOwn + tid] = row < 0 ? 0 : row; stg.info[1 * kOwn + tid] = nxSf & 0x7fffffff;
        stg.info[2 * kOwn + tid] = row < 0 ? 0 : hn;  stg.info[3 * kOwn + tid] = row < 0 ? 0 : 1;
        if (row >= 0) {
          const int c0 = row & ~3, c1 = (row + 1) & ~3;
          float* ch = stg.chunks + tid * 4;
          cp_async16_cg(ch + 0 * kOwn * 4, rp.V + c0);     cp_async16_cg(ch + 1 * kOwn * 4, rp.ADV + c0);
          cp_async16_cg(ch + 2 * kOwn * 4, rp.RHO + c0);   cp_async16_cg(ch + 3 * kOwn * 4, rp.KL + c0);
          cp_async16_cg(ch + 4 * kOwn * 4, rp.DELTA + c0); cp_async16_cg(ch + 7 * kOwn * 4, rp.Q + c0);
          if (hn) { cp_async16_cg(ch + 5 * kOwn * 4, rp.V + c1); cp_async16_cg(ch + 6 * kOwn * 4, rp.ADV + c1); }
        }
      }
      if (tid < nPair) {
        if (nxRowP >= 0) {
          const size_t row = nxRowP;
          cp_async4_ca(stg.pair + tid, rp.A + row * dA + pfPi);
          cp_async4_ca(stg.pair + nPair + tid, rp.MU + row * 2 * dA + pfPi);
          cp_async4_ca(stg.pair + 2 * nPair + tid, rp.MU + row * 2 * dA + dA + pfPi);
        } else { stg.pair[tid] = 0.f; stg.pair[nPair + tid] = 0.f; stg.pair[2 * nPair + tid] = 1.f; }
      }
    }
    if (!hasPass) {          // workers without a pass still need this step's Adam scalars
      if (tid == 0) load_ctrl(c, &a.ctrl[step & 1]);
      __syncthreads();
    }
    DBG_T(a, step, 24);
    cl_p2(a, cp, net, hp, c, st, part, nActive, comb, step);
    DBG_T(a, step, 7);
    if (pfNow) {
      cp_async_wait_all();
      if (tid < kOwn) {
        const int row = nxRowT < 0 ? 0 : nxRowT, e0 = row & 3, e1 = (row + 1) & 3, hn = nxRowT < 0 ? 0 : (nxSf >> 31) & 1;
        const float* ch = stg.chunks + tid * 4;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const bool nextRow = j == 5 || j == 6;
          float v = 0.f;
          if (nxRowT >= 0 && (!nextRow || hn)) v = ch[j * kOwn * 4 + (nextRow ? e1 : e0)];
          stg.old[j * kOwn + tid] = v;
        }
      }
    }
    staged = pfNow;
    cl_grid_barrier(a, barTarget, (unsigned)nw);
  }
  if (isP1) cluster_sync_all();          // no CTA leaves while a peer may still address its shared memory
}

// ------------------------------------------------------------------------------------------
// host: plan, image index maps, launch
// ------------------------------------------------------------------------------------------
static inline int ru(int a, int b) { return (a + b - 1) / b * b; }
constexpr size_t kSmemBudgetCluster = 216 * 1024;     // dynamic shared memory (the kernel also has ~8 KB of static scratch)

// Fills `cp` for `net`; idx = three maps [3][nParams]: position of every parameter in the cluster image (first map; -2 =
// blob padding, -1 = none), its second position (transposed copy; -1 = none) and its position in the tile kernel's image.
void cluster_plan_build(const NetDesc& net, int numWorkersHint, ClusterPlan& cp, std::vector<int>& idx, std::vector<int>& items) {
  memset(&cp, 0, sizeof(cp));
  items.assign((size_t)kCL * kST * kItemInts, 0);
  idx.assign((size_t)3 * net.nParams, -1);
  for (int p = 0; p < net.nParams; ++p) idx[p] = -2;
  if (net.func != 0) return;        // the cluster kernel hard-wires Tanh hidden layers
  if (net.recurrent || net.discrete) return;
  // opt-in (SMB200_CLUSTER=1): at B = 256 the persistent tile kernel is still the faster one ("cluster step kernel")
  const char* on = getenv("SMB200_CLUSTER");
  if (!(on && strcmp(on, "1") == 0)) return;
  // layers: input, (dense tanh, [residual])*, dense linear, param
  int nd = 0;
  for (int l = 1; l < net.nLayers; ++l) {
    const LayerDesc& D = net.L[l];
    if (D.kind == kDenseTanh) {
      if (nd >= SMB200_MAX_HIDDEN) return;
      CDense& L = cp.L[nd];
      L.layer = l; L.resLayer = (l + 1 < net.nLayers && net.L[l + 1].kind == kResidual) ? l + 1 : -1;
      L.K = D.nIn; L.N = D.size; L.needDx = D.needDx;
      ++nd;
    } else if (D.kind == kDenseLinear) cp.oLayer = l;
    else if (D.kind == kParam) cp.pLayer = l;
    else if (D.kind != kResidual) return;
  }
  if (nd < 1 || cp.oLayer == 0 || cp.pLayer == 0) return;
  cp.nDense = nd;
  const LayerDesc& O = net.L[cp.oLayer];
  const LayerDesc& P = net.L[cp.pLayer];
  if (O.size > 128) return;
  // geometry
  for (int li = 0; li < nd; ++li) {
    CDense& L = cp.L[li];
    L.NS = ru((L.N + kCL - 1) / kCL, 4);
    L.Np = ru(kCL * L.NS, 16);
    L.Kp = li == 0 ? ru(L.K, 16) : cp.L[li - 1].Np;
    L.ldf = L.NS + 2; L.ldt = L.Kp + 4;
  }
  cp.oK = O.nIn; cp.oKp = cp.L[nd - 1].Np; cp.oN = O.size; cp.oNp = ru(O.size, 4); cp.ldo = cp.oNp + 2; cp.nP = P.size;
  // weight image: common block, then per rank [forward parts | transposed parts]
  int o = 0;
  cp.iOW = o; o += ru(cp.oKp * cp.ldo, 4); cp.iOB = o; o += cp.oNp; cp.iP = o; o += ru(std::max(cp.nP, 1), 4);
  for (int li = 0; li < nd; ++li) {
    CDense& L = cp.L[li];
    if (L.resLayer >= 0) { L.iRW = o; o += L.Np; L.iRB = o; o += L.Np; } else { L.iRW = -1; L.iRB = -1; }
  }
  cp.commonFloats = ru(o, 4);
  o = 0;
  for (int li = 0; li < nd; ++li) { CDense& L = cp.L[li]; L.iWf = o; o += ru(L.Kp * L.ldf, 4); L.iB = o; o += L.NS; }
  cp.rankFwdFloats = ru(o, 4); o = cp.rankFwdFloats;
  for (int li = 0; li < nd; ++li) { CDense& L = cp.L[li]; L.iWt = o; if (L.needDx) o += ru(L.NS * L.ldt, 4); }
  cp.rankFloats = ru(o, 4);
  // activation area
  o = 0;
  cp.sX0 = o; o += cp.L[0].Kp * kXS;
  for (int li = 0; li < nd; ++li) {
    CDense& L = cp.L[li];
    L.sXout = o; o += L.Np * kXS;
    L.sYs = o; if (L.resLayer >= 0) o += L.NS * kXS;
    L.sDs = o; o += L.NS * kXS;
    L.sE = o; if (li == nd - 1 || L.resLayer >= 0) o += L.Np * kXS;
    L.sEpart = o; if (li < nd - 1) o += kCL * L.NS * kXS;
  }
  cp.sDout = o; o += cp.oNp * kXS;
  cp.sGstd = o; o += ru(std::max(cp.nP, 1), 4) * kXS;
  cp.actFloats = ru(o, 4);
  // gradient accumulator
  o = 0;
  for (int li = 0; li < nd; ++li) {
    CDense& L = cp.L[li];
    L.gW = o; o += L.K * (L.NS + 1); L.gB = o; o += L.NS; L.gRW = o; L.gRB = o;
    if (L.resLayer >= 0) { L.gRW = o; o += L.NS; L.gRB = o; o += L.NS; }
  }
  cp.gOW = o; o += cp.L[nd - 1].NS * (cp.oN + 1); cp.gOB = o; o += cp.oN; cp.gP = o; o += std::max(cp.nP, 1);
  cp.gaccFloats = 4;                                  // (the gradient of a step is accumulated in registers: cl_wgrad)
  // shared-memory carve-up
  size_t b = ((sizeof(DevDescs) + 15) / 16) * 16;
  cp.bPlan = (int)b; b += ru((int)sizeof(ClusterPlan), 16);
  cp.bCommon = (int)b; b += 4 * (size_t)cp.commonFloats;
  cp.bRank = (int)b; b += 4 * (size_t)cp.rankFloats;
  cp.bAct = (int)b; b += 4 * (size_t)cp.actFloats;
  cp.bGacc = (int)b; b += 4 * (size_t)cp.gaccFloats;
  cp.bAct2 = (int)b; b += 4 * (size_t)ru(net.actPerSample * kOwn, 4);
  cp.bErr2 = (int)b; b += 4 * (size_t)ru(net.actPerSample * kOwn, 4);
  cp.bInfo = (int)b; b += 4 * 4 * kOwn;
  cp.bOld = (int)b; b += 4 * 8 * kOwn;
  b = (b + 15) / 16 * 16;
  cp.bPair = (int)b; b += 8 * 16 * (size_t)kOwn * net.dA;
  cp.bSamp = (int)b; b += 8 * 12 * kOwn;
  cp.bBars = 0;
  b = (b + 15) / 16 * 16;
  cp.bStage = (int)b;
  cp.stageFloats = kTS * net.dS + 8 * kOwn + 3 * kOwn * net.dA + 4 * kOwn + kTS + 2 * net.dS + 4 + 8 * kOwn * 4;
  b += 4 * (size_t)ru(cp.stageFloats, 4);
  cp.bTotal = (int)b;
  // helper / statistics CTAs: [common | kCL forward parts] + activations (+ 16 floats for the value heads); the statistics
  // phase uses the image region as its 4096-float scratch
  cp.bHelpImg = cp.bCommon;
  size_t hb = cp.bHelpImg + 4 * ((size_t)cp.commonFloats + (size_t)kCL * cp.rankFwdFloats);
  if (hb < (size_t)cp.bHelpImg + 4 * 4096 + 4 * 256 * 12) hb = (size_t)cp.bHelpImg + 4 * 4096 + 4 * 256 * 12;
  cp.bHelpAct = (int)hb; hb += 4 * ((size_t)cp.actFloats + 16);
  cp.bHelpTotal = (int)hb;
  if (cp.bHelpTotal > cp.bTotal) cp.bTotal = cp.bHelpTotal;
  if ((size_t)cp.bTotal > kSmemBudgetCluster) return;
  // P2 partition
  const int nW = std::max(1, numWorkersHint);
  cp.chunk = (net.nParams + nW - 1) / nW;
  cp.chunkPad = ru(cp.chunk, 32);
  if (cp.chunkPad > kST) return;                                    // more parameters than one per thread: tile kernel
  cp.parts = std::max(1, std::min(4, kST / cp.chunkPad));
  // ---- index maps ----
  int* iA = idx.data(); int* iB = iA + net.nParams; int* iO = iB + net.nParams;
  auto rankBase = [&](int r) { return cp.commonFloats + r * cp.rankFloats; };
  for (int li = 0; li < nd; ++li) {
    const CDense& L = cp.L[li];
    const LayerDesc& D = net.L[L.layer];
    for (int k = 0; k < L.K; ++k)
      for (int n = 0; n < L.N; ++n) {
        const int p = D.wOff + k * D.ld + n, r = n / L.NS, nl = n - r * L.NS;
        iA[p] = rankBase(r) + L.iWf + k * L.ldf + nl;
        iB[p] = L.needDx ? rankBase(r) + L.iWt + nl * L.ldt + k : -1;
        iO[p] = D.imgW + k * D.ldp + n;
      }
    for (int n = 0; n < L.N; ++n) {
      const int r = n / L.NS, nl = n - r * L.NS;
      iA[D.bOff + n] = rankBase(r) + L.iB + nl; iO[D.bOff + n] = D.imgB + n;
      if (L.resLayer >= 0) {
        const LayerDesc& R = net.L[L.resLayer];
        iA[R.wOff + n] = L.iRW + n; iO[R.wOff + n] = R.imgW + n;
        iA[R.bOff + n] = L.iRB + n; iO[R.bOff + n] = R.imgB + n;
      }
    }
  }
  for (int k = 0; k < cp.oK; ++k)
    for (int n = 0; n < cp.oN; ++n) { const int p = O.wOff + k * O.ld + n; iA[p] = cp.iOW + k * cp.ldo + n; iO[p] = O.imgW + k * O.ldp + n; }
  for (int n = 0; n < cp.oN; ++n) { iA[O.bOff + n] = cp.iOB + n; iO[O.bOff + n] = O.imgB + n; }
  for (int i = 0; i < cp.nP; ++i) { iA[P.bOff + i] = cp.iP + i; iO[P.bOff + i] = P.imgB + i; }
  // ---- weight-gradient work items, one per thread and rank (cl_wgrad) ----
  for (int r = 0; r < kCL; ++r) {
    int it = 0;
    int early = 0;
    auto put = [&](int kind, int aoff, int astr, int boff, int bstr, int obase, int ostrA, int ostrB, int na, int nb) -> bool {
      if (it >= kST) return false;
      int* q = items.data() + ((size_t)r * kST + it) * kItemInts;
      q[0] = kind; q[1] = aoff; q[2] = astr; q[3] = boff; q[4] = bstr; q[5] = obase; q[6] = ostrA; q[7] = ostrB; q[8] = na; q[9] = nb;
      q[10] = early;
      ++it; return true;
    };
    bool fits = true;
    for (int li = 0; li < nd && fits; ++li) {                  // dW[k][n] of the slice: rows kq + i KQ, columns nq + j NQ
      const CDense& L = cp.L[li];
      const LayerDesc& D = net.L[L.layer];
      const int KQ = L.Kp >> 2, NQ = L.NS >> 2, n0 = r * L.NS;
      const int xoff = li == 0 ? cp.sX0 : cp.L[li - 1].sXout;
      early = li == nd - 1;
      for (int kq = 0; kq < KQ && fits; ++kq)
        for (int nq = 0; nq < NQ && fits; ++nq) {
          int na = 0, nb = 0;
          for (int i = 0; i < 4; ++i) if (kq + i * KQ < L.K) na = i + 1;
          for (int j = 0; j < 4; ++j) if (n0 + nq + j * NQ < L.N) nb = j + 1;
          if (na == 0 || nb == 0) continue;
          fits = put(1, xoff + kq * kXS, KQ * kXS, L.sDs + nq * kXS, NQ * kXS, D.wOff + kq * D.ld + n0 + nq, KQ * D.ld, NQ, na, nb);
        }
    }
    {                                                            // rows of the output layer owned by this rank
      const CDense& Lt = cp.L[nd - 1];
      const int KQ = Lt.NS >> 2, NQ = cp.oNp >> 2, k0 = r * Lt.NS;
      early = 1;
      for (int kq = 0; kq < KQ && fits; ++kq)
        for (int nq = 0; nq < NQ && fits; ++nq) {
          int na = 0, nb = 0;
          for (int i = 0; i < 4; ++i) if (k0 + kq + i * KQ < cp.oK) na = i + 1;
          for (int j = 0; j < 4; ++j) if (nq + j * NQ < cp.oN) nb = j + 1;
          if (na == 0 || nb == 0) continue;
          fits = put(1, Lt.sXout + (k0 + kq) * kXS, KQ * kXS, cp.sDout + nq * kXS, NQ * kXS, O.wOff + (k0 + kq) * O.ld + nq, KQ * O.ld, NQ, na, nb);
        }
    }
    for (int li = 0; li < nd && fits; ++li) {                  // bias (+ residual parameters) of every column of the slice
      const CDense& L = cp.L[li];
      const LayerDesc& D = net.L[L.layer];
      const int xoff = li == 0 ? cp.sX0 : cp.L[li - 1].sXout;
      early = li == nd - 1;
      for (int nl = 0; nl < L.NS && fits; ++nl) {
        const int n = r * L.NS + nl;
        if (n >= L.N) continue;
        if (L.resLayer >= 0) {
          const LayerDesc& R = net.L[L.resLayer];
          // b_0 = error on the residual output (row n of sE), b_1 = the residual's input x[n]
          fits = put(2, L.sDs + nl * kXS, 0, L.sE + n * kXS, (xoff + n * kXS) - (L.sE + n * kXS), D.bOff + n, R.wOff + n, R.bOff + n, 1, 1);
        } else fits = put(2, L.sDs + nl * kXS, 0, 0, 0, D.bOff + n, 0, 0, 1, 0);
      }
    }
    early = 1;
    if (r == 0) {
      for (int n = 0; n < cp.oN && fits; ++n) fits = put(3, cp.sDout + n * kXS, 0, 0, 0, O.bOff + n, 0, 0, 1, 1);
      for (int i = 0; i < cp.nP && fits; ++i) fits = put(3, cp.sGstd + i * kXS, 0, 0, 0, P.bOff + i, 0, 0, 1, 1);
    }
    if (!fits) return;                                           // more than one item per thread: the tile kernel runs this network
  }
  cp.ok = 1;
}

size_t cluster_image_floats(const ClusterPlan& cp) { return (size_t)cp.commonFloats + (size_t)kCL * cp.rankFloats; }

int cluster_prepare(const ClusterPlan& cp) {
  if (!cp.ok) return 0;
  SMB200_CUDA_CHECK(cudaFuncSetAttribute(k_steps_cluster, cudaFuncAttributeMaxDynamicSharedMemorySize, cp.bTotal));
  SMB200_CUDA_CHECK(cudaFuncSetAttribute(k_steps_cluster, cudaFuncAttributeNonPortableClusterSizeAllowed, 0));
  return 0;
}

// clusters of kCL CTAs that can be co-resident (0: the cluster kernel cannot run here)
int cluster_max_active(const ClusterPlan& cp) {
  if (!cp.ok) return 0;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(kCL * 64); cfg.blockDim = dim3(kST); cfg.dynamicSmemBytes = cp.bTotal;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = kCL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, k_steps_cluster, &cfg) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

int launch_steps_cluster(const StepArgs& a, int p1Clusters, int bytes, int step0, int nSteps, int skipStatsLast, cudaStream_t st) {
  SMB200_CUDA_CHECK(cudaMemsetAsync(a.barrier, 0, 2 * sizeof(unsigned), st));
  StepArgs aa = a;
  aa.cClusters = p1Clusters;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((p1Clusters + 1) * kCL); cfg.blockDim = dim3(kST); cfg.dynamicSmemBytes = bytes; cfg.stream = st;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = kCL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  at[1].id = cudaLaunchAttributeCooperative; at[1].val.cooperative = 1;
  cfg.attrs = at; cfg.numAttrs = 2;
  SMB200_CUDA_CHECK(cudaLaunchKernelEx(&cfg, k_steps_cluster, aa, step0, nSteps, skipStatsLast));
  return 0;
}
