// step_kernels.cuh — kernel argument block and host launchers of the learner step.
#pragma once
#include <vector>
#include "common.cuh"
#include "../../include/smarties_b200.h"

namespace smb200 {

// Shared-memory workspace of ONE sampled transition of a recurrent network: every layer's values
// for all steps of the BPTT window (+ the state after the sampled step), offsets in floats.
struct SeqPlan {
  int T;                                          // time slots = Tc + 1
  int yOff[kMaxLayers], yStride[kMaxLayers];      // layer output y at window step k: ws[yOff + k*yStride ...]
                                                  //   (LSTM: [y | cell state | tanh(state)], each roundUp4(nCells))
  int gOff[kMaxLayers], gStride[kMaxLayers];      // LSTM: [cell input | input gate | forget gate | output gate], later the 4 gate deltas
  int eOff[kMaxLayers], eStride[kMaxLayers];      // hidden layers: error on y at every window step
  int actTop, errTop;                             // activations / output gradient at the sampled step, indexed by LayerDesc::actOff
  int red;                                        // 2*threads floats of reduction scratch
  int total;
};
void seq_plan(const NetDesc& net, SeqPlan& p);
int seq_workspace_floats(const NetDesc& net);

struct DevDescs { NetDesc net; Hyper hp; SeqPlan seq; };

// accumulators of the every-1000-steps sweep (sweep_kernels.cu)
struct SweepSums {
  double sumErr2;           // sum of squared Retrace changes (updateReturnEstimator)
  long long nRet;           // number of refreshed estimates
  long long nFarExact;      // exact far-policy flags over all data rows
  double moments[2 * 512 + 3];   // reward/state moments: sum(s-m)[dS], sum((s-m)^2)[dS], count, sum(r-m), sum((r-m)^2)
};

// Peer-memory view for the fused gradient exchange (one process per GPU, buffers shared through
// CUDA IPC).  Every rank owns one identically laid out block:
//   grad  [2 parity][world][nParamsPad] u64   partial gradients, slot q written by rank q: every
//                                             element carries its own stamp (value | (step+1) << 32), so a
//                                             single 8-byte store is data and flag at once (no fence,
//                                             no second round trip — NCCL's LL idea)
//   cnt   [2 parity][world][4] f64            replay counters of the statistics CTA, cntFlag [world] u32
//   vec   [2 parity][world][kCommVec] f64     small host-driven all-reduces (moments), vecFlag [world] u32
constexpr int kMaxWorld = 8;
constexpr int kCommVec = 2 * 512 + 8;
struct CommView {
  int world, rank;
  int nParamsPad, nTilesPad;
  unsigned char* base[kMaxWorld];      // base[q]: rank q's block as mapped into THIS process
  size_t offGrad, gradBytes, offFlag, offCnt, offCntFlag, offVec, offVecFlag, bytes;
  long long timeoutCycles;
  int* error;                          // set to 1 if a peer wait timed out
  __host__ __device__ unsigned* grad(int q) const { return reinterpret_cast<unsigned*>(base[q] + offGrad); }   // [step & 3][rank][nParamsPad]
  __host__ __device__ unsigned* flag(int q) const { return reinterpret_cast<unsigned*>(base[q] + offFlag); }
  __host__ __device__ double* cnt(int q) const { return reinterpret_cast<double*>(base[q] + offCnt); }
  __host__ __device__ unsigned* cntFlag(int q) const { return reinterpret_cast<unsigned*>(base[q] + offCntFlag); }
  __host__ __device__ double* vec(int q) const { return reinterpret_cast<double*>(base[q] + offVec); }
  __host__ __device__ unsigned* vecFlag(int q) const { return reinterpret_cast<unsigned*>(base[q] + offVecFlag); }
};

// ------------------------------------------------------------------------------------------
// Cluster step kernel (feed-forward nets, cluster_step.cuh): every pass of kTS sampled transitions runs on a thread-block
// CLUSTER of kCL CTAs.  Each CTA keeps 1/kCL of the output columns of every hidden layer (its "slice") in shared memory,
// the layer outputs are exchanged through distributed shared memory, and the weight gradient of the pass is formed inside
// the cluster (one partial sum per cluster and parameter instead of one contraction over the whole mini-batch).
// ------------------------------------------------------------------------------------------
constexpr int kCL = 4;            // CTAs per cluster
constexpr int kTS = 8;            // sampled transitions per cluster pass
constexpr int kXS = 12;           // row stride of the activation buffers [feature][kXS]: 8 samples + 4 pad (conflict-free float4 rows)
constexpr int kOwn = kTS / kCL;   // samples whose loss stage a CTA evaluates
struct CDense {                   // one hidden dense layer (+ the ParametricResidual after it)
  int layer, resLayer;            // NetDesc layer ids (resLayer = -1: no residual)
  int K, Kp, N, NS, Np;           // fan-in (Kp: padded to 16), fan-out, slice width per CTA (multiple of 4), Np = roundUp16(kCL * NS)
  int ldf, ldt;                   // row strides: forward slice image [Kp][ldf = NS + 2], transposed [NS][ldt = Kp + 4]
  int needDx;
  int iWf, iB, iWt;               // float offsets in the CTA's block of the weight image
  int iRW, iRB;                   // residual w / b [Np] in the COMMON block (-1)
  int sXout, sYs, sDs, sE, sEpart;   // shared-memory float offsets in the activation area
  int gW, gB, gRW, gRB;           // offsets in the CTA's gradient accumulator: dW [K][NS + 1], db [NS], dRW [NS], dRB [NS]
};
struct ClusterPlan {
  int ok;                         // the cluster kernel covers this network
  int nDense; CDense L[SMB200_MAX_HIDDEN];
  int oLayer, pLayer, oK, oKp, oN, oNp, ldo;   // linear output layer, replicated in every CTA: [oKp][ldo = oNp + 2]
  int nP;                         // ParamLayer size
  int iOW, iOB, iP;               // common block: output weights, output bias, ParamLayer values
  int commonFloats, rankFloats, rankFwdFloats;   // weight image = [common | kCL rank blocks]; a rank block = [forward part | transposed part]
  int sX0, sDout, sGstd, actFloats;
  int gOW, gOB, gP, gaccFloats;   // dWout rows of the CTA's slice of the top layer [NStop][oN + 1]; bias and ParamLayer on rank 0
  int bPlan, bCommon, bRank, bAct, bGacc, bAct2, bErr2, bInfo, bOld, bPair, bSamp, bBars, bStage, bTotal;   // byte offsets
  int stageFloats;
  int bHelpImg, bHelpAct, bHelpTotal;            // helper CTAs: [common | kCL forward parts] + one activation area
  int chunk, chunkPad, parts;     // P2: parameters per worker CTA, padded to a warp multiple, partial-sum groups per parameter
};

// ------------------------------------------------------------------------------------------
// Wide (large-batch) learner step of feed-forward V-RACER / RACER nets (wide_step.cuh): tiles of 128 sampled transitions, every
// dense product of the network as tcgen05.mma kind::tf32 contractions (3xTF32: f32 accuracy) with the weights stationary in
// shared memory, the activations of a tile as the A operand in TENSOR MEMORY, accumulators in tensor memory.
// ------------------------------------------------------------------------------------------
constexpr int kWideM = 128;                      // samples per tile = M of the forward / input-gradient MMAs
constexpr int kWideMaxD = 4;                     // dense layers (hidden layers + the linear output layer) the wide path takes
constexpr int kWideKS = 16;                      // samples per stage of the weight-gradient contraction
constexpr int kWideLD = 130;                     // float4 rows per k-chunk of a staged weight-gradient operand (== 2 mod 8: conflict-free stores)
struct WDense {
  int layer, res;          // NetDesc ids: the dense layer, the ParametricResidual evaluated after it (-1: none)
  int K, Kp, N, Np;        // fan-in (Kp = roundUp16), fan-out (Np = 16 / 32 / 64 / 128)
  int isTanh;
  int inOff, yOff, zOff;   // scratch rows: the layer's input, its own output y, the block output z (= y unless a residual follows)
  int fImg, bImg;          // float offsets of the [hi | lo] operand images: forward float4 [Kp/4][Np], transposed float4 [Np/4][Kp] (-1)
  int vB, vRW, vRB;        // float offsets in the vector block: bias [Np], residual weight / bias [Np] (-1)
  int gCol;                // weight-gradient kernel: first TMEM column of this layer's accumulator
  int gN;                  // its N: hidden layers Kp (D[n][k], M = delta rows), output layer NpG (D[k][j], M = input rows)
  int gPart;               // float offset of the accumulator in a CTA's partial record: [column][128 rows]
  int vSum;                // float offset of this layer's vector sums in the partial record: delta row sums [Np] (hidden layers) /
                           //   output-gradient row sums [NpG] (output layer); then residual bias / weight sums [Np] each
};
struct WidePlan {
  int ok;                  // the wide path covers this network
  int nD; WDense D[kWideMaxD];
  int NpG;                 // rows of the output-gradient operand: roundUp16(nOut) (dense outputs, then the ParamLayer's)
  int fFloats, bFloats, vFloats;        // sizes of the forward image, the transposed image, the vector block
  int vP;                               // ParamLayer values in the vector block
  int vMsc;                             // forward kernel's shared memory: state mean / scale [2][Kp0] behind the vector block
  int recFloats;                        // partial record of one weight-gradient CTA
  int gCols;                            // TMEM columns of the weight-gradient kernel (power of two)
  int sfVec, sfImg, sfBars, sfTotal;                                               // forward kernel: byte offsets in dynamic shared memory
  int sbVec, sbImg, sbBars, sbTotal;                                               // input-gradient kernel
  int sgStage, sgStageBytes, sgStages, sgBars, sgRaw, sgTotal;                      // weight-gradient kernel (sgRaw: ring of two raw stages)
  int sgOpA[kWideMaxD], sgOpB[kWideMaxD];   // byte offsets of a layer's M-side / N-side operand inside a stage ([hi | lo] each)
  int sgRowsA[kWideMaxD], sgRowsB[kWideMaxD];   // staged rows of those operands
};

struct StepArgs {
  // wide step (wide_step.cuh)
  const WidePlan* wplan; float* wimgF; float* wimgB; float* wvec;   // operand images and the vector block (biases, residual vectors, ParamLayer)
  float* wpart; int wGridG;            // [wGridG][recFloats] partial records of the weight-gradient kernel
  const int* widx;                     // [5][nParams]: record position, image positions (tile image, forward, transposed, vector block)
  int* wcnt;                           // [0] flagged samples of the step (V(s_t+1) list), [1] exact far-policy flag changes
  int* wlist;                          // [B] flagged samples
  // cluster step kernel
  const ClusterPlan* cplan; float* cimg; float* cpart; const int* cidx;   // image, per-cluster partial gradients [clusters][nParams], image positions [3][nParams]
  const int* citems;                 // weight-gradient work items [kCL][threads][12]
  int cClusters;                     // P1 clusters of the launch
  CommView comm;
  const DevDescs* descs;
  ReplayView rp;
  float* W; float* Wimg; float* M1; float* M2; float* G;   // Wimg: weights in the shared-memory image layout
  float* Wtgt; double tgtAlpha; long long tgtPhase;        // "targetDelay" > 0: AdamOptimizer::target_weights, its rate / period, the
                                                           //   update count at which cntUpdateDelay was (re)set to 0 (construction, restart)
  float* actG; float* errG;          // feature-major scratch [actPerSample][Bpad]
  const int* sampRow; const int* sampSlot;  // [nSteps][B]: ring row of (episode,t); slot | hasNext<<31
  SampleRec* rec;                    // [B]
  float* lastO; float* lastG; float* lastX;
  StepCtrl* ctrl;                    // [2]
  smb200_step_stats* statsOut;       // [nSteps] or nullptr
  const GradTile* tiles; int nTiles;
  int B, Bpad;
  int nEpisodes; long long nTransitions;   // table during the segment (before this step's pruning)
  long long nTransitionsPost;              // nStoredSteps() after the LAST step's pruning (beta update)
  int stepBase;                            // samples / statsOut are indexed by (step - stepBase)
  int lastStep;                            // absolute index of the segment's last step
  unsigned* barrier;
  long long* dbgT;                         // optional phase timestamps [step][cta][8] (clock64)
  int useTma;                              // weight image to shared memory by cp.async.bulk (1) or ld.global.cg (0)
  int statsIncremental;                    // persistent kernel: launch-resident statistics (StatKeep) instead of a full scan per step
  // recurrent nets: weight gradient of the LSTM layers on the tensor cores (tcgen05, see tc_wgrad_item)
  int useTc; float* tcPartial;             // [item][128][128] f32 partial tiles of the K-slices
};

int step_threads();
// Work decomposition of the tensor-core weight gradient: per LSTM layer, n-tiles of 128 gate columns x K-slices of Wc
// scratch columns; shared by the host (allocation of the partial tiles) and the kernel.
struct TcPlan { int Wc; int nItems; int item0[kMaxLayers]; int slices[kMaxLayers]; int nT[kMaxLayers]; };
__host__ __device__ inline TcPlan tc_plan(const NetDesc& net, int cols, size_t stagingBytes) {
  TcPlan p; p.nItems = 0;
  int Wc = (int)(stagingBytes / 2064) / 8 * 8;          // (Wc/4) * (129 + 129) float4, hi and lo images
  if (Wc > 96) Wc = 96;
  p.Wc = Wc;
  for (int l = 0; l < kMaxLayers; ++l) { p.item0[l] = 0; p.slices[l] = 0; p.nT[l] = 0; }
  if (Wc < 8) return p;
  for (int l = 1; l < net.nLayers; ++l) {
    const LayerDesc& L = net.L[l];
    if (L.kind != kLSTM || L.nIn + L.size + 1 > 128) continue;   // the 128 MMA rows hold [x | h_prev | 1]
    p.item0[l] = p.nItems; p.slices[l] = (cols + Wc - 1) / Wc; p.nT[l] = (4 * L.size + 127) / 128;
    p.nItems += p.slices[l] * p.nT[l];
  }
  return p;
}
size_t tc_staging_bytes(const NetDesc& net);
// eta of the Adam update that follows `adam_step` completed updates (host and device use the same IEEE ops)
__host__ __device__ inline float adam_eta_for(double learnrate, double epsAnneal, long long adam_step, double bt1d, double bt2d) {
  const long long nStep = adam_step + 1;                                      // prepare_update: nStep++ (Optimizer.cpp:119)
  const float etaf = (float)learnrate;
  const float eta0 = (float)((double)etaf / (1.0 + (double)(float)(double)nStep * epsAnneal));   // annealRate<nnReal>
  const float bt1 = (float)bt1d, bt2 = (float)bt2d;
  return eta0 * sqrtf(1.0f - bt2) / (1.0f - bt1);                            // struct Adam ctor (Optimizer.cpp:64-67)
}
size_t step_smem_bytes(const NetDesc& net, int TB);
int step_kernels_prepare(const NetDesc& net);
int launch_step_two_kernels(const StepArgs& a, const NetDesc& net, int step, int skipStats, cudaStream_t st);
int persistent_grid(const StepArgs& a, const NetDesc& net, int numSMs);
int launch_steps_persistent(const StepArgs& a, const NetDesc& net, int grid, int step0, int nSteps, int skipStatsLast, cudaStream_t st);
int launch_finalize_sweep(const StepArgs& a, int step, const SweepSums* sweep, cudaStream_t st);
int launch_forward(const StepArgs& a, const NetDesc& net, const float* states, int n, float* out, cudaStream_t st);
int launch_forward_seq(const StepArgs& a, const NetDesc& net, const float* states, const int* lengths, int n, int maxLen, float* out,
                       cudaStream_t st);
// cluster step kernel (cluster_step.cuh)
void cluster_plan_build(const NetDesc& net, int numWorkersHint, ClusterPlan& cp, std::vector<int>& idx, std::vector<int>& items);
size_t cluster_image_floats(const ClusterPlan& cp);
int cluster_prepare(const ClusterPlan& cp);
int cluster_max_active(const ClusterPlan& cp);
int launch_steps_cluster(const StepArgs& a, int p1Clusters, int bytes, int step0, int nSteps, int skipStatsLast, cudaStream_t st);

// wide step (wide_step.cuh)
void wide_plan_build(const NetDesc& net, const Hyper& hp, WidePlan& wp, std::vector<int>& idx);
void wide_fill_images(const NetDesc& net, const WidePlan& wp, const std::vector<int>& idx, const float* blob,
                      std::vector<float>& imgF, std::vector<float>& imgB, std::vector<float>& vecs);
int wide_prepare(const WidePlan& wp, const NetDesc& net);
int wide_grid_g(const WidePlan& wp, int B, int numSMs);
int launch_steps_wide(const StepArgs& a, const NetDesc& net, const WidePlan& wp, int numSMs, int step0, int nSteps, int skipStatsLast,
                      cudaStream_t st, cudaStream_t aux, cudaEvent_t evF, cudaEvent_t evN, cudaEvent_t evS);

// sweep_kernels.cu
int launch_init_episode(const ReplayView& rp, int slot, float deltaInit, int haveValues, cudaStream_t st);
// Retrace (+ optional exact aggregate recompute) over `nEpisodes` episodes listed in rp.epOrder,
// or over the single episode `oneSlot` when nEpisodes == 0.  gae != 0: the GAE recursion instead (no importance weight,
// no advantage term; MemoryProcessing.cpp:411-417).  estimator = smb200_returns_estimator; 2 (retraceExplore, :402-409) runs
// k_sweep_explore with the baseline `stats.maxAbsError` taken from exploreCtrl (device) or, if null, exploreBaseline.
int launch_sweep(const ReplayView& rp, int nEpisodes, int oneSlot, float gamma, float lambda, int estimator,
                 int recomputeAggregates, float cmax, float cinv, SweepSums* sums, cudaStream_t st,
                 float exploreBaseline = 0.f, const StepCtrl* exploreCtrl = nullptr);
int launch_moments(const ReplayView& rp, long long rowEnd, SweepSums* sums, int numSMs, cudaStream_t st);
// Retrace / GAE + exact aggregates + reward / state moments of every episode in ONE pass (k_sweep_fused)
bool sweep_fused_supported(const ReplayView& rp, int estimator);
int launch_sweep_fused(const ReplayView& rp, int nEpisodes, float gamma, float lambda, int estimator, float cmax, float cinv,
                       SweepSums* sums, int numSMs, cudaStream_t st);
int launch_update_scaling(const ReplayView& rp, const StepCtrl* ctrlCur, const DevDescs* descs, const SweepSums* sums, int bInit, cudaStream_t st);
int launch_clear_sums(SweepSums* sums, cudaStream_t st);
// sum a small f64 vector over all ranks through peer memory (in place, identical result on every rank)
int launch_peer_allreduce(const CommView& comm, double* vec, int n, unsigned stamp, cudaStream_t st);

}  // namespace smb200
