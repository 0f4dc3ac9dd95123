// learner.cu — host side of the C-ABI (include/smarties_b200.h): the replay MemoryBuffer's
// episode table and ring allocator, the bit-exact host sampler, FIFO pruning, weight
// initialisation, and the orchestration of the step / sweep kernels on one CUDA stream.
//
// Mirrors (file:line relative to /root/reference/source/smarties/):
//   Learner::initializeLearner / processMemoryBuffer     Learners/Learner.cpp:47-100
//   Learner_approximator::spawnTrainTasks / applyGradient Learners/Learner_approximator.cpp:36-105
//   MemoryBuffer::pushBackEpisode / removeBackEpisode     ReplayMemory/MemoryBuffer.cpp:469-520
//   Sample_uniform::sample, Sampling::IDtoSeqStep         ReplayMemory/Sampling.cpp:26-47,82-96
//   MemoryProcessing::applyEpisodesRemovalAlgo            ReplayMemory/MemoryProcessing.cpp:327-351
//   Builder::build weight init, BaseLayer::initialize     Network/Builder.cpp:119-170, Layers/Layer_Base.h:115-141
#include "step_kernels.cuh"

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <chrono>
#include <map>
#include <mutex>
#include <random>
#include <limits>
#include <string>
#include <tuple>
#include <vector>

namespace smb200 {

static thread_local std::string g_err;
void set_error(const char* what, cudaError_t e, const char* file, int line) {
  char buf[1024];
  snprintf(buf, sizeof(buf), "CUDA error %d (%s) at %s:%d: %s", (int)e, cudaGetErrorString(e), file, line, what);
  g_err = buf;
}
void set_error_msg(const char* msg) { g_err = msg; }

static inline int round_up(int n, int m) { return (n + m - 1) / m * m; }

struct EpisodeMeta {
  long long id;
  int nRows;       // nsteps(): data steps + terminal row
  int slot;
  long long start; // first ring row
  int terminated;
  int agentId;     // Episode::agentID (only carried through checkpoints)
};

template <typename T>
static int dev_alloc(T** p, size_t n) {
  SMB200_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(p), n * sizeof(T)));
  SMB200_CUDA_CHECK(cudaMemset(*p, 0, n * sizeof(T)));
  return 0;
}

}  // namespace smb200

using namespace smb200;

struct smb200_learner {
  smb200_config cfg;
  DevDescs descs;                 // host copy
  DevDescs* dDescs = nullptr;
  int numSMs = 148;
  int mode = 1;                   // 1 = persistent cooperative kernel, 0 = two kernels per step
  int persistGrid = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  cudaEvent_t evDone[2] = {nullptr, nullptr};   // pipeline: segment buffers free again

  // replay
  ReplayView rp{};
  std::vector<EpisodeMeta> episodes;      // the reference's `episodes` vector order
  std::vector<int> freeSlots;
  std::map<long long, long long> liveRanges;   // start -> end (exclusive) of live ring ranges
  long long head = 0, highWater = 0;
  long long nTransitions = 0;
  long long nGatheredB4Startup = 0;       // counters.nGatheredB4Startup, set by initializeLearner (Learner.cpp:60)
  long long nSeenEps = 0, nSeenObs = 0;   // counters.nSeenEpisodes_loc / nSeenTransitions_loc (ReplayStatsCounters.h:43-50)
  std::vector<float> tgtBlob;             // AdamOptimizer::target_weights: with targetDelay 0 the weights at construction / restart
  bool orderDirty = true;

  // network / optimiser
  float *W = nullptr, *Wimg = nullptr, *M1 = nullptr, *M2 = nullptr, *G = nullptr;
  float* Wtgt = nullptr; long long tgtPhase = 0;     // "targetDelay" > 0: target weights on the device (see StepArgs)
  int statsIncremental = 1;                          // SMB200_STATS_FULL=1: full scan of the episode aggregates every step
  long long* dDbg = nullptr; int useTma = 1;
  // multi-rank gradient exchange over peer memory (CUDA IPC)
  CommView comm{}; unsigned char* commBuf = nullptr; int* dCommErr = nullptr; unsigned vecStamp = 0;
  void* peerMapped[kMaxWorld] = {nullptr};
  float *actG = nullptr, *errG = nullptr;
  float* tcPartial = nullptr; int useTc = 0;     // tensor-core weight gradient of LSTM layers (recurrent nets)
  GradTile* dTiles = nullptr; int nTiles = 0;
  int Bpad = 0;
  // cluster step kernel (feed-forward nets, cluster_step.cuh)
  ClusterPlan cplan{}; ClusterPlan* dCplan = nullptr; std::vector<int> cidx, citems; int* dCidx = nullptr; int* dCitems = nullptr;
  float* cimg = nullptr; float* cpart = nullptr; int clusterP1 = 0;      // clusterP1 > 0: the cluster kernel runs the steps
  // wide (large-batch) step on the tensor cores (wide_step.cuh)
  WidePlan wplan{}; WidePlan* dWplan = nullptr; std::vector<int> widx; int* dWidx = nullptr;
  float *wimgF = nullptr, *wimgB = nullptr, *wvec = nullptr, *wpart = nullptr; int* wcnt = nullptr; int* wlist = nullptr;
  int wGridG = 0; int wideOn = 0;          // wideOn: the wide kernels run the steps
  cudaStream_t wAux = nullptr; cudaEvent_t wEv[3] = {nullptr, nullptr, nullptr};   // auxiliary stream of the wide step (V(s_t+1), records, statistics)
  int dP = 0;                     // columns of the behaviour policy MU: 2 * dim_action (mean, stdev), or the K option probabilities

  // step state
  StepCtrl* dCtrl = nullptr;       // [2]
  StepCtrl hCtrl{};                // host mirror, valid when ctrlSynced
  long long gradStep = 0;          // host mirror of counters.nGradSteps
  bool initialized = false;
  SampleRec* dRec = nullptr;
  float *lastO = nullptr, *lastG = nullptr, *lastX = nullptr;
  SweepSums* dSums = nullptr;
  unsigned* dBarrier = nullptr;
  int* dSampSlot = nullptr; int* dSampT = nullptr;
  int* hSampSlot = nullptr; int* hSampT = nullptr;    // pinned
  smb200_step_stats* dStats = nullptr; smb200_step_stats* hStats = nullptr;   // pinned host
  int maxSeg = 1024;
  int presampled = 0;              // steps resident in dSampSlot/dSampT (benchmark path)

  // prioritized samplers / non-FIFO filters: host copies of the keys, refreshed after every step
  std::vector<float> keyDelta;           // deltaValue of every ring row (PERrank / PERerr: Episode::SquaredError)
  std::vector<float> keyAgg[3];          // per slot: fracFarPolSteps, avgKLDivergence, avgSquaredErr (filters, PERseq)
  std::discrete_distribution<size_t> perDist;
  bool samplerStale = false;             // a restart: the distribution is prepared before the first step
  bool slow_mode() const { return cfg.data_sampling != SMB200_SAMPLE_UNIFORM || cfg.er_filter != SMB200_FILTER_OLDEST; }
  std::vector<std::pair<long long, int>> pendingEvict;   // ring ranges whose live flags are cleared after the segment
  std::mt19937 gen;
  // host sampler scratch + id -> (episode position, t) lookup, rebuilt when the episode table changes
  std::vector<size_t> sampIds, sampTmp;
  std::vector<long long> epPrefix; std::vector<int> bucketFirst; int bucketShift = 0; bool lookupDirty = true;
  double lastMs = 0; long long lastLaunches = 0;
  long long launches = 0;
  // Sample-ahead queue (uniform sampler, FIFO filter): mini-batches of the NEXT learner steps, drawn from a copy of the
  // sampler's generator while the host would otherwise wait for the device at the end of a train call.  `gen` always stays at
  // the true position of the reference's stream: a consumed step advances it by the number of 32-bit draws the step took
  // (discard), so dropping the queue (an episode arrives, the sampler is re-seeded, ...) needs no repair.
  std::vector<int> aheadSlot, aheadRow; std::vector<unsigned> aheadDraws; std::vector<unsigned char> aheadEnds;
  int aheadHead = 0, aheadCnt = 0; std::mt19937 aheadGen; long long aheadGradStep = 0, aheadVersion = -1;
  long long tableVersion = 0;              // bumped whenever the episode table or its order changes
  long long aheadUsed = 0, aheadDrawn = 0; // diagnostics: steps served from the queue / drawn into it
  void ahead_clear() { aheadHead = aheadCnt = 0; aheadVersion = -1; }

  // actor-side policy evaluation (smb200_forward / smb200_forward_seq): staging buffers that grow on demand, and the lock that
  // lets actor threads push episodes and evaluate the policy while the learner thread trains (one stream, one call at a time)
  std::recursive_mutex apiMutex;
  float *fwdIn = nullptr, *fwdOut = nullptr, *hFwdIn = nullptr, *hFwdOut = nullptr; int *fwdLen = nullptr, *hFwdLen = nullptr;
  size_t fwdInCap = 0, fwdOutCap = 0, fwdLenCap = 0;

  // output-gradient statistics file of the reference (StatsTracker, Utils/StatsTracker.cpp:28-107); off until
  // smb200_set_grad_stats names the file.  trackerSteps = StatsTracker::nStep (reduce_stats calls since construction).
  std::string gradStatsBase; long long trackerSteps = 0;
  float* hGradStat[2] = {nullptr, nullptr};      // pinned [B][nOut], one per pipeline half
  bool grad_stats_step(long long nGradSteps) const { return !gradStatsBase.empty() && nGradSteps % 1000 == 0; }

  StepArgs args() const {
    StepArgs a{};
    a.comm = comm;
    a.descs = dDescs; a.rp = rp; a.W = W; a.Wimg = Wimg; a.M1 = M1; a.M2 = M2; a.G = G; a.dbgT = nullptr; a.useTma = useTma;
    a.Wtgt = Wtgt; a.tgtAlpha = cfg.target_delay; a.tgtPhase = tgtPhase;
    a.statsIncremental = statsIncremental;
    a.useTc = useTc; a.tcPartial = tcPartial;
    a.wplan = dWplan; a.wimgF = wimgF; a.wimgB = wimgB; a.wvec = wvec; a.wpart = wpart; a.wGridG = wGridG; a.widx = dWidx;
    a.wcnt = wcnt; a.wlist = wlist;
    a.cplan = dCplan; a.cimg = cimg; a.cpart = cpart; a.cidx = dCidx; a.citems = dCitems; a.cClusters = clusterP1;
    a.actG = actG; a.errG = errG; a.sampRow = dSampT; a.sampSlot = dSampSlot; a.rec = dRec;
    a.lastO = lastO; a.lastG = lastG; a.lastX = lastX; a.ctrl = dCtrl; a.statsOut = dStats;
    a.tiles = dTiles; a.nTiles = nTiles; a.B = cfg.batch_size; a.Bpad = Bpad;
    a.nEpisodes = (int)episodes.size(); a.nTransitions = nTransitions; a.nTransitionsPost = nTransitions;
    a.barrier = dBarrier; a.stepBase = 0; a.lastStep = -1;
    return a;
  }
};

namespace smb200 {

// ---- network description: RACER::setupNet + Approximator::buildFromSettings + Builder::addLayer ----
// does any ParametricResidual link a narrower layer (hidden sizes that grow)?  The cluster kernel and the wide step assume
// equal or shrinking widths; such nets run on the tile kernels.
static bool residual_widens(const NetDesc& net) {
  for (int l = 2; l < net.nLayers; ++l)
    if (net.L[l].kind == kResidual && net.L[l - 2].size < net.L[l].size) return true;
  return false;
}

static int build_net(const smb200_config& c, NetDesc& net, std::vector<GradTile>& tiles) {
  memset(&net, 0, sizeof(net));
  const int dS = c.dim_state, dA = c.dim_action;
  if (c.algo != SMB200_VRACER && c.algo != SMB200_RACER) { set_error_msg("learner must be VRACER or RACER"); return -1; }
  if (dS < 1 || dA < 1 || dA > SMB200_MAX_ACTION || c.n_hidden < 1 || c.n_hidden > SMB200_MAX_HIDDEN) {
    set_error_msg("unsupported dimensions"); return -1; }
  int off = 0, img = 0, act = 0, id = 0, width = dS;
  auto add = [&](int kind, int size) -> LayerDesc& {
    LayerDesc& L = net.L[id]; L.kind = kind; L.size = size; L.actOff = act; L.needDx = 0; L.in = id - 1;
    act += round_up(size, 4); width = std::max(width, size); ++id; return L; };
  // weight image (shared-memory layout): dense rows padded to roundUp4(size)+4 floats
  const int NT = step_threads();
  auto log2_group = [&](int n) { int sh = 3; while ((1 << sh) < std::min(n, NT)) ++sh; return sh; };
  auto img_dense = [&](LayerDesc& L) {
    L.fwdShift = log2_group(L.size); L.bwdShift = log2_group(L.nIn);
    L.ldp = round_up(L.size, 4) + 4;
    L.imgW = img; img += L.ldp * L.nIn; L.imgB = img; img += round_up(L.size, 4); };
  add(kInput, dS);
  int nIn = dS;
  const bool lstm = c.nn_type == SMB200_LSTM || c.nn_type == SMB200_MGU;      // recurrent-cell layers
  const int cellKind = c.nn_type == SMB200_MGU ? kMGU : kLSTM, ng = cell_gates(cellKind);
  if (c.nn_type != SMB200_FFNN && !lstm) { set_error_msg("nnType must be FFNN, LSTM, MGU or GRU"); return -1; }
  if (c.nn_func < SMB200_TANH || c.nn_func > SMB200_LINEAR || (c.nn_func != SMB200_TANH && lstm)) {
    set_error_msg("nnFunc must be one of makeFunction's ten names (recurrent cells: Tanh)"); return -1; }
  if (lstm && (c.nn_bptt_seq < 0 || c.nn_bptt_seq > 255)) { set_error_msg("nnBPTTseq out of range"); return -1; }
  if (lstm) for (int i = 0; i < c.n_hidden; ++i) if (c.hidden[i] > NT) { set_error_msg("LSTM layers wider than the CTA are not supported"); return -1; }
  for (int i = 0; i < c.n_hidden; ++i) {
    const int h = c.hidden[i];
    if (h < 1) { set_error_msg("hidden layer size must be positive"); return -1; }
    if (lstm) {   // LSTMLayer (Layers/Layer_LSTM.h): W[(nIn + nCells)][4 nCells] then 4 nCells biases; MGULayer (Layer_GRU.h): 2 nCells
      LayerDesc& L = add(cellKind, h);
      // activation rows: [y | h_prev | (MGU: h_prev * forget) | --], delta rows: the gates
      act += round_up(4 * h, 4) - round_up(h, 4);
      L.nIn = nIn; L.ld = ng * h;
      L.wOff = off; off += round_up(ng * h * (nIn + h), 8); L.bOff = off; off += round_up(ng * h, 8);
      L.needDx = i > 0;
      L.fwdShift = log2_group(ng * h); L.bwdShift = log2_group(h);
      L.ldp = round_up(ng * h, 4) + 4;
      L.imgW = img; img += L.ldp * (nIn + h); L.imgB = img; img += round_up(ng * h, 4);
    } else {
      LayerDesc& L = add(kDenseTanh, h);
      L.nIn = nIn; L.ld = round_up(h, 8);
      L.wOff = off; off += round_up(L.ld * nIn, 8); L.bOff = off; off += round_up(h, 8);
      L.needDx = i > 0; img_dense(L);
    }
    if (i > 0) {   // ParametricResidualLayer after every hidden layer but the first (Builder.cpp:92-95)
      LayerDesc& R = add(kResidual, h);
      R.wOff = off; off += round_up(h, 8); R.bOff = off; off += round_up(h, 8);
      R.imgW = img; img += round_up(h, 4); R.imgB = img; img += round_up(h, 4);
      // a residual over a NARROWER dense layer (hidden sizes that grow) links the first min(size below, size) units only
      // (ParametricResidualLayer::forward, Layers.h:347-361): handled by the tile kernels; see residual_widens().  Over a
      // narrower recurrent-cell layer the reference's `sizes[ID-2]` is the cell layer's whole work array (4 nCells: output |
      // cell state | ...), so its extra units link to the lower layer's CELL STATES — an accident that is not reproduced.
      if (lstm && net.L[id - 3].size < h) { set_error_msg("residual over a narrower recurrent-cell layer is not supported"); return -1; }
    }
    nIn = h;
  }
  // V-RACER: [V | mean(dA)], RACER: [V | adv coef, p1(dA), p2(dA) | mean(dA)]; stdev is the ParamLayer (RACER_simpleSigma)
  // discrete action with K options (RACER_common.cpp:109-135): [V | advantages(K) | policy(K)], no ParamLayer — kept here as
  // an EMPTY last layer (0 parameters, 0 activations: the blob layout is the reference's) so that "last layer = ParamLayer" holds
  const int K = c.discrete_options;
  if (K < 0 || K > 64 || (K > 0 && (dA != 1 || c.algo != SMB200_RACER || K < 2))) {
    set_error_msg("discrete actions: one action component, 2..64 options, learner RACER"); return -1; }
  const int nOutDense = K > 0 ? 1 + 2 * K : (c.algo == SMB200_RACER ? 2 + 3 * dA : 1 + dA);
  const int nParamOut = K > 0 ? 0 : dA;
  {
    LayerDesc& L = add(kDenseLinear, nOutDense);
    L.nIn = nIn; L.ld = round_up(nOutDense, 8);
    L.wOff = off; off += round_up(L.ld * nIn, 8); L.bOff = off; off += round_up(nOutDense, 8);
    L.needDx = 1; img_dense(L);
    LayerDesc& P = add(kParam, nParamOut);
    P.bOff = off; off += round_up(nParamOut, 8); P.wOff = off;
    P.imgB = img; P.imgW = img; img += round_up(nParamOut, 4);
  }
  net.nLayers = id; net.nParams = off; net.imgFloats = img; net.nOut = nOutDense + nParamOut; net.nOutDense = nOutDense;
  net.dS = dS; net.dA = dA; net.maxWidth = width; net.discrete = K; net.func = c.nn_func;
  net.recurrent = lstm ? 1 : 0; net.bptt = lstm ? c.nn_bptt_seq : 0; net.Tc = net.bptt + 1;
  net.topInOff = act;
  if (lstm) act += round_up(nIn, 4);      // compact copy of the top hidden layer's output at the sampled step
  net.actPerSample = act;
  net.seqFloats = lstm ? seq_workspace_floats(net) : 0;
  tiles.clear();
  for (int l = 1; l < net.nLayers; ++l) {
    const LayerDesc& L = net.L[l];
    if (L.kind == kDenseTanh || L.kind == kDenseLinear) {
      for (int k0 = 0; k0 < L.nIn + 1; k0 += kTileK)
        for (int n0 = 0; n0 < L.size; n0 += kTileN) tiles.push_back(GradTile{0, l, k0, n0});
    } else if (L.kind == kResidual) {
      for (int n0 = 0; n0 < L.size; n0 += 16) tiles.push_back(GradTile{1, l, 0, n0});
    } else if (L.kind == kParam) {
      for (int n0 = 0; n0 < L.size; n0 += 16) tiles.push_back(GradTile{2, l, 0, n0});
    } else if (L.kind == kLSTM) {
      for (int k0 = 0; k0 < L.nIn + L.size + 1; k0 += kTileK)
        for (int n0 = 0; n0 < 4 * L.size; n0 += kTileN) tiles.push_back(GradTile{0, l, k0, n0});
    } else if (L.kind == kMGU) {      // the recurrent rows of the two column halves multiply different inputs: h_prev | h_prev * forget
      for (int k0 = 0; k0 < L.nIn + L.size + 1; k0 += kTileK)
        for (int half = 0; half < 2; ++half)
          for (int n0 = half * L.size; n0 < (half + 1) * L.size; n0 += kTileN) {
            GradTile t{0, l, k0, n0}; t.nLimit = (half + 1) * L.size; tiles.push_back(t);
          }
    }
  }
  // contraction length of every tile: the mini-batch, or for recurrent networks all (sample, window step)
  // columns; the output and stdev layers only receive a gradient at the sampled step (compact columns)
  const int colsB = round_up(c.batch_size, 256);
  const int colsAll = lstm ? round_up(c.batch_size * net.Tc, 256) : colsB;
  for (auto& t : tiles) t.cols = (net.L[t.layer].kind == kDenseLinear || net.L[t.layer].kind == kParam) ? colsB : colsAll;
  return 0;
}

// Builder::build: layers initialise in order from generators[0] (Builder.cpp:133-137);
// BaseLayer::initialize draws weight[o + ld*i] for i then o (Layer_Base.h:115-141).
static void init_weights(const smb200_config& c, const NetDesc& net, std::mt19937& gen, std::vector<float>& blob) {
  blob.assign(net.nParams, 0.f);
  for (int l = 1; l < net.nLayers; ++l) {
    const LayerDesc& L = net.L[l];
    if (L.kind == kDenseTanh || L.kind == kDenseLinear) {
      const bool out = L.kind == kDenseLinear;
      const double prefac = out ? c.out_weights_prefac : 1.0;
      const float fac = prefac > 0 ? (float)prefac : 1.f;
      // Function::initFactor (Functions.h): Linear, LRelu sqrt(1 / in); Tanh, Sigm, HardSign, SoftSign sqrt(6 / (in + out));
      // Relu, ExpPlus, SoftPlus, Exp sqrt(2 / in)
      const bool in1 = out || c.nn_func == SMB200_LRELU || c.nn_func == SMB200_LINEAR;
      const bool in2 = c.nn_func == SMB200_RELU || c.nn_func == SMB200_EXPPLUS || c.nn_func == SMB200_SOFTPLUS || c.nn_func == SMB200_EXP;
      const double initFactor = in1 ? std::sqrt(1. / L.nIn) : (in2 ? std::sqrt(2. / L.nIn) : std::sqrt(6. / (L.nIn + L.size)));
      const float init = (float)(fac * initFactor);
      std::uniform_real_distribution<float> dis(-init, init);
      for (int i = 0; i < L.nIn; ++i)
        for (int o = 0; o < L.size; ++o) blob[L.wOff + o + L.ld * i] = dis(gen);
      if (out && c.algo == SMB200_RACER && c.discrete_options == 0) {   // Gaussian_advantage::setInitial (Gaus_advantage.h:31-34): bias -1 for the coefficient, 1 for the widths
        blob[L.bOff + 1] = -1.f;
        for (int o = 2; o < 2 + 2 * c.dim_action; ++o) blob[L.bOff + o] = 1.f;
      }
    } else if (L.kind == kLSTM) {   // LSTMLayer::initialize (Layer_LSTM.h:168-188): gates primed with LSTM_PRIME_FAC = 1 (Bund.h:63)
      const int nC = L.size;
      const float init = (float)std::sqrt(6. / (L.nIn + nC));           // Tanh::_initFactor
      std::uniform_real_distribution<float> dis(-init, init);
      for (int o = 0; o < nC; ++o) { blob[L.bOff + o] = 0.f; blob[L.bOff + nC + o] = -1.f; blob[L.bOff + 2 * nC + o] = 1.f; blob[L.bOff + 3 * nC + o] = -1.f; }
      for (int w = 0; w < 4 * nC * (L.nIn + nC); ++w) blob[L.wOff + w] = dis(gen);
    } else if (L.kind == kMGU) {    // MGULayer::initialize (Layer_GRU.h:216-231): forget gate primed with LSTM_PRIME_FAC = 1, state bias 0
      const int nC = L.size;
      const float init = (float)std::sqrt(6. / (L.nIn + nC));           // Tanh::_initFactor
      std::uniform_real_distribution<float> dis(-init, init);
      for (int o = 0; o < nC; ++o) { blob[L.bOff + o] = 1.f; blob[L.bOff + nC + o] = 0.f; }
      for (int w = 0; w < 2 * nC * (L.nIn + nC); ++w) blob[L.wOff + w] = dis(gen);
    } else if (L.kind == kResidual) {
      for (int o = 0; o < L.size; ++o) { blob[L.wOff + o] = 1.f; blob[L.bOff + o] = 0.f; }
    } else if (L.kind == kParam) {   // SoftPlus::_inv(explNoise) (Functions.h:564-568, Continuous_policy.h:195-197)
      double S = c.expl_noise;
      if (S < (double)FLT_EPSILON) S = (double)FLT_EPSILON;
      const double inv = (S * S - 0.25) / S;
      for (int o = 0; o < L.size; ++o) blob[L.bOff + o] = (float)inv;
    }
  }
}

static int upload_weights(smb200_learner* h, const float* blob) {
  const NetDesc& net = h->descs.net;
  std::vector<float> im(net.imgFloats, 0.f);
  for (int l = 1; l < net.nLayers; ++l) {
    const LayerDesc& L = net.L[l];
    if (L.kind == kDenseTanh || L.kind == kDenseLinear) {
      for (int k = 0; k < L.nIn; ++k)
        for (int n = 0; n < L.size; ++n) im[L.imgW + k * L.ldp + n] = blob[L.wOff + k * L.ld + n];
      for (int n = 0; n < L.size; ++n) im[L.imgB + n] = blob[L.bOff + n];
    } else if (is_cell_layer(L.kind)) {
      const int ng = cell_gates(L.kind);
      for (int k = 0; k < L.nIn + L.size; ++k)
        for (int n = 0; n < ng * L.size; ++n) im[L.imgW + k * L.ldp + n] = blob[L.wOff + k * L.ld + n];
      for (int n = 0; n < ng * L.size; ++n) im[L.imgB + n] = blob[L.bOff + n];
    } else if (L.kind == kResidual) {
      for (int n = 0; n < L.size; ++n) { im[L.imgW + n] = blob[L.wOff + n]; im[L.imgB + n] = blob[L.bOff + n]; }
    } else if (L.kind == kParam) {
      for (int n = 0; n < L.size; ++n) im[L.imgB + n] = blob[L.bOff + n];
    }
  }
  std::vector<float> cim;
  if (h->cimg) {       // the cluster kernel's image: [common | kCL rank blocks], positions from the index maps
    cim.assign(cluster_image_floats(h->cplan), 0.f);
    const int* iA = h->cidx.data(); const int* iB = iA + net.nParams;
    for (int p = 0; p < net.nParams; ++p) {
      if (iA[p] >= 0) cim[iA[p]] = blob[p];
      if (iB[p] >= 0) cim[iB[p]] = blob[p];
    }
    SMB200_CUDA_CHECK(cudaMemcpyAsync(h->cimg, cim.data(), sizeof(float) * cim.size(), cudaMemcpyHostToDevice, h->stream));
  }
  std::vector<float> wf, wb, wv;
  if (h->wimgF) {      // the wide step's split operand images and vector block
    wide_fill_images(net, h->wplan, h->widx, blob, wf, wb, wv);
    SMB200_CUDA_CHECK(cudaMemcpyAsync(h->wimgF, wf.data(), sizeof(float) * h->wplan.fFloats, cudaMemcpyHostToDevice, h->stream));
    SMB200_CUDA_CHECK(cudaMemcpyAsync(h->wimgB, wb.data(), sizeof(float) * h->wplan.bFloats, cudaMemcpyHostToDevice, h->stream));
    SMB200_CUDA_CHECK(cudaMemcpyAsync(h->wvec, wv.data(), sizeof(float) * h->wplan.vFloats, cudaMemcpyHostToDevice, h->stream));
  }
  SMB200_CUDA_CHECK(cudaMemcpyAsync(h->W, blob, sizeof(float) * net.nParams, cudaMemcpyHostToDevice, h->stream));
  SMB200_CUDA_CHECK(cudaMemcpyAsync(h->Wimg, im.data(), sizeof(float) * net.imgFloats, cudaMemcpyHostToDevice, h->stream));
  SMB200_CUDA_CHECK(cudaStreamSynchronize(h->stream));
  return 0;
}

static int upload_order(smb200_learner* h) {
  if (!h->orderDirty) return 0;
  std::vector<int> ord(h->episodes.size()), pos((size_t)h->rp.maxEpisodes, 0);
  for (size_t i = 0; i < ord.size(); ++i) { ord[i] = h->episodes[i].slot; pos[(size_t)ord[i]] = (int)i; }
  if (!ord.empty()) {
    SMB200_CUDA_CHECK(cudaMemcpyAsync(h->rp.epOrder, ord.data(), sizeof(int) * ord.size(), cudaMemcpyHostToDevice, h->stream));
    SMB200_CUDA_CHECK(cudaMemcpyAsync(h->rp.epPos, pos.data(), sizeof(int) * pos.size(), cudaMemcpyHostToDevice, h->stream));
  }
  SMB200_CUDA_CHECK(cudaStreamSynchronize(h->stream));   // `ord` is pageable and dies here
  h->orderDirty = false;
  return 0;
}

static int push_ctrl(smb200_learner* h) {
  h->hCtrl.adam_eta = adam_eta_for(h->cfg.learnrate, h->cfg.eps_anneal, h->hCtrl.adam_step, h->hCtrl.adam_bt1, h->hCtrl.adam_bt2);
  StepCtrl two[2] = {h->hCtrl, h->hCtrl};
  SMB200_CUDA_CHECK(cudaMemcpyAsync(h->dCtrl, two, sizeof(two), cudaMemcpyHostToDevice, h->stream));
  SMB200_CUDA_CHECK(cudaStreamSynchronize(h->stream));
  return 0;
}
static int pull_ctrl(smb200_learner* h) {
  SMB200_CUDA_CHECK(cudaMemcpyAsync(&h->hCtrl, h->dCtrl + (h->gradStep & 1), sizeof(StepCtrl), cudaMemcpyDeviceToHost, h->stream));
  SMB200_CUDA_CHECK(cudaStreamSynchronize(h->stream));
  return 0;
}

// ring allocation of `n` contiguous rows
static long long ring_alloc(smb200_learner* h, long long n) {
  const long long cap = h->rp.capRows;
  if (n > cap) return -1;
  long long cand = h->head;
  for (int pass = 0; pass < 2; ++pass) {
    while (cand + n <= cap) {
      auto it = h->liveRanges.upper_bound(cand);            // first range starting after cand
      if (it != h->liveRanges.begin()) {
        auto pv = std::prev(it);
        if (pv->second > cand) { cand = pv->second; continue; }   // cand inside a live range
      }
      if (it == h->liveRanges.end() || it->first >= cand + n) return cand;
      cand = it->second;
    }
    cand = 0;
  }
  return -1;
}

static void fill_stats(const StepCtrl& c, smb200_step_stats* o) {
  o->beta = c.beta; o->cmax = c.cmax; o->cinv = c.cinv; o->n_far_policy = c.n_far_ref; o->n_far_exact = c.n_far_exact;
  o->avg_kl = c.avg_kl; o->avg_sq_err = c.avg_sq_err; o->max_abs_err = c.max_abs_err; o->avg_return = c.avg_return;
  o->stdev_q = c.stdev_q; o->avg_q = c.avg_q; o->max_q = c.max_q; o->min_q = c.min_q;
  o->sum_ret_err = c.sum_ret_err; o->cnt_ret = c.cnt_ret; o->grad_step = c.grad_step;
}

// host mirror of the per-step bookkeeping that follows the device work of one step:
// applyEpisodesRemovalAlgo (sort by ID descending, FIFO prune) and the Adam RNG draw.
// Returns true if the episode table changed.
static bool host_post_step(smb200_learner* h) {
  bool changed = false;
  auto cmp = [](const EpisodeMeta& a, const EpisodeMeta& b) { return a.id > b.id; };
  if (h->cfg.er_filter != SMB200_FILTER_OLDEST) {
    // getERfilterAlgo (MemoryProcessing.cpp:261-298): "a goes before b", the episodes to delete end up at the back; the keys are
    // the per-episode aggregates AFTER this step's updates (fetch_keys).  The reference sorts every step with the unstable
    // std::sort — same comparator, same starting order, same libstdc++: same result, ties included.
    const float* far = h->keyAgg[0].data(); const float* kl = h->keyAgg[1].data(); const float* e2 = h->keyAgg[2].data();
    const std::vector<EpisodeMeta> before = h->episodes;
    switch (h->cfg.er_filter) {
      case SMB200_FILTER_FARPOLFRAC:
        std::sort(h->episodes.begin(), h->episodes.end(), [far](const EpisodeMeta& a, const EpisodeMeta& b) { return far[a.slot] < far[b.slot]; }); break;
      case SMB200_FILTER_MAXKLDIV:
        std::sort(h->episodes.begin(), h->episodes.end(), [kl](const EpisodeMeta& a, const EpisodeMeta& b) { return kl[a.slot] < kl[b.slot]; }); break;
      default:
        std::sort(h->episodes.begin(), h->episodes.end(), [e2](const EpisodeMeta& a, const EpisodeMeta& b) { return e2[a.slot] > e2[b.slot]; }); break;
    }
    for (size_t i = 0; i < before.size() && !changed; ++i) changed = before[i].slot != h->episodes[i].slot;
  } else
  if (!std::is_sorted(h->episodes.begin(), h->episodes.end(), cmp)) {
    std::sort(h->episodes.begin(), h->episodes.end(), cmp);
    changed = true;
  }
  while (!h->episodes.empty() && h->nTransitions - (long long)h->episodes.back().nRows > h->cfg.max_tot_obs) {
    const EpisodeMeta e = h->episodes.back();
    h->episodes.pop_back();
    h->nTransitions -= e.nRows - 1;
    h->liveRanges.erase(e.start);
    h->freeSlots.push_back(e.slot);
    h->pendingEvict.emplace_back(e.start, e.nRows);
    changed = true;
  }
  // `Saru gen(nStep, thrID, generators[thrID]())` (Optimizer.cpp:139): thread 0 consumes one
  // 32-bit draw of generators[0] — the sampler's generator — every update.
  (void)h->gen();
  h->gradStep++; h->trackerSteps++;
  if (changed) { h->orderDirty = true; h->lookupDirty = true; h->tableVersion++; h->ahead_clear(); }
  return changed;
}

// ascending LSD radix sort of n ids < 2^bits (8-bit digits).  Any correct sort reproduces the
// reference's std::sort output: the keys are plain integers.
static void radix_sort_ids(size_t* a, size_t* tmp, int n, int bits) {
  size_t* src = a; size_t* dst = tmp;
  for (int sh = 0; sh < bits; sh += 8) {
    unsigned cnt[256]; memset(cnt, 0, sizeof(cnt));
    for (int i = 0; i < n; ++i) cnt[(src[i] >> sh) & 255]++;
    unsigned sum = 0;
    for (int d = 0; d < 256; ++d) { const unsigned c = cnt[d]; cnt[d] = sum; sum += c; }
    for (int i = 0; i < n; ++i) dst[cnt[(src[i] >> sh) & 255]++] = src[i];
    std::swap(src, dst);
  }
  if (src != a) memcpy(a, src, sizeof(size_t) * n);
}

// Sampling::IDtoSeqStep walks the episodes accumulating ndata() (Sampling.cpp:26-47).  The same
// map, id -> (position, t), through prefix sums and a coarse bucket table rebuilt only when the
// episode table changes.
static void rebuild_lookup(smb200_learner* h) {
  const size_t nEp = h->episodes.size();
  h->epPrefix.resize(nEp + 1);
  long long p = 0;
  for (size_t k = 0; k < nEp; ++k) { h->epPrefix[k] = p; p += h->episodes[k].nRows - 1; }
  h->epPrefix[nEp] = p;
  int sh = 0;
  while (((p >> sh) + 1) > (1 << 16)) ++sh;            // at most 64 Ki buckets
  h->bucketShift = sh;
  const size_t nb = (size_t)(p >> sh) + 1;
  h->bucketFirst.assign(nb, 0);
  size_t k = 0;
  for (size_t b = 0; b < nb; ++b) {
    const long long first = (long long)b << sh;
    while (k + 1 < nEp && h->epPrefix[k + 1] <= first) ++k;
    h->bucketFirst[b] = (int)k;
  }
  h->lookupDirty = false;
}

// a generator that counts its 32-bit draws (same min / max / result_type: std::uniform_int_distribution behaves identically)
struct CountingGen {
  using result_type = std::mt19937::result_type;
  std::mt19937& g; unsigned n = 0;
  explicit CountingGen(std::mt19937& g_) : g(g_) {}
  static constexpr result_type min() { return std::mt19937::min(); }
  static constexpr result_type max() { return std::mt19937::max(); }
  result_type operator()() { ++n; return g(); }
};

template <class Gen>
static void host_sample_with(smb200_learner* h, Gen& gen, int* slotOut, int* tOut, int64_t* posOut, int64_t* tOut64) {
  const int B = h->cfg.batch_size;
  const long nData = (long)h->nTransitions;
  std::uniform_int_distribution<size_t> distObs(0, nData - 1);
  if ((int)h->sampIds.size() != B) { h->sampIds.resize(B); h->sampTmp.resize(B); }
  std::vector<size_t>& ret = h->sampIds;
  int bits = 1; while ((1ull << bits) < (unsigned long long)nData) ++bits;
  auto it = ret.begin();
  while (it != ret.end()) {     // Sample_uniform::sample (Sampling.cpp:82-93): draw, sort, drop duplicates, redraw the tail
    std::generate(it, ret.end(), [&]() { return distObs(gen); });
    radix_sort_ids(ret.data(), h->sampTmp.data(), B, bits);
    it = std::unique(ret.begin(), ret.end());
  }
  if (h->lookupDirty) rebuild_lookup(h);
  const long long* prefix = h->epPrefix.data();
  const int sh = h->bucketShift;
  for (int i = 0; i < B; ++i) {
    const long long id = (long long)ret[i];
    size_t k = (size_t)h->bucketFirst[(size_t)(id >> sh)];
    while (prefix[k + 1] <= id) ++k;
    const long long t = id - prefix[k];
    if (slotOut) {   // device form: ring row of (episode, t); slot with Episode::isTruncated(t+1) in bit 31
      const EpisodeMeta& e = h->episodes[k];
      const unsigned hn = ((int)t + 2 == e.nRows && !e.terminated) ? 0x80000000u : 0u;
      slotOut[i] = (int)((unsigned)e.slot | hn); tOut[i] = (int)(e.start + t);
    }
    if (posOut) { posOut[i] = (int64_t)k; tOut64[i] = (int64_t)t; }
  }
}
static void host_sample(smb200_learner* h, int* slotOut, int* tOut, int64_t* posOut, int64_t* tOut64) {
  host_sample_with(h, h->gen, slotOut, tOut, posOut, tOut64);
}

// ---- sample-ahead queue (see smb200_learner::aheadSlot) ----
constexpr int kAheadMax = 128;     // steps
// the episode table is in its steady state: the next host_post_step neither re-sorts nor evicts
static bool table_steady(const smb200_learner* h) {
  if (h->cfg.er_filter != SMB200_FILTER_OLDEST || h->episodes.empty()) return false;
  auto cmp = [](const EpisodeMeta& a, const EpisodeMeta& b) { return a.id > b.id; };
  if (!std::is_sorted(h->episodes.begin(), h->episodes.end(), cmp)) return false;
  return !(h->nTransitions - (long long)h->episodes.back().nRows > h->cfg.max_tot_obs);
}
// draw one more step into the queue; false = the queue is full, closed by a segment-ending step, or not applicable
static bool ahead_push_one(smb200_learner* h) {
  const int B = h->cfg.batch_size;
  // short calls of small mini-batches are what the queue is for; a step of a wide mini-batch takes the host longer to draw
  // than a launch costs (and the queue would hold kAheadMax * B ids)
  if (h->slow_mode() || h->nTransitions < B || B > 4096) return false;
  if (h->aheadCnt > h->aheadHead && h->aheadVersion != h->tableVersion) h->ahead_clear();
  if (h->aheadCnt == h->aheadHead) {                       // empty: start at the true position of the stream
    if (!table_steady(h)) return false;
    h->aheadHead = h->aheadCnt = 0;
    h->aheadGen = h->gen; h->aheadGradStep = h->gradStep; h->aheadVersion = h->tableVersion;
  } else if (h->aheadEnds[h->aheadCnt - 1]) return false;  // the step after a sweep / a statistics step is drawn after it ran
  if (h->aheadCnt == kAheadMax) {
    if (h->aheadHead == 0) return false;
    const int live = h->aheadCnt - h->aheadHead;           // compact
    memmove(h->aheadSlot.data(), h->aheadSlot.data() + (size_t)h->aheadHead * B, sizeof(int) * (size_t)live * B);
    memmove(h->aheadRow.data(), h->aheadRow.data() + (size_t)h->aheadHead * B, sizeof(int) * (size_t)live * B);
    memmove(h->aheadDraws.data(), h->aheadDraws.data() + h->aheadHead, sizeof(unsigned) * live);
    memmove(h->aheadEnds.data(), h->aheadEnds.data() + h->aheadHead, live);
    h->aheadHead = 0; h->aheadCnt = live;
  }
  if (h->aheadSlot.size() != (size_t)kAheadMax * B) {
    h->aheadSlot.resize((size_t)kAheadMax * B); h->aheadRow.resize((size_t)kAheadMax * B);
    h->aheadDraws.resize(kAheadMax); h->aheadEnds.resize(kAheadMax);
  }
  const int i = h->aheadCnt;
  CountingGen cg(h->aheadGen);
  host_sample_with(h, cg, h->aheadSlot.data() + (size_t)i * B, h->aheadRow.data() + (size_t)i * B, nullptr, nullptr);
  (void)cg();                                              // the Adam update's draw (host_post_step)
  const long long stepNo = h->aheadGradStep + 1;
  h->aheadDraws[i] = cg.n;
  h->aheadEnds[i] = (stepNo % 1000 == 0 || h->grad_stats_step(stepNo - 1)) ? 1 : 0;
  h->aheadGradStep = stepNo; h->aheadCnt = i + 1; h->aheadDrawn++;
  return true;
}
// serve up to n steps of a segment from the queue: ids into the pinned half, the stream position and the step counters
// advance exactly like plan_segment's host_sample + host_post_step would have moved them
static int ahead_consume(smb200_learner* h, int n, int off) {
  if (h->aheadCnt == h->aheadHead) return 0;
  if (h->aheadVersion != h->tableVersion || h->aheadGradStep - (h->aheadCnt - h->aheadHead) != h->gradStep) { h->ahead_clear(); return 0; }
  const int B = h->cfg.batch_size;
  int cnt = 0;
  unsigned long long draws = 0;
  while (cnt < n && off + cnt < h->maxSeg && h->aheadHead < h->aheadCnt) {
    const int i = h->aheadHead++;
    memcpy(h->hSampSlot + (size_t)(off + cnt) * B, h->aheadSlot.data() + (size_t)i * B, sizeof(int) * B);
    memcpy(h->hSampT + (size_t)(off + cnt) * B, h->aheadRow.data() + (size_t)i * B, sizeof(int) * B);
    draws += h->aheadDraws[i];
    h->gradStep++; h->trackerSteps++; h->aheadUsed++;
    ++cnt;
    if (h->aheadEnds[i]) break;
  }
  h->gen.discard(draws);
  return cnt;
}

// ---- prioritized samplers (ReplayMemory/Sampling.cpp:101-296) ----
// Host copies of what the samplers / filters read, after a step (or initializeLearner) has completed on the device.
static int fetch_keys(smb200_learner* h) {
  const int ds = h->cfg.data_sampling;
  if (ds == SMB200_SAMPLE_PER_RANK || ds == SMB200_SAMPLE_PER_ERR) {
    h->keyDelta.resize((size_t)h->highWater);
    if (h->highWater > 0)
      SMB200_CUDA_CHECK(cudaMemcpyAsync(h->keyDelta.data(), h->rp.DELTA, sizeof(float) * (size_t)h->highWater, cudaMemcpyDeviceToHost, h->stream));
  }
  if (ds == SMB200_SAMPLE_PER_SEQ || h->cfg.er_filter != SMB200_FILTER_OLDEST) {
    int maxSlot = 0;
    for (const auto& e : h->episodes) maxSlot = std::max(maxSlot, e.slot);
    const int which[3] = {AGG_FAR, AGG_KL, AGG_E2};
    for (int k = 0; k < 3; ++k) {
      h->keyAgg[k].resize((size_t)maxSlot + 1);
      SMB200_CUDA_CHECK(cudaMemcpyAsync(h->keyAgg[k].data(), h->rp.epAgg + (size_t)which[k] * h->rp.maxEpisodes, sizeof(float) * ((size_t)maxSlot + 1),
                                        cudaMemcpyDeviceToHost, h->stream));
    }
  }
  SMB200_CUDA_CHECK(cudaStreamSynchronize(h->stream));
  return 0;
}

// Sampling::prepare of the three prioritized samplers: float weights, the distribution itself is libstdc++'s
// std::discrete_distribution<Uint> (sequential double sums) — the class the reference instantiates.
static void prepare_sampler(smb200_learner* h) {
  const int ds = h->cfg.data_sampling;
  if (ds == SMB200_SAMPLE_UNIFORM) return;
  const float EPS = std::numeric_limits<float>::epsilon();
  const size_t nEp = h->episodes.size();
  std::vector<float> probs;
  if (ds == SMB200_SAMPLE_PER_SEQ) {                 // Sample_impSeq::prepare (:230-255)
    probs.assign(nEp, 1.f);
    for (size_t i = 0; i < nEp; ++i) {
      const EpisodeMeta& e = h->episodes[i];
      const unsigned long ndata = (unsigned long)(e.nRows - 1);
      const float P = std::sqrt(std::sqrt(h->keyAgg[2][e.slot] + EPS)) * ndata;
      probs[i] = P;
    }
  } else {
    const size_t nData = (size_t)h->nTransitions;
    probs.assign(nData, 1.f);
    std::vector<size_t> prefixes(nEp);
    size_t prefix = 0;
    for (size_t i = 0; i < nEp; ++i) { prefixes[i] = prefix; prefix += (size_t)(h->episodes[i].nRows - 1); }
    if (ds == SMB200_SAMPLE_PER_ERR) {               // TSample_impErr::prepare (:169-206)
      for (size_t i = 0; i < nEp; ++i) {
        const EpisodeMeta& e = h->episodes[i];
        const float* d = h->keyDelta.data() + e.start;
        float* probs_i = probs.data() + prefixes[i];
        for (int j = 0; j < e.nRows - 1; ++j) {
          const float deltasq = d[j] * d[j];                           // Episode::SquaredError
          probs_i[j] = std::sqrt(std::sqrt(deltasq + EPS));
        }
      }
    } else {                                         // TSample_impRank::prepare (:101-146)
      using TupEST = std::tuple<float, unsigned, unsigned>;
      std::vector<TupEST> errors(nData);
      for (size_t i = 0; i < nEp; ++i) {
        const EpisodeMeta& e = h->episodes[i];
        const float* d = h->keyDelta.data() + e.start;
        TupEST* err_i = errors.data() + prefixes[i];
        for (int j = 0; j < e.nRows - 1; ++j) err_i[j] = std::make_tuple(d[j] * d[j], (unsigned)i, (unsigned)j);
      }
      std::sort(errors.begin(), errors.end(), [](const TupEST& a, const TupEST& b) { return std::get<0>(a) > std::get<0>(b); });
      for (unsigned i = 0; i < (unsigned)nData; ++i) {
        const float P = std::get<0>(errors[i]) > 0 ? 1 / std::sqrt(std::sqrt(i + 1)) : 1;      // std::sqrt(unsigned): double
        probs[prefixes[std::get<1>(errors[i])] + std::get<2>(errors[i])] = P;
      }
    }
  }
  h->perDist = std::discrete_distribution<size_t>(probs.begin(), probs.end());
}

// TSample_impRank / TSample_impErr::sample (:148-166, :208-225) and Sample_impSeq::sample, transition branch (:276-294):
// fills the device form of the mini-batch like host_sample.
static void host_sample_per(smb200_learner* h, int* slotOut, int* tOut, int64_t* posOut, int64_t* tOut64) {
  const int B = h->cfg.batch_size;
  std::vector<std::pair<size_t, size_t>> S((size_t)B);        // (episode position, t)
  if (h->cfg.data_sampling == SMB200_SAMPLE_PER_SEQ) {
    std::uniform_real_distribution<float> distT(0, 1);
    auto it = S.begin();
    while (it != S.end()) {
      std::generate(it, S.end(), [&]() {
        const size_t _s = h->perDist(h->gen);
        const size_t _t = distT(h->gen) * (unsigned long)(h->episodes[_s].nRows - 1);
        return std::pair<size_t, size_t>{_s, _t};
      });
      std::sort(S.begin(), S.end());
      it = std::unique(S.begin(), S.end());
    }
  } else {
    std::vector<size_t> ret((size_t)B);
    auto it = ret.begin();
    while (it != ret.end()) {
      std::generate(it, ret.end(), [&]() { return h->perDist(h->gen); });
      std::sort(ret.begin(), ret.end());
      it = std::unique(ret.begin(), ret.end());
    }
    if (h->lookupDirty) rebuild_lookup(h);
    const long long* prefix = h->epPrefix.data();
    const int sh = h->bucketShift;
    for (int i = 0; i < B; ++i) {                    // Sampling::IDtoSeqStep
      const long long id = (long long)ret[i];
      size_t k = (size_t)h->bucketFirst[(size_t)(id >> sh)];
      while (prefix[k + 1] <= id) ++k;
      S[i] = {k, (size_t)(id - prefix[k])};
    }
  }
  for (int i = 0; i < B; ++i) {
    const size_t k = S[i].first; const long long t = (long long)S[i].second;
    if (slotOut) {
      const EpisodeMeta& e = h->episodes[k];
      const unsigned hn = ((int)t + 2 == e.nRows && !e.terminated) ? 0x80000000u : 0u;
      slotOut[i] = (int)((unsigned)e.slot | hn); tOut[i] = (int)(e.start + t);
    }
    if (posOut) { posOut[i] = (int64_t)k; tOut64[i] = (int64_t)t; }
  }
}

// cmax the device will hold after the statistics phase of step `gstep` (1-based)
static double cmax_at(const smb200_learner* h, long long gstep) {
  return 1.0 + h->cfg.clip_imp_weight / (1.0 + (double)gstep * h->cfg.eps_anneal);
}

// StatsTracker::advance + update + printToFile (Utils/StatsTracker.cpp:40-89) for the step whose output gradients
// g[B][nOut] were just read back: per net output the mean and the root mean square over the mini-batch (long double
// sums like the reference), appended as 2*nOut floats to <base>_outGrad_stats.raw; the file starts with the float
// nOut + 0.1 when this is the tracker's very first step.  Only learner rank 0 writes (StatsTracker.cpp:70).
static int write_grad_stats_file(const std::string& base, int B, int nOut, const float* g, bool firstTrackerStep);
static int write_grad_stats(const smb200_learner* h, const float* g, bool firstTrackerStep) {
  if (h->cfg.world_rank != 0) return 0;
  return write_grad_stats_file(h->gradStatsBase, h->cfg.batch_size, h->descs.net.nOut, g, firstTrackerStep);
}
static int write_grad_stats_file(const std::string& base, int B, int nOut, const float* g, bool firstTrackerStep) {
  std::vector<float> row(2 * (size_t)nOut);
  const long double cnt = std::max((long double)2.2e-16, (long double)B);
  for (int o = 0; o < nOut; ++o) {
    long double sum = 0, sq = 0;
    for (int b = 0; b < B; ++b) { const long double v = g[(size_t)b * nOut + o]; sum += v; sq += v * v; }
    row[o] = (float)(double)(sum / cnt);
    row[nOut + o] = (float)std::sqrt((double)(sq / cnt));
  }
  const std::string fn = base + "_outGrad_stats.raw";
  FILE* f = fopen(fn.c_str(), firstTrackerStep ? "wb" : "ab");
  if (!f) { set_error_msg(("cannot open " + fn).c_str()); return -1; }
  if (firstTrackerStep) { const float hdr = (float)(nOut + .1); fwrite(&hdr, sizeof(float), 1, f); }
  fwrite(row.data(), sizeof(float), row.size(), f);
  fclose(f);
  return 0;
}
// queue the read-back of the output gradients of the segment that just went onto the stream (its last step)
static int fetch_grad_stats(smb200_learner* h, int half) {
  SMB200_CUDA_CHECK(cudaMemcpyAsync(h->hGradStat[half], h->lastG, sizeof(float) * (size_t)h->cfg.batch_size * h->descs.net.nOut,
                                    cudaMemcpyDeviceToHost, h->stream));
  return 0;
}

// clear the live flags of the ring rows of evicted episodes (enqueued behind the kernels that may still read them)
static int flush_evictions(smb200_learner* h) {
  for (const auto& ev : h->pendingEvict) SMB200_CUDA_CHECK(cudaMemsetAsync(h->rp.rowFlag + ev.first, 0, ev.second, h->stream));
  h->pendingEvict.clear();
  return 0;
}

// device work of `n` consecutive steps whose samples sit at [first, first+n) of dSampSlot/dSampT.
// The caller guarantees the episode table is constant over them and that only the LAST one may
// be an every-1000-steps sweep step.
static int run_segment(smb200_learner* h, int first, int n, long long gstep0, int nEpPre, long long nTrPre, long long nTrPost) {
  StepArgs a = h->args();
  a.sampSlot = h->dSampSlot + (size_t)first * a.B;
  a.sampRow = h->dSampT + (size_t)first * a.B;
  a.statsOut = h->dStats + first;
  a.nEpisodes = nEpPre; a.nTransitions = nTrPre; a.nTransitionsPost = nTrPost;
  // device-side step index == absolute grad step (its parity selects the ctrl buffer)
  a.stepBase = (int)gstep0; a.lastStep = (int)(gstep0 + n - 1);
  const NetDesc& net = h->descs.net;
  const long long lastStep = gstep0 + n;              // nGradSteps()+1 of the last step
  const int sweepLast = (lastStep % 1000) == 0;
  if (h->wideOn) {
    if (launch_steps_wide(a, net, h->wplan, h->numSMs, (int)gstep0, n, sweepLast, h->stream, h->wAux, h->wEv[0], h->wEv[1], h->wEv[2])) return -2;
    h->launches += 8 * n;
  } else if (h->mode == 1 && h->clusterP1 > 0) {
    if (launch_steps_cluster(a, h->clusterP1, h->cplan.bTotal, (int)gstep0, n, sweepLast, h->stream)) return -2;
    h->launches += 1;
  } else if (h->mode == 1 && h->persistGrid > 0) {
    if (launch_steps_persistent(a, net, h->persistGrid, (int)gstep0, n, sweepLast, h->stream)) return -2;
    h->launches += 1;
  } else {
    for (int s = 0; s < n; ++s) {
      if (launch_step_two_kernels(a, net, (int)gstep0 + s, sweepLast && s == n - 1, h->stream)) return -2;
      h->launches += 2;
    }
  }
  if (sweepLast) {
    const int step = (int)(gstep0 + n - 1);
    const double cm = cmax_at(h, lastStep);
    if (launch_clear_sums(h->dSums, h->stream)) return -2;
    // retraceExplore: baseline = stats.maxAbsError BEFORE this step's update (createReturnEstimator at the top of
    // updateTrainingStatistics, MemoryProcessing.cpp:196) = ctrl[step & 1], still device-resident here
    const char* fz = getenv("SMB200_FUSED_SWEEP");
    if (sweep_fused_supported(h->rp, h->cfg.returns_estimator) && !(fz && fz[0] == '0')) {
      // Retrace + aggregates + moments in one pass over the buffer (both read the normalisers of BEFORE this sweep)
      if (launch_sweep_fused(h->rp, (int)h->episodes.size(), (float)h->cfg.gamma, (float)h->cfg.lambda, h->cfg.returns_estimator, (float)cm,
                             (float)(1.0 / cm), h->dSums, h->numSMs, h->stream)) return -2;
      h->launches -= 1;
    } else {
    if (launch_sweep(h->rp, (int)h->episodes.size(), 0, (float)h->cfg.gamma, (float)h->cfg.lambda, h->cfg.returns_estimator, 1, (float)cm, (float)(1.0 / cm),
                     h->dSums, h->stream, 0.f, h->dCtrl + (step & 1))) return -2;
    if (h->cfg.returns_estimator == SMB200_RETRACE_EXPLORE) h->launches += 1;
    if (launch_moments(h->rp, h->highWater, h->dSums, h->numSMs, h->stream)) return -2;
    }
    if (launch_peer_allreduce(h->comm, h->dSums->moments, 2 * h->cfg.dim_state + 3, ++h->vecStamp, h->stream)) return -2;
    if (launch_finalize_sweep(a, step, h->dSums, h->stream)) return -2;
    if (launch_update_scaling(h->rp, h->dCtrl + (step & 1), h->dDescs, h->dSums, 0, h->stream)) return -2;
    h->launches += 5;
  }
  return flush_evictions(h);
}

}  // namespace smb200

// =============================================================================================
extern "C" {

static int ensure_seg_capacity(smb200_learner* h, int n);
int smb200_comm_error(smb200_learner* h);

const char* smb200_last_error(void) { return g_err.c_str(); }

int smb200_default_config(smb200_config* c, int32_t dS, int32_t dA) {
  if (!c || dS < 1 || dA < 1) return SMB200_ERR_INVALID;
  memset(c, 0, sizeof(*c));
  c->algo = SMB200_VRACER; c->dim_state = dS; c->dim_action = dA;
  c->n_hidden = 2; c->hidden[0] = 128; c->hidden[1] = 128;
  c->batch_size = 256; c->batch_size_global = 256;
  c->max_tot_obs = (int64_t)(std::pow(2, 14) * std::sqrt((double)(dA + dS)));   // HyperParameters.h:53
  c->max_tot_obs_global = c->max_tot_obs;
  c->gamma = 0.995; c->lambda = 1; c->clip_imp_weight = std::sqrt(dA / 2.0); c->penal_tol = 0.1; c->eps_anneal = 5e-7;
  c->learnrate = 1e-4; c->nn_lambda = (double)FLT_EPSILON; c->expl_noise = std::sqrt(0.2); c->out_weights_prefac = 1e-3;
  c->refer_reduce_threads = 32; c->world_rank = 0; c->world_size = 1; c->seed = 42;
  c->nn_type = SMB200_FFNN; c->nn_bptt_seq = 16;
  return 0;
}

int smb200_create(const smb200_config* cfg, smb200_learner** out) {
  if (!cfg || !out) return SMB200_ERR_INVALID;
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) {
    set_error_msg("smarties_b200: no CUDA device — this library has no CPU fallback"); return SMB200_ERR_CUDA; }
  smb200_learner* h = new smb200_learner();
  h->cfg = *cfg;
  smb200_config& c = h->cfg;
  if (c.batch_size_global <= 0) c.batch_size_global = c.batch_size;
  if (c.max_tot_obs_global <= 0) c.max_tot_obs_global = c.max_tot_obs;
  if (c.refer_reduce_threads <= 0) c.refer_reduce_threads = 32;
  if (c.refer_reduce_threads > kThreads) c.refer_reduce_threads = kThreads;
  if (c.batch_size < 1 || c.max_tot_obs < c.batch_size) { set_error_msg("bad batch_size / max_tot_obs"); delete h; return SMB200_ERR_INVALID; }
  if (c.returns_estimator != SMB200_RETRACE && c.returns_estimator != SMB200_GAE && c.returns_estimator != SMB200_RETRACE_EXPLORE) {
    set_error_msg("returnsEstimator must be retrace, GAE or retraceExplore"); delete h; return SMB200_ERR_INVALID; }
  if (c.data_sampling < 0 || c.data_sampling > SMB200_SAMPLE_PER_SEQ || c.er_filter < 0 || c.er_filter > SMB200_FILTER_MINERROR) {
    set_error_msg("unknown dataSamplingAlgo / ERoldSeqFilter"); delete h; return SMB200_ERR_INVALID; }
  if ((c.data_sampling != SMB200_SAMPLE_UNIFORM || c.er_filter != SMB200_FILTER_OLDEST) && c.world_size > 1) {
    set_error_msg("prioritized samplers / non-FIFO filters: one learner rank only"); delete h; return SMB200_ERR_INVALID; }
  if (c.discrete_options != 0 && c.nn_type != SMB200_FFNN) {
    set_error_msg("discrete actions: feed-forward networks only"); delete h; return SMB200_ERR_INVALID; }
  std::vector<GradTile> tiles;
  if (build_net(c, h->descs.net, tiles)) { delete h; return SMB200_ERR_INVALID; }
  memset(&h->descs.seq, 0, sizeof(h->descs.seq));
  if (h->descs.net.recurrent) seq_plan(h->descs.net, h->descs.seq);
  Hyper& hp = h->descs.hp; memset(&hp, 0, sizeof(hp));
  hp.gamma = c.gamma; hp.lambda = c.lambda; hp.clipImpWeight = c.clip_imp_weight; hp.penalTol = c.penal_tol;
  hp.epsAnneal = c.eps_anneal; hp.learnrate = c.learnrate; hp.nnLambda = c.nn_lambda;
  hp.maxTotObsGlobal = c.max_tot_obs_global; hp.batchGlobal = c.batch_size_global; hp.batchLocal = c.batch_size;
  hp.referThreads = c.refer_reduce_threads; hp.algo = c.algo;
  for (int i = 0; i < c.dim_action; ++i) hp.bounded[i] = c.action_bounded[i];

#define CK(x) do { if ((x) != 0) { smb200_destroy(h); return SMB200_ERR_CUDA; } } while (0)
#define CKC(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { set_error(#x, e_, __FILE__, __LINE__); smb200_destroy(h); return SMB200_ERR_CUDA; } } while (0)
  CKC(cudaSetDevice(c.device));
  cudaDeviceProp prop; CKC(cudaGetDeviceProperties(&prop, c.device));
  h->numSMs = prop.multiProcessorCount;
  CKC(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  CKC(cudaEventCreate(&h->ev0)); CKC(cudaEventCreate(&h->ev1));
  CKC(cudaEventCreateWithFlags(&h->evDone[0], cudaEventDisableTiming)); CKC(cudaEventCreateWithFlags(&h->evDone[1], cudaEventDisableTiming));
  const NetDesc& net = h->descs.net;
  const int dS = c.dim_state, dA = c.dim_action, B = c.batch_size;
  long long cap = c.capacity_rows > 0 ? c.capacity_rows : c.max_tot_obs + c.max_tot_obs / 8 + 65536;
  cap = (cap + 63) / 64 * 64;
  if (cap > 0x7fffffffLL) { set_error_msg("capacity_rows must fit in 31 bits"); smb200_destroy(h); return SMB200_ERR_INVALID; }
  int maxEp = c.max_episodes > 0 ? c.max_episodes : (int)std::min<long long>(cap / 2, 1 << 20);
  ReplayView& rp = h->rp;
  rp.capRows = cap; rp.maxEpisodes = maxEp; rp.dS = dS; rp.dA = dA;
  h->dP = c.discrete_options > 0 ? c.discrete_options : 2 * dA;
  CK(dev_alloc(&rp.S, (size_t)cap * dS)); CK(dev_alloc(&rp.A, (size_t)cap * dA)); CK(dev_alloc(&rp.MU, (size_t)cap * h->dP));
  CK(dev_alloc(&rp.R, (size_t)cap)); CK(dev_alloc(&rp.V, (size_t)cap)); CK(dev_alloc(&rp.ADV, (size_t)cap));
  CK(dev_alloc(&rp.Q, (size_t)cap)); CK(dev_alloc(&rp.DELTA, (size_t)cap)); CK(dev_alloc(&rp.RHO, (size_t)cap));
  CK(dev_alloc(&rp.KL, (size_t)cap)); CK(dev_alloc(&rp.rowFlag, (size_t)cap));
  CK(dev_alloc(&rp.epStart, (size_t)maxEp)); CK(dev_alloc(&rp.epLen, (size_t)maxEp)); CK(dev_alloc(&rp.epTerm, (size_t)maxEp));
  CK(dev_alloc(&rp.epId, (size_t)maxEp)); CK(dev_alloc(&rp.epAgg, (size_t)maxEp * AGG_N)); CK(dev_alloc(&rp.epOrder, (size_t)maxEp)); CK(dev_alloc(&rp.epPos, (size_t)maxEp));
  CK(dev_alloc(&rp.stateMean, (size_t)dS)); CK(dev_alloc(&rp.stateScale, (size_t)dS)); CK(dev_alloc(&rp.stateStd, (size_t)dS));
  CK(dev_alloc(&rp.rew, (size_t)4));
  {
    std::vector<float> ones(dS, 1.f); const float rw[4] = {0.f, 1.f, 1.f, 0.f};
    CKC(cudaMemcpy(rp.stateScale, ones.data(), sizeof(float) * dS, cudaMemcpyHostToDevice));
    CKC(cudaMemcpy(rp.stateStd, ones.data(), sizeof(float) * dS, cudaMemcpyHostToDevice));
    CKC(cudaMemcpy(rp.rew, rw, sizeof(rw), cudaMemcpyHostToDevice));
  }
  h->freeSlots.reserve(maxEp);
  for (int s = maxEp - 1; s >= 0; --s) h->freeSlots.push_back(s);

  CK(dev_alloc(&h->W, (size_t)net.nParams)); CK(dev_alloc(&h->Wimg, (size_t)net.imgFloats));
  CK(dev_alloc(&h->M1, (size_t)net.nParams)); CK(dev_alloc(&h->M2, (size_t)net.nParams)); CK(dev_alloc(&h->G, (size_t)net.nParams));
  h->Bpad = net.recurrent ? round_up(B * net.Tc, 256) : round_up(B, 256);
  CK(dev_alloc(&h->actG, (size_t)net.actPerSample * h->Bpad)); CK(dev_alloc(&h->errG, (size_t)net.actPerSample * h->Bpad));
  if (net.recurrent) {      // LSTM weight gradients on the tensor cores unless SMB200_TC=0
    const char* e = getenv("SMB200_TC");
    const TcPlan tp = tc_plan(net, h->Bpad, tc_staging_bytes(net));
    if (!(e && strcmp(e, "0") == 0) && tp.nItems > 0) {
      CK(dev_alloc(&h->tcPartial, (size_t)tp.nItems * 128 * 128));
      h->useTc = 1;
    }
  }
  CK(dev_alloc(&h->dCommErr, 1));
  h->comm.error = h->dCommErr;
  h->nTiles = (int)tiles.size();
  CK(dev_alloc(&h->dTiles, tiles.size()));
  CKC(cudaMemcpy(h->dTiles, tiles.data(), sizeof(GradTile) * tiles.size(), cudaMemcpyHostToDevice));
  CK(dev_alloc(&h->dDescs, 1));
  CKC(cudaMemcpy(h->dDescs, &h->descs, sizeof(DevDescs), cudaMemcpyHostToDevice));
  CK(dev_alloc(&h->dCtrl, 2)); CK(dev_alloc(&h->dRec, (size_t)B));
  CK(dev_alloc(&h->lastO, (size_t)B * net.nOut)); CK(dev_alloc(&h->lastG, (size_t)B * net.nOut)); CK(dev_alloc(&h->lastX, (size_t)B * dS));
  CK(dev_alloc(&h->dSums, 1)); CK(dev_alloc(&h->dBarrier, 4));
  h->maxSeg = 0;
  CK(ensure_seg_capacity(h, 1));          // the default segment capacity (2048 steps up to B = 8192)

  // MemoryBuffer.h:41-44 initial ReF-ER state; AdamOptimizer beta powers (Optimizer.h:93)
  StepCtrl& k = h->hCtrl; memset(&k, 0, sizeof(k));
  k.beta = c.clip_imp_weight <= 0 ? 1.0 : 1e-4;
  k.cmax = 1.0 + c.clip_imp_weight; k.cinv = 1.0 / c.clip_imp_weight;
  k.adam_bt1 = 0.9; k.adam_bt2 = 0.999;
  CK(push_ctrl(h));

  // cluster step kernel: plan, co-resident clusters, image / partial-gradient buffers (before the first weight upload)
  {
    const char* m0 = getenv("SMB200_MODE");
    int coop0 = 0; cudaDeviceGetAttribute(&coop0, cudaDevAttrCooperativeLaunch, c.device);
    cluster_plan_build(net, 4 * 33 - 1, h->cplan, h->cidx, h->citems);
    if (h->cplan.ok && coop0 && !(m0 && strcmp(m0, "two") == 0) && !(c.target_delay > 0) && !residual_widens(net)) {
      CK(cluster_prepare(h->cplan));
      const int maxC = std::min(33, cluster_max_active(h->cplan));
      if (maxC >= 2) {
        cluster_plan_build(net, kCL * maxC - 1, h->cplan, h->cidx, h->citems);      // the P2 partition depends on the worker count
        if (h->cplan.ok) h->clusterP1 = maxC - 1;
      }
    }
    if (getenv("SMB200_DEBUG"))
      fprintf(stderr, "smb200: cluster kernel %s: %d P1 clusters, %d B of dynamic shared memory, image %zu floats, P2 chunk %d x %d parts\n",
              h->clusterP1 > 0 ? "on" : "off", h->clusterP1, h->cplan.bTotal, h->cplan.ok ? cluster_image_floats(h->cplan) : (size_t)0,
              h->cplan.chunk, h->cplan.parts);
    if (h->clusterP1 > 0) {
      CK(dev_alloc(&h->dCplan, 1));
      CKC(cudaMemcpy(h->dCplan, &h->cplan, sizeof(ClusterPlan), cudaMemcpyHostToDevice));
      CK(dev_alloc(&h->dCidx, h->cidx.size()));
      CKC(cudaMemcpy(h->dCidx, h->cidx.data(), sizeof(int) * h->cidx.size(), cudaMemcpyHostToDevice));
      CK(dev_alloc(&h->dCitems, h->citems.size()));
      CKC(cudaMemcpy(h->dCitems, h->citems.data(), sizeof(int) * h->citems.size(), cudaMemcpyHostToDevice));
      CK(dev_alloc(&h->cimg, cluster_image_floats(h->cplan)));
      CK(dev_alloc(&h->cpart, (size_t)h->clusterP1 * net.nParams));
    }
  }
  // wide step (tensor cores): batches of >= 2048 sampled transitions per rank unless SMB200_WIDE says otherwise
  // (SMB200_WIDE=0: never, =1: whenever the plan covers the network).  Measured cross-over on B200 (cfg2 network): the
  // persistent tile kernel takes 39 us at B = 1024 and 137 us at B = 4096, the wide step 57 us and 62 us.
  {
    const char* w = getenv("SMB200_WIDE");
    const bool never = w && strcmp(w, "0") == 0, always = w && strcmp(w, "1") == 0;
    if (!never && (always || B >= 2048) && !(c.target_delay > 0) && !residual_widens(net)) {
      wide_plan_build(net, hp, h->wplan, h->widx);
      if (h->wplan.ok) {
        CK(wide_prepare(h->wplan, net));
        h->wGridG = wide_grid_g(h->wplan, B, h->numSMs);
        CK(dev_alloc(&h->dWplan, 1));
        CKC(cudaMemcpy(h->dWplan, &h->wplan, sizeof(WidePlan), cudaMemcpyHostToDevice));
        CK(dev_alloc(&h->dWidx, h->widx.size()));
        CKC(cudaMemcpy(h->dWidx, h->widx.data(), sizeof(int) * h->widx.size(), cudaMemcpyHostToDevice));
        CK(dev_alloc(&h->wimgF, (size_t)h->wplan.fFloats)); CK(dev_alloc(&h->wimgB, (size_t)h->wplan.bFloats));
        CK(dev_alloc(&h->wvec, (size_t)h->wplan.vFloats));
        CK(dev_alloc(&h->wpart, (size_t)h->wGridG * h->wplan.recFloats));
        CK(dev_alloc(&h->wcnt, 4)); CK(dev_alloc(&h->wlist, (size_t)B));
        CKC(cudaStreamCreateWithFlags(&h->wAux, cudaStreamNonBlocking));
        for (int i = 0; i < 3; ++i) CKC(cudaEventCreateWithFlags(&h->wEv[i], cudaEventDisableTiming));
        h->wideOn = 1;
      }
      if (getenv("SMB200_DEBUG"))
        fprintf(stderr, "smb200: wide step %s: %d dense layers, shared memory fwd %d / bwd %d / wgrad %d B (%d stages of %d B), record %d floats x %d CTAs\n",
                h->wideOn ? "on" : "off", h->wplan.nD, h->wplan.sfTotal, h->wplan.sbTotal, h->wplan.sgTotal, h->wplan.sgStages, h->wplan.sgStageBytes,
                h->wplan.recFloats, h->wGridG);
    }
  }
  h->comm.world = 1; h->comm.rank = 0;
  h->gen.seed((unsigned long)c.seed);
  std::vector<float> blob;
  init_weights(c, net, h->gen, blob);
  CK(upload_weights(h, blob.data()));
  h->tgtBlob = blob;
  if (c.target_delay > 0) {      // the Adam epilogue of the tile kernel maintains them (the wide / cluster kernels are not used then)
    CK(dev_alloc(&h->Wtgt, (size_t)net.nParams));
    CKC(cudaMemcpy(h->Wtgt, blob.data(), sizeof(float) * (size_t)net.nParams, cudaMemcpyHostToDevice));
  }
  CK(step_kernels_prepare(net));
  { const char* t = getenv("SMB200_TMA"); h->useTma = (t && strcmp(t, "0") == 0) ? 0 : 1; }
  { const char* t = getenv("SMB200_STATS_FULL"); h->statsIncremental = (t && strcmp(t, "1") == 0) ? 0 : 1; }
  const char* m = getenv("SMB200_MODE");
  h->mode = (m && strcmp(m, "two") == 0) ? 0 : 1;
  int coop = 0; cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, c.device);
  if (!coop) h->mode = 0;
  { StepArgs a = h->args(); h->persistGrid = persistent_grid(a, net, h->numSMs); }
  if (h->persistGrid < 1) h->mode = 0;
#undef CK
#undef CKC
  *out = h;
  return 0;
}

void smb200_destroy(smb200_learner* h) {
  if (!h) return;
  cudaSetDevice(h->cfg.device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  ReplayView& rp = h->rp;
  void* ptrs[] = {rp.S, rp.A, rp.MU, rp.R, rp.V, rp.ADV, rp.Q, rp.DELTA, rp.RHO, rp.KL, rp.rowFlag, rp.epStart, rp.epLen, rp.epTerm,
                  rp.epId, rp.epAgg, rp.epOrder, rp.epPos, rp.stateMean, rp.stateScale, rp.stateStd, rp.rew, h->W, h->Wimg, h->dDbg, h->M1, h->M2, h->G,
                  h->actG, h->errG, h->dTiles, h->dDescs, h->dCtrl, h->dRec, h->lastO, h->lastG, h->lastX, h->dSums, h->dBarrier,
                  h->dSampSlot, h->dSampT, h->dStats, h->dCplan, h->dCidx, h->dCitems, h->cimg, h->cpart,
                  h->dWplan, h->dWidx, h->wimgF, h->wimgB, h->wvec, h->wpart, h->wcnt, h->wlist};
  for (void* p : ptrs) if (p) cudaFree(p);
  for (int q = 0; q < kMaxWorld; ++q) if (h->peerMapped[q]) cudaIpcCloseMemHandle(h->peerMapped[q]);
  if (h->commBuf) cudaFree(h->commBuf);
  if (h->dCommErr) cudaFree(h->dCommErr);
  if (h->tcPartial) cudaFree(h->tcPartial);
  if (h->Wtgt) cudaFree(h->Wtgt);
  if (h->fwdIn) cudaFree(h->fwdIn);
  if (h->fwdOut) cudaFree(h->fwdOut);
  if (h->fwdLen) cudaFree(h->fwdLen);
  if (h->hFwdIn) cudaFreeHost(h->hFwdIn);
  if (h->hFwdOut) cudaFreeHost(h->hFwdOut);
  if (h->hFwdLen) cudaFreeHost(h->hFwdLen);
  if (h->hSampSlot) cudaFreeHost(h->hSampSlot);
  if (h->hSampT) cudaFreeHost(h->hSampT);
  if (h->hStats) cudaFreeHost(h->hStats);
  for (float* p : h->hGradStat) if (p) cudaFreeHost(p);
  if (h->ev0) cudaEventDestroy(h->ev0);
  if (h->ev1) cudaEventDestroy(h->ev1);
  for (int i = 0; i < 2; ++i) if (h->evDone[i]) cudaEventDestroy(h->evDone[i]);
  for (int i = 0; i < 3; ++i) if (h->wEv[i]) cudaEventDestroy(h->wEv[i]);
  if (h->wAux) cudaStreamDestroy(h->wAux);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
}

int64_t smb200_n_params(const smb200_learner* h) { return h ? h->descs.net.nParams : -1; }
int32_t smb200_n_outputs(const smb200_learner* h) { return h ? h->descs.net.nOut : -1; }
int64_t smb200_n_transitions(const smb200_learner* h) { return h ? h->nTransitions : -1; }
int64_t smb200_n_episodes(const smb200_learner* h) { return h ? (int64_t)h->episodes.size() : -1; }
int64_t smb200_n_rows(const smb200_learner* h) {
  if (!h) return -1;
  int64_t n = 0; for (const auto& e : h->episodes) n += e.nRows; return n;
}

int smb200_set_weights(smb200_learner* h, const float* blob, int64_t n) {
  if (!h || !blob || n != h->descs.net.nParams) return SMB200_ERR_INVALID;
  std::lock_guard<std::recursive_mutex> lock(h->apiMutex);
  cudaSetDevice(h->cfg.device);
  // target weights (unused by RACER with targetDelay 0, but part of the checkpoint): they follow the weights
  // until training starts, like `target_weights->copy(weights)` of a restart without a tgt file (Optimizer.cpp:207-210)
  if (h->gradStep == 0) {
    h->tgtBlob.assign(blob, blob + n);
    if (h->Wtgt) SMB200_CUDA_CHECK(cudaMemcpyAsync(h->Wtgt, blob, sizeof(float) * (size_t)n, cudaMemcpyHostToDevice, h->stream));
  }
  return upload_weights(h, blob);
}
static int d2h(smb200_learner* h, void* dst, const void* src, size_t bytes) {
  SMB200_CUDA_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, h->stream));
  SMB200_CUDA_CHECK(cudaStreamSynchronize(h->stream));
  return 0;
}
int smb200_get_weights(smb200_learner* h, float* blob, int64_t n) {
  if (!h || !blob || n != h->descs.net.nParams) return SMB200_ERR_INVALID;
  std::lock_guard<std::recursive_mutex> lock(h->apiMutex);
  return d2h(h, blob, h->W, sizeof(float) * n);
}
int smb200_get_target_weights(smb200_learner* h, float* blob, int64_t n) {
  if (!h || !blob || n != h->descs.net.nParams) return SMB200_ERR_INVALID;
  std::lock_guard<std::recursive_mutex> lock(h->apiMutex);
  if (h->Wtgt) return d2h(h, blob, h->Wtgt, sizeof(float) * n);
  std::copy(h->tgtBlob.begin(), h->tgtBlob.end(), blob);
  return 0;
}
int smb200_set_target_weights(smb200_learner* h, const float* blob, int64_t n) {
  if (!h || !blob || n != h->descs.net.nParams) return SMB200_ERR_INVALID;
  std::lock_guard<std::recursive_mutex> lock(h->apiMutex);
  cudaSetDevice(h->cfg.device);
  h->tgtBlob.assign(blob, blob + n);
  if (h->Wtgt) {
    SMB200_CUDA_CHECK(cudaMemcpyAsync(h->Wtgt, blob, sizeof(float) * (size_t)n, cudaMemcpyHostToDevice, h->stream));
    SMB200_CUDA_CHECK(cudaStreamSynchronize(h->stream));
  }
  return 0;
}
int smb200_get_grad(smb200_learner* h, float* blob, int64_t n) {
  if (!h || !blob || n != h->descs.net.nParams) return SMB200_ERR_INVALID;
  return d2h(h, blob, h->G, sizeof(float) * n);
}
int smb200_get_adam(smb200_learner* h, float* m1, float* m2, int64_t n) {
  if (!h || n != h->descs.net.nParams) return SMB200_ERR_INVALID;
  if (m1 && d2h(h, m1, h->M1, sizeof(float) * n)) return SMB200_ERR_CUDA;
  if (m2 && d2h(h, m2, h->M2, sizeof(float) * n)) return SMB200_ERR_CUDA;
  return 0;
}
int smb200_set_adam(smb200_learner* h, const float* m1, const float* m2, int64_t n, int64_t n_step) {
  if (!h || n != h->descs.net.nParams) return SMB200_ERR_INVALID;
  if (m1) SMB200_CUDA_CHECK(cudaMemcpyAsync(h->M1, m1, sizeof(float) * n, cudaMemcpyHostToDevice, h->stream));
  if (m2) SMB200_CUDA_CHECK(cudaMemcpyAsync(h->M2, m2, sizeof(float) * n, cudaMemcpyHostToDevice, h->stream));
  SMB200_CUDA_CHECK(cudaStreamSynchronize(h->stream));
  if (pull_ctrl(h)) return SMB200_ERR_CUDA;
  h->hCtrl.adam_step = n_step; h->tgtPhase = n_step;
  return push_ctrl(h);
}

int smb200_set_scaling(smb200_learner* h, const float* mean, const float* scale, const float* stdev, const float rewards[3]) {
  if (!h) return SMB200_ERR_INVALID;
  const size_t b = sizeof(float) * h->cfg.dim_state;
  if (mean) SMB200_CUDA_CHECK(cudaMemcpyAsync(h->rp.stateMean, mean, b, cudaMemcpyHostToDevice, h->stream));
  if (scale) SMB200_CUDA_CHECK(cudaMemcpyAsync(h->rp.stateScale, scale, b, cudaMemcpyHostToDevice, h->stream));
  if (stdev) SMB200_CUDA_CHECK(cudaMemcpyAsync(h->rp.stateStd, stdev, b, cudaMemcpyHostToDevice, h->stream));
  if (rewards) SMB200_CUDA_CHECK(cudaMemcpyAsync(h->rp.rew, rewards, sizeof(float) * 3, cudaMemcpyHostToDevice, h->stream));
  SMB200_CUDA_CHECK(cudaStreamSynchronize(h->stream));
  return 0;
}
int smb200_get_scaling(smb200_learner* h, float* mean, float* scale, float* stdev, float rewards[3]) {
  if (!h) return SMB200_ERR_INVALID;
  const size_t b = sizeof(float) * h->cfg.dim_state;
  if (mean && d2h(h, mean, h->rp.stateMean, b)) return SMB200_ERR_CUDA;
  if (scale && d2h(h, scale, h->rp.stateScale, b)) return SMB200_ERR_CUDA;
  if (stdev && d2h(h, stdev, h->rp.stateStd, b)) return SMB200_ERR_CUDA;
  if (rewards && d2h(h, rewards, h->rp.rew, sizeof(float) * 3)) return SMB200_ERR_CUDA;
  return 0;
}

// `restored` (may be null): {Q, DELTA, RHO, KL} of a checkpointed episode (Episode::unpackEpisode, Episode.cpp:95-128):
// host bookkeeping of MemoryBuffer::pushBackEpisode (MemoryBuffer.cpp:479-520): the episode goes to the back of the
// reference's `episodes` vector, its ring range becomes live, counters advance
static void host_add_episode(smb200_learner* h, int64_t id, int32_t N, int32_t terminated, int slot, long long start) {
  h->episodes.push_back(EpisodeMeta{id, N, slot, start, terminated ? 1 : 0});
  h->liveRanges[start] = start + N;
  h->head = start + N; h->highWater = std::max(h->highWater, start + N);
  h->nTransitions += N - 1;
  h->nSeenEps += 1; h->nSeenObs += N - 1;
  h->orderDirty = true; h->lookupDirty = true; h->presampled = 0;
  h->tableVersion++; h->ahead_clear();
}

// copied as they are; the aggregates are recomputed with (cmax, cinv) like MemoryBuffer::restart does
// (Episode::updateCumulative, MemoryBuffer.cpp:257) and the return estimate is NOT re-evaluated.
static int push_episode_impl(smb200_learner* h, int64_t id, int32_t N, int32_t terminated, const float* S, const float* A,
                             const float* MU, const float* R, const float* V, const float* ADV, const float* const* restored,
                             float cmax, float cinv) {
  if (!h || N < 2 || !S || !A || !MU || !R) { set_error_msg("push_episode: an episode needs at least s0 and sT"); return SMB200_ERR_INVALID; }
  std::lock_guard<std::recursive_mutex> lock(h->apiMutex);
  cudaSetDevice(h->cfg.device);
  if (h->freeSlots.empty()) { set_error_msg("episode table full"); return SMB200_ERR_CAPACITY; }
  const long long start = ring_alloc(h, N);
  if (start < 0) { set_error_msg("replay ring full"); return SMB200_ERR_CAPACITY; }
  const int slot = h->freeSlots.back(); h->freeSlots.pop_back();
  const int dS = h->cfg.dim_state, dA = h->cfg.dim_action;
  ReplayView& rp = h->rp;
  const int dP = h->dP;
  cudaStream_t st = h->stream;
  SMB200_CUDA_CHECK(cudaMemcpyAsync(rp.S + (size_t)start * dS, S, sizeof(float) * (size_t)N * dS, cudaMemcpyHostToDevice, st));
  SMB200_CUDA_CHECK(cudaMemcpyAsync(rp.A + (size_t)start * dA, A, sizeof(float) * (size_t)N * dA, cudaMemcpyHostToDevice, st));
  SMB200_CUDA_CHECK(cudaMemcpyAsync(rp.MU + (size_t)start * dP, MU, sizeof(float) * (size_t)N * dP, cudaMemcpyHostToDevice, st));
  SMB200_CUDA_CHECK(cudaMemcpyAsync(rp.R + start, R, sizeof(float) * (size_t)N, cudaMemcpyHostToDevice, st));
  if (V) SMB200_CUDA_CHECK(cudaMemcpyAsync(rp.V + start, V, sizeof(float) * (size_t)N, cudaMemcpyHostToDevice, st));
  if (ADV) SMB200_CUDA_CHECK(cudaMemcpyAsync(rp.ADV + start, ADV, sizeof(float) * (size_t)N, cudaMemcpyHostToDevice, st));
  else if (V) SMB200_CUDA_CHECK(cudaMemsetAsync(rp.ADV + start, 0, sizeof(float) * (size_t)N, st));
  // last row: no action / policy (MemoryBuffer.cpp:124-130); first reward is 0 (Episode.cpp:239)
  SMB200_CUDA_CHECK(cudaMemsetAsync(rp.A + (size_t)(start + N - 1) * dA, 0, sizeof(float) * dA, st));
  SMB200_CUDA_CHECK(cudaMemsetAsync(rp.MU + (size_t)(start + N - 1) * dP, 0, sizeof(float) * dP, st));
  SMB200_CUDA_CHECK(cudaMemsetAsync(rp.R + start, 0, sizeof(float), st));
  const int meta[3] = {(int)start, N, terminated ? 1 : 0};
  SMB200_CUDA_CHECK(cudaMemcpyAsync(rp.epStart + slot, &meta[0], sizeof(int), cudaMemcpyHostToDevice, st));
  SMB200_CUDA_CHECK(cudaMemcpyAsync(rp.epLen + slot, &meta[1], sizeof(int), cudaMemcpyHostToDevice, st));
  SMB200_CUDA_CHECK(cudaMemcpyAsync(rp.epTerm + slot, &meta[2], sizeof(int), cudaMemcpyHostToDevice, st));
  const long long id64 = id;
  SMB200_CUDA_CHECK(cudaMemcpyAsync(rp.epId + slot, &id64, sizeof(long long), cudaMemcpyHostToDevice, st));
  // pre-training TD-error placeholder sqrt(max(EPS, stats.avgSquaredErr)) (MemoryBuffer.cpp:487)
  if (h->initialized && pull_ctrl(h)) return SMB200_ERR_CUDA;
  const float deltaInit = (float)std::sqrt(std::max((double)FLT_EPSILON, h->hCtrl.avg_sq_err));
  if (launch_init_episode(rp, slot, deltaInit, V != nullptr, st)) return SMB200_ERR_CUDA;
  if (restored) {
    float* dst[4] = {rp.Q, rp.DELTA, rp.RHO, rp.KL};
    for (int k = 0; k < 4; ++k)
      SMB200_CUDA_CHECK(cudaMemcpyAsync(dst[k] + start, restored[k], sizeof(float) * (size_t)N, cudaMemcpyHostToDevice, st));
    if (launch_sweep(rp, 0, slot, (float)h->cfg.gamma, (float)h->cfg.lambda, h->cfg.returns_estimator, 2, cmax, cinv, nullptr, st)) return SMB200_ERR_CUDA;
  } else {
    // computeReturnEstimator at insertion (MemoryBuffer.cpp:143)
    if (launch_sweep(rp, 0, slot, (float)h->cfg.gamma, (float)h->cfg.lambda, h->cfg.returns_estimator, 0, 0.f, 0.f, nullptr, st,
                     (float)h->hCtrl.max_abs_err)) return SMB200_ERR_CUDA;
  }
  SMB200_CUDA_CHECK(cudaStreamSynchronize(st));   // host buffers are the caller's: finish the copies
  host_add_episode(h, id, N, terminated, slot, start);
  return 0;
}

int smb200_push_episode(smb200_learner* h, int64_t id, int32_t N, int32_t terminated, const float* S, const float* A,
                        const float* MU, const float* R, const float* V, const float* ADV) {
  return push_episode_impl(h, id, N, terminated, S, A, MU, R, V, ADV, nullptr, 0.f, 0.f);
}

int smb200_initialize_learner(smb200_learner* h) {
  if (!h || h->episodes.empty()) return SMB200_ERR_STATE;
  std::lock_guard<std::recursive_mutex> lock(h->apiMutex);
  cudaSetDevice(h->cfg.device);
  if (h->gradStep > 0) return 0;   // "Skipping initialization for restarted learner" (Learner.cpp:51-54)
  // updateCounters(bInit=true): beta fixed-point step with the initial far-policy fraction; with
  // several learner ranks the counters are the sums over ranks (globalCounterRdx.get(bInit=true))
  StepCtrl& k = h->hCtrl;
  double cnts[2] = {(double)k.n_far_ref, (double)h->nTransitions};
  if (h->comm.world > 1) {
    double* dv = h->dSums->moments;
    SMB200_CUDA_CHECK(cudaMemcpyAsync(dv, cnts, sizeof(cnts), cudaMemcpyHostToDevice, h->stream));
    if (launch_peer_allreduce(h->comm, dv, 2, ++h->vecStamp, h->stream)) return SMB200_ERR_CUDA;
    if (d2h(h, cnts, dv, sizeof(cnts))) return SMB200_ERR_CUDA;
  }
  const double nData = cnts[1];
  const double lr = 0.1 * (double)h->cfg.batch_size_global / std::max((double)h->cfg.max_tot_obs_global, nData);
  const double frac = cnts[0] / std::max(nData, 1.0);
  const double mn = std::min(lr, k.beta);
  k.beta = frac > h->cfg.penal_tol ? (1 - mn) * k.beta : (1 - mn) * k.beta + std::min(lr, 1 - k.beta);
  k.gl_far_prev = cnts[0]; k.gl_stored_prev = cnts[1]; k.cnt_seed_step = h->gradStep;
  if (push_ctrl(h)) return SMB200_ERR_CUDA;
  if (upload_order(h)) return SMB200_ERR_CUDA;
  // updateRewardsStats(bInit=true) (moments summed over ranks: StateRewRdx), then rescaleAllReturnEstimator
  if (launch_clear_sums(h->dSums, h->stream)) return SMB200_ERR_CUDA;
  if (launch_moments(h->rp, h->highWater, h->dSums, h->numSMs, h->stream)) return SMB200_ERR_CUDA;
  if (launch_peer_allreduce(h->comm, h->dSums->moments, 2 * h->cfg.dim_state + 3, ++h->vecStamp, h->stream)) return SMB200_ERR_CUDA;
  if (launch_update_scaling(h->rp, h->dCtrl, h->dDescs, h->dSums, 1, h->stream)) return SMB200_ERR_CUDA;
  if (launch_sweep(h->rp, (int)h->episodes.size(), 0, (float)h->cfg.gamma, (float)h->cfg.lambda, h->cfg.returns_estimator, 0, 0.f, 0.f, nullptr, h->stream,
                   (float)h->hCtrl.max_abs_err))
    return SMB200_ERR_CUDA;
  SMB200_CUDA_CHECK(cudaStreamSynchronize(h->stream));
  h->initialized = true;
  h->nGatheredB4Startup = h->cfg.min_tot_obs > 0 ? h->cfg.min_tot_obs : h->cfg.max_tot_obs;
  if (h->slow_mode()) {                  // data->updateSampler() (Learner.cpp:63)
    if (fetch_keys(h)) return SMB200_ERR_CUDA;
    prepare_sampler(h);
  }
  return 0;
}

int smb200_set_grad_step(smb200_learner* h, int64_t n) {
  if (!h || n < 0) return SMB200_ERR_INVALID;
  if (pull_ctrl(h)) return SMB200_ERR_CUDA;
  h->gradStep = n; h->hCtrl.grad_step = n; h->hCtrl.adam_step = n; h->tgtPhase = n; h->ahead_clear();
  h->hCtrl.gl_far_prev = (double)h->hCtrl.n_far_ref; h->hCtrl.gl_stored_prev = (double)h->nTransitions; h->hCtrl.cnt_seed_step = n;
  return push_ctrl(h);
}
int smb200_seed_sampler(smb200_learner* h, uint64_t seed) {
  if (!h) return SMB200_ERR_INVALID;
  std::lock_guard<std::recursive_mutex> lock(h->apiMutex);
  h->gen.seed((unsigned long)seed); h->presampled = 0; h->ahead_clear();
  return 0;
}

int smb200_set_grad_stats(smb200_learner* h, const char* base) {
  if (!h) return SMB200_ERR_INVALID;
  h->gradStatsBase = base ? base : "";
  h->ahead_clear();                     // which steps end a segment depends on it
  if (!h->gradStatsBase.empty() && !h->hGradStat[0]) {
    cudaSetDevice(h->cfg.device);
    const size_t bytes = sizeof(float) * (size_t)h->cfg.batch_size * h->descs.net.nOut;
    for (float*& p : h->hGradStat) SMB200_CUDA_CHECK(cudaMallocHost(&p, bytes));
  }
  return 0;
}

int smb200_sample(smb200_learner* h, int64_t* pos, int64_t* t) {
  if (!h || !pos || !t || h->nTransitions < h->cfg.batch_size) return SMB200_ERR_STATE;
  h->ahead_clear();
  if (h->cfg.data_sampling != SMB200_SAMPLE_UNIFORM) {
    if (h->samplerStale) { if (fetch_keys(h)) return SMB200_ERR_CUDA; prepare_sampler(h); h->samplerStale = false; }
    host_sample_per(h, nullptr, nullptr, pos, t);
  } else host_sample(h, nullptr, nullptr, pos, t);
  return 0;
}

// (re)allocate the per-step sample / statistics buffers for segments of up to `n` steps
static int ensure_seg_capacity(smb200_learner* h, int n) {
  if (n <= h->maxSeg && h->dSampSlot) return 0;
  const int B = h->cfg.batch_size;
  if (h->stream) cudaStreamSynchronize(h->stream);
  if (h->dSampSlot) cudaFree(h->dSampSlot);
  if (h->dSampT) cudaFree(h->dSampT);
  if (h->dStats) cudaFree(h->dStats);
  if (h->hSampSlot) cudaFreeHost(h->hSampSlot);
  if (h->hSampT) cudaFreeHost(h->hSampT);
  if (h->hStats) cudaFreeHost(h->hStats);
  h->dSampSlot = h->dSampT = nullptr; h->dStats = nullptr; h->hSampSlot = h->hSampT = nullptr; h->hStats = nullptr;
  // 2048 steps per pair of pipeline halves, fewer for very large mini-batches (the pinned sample arrays hold maxSeg * B
  // ints each: 2 MB at B = 256; capped at 64 MB from B = 8192 on)
  h->maxSeg = std::max(n, std::min(2048, std::max(64, (1 << 24) / B)));
  if (dev_alloc(&h->dSampSlot, (size_t)h->maxSeg * B) || dev_alloc(&h->dSampT, (size_t)h->maxSeg * B) ||
      dev_alloc(&h->dStats, (size_t)h->maxSeg)) return -2;
  SMB200_CUDA_CHECK(cudaMallocHost(&h->hSampSlot, sizeof(int) * (size_t)h->maxSeg * B));
  SMB200_CUDA_CHECK(cudaMallocHost(&h->hSampT, sizeof(int) * (size_t)h->maxSeg * B));
  SMB200_CUDA_CHECK(cudaMallocHost(&h->hStats, sizeof(smb200_step_stats) * (size_t)h->maxSeg));
  h->presampled = 0;
  return 0;
}

// samples + host bookkeeping for up to `n` steps; fills pinned arrays from index 0 and returns
// how many steps can run as one device segment (constant episode table, sweep only at the end)
static int plan_segment(smb200_learner* h, int n, int off) {
  const int B = h->cfg.batch_size;
  int cnt = 0;
  while (cnt < n && off + cnt < h->maxSeg) {
    host_sample(h, h->hSampSlot + (size_t)(off + cnt) * B, h->hSampT + (size_t)(off + cnt) * B, nullptr, nullptr);
    const long long stepNo = h->gradStep + 1;
    const bool changed = host_post_step(h);
    ++cnt;
    // a step whose output gradients go to the statistics file ends its segment: smb200's diagnostics
    // buffer (lastG) holds the last step of a launch
    if (changed || stepNo % 1000 == 0 || h->grad_stats_step(stepNo - 1)) break;
  }
  return cnt;
}

static int upload_samples(smb200_learner* h, int off, int cnt) {
  const size_t bytes = sizeof(int) * (size_t)cnt * h->cfg.batch_size, o = (size_t)off * h->cfg.batch_size;
  SMB200_CUDA_CHECK(cudaMemcpyAsync(h->dSampSlot + o, h->hSampSlot + o, bytes, cudaMemcpyHostToDevice, h->stream));
  SMB200_CUDA_CHECK(cudaMemcpyAsync(h->dSampT + o, h->hSampT + o, bytes, cudaMemcpyHostToDevice, h->stream));
  return 0;
}

// Prioritized samplers / non-FIFO filters: the mini-batch of step k+1 and the episode order depend on the TD errors and episode
// aggregates step k wrote, so every step is its own launch followed by one host round trip (keys D2H, std::sort /
// std::discrete_distribution on the host like the reference, which re-prepares its sampler every step: MemoryProcessing.cpp:350).
static int train_steps_slow(smb200_learner* h, int32_t n, smb200_step_stats* stats) {
  const long long l0 = h->launches;
  SMB200_CUDA_CHECK(cudaEventRecord(h->ev0, h->stream));
  if (pull_ctrl(h)) return SMB200_ERR_CUDA;
  for (int i = 0; i < n; ++i) {
    if (h->samplerStale) { if (fetch_keys(h)) return SMB200_ERR_CUDA; prepare_sampler(h); h->samplerStale = false; }
    if (upload_order(h)) return SMB200_ERR_CUDA;
    const long long g0 = h->gradStep, tr0 = h->trackerSteps;
    const int nEpPre = (int)h->episodes.size(); const long long nTrPre = h->nTransitions;
    const double betaPrev = h->hCtrl.beta;
    if (h->cfg.data_sampling == SMB200_SAMPLE_UNIFORM) host_sample(h, h->hSampSlot, h->hSampT, nullptr, nullptr);
    else host_sample_per(h, h->hSampSlot, h->hSampT, nullptr, nullptr);
    if (upload_samples(h, 0, 1)) return SMB200_ERR_CUDA;
    // the device's beta update assumes no pruning (the pruned episodes are only known once the post-step keys are here)
    if (run_segment(h, 0, 1, g0, nEpPre, nTrPre, nTrPre)) return SMB200_ERR_CUDA;
    const bool gradStat = h->grad_stats_step(g0);
    if (gradStat && fetch_grad_stats(h, 0)) return SMB200_ERR_CUDA;
    if (fetch_keys(h)) return SMB200_ERR_CUDA;                       // synchronises the stream
    if (gradStat && write_grad_stats(h, h->hGradStat[0], tr0 == 0)) return SMB200_ERR_STATE;
    host_post_step(h);                                               // filter sort, pruning, Adam's RNG draw, gradStep++
    if (flush_evictions(h)) return SMB200_ERR_CUDA;
    if (pull_ctrl(h)) return SMB200_ERR_CUDA;                        // ctrl[(g0 + 1) & 1]
    if (h->nTransitions != nTrPre) {
      // updateCounters runs after applyEpisodesRemovalAlgo: beta with the post-pruning count (MemoryProcessing.cpp:73-85)
      const double farGlobal = (double)h->hCtrl.n_far_ref, nPost = (double)h->nTransitions;
      const double maxN = (double)h->cfg.max_tot_obs_global, BS = (double)h->cfg.batch_size_global;
      const double fracOff = farGlobal / std::max(nPost, 1.0);
      const double lrB = 0.1 * BS / std::max(maxN, nPost);
      const double mn = std::min(lrB, betaPrev);
      h->hCtrl.beta = fracOff > h->cfg.penal_tol ? (1.0 - mn) * betaPrev : (1.0 - mn) * betaPrev + std::min(lrB, 1.0 - betaPrev);
      h->hCtrl.gl_stored_prev = nPost;
      if (push_ctrl(h)) return SMB200_ERR_CUDA;
    }
    if (stats) { fill_stats(h->hCtrl, stats + i); }
    prepare_sampler(h);                                              // RM.updateSampler() after the removal
  }
  SMB200_CUDA_CHECK(cudaEventRecord(h->ev1, h->stream));
  SMB200_CUDA_CHECK(cudaStreamSynchronize(h->stream));
  float ms = 0; cudaEventElapsedTime(&ms, h->ev0, h->ev1);
  h->lastMs = ms; h->lastLaunches = h->launches - l0;
  return 0;
}

static int train_steps_impl(smb200_learner* h, int32_t n, smb200_step_stats* stats, float* weightsOut);
int smb200_train_steps(smb200_learner* h, int32_t n, smb200_step_stats* stats) { return train_steps_impl(h, n, stats, nullptr); }

// smb200_train_steps + the weights after the last step in the same call: the copy is enqueued behind the last launch and shares
// the call's one stream synchronisation (the actors' host network gets the new policy without a second round trip).
int smb200_train_steps_weights(smb200_learner* h, int32_t n, smb200_step_stats* stats, float* weights, int64_t n_weights) {
  if (!h || !weights || n_weights != h->descs.net.nParams) return SMB200_ERR_INVALID;
  return train_steps_impl(h, n, stats, weights);
}

// Page-lock a host buffer the library copies into / out of (the host network's parameter blob, episode staging buffers):
// asynchronous copies to pageable memory are staged and serialised by the driver.  bytes = 0: unregister.
int smb200_pin_host_buffer(void* ptr, int64_t bytes) {
  if (!ptr || bytes < 0) return SMB200_ERR_INVALID;
  if (bytes == 0) { SMB200_CUDA_CHECK(cudaHostUnregister(ptr)); return 0; }
  SMB200_CUDA_CHECK(cudaHostRegister(ptr, (size_t)bytes, cudaHostRegisterDefault));
  return 0;
}

static int train_steps_impl(smb200_learner* h, int32_t n, smb200_step_stats* stats, float* weightsOut) {
  if (!h || n < 0) return SMB200_ERR_INVALID;
  std::lock_guard<std::recursive_mutex> lock(h->apiMutex);
  if (h->nTransitions < h->cfg.batch_size) { set_error_msg("not enough transitions for one mini-batch"); return SMB200_ERR_STATE; }
  cudaSetDevice(h->cfg.device);
  h->presampled = 0;
  if (h->slow_mode()) {
    const int rc = train_steps_slow(h, n, stats);
    if (!rc && weightsOut) return smb200_get_weights(h, weightsOut, h->descs.net.nParams);
    return rc;
  }
  const long long l0 = h->launches;
  SMB200_CUDA_CHECK(cudaEventRecord(h->ev0, h->stream));
  // Two-deep pipeline: while the GPU runs one segment, the host samples the next one (the sampler
  // is a sequential std::mt19937 stream and must stay on the host to be bit-exact).
  // Each half of the pinned buffers holds one segment.  Segments grow 64 -> 256 -> 1000 steps: a short first one so
  // that the device starts early, long ones afterwards (every launch of the persistent kernel costs ~0.1 ms of
  // start-up and drain; a segment always ends at an every-1000-steps sweep or when the episode table changes).
  const int P = h->maxSeg / 2;
  int pendCnt[2] = {0, 0}, pendDst[2] = {0, 0}, pendGradStat[2] = {0, 0};   // pendGradStat: 1 = append, 2 = new file
  auto reclaim = [&](int b) -> int {
    if (!pendCnt[b]) return 0;
    SMB200_CUDA_CHECK(cudaEventSynchronize(h->evDone[b]));
    if (stats) memcpy(stats + pendDst[b], h->hStats + (size_t)b * P, sizeof(smb200_step_stats) * pendCnt[b]);
    if (pendGradStat[b] && write_grad_stats(h, h->hGradStat[b], pendGradStat[b] == 2)) return -1;
    pendCnt[b] = 0; pendGradStat[b] = 0;
    return 0;
  };
  int done = 0;
  double hostPlan = 0, hostWait = 0;      // SMB200_HOST_TIMING=1: where the host side of the pipeline spends its time
  for (int i = 0; done < n; ++i) {
    const int b = i & 1, off = b * P;
    const auto tw0 = std::chrono::steady_clock::now();
    if (reclaim(b)) return SMB200_ERR_CUDA;
    hostWait += std::chrono::duration<double>(std::chrono::steady_clock::now() - tw0).count();
    const long long g0 = h->gradStep, tr0 = h->trackerSteps;
    // the order in effect for these steps must reach the device before host_post_step re-sorts
    if (upload_order(h)) return SMB200_ERR_CUDA;
    const int nEpPre = (int)h->episodes.size(); const long long nTrPre = h->nTransitions;
    const auto tp0 = std::chrono::steady_clock::now();
    const int segLen = i == 0 ? 64 : (i == 1 ? 256 : 1000);
    const int want = std::min(std::min(n - done, P), segLen);
    int cnt = ahead_consume(h, want, off);                  // steps drawn ahead during the previous call's wait for the device
    if (cnt == 0) cnt = plan_segment(h, want, off);
    hostPlan += std::chrono::duration<double>(std::chrono::steady_clock::now() - tp0).count();
    const bool dirtyAfter = h->orderDirty;
    if (upload_samples(h, off, cnt)) return SMB200_ERR_CUDA;
    if (run_segment(h, off, cnt, g0, nEpPre, nTrPre, h->nTransitions)) return SMB200_ERR_CUDA;
    if (stats)
      SMB200_CUDA_CHECK(cudaMemcpyAsync(h->hStats + off, h->dStats + off, sizeof(smb200_step_stats) * cnt, cudaMemcpyDeviceToHost, h->stream));
    if (h->grad_stats_step(g0 + cnt - 1)) {         // plan_segment ended the segment on that step
      if (fetch_grad_stats(h, b)) return SMB200_ERR_CUDA;
      pendGradStat[b] = tr0 + cnt - 1 == 0 ? 2 : 1;
    }
    SMB200_CUDA_CHECK(cudaEventRecord(h->evDone[b], h->stream));
    pendCnt[b] = cnt; pendDst[b] = done;
    h->orderDirty = dirtyAfter;
    done += cnt;
  }
  if (weightsOut)
    SMB200_CUDA_CHECK(cudaMemcpyAsync(weightsOut, h->W, sizeof(float) * (size_t)h->descs.net.nParams, cudaMemcpyDeviceToHost, h->stream));
  SMB200_CUDA_CHECK(cudaEventRecord(h->ev1, h->stream));
  // the device is still working on the last segment(s): draw the next call's mini-batches meanwhile (never longer than the
  // device takes: the event is polled between steps)
  if (n > 0 && !getenv("SMB200_NO_SAMPLE_AHEAD"))
    while (cudaEventQuery(h->ev1) == cudaErrorNotReady && h->aheadCnt - h->aheadHead < kAheadMax / 2 && ahead_push_one(h)) { }
  if (reclaim(0) || reclaim(1)) return SMB200_ERR_CUDA;
  SMB200_CUDA_CHECK(cudaStreamSynchronize(h->stream));
  float ms = 0; cudaEventElapsedTime(&ms, h->ev0, h->ev1);
  h->lastMs = ms; h->lastLaunches = h->launches - l0;
  // a tensor-core item that never completed, or a peer that timed out in the fused gradient exchange, must not pass silently
  if ((h->useTc || h->comm.world > 1) && smb200_comm_error(h)) return SMB200_ERR_STATE;
  if (getenv("SMB200_HOST_TIMING"))
    fprintf(stderr, "smb200_train_steps(%d): device span %.3f ms, host sampling %.3f ms, host waiting for the device %.3f ms, %lld launches\n",
            n, ms, 1e3 * hostPlan, 1e3 * hostWait, (long long)h->lastLaunches);
  return 0;
}

int smb200_train_step_on(smb200_learner* h, const int64_t* pos, const int64_t* t, int32_t batch, smb200_step_stats* stats) {
  if (!h || !pos || !t || batch != h->cfg.batch_size) return SMB200_ERR_INVALID;
  std::lock_guard<std::recursive_mutex> lock(h->apiMutex);
  if (h->slow_mode()) { set_error_msg("train_step_on: uniform sampling with the FIFO filter only"); return SMB200_ERR_STATE; }
  h->ahead_clear();
  cudaSetDevice(h->cfg.device);
  h->presampled = 0;
  for (int b = 0; b < batch; ++b) {
    if (pos[b] < 0 || pos[b] >= (int64_t)h->episodes.size()) return SMB200_ERR_INVALID;
    const EpisodeMeta& e = h->episodes[pos[b]];
    if (t[b] < 0 || t[b] >= e.nRows - 1) return SMB200_ERR_INVALID;
    const unsigned hn = ((int)t[b] + 2 == e.nRows && !e.terminated) ? 0x80000000u : 0u;
    h->hSampSlot[b] = (int)((unsigned)e.slot | hn); h->hSampT[b] = (int)(e.start + t[b]);
  }
  if (upload_order(h)) return SMB200_ERR_CUDA;
  const long long g0 = h->gradStep, tr0 = h->trackerSteps;
  const int nEpPre = (int)h->episodes.size(); const long long nTrPre = h->nTransitions;
  host_post_step(h);
  const bool dirtyAfter = h->orderDirty;
  if (upload_samples(h, 0, 1)) return SMB200_ERR_CUDA;
  if (run_segment(h, 0, 1, g0, nEpPre, nTrPre, h->nTransitions)) return SMB200_ERR_CUDA;
  h->orderDirty = dirtyAfter;
  const bool gradStat = h->grad_stats_step(g0);
  if (gradStat && fetch_grad_stats(h, 0)) return SMB200_ERR_CUDA;
  if (stats) {
    SMB200_CUDA_CHECK(cudaMemcpyAsync(h->hStats, h->dStats, sizeof(smb200_step_stats), cudaMemcpyDeviceToHost, h->stream));
    SMB200_CUDA_CHECK(cudaStreamSynchronize(h->stream));
    *stats = h->hStats[0];
  } else SMB200_CUDA_CHECK(cudaStreamSynchronize(h->stream));
  if (gradStat && write_grad_stats(h, h->hGradStat[0], tr0 == 0)) return SMB200_ERR_STATE;
  if ((h->useTc || h->comm.world > 1) && smb200_comm_error(h)) return SMB200_ERR_STATE;
  return 0;
}

int smb200_presample(smb200_learner* h, int32_t n) {
  if (!h || n < 1) return SMB200_ERR_INVALID;
  if (h->slow_mode()) { set_error_msg("presample: uniform sampling with the FIFO filter only"); return SMB200_ERR_STATE; }
  h->ahead_clear();
  cudaSetDevice(h->cfg.device);
  if (ensure_seg_capacity(h, n)) return SMB200_ERR_CUDA;
  // benchmark path: the episode table must be in its steady (sorted, un-pruned) state
  const int B = h->cfg.batch_size;
  for (int i = 0; i < n; ++i) {
    host_sample(h, h->hSampSlot + (size_t)i * B, h->hSampT + (size_t)i * B, nullptr, nullptr);
    (void)h->gen();   // the Adam update's draw (host_post_step)
  }
  if (upload_samples(h, 0, n)) return SMB200_ERR_CUDA;
  SMB200_CUDA_CHECK(cudaStreamSynchronize(h->stream));
  h->presampled = n;
  return 0;
}

int smb200_train_presampled(smb200_learner* h, int32_t first, int32_t n) {
  if (!h || first < 0 || n < 1 || first + n > h->presampled) return SMB200_ERR_INVALID;
  std::lock_guard<std::recursive_mutex> lock(h->apiMutex);
  cudaSetDevice(h->cfg.device);
  auto cmp = [](const EpisodeMeta& a, const EpisodeMeta& b) { return a.id > b.id; };
  if (!std::is_sorted(h->episodes.begin(), h->episodes.end(), cmp)) { set_error_msg("presampled path needs a sorted episode table"); return SMB200_ERR_STATE; }
  if (upload_order(h)) return SMB200_ERR_CUDA;
  const long long l0 = h->launches;
  SMB200_CUDA_CHECK(cudaEventRecord(h->ev0, h->stream));
  int done = 0;
  while (done < n) {
    const long long g0 = h->gradStep;
    int cnt = n - done;
    const long long toSweep = 1000 - (g0 % 1000);     // steps until (and including) the next sweep step
    if (cnt > toSweep) cnt = (int)toSweep;
    if (run_segment(h, first + done, cnt, g0, (int)h->episodes.size(), h->nTransitions, h->nTransitions)) return SMB200_ERR_CUDA;
    h->gradStep += cnt; h->trackerSteps += cnt;
    done += cnt;
  }
  SMB200_CUDA_CHECK(cudaEventRecord(h->ev1, h->stream));
  h->lastLaunches = h->launches - l0;
  return 0;
}

// Diagnostics: run `n` presampled steps in ONE persistent launch with phase timestamps
// (clock64 of thread 0 of every CTA at 8 markers per step).  out[n][grid][24].
int smb200_profile_phases(smb200_learner* h, int32_t n, int64_t* out, int64_t capacity, int32_t* grid_out) {
  if (!h || n < 1 || n > h->presampled || !out || h->mode != 1) return SMB200_ERR_INVALID;
  const long long g0 = h->gradStep;
  if ((g0 % 1000) + n >= 1000) return SMB200_ERR_STATE;   // keep the profiled launch free of sweeps
  const int pgrid = h->clusterP1 > 0 ? (h->clusterP1 + 1) * kCL : h->persistGrid;
  const size_t cnt = (size_t)n * pgrid * 48;
  if ((int64_t)cnt > capacity) return SMB200_ERR_INVALID;
  cudaSetDevice(h->cfg.device);
  if (h->dDbg) { cudaFree(h->dDbg); h->dDbg = nullptr; }
  if (dev_alloc(&h->dDbg, cnt)) return SMB200_ERR_CUDA;
  if (upload_order(h)) return SMB200_ERR_CUDA;
  StepArgs a = h->args();
  a.statsOut = h->dStats;
  a.stepBase = (int)g0; a.lastStep = (int)(g0 + n - 1);
  a.dbgT = h->dDbg;
  SMB200_CUDA_CHECK(cudaEventRecord(h->ev0, h->stream));
  if (h->clusterP1 > 0) { if (launch_steps_cluster(a, h->clusterP1, h->cplan.bTotal, (int)g0, n, 0, h->stream)) return SMB200_ERR_CUDA; }
  else if (launch_steps_persistent(a, h->descs.net, h->persistGrid, (int)g0, n, 0, h->stream)) return SMB200_ERR_CUDA;
  SMB200_CUDA_CHECK(cudaEventRecord(h->ev1, h->stream));
  h->gradStep += n; h->trackerSteps += n;
  if (d2h(h, out, h->dDbg, sizeof(long long) * cnt)) return SMB200_ERR_CUDA;
  float ms = 0; cudaEventElapsedTime(&ms, h->ev0, h->ev1); h->lastMs = ms; h->lastLaunches = 1;
  if (grid_out) *grid_out = pgrid;
  return 0;
}

// Host half of the learner step without any device work (diagnostics for the CPU test suite): the SAME host functions the
// learner runs — ring allocator, host_add_episode, host_sample, host_post_step — on an episode table given as
// (id, rows, terminated) in push order.  Per step: sampled (episode id, t) in sample order; after the step's FIFO
// sort / pruning: the number of episodes and their ids in the reference's vector order (padded with -1).
int smb200_host_replay_trace(int32_t batch_size, int64_t max_tot_obs, int64_t capacity_rows, int32_t n_ep, const int64_t* ids,
                             const int32_t* n_rows, const int32_t* terminated, uint64_t seed, int32_t n_steps,
                             int64_t* ep_id_out, int64_t* t_out, int32_t* n_ep_after, int64_t* order_out,
                             const int32_t* push_before_step, int64_t* start_out) {
  if (batch_size < 1 || n_ep < 1 || n_steps < 0 || !ids || !n_rows || !terminated || !ep_id_out || !t_out) return SMB200_ERR_INVALID;
  smb200_learner* h = new smb200_learner();
  memset(&h->cfg, 0, sizeof(h->cfg));
  h->cfg.batch_size = batch_size; h->cfg.max_tot_obs = max_tot_obs;
  long long cap = capacity_rows > 0 ? capacity_rows : max_tot_obs + max_tot_obs / 8 + 65536;   // smb200_create's default
  h->rp.capRows = (cap + 63) / 64 * 64;
  for (int s = n_ep - 1; s >= 0; --s) h->freeSlots.push_back(s);
  int rc = 0, next = 0;
  // episodes arrive in the order given; push_before_step[e] (non-decreasing; absent = 0) is the learner step before which
  // episode e is pushed — the actors keep feeding the buffer while the learner trains
  auto push_until = [&](int step) {
    for (; next < n_ep && !rc && (!push_before_step || push_before_step[next] <= step); ++next) {
      const int e = next;
      if (n_rows[e] < 2) { set_error_msg("push_episode: an episode needs at least s0 and sT"); rc = SMB200_ERR_INVALID; break; }
      if (h->freeSlots.empty()) { set_error_msg("episode table full"); rc = SMB200_ERR_CAPACITY; break; }
      const long long start = ring_alloc(h, n_rows[e]);
      if (start < 0) { set_error_msg("replay ring full"); rc = SMB200_ERR_CAPACITY; break; }
      const int slot = h->freeSlots.back(); h->freeSlots.pop_back();
      host_add_episode(h, ids[e], n_rows[e], terminated[e], slot, start);
    }
  };
  push_until(0);
  if (!rc && h->nTransitions < batch_size) { set_error_msg("not enough transitions for one mini-batch"); rc = SMB200_ERR_STATE; }
  if (!rc) {
    h->gen.seed((unsigned long)seed);
    std::vector<int64_t> pos(batch_size);
    for (int s = 0; s < n_steps && !rc; ++s) {
      if (s > 0) push_until(s);
      if (rc) break;
      host_sample(h, nullptr, nullptr, pos.data(), t_out + (size_t)s * batch_size);
      for (int i = 0; i < batch_size; ++i) ep_id_out[(size_t)s * batch_size + i] = h->episodes[(size_t)pos[i]].id;
      host_post_step(h);
      h->pendingEvict.clear();        // the learner clears the live flags of evicted rows on the device here
      if (n_ep_after) n_ep_after[s] = (int32_t)h->episodes.size();
      for (int k = 0; k < n_ep; ++k) {
        const bool live = k < (int)h->episodes.size();
        if (order_out) order_out[(size_t)s * n_ep + k] = live ? h->episodes[k].id : -1;
        if (start_out) start_out[(size_t)s * n_ep + k] = live ? h->episodes[k].start : -1;
      }
    }
  }
  delete h;
  return rc;
}

// Network construction without a device (diagnostics for the CPU test suite): the padded parameter blob that smb200_create
// would upload — build_net's layout (Parameters.h:159-176) filled by init_weights from mt19937(cfg->seed), i.e.
// Builder::build on generators[0] of a run with randSeed = seed (Builder.cpp:133-137, ExecutionInfo.cpp:391).
// blob == nullptr: only returns the blob size.
int64_t smb200_host_init_weights(const smb200_config* cfg, float* blob, int64_t n) {
  if (!cfg) return SMB200_ERR_INVALID;
  NetDesc* net = new NetDesc();
  std::vector<GradTile> tiles;
  if (build_net(*cfg, *net, tiles)) { delete net; return SMB200_ERR_INVALID; }
  const int64_t np = net->nParams;
  if (blob) {
    if (n != np) { delete net; set_error_msg("init_weights: blob size mismatch"); return SMB200_ERR_INVALID; }
    std::mt19937 gen((unsigned long)cfg->seed);
    std::vector<float> w;
    init_weights(*cfg, *net, gen, w);
    memcpy(blob, w.data(), sizeof(float) * (size_t)np);
  }
  delete net;
  return np;
}

// The wide step's host-built plan, index maps and operand images (wide_plan_build / wide_fill_images, csrc/wide_step.cuh)
// without a device: what smb200_create builds and uploads.
int smb200_host_wide_plan(const smb200_config* cfg, int32_t* info, int32_t* dense, const float* blob, int64_t n_blob, int32_t* idx,
                          float* img_f, float* img_b, float* vec) {
  if (!cfg) return SMB200_ERR_INVALID;
  NetDesc* net = new NetDesc();
  std::vector<GradTile> tiles;
  if (build_net(*cfg, *net, tiles)) { delete net; return SMB200_ERR_INVALID; }
  Hyper hp; memset(&hp, 0, sizeof(hp)); hp.algo = cfg->algo;
  WidePlan* wp = new WidePlan();
  std::vector<int> widx;
  wide_plan_build(*net, hp, *wp, widx);
  const int ok = wp->ok;
  if (info) {
    const int v[16] = {wp->nD, wp->fFloats, wp->bFloats, wp->vFloats, wp->recFloats, wp->gCols, wp->sfTotal, wp->sbTotal, wp->sgTotal,
                       wp->sgStages, wp->NpG, 0, 0, 0, 0, 0};
    for (int i = 0; i < 16; ++i) info[i] = v[i];
  }
  if (dense && ok)
    for (int d = 0; d < wp->nD; ++d) {
      const WDense& D = wp->D[d];
      const int v[8] = {D.K, D.Kp, D.N, D.Np, D.fImg, D.bImg, D.gN, D.gPart};
      for (int i = 0; i < 8; ++i) dense[8 * d + i] = v[i];
    }
  int rc = ok;
  if (ok && blob) {
    if (n_blob != net->nParams) { set_error_msg("wide_plan: blob size mismatch"); rc = SMB200_ERR_INVALID; }
    else {
      if (idx) memcpy(idx, widx.data(), sizeof(int) * widx.size());
      std::vector<float> f, b, v;
      wide_fill_images(*net, *wp, widx, blob, f, b, v);
      if (img_f) memcpy(img_f, f.data(), sizeof(float) * (size_t)wp->fFloats);
      if (img_b) memcpy(img_b, b.data(), sizeof(float) * (size_t)wp->bFloats);
      if (vec) memcpy(vec, v.data(), sizeof(float) * (size_t)wp->vFloats);
    }
  }
  delete wp; delete net;
  return rc;
}

// The StatsTracker file writer without a device (diagnostics for the CPU test suite): g = per-sample output gradients
// [batch][n_out] of a step that starts at nGradSteps % 1000 == 0, reduced and appended exactly as the learner does.
int smb200_host_write_grad_stats(const char* base, int32_t batch, int32_t n_out, const float* g, int32_t first_tracker_step) {
  if (!base || !g || batch < 1 || n_out < 1) return SMB200_ERR_INVALID;
  return write_grad_stats_file(base, batch, n_out, g, first_tracker_step != 0) ? SMB200_ERR_STATE : 0;
}

// Host build of the inline function the statistics phase uses for the reference's `Uint += float`.
uint64_t smb200_uint_plus_float(uint64_t n, float x) { return (uint64_t)uint_plus_float_x86((unsigned long long)n, x); }

// ---- learner ranks sharing the gradient: peer-memory buffers exchanged through CUDA IPC ----
static void comm_layout(CommView& cm, int world, int rank, int nParams, int nTiles) {
  cm.world = world; cm.rank = rank;
  cm.nParamsPad = (nParams + 31) / 32 * 32; cm.nTilesPad = (nTiles + 1 + 31) / 32 * 32;
  size_t o = 0;
  cm.offGrad = o; o += sizeof(unsigned) * 4 * (size_t)world * cm.nParamsPad;        // 4-byte elements, four rotating slots
  cm.gradBytes = o - cm.offGrad;
  cm.offFlag = o; o += sizeof(unsigned) * (size_t)world * cm.nTilesPad;
  o = (o + 255) / 256 * 256;
  cm.offCnt = o; o += sizeof(unsigned long long) * 4 * (size_t)world * 4;   // [step & 3][rank][4] stamped words
  cm.offCntFlag = o; o += 256;
  cm.offVec = o; o += sizeof(double) * 2 * (size_t)world * kCommVec;
  cm.offVecFlag = o; o += 256;
  cm.bytes = o;
  cm.timeoutCycles = 6000000000LL;   // ~3 s at 2 GHz
}

int smb200_comm_init(smb200_learner* h, int32_t world, int32_t rank, uint8_t* handle_out, int32_t handle_bytes) {
  if (!h || world < 1 || world > kMaxWorld || rank < 0 || rank >= world || !handle_out || handle_bytes < (int)sizeof(cudaIpcMemHandle_t))
    return SMB200_ERR_INVALID;
  if (world != h->cfg.world_size || rank != h->cfg.world_rank) { set_error_msg("comm_init: world/rank differ from the config"); return SMB200_ERR_INVALID; }
  cudaSetDevice(h->cfg.device);
  comm_layout(h->comm, world, rank, h->descs.net.nParams, h->nTiles);
  h->comm.world = 1;   // stays single-rank until smb200_comm_attach
  if (!h->commBuf) {
    SMB200_CUDA_CHECK(cudaMalloc(&h->commBuf, h->comm.bytes));
    SMB200_CUDA_CHECK(cudaMemset(h->commBuf, 0, h->comm.bytes));
    SMB200_CUDA_CHECK(cudaMemset(h->commBuf + h->comm.offGrad, 0xFF, h->comm.gradBytes));     // "not arrived yet" (kPoison)
  }
  cudaIpcMemHandle_t hd;
  SMB200_CUDA_CHECK(cudaIpcGetMemHandle(&hd, h->commBuf));
  memset(handle_out, 0, handle_bytes);
  memcpy(handle_out, &hd, sizeof(hd));
  return 0;
}

int smb200_comm_attach(smb200_learner* h, const uint8_t* handles, int32_t handle_bytes) {
  if (!h || !h->commBuf || !handles || handle_bytes < (int)sizeof(cudaIpcMemHandle_t)) return SMB200_ERR_INVALID;
  cudaSetDevice(h->cfg.device);
  const int world = h->cfg.world_size, rank = h->cfg.world_rank;
  for (int q = 0; q < world; ++q) {
    if (q == rank) { h->comm.base[q] = h->commBuf; continue; }
    cudaIpcMemHandle_t hd; memcpy(&hd, handles + (size_t)q * handle_bytes, sizeof(hd));
    void* ptr = nullptr;
    SMB200_CUDA_CHECK(cudaIpcOpenMemHandle(&ptr, hd, cudaIpcMemLazyEnablePeerAccess));
    h->peerMapped[q] = ptr; h->comm.base[q] = reinterpret_cast<unsigned char*>(ptr);
  }
  h->comm.error = h->dCommErr;
  h->comm.world = world;
  return 0;
}

int smb200_comm_error(smb200_learner* h) {
  if (!h) return SMB200_ERR_INVALID;
  if (!h->dCommErr) return 0;
  int e = 0;
  if (d2h(h, &e, h->dCommErr, sizeof(int))) return SMB200_ERR_CUDA;
  if (e == 2) { set_error_msg("a tensor-core weight-gradient item did not complete (tcgen05 commit never arrived)"); return SMB200_ERR_STATE; }
  if (e) { set_error_msg("a peer rank did not answer within the time-out of the fused gradient exchange"); return SMB200_ERR_STATE; }
  return 0;
}

int smb200_sync(smb200_learner* h) {
  if (!h) return SMB200_ERR_INVALID;
  SMB200_CUDA_CHECK(cudaStreamSynchronize(h->stream));
  float ms = 0;
  if (cudaEventElapsedTime(&ms, h->ev0, h->ev1) == cudaSuccess) h->lastMs = ms;
  // smb200_train_presampled only enqueues: its device-side errors (peer time-out, tensor-core item) surface here
  if ((h->useTc || h->comm.world > 1) && smb200_comm_error(h)) return SMB200_ERR_STATE;
  return 0;
}

int smb200_step_kernel(const smb200_learner* h) {
  if (!h) return -1;
  if (h->wideOn) return 3;
  if (h->mode == 1 && h->clusterP1 > 0) return 2;
  return (h->mode == 1 && h->persistGrid > 0) ? 1 : 0;
}

int smb200_last_timing(smb200_learner* h, double* ms, int64_t* launches) {
  if (!h) return SMB200_ERR_INVALID;
  if (ms) *ms = h->lastMs;
  if (launches) *launches = h->lastLaunches;
  return 0;
}

int smb200_get_last_batch(smb200_learner* h, float* O, float* g, float* X) {
  if (!h) return SMB200_ERR_INVALID;
  const int B = h->cfg.batch_size, nOut = h->descs.net.nOut, dS = h->cfg.dim_state;
  if (O && d2h(h, O, h->lastO, sizeof(float) * B * nOut)) return SMB200_ERR_CUDA;
  if (g && d2h(h, g, h->lastG, sizeof(float) * B * nOut)) return SMB200_ERR_CUDA;
  if (X && d2h(h, X, h->lastX, sizeof(float) * B * dS)) return SMB200_ERR_CUDA;
  return 0;
}

int smb200_retrace_sweep(smb200_learner* h, double* sumErr2) {
  if (!h || h->episodes.empty()) return SMB200_ERR_STATE;
  cudaSetDevice(h->cfg.device);
  if (upload_order(h)) return SMB200_ERR_CUDA;
  if (launch_clear_sums(h->dSums, h->stream)) return SMB200_ERR_CUDA;
  SMB200_CUDA_CHECK(cudaEventRecord(h->ev0, h->stream));
  if (launch_sweep(h->rp, (int)h->episodes.size(), 0, (float)h->cfg.gamma, (float)h->cfg.lambda, h->cfg.returns_estimator, 0, 0.f, 0.f, h->dSums, h->stream,
                   0.f, h->dCtrl + (h->gradStep & 1)))
    return SMB200_ERR_CUDA;
  SMB200_CUDA_CHECK(cudaEventRecord(h->ev1, h->stream));
  double e = 0;
  if (d2h(h, &e, &h->dSums->sumErr2, sizeof(double))) return SMB200_ERR_CUDA;
  float ms = 0; cudaEventElapsedTime(&ms, h->ev0, h->ev1); h->lastMs = ms; h->lastLaunches = 1;
  if (sumErr2) *sumErr2 = e;
  return 0;
}

// The every-1000-steps pass as ONE kernel (k_sweep_fused): Retrace / GAE over all episodes, exact recompute of the episode
// aggregates, reward and state moments; without the statistics / normaliser update that follows it inside a learner step.
int smb200_fused_sweep(smb200_learner* h, double* sumErr2, double* moments) {
  if (!h || h->episodes.empty()) return SMB200_ERR_STATE;
  if (!sweep_fused_supported(h->rp, h->cfg.returns_estimator)) { set_error_msg("fused sweep: unsupported state width or estimator"); return SMB200_ERR_INVALID; }
  cudaSetDevice(h->cfg.device);
  const int dS = h->cfg.dim_state;
  if (upload_order(h)) return SMB200_ERR_CUDA;
  if (launch_clear_sums(h->dSums, h->stream)) return SMB200_ERR_CUDA;
  const double cm = cmax_at(h, h->gradStep);
  SMB200_CUDA_CHECK(cudaEventRecord(h->ev0, h->stream));
  if (launch_sweep_fused(h->rp, (int)h->episodes.size(), (float)h->cfg.gamma, (float)h->cfg.lambda, h->cfg.returns_estimator, (float)cm,
                         (float)(1.0 / cm), h->dSums, h->numSMs, h->stream)) return SMB200_ERR_CUDA;
  SMB200_CUDA_CHECK(cudaEventRecord(h->ev1, h->stream));
  double e = 0;
  std::vector<double> m(2 * dS + 3);
  if (d2h(h, &e, &h->dSums->sumErr2, sizeof(double))) return SMB200_ERR_CUDA;
  if (d2h(h, m.data(), h->dSums->moments, sizeof(double) * m.size())) return SMB200_ERR_CUDA;
  float ms = 0; cudaEventElapsedTime(&ms, h->ev0, h->ev1); h->lastMs = ms; h->lastLaunches = 1;
  if (sumErr2) *sumErr2 = e;
  if (moments) memcpy(moments, m.data(), sizeof(double) * m.size());
  return 0;
}

int smb200_reward_state_moments(smb200_learner* h, double* out) {
  if (!h || h->episodes.empty()) return SMB200_ERR_STATE;
  cudaSetDevice(h->cfg.device);
  const int dS = h->cfg.dim_state;
  if (launch_clear_sums(h->dSums, h->stream)) return SMB200_ERR_CUDA;
  SMB200_CUDA_CHECK(cudaEventRecord(h->ev0, h->stream));
  if (launch_moments(h->rp, h->highWater, h->dSums, h->numSMs, h->stream)) return SMB200_ERR_CUDA;
  SMB200_CUDA_CHECK(cudaEventRecord(h->ev1, h->stream));
  std::vector<double> m(2 * dS + 3);
  if (d2h(h, m.data(), h->dSums->moments, sizeof(double) * m.size())) return SMB200_ERR_CUDA;
  float ms = 0; cudaEventElapsedTime(&ms, h->ev0, h->ev1); h->lastMs = ms; h->lastLaunches = 1;
  if (out) memcpy(out, m.data(), sizeof(double) * m.size());
  return 0;
}

int smb200_read_field(smb200_learner* h, int32_t field, float* out, int64_t n) {
  if (!h || !out || n != smb200_n_rows(h)) return SMB200_ERR_INVALID;
  const float* src = nullptr;
  switch (field) {
    case SMB200_F_V: src = h->rp.V; break;       case SMB200_F_ADV: src = h->rp.ADV; break;
    case SMB200_F_QRET: src = h->rp.Q; break;    case SMB200_F_DELTA: src = h->rp.DELTA; break;
    case SMB200_F_RHO: src = h->rp.RHO; break;   case SMB200_F_KL: src = h->rp.KL; break;
    case SMB200_F_REWARD: src = h->rp.R; break;  default: return SMB200_ERR_INVALID;
  }
  std::vector<float> all(h->highWater);
  if (d2h(h, all.data(), src, sizeof(float) * h->highWater)) return SMB200_ERR_CUDA;
  int64_t o = 0;
  for (const auto& e : h->episodes) { memcpy(out + o, all.data() + e.start, sizeof(float) * e.nRows); o += e.nRows; }
  return 0;
}

int smb200_read_episodes(smb200_learner* h, int64_t* ids, int64_t* nRows, float* agg, int64_t nEp) {
  if (!h || nEp != (int64_t)h->episodes.size()) return SMB200_ERR_INVALID;
  const int ME = h->rp.maxEpisodes;
  std::vector<float> all;
  if (agg) { all.resize((size_t)ME * AGG_N); if (d2h(h, all.data(), h->rp.epAgg, sizeof(float) * all.size())) return SMB200_ERR_CUDA; }
  for (int64_t i = 0; i < nEp; ++i) {
    const EpisodeMeta& e = h->episodes[i];
    if (ids) ids[i] = e.id;
    if (nRows) nRows[i] = e.nRows;
    if (agg) for (int k = 0; k < AGG_N; ++k) agg[i * AGG_N + k] = all[(size_t)k * ME + e.slot];
  }
  return 0;
}

int smb200_get_stats(smb200_learner* h, smb200_step_stats* out) {
  if (!h || !out) return SMB200_ERR_INVALID;
  if (pull_ctrl(h)) return SMB200_ERR_CUDA;
  fill_stats(h->hCtrl, out);
  return 0;
}

// device + page-locked host staging of the actors' requests, grown geometrically and kept for the life of the learner
static int forward_buffers(smb200_learner* h, size_t nIn, size_t nOut, size_t nLen) {
  if (nIn > h->fwdInCap) {
    if (h->fwdIn) cudaFree(h->fwdIn);
    if (h->hFwdIn) cudaFreeHost(h->hFwdIn);
    h->fwdIn = nullptr; h->hFwdIn = nullptr; h->fwdInCap = 0;
    const size_t cap = std::max(nIn, (size_t)4096);
    SMB200_CUDA_CHECK(cudaMalloc(&h->fwdIn, sizeof(float) * cap));
    SMB200_CUDA_CHECK(cudaMallocHost(&h->hFwdIn, sizeof(float) * cap));
    h->fwdInCap = cap;
  }
  if (nOut > h->fwdOutCap) {
    if (h->fwdOut) cudaFree(h->fwdOut);
    if (h->hFwdOut) cudaFreeHost(h->hFwdOut);
    h->fwdOut = nullptr; h->hFwdOut = nullptr; h->fwdOutCap = 0;
    const size_t cap = std::max(nOut, (size_t)4096);
    SMB200_CUDA_CHECK(cudaMalloc(&h->fwdOut, sizeof(float) * cap));
    SMB200_CUDA_CHECK(cudaMallocHost(&h->hFwdOut, sizeof(float) * cap));
    h->fwdOutCap = cap;
  }
  if (nLen > h->fwdLenCap) {
    if (h->fwdLen) cudaFree(h->fwdLen);
    if (h->hFwdLen) cudaFreeHost(h->hFwdLen);
    h->fwdLen = nullptr; h->hFwdLen = nullptr; h->fwdLenCap = 0;
    const size_t cap = std::max(nLen, (size_t)256);
    SMB200_CUDA_CHECK(cudaMalloc(&h->fwdLen, sizeof(int) * cap));
    SMB200_CUDA_CHECK(cudaMallocHost(&h->hFwdLen, sizeof(int) * cap));
    h->fwdLenCap = cap;
  }
  return 0;
}

int smb200_forward(smb200_learner* h, const float* states, int32_t n, float* outputs) {
  if (!h || !states || !outputs || n < 1) return SMB200_ERR_INVALID;
  if (h->descs.net.recurrent) { set_error_msg("smb200_forward: stateless evaluation is undefined for a recurrent network (smb200_forward_seq takes the window)"); return SMB200_ERR_STATE; }
  std::lock_guard<std::recursive_mutex> lock(h->apiMutex);
  cudaSetDevice(h->cfg.device);
  const int dS = h->cfg.dim_state, nOut = h->descs.net.nOut;
  const size_t nIn = (size_t)n * dS, nO = (size_t)n * nOut;
  if (forward_buffers(h, nIn, nO, 0)) return SMB200_ERR_CUDA;
  std::memcpy(h->hFwdIn, states, sizeof(float) * nIn);
  SMB200_CUDA_CHECK(cudaMemcpyAsync(h->fwdIn, h->hFwdIn, sizeof(float) * nIn, cudaMemcpyHostToDevice, h->stream));
  StepArgs a = h->args();
  if (launch_forward(a, h->descs.net, h->fwdIn, n, h->fwdOut, h->stream)) return SMB200_ERR_CUDA;
  if (d2h(h, h->hFwdOut, h->fwdOut, sizeof(float) * nO)) return SMB200_ERR_CUDA;
  std::memcpy(outputs, h->hFwdOut, sizeof(float) * nO);
  return 0;
}

int smb200_forward_seq(smb200_learner* h, const float* states, const int32_t* lengths, int32_t n, int32_t max_len, float* outputs) {
  if (!h || !states || !lengths || !outputs || n < 1 || max_len < 1) return SMB200_ERR_INVALID;
  for (int32_t i = 0; i < n; ++i)
    if (lengths[i] < 1 || lengths[i] > max_len) { set_error_msg("smb200_forward_seq: window lengths must be in [1, max_len]"); return SMB200_ERR_INVALID; }
  std::lock_guard<std::recursive_mutex> lock(h->apiMutex);
  cudaSetDevice(h->cfg.device);
  const NetDesc& net = h->descs.net;
  const int dS = h->cfg.dim_state, nOut = net.nOut;
  // feed-forward nets see the newest state only; recurrent nets the newest min(length, nnBPTTseq + 1) states
  const int keep = net.recurrent ? std::min<int>(max_len, net.Tc) : 1;
  const size_t nIn = (size_t)n * keep * dS, nO = (size_t)n * nOut;
  if (forward_buffers(h, nIn, nO, (size_t)n)) return SMB200_ERR_CUDA;
  for (int32_t i = 0; i < n; ++i) {
    const int len = std::min<int>(lengths[i], keep);
    std::memcpy(h->hFwdIn + (size_t)i * keep * dS, states + ((size_t)i * max_len + (lengths[i] - len)) * dS, sizeof(float) * (size_t)len * dS);
    h->hFwdLen[i] = len;
  }
  SMB200_CUDA_CHECK(cudaMemcpyAsync(h->fwdIn, h->hFwdIn, sizeof(float) * nIn, cudaMemcpyHostToDevice, h->stream));
  StepArgs a = h->args();
  if (net.recurrent) {
    SMB200_CUDA_CHECK(cudaMemcpyAsync(h->fwdLen, h->hFwdLen, sizeof(int) * (size_t)n, cudaMemcpyHostToDevice, h->stream));
    if (launch_forward_seq(a, net, h->fwdIn, h->fwdLen, n, keep, h->fwdOut, h->stream)) return SMB200_ERR_CUDA;
  } else if (launch_forward(a, net, h->fwdIn, n, h->fwdOut, h->stream)) return SMB200_ERR_CUDA;
  if (d2h(h, h->hFwdOut, h->fwdOut, sizeof(float) * nO)) return SMB200_ERR_CUDA;
  std::memcpy(outputs, h->hFwdOut, sizeof(float) * nO);
  return 0;
}

}  // extern "C"

// =============================================================================================
// Checkpoint files of the reference (row f2): Learner_approximator::save / restart
// (Learners/Learner_approximator.cpp:118-142) = Network::save/restart of weights, target weights and
// the two Adam moments (Network/Network.cpp:22-67, Optimizer.cpp:180-215; per layer, padding
// stripped: Layer_Base.h:143-169, Layers.h:401-418,554-566, Layer_LSTM.h:189-211) +
// MemoryBuffer::save/restart (ReplayMemory/MemoryBuffer.cpp:172-324: scaling, counters, episodes
// packed by Episode::packEpisode, Episode.cpp:24-93).  Byte-compatible both ways for MDPs whose
// state is fully observed (the device replay does not store latent state components).
// =============================================================================================
namespace smb200 {

static size_t stripped_size(const NetDesc& net) {
  size_t n = 0;
  for (int l = 1; l < net.nLayers; ++l) {
    const LayerDesc& L = net.L[l];
    if (L.kind == kDenseTanh || L.kind == kDenseLinear) n += (size_t)L.size * (L.nIn + 1);
    else if (is_cell_layer(L.kind)) n += (size_t)cell_gates(L.kind) * L.size * (L.nIn + L.size + 1);
    else if (L.kind == kResidual) n += 2 * (size_t)L.size;
    else if (L.kind == kParam) n += L.size;
  }
  return n;
}
// dir = +1: padded blob -> file order; dir = -1: file order -> padded blob
static void strip_copy(const NetDesc& net, float* blob, float* flat, int dir) {
  size_t o = 0;
  auto mv = [&](int blobIdx) { if (dir > 0) flat[o] = blob[blobIdx]; else blob[blobIdx] = flat[o]; ++o; };
  for (int l = 1; l < net.nLayers; ++l) {
    const LayerDesc& L = net.L[l];
    if (L.kind == kDenseTanh || L.kind == kDenseLinear) {
      for (int i = 0; i < L.nIn; ++i) for (int n = 0; n < L.size; ++n) mv(L.wOff + n + L.ld * i);
      for (int n = 0; n < L.size; ++n) mv(L.bOff + n);
    } else if (is_cell_layer(L.kind)) {
      const int ng = cell_gates(L.kind);
      for (int w = 0; w < ng * L.size * (L.nIn + L.size); ++w) mv(L.wOff + w);
      for (int n = 0; n < ng * L.size; ++n) mv(L.bOff + n);
    } else if (L.kind == kResidual) {
      for (int n = 0; n < L.size; ++n) mv(L.wOff + n);
      for (int n = 0; n < L.size; ++n) mv(L.bOff + n);
    } else if (L.kind == kParam) {
      for (int n = 0; n < L.size; ++n) mv(L.bOff + n);
    }
  }
}
// the reference writes <name>_backup.raw first and then copies it to <name>.raw (Network.cpp:27-39)
static int write_both(const std::string& stem, const void* data, size_t bytes, bool text = false) {
  for (const char* suffix : {"_backup.raw", ".raw"}) {
    FILE* f = fopen((stem + suffix).c_str(), text ? "w" : "wb");
    if (!f) { set_error_msg(("cannot write " + stem + suffix).c_str()); return -1; }
    const size_t w = fwrite(data, 1, bytes, f);
    fclose(f);
    if (w != bytes) { set_error_msg(("short write on " + stem + suffix).c_str()); return -1; }
  }
  return 0;
}
// One episode of <...>_learner_data.raw: `size_t N` followed by Episode::packEpisode's floats (Episode.cpp:24-85;
// size Episode::computeTotalEpisodeSize, Episode.h:211-219): N x [state | reward | action | policy], six length-N arrays
// (Qret, A, V, delta, rho, KL) and a 10-float tail {bool terminated; ssize_t ID, just_sampled, agentID}.
// F = {reward, Qret, A, V, delta, rho, KL}, N values each.
static void pack_episode(std::vector<unsigned char>& out, int dS, int dA, int dP, size_t N, const float* S, const float* A,
                         const float* MU, const float* const F[7], bool term, ptrdiff_t id, ptrdiff_t agentId) {
  const size_t tot = (size_t)(dS + dA + dP + 1 + 6) * N + 10;
  std::vector<float> buf(tot, 0.f);
  float* b = buf.data();
  for (size_t t = 0; t < N; ++t) {
    memcpy(b, S + t * dS, sizeof(float) * dS); b[dS] = F[0][t]; b += dS + 1;
    memcpy(b, A + t * dA, sizeof(float) * dA); b += dA;
    memcpy(b, MU + t * dP, sizeof(float) * dP); b += dP;
  }
  for (int k = 1; k < 7; ++k) { memcpy(b, F[k], sizeof(float) * N); b += N; }
  char* c = reinterpret_cast<char*>(b);
  const ptrdiff_t js = -1;                                           // just_sampled: reset by updateTrainingStatistics every step
  memcpy(c, &term, sizeof(bool)); c += sizeof(bool);
  memcpy(c, &id, sizeof(ptrdiff_t)); c += sizeof(ptrdiff_t);
  memcpy(c, &js, sizeof(ptrdiff_t)); c += sizeof(ptrdiff_t);
  memcpy(c, &agentId, sizeof(ptrdiff_t));
  const size_t seqLen = N;
  const unsigned char* p0 = reinterpret_cast<const unsigned char*>(&seqLen);
  out.insert(out.end(), p0, p0 + sizeof(size_t));
  const unsigned char* p1 = reinterpret_cast<const unsigned char*>(buf.data());
  out.insert(out.end(), p1, p1 + sizeof(float) * tot);
}
// The inverse (Episode::unpackEpisode, Episode.cpp:87-130): reads the episode at `pos` and advances it.  S/A/MU/R are copied
// out of the interleaved tuples; Q, ADV, V, delta, rho, KL point into `dat`.
struct UnpackedEpisode {
  size_t N = 0;
  std::vector<float> S, A, MU, R;
  const float *Q = nullptr, *ADV = nullptr, *V = nullptr, *delta = nullptr, *rho = nullptr, *KL = nullptr;
  bool term = false; ptrdiff_t id = 0, js = 0, ag = 0;
};
static int unpack_episode(const std::vector<unsigned char>& dat, size_t& pos, int dS, int dA, int dP, UnpackedEpisode& u) {
  if (pos + sizeof(size_t) > dat.size()) { set_error_msg("Unable to find sequence in learner_data.raw"); return SMB200_ERR_STATE; }
  size_t N; memcpy(&N, dat.data() + pos, sizeof(size_t)); pos += sizeof(size_t);
  const size_t tot = (size_t)(dS + dA + dP + 1 + 6) * N + 10;
  if (N < 2 || pos + sizeof(float) * tot > dat.size()) { set_error_msg("mismatch in learner_data.raw"); return SMB200_ERR_STATE; }
  const float* b = reinterpret_cast<const float*>(dat.data() + pos); pos += sizeof(float) * tot;
  u.N = N;
  u.S.resize(N * dS); u.A.resize(N * dA); u.MU.resize(N * dP); u.R.resize(N);
  for (size_t t = 0; t < N; ++t) {
    memcpy(u.S.data() + t * dS, b, sizeof(float) * dS); u.R[t] = b[dS]; b += dS + 1;
    memcpy(u.A.data() + t * dA, b, sizeof(float) * dA); b += dA;
    memcpy(u.MU.data() + t * dP, b, sizeof(float) * dP); b += dP;
  }
  u.Q = b; u.ADV = b + N; u.V = b + 2 * N; u.delta = b + 3 * N; u.rho = b + 4 * N; u.KL = b + 5 * N;
  const char* c = reinterpret_cast<const char*>(b + 6 * N);
  memcpy(&u.term, c, sizeof(bool)); c += sizeof(bool);
  memcpy(&u.id, c, sizeof(ptrdiff_t)); c += sizeof(ptrdiff_t);
  memcpy(&u.js, c, sizeof(ptrdiff_t)); c += sizeof(ptrdiff_t);
  memcpy(&u.ag, c, sizeof(ptrdiff_t));
  return 0;
}
static int read_all(const std::string& path, std::vector<unsigned char>& out) {
  FILE* f = fopen(path.c_str(), "rb");
  if (!f) return 1;
  fseek(f, 0, SEEK_END); const long n = ftell(f); fseek(f, 0, SEEK_SET);
  out.resize(n > 0 ? (size_t)n : 0);
  const size_t r = out.empty() ? 0 : fread(out.data(), 1, out.size(), f);
  fclose(f);
  return r == out.size() ? 0 : -1;
}
static const float* field_ptr(const ReplayView& rp, int field) {
  switch (field) {
    case SMB200_F_V: return rp.V;         case SMB200_F_ADV: return rp.ADV;   case SMB200_F_QRET: return rp.Q;
    case SMB200_F_DELTA: return rp.DELTA; case SMB200_F_RHO: return rp.RHO;   case SMB200_F_KL: return rp.KL;
    case SMB200_F_REWARD: return rp.R;    default: return nullptr;
  }
}

}  // namespace smb200

extern "C" {

// The episode file format without a device (diagnostics for the CPU test suite): every episode of a <...>_learner_data.raw
// image is read by the parser smb200_restart uses and written again by the packer smb200_save uses.  Returns the bytes
// written to `out` (negative on error); ids / n_rows / terminated (each optional, capacity max_eps) describe what was read.
int64_t smb200_host_repack_episodes(int32_t dim_state, int32_t dim_action, const uint8_t* in, int64_t n_in, uint8_t* out, int64_t capacity,
                                    int64_t max_eps, int64_t* n_episodes, int64_t* ids, int32_t* n_rows, int32_t* terminated) {
  if (dim_state < 1 || dim_action < 1 || !in || n_in < 0 || !out) return SMB200_ERR_INVALID;
  const int dS = dim_state, dA = dim_action, dP = 2 * dA;
  const std::vector<unsigned char> dat(in, in + n_in);
  std::vector<unsigned char> packed;
  size_t pos = 0; int64_t n = 0;
  UnpackedEpisode u;
  while (pos < dat.size()) {
    const int rc = unpack_episode(dat, pos, dS, dA, dP, u);
    if (rc) return rc;
    const float* F[7] = {u.R.data(), u.Q, u.ADV, u.V, u.delta, u.rho, u.KL};
    pack_episode(packed, dS, dA, dP, u.N, u.S.data(), u.A.data(), u.MU.data(), F, u.term, u.id, u.ag);
    if (n < max_eps) {
      if (ids) ids[n] = (int64_t)u.id;
      if (n_rows) n_rows[n] = (int32_t)u.N;
      if (terminated) terminated[n] = u.term ? 1 : 0;
    }
    ++n;
  }
  if ((int64_t)packed.size() > capacity) { set_error_msg("repack_episodes: output buffer too small"); return SMB200_ERR_CAPACITY; }
  memcpy(out, packed.data(), packed.size());
  if (n_episodes) *n_episodes = n;
  return (int64_t)packed.size();
}

// Checkpoint weight order without a device (diagnostics for the CPU test suite): the library's own strip_copy between the
// padded parameter blob (Parameters.h:159-176) and the order Network::save writes (Network.cpp:22-67; per layer, padding
// stripped: Layer_Base.h:143-169, Layers.h:401-418,554-566, Layer_LSTM.h:189-211).  dir = +1: blob -> flat, -1: flat -> blob
// (padding left untouched).  flat == nullptr: returns the stripped size.
int64_t smb200_host_strip_weights(const smb200_config* cfg, float* blob, int64_t n_blob, float* flat, int64_t n_flat, int32_t dir) {
  if (!cfg) return SMB200_ERR_INVALID;
  NetDesc* net = new NetDesc();
  std::vector<GradTile> tiles;
  if (build_net(*cfg, *net, tiles)) { delete net; return SMB200_ERR_INVALID; }
  const int64_t ns = (int64_t)stripped_size(*net);
  if (flat) {
    if (!blob || n_blob != net->nParams || n_flat != ns || (dir != 1 && dir != -1)) {
      delete net; set_error_msg("strip_weights: size mismatch"); return SMB200_ERR_INVALID; }
    strip_copy(*net, blob, flat, dir);
  }
  delete net;
  return ns;
}


int smb200_write_field(smb200_learner* h, int32_t field, const float* in, int64_t n) {
  if (!h || !in || n != smb200_n_rows(h)) return SMB200_ERR_INVALID;
  float* dst = const_cast<float*>(field_ptr(h->rp, field));
  if (!dst) return SMB200_ERR_INVALID;
  cudaSetDevice(h->cfg.device);
  int64_t o = 0;
  for (const auto& e : h->episodes) {
    SMB200_CUDA_CHECK(cudaMemcpyAsync(dst + e.start, in + o, sizeof(float) * e.nRows, cudaMemcpyHostToDevice, h->stream));
    o += e.nRows;
  }
  SMB200_CUDA_CHECK(cudaStreamSynchronize(h->stream));
  return 0;
}

int smb200_push_episode_restored(smb200_learner* h, int64_t id, int32_t N, int32_t terminated, const float* S, const float* A,
                                 const float* MU, const float* R, const float* V, const float* ADV, const float* Q,
                                 const float* DELTA, const float* RHO, const float* KL, double cmax) {
  if (!V || !ADV || !Q || !DELTA || !RHO || !KL || !(cmax > 0)) return SMB200_ERR_INVALID;
  const float* rest[4] = {Q, DELTA, RHO, KL};
  return push_episode_impl(h, id, N, terminated, S, A, MU, R, V, ADV, rest, (float)cmax, (float)(1.0 / cmax));
}

int smb200_set_refer(smb200_learner* h, double beta, double cmax) {
  if (!h || !(cmax > 0) || beta < 0 || beta > 1) return SMB200_ERR_INVALID;
  cudaSetDevice(h->cfg.device);
  if (pull_ctrl(h)) return SMB200_ERR_CUDA;
  h->hCtrl.beta = beta; h->hCtrl.cmax = cmax; h->hCtrl.cinv = 1.0 / cmax;
  h->initialized = true;
  return push_ctrl(h);
}

int smb200_save(smb200_learner* h, const char* base_c) {
  if (!h || !base_c) return SMB200_ERR_INVALID;
  cudaSetDevice(h->cfg.device);
  const std::string base(base_c);
  const NetDesc& net = h->descs.net;
  const size_t nP = net.nParams, nF = stripped_size(net);
  // ---- Approximator::save -> AdamOptimizer::save (Optimizer.cpp:180-197) ----
  std::vector<float> blob(nP), flat(nF);
  const float* src[4] = {h->W, h->Wtgt /* null with targetDelay 0: the host copy */, h->M1, h->M2};
  const char* name[4] = {"_net_weights", "_net_tgt_weights", "_net_1stMom", "_net_2ndMom"};
  for (int k = 0; k < 4; ++k) {
    if (src[k]) { if (d2h(h, blob.data(), src[k], sizeof(float) * nP)) return SMB200_ERR_CUDA; }
    else blob = h->tgtBlob;
    strip_copy(net, blob.data(), flat.data(), +1);
    if (write_both(base + name[k], flat.data(), sizeof(float) * nF)) return SMB200_ERR_STATE;
  }
  // ---- MemoryBuffer::save (MemoryBuffer.cpp:267-324) ----
  const int dS = h->cfg.dim_state, dA = h->cfg.dim_action, dP = h->dP;
  {
    std::vector<float> mean(dS), scale(dS), stdev(dS); float rew[4];
    if (smb200_get_scaling(h, mean.data(), scale.data(), stdev.data(), rew)) return SMB200_ERR_CUDA;
    std::vector<double> V;
    V.insert(V.end(), mean.begin(), mean.end()); V.insert(V.end(), scale.begin(), scale.end()); V.insert(V.end(), stdev.begin(), stdev.end());
    V.push_back((double)rew[2]); V.push_back((double)rew[1]); V.push_back((double)rew[0]);    // rewardsStdDev, rewardsScale, rewardsMean
    if (write_both(base + "_scaling", V.data(), sizeof(double) * V.size())) return SMB200_ERR_STATE;
  }
  if (pull_ctrl(h)) return SMB200_ERR_CUDA;
  char rk[16]; snprintf(rk, sizeof(rk), "%03d", h->cfg.world_rank);
  const std::string stem = base + "_rank_" + rk + "_learner_";
  {
    char txt[512];
    // `doneGradSteps = counters.nGradSteps + 1`: Learner::save runs inside logStats, before the step counter moves
    const int len = snprintf(txt, sizeof(txt), "nStoredEps: %lu\nnStoredObs: %lu\nnLocalSeenEps: %lu\nnLocalSeenObs: %lu\n"
                             "nInitialData: %ld\nnGradSteps: %ld\nCmaxReFER: %le\nbeta: %le\n",
                             (unsigned long)h->episodes.size(), (unsigned long)h->nTransitions, (unsigned long)h->nSeenEps,
                             (unsigned long)h->nSeenObs, (long)h->nGatheredB4Startup, (long)(h->gradStep + 1), h->hCtrl.cmax, h->hCtrl.beta);
    if (write_both(stem + "status", txt, (size_t)len, true)) return SMB200_ERR_STATE;
  }
  {
    const long long hw = h->highWater;
    std::vector<float> S((size_t)hw * dS), A((size_t)hw * dA), MU((size_t)hw * dP), F[7];
    if (d2h(h, S.data(), h->rp.S, sizeof(float) * S.size()) || d2h(h, A.data(), h->rp.A, sizeof(float) * A.size()) ||
        d2h(h, MU.data(), h->rp.MU, sizeof(float) * MU.size())) return SMB200_ERR_CUDA;
    const int fid[7] = {SMB200_F_REWARD, SMB200_F_QRET, SMB200_F_ADV, SMB200_F_V, SMB200_F_DELTA, SMB200_F_RHO, SMB200_F_KL};
    for (int k = 0; k < 7; ++k) { F[k].resize(hw); if (d2h(h, F[k].data(), field_ptr(h->rp, fid[k]), sizeof(float) * hw)) return SMB200_ERR_CUDA; }
    std::vector<unsigned char> out;
    for (const auto& e : h->episodes) {
      const size_t r = (size_t)e.start;
      const float* Fe[7];
      for (int k = 0; k < 7; ++k) Fe[k] = F[k].data() + r;
      pack_episode(out, dS, dA, dP, (size_t)e.nRows, S.data() + r * dS, A.data() + r * dA, MU.data() + r * dP, Fe, e.terminated != 0,
                   (ptrdiff_t)e.id, (ptrdiff_t)e.agentId);
    }
    if (write_both(stem + "data", out.data(), out.size())) return SMB200_ERR_STATE;
  }
  return 0;
}

int smb200_restart(smb200_learner* h, const char* base_c) {
  if (!h || !base_c) return SMB200_ERR_INVALID;
  if (!h->episodes.empty() || h->gradStep != 0) { set_error_msg("restart needs a freshly created learner"); return SMB200_ERR_STATE; }
  cudaSetDevice(h->cfg.device);
  const std::string base(base_c);
  const NetDesc& net = h->descs.net;
  const size_t nP = net.nParams, nF = stripped_size(net);
  std::vector<unsigned char> raw;
  // ---- AdamOptimizer::restart (Optimizer.cpp:199-215): weights are mandatory, the rest optional ----
  std::vector<float> blob(nP, 0.f), m1(nP, 0.f), m2(nP, 0.f);
  auto load_net = [&](const char* name, std::vector<float>& dst) -> int {
    const int rc = read_all(base + name + ".raw", raw);
    if (rc) return rc;
    if (raw.size() != sizeof(float) * nF) { set_error_msg((std::string("Mismatch in restarted file ") + base + name).c_str()); return -1; }
    strip_copy(net, dst.data(), reinterpret_cast<float*>(raw.data()), -1);
    return 0;
  };
  if (d2h(h, blob.data(), h->W, sizeof(float) * nP)) return SMB200_ERR_CUDA;     // padding keeps its current (zero) content
  int rc = load_net("_net_weights", blob);
  if (rc > 0) { set_error_msg(("Parameters restart file " + base + "_net_weights.raw not found").c_str()); return SMB200_ERR_STATE; }
  if (rc < 0) return SMB200_ERR_STATE;
  if (upload_weights(h, blob.data())) return SMB200_ERR_CUDA;
  h->tgtBlob = blob;
  { std::vector<float> t = blob; const int r2 = load_net("_net_tgt_weights", t); if (r2 < 0) return SMB200_ERR_STATE; if (r2 == 0) h->tgtBlob = t; }
  if (h->Wtgt) SMB200_CUDA_CHECK(cudaMemcpyAsync(h->Wtgt, h->tgtBlob.data(), sizeof(float) * nP, cudaMemcpyHostToDevice, h->stream));
  if (load_net("_net_1stMom", m1) < 0 || load_net("_net_2ndMom", m2) < 0) return SMB200_ERR_STATE;
  SMB200_CUDA_CHECK(cudaMemcpyAsync(h->M1, m1.data(), sizeof(float) * nP, cudaMemcpyHostToDevice, h->stream));
  SMB200_CUDA_CHECK(cudaMemcpyAsync(h->M2, m2.data(), sizeof(float) * nP, cudaMemcpyHostToDevice, h->stream));
  SMB200_CUDA_CHECK(cudaStreamSynchronize(h->stream));
  // ---- MemoryBuffer::restart (MemoryBuffer.cpp:172-265) ----
  const int dS = h->cfg.dim_state, dA = h->cfg.dim_action, dP = h->dP;
  if (read_all(base + "_scaling.raw", raw)) return 0;                  // "Parameters restart file ... not found": nothing else is read
  if (raw.size() != sizeof(double) * (3 * (size_t)dS + 3)) { set_error_msg("Mismatch in restarted file _scaling.raw"); return SMB200_ERR_STATE; }
  {
    const double* V = reinterpret_cast<const double*>(raw.data());
    std::vector<float> mean(V, V + dS), scale(V + dS, V + 2 * dS), stdev(V + 2 * dS, V + 3 * dS);
    const float rew[3] = {(float)V[3 * dS + 2], (float)V[3 * dS + 1], (float)V[3 * dS + 0]};
    if (smb200_set_scaling(h, mean.data(), scale.data(), stdev.data(), rew)) return SMB200_ERR_CUDA;
  }
  char rk[16]; snprintf(rk, sizeof(rk), "%03d", h->cfg.world_rank);
  const std::string stem = base + "_rank_" + rk + "_learner_";
  FILE* fs = fopen((stem + "status.raw").c_str(), "r");
  std::vector<unsigned char> dat;
  if (!fs || read_all(stem + "data.raw", dat)) { if (fs) fclose(fs); return 0; }   // scaling only (evaluation runs)
  unsigned long nEps = 0, nObs = 0, seenEps = 0, seenObs = 0; long nInit = 0, grad = 0; double cmax = 0, beta = 0;
  int pass = 1;
  pass = pass && 1 == fscanf(fs, "nStoredEps: %lu\n", &nEps);     pass = pass && 1 == fscanf(fs, "nStoredObs: %lu\n", &nObs);
  pass = pass && 1 == fscanf(fs, "nLocalSeenEps: %lu\n", &seenEps); pass = pass && 1 == fscanf(fs, "nLocalSeenObs: %lu\n", &seenObs);
  pass = pass && 1 == fscanf(fs, "nInitialData: %ld\n", &nInit);    pass = pass && 1 == fscanf(fs, "nGradSteps: %ld\n", &grad);
  pass = pass && 1 == fscanf(fs, "CmaxReFER: %le\n", &cmax);        pass = pass && 1 == fscanf(fs, "beta: %le\n", &beta);
  fclose(fs);
  if (!pass || grad < 0) { set_error_msg("malformed learner_status.raw"); return SMB200_ERR_STATE; }
  const float C = (float)cmax, invC = (float)(1.0 / cmax);
  size_t pos = 0;
  UnpackedEpisode u;
  for (unsigned long e = 0; e < nEps; ++e) {
    const int r1 = unpack_episode(dat, pos, dS, dA, dP, u);
    if (r1) return r1;
    const float* rest[4] = {u.Q, u.delta, u.rho, u.KL};
    const int r2 = push_episode_impl(h, (int64_t)u.id, (int32_t)u.N, u.term ? 1 : 0, u.S.data(), u.A.data(), u.MU.data(), u.R.data(), u.V, u.ADV, rest, C, invC);
    if (r2) return r2;
    h->episodes.back().agentId = (int)u.ag;
  }
  if ((unsigned long)h->nTransitions != nObs) { set_error_msg("learner_status.raw and learner_data.raw disagree on nStoredObs"); return SMB200_ERR_STATE; }
  h->nSeenEps = (long long)seenEps; h->nSeenObs = (long long)seenObs; h->nGatheredB4Startup = nInit;
  // counters.nGradSteps = doneGradSteps; Approximator::setNgradSteps(nGradSteps()) (Learner_approximator.cpp:130);
  // the running beta powers of Adam are NOT part of a checkpoint: they restart from beta_1, beta_2 (Optimizer.h:94)
  if (pull_ctrl(h)) return SMB200_ERR_CUDA;
  StepCtrl& k = h->hCtrl;
  k.beta = beta; k.cmax = cmax; k.cinv = 1.0 / cmax;
  h->gradStep = grad; k.grad_step = grad; k.adam_step = grad;
  h->tgtPhase = grad;                                             // cntUpdateDelay is not part of a checkpoint: 0 in the new process
  k.n_far_ref = 0; k.avg_sq_err = 0;                              // ReplayStats are not restored (zero until the next statistics pass)
  k.gl_far_prev = 0; k.gl_stored_prev = (double)h->nTransitions; k.cnt_seed_step = grad;
  if (push_ctrl(h)) return SMB200_ERR_CUDA;
  h->initialized = true;
  h->orderDirty = true; h->lookupDirty = true; h->tableVersion++; h->ahead_clear();
  h->samplerStale = h->slow_mode();      // (the reference never prepares the sampler of a restarted learner before its first step)
  return 0;
}

}  // extern "C"
