// common.cuh — device-visible descriptors shared by the step and sweep kernels.
//
// Vocabulary follows the reference (cselab/smarties): episodes, transitions (rows), the
// replay MemoryBuffer, ReF-ER coefficients (beta, Cmax), Retrace return estimates.
// Reference citations are relative to /root/reference/source/smarties/.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace smb200 {

constexpr int kMaxLayers = 2 * 8 + 4;   // input + (dense, residual)*hidden + out + param
constexpr int kThreads   = 256;         // every kernel here uses 256-thread CTAs
constexpr int kTileK     = 16;          // weight-gradient tile: 16 input rows x 16 output cols
constexpr int kTileN     = 16;

enum LayerKind : int { kInput = 0, kDenseTanh = 1, kResidual = 2, kDenseLinear = 3, kParam = 4, kLSTM = 5, kMGU = 6 };
// gate columns per cell of a recurrent-cell layer: LSTMLayer 4 (Layer_LSTM.h:24-29), MGULayer 2 (forget | state, Layer_GRU.h:27-33)
__host__ __device__ inline int cell_gates(int kind) { return kind == kLSTM ? 4 : (kind == kMGU ? 2 : 1); }
__host__ __device__ inline bool is_cell_layer(int kind) { return kind == kLSTM || kind == kMGU; }

// One layer of the network built by RACER::setupNet (Learners/RACER_common.cpp:70-115,
// Network/Builder.cpp:48-99).  Layer ids equal the reference's (0 = input).
struct LayerDesc {
  int kind;
  int size;        // number of neurons = size of this layer's activation (LSTM: nCells)
  int nIn;         // dense / LSTM: fan-in from the layer below
  int ld;          // dense: row stride of W[nIn][ld] (roundUp8(size), Layer_Base.h:46); LSTM: 4*nCells, MGU: 2*nCells, rows nIn+nCells
  int wOff, bOff;  // offsets into the padded parameter blob (Parameters.h:159-176)
  int needDx;      // 1 if the backward pass propagates into this layer's input (not the first layer)
  int imgW, imgB;  // offsets into the weight IMAGE (the shared-memory layout, see NetDesc::imgFloats)
  int ldp;         // dense: row stride of W in the image = roundUp4(size) + 4 (bank-conflict-free both ways)
  int fwdShift;    // dense: log2 of the thread-group width of the forward pass  (power of two >= min(size, threads))
  int bwdShift;    // dense: log2 of the thread-group width of the input-gradient pass (>= min(nIn, threads))
  int in;          // id of the input layer (ID - link); residual: ID-1 and ID-2 implied
  int actOff;      // offset (in floats per sample) of this layer's activation in act buffers
};

struct NetDesc {
  int nLayers;
  int nParams;       // padded blob size
  int imgFloats;     // size of the weight image (multiple of 4 floats)
  int nOut;          // network outputs (dense-out + param layer)
  int nOutDense;     // outputs of the linear output layer
  int dS, dA;
  int actPerSample;  // floats per sample over all layers
  int maxWidth;      // widest layer
  // recurrent networks (nnType LSTM): sampled transition t is evaluated on the window
  // [t - min(bptt, t), t] (MemoryBuffer.cpp:393-402) with BPTT over the window (Network.h:155-193)
  int recurrent;     // 1 if the hidden layers are LSTM layers
  int bptt;          // nnBPTTseq
  int Tc;            // bptt + 1 = longest window; P2 column of (sample b, window step k) = b*Tc + k
  int topInOff;      // row of the compact copy (column = b) of the top hidden layer's output at the sampled step
  int seqFloats;     // shared-memory floats of the per-sample sequence workspace (step_kernels.cu: SeqPlan)
  int func;          // "nnFunc" of the hidden dense layers: 0 Tanh, 1 SoftSign, 2 HardSign, 3 Sigm, 4 Relu, 5 LRelu
                     // (Network/Layers/Functions.h:90-480)
  int discrete;      // K > 0: one discrete action component with K options — outputs [V | advantages(K) | policy(K)], MU rows of K
                     // probabilities (RACER<Discrete_advantage, Discrete_policy, Uint>, Learners/RACER.cpp:114)
  LayerDesc L[kMaxLayers];
};

// Work item of the weight-gradient + Adam phase.
struct GradTile {
  int kind;    // 0: dense / LSTM 16x16 tile (row K = bias row), 1: residual vector, 2: param-layer bias
  int layer;
  int k0, n0;
  int cols;    // contraction length (columns of the feature-major scratch, multiple of 256)
  int nLimit;  // > 0: the tile's columns end here (MGU layers: tiles do not straddle the forget | state halves)
  int pad_[2];
};

// Scalars a step needs; double-buffered by step parity: step k reads ctrl[k&1], its
// statistics phase writes ctrl[(k+1)&1].
struct StepCtrl {
  double beta, cmax, cinv;            // ReF-ER (MemoryBuffer.h:41-44)
  double adam_bt1, adam_bt2;          // running beta powers (Optimizer.h:93, Optimizer.cpp:155-158)
  long long adam_step;                // AdamOptimizer::nStep BEFORE this step's prepare_update
  long long grad_step;                // counters.nGradSteps before this step
  long long n_far_ref, n_far_exact;   // stats.nFarPolicySteps (reference formula) / exact flags
  float adam_eta; float pad_;         // learning rate of THIS step incl. bias correction (struct Adam ctor, Optimizer.cpp:64-67)
  // multi-rank: global far-policy / stored counts of the previous step (DelayedReductor, MemoryProcessing.cpp:48-58)
  double gl_far_prev, gl_stored_prev; long long cnt_seed_step;
  double avg_kl, avg_sq_err, max_abs_err, avg_return, stdev_q, avg_q, max_q, min_q;
  double sum_ret_err; long long cnt_ret;
};

// Replay MemoryBuffer in HBM: structure-of-arrays over ring rows (one row = one time step of
// one episode, terminal/truncated row included), plus the episode table indexed by slot.
struct ReplayView {
  // per row
  float* S;  float* A;  float* MU;  float* R;           // states [rows][dS], actions [rows][dA], mu [rows][2dA], reward
  float* V;  float* ADV; float* Q;  float* DELTA; float* RHO; float* KL;  // Episode.h:66-75
  uint8_t* rowFlag;                                       // bit0 first row, bit1 last row, bit2 live
  // per episode slot
  int* epStart; int* epLen; int* epTerm; long long* epId;
  float* epAgg;       // [9][maxEpisodes]: avgKL, fracFar, avgSqErr, maxAbsErr, sumQ2, sumQ, maxQ, minQ, totR
  int* epOrder;       // [nEpisodes] slot at each position of the reference's `episodes` vector
  int* epPos;         // [maxEpisodes] the inverse: position of a slot
  int maxEpisodes;
  long long capRows;
  int dS, dA;
  float* stateMean; float* stateScale; float* stateStd;   // MDPdescriptor (Core/StateAction.h:56-58)
  float* rew;         // {rewardsMean, rewardsScale, rewardsStdDev}
};

enum { AGG_KL = 0, AGG_FAR = 1, AGG_E2 = 2, AGG_MAXE = 3, AGG_Q2 = 4, AGG_Q1 = 5, AGG_MAXQ = 6, AGG_MINQ = 7, AGG_TOTR = 8, AGG_N = 9 };

// Per-sample record written by the loss phase and consumed by the statistics phase
// (Episode::updateCumulative_atomic / updateValues_atomic, Episode.h:112-145).
struct __align__(16) SampleRec {
  int slot; int hasNext; int farDelta; int pad;
  float dKL, dFar, dE2, absE;
  float qOld, qNew, qNextOld, qNextNew;
};

struct Hyper {
  double gamma, lambda, clipImpWeight, penalTol, epsAnneal, learnrate, nnLambda;
  long long maxTotObsGlobal; int batchGlobal; int batchLocal;
  int referThreads; int algo;
  unsigned char bounded[64];
};

// `Uint n; n += x;` with float x exactly as gcc compiles it for x86-64 in the reference's far-policy
// count (MemoryProcessing.cpp:202-227, `nOffPol += Nsteps * EP.fracFarPolSteps`): n is rounded to float,
// added, and converted back with cvttss2si — directly below 2^63 (a negative sum wraps to 2^64 - |sum|,
// out of range gives the "integer indefinite" 2^63), as cvttss2si(f - 2^63) ^ 2^63 from there on (anything
// >= 2^64, such as the float nearest to a wrapped value, becomes 0).  CUDA's own float -> unsigned
// conversion saturates instead.  Host and device: tests/test_host_logic.py pins this against the oracle's
// `uint_plus_float` through smb200_uint_plus_float.
__host__ __device__ inline unsigned long long uint_plus_float_x86(unsigned long long n, float x) {
  const float two63 = 9223372036854775808.0f;
  const unsigned long long indefinite = 0x8000000000000000ull;
#ifdef __CUDA_ARCH__
  const float f = __ull2float_rn(n) + x;
#else
  const volatile float nf = (float)n;          // volatile: no double-precision contraction of the float sum
  const volatile float fv = nf + x;
  const float f = fv;
#endif
  if (!(f >= two63)) {                         // gcc: `comiss; jae` — an unordered compare (NaN) takes this branch too
    if (!(f > -two63)) return indefinite;      // below the signed range, or NaN
    return (unsigned long long)(long long)f;   // truncation; negative values wrap
  }
  const float g = f - two63;
  const unsigned long long t = g < two63 ? (unsigned long long)(long long)g : indefinite;
  return t ^ indefinite;
}

#define SMB200_CUDA_CHECK(expr) do { cudaError_t e__ = (expr); if (e__ != cudaSuccess) { \
  smb200::set_error(#expr, e__, __FILE__, __LINE__); return -2; } } while (0)

void set_error(const char* what, cudaError_t e, const char* file, int line);
void set_error_msg(const char* msg);

}  // namespace smb200
