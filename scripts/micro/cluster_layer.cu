// Microbenchmark for DESIGN.md §7 item 0: one dense Tanh layer of the P1 tile, y[n][4] = tanh(b[n] + sum_k x[k][4] W[k][n]),
// K = N = 128, 4 samples, 512 threads per CTA, weights resident in shared memory — computed
//   CS = 1: by ONE CTA (the round-1 kernel's situation: the whole 67 KB matrix goes through one shared-memory pipe), or
//   CS = 2, 4: by a thread-block cluster of CS CTAs, each holding N/CS output columns of W; after every layer each CTA
//              stores its slice of the activations into every peer's input buffer through distributed shared memory
//              (st.shared::cluster) and the cluster synchronises once (double-buffered activations).
// The layers form a dependency chain (layer l+1 reads layer l's output) like the forward pass of the step kernel.
// Prints SM cycles per layer and the error against a host f64 reference of the same chain.
// NOT YET RUN ON A GPU (written when the round's GPU budget was spent); build check only:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o scripts/micro/cluster_layer scripts/micro/cluster_layer.cu
#include <cooperative_groups.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <vector>
namespace cg = cooperative_groups;

constexpr int T = 512, K = 128, N = 128, S = 4, LAYERS = 3, REP = 100;

__device__ __forceinline__ float tanh_ref(float x) {
  const float e = __expf(-2.0f * fabsf(x));
  const float y = __fdividef(1.0f - e, 1.0f + e);
  return x > 0.0f ? y : -y;
}

// dynamic shared memory of one CTA:  W slice [K][NL + 4] | b slice [NL] | x[2][K][S] | red[G][NL][S]
template <int CS>
struct Layout {
  static constexpr int NL = N / CS;          // output columns of this CTA
  static constexpr int LDP = NL + 4;         // padded row: conflict-free float accesses with lanes over columns
  static constexpr int G = T / NL;           // K-groups
  static constexpr int KC = K / G;           // k per group
  static constexpr int offB = K * LDP, offX = offB + NL, offRed = offX + 2 * K * S, total = offRed + G * NL * S;
};

template <int CS>
__global__ void __launch_bounds__(T, 1) chain(const float* __restrict__ Wg, const float* __restrict__ bg, const float* __restrict__ xg,
                                              float* __restrict__ yg, long long* cycles, int reps) {
  using L = Layout<CS>;
  extern __shared__ __align__(16) float sm[];
  float* W = sm; float* b = sm + L::offB; float* x = sm + L::offX; float* red = sm + L::offRed;
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = CS > 1 ? (int)cluster.block_rank() : 0;
  const int tid = threadIdx.x, col0 = rank * L::NL;
  // stage this CTA's weight slice (the same matrix serves every layer of the chain), bias slice and the input
  for (int i = tid; i < K * L::NL; i += T) { const int k = i / L::NL, n = i - k * L::NL; W[k * L::LDP + n] = Wg[k * N + col0 + n]; }
  for (int i = tid; i < L::NL; i += T) b[i] = bg[col0 + i];
  for (int i = tid; i < K * S; i += T) x[i] = xg[i];
  __syncthreads();
  if (CS > 1) cluster.sync();
  const int g = tid / L::NL, n = tid - g * L::NL, kb = g * L::KC;
  float* peerX[CS];
#pragma unroll
  for (int r = 0; r < CS; ++r) peerX[r] = CS > 1 ? cluster.map_shared_rank(x, r) : x;
  long long t0 = 0;
  int cur = 0;
  for (int it = 0; it < reps * LAYERS + LAYERS; ++it) {
    if (it == LAYERS) t0 = clock64();                 // first chain = warm-up
    const float* xin = x + cur * K * S;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    const float* w = W + kb * L::LDP + n;
    const float4* x4 = reinterpret_cast<const float4*>(xin) + kb;
#pragma unroll 8
    for (int k = 0; k < L::KC; ++k) {
      const float wv = w[k * L::LDP]; const float4 xv = x4[k];
      a0 = fmaf(xv.x, wv, a0); a1 = fmaf(xv.y, wv, a1); a2 = fmaf(xv.z, wv, a2); a3 = fmaf(xv.w, wv, a3);
    }
    reinterpret_cast<float4*>(red)[g * L::NL + n] = make_float4(a0, a1, a2, a3);
    __syncthreads();
    if (tid < L::NL) {                                // combine the K-groups, activation, publish to every CTA of the cluster
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int gg = 0; gg < L::G; ++gg) { const float4 p = reinterpret_cast<const float4*>(red)[gg * L::NL + tid]; v.x += p.x; v.y += p.y; v.z += p.z; v.w += p.w; }
      const float bv = b[tid];
      const float4 y4 = make_float4(tanh_ref(v.x + bv), tanh_ref(v.y + bv), tanh_ref(v.z + bv), tanh_ref(v.w + bv));
#pragma unroll
      for (int r = 0; r < CS; ++r) reinterpret_cast<float4*>(peerX[r] + (cur ^ 1) * K * S)[col0 + tid] = y4;
    }
    if (CS > 1) cluster.sync(); else __syncthreads();  // the next layer reads every CTA's slice
    cur ^= 1;
  }
  const long long t1 = clock64();
  if (tid == 0 && rank == 0) *cycles = (t1 - t0) / (long long)(reps * LAYERS);
  if (rank == 0) for (int i = tid; i < K * S; i += T) yg[i] = x[cur * K * S + i];
}

template <int CS>
static void run(const float* W, const float* b, const float* x, float* y, long long* cyc, const std::vector<double>& ref, int reps) {
  using L = Layout<CS>;
  const size_t smem = sizeof(float) * L::total;
  cudaFuncSetAttribute(chain<CS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(CS); cfg.blockDim = dim3(T); cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = CS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, chain<CS>, W, b, x, y, cyc, reps);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  double err = 0;
  for (int i = 0; i < K * S; ++i) err = fmax(err, fabs((double)y[i] - ref[i]));
  printf("cluster of %d CTA(s), %3d columns each, %2d K-groups: %6lld cycles/layer, %5.1f KB smem, max err %.2e (%s)\n", CS, L::NL, L::G, *cyc,
         smem / 1024.0, err, cudaGetErrorString(e));
}

int main() {
  float *W, *b, *x, *y; long long* cyc;
  cudaMallocManaged(&W, K * N * 4); cudaMallocManaged(&b, N * 4); cudaMallocManaged(&x, K * S * 4); cudaMallocManaged(&y, K * S * 4);
  cudaMallocManaged(&cyc, 8);
  for (int k = 0; k < K; ++k) for (int n = 0; n < N; ++n) W[k * N + n] = 0.05f * sinf(0.37f * k + 0.11f * n);
  for (int n = 0; n < N; ++n) b[n] = 0.001f * n;
  for (int i = 0; i < K * S; ++i) x[i] = cosf(0.3f * i);
  // host reference of the same chain (warm-up chain + REP chains of LAYERS layers), f64
  std::vector<double> cur(K * S), nxt(K * S);
  for (int i = 0; i < K * S; ++i) cur[i] = x[i];
  for (int it = 0; it < REP * LAYERS + LAYERS; ++it) {
    for (int n = 0; n < N; ++n) for (int s = 0; s < S; ++s) {
      double a = b[n];
      for (int k = 0; k < K; ++k) a += cur[k * S + s] * (double)W[k * N + n];
      nxt[n * S + s] = tanh(a);
    }
    cur.swap(nxt);
  }
  run<1>(W, b, x, y, cyc, cur, REP);
  run<2>(W, b, x, y, cyc, cur, REP);
  run<4>(W, b, x, y, cyc, cur, REP);
  return 0;
}
