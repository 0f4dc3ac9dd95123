// Microbenchmark of one dense layer of the P1 tile: y[n][4] = tanh(b[n] + sum_k x[k][4] W[k][n]),
// K = N = 128, 4 samples, 512 threads, everything in shared memory (W row stride 132 floats).
// Compares thread mappings; prints SM cycles per layer call.
#include <cstdio>
#include <cmath>
#include <cuda_runtime.h>
constexpr int T = 512, K = 128, N = 128, LDP = 132, REP = 200;
__device__ __forceinline__ float tanh_ref(float x) { const float e = __expf(-2.0f * fabsf(x)); const float y = __fdividef(1.0f - e, 1.0f + e); return x > 0.0f ? y : -y; }

// V0: lanes over n, K split over 4 groups, shared-memory combine (the round-1 kernel)
__device__ void v0(const float* W, const float* b, const float* x, float* y, float* red) {
  const int tid = threadIdx.x, g = tid >> 7, n = tid & 127, kb = g * 32;
  float a0 = 0, a1 = 0, a2 = 0, a3 = 0;
  const float* w = W + kb * LDP + n; const float4* x4 = reinterpret_cast<const float4*>(x) + kb;
#pragma unroll 8
  for (int k = 0; k < 32; ++k) { const float wv = w[k * LDP]; const float4 xv = x4[k]; a0 = fmaf(xv.x, wv, a0); a1 = fmaf(xv.y, wv, a1); a2 = fmaf(xv.z, wv, a2); a3 = fmaf(xv.w, wv, a3); }
  reinterpret_cast<float4*>(red)[g * 128 + n] = make_float4(a0, a1, a2, a3);
  __syncthreads();
  { const int n2 = tid >> 2, s = tid & 3; float v = b[n2];
    for (int gg = 0; gg < 4; ++gg) v += red[(gg * 128 + n2) * 4 + s];
    y[n2 * 4 + s] = tanh_ref(v); }
  __syncthreads();
}
// V1: quad of outputs x 4 samples per thread, K split over G groups of 32 threads, shared-memory combine
template <int G>
__device__ void v1(const float* W, const float* b, const float* x, float* y, float* red) {
  const int tid = threadIdx.x, g = tid >> 5, nq = tid & 31;
  constexpr int Kc = K / G;
  if (g < G) {
    float acc[4][4] = {};
    const float* w = W + g * Kc * LDP + nq * 4; const float4* x4 = reinterpret_cast<const float4*>(x) + g * Kc;
#pragma unroll 4
    for (int k = 0; k < Kc; ++k) {
      const float4 wv = *reinterpret_cast<const float4*>(w + k * LDP), xv = x4[k];
      const float ww[4] = {wv.x, wv.y, wv.z, wv.w}, xx[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int s = 0; s < 4; ++s) acc[i][s] = fmaf(ww[i], xx[s], acc[i][s]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) reinterpret_cast<float4*>(red)[(g * 4 + i) * 32 + nq] = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
  }
  __syncthreads();
  { const int i = tid >> 5 & 3, q = tid & 31;   // 128 outputs as float4 over samples: threads 0..127
    if (tid < 128) {
      float4 v = make_float4(0, 0, 0, 0);
#pragma unroll
      for (int gg = 0; gg < G; ++gg) { const float4 t = reinterpret_cast<const float4*>(red)[(gg * 4 + i) * 32 + q]; v.x += t.x; v.y += t.y; v.z += t.z; v.w += t.w; }
      const int n = q * 4 + i; const float bv = b[n];
      *reinterpret_cast<float4*>(y + n * 4) = make_float4(tanh_ref(v.x + bv), tanh_ref(v.y + bv), tanh_ref(v.z + bv), tanh_ref(v.w + bv));
    } }
  __syncthreads();
}
// V3: a warp owns 8 outputs for all K: lane = (quad = lane/16, kgroup = lane%16), k = kgroup + 16 j;
// butterfly reduce-scatter over the 16 k-groups with 15 shuffles; no block-wide combine.
__device__ void v3(const float* W, const float* b, const float* x, float* y, float* /*red*/) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane & 15, quad = lane >> 4;
  float acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = 0.f;
  const float* w = W + g * LDP + warp * 8 + quad * 4; const float4* x4 = reinterpret_cast<const float4*>(x) + g;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float4 wv = *reinterpret_cast<const float4*>(w + j * 16 * LDP), xv = x4[j * 16];
    const float ww[4] = {wv.x, wv.y, wv.z, wv.w}, xx[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int s = 0; s < 4; ++s) acc[i * 4 + s] = fmaf(ww[i], xx[s], acc[i * 4 + s]);
  }
  // reduce-scatter over lane bits 3,2,1,0: after step with bit d the lane keeps the half of the values selected by its bit
#pragma unroll
  for (int h = 8, d = 8; h >= 1; h >>= 1, d >>= 1) {
    const bool up = (lane & d) != 0;
#pragma unroll
    for (int i = 0; i < h; ++i) {
      const float send = up ? acc[i] : acc[i + h];
      const float keep = up ? acc[i + h] : acc[i];
      acc[i] = keep + __shfl_xor_sync(0xffffffffu, send, d);
    }
  }
  // lane now owns value index v = bits(lane&15) reversed-ish: index = sum over steps; recover (i, s)
  int v = 0;
  { int h = 8; for (int d = 8; d >= 1; d >>= 1, h >>= 1) if (lane & d) v += h; }
  const int i = v >> 2, s = v & 3, n = warp * 8 + quad * 4 + i;
  y[n * 4 + s] = tanh_ref(acc[0] + b[n]);
  __syncthreads();
}

// V5: tensor cores, mma.sync.m16n8k8 TF32 with the 3xTF32 split (a = hi + lo, hi*hi + hi*lo + lo*hi accumulated in f32):
// m = output neuron (8 m-tiles of 16), n = sample (8 columns, 4 used), k = input; warp = (m-tile, K half); the weight
// image needs row stride == 8 (mod 32) floats (136) for conflict-free fragment loads.
constexpr int LDM = 136;
__device__ __forceinline__ unsigned tf32_of(float f) { unsigned u; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(f)); return u; }
__device__ __forceinline__ void mma_tf32(float (&c)[4], const unsigned (&a)[4], unsigned b0, unsigned b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ void v5(const float* W, const float* b, const float* x, float* y, float* red) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, gid = lane >> 2, tig = lane & 3;
  const int mt = warp & 7, kh = warp >> 3;
  float c[4] = {0.f, 0.f, 0.f, 0.f}, c1[4] = {0.f, 0.f, 0.f, 0.f}, c2[4] = {0.f, 0.f, 0.f, 0.f};
  float d0[4] = {0.f, 0.f, 0.f, 0.f}, d1[4] = {0.f, 0.f, 0.f, 0.f}, d2[4] = {0.f, 0.f, 0.f, 0.f};
  const float* w = W + (kh * 64 + tig) * LDM + mt * 16 + gid;
  const float* xx = x + (kh * 64 + tig) * 4 + gid;
#pragma unroll
  for (int ks = 0; ks < 8; ++ks) {
    const float af[4] = {w[ks * 8 * LDM], w[ks * 8 * LDM + 8], w[(ks * 8 + 4) * LDM], w[(ks * 8 + 4) * LDM + 8]};
    const float b0f = gid < 4 ? xx[ks * 32] : 0.f, b1f = gid < 4 ? xx[ks * 32 + 16] : 0.f;
    unsigned ah[4], al[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { ah[i] = tf32_of(af[i]); al[i] = tf32_of(af[i] - __uint_as_float(ah[i])); }
    const unsigned bh0 = tf32_of(b0f), bh1 = tf32_of(b1f);
    const unsigned bl0 = tf32_of(b0f - __uint_as_float(bh0)), bl1 = tf32_of(b1f - __uint_as_float(bh1));
    if (ks & 1) { mma_tf32(d1, al, bh0, bh1); mma_tf32(d2, ah, bl0, bl1); mma_tf32(d0, ah, bh0, bh1); }
    else        { mma_tf32(c1, al, bh0, bh1); mma_tf32(c2, ah, bl0, bl1); mma_tf32(c, ah, bh0, bh1); }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) c[i] = (c[i] + d0[i]) + ((c1[i] + d1[i]) + (c2[i] + d2[i]));
  if (tig < 2) {      // columns 2*tig, 2*tig+1 are real samples
    float* r0 = red + ((kh * 128) + mt * 16 + gid) * 4 + 2 * tig;
    r0[0] = c[0]; r0[1] = c[1]; r0[8 * 4] = c[2]; r0[8 * 4 + 1] = c[3];
  }
  __syncthreads();
  { const int n2 = tid >> 2, s = tid & 3; y[n2 * 4 + s] = tanh_ref(b[n2] + red[n2 * 4 + s] + red[(128 + n2) * 4 + s]); }
  __syncthreads();
}

template <int V>
__global__ void bench(const float* Wg, const float* bg, const float* xg, float* yg, long long* cyc) {
  extern __shared__ __align__(16) float sm[];
  float* W = sm; float* b = W + K * LDP; float* x = b + N; float* y = x + K * 4; float* red = y + N * 4; float* Wm = red + T * 16;
  for (int i = threadIdx.x; i < K * LDP; i += T) W[i] = Wg[i];
  for (int i = threadIdx.x; i < K * LDM; i += T) { const int k = i / LDM, n = i - k * LDM; Wm[i] = n < N ? Wg[k * LDP + n] : 0.f; }
  for (int i = threadIdx.x; i < N; i += T) b[i] = bg[i];
  for (int i = threadIdx.x; i < K * 4; i += T) x[i] = xg[i];
  __syncthreads();
  long long t0 = clock64();
  for (int r = 0; r < REP; ++r) {
    if (V == 0) v0(W, b, x, y, red);
    if (V == 1) v1<8>(W, b, x, y, red);
    if (V == 2) v1<16>(W, b, x, y, red);
    if (V == 3) v3(W, b, x, y, red);
    if (V == 4) v1<4>(W, b, x, y, red);
    if (V == 5) v5(Wm, b, x, y, red);
  }
  long long t1 = clock64();
  for (int i = threadIdx.x; i < N * 4; i += T) yg[i] = y[i];
  if (threadIdx.x == 0) *cyc = (t1 - t0) / REP;
}
int main() {
  float *W, *b, *x, *y; long long* cyc;
  cudaMallocManaged(&W, K * LDP * 4); cudaMallocManaged(&b, N * 4); cudaMallocManaged(&x, K * 16); cudaMallocManaged(&y, N * 16); cudaMallocManaged(&cyc, 8);
  for (int k = 0; k < K; ++k) for (int n = 0; n < LDP; ++n) W[k * LDP + n] = n < N ? 0.05f * sinf(0.37f * k + 0.11f * n) : 0.f;
  for (int n = 0; n < N; ++n) b[n] = 0.01f * n;
  for (int i = 0; i < K * 4; ++i) x[i] = cosf(0.3f * i);
  static float ref[N * 4];
  for (int n = 0; n < N; ++n) for (int s = 0; s < 4; ++s) { double a = b[n]; for (int k = 0; k < K; ++k) a += (double)x[k * 4 + s] * W[k * LDP + n]; ref[n * 4 + s] = (float)tanh(a); }
  const size_t smem = (K * LDP + N + K * 4 + N * 4 + T * 16 + K * LDM) * 4;
  const char* names[] = {"V0 lanes-over-n, 4 K-groups", "V1 quads, 8 K-groups", "V1 quads, 16 K-groups", "V3 warp-owned columns + shuffle reduce-scatter", "V1 quads, 4 K-groups", "V5 mma.sync m16n8k8 3xTF32, 8 m-tiles x 2 K-halves"};
#define RUN(V) { cudaFuncSetAttribute(bench<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); bench<V><<<1, T, smem>>>(W, b, x, y, cyc); cudaDeviceSynchronize(); \
    double e = 0; for (int i = 0; i < N * 4; ++i) e = fmax(e, fabs((double)y[i] - ref[i])); printf("%-48s %6lld cycles/layer  max err %.2e  (%s)\n", names[V], *cyc, e, cudaGetErrorString(cudaGetLastError())); }
  RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5)
  return 0;
}
