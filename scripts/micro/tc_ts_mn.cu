// scripts/micro/tc_ts_mn.cu — the two tcgen05 features the large-batch ("wide") learner step is built on, checked in
// isolation before the kernel relies on them:
//   * A operand read from TENSOR MEMORY (tcgen05.mma ... [d_tmem], [a_tmem], b_desc — the ".ts" form), written there by
//     tcgen05.st.32x32b (lane = row m, column = k, one 32-bit column per tf32 element);
//   * B operand MN-major: the SAME shared-memory image of a weight matrix W[k][n] that the forward product reads K-major
//     (contraction over k) is read with the transposed descriptor by the input-gradient product (contraction over n).
// Storage of an operand with R rows and Kk contraction columns: float4 smem[Kk/4][R]  (the K-major no-swizzle canonical
// layout ((8,n),2):((1,SBO),LBO) of cute/atom/mma_traits_sm100.hpp:192-199 with SBO = 128 B, LBO = R*16 B).  Read MN-major
// (rows become the contraction) it is ((1,n),(8,k)):((X,SBO'),(1,LBO')) with SBO' = R*16 B, LBO' = 128 B (:169-187, :240-269).
// Inputs are small multiples of 1/8 (exact in tf32, sums exact in f32): every variant must reproduce the CPU sum EXACTLY.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/micro/tc_ts_mn scripts/micro/tc_ts_mn.cu && ./scripts/micro/tc_ts_mn
#include <cstdint>
#include <cstdio>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>

constexpr int T = 128;
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t umma_desc(const void* base, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_u32(base) >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
__host__ __device__ constexpr uint32_t idesc_tf32(int m, int n, int bMN) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)bMN << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void umma_ss(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
               ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_ts(uint32_t d, uint32_t ta, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n"
               ::"r"(d), "r"(ta), "l"(db), "r"(idesc), "r"(acc) : "memory");
}

// D[128][N] = A[128][K] * Bop, where the B image holds S[R][Kk]:
//   bMN = 0: N = R,  K = Kk, D[m][n] = sum_k A[m][k] S[n][k]
//   bMN = 1: N = Kk, K = R,  D[m][j] = sum_r A[m][r] S[r][j]
// aTmem: A through tensor memory instead of shared memory.
__global__ void __launch_bounds__(T) k_test(const float* __restrict__ A, const float* __restrict__ S, float* __restrict__ D,
                                            int R, int Kk, int bMN, int aTmem, int* fail) {
  extern __shared__ __align__(128) unsigned char smraw[];
  const int K = bMN ? R : Kk, N = bMN ? Kk : R;
  float4* Sb = reinterpret_cast<float4*>(smraw);                   // [Kk/4][R]
  float4* As = Sb + (size_t)(Kk / 4) * R;                          // [K/4][128]
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < (Kk / 4) * R; i += T) { const int c = i / R, r = i - c * R; Sb[i] = *reinterpret_cast<const float4*>(S + (size_t)r * Kk + 4 * c); }
  for (int i = tid; i < (K / 4) * 128; i += T) { const int c = i / 128, r = i - c * 128; As[i] = *reinterpret_cast<const float4*>(A + (size_t)r * K + 4 * c); }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s, tD = tmem, tA = tmem + 256;
  if (aTmem) {           // row m = warp*32 + lane -> TMEM lane m, columns tA + k
    const int m = warp * 32 + lane;
    for (int k0 = 0; k0 < K; k0 += 8) {
      uint32_t v[8];
      for (int j = 0; j < 8; ++j) v[j] = __float_as_uint(A[(size_t)m * K + k0 + j]);
      asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                   ::"r"(tA + ((uint32_t)(warp * 32) << 16) + k0), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }
  if (tid == 0) {
    const uint32_t idesc = idesc_tf32(128, N, bMN);
    for (int kk = 0; kk < K / 8; ++kk) {
      // K-major: two k-chunks (LBO apart) per MMA; MN-major: 8 contraction rows = 8 consecutive float4 (128 B) per MMA
      const uint64_t db = bMN ? umma_desc(Sb + (size_t)kk * 8, 128, R * 16) : umma_desc(Sb + (size_t)(2 * kk) * R, R * 16, 128);
      if (aTmem) umma_ts(tD, tA + kk * 8, db, idesc, kk > 0);
      else umma_ss(tD, umma_desc(As + (size_t)(2 * kk) * 128, 128 * 16, 128), db, idesc, kk > 0);
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  uint32_t ok = 0;
  for (int spin = 0; spin < (1 << 22) && !ok; ++spin)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
  if (!ok) { if (tid == 0) *fail = 1; }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (ok) {
    const int row = warp * 32 + lane;
    for (int c0 = 0; c0 < N; c0 += 8) {
      uint32_t v[8];
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                   : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                   : "r"(tD + ((uint32_t)(warp * 32) << 16) + c0));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int j = 0; j < 8; ++j) D[(size_t)row * N + c0 + j] = __uint_as_float(v[j]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

static int run(const char* name, int R, int Kk, int bMN, int aTmem) {
  const int K = bMN ? R : Kk, N = bMN ? Kk : R;
  std::vector<float> A((size_t)128 * K), S((size_t)R * Kk);
  unsigned s = 12345u + R * 7 + Kk * 13 + bMN;
  auto rnd = [&]() { s = s * 1664525u + 1013904223u; return (float)((int)((s >> 20) & 31) - 16) / 8.0f; };
  for (auto& x : A) x = rnd();
  for (auto& x : S) x = rnd();
  float *dA, *dS, *dD; int* dF;
  cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dS, S.size() * 4); cudaMalloc(&dD, (size_t)128 * N * 4); cudaMalloc(&dF, 4);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(dS, S.data(), S.size() * 4, cudaMemcpyHostToDevice);
  cudaMemset(dD, 0xff, (size_t)128 * N * 4); cudaMemset(dF, 0, 4);
  const size_t smem = (size_t)(Kk / 4) * R * 16 + (size_t)(K / 4) * 128 * 16;
  cudaFuncSetAttribute(k_test, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  k_test<<<1, T, smem>>>(dA, dS, dD, R, Kk, bMN, aTmem, dF);
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  std::vector<float> D((size_t)128 * N); int fail = 0;
  cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost); cudaMemcpy(&fail, dF, 4, cudaMemcpyDeviceToHost);
  double err = 0; int bad = 0;
  for (int i = 0; i < 128; ++i) for (int j = 0; j < N; ++j) {
    double ref = 0;
    for (int k = 0; k < K; ++k) ref += (double)A[(size_t)i * K + k] * (bMN ? S[(size_t)k * Kk + j] : S[(size_t)j * Kk + k]);
    const double d = fabs(ref - D[(size_t)i * N + j]);
    if (!(d == 0)) ++bad;
    if (d == d) err = fmax(err, d); else err = 1e30;
  }
  printf("%-44s M=128 N=%3d K=%3d  max err %.3e  mismatches %d / %d  %s%s  => %s\n", name, N, K, err, bad, 128 * N,
         cudaGetErrorString(e), fail ? "  WAIT TIMED OUT" : "", (bad == 0 && e == cudaSuccess && !fail) ? "PASS" : "FAIL");
  cudaFree(dA); cudaFree(dS); cudaFree(dD); cudaFree(dF);
  return bad == 0 && e == cudaSuccess && !fail;
}

int main() {
  int ok = 1;
  ok &= run("SS  A smem K-major, B K-major", 128, 32, 0, 0);
  ok &= run("TS  A TMEM,         B K-major", 128, 32, 0, 1);
  ok &= run("TS  A TMEM,         B K-major (K=128)", 128, 128, 0, 1);
  ok &= run("TS  A TMEM,         B K-major (N=16)", 16, 128, 0, 1);
  ok &= run("SS  A smem K-major, B MN-major", 128, 32, 1, 0);
  ok &= run("TS  A TMEM,         B MN-major", 128, 32, 1, 1);
  ok &= run("TS  A TMEM,         B MN-major (W2: 128x128)", 128, 128, 1, 1);
  ok &= run("TS  A TMEM,         B MN-major (W3: K=16)", 16, 128, 1, 1);
  printf(ok ? "ALL PASS\n" : "SOME FAILED\n");
  return 0;
}
