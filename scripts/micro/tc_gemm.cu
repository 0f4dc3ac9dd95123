// scripts/micro/tc_gemm.cu — first tcgen05 / TMEM kernel of this repository: the building block of the
// weight-gradient contraction of recurrent nets,  D[M=128][N=64] (f32) = A[128][K] * B[64][K]^T,  contraction over the
// K = (sample, window step) columns, operands in shared memory, accumulator in tensor memory.
//
//   * operands: K-major, UMMA "no swizzle" canonical layout (cute/atom/mma_traits_sm100.hpp:192-199,273-303): in units
//     of 16 bytes ((8,n),2):((1,SBO),LBO).  Stored here as smem[kchunk][row] (one float4 = 4 consecutive k of one row),
//     i.e. SBO = 128 B (next 8-row group), LBO = rows*16 B (next k-chunk); one MMA consumes 2 k-chunks (UMMA_K = 8 tf32).
//   * instruction: tcgen05.mma.cta_group::1.kind::tf32, M = 128, N = 64, issued by ONE thread, accumulate flag per call;
//     3xTF32 variant: A = Ah + Al, B = Bh + Bl split on the way into shared memory, Al*Bh + Ah*Bl + Ah*Bh per k-step.
//   * completion: tcgen05.commit -> mbarrier; epilogue: tcgen05.ld.32x32b (warp w owns TMEM lanes 32w..32w+31).
// Every wait is bounded; a failure prints instead of hanging the box.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/micro/tc_gemm scripts/micro/tc_gemm.cu && ./scripts/micro/tc_gemm
#include <cstdint>
#include <cstdio>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>

constexpr int M = 128, N = 64, T = 128;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// UMMA shared-memory descriptor (cute/arch/mma_sm100_desc.hpp:98-123): start address, leading / stride byte offsets
// in 16-byte units, version 1 (Blackwell), no swizzle.
__device__ __forceinline__ uint64_t umma_desc(const void* base, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_u32(base) >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;                       // version_
  return d;                                     // base_offset 0, lbo_mode 0, layout_type 0 = SWIZZLE_NONE
}
// instruction descriptor (mma_sm100_desc.hpp:412-439): D f32, A/B tf32, both K-major, N>>3, M>>4
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int m, int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (0u << 15) | (0u << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
               ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ float tf32_hi(float x) { uint32_t u; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x)); return __uint_as_float(u); }

template <bool SPLIT>
__global__ void __launch_bounds__(T) tc_gemm(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D, int K,
                                             long long* cycles, int* fail) {
  extern __shared__ __align__(128) unsigned char smraw[];
  const int KC = K / 4;                                            // k-chunks of 16 bytes
  float4* Ah = reinterpret_cast<float4*>(smraw);                   // [KC][M]
  float4* Bh = Ah + (size_t)KC * M;                                // [KC][N]
  float4* Al = Bh + (size_t)KC * N;                                // SPLIT only
  float4* Bl = Al + (size_t)KC * M;
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  for (int i = tid; i < KC * M; i += T) {
    const int c = i / M, r = i - c * M;
    const float4 v = *reinterpret_cast<const float4*>(A + (size_t)r * K + 4 * c);
    if (SPLIT) {
      const float4 h = make_float4(tf32_hi(v.x), tf32_hi(v.y), tf32_hi(v.z), tf32_hi(v.w));
      Ah[i] = h; Al[i] = make_float4(v.x - h.x, v.y - h.y, v.z - h.z, v.w - h.w);
    } else Ah[i] = v;
  }
  for (int i = tid; i < KC * N; i += T) {
    const int c = i / N, r = i - c * N;
    const float4 v = *reinterpret_cast<const float4*>(B + (size_t)r * K + 4 * c);
    if (SPLIT) {
      const float4 h = make_float4(tf32_hi(v.x), tf32_hi(v.y), tf32_hi(v.z), tf32_hi(v.w));
      Bh[i] = h; Bl[i] = make_float4(v.x - h.x, v.y - h.y, v.z - h.z, v.w - h.w);
    } else Bh[i] = v;
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {      // one warp allocates 64 TMEM columns (128 lanes x 64 x 32 bit = the 128 x 64 f32 accumulator)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(64u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // generic-proxy writes of the operands -> async proxy reads of the tensor core
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_d = tmem_base_s;

  const long long t0 = clock64();
  if (tid == 0) {
    const uint32_t idesc = umma_idesc_tf32(M, N);
    for (int kk = 0; kk < K / 8; ++kk) {
      const uint64_t dah = umma_desc(Ah + (size_t)(2 * kk) * M, M * 16, 128), dbh = umma_desc(Bh + (size_t)(2 * kk) * N, N * 16, 128);
      if (SPLIT) {
        const uint64_t dal = umma_desc(Al + (size_t)(2 * kk) * M, M * 16, 128), dbl = umma_desc(Bl + (size_t)(2 * kk) * N, N * 16, 128);
        umma_tf32(tmem_d, dal, dbh, idesc, kk > 0);
        umma_tf32(tmem_d, dah, dbl, idesc, 1);
        umma_tf32(tmem_d, dah, dbh, idesc, 1);
      } else umma_tf32(tmem_d, dah, dbh, idesc, kk > 0);
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  // bounded wait for the accumulator
  uint32_t ok = 0;
  for (int spin = 0; spin < (1 << 22) && !ok; ++spin)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
  const long long t1 = clock64();
  if (!ok) { if (tid == 0) *fail = 1; }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (ok) {             // epilogue: warp w reads its 32 TMEM lanes (rows 32w .. 32w+31), 64 columns
    uint32_t v[64];
    const uint32_t taddr = tmem_d + ((uint32_t)(warp * 32) << 16);
#define LD32(off) asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 " \
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];" \
      : "=r"(v[off+0]),"=r"(v[off+1]),"=r"(v[off+2]),"=r"(v[off+3]),"=r"(v[off+4]),"=r"(v[off+5]),"=r"(v[off+6]),"=r"(v[off+7]), \
        "=r"(v[off+8]),"=r"(v[off+9]),"=r"(v[off+10]),"=r"(v[off+11]),"=r"(v[off+12]),"=r"(v[off+13]),"=r"(v[off+14]),"=r"(v[off+15]), \
        "=r"(v[off+16]),"=r"(v[off+17]),"=r"(v[off+18]),"=r"(v[off+19]),"=r"(v[off+20]),"=r"(v[off+21]),"=r"(v[off+22]),"=r"(v[off+23]), \
        "=r"(v[off+24]),"=r"(v[off+25]),"=r"(v[off+26]),"=r"(v[off+27]),"=r"(v[off+28]),"=r"(v[off+29]),"=r"(v[off+30]),"=r"(v[off+31]) \
      : "r"(taddr + off))
    LD32(0); LD32(32);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    const int row = warp * 32 + lane;
#pragma unroll
    for (int j = 0; j < 64; ++j) D[(size_t)row * N + j] = __uint_as_float(v[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(64u) : "memory");
  if (tid == 0) *cycles = t1 - t0;
}

template <bool SPLIT>
static void run(const char* name, int K) {
  std::vector<float> A((size_t)M * K), B((size_t)N * K);
  for (size_t i = 0; i < A.size(); ++i) A[i] = sinf(0.37f * (float)i) * 0.8f;
  for (size_t i = 0; i < B.size(); ++i) B[i] = cosf(0.11f * (float)i) * 0.6f;
  float *dA, *dB, *dD; long long* dC; int* dF;
  cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, (size_t)M * N * 4); cudaMalloc(&dC, 8); cudaMalloc(&dF, 4);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
  cudaMemset(dD, 0, (size_t)M * N * 4); cudaMemset(dF, 0, 4);
  const size_t smem = (size_t)(K / 4) * (M + N) * 16 * (SPLIT ? 2 : 1);
  cudaFuncSetAttribute(tc_gemm<SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  tc_gemm<SPLIT><<<1, T, smem>>>(dA, dB, dD, K, dC, dF);
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  std::vector<float> D((size_t)M * N); long long cyc = 0; int fail = 0;
  cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost); cudaMemcpy(&cyc, dC, 8, cudaMemcpyDeviceToHost); cudaMemcpy(&fail, dF, 4, cudaMemcpyDeviceToHost);
  double err = 0, ref_max = 0;
  for (int i = 0; i < M; ++i) for (int j = 0; j < N; ++j) {
    double s = 0; for (int k = 0; k < K; ++k) s += (double)A[(size_t)i * K + k] * B[(size_t)j * K + k];
    err = fmax(err, fabs(s - D[(size_t)i * N + j])); ref_max = fmax(ref_max, fabs(s));
  }
  const double flop = 2.0 * M * N * K * (SPLIT ? 3 : 1);
  printf("%-34s K=%4d  %7lld cycles (MMA issue -> accumulator ready)  %.1f flop/cycle/SM  max err %.3e (max |D| %.2f)  %s%s\n",
         name, K, cyc, cyc ? flop / cyc : 0.0, err, ref_max, cudaGetErrorString(e), fail ? "  WAIT TIMED OUT" : "");
  cudaFree(dA); cudaFree(dB); cudaFree(dD); cudaFree(dC); cudaFree(dF);
}

int main() {
  run<false>("tcgen05 kind::tf32 128x64", 64);
  run<false>("tcgen05 kind::tf32 128x64", 256);
  run<false>("tcgen05 kind::tf32 128x64", 128);
  run<true>("tcgen05 3xTF32 split 128x64", 128);
  return 0;
}
