// scripts/micro/tc_mn_probe.cu — which shared-memory float does tcgen05.mma read as B[k][n] for an MN-major no-swizzle
// descriptor?  A = identity (through TMEM), the B image holds its own float index (split in two exact-in-tf32 passes), so
// D[k][n] = index of the float the tensor core used.  Prints the decoded byte offsets for a few (k, n) and LBO / SBO pairs.
#include <cstdint>
#include <cstdio>
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
__global__ void __launch_bounds__(T) k_probe(float* __restrict__ D, int N, int lbo, int sbo, int pass, int floats, int* fail, int bmn) {
  extern __shared__ __align__(128) unsigned char smraw[];
  float* Sb = reinterpret_cast<float*>(smraw);
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < floats; i += T) Sb[i] = (float)(pass ? (i >> 10) : (i & 1023));
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
  {   // A = identity on the first 8 rows: A[m][k] = (m == k), K = 8
    const int m = warp * 32 + lane;
    uint32_t v[8];
    for (int j = 0; j < 8; ++j) v[j] = __float_as_uint(m == j ? 1.0f : 0.0f);
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 ::"r"(tA + ((uint32_t)(warp * 32) << 16)), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }
  if (tid == 0) {
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)bmn << 16) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
    const uint64_t db = umma_desc(Sb, lbo, sbo);
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n"
                 ::"r"(tD), "r"(tA), "l"(db), "r"(idesc), "r"(0u) : "memory");
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
static void probe(int N, int lbo, int sbo, int bmn = 1) {
  const int floats = 16384;
  float* dD; int* dF; cudaMalloc(&dD, 128 * N * 4); cudaMalloc(&dF, 4); cudaMemset(dF, 0, 4);
  std::vector<float> D0(128 * N), D1(128 * N);
  cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, floats * 4);
  for (int pass = 0; pass < 2; ++pass) {
    k_probe<<<1, T, floats * 4>>>(dD, N, lbo, sbo, pass, floats, dF, bmn);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return; }
    cudaMemcpy(pass ? D1.data() : D0.data(), dD, 128 * N * 4, cudaMemcpyDeviceToHost);
  }
  int fail; cudaMemcpy(&fail, dF, 4, cudaMemcpyDeviceToHost);
  printf("bMN=%d N=%d LBO=%d SBO=%d%s: byte offset read for B[k][n]\n", bmn, N, lbo, sbo, fail ? " TIMEOUT" : "");
  for (int k = 0; k < 8; ++k) {
    printf("  k=%d:", k);
    for (int n = 0; n < N && n < 40; ++n) printf(" %5d", 4 * ((int)D0[k * N + n] + 1024 * (int)D1[k * N + n]));
    printf("\n");
  }
  cudaFree(dD); cudaFree(dF);
}
int main() {
  probe(32, 2048, 128, 0);
  probe(32, 128, 2048);
  probe(32, 2048, 128);
  probe(32, 4096, 1024);
  probe(16, 1024, 4096);
  return 0;
}
