// Shared-memory load throughput microbenchmark (one CTA of 512 threads on one SM):
// cycles per warp-wide LDS instruction for broadcast / consecutive, 32-bit / 128-bit accesses.
#include <cstdio>
#include <cuda_runtime.h>
constexpr int T = 512, IT = 2048;
template <int MODE>
__global__ void k(float* out, long long* cyc) {
  __shared__ __align__(16) float sm[8192];
  for (int i = threadIdx.x; i < 8192; i += T) sm[i] = (float)i;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float acc0 = 0, acc1 = 0, acc2 = 0, acc3 = 0;
  long long t0 = clock64();
#pragma unroll 8
  for (int i = 0; i < IT; ++i) {
    const int base = ((i * 37 + warp * 5) & 255) * 16;   // float index, multiple of 16
    if (MODE == 0) { const float4 v = *reinterpret_cast<const float4*>(sm + base); acc0 += v.x; acc1 += v.y; acc2 += v.z; acc3 += v.w; }          // LDS.128 broadcast
    if (MODE == 1) { const float v = sm[base + lane]; acc0 += v; }                                                                                 // LDS.32 consecutive
    if (MODE == 2) { const float4 v = *reinterpret_cast<const float4*>(sm + base + lane * 4); acc0 += v.x; acc1 += v.y; acc2 += v.z; acc3 += v.w; }  // LDS.128 consecutive (512 B)
    if (MODE == 3) { const float v = sm[base]; acc0 += v; }                                                                                        // LDS.32 broadcast
    if (MODE == 4) { const float2 v = *reinterpret_cast<const float2*>(sm + base + lane * 2); acc0 += v.x; acc1 += v.y; }                           // LDS.64 consecutive
    if (MODE == 5) { const float4 v = *reinterpret_cast<const float4*>(sm + base + (lane >> 3) * 4); acc0 += v.x; acc1 += v.y; acc2 += v.z; acc3 += v.w; }  // LDS.128, 4 distinct addresses
  }
  long long t1 = clock64();
  out[threadIdx.x] = acc0 + acc1 + acc2 + acc3;
  if (threadIdx.x == 0) *cyc = t1 - t0;
}
int main() {
  float* out; long long* cyc; cudaMalloc(&out, T * 4); cudaMalloc(&cyc, 8);
  const char* names[] = {"LDS.128 broadcast", "LDS.32 consecutive", "LDS.128 consecutive", "LDS.32 broadcast", "LDS.64 consecutive", "LDS.128 4 addrs"};
  for (int m = 0; m < 6; ++m) {
    long long h = 0;
    for (int rep = 0; rep < 2; ++rep) {
      switch (m) { case 0: k<0><<<1, T>>>(out, cyc); break; case 1: k<1><<<1, T>>>(out, cyc); break; case 2: k<2><<<1, T>>>(out, cyc); break;
                   case 3: k<3><<<1, T>>>(out, cyc); break; case 4: k<4><<<1, T>>>(out, cyc); break; case 5: k<5><<<1, T>>>(out, cyc); break; }
      cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    }
    printf("%-22s %7.2f cycles per warp-instruction (16 warps on one SM: SM-wide rate %.2f cycles/instr)\n", names[m], (double)h / IT, (double)h / IT / 16);
  }
  return 0;
}
