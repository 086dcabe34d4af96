// Micro-benchmark: cycles per tcgen05.mma (kind::f16, M = 128, K = 16) issued back to back by one thread, by N and by
// where A lives (shared memory descriptor vs tensor memory), plus the cost of the mbarrier wait / commit pair around
// a group of G MMAs (the shape of the issue loops in enc_mvn.cu / mirt_tc.cu).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I vipsy_b200/csrc -o scripts/ubench/bin/mma_rate scripts/ubench/mma_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "umma.cuh"
using namespace vipsy;
__global__ void __launch_bounds__(128, 1) k(int N, int a_tmem, int G, int groups, int R, int uniform, long long* out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar[2];
  __shared__ uint32_t tbase;
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x < 32) umma::tmem_alloc(&tbase, 512);
  if (threadIdx.x == 0) { umma::mbar_init(&bar[0], 1); umma::mbar_init(&bar[1], 1); umma::fence_mbar_init(); }
  umma::fence_proxy_async();
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem = tbase;
  if (!uniform && threadIdx.x == 0) {
    const uint32_t idesc = umma::make_idesc_f16(128, N, umma::kMajorK, umma::kMajorK);
    const uint64_t da = umma::make_desc(umma::smem_u32(smem), 128, 256);
    const uint64_t db = umma::make_desc(umma::smem_u32(smem) + 16384, 128, 256);
    long long t0 = clock64();
    for (int g = 0; g < groups; ++g) {
      if (g >= 2) umma::mbar_wait(&bar[g & 1], ((g - 2) >> 1) & 1);        // the group two back has completed
      umma::fence_after_sync();
      for (int i = 0; i < G; ++i) {
        if (a_tmem) umma::mma_f16_ts(tmem + (i & (R - 1)) * 64, tmem + 384 + 8 * (i & 3), db + (uint64_t)(i & 3) * 2, idesc, 1u);
        else umma::mma_f16_ss(tmem + (i & (R - 1)) * 64, da + (uint64_t)(i & 3) * 2, db + (uint64_t)(i & 3) * 2, idesc, 1u);
      }
      umma::commit(&bar[g & 1]);
    }
    long long t1 = clock64();
    umma::mbar_wait(&bar[(groups - 1) & 1], ((groups - 1) >> 1) & 1);
    long long t2 = clock64();
    if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  if (uniform && threadIdx.x < 32) {                     // the whole warp walks the loop, one elected lane issues
    const uint32_t tmem = uniform == 2 ? __reduce_or_sync(0xffffffffu, tbase) : tbase;     // 2: provably warp-uniform
    const uint32_t idesc = umma::make_idesc_f16(128, N, umma::kMajorK, umma::kMajorK);
    const uint64_t da = umma::make_desc(umma::smem_u32(smem), 128, 256);
    const uint64_t db = umma::make_desc(umma::smem_u32(smem) + 16384, 128, 256);
    long long t0 = clock64();
    for (int g = 0; g < groups; ++g) {
      if (g >= 2) umma::mbar_wait(&bar[g & 1], ((g - 2) >> 1) & 1);
      umma::fence_after_sync();
      if (umma::elect_one()) {
        for (int i = 0; i < G; ++i) {
          if (a_tmem) umma::mma_f16_ts(tmem + (i & (R - 1)) * 64, tmem + 384 + 8 * (i & 3), db + (uint64_t)(i & 3) * 2, idesc, 1u);
          else umma::mma_f16_ss(tmem + (i & (R - 1)) * 64, da + (uint64_t)(i & 3) * 2, db + (uint64_t)(i & 3) * 2, idesc, 1u);
        }
        umma::commit(&bar[g & 1]);
      }
      __syncwarp();
    }
    long long t1 = clock64();
    umma::mbar_wait(&bar[(groups - 1) & 1], ((groups - 1) >> 1) & 1);
    long long t2 = clock64();
    if (blockIdx.x == 0 && threadIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  __syncthreads();
  if (threadIdx.x < 32) umma::tmem_dealloc(tmem, 512);
}
int main() {
  long long* out; long long h[2];
  cudaMalloc(&out, 16);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  for (int a_tmem : {0, 1})
   for (int uniform : {1, 2})
    for (int N : {64, 128})
     for (int R : {1, 4})
      for (int G : {12, 48}) {
        const int groups = 4800 / G;
        k<<<148, 128, 64 * 1024>>>(N, a_tmem, G, groups, R, uniform, out);
        cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
        printf("%s A in %s  N=%3d  accumulators %d  %2d MMAs per wait/commit group: %6.1f cycles per MMA issued, %6.1f to completion (tensor floor %d)\n",
               uniform == 2 ? "uniform + elect + redux" : uniform ? "warp-uniform + elect   " : "if (tid == 0)       ", a_tmem ? "TMEM" : "smem", N, R, G, (double)h[0] / (G * groups), (double)h[1] / (G * groups), N / 2);
      }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
