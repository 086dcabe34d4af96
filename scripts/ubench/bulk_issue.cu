// Where do the ~730 cycles per cp.async.bulk go in a one-thread issue loop?  clock64 around each piece.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/ubench/bin/bulk_issue scripts/ubench/bulk_issue.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void __launch_bounds__(64, 1) k(const unsigned char* img, uint32_t chunk, int nst, int iters, int mode, long long* out) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t full[16];
  if (threadIdx.x == 0) {
    for (int i = 0; i < 16; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&full[i])));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    long long tw = 0, te = 0, ti = 0;
    long long t0 = clock64();
    for (int it = 0; it < iters + nst; ++it) {
      const int st = it % nst;
      long long a = clock64();
      if (it >= nst) {
        const uint32_t par = ((it - nst) / nst) & 1;
        asm volatile("{ .reg .pred p; W: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1; @p bra D; bra W; D: }" ::"r"(
                         s32(&full[st])), "r"(par) : "memory");
      }
      long long b = clock64();
      if (it < iters) {
        const unsigned char* src = img + ((size_t)(it + blockIdx.x * 7) % 64) * chunk;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&full[st])), "r"(chunk) : "memory");
        long long c = clock64();
        if (mode == 0)
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                           s32(smem + (size_t)st * chunk)), "l"(src), "r"(chunk), "r"(s32(&full[st])) : "memory");
        else
          asm volatile("cp.async.bulk.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                           s32(smem + (size_t)st * chunk)), "l"(src), "r"(chunk), "r"(s32(&full[st])) : "memory");
        long long d = clock64();
        te += c - b; ti += d - c;
      }
      tw += b - a;
    }
    long long t1 = clock64();
    if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = tw; out[2] = te; out[3] = ti; }
  }
}
int main() {
  unsigned char* img; long long* out; long long h[4];
  cudaMalloc(&img, 64u << 20); cudaMemset(img, 1, 64u << 20); cudaMalloc(&out, 32);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int iters = 4000;
  for (int mode : {0, 1})
    for (uint32_t chunk : {4096u, 16384u})
      for (int nst : {2, 8}) {
        k<<<148, 64, nst * chunk>>>(img, chunk, nst, iters, mode, out);
        cudaMemcpy(h, out, 32, cudaMemcpyDeviceToHost);
        printf("%s chunk %5u stages %d: %5.0f cycles per copy = wait %5.0f + expect_tx %5.0f + issue %5.0f\n",
               mode ? "shared::cta    " : "shared::cluster", chunk, nst, (double)h[0] / iters, (double)h[1] / iters, (double)h[2] / iters,
               (double)h[3] / iters);
      }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
