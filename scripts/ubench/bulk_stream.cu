// Micro-benchmark: how fast can every SM stream one small, L2-resident operand image into shared memory with
// cp.async.bulk (the fc22 operand image of the amortised-guide kernels: 1.5 - 1.9 MB read once per 128-person tile)?
// NW issuing warps (one elected lane each, own ring of `nst` stages and mbarriers), copies of `chunk` bytes.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/ubench/bin/bulk_stream scripts/ubench/bulk_stream.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void __launch_bounds__(128, 1) k(const unsigned char* img, size_t img_bytes, uint32_t chunk, int nst, int iters,
                                             int nw, int nl, long long* cyc) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t full[4][16];
  if (threadIdx.x == 0) {
    for (int w = 0; w < 4; ++w)
      for (int i = 0; i < 16; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&full[w][i])));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  __syncthreads();
  const int w = threadIdx.x >> 5;
  long long t0 = clock64();
  const int lane = threadIdx.x & 31;
  if (lane < nl && w < nw) {
    const size_t n_chunks = img_bytes / chunk;
    unsigned char* ring = smem + (size_t)w * nst * chunk;
    for (int it = 0; it < iters + nst; ++it) {
      const int st = it % nst;
      if (it >= nst) {        // wait for the copy issued nst iterations ago, then reuse the stage
        const uint32_t par = ((it - nst) / nst) & 1;
        asm volatile("{ .reg .pred p; W: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1; @p bra D; bra W; D: }" ::"r"(
                         s32(&full[w][st])), "r"(par) : "memory");
      }
      if (it < iters) {
        const unsigned char* src = img + ((size_t)(it * nw + w + blockIdx.x * 7) % n_chunks) * chunk;
        if (lane == 0)
          asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&full[w][st])), "r"(chunk) : "memory");
        __syncwarp((1u << nl) - 1);
        const uint32_t piece = chunk / nl;
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         s32(ring + (size_t)st * chunk + lane * piece)), "l"(src + lane * piece), "r"(piece), "r"(s32(&full[w][st])) : "memory");
      }
    }
  }
  __syncthreads();
  long long t1 = clock64();
  if (blockIdx.x == 0 && threadIdx.x == 0) *cyc = t1 - t0;
}
int main() {
  unsigned char* img; long long* cyc; long long h;
  const size_t big = 64u << 20;
  cudaMalloc(&img, big); cudaMemset(img, 1, big); cudaMalloc(&cyc, 8);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int iters = 4000;
  const size_t img_bytes = (size_t)1536 << 10;
  for (uint32_t chunk : {4096u, 16384u, 32768u}) {
    for (int nw : {1, 2})
      for (int nl : {1, 2, 4, 8})
        for (int nst : {2, 4}) {
          if ((size_t)nw * nst * chunk > 200 * 1024) continue;
          const int grid = 148;
          k<<<grid, 128, nw * nst * chunk>>>(img, img_bytes, chunk, nst, iters, nw, nl, cyc);
          cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
          printf("chunk %5u B issuing warps %d lanes %d stages %d CTAs %3d: %6.1f B/clk/SM  (%.0f cycles per chunk per warp)\n", chunk,
                 nw, nl, nst, grid, (double)iters * nw * chunk / h, (double)h / iters);
        }
  }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
