// Micro-benchmark: MUFU throughput per SM (lanes / clk) for ex2, rcp, lg2, rsqrt, and mixes.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/mufu scripts/ubench/mufu.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int OP>
__device__ __forceinline__ float f(float x) {
  float y;
  if (OP == 0) asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  if (OP == 1) asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  if (OP == 2) asm volatile("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  if (OP == 3) asm volatile("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  if (OP == 4) asm volatile("fma.rn.f32 %0, %1, %1, %1;" : "=f"(y) : "f"(x));
  if (OP == 5) asm volatile("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  if (OP == 6) asm volatile("{ .reg .f32 t; rcp.approx.ftz.f32 t, %1; add.f32 %0, t, 0f3F800000; }" : "=f"(y) : "f"(x));
  if (OP == 7) asm volatile("{ .reg .f32 t; ex2.approx.ftz.f32 t, %1; add.f32 %0, t, 0fBF000000; }" : "=f"(y) : "f"(x));
  if (OP == 8) asm volatile("{ .reg .f32 t; sqrt.approx.ftz.f32 t, %1; add.f32 %0, t, 0f3F800000; }" : "=f"(y) : "f"(x));
  return y;
}
template <int OP, int OP2>
__global__ void k(float* out, int iters, long long* cyc) {
  float v[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = 1.0f + threadIdx.x * 1e-3f + i;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = (i & 1) ? f<OP2>(v[i]) : f<OP>(v[i]);
  }
  long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int OP, int OP2>
void run(const char* name) {
  float* out; long long* cyc; long long h;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
  const int iters = 20000;
  for (int warps = 4; warps <= 32; warps *= 2) {
    k<OP, OP2><<<148, warps * 32>>>(out, iters, cyc);
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    const double ops = (double)iters * 8 * warps * 32;   // per SM
    printf("%-12s warps/SM=%2d  %.2f lane-ops/clk/SM  (%.2f cycles per warp-instr per SMSP)\n", name, warps, ops / h,
           (double)h / (iters * 8.0 * warps / 4));
  }
}
int main() {
  run<0, 0>("ex2"); run<6, 6>("rcp(+add)"); run<7, 7>("ex2(+add)"); run<8, 8>("sqrt(+add)");
  run<6, 7>("rcp,ex2 mix");
  return 0;
}
