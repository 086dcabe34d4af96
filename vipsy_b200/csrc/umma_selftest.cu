// Diagnostic: one 128-row tcgen05 tile product in each operand orientation the MIRT tensor-core
// kernel (mirt_tc.cu) uses, through the same helpers (umma.cuh), with the fp16 hi/lo split.
//   variant 0  D[m][n] = sum_k T1[m][k] T2[k][n]   A K-major,  B MN-major   (logits = x a)
//   variant 1  D[m][n] = sum_k T1[m][k] T2[n][k]   A K-major,  B K-major    (d x = G a^T)
//   variant 2  D[m][n] = sum_k T1[k][m] T2[k][n]   A MN-major, B MN-major   (d a = x^T G)
//   variant 3  as 0, variant 4 as 1, with A in tensor memory (tcgen05.st by the thread that owns the row): the
//              operand path of the amortised-guide kernels (enc_mvn.cu)
// T1 is R1 x C1, T2 is R2 x C2 (fp32, row-major, dimensions multiples of 8 / 16 as the MMA needs).
#include "common.cuh"
#include "umma.cuh"

namespace vipsy {

constexpr int kStThreads = 128;

__global__ void __launch_bounds__(kStThreads) umma_selftest_kernel(const float* __restrict__ T1, int R1, int C1,
                                                                   const float* __restrict__ T2, int R2, int C2,
                                                                   float* __restrict__ D, int variant, int swap,
                                                                   int n_split) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_sm;
  const uint32_t CS = 128;
  const uint32_t RS1 = (uint32_t)(C1 / 8) * CS, RS2 = (uint32_t)(C2 / 8) * CS;
  const uint32_t bytes1 = (uint32_t)(R1 / 8) * RS1, bytes2 = (uint32_t)(R2 / 8) * RS2;
  unsigned char* t1h = smem;
  unsigned char* t1l = t1h + bytes1;
  unsigned char* t2h = t1l + bytes1;
  unsigned char* t2l = t2h + bytes2;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  for (int e = tid; e < R1 * C1; e += kStThreads) {
    const int r = e / C1, c = e - r * C1;
    __half hi, lo;
    umma::split_f16(T1[e], hi, lo);
    const uint32_t o = umma::tile_off(r, c, RS1, CS);
    *reinterpret_cast<__half*>(t1h + o) = hi;
    *reinterpret_cast<__half*>(t1l + o) = lo;
  }
  for (int e = tid; e < R2 * C2; e += kStThreads) {
    const int r = e / C2, c = e - r * C2;
    __half hi, lo;
    umma::split_f16(T2[e], hi, lo);
    const uint32_t o = umma::tile_off(r, c, RS2, CS);
    *reinterpret_cast<__half*>(t2h + o) = hi;
    *reinterpret_cast<__half*>(t2l + o) = lo;
  }
  if (warp == 0) umma::tmem_alloc(&tmem_base_sm, 256);
  if (tid == 0) {
    umma::mbar_init(&bar, 1);
    umma::fence_mbar_init();
  }
  umma::fence_proxy_async();
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem = tmem_base_sm;
  const bool a_tmem = variant >= 3;
  if (a_tmem) variant -= 3;

  int N, ksteps;
  if (variant == 0) { N = C2; ksteps = C1 / 16; }
  else if (variant == 1) { N = R2; ksteps = C1 / 16; }
  else { N = C2; ksteps = R1 / 16; }

  if (a_tmem) {                                          // row tid of T1 -> lane tid: hi words at column 128, lo at 192
    const uint32_t tl = tmem + ((uint32_t)(warp * 32) << 16);
    for (int c0 = 0; c0 < C1; c0 += 16) {
      uint32_t h[8], l[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        __half h0, l0, h1, l1;
        umma::split_f16(tid < R1 ? T1[(size_t)tid * C1 + c0 + 2 * q] : 0.f, h0, l0);
        umma::split_f16(tid < R1 ? T1[(size_t)tid * C1 + c0 + 2 * q + 1] : 0.f, h1, l1);
        h[q] = (uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16);
        l[q] = (uint32_t)__half_as_ushort(l0) | ((uint32_t)__half_as_ushort(l1) << 16);
      }
      umma::tmem_st8(tl + 128 + c0 / 2, h);
      umma::tmem_st8(tl + 192 + c0 / 2, l);
    }
    umma::tmem_st_wait();
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
  }

  if (tid == 0) {
    const uint32_t a_major = variant == 2 ? umma::kMajorMN : umma::kMajorK;
    const uint32_t b_major = variant == 1 ? umma::kMajorK : umma::kMajorMN;
    const uint32_t idesc = umma::make_idesc_f16(128, N, a_major, b_major);
    uint32_t acc = 0;
    for (int ks = 0; ks < ksteps; ++ks) {
      for (int term = 0; term < n_split; ++term) {       // hi*hi, hi*lo, lo*hi
        const unsigned char* ta = term == 2 ? t1l : t1h;
        const unsigned char* tb = term == 1 ? t2l : t2h;
        uint64_t da, db;
        if (a_major == umma::kMajorK) {
          da = umma::make_desc(umma::smem_u32(ta) + ks * 2 * CS, CS, RS1);
        } else {
          const uint32_t lbo = swap ? CS : RS1, sbo = swap ? RS1 : CS;
          da = umma::make_desc(umma::smem_u32(ta) + ks * 2 * RS1, lbo, sbo);
        }
        if (b_major == umma::kMajorK) {
          db = umma::make_desc(umma::smem_u32(tb) + ks * 2 * CS, CS, RS2);
        } else {
          const uint32_t lbo = swap ? CS : RS2, sbo = swap ? RS2 : CS;
          db = umma::make_desc(umma::smem_u32(tb) + ks * 2 * RS2, lbo, sbo);
        }
        if (a_tmem) umma::mma_f16_ts(tmem, tmem + (term == 2 ? 192 : 128) + ks * 8, db, idesc, acc);
        else umma::mma_f16_ss(tmem, da, db, idesc, acc);
        acc = 1;
      }
    }
    umma::commit(&bar);
  }
  umma::mbar_wait(&bar, 0);
  umma::fence_after_sync();
  for (int c0 = 0; c0 < N; c0 += 16) {
    float v[16];
    umma::tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
    const int row = warp * 32 + lane;
#pragma unroll
    for (int i = 0; i < 16; ++i) D[(size_t)row * N + c0 + i] = v[i];
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem, 256);
}

int umma_selftest_launch(const float* t1, int R1, int C1, const float* t2, int R2, int C2, float* out, int variant,
                         int swap, int n_split, cudaStream_t st) {
  VIPSY_REQUIRE(t1 && t2 && out, VIPSY_ERR_ARG, "vipsy_selftest_umma: null pointer");
  VIPSY_REQUIRE(variant >= 0 && variant <= 4 && n_split >= 1 && n_split <= 3, VIPSY_ERR_ARG,
                "vipsy_selftest_umma: variant in 0..4, n_split in 1..3");
  VIPSY_REQUIRE(R1 % 8 == 0 && C1 % 8 == 0 && R2 % 8 == 0 && C2 % 8 == 0 && R1 <= 128 && C1 <= 128 && R2 <= 128 &&
                    C2 <= 128,
                VIPSY_ERR_ARG, "vipsy_selftest_umma: tile dimensions must be multiples of 8, at most 128");
  const size_t smem = 2 * ((size_t)R1 * C1 + (size_t)R2 * C2) * 2 + 4096;
  VIPSY_CUDA_CHECK(cudaFuncSetAttribute(umma_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  umma_selftest_kernel<<<1, kStThreads, smem, st>>>(t1, R1, C1, t2, R2, C2, out, variant, swap, n_split);
  VIPSY_CUDA_CHECK(cudaGetLastError());
  return VIPSY_OK;
}

}  // namespace vipsy
