// NormEncoder products on the 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM), J <= 127, H <= 64.
//
// The response codes are exactly representable in bf16 (0, 1, -1); the fp32 factor of each product (fc1.weight
// forward, d loss / d pre backward) is split into three bf16 terms v = hi + mid + lo (24 mantissa bits, fp32
// exponent range), so three MMAs per 16-deep k-step reproduce the fp32 product and the fp32 accumulator in TMEM
// adds them -- same numerics class as the SIMT kernels (enc_norm.cu), a tenth of the issue slots.
//
//   forward : pre[128 persons x 64] = X[128 x JP] . W1^T        A = code tile (K-major), B = W1 tiles (K-major),
//             epilogue TMEM -> registers: bias, store pre, softplus, the two 64 -> 1 heads
//   backward: D[128 x 64] += X^T[JP128 x 128 persons] . dpre    A = code tile (MN-major, + a ones column -> d b1),
//             B = dpre tiles (MN-major); D stays in TMEM over every tile of the persistent CTA and is read once
//
// Operand tiles use the no-swizzle core-matrix layout of umma.cuh (verified by tests/test_umma_selftest.py in
// the same orientations: variant 1 = forward, variant 2 = backward).
#include <cuda_bf16.h>

#include "enc_norm.cuh"
#include "umma.cuh"

namespace vipsy {

constexpr int kTB = 128;                  // persons per tile = UMMA M (forward) / K extent (backward)
constexpr int kH = 64;                    // hidden slots = UMMA N
constexpr int kThreads = 256;
constexpr uint32_t kCS = 128;             // bytes between core matrices along the contiguous dimension
constexpr uint32_t kTmemCols = 64;
constexpr int kBwdThreads = 512;          // backward: 4 threads per person (16 hidden units each)

bool enc_tc_supported(int H, int J) { return H >= 1 && H <= kH && J >= 1 && J <= 127; }

__device__ __forceinline__ uint32_t bf16_bits(float v) {
  return (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(v));
}
// v = hi + mid + lo, each a bf16 (exact: every residual is representable in fp32)
__device__ __forceinline__ void split3(float v, uint32_t& hi, uint32_t& mid, uint32_t& lo) {
  const __nv_bfloat16 h = __float2bfloat16_rn(v);
  const float r1 = v - __bfloat162float(h);
  const __nv_bfloat16 m = __float2bfloat16_rn(r1);
  const float r2 = r1 - __bfloat162float(m);
  hi = __bfloat16_as_ushort(h); mid = __bfloat16_as_ushort(m); lo = bf16_bits(r2);
}
// two response codes (0 / 1 / 255) -> two bf16 (0, 1, -1) in one word
__device__ __forceinline__ uint32_t code_bf16(uint32_t c) { return (c & 1u) * 0x3f80u | (c & 0x80u) << 8; }
__device__ __forceinline__ uint32_t codes_bf16x2(uint32_t word, int k) {
  return code_bf16((word >> (16 * k)) & 255u) | code_bf16((word >> (16 * k + 8)) & 255u) << 16;
}

// Codes of person (tile0 + p), items [8 j8, 8 j8 + 8) for this thread's chunks j8 = half + 2 q -> registers.
// `ones_col` >= 0: that item position reads as code 1 for valid persons (the bias column of the backward product).
// NT threads share a person (16-byte chunks j8 = (threadIdx.x >> 7) + NT q).
template <int NT>
__device__ __forceinline__ void tc_load_codes(uint32_t (&wd)[16 / NT][2], const EncParams& P, int64_t tile0, int n_chunks,
                                              bool aligned) {
  constexpr int kNQ = 16 / NT;
  const int p = threadIdx.x & (kTB - 1), half = threadIdx.x >> 7;
  const int64_t r = tile0 + p;
  const bool ok = r < P.B;
  const uint8_t* g = ok ? P.y + (P.idx ? __ldg(P.idx + r) : r) * P.ld : P.y;
#pragma unroll
  for (int q = 0; q < kNQ; ++q) {
    const int j8 = half + NT * q;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int j = 8 * j8 + 4 * k;
      uint32_t word = 0;
      if (ok && j8 < n_chunks && j < P.J) {
        if (aligned && j + 3 < P.J) {
          word = __ldg(reinterpret_cast<const uint32_t*>(g + j));
        } else {
#pragma unroll
          for (int b = 0; b < 4; ++b)
            if (j + b < P.J) word |= (uint32_t)__ldg(g + j + b) << (8 * b);
        }
      }
      wd[q][k] = word;
    }
  }
}
template <int NT>
__device__ __forceinline__ void tc_store_x(unsigned char* Xt, uint32_t RS, const uint32_t (&wd)[16 / NT][2], int n_chunks,
                                           int ones_col, bool person_ok) {
  constexpr int kNQ = 16 / NT;
  const int p = threadIdx.x & (kTB - 1), half = threadIdx.x >> 7;
#pragma unroll
  for (int q = 0; q < kNQ; ++q) {
    const int j8 = half + NT * q;
    if (j8 < n_chunks) {
      uint4 v = make_uint4(codes_bf16x2(wd[q][0], 0), codes_bf16x2(wd[q][0], 1), codes_bf16x2(wd[q][1], 0),
                           codes_bf16x2(wd[q][1], 1));
      if (ones_col >= 0 && (ones_col >> 3) == j8 && person_ok) {
        const uint32_t one = 0x3f80u << (16 * (ones_col & 1));
        const int wsel = (ones_col & 7) >> 1;
        if (wsel == 0) v.x |= one; else if (wsel == 1) v.y |= one; else if (wsel == 2) v.z |= one; else v.w |= one;
      }
      *reinterpret_cast<uint4*>(Xt + (uint32_t)(p >> 3) * RS + (uint32_t)j8 * kCS + (uint32_t)(p & 7) * 16u) = v;
    }
  }
}

// ------------------------------------------------------------------------------------------------ forward
__global__ void __launch_bounds__(kThreads, 2) enc_tc_fwd_kernel(const EncParams P, const int JP) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_sm;
  __shared__ float hv[3 * kH];                    // b1 | w21 | w22, zero beyond H
  __shared__ float2 part[2][kTB];                 // per column half: partial (mu, log sigma) of each person
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int n_chunks = JP >> 3;
  const uint32_t RS = (uint32_t)n_chunks * kCS;   // same pitch for the code tile (16 row blocks) and W tiles (8)
  unsigned char* Xt = smem;
  unsigned char* Wt[3] = {Xt + 16 * RS, Xt + 24 * RS, Xt + 32 * RS};
  const bool aligned = ((reinterpret_cast<uintptr_t>(P.y) | (uintptr_t)P.ld) & 3) == 0;

  for (int e = t; e < kH * JP; e += kThreads) {   // fc1.weight -> three bf16 operand tiles, once per CTA
    const int h = e / JP, j = e - h * JP;
    const float v = (h < P.H && j < P.J) ? __ldg(P.w1 + (size_t)h * P.J + j) : 0.0f;
    uint32_t a, b, c;
    split3(v, a, b, c);
    const uint32_t o = umma::tile_off(h, j, RS, kCS);
    *reinterpret_cast<uint16_t*>(Wt[0] + o) = (uint16_t)a;
    *reinterpret_cast<uint16_t*>(Wt[1] + o) = (uint16_t)b;
    *reinterpret_cast<uint16_t*>(Wt[2] + o) = (uint16_t)c;
  }
  if (t < 3 * kH) {
    const int h = t & 63, which = t >> 6;
    const float* src = which == 0 ? P.b1 : which == 1 ? P.w21 : P.w22;
    hv[t] = h < P.H ? __ldg(src + h) : 0.0f;
  }
  if (warp == 0) umma::tmem_alloc(&tmem_base_sm, kTmemCols);
  if (t == 0) {
    umma::mbar_init(&bar, 1);
    umma::fence_mbar_init();
  }
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem = tmem_base_sm;
  const uint32_t idesc = umma::make_idesc_bf16(kTB, kH, umma::kMajorK, umma::kMajorK);
  const uint32_t sX = umma::smem_u32(Xt);
  const uint32_t sW[3] = {umma::smem_u32(Wt[0]), umma::smem_u32(Wt[1]), umma::smem_u32(Wt[2])};
  const float b21 = __ldg(P.b21), b22 = __ldg(P.b22);
  const int quad = warp & 3, half = warp >> 2;    // TMEM lane quadrant, column half of this warp
  const int64_t n_tiles = (P.B + kTB - 1) / kTB;
  uint32_t wd[8][2];
  if ((int64_t)blockIdx.x < n_tiles) tc_load_codes<2>(wd, P, (int64_t)blockIdx.x * kTB, n_chunks, aligned);
  uint32_t phase = 0;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t tile0 = tile * kTB;
    tc_store_x<2>(Xt, RS, wd, n_chunks, -1, true);
    umma::fence_proxy_async();
    __syncthreads();
    if (t == 0) {
      umma::fence_after_sync();
      uint32_t acc = 0;
      for (int ks = 0; ks < JP / 16; ++ks) {
        const uint64_t da = umma::make_desc(sX + ks * 2 * kCS, kCS, RS);
#pragma unroll
        for (int term = 0; term < 3; ++term) {
          umma::mma_f16_ss(tmem, da, umma::make_desc(sW[term] + ks * 2 * kCS, kCS, RS), idesc, acc);
          acc = 1;
        }
      }
      umma::commit(&bar);
    }
    if (tile + gridDim.x < n_tiles) tc_load_codes<2>(wd, P, (tile + gridDim.x) * kTB, n_chunks, aligned);
    umma::mbar_wait(&bar, phase);
    phase ^= 1u;
    umma::fence_after_sync();
    {
      float v[32];
      umma::tmem_ld32(tmem + ((uint32_t)(quad * 32) << 16) + (uint32_t)(half * 32), v);
      const int p = quad * 32 + lane;
      const int64_t r = tile0 + p;
      const int h0 = half * 32;
      float m = 0.0f, l = 0.0f;
#pragma unroll
      for (int c = 0; c < 32; ++c) {
        v[c] += hv[h0 + c];
        float sp, sg;
        softplus_parts(v[c], sp, sg);
        m = fmaf(sp, hv[kH + h0 + c], m);           // weights are 0 beyond H
        l = fmaf(sp, hv[2 * kH + h0 + c], l);
      }
      if (r < P.B) {
        float* dst = P.pre + (size_t)r * P.H + h0;
        if ((P.H & 3) == 0) {
#pragma unroll
          for (int c = 0; c < 32; c += 4)
            if (h0 + c < P.H) *reinterpret_cast<float4*>(dst + c) = make_float4(v[c], v[c + 1], v[c + 2], v[c + 3]);
        } else {
#pragma unroll
          for (int c = 0; c < 32; ++c)
            if (h0 + c < P.H) dst[c] = v[c];
        }
      }
      part[half][p] = make_float2(m, l);
    }
    umma::fence_before_sync();
    __syncthreads();                                // TMEM and the code tile are free; both halves are in `part`
    if (t < kTB && tile0 + t < P.B) {
      const float2 a = part[0][t], b = part[1][t];
      P.mu[tile0 + t] = (a.x + b.x) + b21;
      P.sigma[tile0 + t] = expf((a.y + b.y) + b22);
    }
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem, kTmemCols);
}

// ------------------------------------------------------------------------------------------------ backward
__global__ void __launch_bounds__(kBwdThreads) enc_tc_bwd_kernel(const EncParams P) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_sm;
  __shared__ float gm[kTB], gl[kTB];              // d loss / d mu, d loss / d log sigma of the tile's persons
  __shared__ float red[2][4][kH];                 // end of kernel: [w21 | w22][warp of the half][h]
  __shared__ float wh[2][kH];                     // fc21.weight | fc22.weight, zero beyond H
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  constexpr uint32_t RSx = (kTB / 8) * kCS;       // code tile: 128 persons x 128 item slots (item J = ones column)
  constexpr uint32_t RSd = (kH / 8) * kCS;        // dpre tiles: 128 persons x 64 hidden
  unsigned char* Xt = smem;
  unsigned char* Dt[3] = {Xt + 16 * RSx, Xt + 16 * RSx + 16 * RSd, Xt + 16 * RSx + 32 * RSd};
  const bool aligned = ((reinterpret_cast<uintptr_t>(P.y) | (uintptr_t)P.ld) & 3) == 0;
  const int p = t & (kTB - 1), hq = t >> 7;       // this thread: person p, hidden chunks hc = hq + 4 q (8 units each)

  if (t < 2 * kH) {
    const int h = t & 63;
    wh[t >> 6][h] = h < P.H ? __ldg((t >> 6 ? P.w22 : P.w21) + h) : 0.0f;
  }
  if (warp == 0) umma::tmem_alloc(&tmem_base_sm, kTmemCols);
  if (t == 0) {
    umma::mbar_init(&bar, 1);
    umma::fence_mbar_init();
  }
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem = tmem_base_sm;
  const uint32_t idesc = umma::make_idesc_bf16(kTB, kH, umma::kMajorMN, umma::kMajorMN);
  const uint32_t sX = umma::smem_u32(Xt);
  const uint32_t sD[3] = {umma::smem_u32(Dt[0]), umma::smem_u32(Dt[1]), umma::smem_u32(Dt[2])};

  float s21[2][8], s22[2][8];
#pragma unroll
  for (int q = 0; q < 2; ++q)
#pragma unroll
    for (int i = 0; i < 8; ++i) s21[q][i] = s22[q][i] = 0.0f;
  float s_gm = 0.0f, s_gl = 0.0f;
  const int64_t n_tiles = (P.B + kTB - 1) / kTB;
  uint32_t wd[4][2];
  float n_dm = 0.0f, n_ds = 0.0f, n_sg = 0.0f;
  auto load_heads = [&](int64_t tile0) {
    const int64_t r = tile0 + p;
    n_dm = n_ds = n_sg = 0.0f;
    if (r < P.B) { n_dm = __ldg(P.d_mu + r); n_ds = __ldg(P.d_sigma + r); n_sg = __ldg(P.sigma + r); }
  };
  if ((int64_t)blockIdx.x < n_tiles) {
    tc_load_codes<4>(wd, P, (int64_t)blockIdx.x * kTB, 16, aligned);
    load_heads((int64_t)blockIdx.x * kTB);
  }
  uint32_t phase = 0, acc = 0;
  bool pending = false;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t tile0 = tile * kTB;
    const int64_t r = tile0 + p;
    const bool ok = r < P.B;
    // this person's dpre row from `pre` (global loads first; they fly while the previous tile's MMAs finish)
    const float a = n_dm * P.gscale, b = n_ds * n_sg * P.gscale;       // exp chain rule; 0 beyond B
    float pv[2][8];
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int h0 = 8 * (hq + 4 * q);
#pragma unroll
      for (int i = 0; i < 8; ++i) pv[q][i] = 0.0f;
      if (ok) {
        const float* src = P.pre + (size_t)r * P.H + h0;
        if ((P.H & 3) == 0) {
          if (h0 < P.H) {
            const float4 u = __ldg(reinterpret_cast<const float4*>(src));
            pv[q][0] = u.x; pv[q][1] = u.y; pv[q][2] = u.z; pv[q][3] = u.w;
          }
          if (h0 + 4 < P.H) {
            const float4 u = __ldg(reinterpret_cast<const float4*>(src + 4));
            pv[q][4] = u.x; pv[q][5] = u.y; pv[q][6] = u.z; pv[q][7] = u.w;
          }
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i)
            if (h0 + i < P.H) pv[q][i] = __ldg(src + i);
        }
      }
    }
    if (pending) {                                  // previous tile's MMAs have read the operand tiles
      umma::mbar_wait(&bar, phase);
      phase ^= 1u;
      pending = false;
    }
    tc_store_x<4>(Xt, RSx, wd, 16, P.J, ok);
    if (hq == 0) { s_gm += a; s_gl += b; }
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      uint32_t hi[8], mid[8], lo[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float sp, sg;
        softplus_parts(pv[q][i], sp, sg);
        const int h = 8 * (hq + 4 * q) + i;
        const float dp = fmaf(a, wh[0][h], b * wh[1][h]) * sg;          // 0 beyond H (weights) and beyond B (a, b)
        s21[q][i] = fmaf(a, sp, s21[q][i]);
        s22[q][i] = fmaf(b, sp, s22[q][i]);
        split3(dp, hi[i], mid[i], lo[i]);
      }
      const uint32_t o = (uint32_t)(p >> 3) * RSd + (uint32_t)(hq + 4 * q) * kCS + (uint32_t)(p & 7) * 16u;
      *reinterpret_cast<uint4*>(Dt[0] + o) =
          make_uint4(hi[0] | hi[1] << 16, hi[2] | hi[3] << 16, hi[4] | hi[5] << 16, hi[6] | hi[7] << 16);
      *reinterpret_cast<uint4*>(Dt[1] + o) =
          make_uint4(mid[0] | mid[1] << 16, mid[2] | mid[3] << 16, mid[4] | mid[5] << 16, mid[6] | mid[7] << 16);
      *reinterpret_cast<uint4*>(Dt[2] + o) =
          make_uint4(lo[0] | lo[1] << 16, lo[2] | lo[3] << 16, lo[4] | lo[5] << 16, lo[6] | lo[7] << 16);
    }
    umma::fence_proxy_async();
    __syncthreads();
    if (t == 0) {
      umma::fence_after_sync();
      for (int ks = 0; ks < kTB / 16; ++ks) {
        const uint64_t da = umma::make_desc(sX + ks * 2 * RSx, RSx, kCS);
#pragma unroll
        for (int term = 0; term < 3; ++term) {
          umma::mma_f16_ss(tmem, da, umma::make_desc(sD[term] + ks * 2 * RSd, RSd, kCS), idesc, acc);
          acc = 1;
        }
      }
      umma::commit(&bar);
    }
    pending = true;
    if (tile + gridDim.x < n_tiles) {
      tc_load_codes<4>(wd, P, (tile + gridDim.x) * kTB, 16, aligned);
      load_heads((tile + gridDim.x) * kTB);
    }
  }
  if (pending) umma::mbar_wait(&bar, phase);
  umma::fence_after_sync();
  // D[j][h] (TMEM lane = item slot j, column = h) -> this CTA's partial row
  // [d_w1 (H x J) | d_b1 (H) | d_w21 (H) | d_b21 | d_w22 (H) | d_b22]; slot J is the ones column = d_b1
  float* row = P.partial + (size_t)blockIdx.x * P.cols;
  float* tail = row + (size_t)P.H * P.J;
  if (warp < 8) {
    const int quad = warp & 3, half = warp >> 2;
    float v[32];
    umma::tmem_ld32(tmem + ((uint32_t)(quad * 32) << 16) + (uint32_t)(half * 32), v);
    const int j = quad * 32 + lane;
#pragma unroll
    for (int c = 0; c < 32; ++c) {
      const int h = half * 32 + c;
      if (h < P.H) {
        if (j < P.J) row[(size_t)h * P.J + j] = v[c];
        else if (j == P.J) tail[h] = v[c];
      }
    }
  }
  // d fc21.weight / d fc22.weight: sum over the persons (the four warps of one quarter) of the register partials
#pragma unroll
  for (int q = 0; q < 2; ++q)
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float x = warp_sum(s21[q][i]), y = warp_sum(s22[q][i]);
      if (lane == 0) {
        red[0][warp & 3][8 * (hq + 4 * q) + i] = x;
        red[1][warp & 3][8 * (hq + 4 * q) + i] = y;
      }
    }
  s_gm = warp_sum(s_gm); s_gl = warp_sum(s_gl);
  if (lane == 0 && hq == 0) { gm[warp] = s_gm; gl[warp] = s_gl; }
  umma::fence_before_sync();
  __syncthreads();
  if (t < P.H) {
    tail[P.H + t] = (red[0][0][t] + red[0][1][t]) + (red[0][2][t] + red[0][3][t]);
    tail[2 * P.H + 1 + t] = (red[1][0][t] + red[1][1][t]) + (red[1][2][t] + red[1][3][t]);
  }
  if (t == 0) {
    tail[2 * P.H] = (gm[0] + gm[1]) + (gm[2] + gm[3]);
    tail[3 * P.H + 1] = (gl[0] + gl[1]) + (gl[2] + gl[3]);
  }
  if (warp == 0) umma::tmem_dealloc(tmem, kTmemCols);
}

static size_t tc_fwd_smem(int JP) { return (size_t)40 * (JP / 8) * kCS + 1024; }
constexpr size_t kTcBwdSmem = 16 * (kTB / 8) * kCS + 3 * 16 * (kH / 8) * kCS + 1024;

int enc_tc_fwd_launch(const EncParams& P, cudaStream_t st) {
  const int JP = (P.J + 15) / 16 * 16;
  static size_t attr = 0;
  const size_t smem = tc_fwd_smem(JP);
  if (smem > attr) {
    VIPSY_CUDA_CHECK(cudaFuncSetAttribute(enc_tc_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = smem;
  }
  const int64_t tiles = (P.B + kTB - 1) / kTB;
  const int64_t cap = 2 * (int64_t)num_sms();
  enc_tc_fwd_kernel<<<(int)(tiles < cap ? tiles : cap), kThreads, smem, st>>>(P, JP);
  VIPSY_CUDA_CHECK(cudaGetLastError());
  return VIPSY_OK;
}

int enc_tc_bwd_launch(const EncParams& P, int grid_x, cudaStream_t st) {
  static bool attr = false;
  if (!attr) {
    VIPSY_CUDA_CHECK(cudaFuncSetAttribute(enc_tc_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)kTcBwdSmem));
    attr = true;
  }
  enc_tc_bwd_kernel<<<grid_x, kBwdThreads, kTcBwdSmem, st>>>(P);
  VIPSY_CUDA_CHECK(cudaGetLastError());
  return VIPSY_OK;
}

}  // namespace vipsy
