// NormEncoder products on the 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM), J <= 127, H <= 64.
//
// The response codes are exactly representable in bf16 (0, 1, -1); the fp32 factor of each product (fc1.weight
// forward, d loss / d pre backward) is split into three bf16 terms v = hi + mid + lo (24 mantissa bits, fp32
// exponent range), so three MMAs per 16-deep k-step reproduce the fp32 product and the fp32 accumulator in TMEM
// adds them -- same numerics class as the SIMT kernels (enc_norm.cu), a tenth of the issue slots.
//
//   forward : pre[128 persons x 64] = [X | 1][128 x JP] . [W1 | b1]^T   A = code tile (K-major), B = W1 tiles (K-major),
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
constexpr int kPitch = kH + 4;               // fp32 [person][hidden] staging tiles: conflict-free rows
constexpr int kBwdThreads = 512;          // backward: 4 threads per person (16 hidden units each)

bool enc_tc_supported(int H, int J) { return H >= 1 && H <= kH && J >= 1 && J <= 127; }
// with resident operand tiles the item axis may span two 128-slot blocks (slot J is the ones column)
bool enc_tc_resident_supported(int H, int J) { return H >= 1 && H <= kH && J >= 1 && J <= 255; }
static int enc_tc_slots(int J) { return J <= 127 ? 128 : 256; }
size_t enc_tc_tile_bytes(int J) { return (size_t)kTB * enc_tc_slots(J) * 2; }

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
// Same split by truncation, two values at a time, integer pipe only: the upper halves of v, v - hi, v - hi - mid
// are the three bf16 terms (8 + 8 + 8 mantissa bits, every residual exact); packed as (v0 | v1 << 16) per term.
__device__ __forceinline__ void split3_pair(float v0, float v1, uint32_t& hi, uint32_t& mid, uint32_t& lo) {
  const uint32_t h0 = __float_as_uint(v0) & 0xffff0000u, h1 = __float_as_uint(v1) & 0xffff0000u;
  const float r0 = v0 - __uint_as_float(h0), r1 = v1 - __uint_as_float(h1);
  const uint32_t m0 = __float_as_uint(r0) & 0xffff0000u, m1 = __float_as_uint(r1) & 0xffff0000u;
  const float q0 = r0 - __uint_as_float(m0), q1 = r1 - __uint_as_float(m1);
  hi = __byte_perm(h0, h1, 0x7632);
  mid = __byte_perm(m0, m1, 0x7632);
  lo = __byte_perm(__float_as_uint(q0), __float_as_uint(q1), 0x7632);
}
// four response codes (0 / 1 / 255) -> four bf16 (0, 1, -1) by byte lookup: code & 3 in {0, 1, 3} selects the low
// byte {00, 80, 80} and the high byte {00, 3f, bf} of the bf16 through one PRMT per pair of codes
__device__ __forceinline__ void codes_bf16x4(uint32_t word, uint32_t& lo_pair, uint32_t& hi_pair) {
  const uint32_t m = word & 0x03030303u;
  const uint32_t s0 = (m & 0xffffu) * 0x11u + 0x4040u, s1 = (m >> 16) * 0x11u + 0x4040u;
  lo_pair = __byte_perm(0x80008000u, 0xbf003f00u, s0);
  hi_pair = __byte_perm(0x80008000u, 0xbf003f00u, s1);
}
// Response codes of a 128-person tile -> registers -> bf16 operand tile.  A warp instruction covers 8 persons (one
// row block of core matrices) x 4 consecutive 4-byte words, so a request touches 8 short segments instead of 32
// rows; warp w owns row blocks w, w + NW, ...; per row block 8 iterations cover 32 words (128 items).
// `ones_col` >= 0: that item slot reads as 1 for persons below b_ones (bias column of either product).
template <int NW>
__device__ __forceinline__ void tc_load_codes(uint32_t (&wd)[16 / NW][8], const EncParams& P, int64_t tile0, int n_words,
                                              bool aligned) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, pp = lane & 7, w0 = lane >> 3;
  const int jt = P.J & ~3;                          // start of the one partial word of a row (J % 4 != 0)
#pragma unroll
  for (int rbi = 0; rbi < 16 / NW; ++rbi) {
    const int64_t r = tile0 + (warp + NW * rbi) * 8 + pp;
    const bool ok = r < P.B;
    const uint8_t* g = ok ? P.y + (P.idx ? __ldg(P.idx + r) : r) * P.ld : P.y;
    if (aligned) {
      uint32_t tail = 0;
      if ((P.J & 3) && ok && ((jt >> 2) & 3) == w0 && (jt >> 2) < n_words)
        for (int b = 0; b < (P.J & 3); ++b) tail |= (uint32_t)__ldg(g + jt + b) << (8 * b);
#pragma unroll
      for (int it = 0; it < 8; ++it) {              // n_words is a multiple of 4: `it` alone bounds the word
        const int j = 4 * (w0 + 4 * it);
        uint32_t v = j == jt ? tail : 0u;
        if (4 * it < n_words && ok && j + 3 < P.J) v = __ldg(reinterpret_cast<const uint32_t*>(g + j));
        wd[rbi][it] = v;
      }
    } else {
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int word = w0 + 4 * it, j = 4 * word;
        uint32_t v = 0;
        if (ok && word < n_words && j < P.J) {
#pragma unroll
          for (int b = 0; b < 4; ++b)
            if (j + b < P.J) v |= (uint32_t)__ldg(g + j + b) << (8 * b);
        }
        wd[rbi][it] = v;
      }
    }
  }
}
template <int NW>
__device__ __forceinline__ void tc_store_x(unsigned char* Xt, uint32_t RS, const uint32_t (&wd)[16 / NW][8], int n_words,
                                           int ones_col, int64_t tile0, int64_t b_ones) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, pp = lane & 7, w0 = lane >> 3;
#pragma unroll
  for (int rbi = 0; rbi < 16 / NW; ++rbi) {
    const int rb = warp + NW * rbi;
    const bool one_ok = tile0 + rb * 8 + pp < b_ones;
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const int word = w0 + 4 * it;
      if (4 * it < n_words) {                       // uniform: n_words is a multiple of 4
        uint2 v;
        codes_bf16x4(wd[rbi][it], v.x, v.y);
        if (ones_col >= 0 && (ones_col >> 2) == word && one_ok) {
          const uint32_t one = 0x3f80u << (16 * (ones_col & 1));
          if (ones_col & 2) v.y |= one; else v.x |= one;
        }
        *reinterpret_cast<uint2*>(Xt + (uint32_t)rb * RS + (uint32_t)(word >> 1) * kCS + (uint32_t)pp * 16u +
                                  (uint32_t)(word & 1) * 8u) = v;
      }
    }
  }
}

// ---- resident operand tiles ---------------------------------------------------------------------------------
// The response matrix never changes between SVI steps, so its bf16 operand tiles can live in HBM in exactly the
// shared-memory layout (per 128 persons: 16 row blocks x NS/8 chunks x 8 rows x 8 bf16, NS = 128 or 256 item slots = 32 / 64 KB, item
// slot J = 1 for existing persons).  A tile then reaches shared memory by ONE bulk asynchronous copy that completes on an mbarrier.
__global__ void __launch_bounds__(256) enc_tc_pack_kernel(const uint8_t* __restrict__ y, int64_t n_rows, int J,
                                                          int64_t ld, unsigned char* __restrict__ tiles, int NS) {
  const int64_t tile0 = (int64_t)blockIdx.x * kTB;
  unsigned char* out = tiles + (size_t)blockIdx.x * ((size_t)kTB * NS * 2);
  for (int id = threadIdx.x; id < 16 * NS; id += 256) {        // 16-byte chunk id = rb * NS + ch * 8 + pp
    const int rb = id / NS, rem = id - rb * NS, ch = rem >> 3, pp = rem & 7;
    const int64_t r = tile0 + rb * 8 + pp;
    uint32_t w[2] = {0u, 0u};
    if (r < n_rows) {
      const uint8_t* g = y + r * ld;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int j = 8 * ch + k;
        const uint32_t c = j < J ? (uint32_t)g[j] : (j == J ? 1u : 0u);
        w[k >> 2] |= c << (8 * (k & 3));
      }
    }
    uint4 v;
    codes_bf16x4(w[0], v.x, v.y);
    codes_bf16x4(w[1], v.z, v.w);
    *reinterpret_cast<uint4*>(out + (size_t)id * 16) = v;
  }
}

int enc_tc_pack_launch(const uint8_t* y, int64_t n_rows, int J, int64_t ld, void* tiles, cudaStream_t st) {
  const int64_t n_tiles = (n_rows + kTB - 1) / kTB;
  enc_tc_pack_kernel<<<(unsigned)n_tiles, 256, 0, st>>>(y, n_rows, J, ld, reinterpret_cast<unsigned char*>(tiles),
                                                        enc_tc_slots(J));
  VIPSY_CUDA_CHECK(cudaGetLastError());
  return VIPSY_OK;
}

// one thread: expect `bytes` on the mbarrier and start the bulk copy global -> shared that will deliver them
__device__ __forceinline__ void bulk_load(unsigned char* dst_smem, const unsigned char* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(umma::smem_u32(bar)), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   umma::smem_u32(dst_smem)),
               "l"(src), "r"(bytes), "r"(umma::smem_u32(bar))
               : "memory");
}

// ------------------------------------------------------------------------------------------------ forward
template <int NS>
__global__ void __launch_bounds__(kThreads, NS == 128 ? 2 : 1) enc_tc_fwd_kernel(const EncParams P, const int JP) {
  constexpr uint32_t kRSTile = (NS / 8) * kCS;    // row-block pitch of a resident tile
  constexpr uint32_t kTileBytes = kTB * NS * 2;
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar, barx;
  __shared__ uint32_t tmem_base_sm;
  __shared__ float hv[3 * kH];                    // b1 | w21 | w22, zero beyond H
  __shared__ float2 part[2][kTB];                 // per column half: partial (mu, log sigma) of each person
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int n_chunks = JP >> 3;
  const bool resident = NS == 256 || P.xtiles != nullptr;          // two slot blocks exist only as resident tiles
  const uint32_t RS = (uint32_t)n_chunks * kCS;   // row-block pitch of the W tiles (8 row blocks each)
  const uint32_t RSx = resident ? kRSTile : RS;   // ... and of the code tile (16 row blocks)
  unsigned char* Xt = smem;
  unsigned char* Wt[3] = {Xt + 16 * RSx, Xt + 16 * RSx + 8 * RS, Xt + 16 * RSx + 16 * RS};
  float* ps = reinterpret_cast<float*>(Xt + 16 * RSx + 24 * RS);     // [128][kPitch] pre tile on its way to global memory
  const bool aligned = ((reinterpret_cast<uintptr_t>(P.y) | (uintptr_t)P.ld) & 3) == 0;

  // fc1.weight -> three bf16 operand tiles, once per CTA; eight loads in flight per thread (the CTA's first
  // MMA waits for this, and a one-tile minibatch launch is nothing but this prologue)
  for (int e0 = t; e0 < kH * JP; e0 += 8 * kThreads) {
    float v[8];
    uint32_t o[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int e = e0 + u * kThreads;
      const int h = e / JP, j = e - h * JP;
      v[u] = 0.0f;                                  // item slot J multiplies the ones column: the bias row
      if (e < kH * JP && h < P.H) v[u] = j < P.J ? __ldg(P.w1 + (size_t)h * P.J + j) : (j == P.J ? __ldg(P.b1 + h) : 0.0f);
      o[u] = umma::tile_off(h, j, RS, kCS);
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      if (e0 + u * kThreads < kH * JP) {
        uint32_t a, b, c;
        split3(v[u], a, b, c);
        *reinterpret_cast<uint16_t*>(Wt[0] + o[u]) = (uint16_t)a;
        *reinterpret_cast<uint16_t*>(Wt[1] + o[u]) = (uint16_t)b;
        *reinterpret_cast<uint16_t*>(Wt[2] + o[u]) = (uint16_t)c;
      }
    }
  }
  if (t < 3 * kH) {
    const int h = t & 63, which = t >> 6;
    const float* src = which == 0 ? P.b1 : which == 1 ? P.w21 : P.w22;
    hv[t] = h < P.H ? __ldg(src + h) : 0.0f;
  }
  if (warp == 0) umma::tmem_alloc(&tmem_base_sm, kTmemCols);
  if (t == 0) {
    umma::mbar_init(&bar, 1);
    umma::mbar_init(&barx, 1);
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
  uint32_t wd[2][8];
  if (resident) {
    if (t == 0) bulk_load(Xt, P.xtiles + (size_t)blockIdx.x * kTileBytes, kTileBytes, &barx);
  } else {
    if constexpr (NS == 128) tc_load_codes<8>(wd, P, (int64_t)blockIdx.x * kTB, 2 * n_chunks, aligned);
  }
  uint32_t phase = 0;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t tile0 = tile * kTB;
    if (!resident) {
      if constexpr (NS == 128) tc_store_x<8>(Xt, RS, wd, 2 * n_chunks, P.J, tile0, INT64_MAX);
      umma::fence_proxy_async();
    }
    __syncthreads();
    if (warp == 0) {                                // the warp waits, one elected lane issues (see umma::elect_one)
      if (resident) umma::mbar_wait(&barx, phase);  // this tile's bulk copy has landed
      umma::fence_after_sync();
      if (umma::elect_one()) {
        uint32_t acc = 0;
        for (int ks = 0; ks < JP / 16; ++ks) {
          const uint64_t da = umma::make_desc(sX + ks * 2 * kCS, kCS, RSx);
#pragma unroll
          for (int term = 0; term < 3; ++term) {
            umma::mma_f16_ss(tmem, da, umma::make_desc(sW[term] + ks * 2 * kCS, kCS, RS), idesc, acc);
            acc = 1;
          }
        }
        umma::commit(&bar);
      }
      __syncwarp();
    }
    if constexpr (NS == 128) {
      if (!resident && tile + gridDim.x < n_tiles)
        tc_load_codes<8>(wd, P, (tile + gridDim.x) * kTB, 2 * n_chunks, aligned);
    }
    umma::mbar_wait(&bar, phase);
    phase ^= 1u;
    umma::fence_after_sync();
    if (resident && t == 0 && tile + gridDim.x < n_tiles)         // the MMAs are done with the code tile: refill it
      bulk_load(Xt, P.xtiles + (size_t)(tile + gridDim.x) * kTileBytes, kTileBytes, &barx);
    {
      float v[32];
      umma::tmem_ld32(tmem + ((uint32_t)(quad * 32) << 16) + (uint32_t)(half * 32), v);
      const int p = quad * 32 + lane;
      const int h0 = half * 32;
      float m = 0.0f, l = 0.0f;
#pragma unroll
      for (int c = 0; c < 32; ++c) {
        float sp, sg;
        softplus_parts(v[c], sp, sg);
        m = fmaf(sp, hv[kH + h0 + c], m);           // weights are 0 beyond H
        l = fmaf(sp, hv[2 * kH + h0 + c], l);
      }
#pragma unroll
      for (int c = 0; c < 32; c += 4)
        *reinterpret_cast<float4*>(ps + p * kPitch + h0 + c) = make_float4(v[c], v[c + 1], v[c + 2], v[c + 3]);
      part[half][p] = make_float2(m, l);
    }
    umma::fence_before_sync();
    __syncthreads();                                // TMEM and the code tile are free; both halves are in `part`
    if (t < kTB && tile0 + t < P.B) {
      const float2 a = part[0][t], b = part[1][t];
      P.mu[tile0 + t] = (a.x + b.x) + b21;
      P.sigma[tile0 + t] = expf((a.y + b.y) + b22);
    }
    // pre tile -> global memory, whole rows per warp request (the tile is one contiguous block of `pre`)
    if (P.H == kH) {                                // the tile is 32 KB contiguous: 8 float4 per thread
      float* dst = P.pre + (size_t)tile0 * kH;
      const int64_t lim = (P.B - tile0) * (kH / 4);
#pragma unroll
      for (int q = 0; q < kTB * (kH / 4) / kThreads; ++q) {
        const int e = t + q * kThreads;
        if (e < lim)
          *reinterpret_cast<float4*>(dst + 4 * e) = *reinterpret_cast<const float4*>(ps + (e >> 4) * kPitch + 4 * (e & 15));
      }
    } else if ((P.H & 3) == 0) {
      const int h4 = P.H >> 2;
      for (int e = t; e < kTB * h4; e += kThreads) {
        const int row = e / h4, c4 = e - row * h4;
        if (tile0 + row < P.B)
          *reinterpret_cast<float4*>(P.pre + (size_t)(tile0 + row) * P.H + 4 * c4) =
              *reinterpret_cast<const float4*>(ps + row * kPitch + 4 * c4);
      }
    } else {
      for (int e = t; e < kTB * P.H; e += kThreads) {
        const int row = e / P.H, c = e - row * P.H;
        if (tile0 + row < P.B) P.pre[(size_t)(tile0 + row) * P.H + c] = ps[row * kPitch + c];
      }
    }
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem, kTmemCols);
}

// ------------------------------------------------------------------------------------------------ backward
template <int NS>
__global__ void __launch_bounds__(kBwdThreads) enc_tc_bwd_kernel(const EncParams P) {
  constexpr uint32_t kTileBytes = kTB * NS * 2;
  constexpr int MB = NS / 128;                    // 128-slot blocks of the item axis = accumulators in TMEM
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar, barx;
  __shared__ uint32_t tmem_base_sm;
  __shared__ float gm[kTB], gl[kTB];              // d loss / d mu, d loss / d log sigma of the tile's persons
  __shared__ float red[2][4][kH];                 // end of kernel: [w21 | w22][warp of the half][h]
  __shared__ float wh[2][kH];                     // fc21.weight | fc22.weight, zero beyond H
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  constexpr uint32_t RSx = (NS / 8) * kCS;        // code tile: 128 persons x NS item slots (item J = ones column)
  constexpr uint32_t RSd = (kH / 8) * kCS;        // dpre tiles: 128 persons x 64 hidden
  unsigned char* Xt = smem;
  unsigned char* Dt[3] = {Xt + 16 * RSx, Xt + 16 * RSx + 16 * RSd, Xt + 16 * RSx + 32 * RSd};
  float* pb0 = reinterpret_cast<float*>(Xt + 16 * RSx + 48 * RSd);     // 2 x [128][kPitch]: `pre` rows, this tile / next
  const bool aligned = ((reinterpret_cast<uintptr_t>(P.y) | (uintptr_t)P.ld) & 3) == 0;
  const bool resident = NS == 256 || P.xtiles != nullptr;
  const int p = t & (kTB - 1), hq = t >> 7;       // this thread: person p, hidden chunks hc = hq + 4 q (8 units each)

  if (t < 2 * kH) {
    const int h = t & 63;
    wh[t >> 6][h] = h < P.H ? __ldg((t >> 6 ? P.w22 : P.w21) + h) : 0.0f;
  }
  for (int e = t; e < 2 * kTB * kPitch; e += kBwdThreads) pb0[e] = 0.0f;        // columns beyond H stay 0
  // The tensor core adds into its fp32 accumulator with truncation: a chain of accumulating MMAs drifts by about
  // 2^-24 of the running sum per instruction (measured at 1M persons: d fc1.weight off by 1.6e-5 after 1272 MMAs).
  // The CTA's tiles therefore go round-robin into NACC independent accumulator sets, summed at the drain.
  constexpr uint32_t NACC = 512 / (MB * kTmemCols);
  if (warp == 0) umma::tmem_alloc(&tmem_base_sm, 512);
  if (t == 0) {
    umma::mbar_init(&bar, 1);
    umma::mbar_init(&barx, 1);
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
  uint32_t wd[1][8];
  float n_dm = 0.0f, n_ds = 0.0f, n_sg = 0.0f;
  auto load_heads = [&](int64_t tile0) {
    const int64_t r = tile0 + p;
    n_dm = n_ds = n_sg = 0.0f;
    if (r < P.B) { n_dm = __ldg(P.d_mu + r); n_ds = __ldg(P.d_sigma + r); n_sg = __ldg(P.sigma + r); }
  };
  // `pre` rows of a tile -> pb: the tile is one contiguous block of the [B x H] matrix, copied by 16-byte
  // asynchronous copies (whole rows per warp request); rows beyond B read as 0
  auto fetch_pre = [&](int64_t tile0, int buf) {
    float* pb = pb0 + buf * (kTB * kPitch);
    if (P.H == kH) {                                // 32 KB contiguous: 4 asynchronous 16-byte copies per thread
      const float* src = P.pre + (size_t)tile0 * kH;
      const int64_t lim = (P.B - tile0) * (kH / 4);
#pragma unroll
      for (int q = 0; q < kTB * (kH / 4) / kBwdThreads; ++q) {
        const int e = t + q * kBwdThreads;
        float* dst = pb + (e >> 4) * kPitch + 4 * (e & 15);
        if (e < lim) {
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(umma::smem_u32(dst)), "l"(src + 4 * e)
                       : "memory");
        } else {
          *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
    } else if ((P.H & 3) == 0) {
      const int h4 = P.H >> 2;
      for (int e = t; e < kTB * h4; e += kBwdThreads) {
        const int row = e / h4, c4 = e - row * h4;
        float* dst = pb + row * kPitch + 4 * c4;
        if (tile0 + row < P.B) {
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(umma::smem_u32(dst)),
                       "l"(P.pre + (size_t)(tile0 + row) * P.H + 4 * c4)
                       : "memory");
        } else {
          *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
    } else {
      for (int e = t; e < kTB * P.H; e += kBwdThreads) {
        const int row = e / P.H, c = e - row * P.H;
        pb[row * kPitch + c] = tile0 + row < P.B ? __ldg(P.pre + (size_t)(tile0 + row) * P.H + c) : 0.0f;
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  if ((int64_t)blockIdx.x < n_tiles) {
    if (resident) {
      if (t == 0) bulk_load(Xt, P.xtiles + (size_t)blockIdx.x * kTileBytes, kTileBytes, &barx);
    } else {
      if constexpr (NS == 128) tc_load_codes<16>(wd, P, (int64_t)blockIdx.x * kTB, 32, aligned);
    }
    fetch_pre((int64_t)blockIdx.x * kTB, 0);
    load_heads((int64_t)blockIdx.x * kTB);
  }
  uint32_t phase = 0, phx = 0, nt = 0;           // nt: tiles this CTA has issued
  int nb = 0;                                     // `pre` buffer of the current tile
  bool pending = false;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t tile0 = tile * kTB;
    const float a = n_dm * P.gscale, b = n_ds * n_sg * P.gscale;       // exp chain rule; 0 beyond B
    const bool has_next = tile + gridDim.x < n_tiles;
    if (has_next) {                                 // next tile's `pre` rows: a whole iteration ahead
      fetch_pre((tile + gridDim.x) * kTB, (nb ^ 1));
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    const float* pb = pb0 + nb * (kTB * kPitch);
    if (pending) {                                  // previous tile's MMAs have read the operand tiles
      umma::mbar_wait(&bar, phase);
      phase ^= 1u;
      pending = false;
      if (resident && t == 0) bulk_load(Xt, P.xtiles + (size_t)tile * kTileBytes, kTileBytes, &barx);
    }
    __syncthreads();                                // every thread's share of `pre` has landed in pb
    float pv[2][8];
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const float* src = pb + p * kPitch + 8 * (hq + 4 * q);
      const float4 u0 = *reinterpret_cast<const float4*>(src), u1 = *reinterpret_cast<const float4*>(src + 4);
      pv[q][0] = u0.x; pv[q][1] = u0.y; pv[q][2] = u0.z; pv[q][3] = u0.w;
      pv[q][4] = u1.x; pv[q][5] = u1.y; pv[q][6] = u1.z; pv[q][7] = u1.w;
    }
    if constexpr (NS == 128) {
      if (!resident) tc_store_x<16>(Xt, RSx, wd, 32, P.J, tile0, P.B);
    }
    if (hq == 0) { s_gm += a; s_gl += b; }
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      uint32_t hi[4], mid[4], lo[4];
      float dp[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float sp, sg;
        softplus_parts(pv[q][i], sp, sg);
        const int h = 8 * (hq + 4 * q) + i;
        dp[i] = fmaf(a, wh[0][h], b * wh[1][h]) * sg;                   // 0 beyond H (weights) and beyond B (a, b)
        s21[q][i] = fmaf(a, sp, s21[q][i]);
        s22[q][i] = fmaf(b, sp, s22[q][i]);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) split3_pair(dp[2 * i], dp[2 * i + 1], hi[i], mid[i], lo[i]);
      const uint32_t o = (uint32_t)(p >> 3) * RSd + (uint32_t)(hq + 4 * q) * kCS + (uint32_t)(p & 7) * 16u;
      *reinterpret_cast<uint4*>(Dt[0] + o) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
      *reinterpret_cast<uint4*>(Dt[1] + o) = make_uint4(mid[0], mid[1], mid[2], mid[3]);
      *reinterpret_cast<uint4*>(Dt[2] + o) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    }
    umma::fence_proxy_async();
    __syncthreads();
    if (warp == 0) {                                // the warp waits, one elected lane issues (see umma::elect_one)
      if (resident) umma::mbar_wait(&barx, phx);    // this tile's bulk copy has landed
      umma::fence_after_sync();
      if (umma::elect_one()) {
        for (int ks = 0; ks < kTB / 16; ++ks) {
#pragma unroll
          for (int mb = 0; mb < MB; ++mb) {           // item-slot block mb: 16 core-matrix columns further along
            const uint64_t da = umma::make_desc(sX + ks * 2 * RSx + mb * 16 * kCS, RSx, kCS);
#pragma unroll
            for (int term = 0; term < 3; ++term)        // only the very first MMA of an accumulator overwrites it
              umma::mma_f16_ss(tmem + ((nt % NACC) * MB + mb) * kTmemCols, da,
                               umma::make_desc(sD[term] + ks * 2 * RSd, RSd, kCS), idesc,
                               ((nt >= NACC ? 1u : 0u) | (uint32_t)ks | (uint32_t)term) != 0u ? 1u : 0u);
          }
        }
        umma::commit(&bar);
      }
      __syncwarp();
    }
    if (resident) phx ^= 1u;
    ++nt;
    pending = true;
    nb ^= 1;
    if (has_next) {
      if constexpr (NS == 128) {
        if (!resident) tc_load_codes<16>(wd, P, (tile + gridDim.x) * kTB, 32, aligned);
      }
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
#pragma unroll
    for (int mb = 0; mb < MB; ++mb) {
      float v[32];
      umma::tmem_ld32(tmem + ((uint32_t)(quad * 32) << 16) + (uint32_t)(mb * kTmemCols + half * 32), v);
      for (uint32_t s = 1; s < (nt < NACC ? nt : NACC); ++s) {       // the other accumulator sets, in order
        float w[32];
        umma::tmem_ld32(tmem + ((uint32_t)(quad * 32) << 16) + (uint32_t)((s * MB + mb) * kTmemCols + half * 32), w);
#pragma unroll
        for (int c = 0; c < 32; ++c) v[c] += w[c];
      }
      const int j = mb * 128 + quad * 32 + lane;
#pragma unroll
      for (int c = 0; c < 32; ++c) {
        const int h = half * 32 + c;
        if (h < P.H) {
          if (j < P.J) row[(size_t)h * P.J + j] = v[c];
          else if (j == P.J) tail[h] = v[c];
        }
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
  if (warp == 0) umma::tmem_dealloc(tmem, 512);
}

static size_t tc_fwd_smem(int JP, bool resident, int NS) {
  const size_t RS = (size_t)(JP / 8) * kCS;
  return 16 * (resident ? (size_t)(NS / 8) * kCS : RS) + 24 * RS + (size_t)kTB * kPitch * 4 + 1024;
}
static size_t tc_bwd_smem(int NS) {
  return 16 * (size_t)(NS / 8) * kCS + 3 * 16 * (kH / 8) * kCS + 2 * (size_t)kTB * kPitch * 4 + 1024;
}

int enc_tc_fwd_launch(const EncParams& P, cudaStream_t st) {
  const int JP = (P.J + 1 + 15) / 16 * 16;        // + the ones column that carries fc1.bias
  const int NS = enc_tc_slots(P.J);
  const bool resident = P.xtiles != nullptr;
  if (NS == 256 && !resident) {
    set_error("enc_tc_fwd_launch: J > 127 needs resident operand tiles");
    return VIPSY_ERR_UNSUPPORTED;
  }
  const size_t smem = tc_fwd_smem(JP, resident, NS);
  const int64_t tiles = (P.B + kTB - 1) / kTB;
  if (NS == 128) {
    // per-device attribute: set on every launch, no process-wide cache
    VIPSY_CUDA_CHECK(cudaFuncSetAttribute(enc_tc_fwd_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t cap = 2 * (int64_t)num_sms();
    enc_tc_fwd_kernel<128><<<(int)(tiles < cap ? tiles : cap), kThreads, smem, st>>>(P, JP);
  } else {
    VIPSY_CUDA_CHECK(cudaFuncSetAttribute(enc_tc_fwd_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t cap = (int64_t)num_sms();        // one CTA per SM: 64 KB code tile + up to 96 KB of W tiles
    enc_tc_fwd_kernel<256><<<(int)(tiles < cap ? tiles : cap), kThreads, smem, st>>>(P, JP);
  }
  VIPSY_CUDA_CHECK(cudaGetLastError());
  return VIPSY_OK;
}

int enc_tc_bwd_launch(const EncParams& P, int grid_x, cudaStream_t st) {
  const int NS = enc_tc_slots(P.J);
  if (NS == 256 && P.xtiles == nullptr) {
    set_error("enc_tc_bwd_launch: J > 127 needs resident operand tiles");
    return VIPSY_ERR_UNSUPPORTED;
  }
  if (NS == 128) {
    VIPSY_CUDA_CHECK(cudaFuncSetAttribute(enc_tc_bwd_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)tc_bwd_smem(128)));
    enc_tc_bwd_kernel<128><<<grid_x, kBwdThreads, tc_bwd_smem(128), st>>>(P);
  } else {
    VIPSY_CUDA_CHECK(cudaFuncSetAttribute(enc_tc_bwd_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)tc_bwd_smem(256)));
    enc_tc_bwd_kernel<256><<<grid_x, kBwdThreads, tc_bwd_smem(256), st>>>(P);
  }
  VIPSY_CUDA_CHECK(cudaGetLastError());
  return VIPSY_OK;
}

}  // namespace vipsy
