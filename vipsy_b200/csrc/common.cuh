// Shared device helpers for the vipsy_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <stdint.h>
#include <stdio.h>

#include "../../include/vipsy_b200.h"

namespace vipsy {

constexpr float kEps32 = 1.1920928955078125e-07f;        // torch clamp_probs eps
constexpr float kOneMinusEps32 = 1.0f - kEps32;
constexpr float kLogEps32 = -15.942385152878742f;
// A masked cell is scored as Bernoulli(probs=0).log_prob(0) (vi.py:621-624): probs clamp to eps and the
// cell contributes log(1 - eps) = -1.19e-7 (torch 1.6 fp32 and the fp64 oracle; no gradient).
constexpr float kMaskedLogProb = -1.1920929665620834e-07f;
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;
constexpr int kNumSMs = 148;

void set_error(const char* fmt, ...);
int check_irt(const VipsyIrtArgs* a);      // argument validation shared by the IRT entry points (api.cu)

#define VIPSY_CUDA_CHECK(expr)                                                       \
  do {                                                                               \
    cudaError_t _e = (expr);                                                         \
    if (_e != cudaSuccess) {                                                         \
      ::vipsy::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),     \
                         __FILE__, __LINE__);                                        \
      return VIPSY_ERR_CUDA;                                                         \
    }                                                                                \
  } while (0)

#define VIPSY_REQUIRE(cond, code, ...)   \
  do {                                   \
    if (!(cond)) {                       \
      ::vipsy::set_error(__VA_ARGS__);   \
      return code;                       \
    }                                    \
  } while (0)

// ---- SFU wrappers: one MUFU each, ~2 ulp ----------------------------------
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float lg2_approx(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// ---- warp reductions -------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Sum NQ per-lane values across the warp with a transposing butterfly: after the
// call, lane q * (32 / P) holds the warp total of v[q], P = 4 for NQ in {3,4},
// 2 for NQ == 2, 1 for NQ == 1 (every lane of that group holds it, in fact).
// 6 shuffles for NQ=4 instead of 20; fixed order, hence deterministic.
template <int NQ>
__device__ __forceinline__ float warp_sum_multi(const float (&v)[NQ], int lane) {
  static_assert(NQ >= 1 && NQ <= 4, "NQ in 1..4");
  float r;
  if constexpr (NQ >= 3) {
    const bool hi = lane & 16;
    float v3 = 0.f;
    if constexpr (NQ == 4) v3 = v[3];
    float keep0 = hi ? v[2] : v[0], send0 = hi ? v[0] : v[2];
    float keep1 = hi ? v3 : v[1], send1 = hi ? v[1] : v3;
    keep0 += __shfl_xor_sync(0xffffffffu, send0, 16);
    keep1 += __shfl_xor_sync(0xffffffffu, send1, 16);
    const bool mid = lane & 8;
    r = (mid ? keep1 : keep0) + __shfl_xor_sync(0xffffffffu, mid ? keep0 : keep1, 8);
  } else if constexpr (NQ == 2) {
    const bool hi = lane & 16;
    r = (hi ? v[1] : v[0]) + __shfl_xor_sync(0xffffffffu, hi ? v[0] : v[1], 16);
    r += __shfl_xor_sync(0xffffffffu, r, 8);
  } else {
    r = v[0] + __shfl_xor_sync(0xffffffffu, v[0], 16);
    r += __shfl_xor_sync(0xffffffffu, r, 8);
  }
  r += __shfl_xor_sync(0xffffffffu, r, 4);
  r += __shfl_xor_sync(0xffffffffu, r, 2);
  r += __shfl_xor_sync(0xffffffffu, r, 1);
  return r;
}
template <int NQ>
__device__ __forceinline__ int warp_sum_multi_owner(int lane) {   // q owned by this lane, or -1
  constexpr int P = NQ >= 3 ? 4 : NQ;
  constexpr int W = 32 / P;
  return (lane % W == 0 && lane / W < NQ) ? lane / W : -1;
}

// ---- Philox4x32-10 (Salmon et al., SC'11) + Box-Muller ---------------------
struct Philox {
  uint32_t k0, k1;
  __device__ __forceinline__ Philox(uint64_t seed) : k0(uint32_t(seed)), k1(uint32_t(seed >> 32)) {}
  __device__ __forceinline__ uint4 operator()(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3) const {
    uint32_t a = k0, b = k1;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      const uint32_t h0 = __umulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
      const uint32_t h1 = __umulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
      c0 = h1 ^ c1 ^ a; c1 = l1; c2 = h0 ^ c3 ^ b; c3 = l0;
      a += 0x9E3779B9u; b += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
  }
};

// counter layout: (person_lo, person_hi, draw_block, offset_lo) keyed by seed ^ (offset_hi << 32)
__device__ __forceinline__ uint4 philox_draw(uint64_t seed, uint64_t offset, uint64_t person, uint32_t block) {
  Philox ph(seed ^ (offset & 0xffffffff00000000ull));
  return ph(uint32_t(person), uint32_t(person >> 32), block, uint32_t(offset));
}
__device__ __forceinline__ float u01(uint32_t r) {   // (0,1), 24 significant bits
  return (float(r >> 8) + 0.5f) * (1.0f / 16777216.0f);
}
// four N(0,1) draws from one Philox block
__device__ __forceinline__ float4 box_muller4(uint4 r) {
  const float ra = sqrtf(-2.0f * __logf(u01(r.x)));
  const float rb = sqrtf(-2.0f * __logf(u01(r.z)));
  float sa, ca, sb, cb;
  __sincosf(6.283185307179586f * u01(r.y), &sa, &ca);
  __sincosf(6.283185307179586f * u01(r.w), &sb, &cb);
  return make_float4(ra * ca, ra * sa, rb * cb, rb * sb);
}
// Noise element n (flattened (particle, dim) index: n = s*K + k) of one person.
__device__ __forceinline__ float4 normal4(uint64_t seed, uint64_t offset, uint64_t person, uint32_t block) {
  return box_muller4(philox_draw(seed, offset, person, block));
}

// Where element (particle s, dimension k) of a person's N(0,1) draw lives in the Philox stream.
// K <= 32 (and K > 128): flattened n = s*K + k, four consecutive n per block (a thread that owns one
// (person, particle) draws its K values from K/4 calls).  32 < K <= 128: block s*32 + (k mod 32),
// component k / 32 -- lane l of a warp that owns dimensions l, l+32, l+64, l+96 gets all four of its
// values from one Philox call.
__device__ __forceinline__ void noise_coord(int s, int k, int K, uint32_t& block, int& comp) {
  if (K > 32 && K <= 128) {
    block = (uint32_t)s * 32u + (uint32_t)(k & 31);
    comp = k >> 5;
  } else {
    const uint32_t n = (uint32_t)s * (uint32_t)K + (uint32_t)k;
    block = n >> 2;
    comp = (int)(n & 3u);
  }
}

// ---- launch helpers ---------------------------------------------------------
inline int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = kNumSMs;
  }
  return n;
}

// Resident CTAs per SM of one kernel instantiation at one dynamic shared-memory size, cached per call site in
// one atomic word {smem + 1 : 40 bits | occupancy : 24 bits}.  Lock-free and device-agnostic: racing threads and
// other devices of the box (one GPU model) compute the same value.
struct OccCache {
  std::atomic<uint64_t> w{0};
};
template <class Kern>
inline cudaError_t cached_occupancy(OccCache& c, Kern k, int threads, size_t smem, int hi, int* out) {
  uint64_t w = c.w.load(std::memory_order_relaxed);
  if ((w >> 24) != (uint64_t)smem + 1) {
    int occ = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, threads, smem);
    if (e != cudaSuccess) return e;
    if (occ < 1) occ = 1;
    if (occ > hi) occ = hi;
    w = (((uint64_t)smem + 1) << 24) | (uint64_t)occ;
    c.w.store(w, std::memory_order_relaxed);
  }
  *out = (int)(w & 0xffffffu);
  return cudaSuccess;
}

// Column sums of a row-major [rows x cols] fp32 partial matrix in a fixed order
// (fp64 accumulation), scattered to up to 8 output segments, plus the sum of the
// per-CTA fp64 loss partials: out_seg[c] = alpha * sum_r part[r][c].
struct FinalizeSegs {
  float* out[8];           // may be null: segment skipped
  const uint8_t* mask[8];  // optional 0/1 mask multiplied into the segment (a_free, vi.py:511-512)
  int cols[8];
  int n;
};
int launch_finalize(const float* partial, int rows, int cols, float alpha, const FinalizeSegs& segs,
                    const double* loss_partial, int loss_rows, double* loss_out, cudaStream_t stream);


// -scale * log(1-eps) * (masked cells among the batch rows) as extra fp64 loss partial rows (util.cu);
// known >= 0: the caller's count is used instead of a pass over y.  Writes *n_rows <= kMaskRows rows.
constexpr int kMaskRows = 296;
int launch_masked_term(const uint8_t* y, int64_t ld, const int64_t* idx, int64_t B, int J, float scale, int64_t known,
                       double* out_rows, int* n_rows, cudaStream_t st);
// the same constant from rows in the 2-bit transport format (no padding bits: J % 4 == 0), n_bytes = B J / 4
int launch_masked_term_2bit(const uint8_t* y2, int64_t n_bytes, float scale, int64_t known, double* out_rows, int* n_rows,
                            cudaStream_t st);

}  // namespace vipsy
