// Tensor-core (tcgen05 / TMEM) likelihood contraction of the multidimensional 2PL model
// (vi.py:32-41 irt_2pl `x.mm(a)`, vi.py:596-615 model, vi.py:617-625 masked Bernoulli likelihood).
//
// One CTA owns one 128-item tile (column block of `a`) and walks 128-row tiles of person-samples:
//
//   logits[128 x 128] = [x | 1 | 0] [a ; b ; 0]          tcgen05.mma, fp16 hi/lo split (3 MMAs per
//                                                         16-deep k-step), fp32 accumulation in TMEM
//   epilogue (TMEM -> registers): 2PL link, clamp, masked Bernoulli log-lik, G = d lb / d logit
//                                 -> G written back to shared memory as fp16 hi/lo operand tiles
//   d x [128 x KP]   = G a^T                              (same `a` tile, K-major this time)
//   d a [KP  x 128] += x^T G                              (same x and G tiles, MN-major), accumulated
//                                                         in TMEM over every row tile of the CTA;
//                                                         row K of it is d b (the ones column of x)
//
// Neither logits nor G ever reach HBM.  fp16 hi/lo: v = hi + lo with |v - hi - lo| <= 2^-22 |v|
// (absolute floor 2^-25); the three products hi*hi + hi*lo + lo*hi are exact in the fp32
// accumulator, so the contraction is fp32-class accurate (measured 6e-7, tests/test_umma_selftest.py)
// at half the shared-memory footprint and twice the tensor throughput of a 3xTF32 split.
#include "mirt.cuh"
#include "umma.cuh"

namespace vipsy {

__device__ long long g_tc_dbg[3 * 8 * 4];   // development: clock64 stamps [role][tile][event] of CTA 0
#define TC_STAMP(role, ev)                                                             \
  do {                                                                                 \
    if ((p.dbg & 32) && blockIdx.x == 0 && n < 8) g_tc_dbg[((role) * 8 + n) * 4 + (ev)] = clock64(); \
  } while (0)

constexpr int kTcEpiThreads = 256;         // warps 0-7: epilogue
constexpr int kTcXThreads = 256;           // warps 8-15: x-tile producers + d x drain
constexpr int kTcYThreads = 128;           // warps 16-19: response-tile producers
constexpr int kTcMmaWarp = (kTcEpiThreads + kTcXThreads + kTcYThreads) / 32;      // warp 20
constexpr int kTcThreads = kTcEpiThreads + kTcXThreads + kTcYThreads + 32;
constexpr int kTcTM = 128;                 // person-sample rows per tile (UMMA M)
constexpr int kTcTN = 128;                 // items per CTA (UMMA N of the forward product)
constexpr int kTcYStride = kTcTN + 4;      // response tile row stride (bank-conflict-free row reads)
constexpr uint32_t kCS = 128;              // byte stride between core matrices along the contiguous dim
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kColLogit = 0, kColDx = 128, kColDa = 256;   // + a second d a accumulator at kColDa + 128
// The tensor core adds into its fp32 accumulator with truncation, so a chain of accumulating MMAs drifts by about
// 2^-24 of the running sum per instruction (measured at 200 000 persons: d a off by 1.7e-5 after 1008 MMAs into
// one accumulator).  d a therefore accumulates in TMEM for kTcFlushTiles row tiles only; the epilogue adds each
// finished group to the CTA's partial row (round-to-nearest) while the next group builds up in the other accumulator.
constexpr int kTcFlushTiles = 8;

__host__ __device__ inline int tc_kp(int K) { return ((K + 1 + 15) / 16) * 16; }   // + ones column, k-step 16

struct TcSmem {
  uint32_t rsA, rsX, rsG, bytesA, bytesX, bytesG;
  uint32_t offAh, offAl, offXh, offXl, offGh, offGl, offY, offLL, offRow, offItc, offCs, total;
};
__host__ __device__ inline TcSmem tc_smem(int K) {
  const int KP = tc_kp(K);
  TcSmem s;
  s.rsA = (kTcTN / 8) * kCS;              // a tile: rows k (KP), cols j (128)
  s.rsX = (uint32_t)(KP / 8) * kCS;       // x tile: rows person-sample (128), cols k (KP)
  s.rsG = (kTcTN / 8) * kCS;              // G tile: rows person-sample (128), cols j (128)
  s.bytesA = (uint32_t)(KP / 8) * s.rsA;
  s.bytesX = (kTcTM / 8) * s.rsX;
  s.bytesG = (kTcTM / 8) * s.rsG;
  uint32_t o = 0;
  s.offAh = o; o += s.bytesA;
  s.offAl = o; o += s.bytesA;
  s.offXh = o; o += s.bytesX;
  s.offXl = o; o += s.bytesX;             // the M = 128 read of x^T runs up to 2 KB past a tile: lands in Xl / Gh
  s.offGh = o; o += s.bytesG;
  s.offGl = o; o += s.bytesG;
  s.offY = o; o += 2 * kTcTM * kTcYStride;       // response tile, double-buffered
  s.offLL = o; o += 2 * kTcTM * 4;
  s.offRow = o; o += kTcTM * 8;
  s.offItc = o; o += 3 * kTcTN * 4;              // c | d - c | 1 - d per item (3PL / 4PL)
  s.offCs = o; o += 4 * 2 * kTcTN * 4;           // [quadrant][d_c | d_d][item] column sums at the end
  s.total = o;
  return s;
}

__device__ __forceinline__ uint32_t pack_half2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
// 8 floats -> one 16-byte chunk of fp16 "hi" and one of fp16 "lo"
__device__ __forceinline__ void split_store8(const float (&v)[8], unsigned char* hi_dst, unsigned char* lo_dst) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    __half2 hh = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
    const float2 hf = __half22float2(hh);
    h[i] = *reinterpret_cast<uint32_t*>(&hh);
    l[i] = pack_half2(v[2 * i] - hf.x, v[2 * i + 1] - hf.y);
  }
  *reinterpret_cast<uint4*>(hi_dst) = make_uint4(h[0], h[1], h[2], h[3]);
  *reinterpret_cast<uint4*>(lo_dst) = make_uint4(l[0], l[1], l[2], l[3]);
}

__device__ __forceinline__ void split8_regs(const float (&v)[8], uint4& hi, uint4& lo) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    __half2 hh = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
    const float2 hf = __half22float2(hh);
    h[i] = *reinterpret_cast<uint32_t*>(&hh);
    l[i] = pack_half2(v[2 * i] - hf.x, v[2 * i + 1] - hf.y);
  }
  hi = make_uint4(h[0], h[1], h[2], h[3]);
  lo = make_uint4(l[0], l[1], l[2], l[3]);
}

// mbarrier slots of the pipeline (one completion per row tile each; the response ring has one pair per buffer)
enum { kBarXFull = 0,   // x producers -> MMA: x tile staged                                    (8 arrivals)  
       kBarXFree,       // MMA commit -> x producers: backward MMAs done, x and G tiles reusable
       kBarLFull,       // MMA commit -> epilogue: logits in TMEM
       kBarGFull,       // epilogue -> MMA: G tile written, logits consumed                     (8 arrivals)  
       kBarDxFull,      // MMA commit -> x producers (+ epilogue at the end): d x / d a in TMEM
       kBarDxFree,      // x producers -> MMA: d x drained to global memory                     (8 arrivals)  
       kBarYFull0, kBarYFull1,   // y producers -> epilogue: response buffer b staged           (4 arrivals)  
       kBarYFree0, kBarYFree1,   // epilogue -> y producers: response buffer b consumed         (8 arrivals)  
       kNumBars };

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(umma::smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void warp_arrive(uint64_t* bar, int lane) {
  __syncwarp();
  if (lane == 0) mbar_arrive(bar);
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// Sum of v[c] over the 32 lanes of the warp for each of 16 columns; on return every lane l holds the
// total of column l & 15.  Fixed butterfly order -> deterministic.
__device__ __forceinline__ float colsum16(float (&v)[16], int lane) {
#pragma unroll
  for (int j = 0; j < 16; ++j) v[j] += __shfl_xor_sync(0xffffffffu, v[j], 16);
#pragma unroll
  for (int off = 8; off >= 1; off >>= 1) {
    const bool up = lane & off;
#pragma unroll
    for (int j = 0; j < off; ++j) {
      const float keep = up ? v[j + off] : v[j];
      const float send = up ? v[j] : v[j + off];
      v[j] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  return v[0];
}

// Warp roles: warps 0-7 epilogue (TMEM lane quadrant = warp & 3, column half = warp >> 2),
// warps 8-11 producers (stage the next x / response tiles while the epilogue runs), warp 12 issues the MMAs.
template <int MODEL>
__global__ void __launch_bounds__(kTcThreads, 1) mirt_tc_lik_kernel(const MirtParams p, const int n_jt, const int G) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t bars[kNumBars];
  __shared__ uint32_t tmem_base_sm;
  const int K = p.K, J = p.J, KP = tc_kp(K);
  const TcSmem L = tc_smem(K);
  unsigned char* Ah = smem + L.offAh;
  unsigned char* Al = smem + L.offAl;
  unsigned char* Xh = smem + L.offXh;
  unsigned char* Xl = smem + L.offXl;
  unsigned char* Gh = smem + L.offGh;
  unsigned char* Gl = smem + L.offGl;
  float* llsm_all = reinterpret_cast<float*>(smem + L.offLL); // [2][128]: column-half 1's log-lik partial, by tile parity

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int jt = blockIdx.x % n_jt, g = blockIdx.x / n_jt;
  const int j0 = jt * kTcTN;
  const int64_t R = (int64_t)p.S * p.B;
  const int64_t ntile_m = (R + kTcTM - 1) / kTcTM;
  const int nck = KP / 8;                                      // 16-byte chunks per x row

  // ---- one-time: a tile (rows k < K: a[k][j]; row K: b[j]; rows > K: 0), barriers, TMEM ----------
  for (int e = tid; e < KP * (kTcTN / 8); e += kTcThreads) {
    const int k = e / (kTcTN / 8), c8 = e - k * (kTcTN / 8);
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int j = j0 + c8 * 8 + i;
      float t = 0.f;
      if (j < J) {
        if (k < K) t = p.a[(size_t)k * J + j];
        else if (k == K) t = p.b[j];
      }
      v[i] = t;
    }
    const uint32_t o = umma::tile_off(k, c8 * 8, L.rsA, kCS);
    split_store8(v, Ah + o, Al + o);
  }
  float* itc = reinterpret_cast<float*>(smem + L.offItc);
  if (MODEL >= VIPSY_IRT_3PL && tid < kTcTN) {
    const int j = j0 + tid;
    const bool ok = j < J;
    const float cj = ok ? p.c[j] : 0.f;
    const float dj = (ok && MODEL == VIPSY_IRT_4PL) ? p.d[j] : 1.f;
    itc[tid] = cj;
    itc[kTcTN + tid] = dj - cj;
    itc[2 * kTcTN + tid] = 1.f - dj;
  }
  if (warp == kTcMmaWarp) umma::tmem_alloc(&tmem_base_sm, kTmemCols);
  if (tid == 0) {
    // one arrival per WARP everywhere (all lanes fence, lane 0 signals): hundreds of separate arrivals on one mbarrier
    // serialise on the shared-memory atomic unit -- 256 of them cost 1-2 k cycles per phase
    umma::mbar_init(&bars[kBarXFull], kTcXThreads / 32);
    umma::mbar_init(&bars[kBarYFull0], kTcYThreads / 32);
    umma::mbar_init(&bars[kBarYFull1], kTcYThreads / 32);
    umma::mbar_init(&bars[kBarYFree0], kTcEpiThreads / 32);
    umma::mbar_init(&bars[kBarYFree1], kTcEpiThreads / 32);
    umma::mbar_init(&bars[kBarXFree], 1);
    umma::mbar_init(&bars[kBarLFull], 1);
    umma::mbar_init(&bars[kBarGFull], kTcEpiThreads / 32);
    umma::mbar_init(&bars[kBarDxFull], 1);
    umma::mbar_init(&bars[kBarDxFree], kTcXThreads / 32);
    umma::fence_mbar_init();
  }
  umma::fence_proxy_async();
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem = tmem_base_sm;

  if (warp >= kTcEpiThreads / 32 && warp < (kTcEpiThreads + kTcXThreads) / 32) {
    // =========================== X PRODUCERS =========================================================
    // Per tile n: put the global loads of x tile n in flight (registers) while tile n-1 is still being
    // computed, wait for the single x buffer to be released by the backward MMAs of tile n-1, convert to
    // fp16 hi/lo + store, then drain d x of tile n-1 from TMEM to global memory.
    const int ptid = tid - kTcEpiThreads;
    const int pquad = warp & 3, phalf = (warp >> 2) & 1;     // TMEM lane quadrant / column interleave of this warp
    const uint32_t ptmem_row = tmem + ((uint32_t)(pquad * 32) << 16);
    constexpr int kPre = 8;                                  // 32-byte chunks per thread (16 x 128 / 256)
    const bool xvec = (K & 3) == 0;

    // chunk e of the tile -> (row, 8 columns): the 8 lanes of a quarter-warp take 8 consecutive ROWS of one column
    // group, so that their 16-byte pieces of the operand tile are one contiguous 128-byte core matrix (with
    // consecutive columns per lane the pieces lie 128 B apart: an 8-way bank conflict on every store); the four
    // quarter-warps take four consecutive column groups of those rows, so that a warp's load touches 8 lines of
    // 128 contiguous bytes (one row each) instead of 32 lines (measured: 8.7 k -> 2 k cycles to load a tile).
    // e = ((cg * 16 + rb) * 4 + cq) * 8 + rl:  row = 8 rb + rl,  column group = 4 cg + cq.
    auto chunk_rc = [&](int e, int& pr, int& c8) {
      const int rl = e & 7, cq = (e >> 3) & 3, rb = (e >> 5) & 15, cg = e >> 9;
      pr = 8 * rb + rl;
      c8 = 4 * cg + cq;
    };
    auto load_chunk = [&](int e, int64_t r0, float (&v)[8]) {
      int pr, c8;
      chunk_rc(e, pr, c8);
      const int64_t r = r0 + pr;
      if (c8 < nck && r < R && !(p.dbg & 4)) {
        const float* src = p.x + r * K + c8 * 8;
        if (xvec && c8 * 8 + 8 <= K) {
          const float4 u0 = *reinterpret_cast<const float4*>(src);
          const float4 u1 = *reinterpret_cast<const float4*>(src + 4);
          v[0] = u0.x; v[1] = u0.y; v[2] = u0.z; v[3] = u0.w;
          v[4] = u1.x; v[5] = u1.y; v[6] = u1.z; v[7] = u1.w;
        } else if (xvec && c8 * 8 + 4 == K) {                 // the chunk that holds the last 4 columns and the ones column
          const float4 u0 = *reinterpret_cast<const float4*>(src);
          v[0] = u0.x; v[1] = u0.y; v[2] = u0.z; v[3] = u0.w;
          v[4] = 1.f; v[5] = 0.f; v[6] = 0.f; v[7] = 0.f;
        } else if (c8 * 8 > K) {
#pragma unroll
          for (int i = 0; i < 8; ++i) v[i] = 0.f;
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int k = c8 * 8 + i;
            v[i] = k < K ? src[i] : (k == K ? 1.f : 0.f);
          }
        }
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = 0.f;
      }
    };
    auto drain_dx = [&](int64_t r0) {                        // this thread's row of d x: TMEM -> global
      const int dpr = pquad * 32 + lane;
      float* dst = p.gx + ((int64_t)jt * R + r0 + dpr) * K;
      const bool rvalid = r0 + dpr < R && !(p.dbg & 16);
      for (int c0 = phalf * 16; c0 < K; c0 += 32) {          // the two warps of a quadrant interleave 16-column groups
        float v[16];
        umma::tmem_ld16(ptmem_row + kColDx + c0, v);
        if (rvalid) {
          if (xvec) {
#pragma unroll
            for (int i = 0; i < 16; i += 4)
              if (c0 + i < K) *reinterpret_cast<float4*>(dst + c0 + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i)
              if (c0 + i < K) dst[c0 + i] = v[i];
          }
        }
      }
      umma::fence_before_sync();
      warp_arrive(&bars[kBarDxFree], lane);
    };

    int64_t n = 0, r0_prev = 0;
    for (int64_t tm = g; tm < ntile_m; tm += G, ++n) {
      const int64_t r0 = tm * kTcTM;
      if (ptid == 0) TC_STAMP(0, 0);
      // loaded AND split into the fp16 hi / lo pieces before the buffer is waited for: what is left on the critical
      // path between "x buffer free" and "x tile staged" are the 16 stores (the split is ~200 instructions per thread).
      // (Issuing these loads a stage earlier, before the d x drain, needs 64 more live registers: spills at 672 threads.)
      uint4 xh[kPre], xl[kPre];
      {
        float xv[kPre][8];                                   // every load in flight before the first split needs one
#pragma unroll
        for (int u = 0; u < kPre; ++u) load_chunk(ptid + u * kTcXThreads, r0, xv[u]);
#pragma unroll
        for (int u = 0; u < kPre; ++u) split8_regs(xv[u], xh[u], xl[u]);
      }
      if (ptid == 0) TC_STAMP(0, 1);
      // ---- the single x buffer is free once the backward MMAs of tile n-1 have completed ---------------
      if (n > 0) umma::mbar_wait(&bars[kBarXFree], (uint32_t)((n - 1) & 1));
      if (ptid == 0) TC_STAMP(0, 2);
#pragma unroll
      for (int u = 0; u < kPre; ++u) {
        const int e = ptid + u * kTcXThreads;
        int pr, c8;
        chunk_rc(e, pr, c8);
        if (c8 < nck) {
          const uint32_t o = umma::tile_off(pr, c8 * 8, L.rsX, kCS);
          *reinterpret_cast<uint4*>(Xh + o) = xh[u];
          *reinterpret_cast<uint4*>(Xl + o) = xl[u];
        }
      }
      umma::fence_proxy_async();                            // generic-proxy writes -> visible to the tensor core
      warp_arrive(&bars[kBarXFull], lane);
      if (ptid == 0) TC_STAMP(0, 3);
      // ---- d x of tile n-1 (kBarXFree and kBarDxFull complete together) --------------------------------
      if (n > 0) {
        umma::mbar_wait(&bars[kBarDxFull], (uint32_t)((n - 1) & 1));
        umma::fence_after_sync();
        drain_dx(r0_prev);
      }
      r0_prev = r0;
    }
    if (n > 0) {
      umma::mbar_wait(&bars[kBarDxFull], (uint32_t)((n - 1) & 1));
      umma::fence_after_sync();
      drain_dx(r0_prev);
    }
  } else if (warp >= (kTcEpiThreads + kTcXThreads) / 32 && warp < kTcMmaWarp) {
    // =========================== Y PRODUCERS =========================================================
    // Response tile n -> ring buffer n & 1, as far ahead of the epilogue as the two buffers allow.
    const int ptid = tid - kTcEpiThreads - kTcXThreads;
    const bool y_words = ((p.ld_y & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.y) & 3) == 0);
    int64_t* rowsm = reinterpret_cast<int64_t*>(smem + L.offRow);
    int64_t n = 0;
    for (int64_t tm = g; tm < ntile_m; tm += G, ++n) {
      const int64_t r0 = tm * kTcTM;
      const int buf = (int)(n & 1);
      unsigned char* ytile = smem + L.offY + buf * (kTcTM * kTcYStride);
      // y-row index of each tile row (one 64-bit modulo per row, not per cell)
      {
        const int64_t r = r0 + ptid;
        int64_t row = -1;
        if (r < R) {
          const int64_t bb = r % p.B;
          row = p.idx ? p.idx[bb] : bb;
        }
        named_bar_sync(1, kTcYThreads);                      // previous tile's readers of rowsm are done
        rowsm[ptid] = row;
      }
      named_bar_sync(1, kTcYThreads);
      if (y_words) {
        // buffer `buf` was last read by the epilogue of tile n-2 (waited for after the first loads are in flight)
#pragma unroll 1
        for (int h = 0; h < 2; ++h) {
          uint32_t code4[16];                                // 16 words in flight per thread and round
#pragma unroll
          for (int u = 0; u < 16; ++u) {
            const int e = ptid + (h * 16 + u) * kTcYThreads;
            const int pr = e / (kTcTN / 4), w = e - pr * (kTcTN / 4);
            const int64_t row = rowsm[pr];
            const int j = j0 + w * 4;
            uint32_t c4 = 0xffffffffu;
            if (row >= 0 && j < J) {
              const unsigned char* src = p.y + row * p.ld_y + j;
              if (j + 4 <= J) {
                c4 = *reinterpret_cast<const uint32_t*>(src);
              } else {                                        // ragged tail word of the last item tile
                c4 = 0;
#pragma unroll
                for (int i = 0; i < 4; ++i) c4 |= (uint32_t)(j + i < J ? src[i] : 255u) << (8 * i);
              }
            }
            code4[u] = c4;
          }
          if (h == 0 && n >= 2) umma::mbar_wait(&bars[kBarYFree0 + buf], (uint32_t)(((n >> 1) - 1) & 1));
#pragma unroll
          for (int u = 0; u < 16; ++u) {
            const int e = ptid + (h * 16 + u) * kTcYThreads;
            const int pr = e / (kTcTN / 4), w = e - pr * (kTcTN / 4);
            *reinterpret_cast<uint32_t*>(ytile + pr * kTcYStride + w * 4) = code4[u];
          }
        }
      } else {
        if (n >= 2) umma::mbar_wait(&bars[kBarYFree0 + buf], (uint32_t)(((n >> 1) - 1) & 1));
        for (int e = ptid; e < kTcTM * kTcTN; e += kTcYThreads) {
          const int pr = e / kTcTN, jj = e - pr * kTcTN;
          const int64_t row = rowsm[pr];
          unsigned char code = 255;
          if (row >= 0 && j0 + jj < J) code = p.y[row * p.ld_y + j0 + jj];
          ytile[pr * kTcYStride + jj] = code;
        }
      }
      warp_arrive(&bars[kBarYFull0 + buf], lane);
    }
  } else if (warp == kTcMmaWarp) {
    // =========================== MMA ISSUER ===========================================================
    // the whole warp walks the loop (every lane waits), one elected lane issues: see umma::elect_one
    {
      const uint32_t idesc_fwd = umma::make_idesc_f16(kTcTM, kTcTN, umma::kMajorK, umma::kMajorMN);
      const uint32_t idesc_dx = umma::make_idesc_f16(kTcTM, KP, umma::kMajorK, umma::kMajorK);
      const uint32_t idesc_da = umma::make_idesc_f16(128, kTcTN, umma::kMajorMN, umma::kMajorMN);
      const uint32_t sAh = umma::smem_u32(Ah), sAl = umma::smem_u32(Al), sXh = umma::smem_u32(Xh),
                     sXl = umma::smem_u32(Xl), sGh = umma::smem_u32(Gh), sGl = umma::smem_u32(Gl);
      int64_t n = 0;
      for (int64_t tm = g; tm < ntile_m; tm += G, ++n) {
        const uint32_t par = (uint32_t)(n & 1);
        // forward: logits = [x|1] [a;b]  (A operand K-major from the x tile, B MN-major from the a tile)
        umma::mbar_wait(&bars[kBarXFull], par);
        umma::fence_after_sync();
        uint32_t acc = 0;
        if (umma::elect_one()) {
        TC_STAMP(1, 0);
        for (int ks = 0; ks < KP / 16; ++ks) {
#pragma unroll
          for (int term = 0; term < 3; ++term) {             // hi*hi, hi*lo, lo*hi
            const uint64_t da = umma::make_desc((term == 2 ? sXl : sXh) + ks * 2 * kCS, kCS, L.rsX);
            const uint64_t db = umma::make_desc((term == 1 ? sAl : sAh) + ks * 2 * L.rsA, L.rsA, kCS);
            umma::mma_f16_ss(tmem + kColLogit, da, db, idesc_fwd, acc);
            acc = 1;
          }
        }
        umma::commit(&bars[kBarLFull]);
        TC_STAMP(1, 1);
        }
        __syncwarp();
        // backward: d x = G a^T (cols kColDx..), d a += x^T G (cols kColDa.., lives across row tiles)
        umma::mbar_wait(&bars[kBarGFull], par);
        if (n > 0) umma::mbar_wait(&bars[kBarDxFree], (uint32_t)((n - 1) & 1));
        umma::fence_after_sync();
        if (umma::elect_one()) {
        TC_STAMP(1, 2);
        acc = 0;
        if (!(p.dbg & 8))
        for (int ks = 0; ks < kTcTN / 16; ++ks) {
#pragma unroll
          for (int term = 0; term < 3; ++term) {
            const uint64_t da = umma::make_desc((term == 2 ? sGl : sGh) + ks * 2 * kCS, kCS, L.rsG);
            const uint64_t db = umma::make_desc((term == 1 ? sAl : sAh) + ks * 2 * kCS, kCS, L.rsA);
            umma::mma_f16_ss(tmem + kColDx, da, db, idesc_dx, acc);
            acc = 1;
          }
        }
        acc = (n % kTcFlushTiles) != 0 ? 1u : 0u;          // a new group starts a fresh accumulator
        const uint32_t col_da = kColDa + (uint32_t)((n / kTcFlushTiles) & 1) * 128u;
        if (!(p.dbg & 8))
        for (int ks = 0; ks < kTcTM / 16; ++ks) {
#pragma unroll
          for (int term = 0; term < 3; ++term) {
            const uint64_t da = umma::make_desc((term == 2 ? sXl : sXh) + ks * 2 * L.rsX, L.rsX, kCS);
            const uint64_t db = umma::make_desc((term == 1 ? sGl : sGh) + ks * 2 * L.rsG, L.rsG, kCS);
            umma::mma_f16_ss(tmem + col_da, da, db, idesc_da, acc);
            acc = 1;
          }
        }
        umma::commit(&bars[kBarXFree]);
        umma::commit(&bars[kBarDxFull]);
        TC_STAMP(1, 3);
        }
        __syncwarp();
      }
    }
  } else {
    // =========================== EPILOGUE (256 threads) ===============================================
    const int quad = warp & 3, half = warp >> 2;             // TMEM lane quadrant, column half
    const uint32_t tmem_row = tmem + ((uint32_t)(quad * 32) << 16);
    const int pr = quad * 32 + lane;                         // this thread's tile row
    const float negDl2e = -p.D * kLog2e;
    float cs_acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};   // lane (l & 15): column l & 15 of chunk cc: d_c | d_d
    // d a / d b of one finished group of row tiles -> this CTA's partial row (group 0 stores, later groups add)
    auto flush_da = [&](uint32_t taddr, int64_t grp) {
      const int k = quad * 32 + lane;
      float* part = p.item_partial + (size_t)g * ((size_t)K * J + 3 * (size_t)J);
      umma::fence_after_sync();
      for (int cc = 0; cc < 4; ++cc) {
        const int c0 = half * 64 + cc * 16;
        float v[16];
        float* dst = part + (size_t)(k <= K ? k : 0) * J + j0 + c0;        // k == K lands on the d_b segment (offset K*J)
        const bool vec = (J & 3) == 0 && j0 + c0 + 16 <= J && k <= K;      // 16-byte pieces: 4 requests instead of 16
        float4 old[4];
#pragma unroll
        for (int q = 0; q < 4; ++q)                            // the old partial sums are on their way before TMEM is read
          old[q] = (vec && grp > 0) ? *reinterpret_cast<const float4*>(dst + 4 * q) : make_float4(0.f, 0.f, 0.f, 0.f);
        umma::tmem_ld16(taddr + c0, v);
        if (vec) {
#pragma unroll
          for (int q = 0; q < 4; ++q)
            *reinterpret_cast<float4*>(dst + 4 * q) = make_float4(old[q].x + v[4 * q], old[q].y + v[4 * q + 1],
                                                                  old[q].z + v[4 * q + 2], old[q].w + v[4 * q + 3]);
        } else if (k <= K) {
#pragma unroll
          for (int i = 0; i < 16; ++i)
            if (j0 + c0 + i < J) dst[i] = grp > 0 ? dst[i] + v[i] : v[i];
        }
      }
      umma::fence_before_sync();
    };
    int64_t n = 0;
    for (int64_t tm = g; tm < ntile_m; tm += G, ++n) {
      const int64_t r0 = tm * kTcTM;
      const uint32_t par = (uint32_t)(n & 1);
      const unsigned char* ytile = smem + L.offY + (n & 1) * (kTcTM * kTcYStride);
      umma::mbar_wait(&bars[kBarYFull0 + (n & 1)], (uint32_t)((n >> 1) & 1));   // response tile n staged
      umma::mbar_wait(&bars[kBarLFull], par);
      umma::fence_after_sync();
      if (tid == 0) TC_STAMP(2, 0);
      float ll = 0.f;
#pragma unroll 1
      for (int cc = 0; cc < 4; ++cc) {
        const int c0 = half * 64 + cc * 16;
        float v[16];
        umma::tmem_ld16(tmem_row + kColLogit + c0, v);
        uint32_t yw[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) yw[i] = *reinterpret_cast<const uint32_t*>(ytile + pr * kTcYStride + c0 + i * 4);
        float csc[16], csd[16];                              // 3PL / 4PL: this row's d lb/dc, d lb/dd per column
        if constexpr (MODEL == VIPSY_IRT_2PL) {
          // 2PL: no per-cell selects (the technique of the K = 1 row kernels, irt_uni_row.cu).  sg = +2 (y = 1),
          // -2 (y = 0), 0 (missing) is assembled from the response byte with integer ops (top byte of the float
          // 0x40 / 0xc0 / 0x00); t = (sg D / 2)(-log2e l) is the logit signed by the response and 0 for a missing
          // cell, whose sp = 1/2 multiplies the likelihood product by exactly 2^-1 (added back through n_half) while
          // its gradient (sg D / 2) e sp vanishes by itself.  The probability clamp (sp < eps <=> t >= 23) comes from
          // the min on t; the gradient of a clamped cell is switched off in a rare branch, one test per word.
          // 16 instructions per cell instead of 29: the epilogue is bound by issue slots (8 warps, 64 cells a thread).
          const float halfD = 0.5f * p.D;
          int n_half = 0;
          float gl16[16];
#pragma unroll
          for (int w = 0; w < 4; ++w) {
            const uint32_t word = yw[w];
            n_half += __popc(word & 0x02020202u);
            const uint32_t nw = ~word;
            const uint32_t T = ((nw << 7) & 0x80808080u) | (nw & 0x40404040u);
            float sgD[4], tt[4], es[4];
            sgD[0] = halfD * __uint_as_float(__byte_perm(T, 0u, 0x0444u));
            sgD[1] = halfD * __uint_as_float(__byte_perm(T, 0u, 0x1444u));
            sgD[2] = halfD * __uint_as_float(__byte_perm(T, 0u, 0x2444u));
            sgD[3] = halfD * __uint_as_float(__byte_perm(T, 0u, 0x3444u));
            float prod = 1.f;                                // one lg2 per 4 cells: each factor >= 1.19e-7
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              tt[i] = sgD[i] * (-kLog2e * v[4 * w + i]);
              const float e = ex2_approx(fminf(tt[i], 23.f));
              const float sp = rcp_approx(1.f + e);           // s (y=1) | 1-s (y=0) | 1/2 (missing)
              es[i] = e * sp;
              prod *= sp;
            }
            if (fmaxf(fmaxf(tt[0], tt[1]), fmaxf(tt[2], tt[3])) >= 23.f) {
#pragma unroll
              for (int i = 0; i < 4; ++i)
                if (tt[i] >= 23.f) es[i] = 0.f;
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) gl16[4 * w + i] = sgD[i] * es[i];
            ll = fmaf(kLn2, lg2_approx(prod), ll);
          }
          ll = fmaf(kLn2, (float)n_half, ll);
#pragma unroll
          for (int c8 = 0; c8 < 2; ++c8) {
            float gl8[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) gl8[i] = gl16[8 * c8 + i];
            const uint32_t o = umma::tile_off(pr, c0 + c8 * 8, L.rsG, kCS);
            split_store8(gl8, Gh + o, Gl + o);
          }
        } else
#pragma unroll
        for (int c8 = 0; c8 < 2; ++c8) {
          float gl8[8];
#pragma unroll
          for (int h4 = 0; h4 < 2; ++h4) {
            float prod = 1.f;                                // one lg2 per 4 cells: each factor >= 1.19e-7
#pragma unroll
            for (int i4 = 0; i4 < 4; ++i4) {
              const int i = h4 * 4 + i4;
              const int cidx = c8 * 8 + i;
              const uint32_t code = (yw[cidx >> 2] >> ((cidx & 3) * 8)) & 0xffu;
              // branch-free: every cell runs the same chain (the chains interleave), missing cells are
              // neutralised by selects at the end
              const bool valid = code != 255u, y1 = code == 1u;
              float t = negDl2e * v[cidx];                    // -l log2e, l = D (x.a + b)
              t = y1 ? t : -t;
              t = fminf(t, 126.f);
              const float e = ex2_approx(t);
              const float sp = rcp_approx(1.f + e);           // s (y=1) | 1-s (y=0)
              const float es = e * sp;
              float gl;
              if (MODEL >= VIPSY_IRT_3PL) {                   // q = c + (d-c) s  |  (1-d) + (d-c)(1-s)   (vi.py:44-66)
                const int jl = c0 + cidx;
                const float dmc = itc[kTcTN + jl];
                const float base = y1 ? itc[jl] : itc[2 * kTcTN + jl];
                const float q = fmaf(dmc, sp, base);
                const float qc = fminf(fmaxf(q, kEps32), kOneMinusEps32);
                prod *= valid ? qc : 1.f;
                const bool live = valid && q == qc;           // clamp fired: zero gradient
                const float r = rcp_approx(qc);
                const float us = r * sp, uc = r * es;
                gl = live ? (y1 ? p.D : -p.D) * dmc * us * es : 0.f;
                csc[cidx] = live ? (y1 ? uc : -us) : 0.f;
                csd[cidx] = live ? (y1 ? us : -uc) : 0.f;
              } else {
                const float qc = fminf(fmaxf(sp, kEps32), kOneMinusEps32);
                prod *= valid ? qc : 1.f;
                gl = (valid && sp == qc) ? (y1 ? p.D : -p.D) * es : 0.f;
              }
              gl8[i] = gl;
            }
            ll = fmaf(kLn2, lg2_approx(prod), ll);
          }
          const uint32_t o = umma::tile_off(pr, c0 + c8 * 8, L.rsG, kCS);
          split_store8(gl8, Gh + o, Gl + o);
        }
        if (MODEL >= VIPSY_IRT_3PL) {                         // column sums over the warp's 32 rows (fixed order)
          const float tc = colsum16(csc, lane);
          cs_acc[cc] += tc;
          if (MODEL == VIPSY_IRT_4PL) cs_acc[4 + cc] += colsum16(csd, lane);
        }
      }
      umma::fence_before_sync();
      umma::fence_proxy_async();
      warp_arrive(&bars[kBarGFull], lane);
      warp_arrive(&bars[kBarYFree0 + (n & 1)], lane);
      if (tid == 0) TC_STAMP(2, 1);
      // first tile of a new group: the previous group's d a is complete (the logits of this tile were committed after
      // every earlier MMA) and nothing writes that accumulator before the group after next -> add it to the partial row
      if (n > 0 && n % kTcFlushTiles == 0) {
        flush_da(tmem_row + kColDa + (uint32_t)(((n / kTcFlushTiles) - 1) & 1) * 128u, n / kTcFlushTiles - 1);
      }
      // per-row log-likelihood partial of this item tile: column half 1 hands its sum to half 0
      float* llsm = llsm_all + (n & 1) * kTcTM;
      if (half == 1) llsm[pr] = ll;
      named_bar_sync(2, kTcEpiThreads);
      if (half == 0 && r0 + pr < R) p.ell[(int64_t)jt * R + r0 + pr] = ll + llsm[pr];

    }
    // every MMA of the CTA must have completed before d a is read
    if (n > 0) umma::mbar_wait(&bars[kBarDxFull], (uint32_t)((n - 1) & 1));
    umma::fence_after_sync();

    // ---- flush the last group of d a (rows k < K) and d b (row K) of this CTA's item tile: the last kBarDxFull
    //      wait above covers every MMA of the CTA ------------------------------------------------------------
    {
      float* part = p.item_partial + (size_t)g * ((size_t)K * J + 3 * (size_t)J);
      if (n > 0) {
        const int64_t grp = (n - 1) / kTcFlushTiles;
        flush_da(tmem_row + kColDa + (uint32_t)(grp & 1) * 128u, grp);
      } else {
        const int k = quad * 32 + lane;
        for (int c = half * 64; c < half * 64 + 64; ++c)
          if (k <= K && j0 + c < J) part[(size_t)k * J + j0 + c] = 0.f;
      }
      if (MODEL >= VIPSY_IRT_3PL) {                           // d_c, d_d: combine the four row quadrants in order
        float* cs = reinterpret_cast<float*>(smem + L.offCs);
        if (lane < 16) {
#pragma unroll
          for (int cc = 0; cc < 4; ++cc) {
            cs[(quad * 2 + 0) * kTcTN + half * 64 + cc * 16 + lane] = cs_acc[cc];
            cs[(quad * 2 + 1) * kTcTN + half * 64 + cc * 16 + lane] = cs_acc[4 + cc];
          }
        }
        named_bar_sync(2, kTcEpiThreads);
        if (tid < 2 * kTcTN) {
          const int which = tid / kTcTN, jl = tid - which * kTcTN;
          if (j0 + jl < J && (which == 0 || MODEL == VIPSY_IRT_4PL)) {
            float t = 0.f;
#pragma unroll
            for (int q = 0; q < 4; ++q) t += cs[(q * 2 + which) * kTcTN + jl];
            part[(size_t)K * J + (size_t)(1 + which) * J + j0 + jl] = t;
          }
        }
      }
    }
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == kTcMmaWarp) umma::tmem_dealloc(tmem, kTmemCols);
}

// ----------------------------------------------------------------------------------------------------
int mirt_tc_parts(int J) { return (J + kTcTN - 1) / kTcTN; }

bool mirt_tc_supported(const MirtParams& p) {
  return p.model >= VIPSY_IRT_2PL && p.model <= VIPSY_IRT_4PL && p.K >= 8 && tc_kp(p.K) <= 128 && tc_smem(p.K).total + 1024u <= 227u * 1024u;   // + static barriers
}

int launch_lik_tc(const MirtParams& p, int* rows_out, int* parts_out, cudaStream_t st) {
  const int n_jt = mirt_tc_parts(p.J);
  const int64_t ntile_m = ((int64_t)p.S * p.B + kTcTM - 1) / kTcTM;
  int64_t G = num_sms() / n_jt;
  if (G < 1) G = 1;
  if (G > ntile_m) G = ntile_m;
  const TcSmem L = tc_smem(p.K);
  VIPSY_REQUIRE(L.total <= 227 * 1024, VIPSY_ERR_UNSUPPORTED, "mirt_tc: shared memory %u B exceeds 227 KB", L.total);
  auto kern = p.model == VIPSY_IRT_2PL ? mirt_tc_lik_kernel<VIPSY_IRT_2PL>
              : (p.model == VIPSY_IRT_3PL ? mirt_tc_lik_kernel<VIPSY_IRT_3PL> : mirt_tc_lik_kernel<VIPSY_IRT_4PL>);
  VIPSY_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total));
  kern<<<(int)(G * n_jt), kTcThreads, L.total, st>>>(p, n_jt, (int)G);
  VIPSY_CUDA_CHECK(cudaGetLastError());
  if (p.dbg & 32) {
    long long h[3 * 8 * 4];
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(h, g_tc_dbg, sizeof(h));
    const long long t0 = h[0];
    const char* names[3] = {"producer: start | y staged | x free | x staged", "mma: x full | fwd issued | G full | bwd issued",
                            "epilogue: logits | G done | dx full | dx stored"};
    for (int r = 0; r < 3; ++r) {
      fprintf(stderr, "%s\n", names[r]);
      for (int n = 0; n < 6; ++n)
        fprintf(stderr, "  tile %d: %8lld %8lld %8lld %8lld\n", n, h[(r * 8 + n) * 4] - t0, h[(r * 8 + n) * 4 + 1] - t0,
                h[(r * 8 + n) * 4 + 2] - t0, h[(r * 8 + n) * 4 + 3] - t0);
    }
  }
  *rows_out = (int)G;
  *parts_out = n_jt;
  return VIPSY_OK;
}

}  // namespace vipsy
