// Amortised multivariate-normal guide of VaeIRT, K > 1 (vi.py:685-693), with MvnEncoder's tail fused in:
//
//   raw = fc22(hidden)                       vi.py:444, 451     B x K(K+1)/2, packed in torch.tril_indices order
//   L   = tril(raw, -1) + diag(exp(raw_ii))  vi.py:452-454 + LowerCholeskyTransform (vi.py:693)
//   x   = loc + L eps                        MultivariateNormal(loc, scale_tril=L).rsample()
//
// forward  (mvn_fwd_kernel):  raw[128 persons x 128 factor columns] = [hidden | 1] [W22 | b22]^T on tcgen05 (fp16
//          hi/lo split of both operands, 3 MMAs per 16-deep k-step, fp32 accumulation in TMEM).  The epilogue thread
//          of a person reads its factor out of tensor memory 64 columns at a time and folds it into x on the fly: raw never
//          reaches shared or global memory.  The factor columns are ordered by 8 x 8 blocks of L (mv_col): 64
//          columns are one 8-vector of the noise against eight rows -- eight independent accumulators, the noise
//          read once per 64 products, the diagonal at fixed slots.  fc22's weights arrive as a pre-split operand image (mvn_prep_kernel, once per
//          step, 1.7 MB, L2-resident) through one bulk copy per 128-column chunk.
// backward (mvn_bwd_kernel):  V[person][(i, j)] = g_i eps_j (the gradient of raw, with the exp chain rule on the
//          diagonal) is generated tile by tile in shared memory, split into fp16 hi/lo, and feeds two products
//          from the same bytes:  d W22[(i,j)][h] += V^T [hidden | 1]  (accumulated in TMEM over all person tiles of
//          the CTA) and  d hidden[person][h] += V (64 W22).  CTAs own a slab of factor columns x a group of
//          persons; partial sums are reduced in a fixed order (deterministic).
#include "enc_mvn.cuh"
#include "umma.cuh"

namespace vipsy {
namespace {

constexpr uint32_t kCSb = 128;                 // byte stride between core matrices along the contiguous dimension
constexpr int kFwEpiWarps = 8;                 // warps 0-7: epilogue, two per TMEM lane quadrant (warp & 3), one unit of a chunk each
constexpr int kFwMmaWarps = 2;                 // MMA-issuing warps (alternate chunks)
constexpr int kFwThreads = (kFwEpiWarps + 1 + kFwMmaWarps) * 32;   // + a bulk-copy producer warp + the MMA warps
constexpr int kFwTmemBufs = 3;                 // 3 x 128 accumulator columns; columns 384.. : the [hidden | 1] operand
constexpr uint32_t kFwACol = 384;              // A operand in tensor memory: KH / 2 words of fp16 hi pairs, then KH / 2 of lo

// One arrival per warp: every lane has finished (and fenced) its part, lane 0 signals.  512 separate arrivals on one
// mbarrier serialise on the shared-memory atomic unit (measured: ~2000 cycles per phase).
__device__ __forceinline__ void warp_arrive(uint64_t* bar, int lane) {
  __syncwarp();
  if (lane == 0) umma::mbar_arrive(bar);
}

__device__ __forceinline__ uint32_t pk_half2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
// 8 floats -> one 16-byte chunk of fp16 "hi" and one of fp16 "lo" (v = hi + lo to ~2^-22 relative)
__device__ __forceinline__ void split8(const float (&v)[8], uint4& hi, uint4& lo) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    __half2 hh = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
    const float2 hf = __half22float2(hh);
    h[i] = *reinterpret_cast<uint32_t*>(&hh);
    l[i] = pk_half2(v[2 * i] - hf.x, v[2 * i + 1] - hf.y);
  }
  hi = make_uint4(h[0], h[1], h[2], h[3]);
  lo = make_uint4(l[0], l[1], l[2], l[3]);
}

// The draws in "split" form: a row of K words, word 2m = half2(hi_2m, hi_2m+1), word 2m+1 = half2(lo_2m, lo_2m+1)
// with z = hi + lo to 2^-22: eight consecutive draws are 32 bytes {H01, L01, H23, L23, H45, L45, H67, L67}.
__device__ __forceinline__ void split_pair(float z0, float z1, uint32_t& H, uint32_t& L) {
  const __half2 h = __floats2half2_rn(z0, z1);
  const float2 hf = __half22float2(h);
  H = *reinterpret_cast<const uint32_t*>(&h);
  L = pk_half2(z0 - hf.x, z1 - hf.y);
}
__device__ __forceinline__ __half2 as_h2(uint32_t w) { return *reinterpret_cast<const __half2*>(&w); }
__device__ __forceinline__ uint32_t h2_bits(__half2 h) { return *reinterpret_cast<const uint32_t*>(&h); }

// One row-octet of V = g * eps straight into fp16 hi / lo operand pieces, in packed half arithmetic: with g = gh + gl
// and eps = eh + el,  hi = fl16(gh eh),  lo = fl16(gl eh + fl16(gh el + (gh eh - hi)))  -- the residual of the first
// product is exact (one fused rounding), the dropped gl el term and the roundings of lo are 2^-22 of V.  16 HFMA2-class
// instructions per 8 elements instead of 8 FMUL + 4 + 16 + 4 conversions.
__device__ __forceinline__ void split_g(float g, __half2& gh, __half2& gl) {
  gh = __float2half2_rn(g);
  gl = __float2half2_rn(g - __low2float(gh));
}
__device__ __forceinline__ void gen_octet(const __half2 gh, const __half2 gl, const uint4& w0, const uint4& w1, uint4& hi, uint4& lo) {
  auto one = [&](uint32_t Hw, uint32_t Lw, uint32_t& oh, uint32_t& ol) {
    const __half2 eh = as_h2(Hw), el = as_h2(Lw);
    const __half2 h = __hmul2(gh, eh);
    __half2 r = __hfma2(gh, eh, __hneg2(h));
    r = __hfma2(gh, el, r);
    r = __hfma2(gl, eh, r);
    oh = h2_bits(h);
    ol = h2_bits(r);
  };
  one(w0.x, w0.y, hi.x, lo.x);
  one(w0.z, w0.w, hi.y, lo.y);
  one(w1.x, w1.y, hi.z, lo.z);
  one(w1.z, w1.w, hi.w, lo.w);
}
// Diagonal block, row r of the unit: element r becomes (g eps_r) L_rr + dadd (exp chain rule + entropy term, in
// fp32), the elements above the diagonal are padding -> 0.
__device__ __forceinline__ void fix_diag(float g, float lii, float dadd, const uint32_t* ep, int r, uint4& hi, uint4& lo) {
  const uint32_t Hw = ep[(r >> 1) * 2], Lw = ep[(r >> 1) * 2 + 1];
  const float er = (r & 1) ? __high2float(as_h2(Hw)) + __high2float(as_h2(Lw)) : __low2float(as_h2(Hw)) + __low2float(as_h2(Lw));
  const float vd = fmaf(g * er, lii, dadd);
  const __half hd = __float2half_rn(vd);
  const __half ld = __float2half_rn(vd - __half2float(hd));
  const uint32_t hb = (uint32_t)__half_as_ushort(hd), lb = (uint32_t)__half_as_ushort(ld);
  uint32_t* hw = reinterpret_cast<uint32_t*>(&hi);
  uint32_t* lw = reinterpret_cast<uint32_t*>(&lo);
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    // keep positions < r, replace position r, clear positions > r  (position 2q = low half, 2q + 1 = high half)
    const uint32_t keep = (2 * q + 1 < r) ? 0xffffffffu : ((2 * q + 1 == r || 2 * q == r) ? 0x0000ffffu : 0u);
    uint32_t h = hw[q] & keep, l = lw[q] & keep;
    if (2 * q + 1 == r) { h |= hb << 16; l |= lb << 16; }
    if (2 * q == r) { h = hb; l = lb; }
    hw[q] = h;
    lw[q] = l;
  }
}

// element c (0..7) of an octet := val (fp16 hi / lo pair)
__device__ __forceinline__ void put_elem(uint4& hi, uint4& lo, int c, float val) {
  const __half hd = __float2half_rn(val);
  const __half ld = __float2half_rn(val - __half2float(hd));
  const uint32_t sh = (c & 1) ? 16u : 0u, keep = ~(0xffffu << sh);
  uint32_t* hw = reinterpret_cast<uint32_t*>(&hi);
  uint32_t* lw = reinterpret_cast<uint32_t*>(&lo);
  hw[c >> 1] = (hw[c >> 1] & keep) | ((uint32_t)__half_as_ushort(hd) << sh);
  lw[c >> 1] = (lw[c >> 1] & keep) | ((uint32_t)__half_as_ushort(ld) << sh);
}

// ------------------------------------------------------------------------------------------------ operand image
// Padded column pc of the factor -> (row i, column j) (j > i or i >= K: padding, weight 0).
__device__ __forceinline__ void mv_row_of(int pc, int& i, int& j) {
  const int unit = pc >> 6;
  int I = 0;
  while ((I + 1) * (I + 2) / 2 <= unit) ++I;
  const int Jc = unit - I * (I + 1) / 2;
  i = 8 * I + ((pc >> 3) & 7);
  j = 8 * Jc + (pc & 7);
}

// img  (forward):  chunks of 128 padded columns x KH = [hidden | 1 | 0], hi tile then lo tile, rows = columns of L
// imgb (backward): units of 64 padded columns x NH hidden units (no bias column), hi tile then lo tile
__global__ void mvn_prep_kernel(const float* __restrict__ w22, const float* __restrict__ b22, const float* __restrict__ w21,
                                const float* __restrict__ b21, int K, int H, int KH,
                                int NH, int n_chunks, unsigned char* __restrict__ img, unsigned char* __restrict__ imgb) {
  const int n_oct = KH / 8;
  const int total = n_chunks * kMvCN * n_oct;
  const uint32_t RS = (uint32_t)n_oct * kCSb, RSb = (uint32_t)(NH / 8) * kCSb;
  const size_t tile_bytes = (size_t)kMvCN * KH * 2, tileb_bytes = (size_t)64 * NH * 2;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
    const int pc = e / n_oct, o = e - pc * n_oct;
    int i, j;
    mv_row_of(pc, i, j);
    float v[8];
#pragma unroll
    for (int t = 0; t < 8; ++t) v[t] = 0.f;
    if (i < K && j <= i) {
      const size_t p = (size_t)tri_off(i) + j;
#pragma unroll
      for (int t = 0; t < 8; ++t) {
        const int k = o * 8 + t;
        if (k < H) v[t] = kMvWScale * w22[p * H + k];
        else if (k == H) v[t] = kMvWScale * b22[p];
      }
    } else if (w21 && (i >> 3) == (j >> 3)) {         // padding of a diagonal unit: a row of fc21 in its slot
      const int r = mv_loc_row_of_slot(pc & 63), il = 8 * (i >> 3) + r;
      if (r >= 0 && il < K) {
#pragma unroll
        for (int t = 0; t < 8; ++t) {
          const int k = o * 8 + t;
          if (k < H) v[t] = kMvWScale * w21[(size_t)il * H + k];
          else if (k == H) v[t] = kMvWScale * b21[il];
        }
      }
    }
    uint4 hi, lo;
    split8(v, hi, lo);
    const int c = pc / kMvCN, r = pc - c * kMvCN;
    unsigned char* base = img + (size_t)c * 2 * tile_bytes + (size_t)(r >> 3) * RS + (size_t)o * kCSb + (r & 7) * 16;
    *reinterpret_cast<uint4*>(base) = hi;
    *reinterpret_cast<uint4*>(base + tile_bytes) = lo;
    if (imgb && o * 8 < NH) {
      if (o * 8 + 8 > H) {                      // the octet that holds the bias column forward: weights only here
#pragma unroll
        for (int t = 0; t < 8; ++t)
          if (o * 8 + t >= H) v[t] = 0.f;
        split8(v, hi, lo);
      }
      const int u = pc >> 6, ru = pc & 63;
      unsigned char* bb = imgb + (size_t)u * 2 * tileb_bytes + (size_t)(ru >> 3) * RSb + (size_t)o * kCSb + (ru & 7) * 16;
      *reinterpret_cast<uint4*>(bb) = hi;
      *reinterpret_cast<uint4*>(bb + tileb_bytes) = lo;
    }
  }
}

// ------------------------------------------------------------------------------------------------ forward
constexpr uint32_t kFwXchBytes = (2 * 8 + 1) * kMvTP * 4;     // partial row sums [parity][r][person] + log det [person]
struct FwSmem {
  uint32_t offW, offEps, offX, total;
  uint32_t chunk_bytes;
  int nst, ES;
};
__host__ __device__ inline FwSmem fw_smem(int K, int H) {
  FwSmem s;
  s.chunk_bytes = (uint32_t)mv_chunk_bytes(H);
  s.ES = (K + 7) / 8 * 8 + 4;
  const uint32_t fixed = (uint32_t)kMvTP * s.ES * 4 + kFwXchBytes;
  const uint32_t cap = 227 * 1024 - 1024;
  int nst = (int)((cap - fixed) / s.chunk_bytes);
  if (nst > 4) nst = 4;
  s.nst = nst;
  uint32_t o = 0;
  s.offW = o; o += (uint32_t)(nst > 0 ? nst : 0) * s.chunk_bytes;
  s.offEps = o; o += (uint32_t)kMvTP * s.ES * 4;
  s.offX = o; o += kFwXchBytes;
  s.total = o;
  return s;
}

// N(0,1) draws of particle s for one person -> er[0..K) scaled by `mul` (same Philox coordinates as noise_at)
__device__ __forceinline__ float fill_noise_row(const MirtParams& m, int s, int64_t b, float* er, float mul) {
  const int K = m.K;
  float ss = 0.f;
  if (m.eps) {
    const float* src = m.eps + ((int64_t)s * m.B + b) * K;
    for (int k = 0; k < K; ++k) {
      const float z = src[k];
      er[k] = z * mul;
      ss = fmaf(z, z, ss);
    }
  } else if (K > 32 && K <= 128) {
    for (int l = 0; l < 32; ++l) {
      const float4 z = normal4(m.seed, m.offset, (uint64_t)(m.person_offset + b), (uint32_t)s * 32u + (uint32_t)l);
      er[l] = z.x * mul; ss = fmaf(z.x, z.x, ss);
      if (l + 32 < K) { er[l + 32] = z.y * mul; ss = fmaf(z.y, z.y, ss); }
      if (l + 64 < K) { er[l + 64] = z.z * mul; ss = fmaf(z.z, z.z, ss); }
      if (l + 96 < K) { er[l + 96] = z.w * mul; ss = fmaf(z.w, z.w, ss); }
    }
  } else {
    const int n0 = s * K, n1 = n0 + K;
    for (int blk = n0 >> 2; blk <= (n1 - 1) >> 2; ++blk) {
      const float4 z = normal4(m.seed, m.offset, (uint64_t)(m.person_offset + b), (uint32_t)blk);
      const float zz[4] = {z.x, z.y, z.z, z.w};
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int n = 4 * blk + c;
        if (n >= n0 && n < n1) { er[n - n0] = zz[c] * mul; ss = fmaf(zz[c], zz[c], ss); }
      }
    }
  }
  return ss;
}

// The draws of the step, once, ahead of the forward kernel (the library's Philox stream: same coordinates as every
// other kernel): [S*B][K] fp32 (eps_out) or fp16 (hi, lo) pair words (eps_split, K % 4 == 0).  A thread fills the
// row of its person in shared memory, a warp writes its 32 rows with whole-line stores.  Inside the forward kernel
// this work (~1600 instructions per person) sat in front of every tile's first MMA.
constexpr int kNzThreads = 128;
__global__ void __launch_bounds__(kNzThreads) mvn_noise_kernel(const MvnEncParams P) {
  extern __shared__ __align__(16) unsigned char smem[];
  const MirtParams& m = P.m;
  const int K = m.K, ES = K + 1;                       // odd pitch: the transposed reads below are conflict-free
  float* rows = reinterpret_cast<float*>(smem);
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int64_t n_tiles = (m.B + kNzThreads - 1) / kNzThreads;
  for (int64_t step = blockIdx.x; step < n_tiles * m.S; step += gridDim.x) {
    const int s = (int)(step / n_tiles);
    const int64_t tile = step - (int64_t)s * n_tiles;
    const int64_t b = tile * kNzThreads + t;
    fill_noise_row(m, s, b < m.B ? b : 0, rows + (size_t)t * ES, 1.f);
    __syncwarp();
    const int64_t b0 = tile * kNzThreads + warp * 32;
    const float* e0 = rows + (size_t)(warp * 32) * ES;
    if (P.eps_split) {
      for (int pp = 0; pp < 32 && b0 + pp < m.B; ++pp) {
        uint2* dst = reinterpret_cast<uint2*>(P.eps_split + ((int64_t)s * m.B + b0 + pp) * K);
        for (int mm = lane; mm < K / 2; mm += 32) {
          uint2 w;
          split_pair(e0[(size_t)pp * ES + 2 * mm], e0[(size_t)pp * ES + 2 * mm + 1], w.x, w.y);
          dst[mm] = w;
        }
      }
    } else {
      for (int pp = 0; pp < 32 && b0 + pp < m.B; ++pp) {
        float* dst = P.eps_out + ((int64_t)s * m.B + b0 + pp) * K;
        for (int k = lane; k < K; k += 32) dst[k] = e0[(size_t)pp * ES + k];
      }
    }
    __syncwarp();
  }
}

// 32 < K <= 128 (the library's stream gives lane l of a person the draws l, l + 32, l + 64, l + 96 from one Philox
// call): a warp per person, nothing staged -- neighbouring lanes exchange their values to form the fp16 pair words.
__global__ void __launch_bounds__(256) mvn_noise_mid_kernel(const MvnEncParams P) {
  const MirtParams& m = P.m;
  const int K = m.K;
  const int lane = threadIdx.x & 31;
  const int64_t n_rows = (int64_t)m.S * m.B;
  for (int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); row < n_rows; row += (int64_t)gridDim.x * 8) {
    const int s = (int)(row / m.B);
    const int64_t b = row - (int64_t)s * m.B;
    const float4 z = normal4(m.seed, m.offset, (uint64_t)(m.person_offset + b), (uint32_t)s * 32u + (uint32_t)lane);
    const float zz[4] = {z.x, z.y, z.z, z.w};
    if (P.eps_split) {
      uint2* dst = reinterpret_cast<uint2*>(P.eps_split + row * K);
      const bool odd = lane & 1;
#pragma unroll
      for (int seg = 0; seg < 4; ++seg) {
        const float other = __shfl_xor_sync(0xffffffffu, zz[seg], 1);
        const int k = 32 * seg + (lane & ~1);                      // the even draw of the pair
        if ((seg >> 1) == (int)odd && k < K) {                    // even lanes write segments 0, 1; odd lanes 2, 3
          uint2 w;
          split_pair(odd ? other : zz[seg], odd ? zz[seg] : other, w.x, w.y);
          dst[k >> 1] = w;
        }
      }
    } else {
      float* dst = P.eps_out + row * K;
#pragma unroll
      for (int seg = 0; seg < 4; ++seg)
        if (32 * seg + lane < K) dst[32 * seg + lane] = zz[seg];
    }
  }
}

__global__ void __launch_bounds__(kFwThreads, 1) mvn_fwd_kernel(const MvnEncParams P) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t bar_wfull[4], bar_wempty[4], bar_tfull[kFwTmemBufs], bar_tempty[kFwTmemBufs], bar_hfull;
  __shared__ uint32_t tmem_base_sm;
  const MirtParams& m = P.m;
  const int K = m.K, S = m.S, H = P.H, KH = mv_kh(H);
  const FwSmem L = fw_smem(K, H);
  const int NST = L.nst, ES = L.ES;
  const int n_chunks = mv_chunks(K), NB = mv_nb(K);
  const uint32_t RS = (uint32_t)(KH / 8) * kCSb;
  unsigned char* Wst = smem + L.offW;
  float* eps_s = reinterpret_cast<float*>(smem + L.offEps);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t n_tiles = (m.B + kMvTP - 1) / kMvTP;

  if (tid == 0) {
    for (int i = 0; i < 4; ++i) { umma::mbar_init(&bar_wfull[i], 1); umma::mbar_init(&bar_wempty[i], 1); }
    for (int i = 0; i < kFwTmemBufs; ++i) { umma::mbar_init(&bar_tfull[i], 1); umma::mbar_init(&bar_tempty[i], kFwEpiWarps); }
    umma::mbar_init(&bar_hfull, kFwEpiWarps);
    umma::fence_mbar_init();
  }
  if (warp == kFwEpiWarps + 1) umma::tmem_alloc(&tmem_base_sm, 512);
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem = tmem_base_sm;

  // Producer and MMA warp: whole-warp loops, only the issuing instructions under elect_one() (see umma::elect_one).
  // The [hidden | 1] tile is the A operand in TENSOR memory (written by the epilogue threads, lane = person): an
  // SS-mode 128 x 128 x 16 MMA needs 8 KB of shared-memory reads per 64 cycles -- the whole port.
  if (warp == kFwEpiWarps) {
    // =========================== bulk-copy producer: the operand image, chunk by chunk ========================
    uint32_t it = 0;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x)
      for (int s = 0; s < S; ++s)
        for (int c = 0; c < n_chunks; ++c, ++it) {
          const uint32_t st = it % NST, round = it / NST;
          umma::mbar_wait(&bar_wempty[st], (round & 1) ^ 1);
          if (umma::elect_one())
            umma::bulk_load(Wst + (size_t)st * L.chunk_bytes, P.wimg + (size_t)c * L.chunk_bytes, L.chunk_bytes,
                            &bar_wfull[st]);
          __syncwarp();
        }
  } else if (warp > kFwEpiWarps) {
    // =========================== MMA issuers: two warps on alternate chunks (see mvn_bwd_dw_kernel) ===========
    const uint32_t mw = (uint32_t)(warp - (kFwEpiWarps + 1));
    const uint32_t idesc = umma::make_idesc_f16(kMvTP, kMvCN, umma::kMajorK, umma::kMajorK);
    const uint32_t aH = tmem + kFwACol, aL = aH + (uint32_t)(KH / 2);
    uint32_t it = 0, ntile = 0;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++ntile) {
      umma::mbar_wait(&bar_hfull, ntile & 1);
      umma::fence_after_sync();
      for (int s = 0; s < S; ++s)
        for (int c = 0; c < n_chunks; ++c, ++it) {
          if ((it % kFwMmaWarps) != mw) continue;
          const uint32_t st = it % NST, round = it / NST;
          const uint32_t buf = it % kFwTmemBufs, tround = it / kFwTmemBufs;
          umma::mbar_wait(&bar_wfull[st], round & 1);
          umma::mbar_wait(&bar_tempty[buf], (tround & 1) ^ 1);
          umma::fence_after_sync();
          if (umma::elect_one()) {
            const uint32_t sWh = umma::smem_u32(Wst + (size_t)st * L.chunk_bytes), sWl = sWh + L.chunk_bytes / 2;
            uint32_t acc = 0;
            for (int ks = 0; ks < KH / 16; ++ks) {
#pragma unroll
              for (int term = 0; term < 3; ++term) {            // hi*hi, hi*lo, lo*hi
                if (term == 2 && 16 * ks >= H) continue;        // only the ones column and padding: the lo part is 0
                const uint64_t db = umma::make_desc((term == 1 ? sWl : sWh) + ks * 2 * kCSb, kCSb, RS);
                umma::mma_f16_ts(tmem + buf * kMvCN, (term == 2 ? aL : aH) + 8 * ks, db, idesc, acc);
                acc = 1;
              }
            }
            umma::commit(&bar_wempty[st]);
            umma::commit(&bar_tfull[buf]);
          }
          __syncwarp();
        }
    }
  } else {
    // =========================== epilogue: thread = person = TMEM lane =========================================
    // Two warps per lane quadrant (a lone warp per scheduler issues at a quarter of the rate): warp `half` takes unit
    // `half` of every chunk.  A block row of L ends with its diagonal unit; the warp that holds that unit finishes
    // the row, the other one hands over its partial sums through shared memory right after the unit before it.
    const int quad = warp & 3, half = warp >> 2;
    const int t = quad * 32 + lane;                      // person of the tile = TMEM lane
    float* er = eps_s + (size_t)t * ES;
    float* xch = reinterpret_cast<float*>(smem + L.offX);
    const uint32_t tlane = tmem + ((uint32_t)(quad * 32) << 16);
    const float inv_scale = 1.0f / kMvWScale;
    const bool has_loc = P.w21 != nullptr;               // fc21 fused: the location comes out of the same product
    auto pair_sync = [&]() { asm volatile("bar.sync %0, 64;" ::"r"(1 + quad) : "memory"); };
    uint32_t it = 0;
    if (half == 0)
      for (int k = K; k < ES; ++k) er[k] = 0.f;          // padding columns of a row's last group: finite, weight 0
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const int64_t b = tile * kMvTP + t;
      const bool live = b < m.B;
      // ---- [hidden | 1 | 0] row of this person -> fp16 hi / lo words of the A operand in tensor memory ------------
      // (every MMA of the previous tile has completed: both warps have consumed their last accumulator)
      {
        const float* hr = P.hidden + (live ? b : 0) * H;
        const bool vec = (H & 3) == 0;
        float hv[3][16];                                 // this warp's k-steps (at most 3 of 5), every load in flight first
#pragma unroll
        for (int q2 = 0; q2 < 3; ++q2) {
          const int o2 = half + 2 * q2;
#pragma unroll
          for (int q = 0; q < 16; q += 4) {
            const int k = 16 * o2 + q;
            if (o2 < KH / 16 && vec && live && k + 4 <= H) {
              const float4 a0 = *reinterpret_cast<const float4*>(hr + k);
              hv[q2][q] = a0.x; hv[q2][q + 1] = a0.y; hv[q2][q + 2] = a0.z; hv[q2][q + 3] = a0.w;
            } else {
#pragma unroll
              for (int e = 0; e < 4; ++e)
                hv[q2][q + e] = (o2 >= KH / 16 || !live) ? 0.f : (k + e < H ? hr[k + e] : (k + e == H ? 1.f : 0.f));
            }
          }
        }
#pragma unroll
        for (int q2 = 0; q2 < 3; ++q2) {
          const int o2 = half + 2 * q2;                  // a k-step: 16 elements = 8 words
          if (o2 < KH / 16) {
            uint32_t h8[8], l8[8];
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
              float v[8];
#pragma unroll
              for (int q = 0; q < 8; ++q) v[q] = hv[q2][8 * hh + q];
              uint4 hi, lo;
              split8(v, hi, lo);
              h8[4 * hh] = hi.x; h8[4 * hh + 1] = hi.y; h8[4 * hh + 2] = hi.z; h8[4 * hh + 3] = hi.w;
              l8[4 * hh] = lo.x; l8[4 * hh + 1] = lo.y; l8[4 * hh + 2] = lo.z; l8[4 * hh + 3] = lo.w;
            }
            umma::tmem_st8(tlane + kFwACol + 8 * o2, h8);
            umma::tmem_st8(tlane + kFwACol + (uint32_t)(KH / 2) + 8 * o2, l8);
          }
        }
        umma::tmem_st_wait();
        umma::fence_before_sync();
        warp_arrive(&bar_hfull, lane);
      }
      float logdet = 0.f, auxp = 0.f;
      for (int s = 0; s < S; ++s) {
        pair_sync();                                     // the partner has finished with the previous draws
        // the person's draws (mvn_noise_kernel, or injected) -> its row in shared memory, prescaled; half a row per warp
        float ss = 0.f;
        {
          const int64_t row = ((int64_t)s * m.B + (live ? b : 0)) * K;
          const int kh = (K / 2 + 3) & ~3;               // split point: a multiple of 4
          const int k0 = half ? kh : 0, k1 = half ? K : (kh < K ? kh : K);
          if (P.eps_split && !m.eps) {                   // K % 4 == 0: whole 16-byte pieces = 4 draws
            for (int kk = k0; kk < k1; kk += 32) {       // eight pieces in flight before the first one is used
              uint4 w[8];
#pragma unroll
              for (int q = 0; q < 8; ++q)
                w[q] = kk + 4 * q < k1 ? *reinterpret_cast<const uint4*>(P.eps_split + row + kk + 4 * q) : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
              for (int q = 0; q < 8; ++q) {
                if (kk + 4 * q < k1) {
                  const float2 h0 = __half22float2(as_h2(w[q].x)), l0 = __half22float2(as_h2(w[q].y));
                  const float2 h1 = __half22float2(as_h2(w[q].z)), l1 = __half22float2(as_h2(w[q].w));
                  const float z0 = h0.x + l0.x, z1 = h0.y + l0.y, z2 = h1.x + l1.x, z3 = h1.y + l1.y;
                  ss = fmaf(z0, z0, fmaf(z1, z1, fmaf(z2, z2, fmaf(z3, z3, ss))));
                  *reinterpret_cast<float4*>(er + kk + 4 * q) = make_float4(z0 * inv_scale, z1 * inv_scale, z2 * inv_scale, z3 * inv_scale);
                }
              }
            }
          } else {
            const float* src = (m.eps ? m.eps : P.eps_out) + row;
            for (int k = k0; k < k1; ++k) {
              const float z = src[k];
              ss = fmaf(z, z, ss);
              er[k] = z * inv_scale;
            }
          }
        }
        pair_sync();                                     // both halves of the row are in place
        auxp = fmaf(0.5f, ss, auxp);
        const float* mur = m.mu + (live ? b : 0) * K;
        float* xr = m.x ? m.x + ((int64_t)s * m.B + (live ? b : 0)) * K : nullptr;
        int I = half, Jc = 0;                            // unit `half` of chunk 0: (0, 0) or (1, 0)
        int ug = half;                                   // its index in the order of the units
        float acc[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) acc[r] = 0.f;
        for (int c = 0; c < n_chunks; ++c, ++it) {
          const uint32_t buf = it % kFwTmemBufs, tround = it / kFwTmemBufs;
          umma::mbar_wait(&bar_tfull[buf], tround & 1);
          umma::fence_after_sync();
          if (I < NB) {                                          // warp-uniform
            const bool diag = Jc == I;
            float mu_r[8];
            if (diag && !has_loc) {
#pragma unroll
              for (int r = 0; r < 8; ++r) mu_r[r] = 8 * I + r < K ? mur[8 * I + r] : 0.f;
            }
            float v[64];
            umma::tmem_ld64(tlane + buf * kMvCN + half * 64, v);
            const float4 e0v = *reinterpret_cast<const float4*>(er + 8 * Jc);
            const float4 e1v = *reinterpret_cast<const float4*>(er + 8 * Jc + 4);
            if (diag && has_loc) {
              // the padding above the diagonal carries fc21's rows (mv_loc_slot): only columns j <= i meet the draws,
              // the location of row r is read from its slot
              const float ev[8] = {e0v.x, e0v.y, e0v.z, e0v.w, e1v.x, e1v.y, e1v.z, e1v.w};
#pragma unroll
              for (int r = 0; r < 8; ++r) {
                float d = 0.f;
#pragma unroll
                for (int jj = 0; jj <= r; ++jj) d = fmaf(v[8 * r + jj], ev[jj], d);
                acc[r] += d;
                mu_r[r] = v[mv_loc_slot(r)] * inv_scale;
              }
            } else {
#pragma unroll
              for (int r = 0; r < 8; ++r) {
                float d0 = v[8 * r] * e0v.x, d1 = v[8 * r + 4] * e1v.x;
                d0 = fmaf(v[8 * r + 1], e0v.y, d0); d1 = fmaf(v[8 * r + 5], e1v.y, d1);
                d0 = fmaf(v[8 * r + 2], e0v.z, d0); d1 = fmaf(v[8 * r + 6], e1v.z, d1);
                d0 = fmaf(v[8 * r + 3], e0v.w, d0); d1 = fmaf(v[8 * r + 7], e1v.w, d1);
                acc[r] += d0 + d1;
              }
            }
            float* xs = xch + (size_t)((I & 1) * 8) * kMvTP + t;
            if (diag) {                                         // rows 8I .. 8I+7 are complete once the partner's share is in
              if (I > 0) {
                pair_sync();
#pragma unroll
                for (int r = 0; r < 8; ++r) acc[r] += xs[r * kMvTP];
              }
              const float ev[8] = {e0v.x, e0v.y, e0v.z, e0v.w, e1v.x, e1v.y, e1v.z, e1v.w};
              float xo[8], lo8[8];
#pragma unroll
              for (int r = 0; r < 8; ++r) {
                const int i = 8 * I + r;
                const float dv = v[9 * r];                     // 64 raw_ii
                const float dr = dv * inv_scale;
                const float lii = __expf(dr);
                const float xv = mu_r[r] + fmaf(lii * kMvWScale - dv, ev[r], acc[r]);    // exp(raw_ii) eps_i for raw_ii eps_i
                xo[r] = xv; lo8[r] = lii;
                if (i < K) {
                  auxp = fmaf(-0.5f * xv, xv, auxp);
                  if (s == 0) logdet += dr;
                }
                acc[r] = 0.f;
              }
              if (live) {                                       // 32-byte pieces of the person's rows: whole sectors
                const int i0 = 8 * I;
                const bool wd = s == 0 && m.diagL != nullptr;
                if ((K & 3) == 0) {
                  if (xr) {
                    *reinterpret_cast<float4*>(xr + i0) = make_float4(xo[0], xo[1], xo[2], xo[3]);
                    if (i0 + 4 < K) *reinterpret_cast<float4*>(xr + i0 + 4) = make_float4(xo[4], xo[5], xo[6], xo[7]);
                  }
                  if (wd) {
                    float* dg = m.diagL + b * K + i0;
                    *reinterpret_cast<float4*>(dg) = make_float4(lo8[0], lo8[1], lo8[2], lo8[3]);
                    if (i0 + 4 < K) *reinterpret_cast<float4*>(dg + 4) = make_float4(lo8[4], lo8[5], lo8[6], lo8[7]);
                  }
                } else {
#pragma unroll
                  for (int r = 0; r < 8; ++r)
                    if (i0 + r < K) {
                      if (xr) xr[i0 + r] = xo[r];
                      if (wd) m.diagL[b * K + i0 + r] = lo8[r];
                    }
                }
              }
            } else if (Jc == I - 1) {                           // the unit before the diagonal one: hand the row over
#pragma unroll
              for (int r = 0; r < 8; ++r) { xs[r * kMvTP] = acc[r]; acc[r] = 0.f; }
              pair_sync();
            }
          }
          umma::fence_before_sync();
          warp_arrive(&bar_tempty[buf], lane);
#pragma unroll
          for (int st = 0; st < 2; ++st) {                      // two units on: this warp's unit of the next chunk
            if (Jc == I) { ++I; Jc = 0; } else ++Jc;
          }
          ug += 2;
        }
      }
      // loss pieces of the person: the tail sums the four words
      if (live && m.aux) *reinterpret_cast<float2*>(m.aux + 4 * b + 2 * half) = make_float2(auxp + (float)S * logdet, 0.f);
      if (P.logdet_out) {
        float* ls = xch + (size_t)16 * kMvTP + t;
        if (half == 1) *ls = logdet;
        pair_sync();
        if (half == 0 && live) P.logdet_out[b] = logdet + *ls;
        pair_sync();
      }
    }
  }
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  if (warp == kFwEpiWarps + 1) umma::tmem_dealloc(tmem, 512);
}


// ------------------------------------------------------------------------------------------------ tail of the likelihood
// G = -(scale / S) (sum over item tiles of d lb / d x  -  x)  (the prior's share is -x), written over plane 0 of gx;
// d_mu = sum_s G (the pathwise gradient of the location is the gradient of x); loss partial per CTA from the
// log-likelihood rows and the forward kernel's aux word.  What mirt_cholgrad_* does, minus the factor gradient.
constexpr int kTailThreads = 256;
__global__ void __launch_bounds__(kTailThreads) mvn_tail_kernel(const MirtParams p, float* __restrict__ gmax) {
  __shared__ double red[kTailThreads / 32];
  const int K = p.K, S = p.S;
  const int64_t BK = p.B * K, R = (int64_t)S * p.B;
  const float kf = -p.scale / (float)S;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (int64_t)gridDim.x * blockDim.x;
  if (p.d_mu) {
    float gm = 0.f;
    if ((BK & 3) == 0 && (reinterpret_cast<uintptr_t>(p.d_mu) & 15) == 0) {      // 16-byte pieces, every load of a piece in flight before the first use
      const int np = p.n_parts;
      for (int64_t e4 = tid; e4 < (BK >> 2); e4 += nth) {
        float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int s = 0; s < S; ++s) {
          const int64_t at = (int64_t)s * BK + 4 * e4;
          const float4 xv = __ldcs(reinterpret_cast<const float4*>(p.x + at));
          float4 v[4];
#pragma unroll
          for (int part = 0; part < 4; ++part)
            v[part] = part < np ? __ldcs(reinterpret_cast<const float4*>(p.gx + (int64_t)part * R * K + at))
                                : make_float4(0.f, 0.f, 0.f, 0.f);
          float4 g = make_float4(v[0].x - xv.x, v[0].y - xv.y, v[0].z - xv.z, v[0].w - xv.w);
#pragma unroll
          for (int part = 1; part < 4; ++part) { g.x += v[part].x; g.y += v[part].y; g.z += v[part].z; g.w += v[part].w; }
          for (int part = 4; part < np; ++part) {
            const float4 w = __ldcs(reinterpret_cast<const float4*>(p.gx + (int64_t)part * R * K + at));
            g.x += w.x; g.y += w.y; g.z += w.z; g.w += w.w;
          }
          g.x *= kf; g.y *= kf; g.z *= kf; g.w *= kf;
          *reinterpret_cast<float4*>(p.gx + at) = g;
          gm = fmaxf(fmaxf(gm, fmaxf(fabsf(g.x), fabsf(g.y))), fmaxf(fabsf(g.z), fabsf(g.w)));
          t.x += g.x; t.y += g.y; t.z += g.z; t.w += g.w;
        }
        __stcs(reinterpret_cast<float4*>(p.d_mu + 4 * e4), t);
      }
    } else {
      for (int64_t e = tid; e < BK; e += nth) {
        float t = 0.f;
        for (int s = 0; s < S; ++s) {
          const int64_t at = (int64_t)s * BK + e;
          float g = -p.x[at];
          for (int part = 0; part < p.n_parts; ++part) g += p.gx[(int64_t)part * R * K + at];
          g *= kf;
          p.gx[at] = g;
          gm = fmaxf(gm, fabsf(g));
          t += g;
        }
        p.d_mu[e] = t;
      }
    }
    gm = warp_max(gm);                                  // a maximum: independent of the order, hence deterministic
    if ((threadIdx.x & 31) == 0 && gm > 0.f) atomicMax(reinterpret_cast<int*>(gmax), __float_as_int(gm));
  }
  double acc = 0.0;
  for (int64_t b = tid; b < p.B; b += nth) {
    float ell = 0.f;
    for (int s = 0; s < S; ++s)
      for (int part = 0; part < p.n_parts; ++part) ell += p.ell[(int64_t)part * R + (int64_t)s * p.B + b];
    const float4 ax = *reinterpret_cast<const float4*>(p.aux + 4 * b);
    acc += (double)(kf * (ell + ((ax.x + ax.y) + (ax.z + ax.w))));
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < kTailThreads / 32; ++w) t += red[w];
    p.loss_partial[blockIdx.x] = t;
  }
}

// ------------------------------------------------------------------------------------------------ backward
constexpr int kBwGenWarps = 8;                         // warps 0-7 generate V: thread = (person, 2 or 4 rows of a unit)
constexpr int kBwGen = kBwGenWarps * 32;
constexpr int kBwThreads = kBwGen + 64;                // + a bulk-copy producer warp and an MMA issuer warp

__device__ __forceinline__ void gen_bar() { asm volatile("bar.sync 1, %0;" ::"n"(kBwGen) : "memory"); }

__device__ __forceinline__ void unit_to_block(int u, int& I, int& Jc) {
  I = 0;
  while ((I + 1) * (I + 2) / 2 <= u) ++I;
  Jc = u - I * (I + 1) / 2;
}

// G rows 8I + NR sub + {0 .. NR-1} of the thread's person (0 beyond K / B), the diagonal of L for them and the entropy term
template <int NR>
__device__ __forceinline__ void load_g(const MirtParams& m, int s, int64_t b, bool live, int I, int sub, float ginv,
                                       float (&gI)[NR], float (&lI)[NR], float (&dadd)[NR]) {
#pragma unroll
  for (int rr = 0; rr < NR; ++rr) {
    const int i = 8 * I + NR * sub + rr;
    const bool ok = live && i < m.K;
    gI[rr] = ok ? ginv * m.gx[((int64_t)s * m.B + b) * m.K + i] : 0.f;
    lI[rr] = ok ? m.diagL[b * m.K + i] : 0.f;
    dadd[rr] = (ok && s == 0) ? -m.scale * ginv : 0.f;
  }
}

// ---- d hidden[person][h] = (1 / 64) sum_p V[person][p] (64 W22)[p][h]: CTAs own person tiles, every unit of V ------
constexpr int kDhMaxStages = 8;
constexpr int kDhMmaWarps = 2;                        // MMA-issuing warps (alternate units)
constexpr int kDhThreads = kBwGen + 32 + 32 * kDhMmaWarps;
constexpr int kDhAStages = 4;                          // V units in tensor memory: 4 x (32 hi + 32 lo) columns
struct DhSmem {
  uint32_t offEps, offW, total, w_bytes, eps_bytes;
  int ES, nst, neb;
};
__host__ __device__ inline DhSmem dh_smem(int K, int H, bool bulk) {
  DhSmem s;
  s.ES = bulk ? K : (K + 7) / 8 * 8 + 4;               // words per row of the split draws
  s.eps_bytes = ((uint32_t)kMvTP * s.ES * 4 + 64 + 127) & ~127u;
  s.w_bytes = 2u * 64 * (uint32_t)mv_nh(H) * 2;         // hi + lo, 64 columns x NH
  // A bulk copy of a unit takes ~1.2 us from L2: the stream of units needs depth, the draws (one copy per tile)
  // do not -- so the draws get one buffer and the units every stage that fits (measured: 3 -> 6 stages).
  s.neb = 1;
  int nst = (int)((226u * 1024u - s.eps_bytes) / s.w_bytes);
  s.nst = nst > kDhMaxStages ? kDhMaxStages : nst;
  uint32_t o = 0;
  s.offEps = o; o += (uint32_t)s.neb * s.eps_bytes;
  s.offW = o; o += (uint32_t)s.nst * s.w_bytes;
  s.total = o;
  return s;
}

__global__ void __launch_bounds__(kDhThreads, 1) mvn_bwd_dh_kernel(const MvnEncParams P, const int bulk) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t bar_vfull[kDhAStages], bar_vfree[kDhAStages], bar_wfull[kDhMaxStages], bar_wempty[kDhMaxStages], bar_efull[2], bar_efree[2], bar_dfull, bar_dfree;
  __shared__ uint32_t tmem_base_sm;
  const MirtParams& m = P.m;
  const int K = m.K, S = m.S, H = P.H, NH = mv_nh(H), NB = mv_nb(K), NU = NB * (NB + 1) / 2;
  const DhSmem L = dh_smem(K, H, bulk != 0);
  const int ES = L.ES, NEB = L.neb;
  const uint32_t NST = (uint32_t)L.nst;
  unsigned char* Wb = smem + L.offW;
  const uint32_t RSW = (uint32_t)(NH / 8) * kCSb;
  // V = g eps^T is the A operand and lives in TENSOR memory (lane = person, a unit of 64 columns = 32 + 32 words of
  // fp16 hi / lo pairs, written by the generating threads with tcgen05.st): it never crosses shared memory, whose
  // read port an SS-mode 128 x 64 x 16 MMA would saturate (6 KB per 32 cycles).  Columns [0, 256): accumulators,
  // [256, 512): kDhAStages units of V.
  // The tensor core adds into its fp32 accumulator with truncation: a long chain of accumulating MMAs drifts by
  // ~2^-24 of the running sum per instruction (measured: 2e-5 relative over the 1092 MMAs of a K = 100 tile).
  // The units therefore go round-robin into NACC independent accumulators, summed (round-to-nearest) at the drain.
  // (an even number: the two issuing warps take alternate units and must never share an accumulator -- the order of
  //  two warps' MMAs is not fixed, and the results are promised bit-identical run to run)
  const uint32_t NACC = (256u / (uint32_t)NH < 8u ? 256u / (uint32_t)NH : 8u) & ~1u;
  constexpr uint32_t kACol = 256;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t n_tiles = (m.B + kMvTP - 1) / kMvTP;
  const int64_t my_tiles = blockIdx.x < n_tiles ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const int64_t n_steps = my_tiles * S;                // (tile, particle) pairs of this CTA
  const float* eps_src = m.eps ? m.eps : P.eps_out;
  // V is an fp16 hi/lo operand: G is scaled by a power of two so that max |G'| <= 16 (exact; undone at the drain)
  int gexp;
  frexpf(fmaxf(*P.gmax, m.scale), &gexp);
  const float ginv = ldexpf(1.f, 4 - gexp), gfix = ldexpf(1.f, gexp - 4);

  if (tid == 0) {
    for (int i = 0; i < kDhAStages; ++i) {
      umma::mbar_init(&bar_vfull[i], kBwGenWarps);
      umma::mbar_init(&bar_vfree[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      umma::mbar_init(&bar_efull[i], 1);
      umma::mbar_init(&bar_efree[i], kBwGenWarps);
    }
    for (int i = 0; i < kDhMaxStages; ++i) { umma::mbar_init(&bar_wfull[i], 1); umma::mbar_init(&bar_wempty[i], 1); }
    umma::mbar_init(&bar_dfull, kDhMmaWarps);
    umma::mbar_init(&bar_dfree, 4);
    umma::fence_mbar_init();
  }
  // The last row-octet of the last person of a tile reaches past the rows a bulk copy brings in (its products meet
  // zero weights, so the values only have to be finite): clear that tail once.
  for (int e = 0; e < NEB; ++e)
    for (uint32_t o = (uint32_t)kMvTP * ES * 4 + 4 * tid; o < L.eps_bytes; o += 4 * kDhThreads)
      *reinterpret_cast<uint32_t*>(smem + L.offEps + (size_t)e * L.eps_bytes + o) = 0u;
  if (warp == kBwGenWarps + 1) umma::tmem_alloc(&tmem_base_sm, 512);
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem = tmem_base_sm;

  // The producer and the MMA warp walk their loops as whole warps (waits by every lane) and put only the issuing
  // instructions under elect_one(): see umma::elect_one -- behind `if (lane == 0)` one MMA costs ~115 cycles to
  // issue, more than three times what a 128 x 64 x 16 MMA takes to execute.
  if (warp == kBwGenWarps) {
    {                                                  // bulk-copy producer: the units of the backward image, the draws
      auto issue_eps = [&](int64_t e) {                // split draws of step e -> buffer e % NEB
        const uint32_t eb = (uint32_t)(e % NEB);
        const int64_t tile = blockIdx.x + (e / S) * gridDim.x;
        umma::mbar_wait(&bar_efree[eb], (uint32_t)((((e / NEB) & 1)) ^ 1));
        if (umma::elect_one())
          umma::bulk_load(smem + L.offEps + (size_t)eb * L.eps_bytes,
                          P.eps_split + ((int64_t)(e % S) * m.B + tile * kMvTP) * K, (uint32_t)kMvTP * K * 4, &bar_efull[eb]);
        __syncwarp();
      };
      uint32_t it = 0;
      if (bulk && n_steps > 0) issue_eps(0);
      for (int64_t n = 0; n < n_steps; ++n) {
        if (bulk && NEB == 2 && n + 1 < n_steps) issue_eps(n + 1);
        for (int u = 0; u < NU; ++u, ++it) {
          const uint32_t st = it % NST, round = it / NST;
          umma::mbar_wait(&bar_wempty[st], (round & 1) ^ 1);
          if (umma::elect_one())
            umma::bulk_load(Wb + (size_t)st * L.w_bytes, P.wimg_b + (size_t)u * L.w_bytes, L.w_bytes, &bar_wfull[st]);
          __syncwarp();
        }
        if (bulk && NEB == 1 && n + 1 < n_steps) issue_eps(n + 1);
      }
    }
  } else if (warp > kBwGenWarps) {
    {                                                  // MMA issuers: two warps on alternate units (see mvn_bwd_dw_kernel)
      const uint32_t mw = (uint32_t)(warp - (kBwGenWarps + 1));
      const uint32_t idesc = umma::make_idesc_f16(kMvTP, NH, umma::kMajorK, umma::kMajorMN);
      const uint64_t dW0 = umma::make_desc(umma::smem_u32(Wb), RSW, kCSb);     // 64 W22 unit: MN-major B
      uint32_t it = 0, ntile = 0;
      for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++ntile) {
        umma::mbar_wait(&bar_dfree, (ntile & 1) ^ 1);
        umma::fence_after_sync();
        uint32_t nu = 0;
        for (int s = 0; s < S; ++s)
          for (int u = 0; u < NU; ++u, ++it, ++nu) {
            if ((it % kDhMmaWarps) != mw) continue;
            const uint32_t vb = it % kDhAStages, st = it % NST;
            const uint32_t ai = nu % NACC;               // round-robin over the accumulators (see NACC)
            umma::mbar_wait(&bar_vfull[vb], (it / kDhAStages) & 1);
            umma::mbar_wait(&bar_wfull[st], (it / NST) & 1);
            umma::fence_after_sync();
            if (umma::elect_one()) {
              uint32_t acc = nu < NACC ? 0u : 1u;
              // B descriptors: one base per tile, the k-step only moves the 14-bit start-address field (units of 16 bytes)
              const uint64_t dWh = dW0 + (uint64_t)((st * L.w_bytes) >> 4), dWl = dWh + (uint64_t)(L.w_bytes >> 5);
              const uint32_t aH = tmem + kACol + vb * 64, aL = aH + 32;      // a k-step = 16 columns of V = 8 words
              const uint32_t td = tmem + ai * NH;
#pragma unroll
              for (int ks = 0; ks < 4; ++ks) {
                const uint64_t kb = (uint64_t)ks * (uint64_t)((2 * RSW) >> 4);
                umma::mma_f16_ts(td, aH + 8 * ks, dWh + kb, idesc, acc);       // hi * hi
                umma::mma_f16_ts(td, aH + 8 * ks, dWl + kb, idesc, 1u);        // hi * lo
                umma::mma_f16_ts(td, aL + 8 * ks, dWh + kb, idesc, 1u);        // lo * hi
                acc = 1;
              }
              umma::commit(&bar_vfree[vb]);
              umma::commit(&bar_wempty[st]);
            }
            __syncwarp();
          }
        if (umma::elect_one()) umma::commit(&bar_dfull);
        __syncwarp();
      }
    }
  } else {
    const int t = tid & 127, sub = tid >> 7;           // thread = (person, rows 4 sub .. 4 sub + 3 of every unit)
    const bool has_loc = P.w21 != nullptr;
    // TMEM lane = person (warp w reaches lanes 32 (w % 4) ..: t >> 5 == warp % 4); row r of a unit = words 4 r .. 4 r + 3
    const uint32_t ta = tmem + ((uint32_t)((warp & 3) * 32) << 16) + kACol + 16u * (uint32_t)sub;
    uint32_t it = 0;
    for (int64_t n = 0; n < n_steps; ++n) {
      const uint32_t eb = (uint32_t)(n % NEB);
      const int64_t tile = blockIdx.x + (n / S) * gridDim.x;
      const int s = (int)(n % S);
      const int64_t b = tile * kMvTP + t;
      const bool live = b < m.B;
      uint32_t* eps_t = reinterpret_cast<uint32_t*>(smem + L.offEps + (size_t)eb * L.eps_bytes);
      if (!bulk) {                                     // injected noise / K % 4 != 0: rows split here, padded pitch
        gen_bar();                                     // everyone has finished with the previous draws
        const float* src = eps_src + (int64_t)s * m.B * K;
        for (int pp = warp * (kMvTP / kBwGenWarps); pp < (warp + 1) * (kMvTP / kBwGenWarps); ++pp) {
          const int64_t bb = tile * kMvTP + pp;
          for (int mm = lane; mm < ES / 2; mm += 32) {
            const float z0 = (bb < m.B && 2 * mm < K) ? src[bb * K + 2 * mm] : 0.f;
            const float z1 = (bb < m.B && 2 * mm + 1 < K) ? src[bb * K + 2 * mm + 1] : 0.f;
            split_pair(z0, z1, eps_t[(size_t)pp * ES + 2 * mm], eps_t[(size_t)pp * ES + 2 * mm + 1]);
          }
        }
        gen_bar();
      } else {
        umma::mbar_wait(&bar_efull[eb], (uint32_t)((n / NEB) & 1));
      }
      const uint32_t* er = eps_t + (size_t)t * ES;
      float gI[4], lI[4], dadd[4], gN[4], lN[4], dN[4];      // this block row / the next one, fetched a block row ahead
      load_g<4>(m, s, b, live, 0, sub, ginv, gI, lI, dadd);
      for (int I = 0; I < NB; ++I) {
        load_g<4>(m, s, b, live && I + 1 < NB, I + 1, sub, ginv, gN, lN, dN);
        __half2 gh[4], gl[4];                                // the row factors, split once per block row
#pragma unroll
        for (int rr = 0; rr < 4; ++rr) split_g(gI[rr], gh[rr], gl[rr]);
        for (int Jc = 0; Jc <= I; ++Jc, ++it) {
          const uint32_t vb = it % kDhAStages;
          const uint4 w0 = *reinterpret_cast<const uint4*>(er + 8 * Jc), w1 = *reinterpret_cast<const uint4*>(er + 8 * Jc + 4);
          uint32_t vh[16], vl[16];
          if (Jc < I) {                                      // CTA-uniform: the plain units carry no diagonal code
#pragma unroll
            for (int rr = 0; rr < 4; ++rr) {
              uint4 hi, lo;
              gen_octet(gh[rr], gl[rr], w0, w1, hi, lo);
              vh[4 * rr] = hi.x; vh[4 * rr + 1] = hi.y; vh[4 * rr + 2] = hi.z; vh[4 * rr + 3] = hi.w;
              vl[4 * rr] = lo.x; vl[4 * rr + 1] = lo.y; vl[4 * rr + 2] = lo.z; vl[4 * rr + 3] = lo.w;
            }
          } else {
            // fc21 fused: the slots above the diagonal hold V = G_r x 1 for the location rows (mv_loc_slot): rows 0 and 1
            // of the unit, i.e. the sub == 0 thread, which fetches the block row's other four G values for it
            float gloc[8];
            if (has_loc && sub == 0) {
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const int i = 8 * I + 4 + q;
                gloc[q] = gI[q];
                gloc[4 + q] = (live && i < K) ? ginv * m.gx[((int64_t)s * m.B + b) * K + i] : 0.f;
              }
            }
#pragma unroll
            for (int rr = 0; rr < 4; ++rr) {
              uint4 hi, lo;
              gen_octet(gh[rr], gl[rr], w0, w1, hi, lo);
              fix_diag(gI[rr], lI[rr], dadd[rr], er + 8 * Jc, 4 * sub + rr, hi, lo);
              if (has_loc && sub == 0) {
                if (rr == 0) {
#pragma unroll
                  for (int c = 1; c < 8; ++c) put_elem(hi, lo, c, gloc[c]);
                } else if (rr == 1) {
                  put_elem(hi, lo, 2, gloc[0]);
                }
              }
              vh[4 * rr] = hi.x; vh[4 * rr + 1] = hi.y; vh[4 * rr + 2] = hi.z; vh[4 * rr + 3] = hi.w;
              vl[4 * rr] = lo.x; vl[4 * rr + 1] = lo.y; vl[4 * rr + 2] = lo.z; vl[4 * rr + 3] = lo.w;
            }
          }
          umma::mbar_wait(&bar_vfree[vb], ((it / kDhAStages) & 1) ^ 1);
          umma::fence_after_sync();
          umma::tmem_st16(ta + vb * 64, vh);
          umma::tmem_st16(ta + vb * 64 + 32, vl);
          umma::tmem_st_wait();
          umma::fence_before_sync();
          warp_arrive(&bar_vfull[vb], lane);
        }
#pragma unroll
        for (int rr = 0; rr < 4; ++rr) { gI[rr] = gN[rr]; lI[rr] = lN[rr]; dadd[rr] = dN[rr]; }
      }
      if (bulk) warp_arrive(&bar_efree[eb], lane);     // done with this step's draws
      if (s == S - 1 && warp < 4) {                    // drain d hidden of this tile: thread = person = TMEM lane
        const uint32_t ntile = (uint32_t)(n / S);
        umma::mbar_wait(&bar_dfull, ntile & 1);
        umma::fence_after_sync();
        const uint32_t tl = tmem + ((uint32_t)(warp * 32) << 16);
        const float dfix = gfix / kMvWScale;
        const uint32_t n_used = (uint32_t)(S * NU) < NACC ? (uint32_t)(S * NU) : NACC;
        for (int c0 = 0; c0 < NH; c0 += 16) {
          float v[16];
          umma::tmem_ld16(tl + c0, v);
          for (uint32_t ai = 1; ai < n_used; ++ai) {
            float w[16];
            umma::tmem_ld16(tl + ai * NH + c0, w);
#pragma unroll
            for (int q = 0; q < 16; ++q) v[q] += w[q];
          }
          if (live && P.d_hidden) {
            float* dst = P.d_hidden + b * H + c0;
            if ((H & 3) == 0 && c0 + 16 <= H) {
#pragma unroll
              for (int q = 0; q < 4; ++q)
                *reinterpret_cast<float4*>(dst + 4 * q) = make_float4(v[4 * q] * dfix, v[4 * q + 1] * dfix,
                                                                      v[4 * q + 2] * dfix, v[4 * q + 3] * dfix);
            } else {
#pragma unroll
              for (int q = 0; q < 16; ++q)
                if (c0 + q < H) dst[q] = v[q] * dfix;
            }
          }
        }
        umma::fence_before_sync();
        warp_arrive(&bar_dfree, lane);
      }
    }
  }
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  if (warp == kBwGenWarps + 1) umma::tmem_dealloc(tmem, 512);
}

// ---- d [W22 | b22][(i,j)][h] = sum over persons of V[person][(i,j)] [hidden | 1][person][h] -------------------------
// CTA (pg, qg): chunks [pg * CPB, ..) of the factor columns x 64-person tiles qg, qg + n_qg, ..; the accumulators of
// the slab stay in tensor memory; partial results [qg] are summed in a fixed order by mvn_dw_finalize.
//
// V^T is the A operand (M = factor column, K = person) and lives in TENSOR memory: a generating thread owns one factor
// column (i, j) = its TMEM lane and forms g_i[b] eps_j[b] for 32 persons in packed half arithmetic from two rows of
// person-contiguous, pre-split fp16 pairs -- rows that mvn_dwprep_kernel lays out once per step (transposed tiles of
// G, of the draws and of the diagonal terms).  Nothing is transposed, split or converted in this kernel, V never
// crosses shared memory (whose read port an SS-mode MMA would saturate), and a chunk costs ~90 instructions per
// thread instead of ~290.
constexpr int kDwTP = 64;                  // persons per tile (UMMA K extent: 4 k-steps)
constexpr int kDwFlushTiles = 16;          // accumulate in TMEM for at most 16 tiles = 192 MMAs per accumulator
constexpr int kDwMaxRows = 4;              // block rows of L that a slab of <= 4 chunks (8 consecutive units) can touch
constexpr int kDwAStages = 3;              // V^T chunks in tensor memory: 3 x (32 hi + 32 lo) columns at 320..511
constexpr uint32_t kDwACol = 512 - 64 * kDwAStages;
constexpr uint32_t kDwGRec = 512;          // per row i of a tile: [g hi | g lo | diag hi | diag lo], 64 persons x fp16 each
constexpr int kDwGenWarps = 16;            // generator warps: 4 per TMEM lane quadrant, 16 persons of the tile each
constexpr int kDwMmaWarps = 2;             // MMA-issuing warps (alternate chunks)
constexpr int kDwThreads = (kDwGenWarps + 1 + kDwMmaWarps) * 32;    // + bulk-copy producer warp + MMA warps
constexpr int kDwPH = kDwGenWarps / 4;     // person groups per tile
constexpr int kDwPPT = kDwTP / kDwPH;      // persons per generator thread
constexpr int kDwFG = (5 + kDwPH - 1) / kDwPH;       // groups of 16 accumulator columns per warp in a flush (KH <= 80)
constexpr uint32_t kDwERec = 272;          // per column j of a tile: [eps hi | eps lo | 16 B pad] (pad: 8 columns -> 32 banks)
__host__ __device__ inline int dw_cpb(int H) {                                     // chunks per CTA slab
  const int c = (int)kDwACol / mv_kh(H);
  return c < 4 ? c : 4;
}
__host__ __device__ inline size_t mv_himg_tile_bytes(int H) { return (size_t)2 * kDwTP * mv_kh(H) * 2; }
__host__ __device__ inline size_t mv_vimg_tile_bytes(int K) { return (size_t)8 * mv_nb(K) * (kDwGRec + kDwERec); }

constexpr int kDwBufs = 3;                 // step buffers ([hidden | 1] tile, draws, G rows): copies run two steps ahead
struct DwSmem {
  uint32_t offH, offE, offG, buf_bytes, offConst, total, h_bytes, e_bytes, g_bytes;
  int nbuf;
};
__host__ __device__ inline DwSmem dw_smem(int K, int H) {
  DwSmem s;
  s.h_bytes = (uint32_t)mv_himg_tile_bytes(H);
  s.e_bytes = (uint32_t)(8 * mv_nb(K)) * kDwERec;
  s.g_bytes = (uint32_t)kDwMaxRows * 8 * kDwGRec;
  s.offH = 0;
  s.offE = s.h_bytes;
  s.offG = s.offE + s.e_bytes;
  s.buf_bytes = (s.offG + s.g_bytes + 127) & ~127u;
  s.nbuf = (227u * 1024u - 1024u) / s.buf_bytes >= (uint32_t)kDwBufs ? kDwBufs : 2;
  s.offConst = (uint32_t)s.nbuf * s.buf_bytes;        // 512 B: a row of ones | zeros (fp16), then 256 B of zeros
  s.total = s.offConst + 512;
  return s;
}

// [hidden | 1 | 0] of 64 persons as fp16 hi / lo operand tiles T[r = person][c = h], one 2 x 64 x KH x 2 byte image per tile
__global__ void mvn_himg_kernel(const float* __restrict__ hidden, int64_t B, int H, int KH, unsigned char* __restrict__ img) {
  const int n_oct = KH / 8;
  const int64_t n_rows = (B + kDwTP - 1) / kDwTP * kDwTP;
  const uint32_t RSH = (uint32_t)n_oct * kCSb;
  const size_t tile_bytes = (size_t)kDwTP * KH * 2;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n_rows * n_oct; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = e / n_oct;
    const int o = (int)(e - b * n_oct);
    float v[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int k = o * 8 + q;
      v[q] = b >= B ? 0.f : (k < H ? hidden[b * H + k] : (k == H ? 1.f : 0.f));
    }
    uint4 hi, lo;
    split8(v, hi, lo);
    const int r = (int)(b % kDwTP);
    unsigned char* base = img + (size_t)(b / kDwTP) * 2 * tile_bytes + (size_t)(r >> 3) * RSH + (size_t)o * kCSb + (r & 7) * 16;
    *reinterpret_cast<uint4*>(base) = hi;
    *reinterpret_cast<uint4*>(base + tile_bytes) = lo;
  }
}

// The generator's operands, once per step: for every (particle, 64-person tile) a block of person-contiguous rows
//   row i   : g_i[b] hi | g_i[b] lo | v_ii[b] hi | v_ii[b] lo      v_ii = g_i eps_i L_ii - scale [s == 0]   (diagonal of V)
//   column j: eps_j[b] hi | eps_j[b] lo
// with g = G * 2^-e such that max |g| <= 16 (exact; undone by the drain), zeros for dead persons and rows beyond K.
__global__ void __launch_bounds__(256) mvn_dwprep_kernel(const MvnEncParams P, const float* __restrict__ eps_f32,
                                                          unsigned char* __restrict__ vimg) {
  const MirtParams& m = P.m;
  const int K = m.K, NB = mv_nb(K), K8 = 8 * NB;
  const int64_t n_tiles = (m.B + kDwTP - 1) / kDwTP;
  int gexp;
  frexpf(fmaxf(*P.gmax, m.scale), &gexp);
  const float ginv = ldexpf(1.f, 4 - gexp);
  // a warp = eight rows i (a block row of L), a lane = two persons: 32-byte pieces of their G / draw / diag L rows in,
  // 128-byte row segments out
  const int bp = threadIdx.x & 31, w = threadIdx.x >> 5;
  const bool vec = P.eps_split != nullptr;               // K % 4 == 0: whole 16-byte pieces
  for (int64_t step = blockIdx.x; step < n_tiles * m.S; step += gridDim.x) {
    const int s = (int)(step / n_tiles);
    const int64_t tile = step - (int64_t)s * n_tiles;
    unsigned char* grec = vimg + (size_t)step * mv_vimg_tile_bytes(K);
    unsigned char* erec = grec + (size_t)K8 * kDwGRec;
    for (int o = w; o < NB; o += 8) {
      float g[2][8], e[2][8], l[2][8];
      uint32_t ehw[2][4], elw[2][4];                       // the draws as fp16 pairs along i (hi / lo)
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int64_t b = tile * kDwTP + 2 * bp + q;
        const int64_t row = ((int64_t)s * m.B + b) * K + 8 * o;
        if (vec) {
#pragma unroll
          for (int h4 = 0; h4 < 2; ++h4) {
            const bool ok = b < m.B && 8 * o + 4 * h4 < K;
            const float4 gv = ok ? *reinterpret_cast<const float4*>(m.gx + row + 4 * h4) : make_float4(0.f, 0.f, 0.f, 0.f);
            const float4 lv = ok ? *reinterpret_cast<const float4*>(m.diagL + b * K + 8 * o + 4 * h4) : make_float4(0.f, 0.f, 0.f, 0.f);
            const uint4 ev = ok ? *reinterpret_cast<const uint4*>(P.eps_split + row + 4 * h4) : make_uint4(0u, 0u, 0u, 0u);
            g[q][4 * h4] = gv.x; g[q][4 * h4 + 1] = gv.y; g[q][4 * h4 + 2] = gv.z; g[q][4 * h4 + 3] = gv.w;
            l[q][4 * h4] = lv.x; l[q][4 * h4 + 1] = lv.y; l[q][4 * h4 + 2] = lv.z; l[q][4 * h4 + 3] = lv.w;
            ehw[q][2 * h4] = ev.x; elw[q][2 * h4] = ev.y; ehw[q][2 * h4 + 1] = ev.z; elw[q][2 * h4 + 1] = ev.w;
          }
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const float2 hf = __half22float2(as_h2(ehw[q][t])), lf = __half22float2(as_h2(elw[q][t]));
            e[q][2 * t] = hf.x + lf.x;
            e[q][2 * t + 1] = hf.y + lf.y;
          }
        } else {
#pragma unroll
          for (int t = 0; t < 8; ++t) {
            const bool ok = b < m.B && 8 * o + t < K;
            g[q][t] = ok ? m.gx[row + t] : 0.f;
            l[q][t] = ok ? m.diagL[b * K + 8 * o + t] : 0.f;
            e[q][t] = ok ? eps_f32[row + t] : 0.f;
          }
#pragma unroll
          for (int t = 0; t < 4; ++t) split_pair(e[q][2 * t], e[q][2 * t + 1], ehw[q][t], elw[q][t]);
        }
      }
#pragma unroll
      for (int t = 0; t < 8; ++t) {
        const int i = 8 * o + t;
        float gg[2], vd[2];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const bool live = tile * kDwTP + 2 * bp + q < m.B && i < K;
          gg[q] = ginv * g[q][t];
          vd[q] = fmaf(gg[q] * e[q][t], l[q][t], (live && s == 0) ? -m.scale * ginv : 0.f);
        }
        uint32_t wv[4];
        split_pair(gg[0], gg[1], wv[0], wv[1]);
        split_pair(vd[0], vd[1], wv[2], wv[3]);
        // the two persons' halves of element t: low / high half of word t >> 1
        const uint32_t sh = (t & 1) ? 16u : 0u;
        const uint32_t eh = ((ehw[0][t >> 1] >> sh) & 0xffffu) | (((ehw[1][t >> 1] >> sh) & 0xffffu) << 16);
        const uint32_t el = ((elw[0][t >> 1] >> sh) & 0xffffu) | (((elw[1][t >> 1] >> sh) & 0xffffu) << 16);
        uint32_t* gr = reinterpret_cast<uint32_t*>(grec + (size_t)i * kDwGRec) + bp;
        gr[0] = wv[0]; gr[32] = wv[1]; gr[64] = wv[2]; gr[96] = wv[3];
        uint32_t* er = reinterpret_cast<uint32_t*>(erec + (size_t)i * kDwERec) + bp;
        er[0] = eh; er[32] = el;
      }
    }
  }
}

__device__ __forceinline__ void gen_pairs(const uint4& gh, const uint4& gl, const uint4& eh, const uint4& el, uint32_t* vh,
                                          uint32_t* vl) {
  auto one = [&](uint32_t a, uint32_t al, uint32_t e, uint32_t e_l, uint32_t& oh, uint32_t& ol) {
    const __half2 h = __hmul2(as_h2(a), as_h2(e));
    __half2 r = __hfma2(as_h2(a), as_h2(e), __hneg2(h));
    r = __hfma2(as_h2(a), as_h2(e_l), r);
    r = __hfma2(as_h2(al), as_h2(e), r);
    oh = h2_bits(h);
    ol = h2_bits(r);
  };
  one(gh.x, gl.x, eh.x, el.x, vh[0], vl[0]);
  one(gh.y, gl.y, eh.y, el.y, vh[1], vl[1]);
  one(gh.z, gl.z, eh.z, el.z, vh[2], vl[2]);
  one(gh.w, gl.w, eh.w, el.w, vh[3], vl[3]);
}

__global__ void __launch_bounds__(kDwThreads, 1) mvn_bwd_dw_kernel(const MvnEncParams P, const unsigned char* __restrict__ vimg,
                                                                    const int n_pg, const int n_qg) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t bar_efull[kDwBufs], bar_tfree[kDwBufs], bar_vfull[kDwAStages], bar_vfree[kDwAStages], bar_done, bar_drained;
  __shared__ uint32_t tmem_base_sm;
  __shared__ int utab[8][2];                          // (I, Jc) of the slab's units; I >= NB: beyond the factor
  const MirtParams& m = P.m;
  const int K = m.K, S = m.S, H = P.H, KH = mv_kh(H), NB = mv_nb(K), NU = NB * (NB + 1) / 2;
  const DwSmem L = dw_smem(K, H);
  const int CPB = dw_cpb(H), n_chunks = mv_chunks(K);
  const uint32_t RSH = (uint32_t)(KH / 8) * kCSb;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int pg = blockIdx.x % n_pg, qg = blockIdx.x / n_pg;
  const int c0 = pg * CPB, c1 = min(c0 + CPB, n_chunks), nblk = c1 - c0;
  const int64_t n_tiles = (m.B + kDwTP - 1) / kDwTP;
  const int64_t my_tiles = qg < n_tiles ? (n_tiles - qg + n_qg - 1) / n_qg : 0;       // tiles qg, qg + n_qg, ..
  const int64_t n_steps = my_tiles * S;                                              // (tile, particle) pairs
  int gexp;
  frexpf(fmaxf(*P.gmax, m.scale), &gexp);
  const float gfix = ldexpf(1.f, gexp - 4);

  if (tid == 0) {
    for (int i = 0; i < kDwBufs; ++i) {
      umma::mbar_init(&bar_efull[i], 3);               // three bulk copies per step: [hidden | 1], draws, G rows
      umma::mbar_init(&bar_tfree[i], kDwGenWarps + kDwMmaWarps);
    }
    for (int i = 0; i < kDwAStages; ++i) {
      umma::mbar_init(&bar_vfull[i], kDwGenWarps);
      umma::mbar_init(&bar_vfree[i], 1);
    }
    umma::mbar_init(&bar_done, kDwMmaWarps);
    umma::mbar_init(&bar_drained, kDwGenWarps);
    umma::fence_mbar_init();
  }
  if (tid < 128) {                                     // constant rows: 64 x 1.0 | 64 x 0 (fp16), then 256 B of zeros
    uint32_t* cr = reinterpret_cast<uint32_t*>(smem + L.offConst);
    cr[tid] = tid < 32 ? 0x3c003c00u : 0u;
  }
  if (tid < 2 * nblk) {
    int I, Jc;
    unit_to_block(2 * c0 + tid, I, Jc);
    utab[tid][0] = 2 * c0 + tid < NU ? I : NB;
    utab[tid][1] = Jc;
  }
  if (warp == kDwGenWarps + 1) umma::tmem_alloc(&tmem_base_sm, 512);
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem = tmem_base_sm;
  const int I_first = utab[0][0];
  const int I_last = utab[2 * nblk - 1][0] < NB ? utab[2 * nblk - 1][0] : NB - 1;
  const int n_rows = I_last - I_first + 1;             // block rows of the slab (<= kDwMaxRows)
  const uint32_t e_copy = (uint32_t)(8 * (I_last + 1)) * kDwERec, g_copy = (uint32_t)(8 * n_rows) * kDwGRec;
  const int64_t NBUF = L.nbuf;

  if (warp == kDwGenWarps) {
    // bulk-copy producer: [hidden | 1] tile, the columns of the draws and the rows of G this slab needs, of step n
    for (int64_t n = 0; n < n_steps; ++n) {
      const uint32_t buf = (uint32_t)(n % NBUF);
      const int64_t tile = qg + (n / S) * n_qg;
      const int s = (int)(n % S);
      umma::mbar_wait(&bar_tfree[buf], (uint32_t)(((n / NBUF) & 1) ^ 1));
      if (umma::elect_one()) {
        unsigned char* dst = smem + (size_t)buf * L.buf_bytes;
        const unsigned char* src = vimg + (size_t)((int64_t)s * n_tiles + tile) * mv_vimg_tile_bytes(K);
        umma::bulk_load(dst + L.offH, P.himg + (size_t)tile * L.h_bytes, L.h_bytes, &bar_efull[buf]);
        umma::bulk_load(dst + L.offE, src + (size_t)(8 * NB) * kDwGRec, e_copy, &bar_efull[buf]);
        umma::bulk_load(dst + L.offG, src + (size_t)(8 * I_first) * kDwGRec, g_copy, &bar_efull[buf]);
      }
      __syncwarp();
    }
  } else if (warp > kDwGenWarps) {
    // MMA issuers: whole-warp loops, issue under elect_one().  TWO warps take alternate chunks: a wait / commit pair
    // around a group of MMAs costs the issuing warp ~350 cycles in which it feeds nothing to the tensor pipe -- with two
    // issuers one warp's round trip hides under the other's MMAs (the chunks go to different accumulators; commits
    // track the MMAs of the committing thread).
    const uint32_t mw = (uint32_t)(warp - (kDwGenWarps + 1));
    const uint32_t idesc = umma::make_idesc_f16(kMvCN, KH, umma::kMajorK, umma::kMajorMN);
    const uint64_t dH0 = umma::make_desc(umma::smem_u32(smem + L.offH), RSH, kCSb);     // [hidden | 1]: MN-major B
    uint32_t it = 0, nflush = 0;
    bool fresh = true;                                 // the accumulators hold nothing yet (start, or just flushed)
    for (int64_t n = 0; n < n_steps; ++n) {
      const uint32_t buf = (uint32_t)(n % NBUF);
      umma::mbar_wait(&bar_efull[buf], (uint32_t)((n / NBUF) & 1));
      for (int k = 0; k < nblk; ++k, ++it) {
        if ((it % kDwMmaWarps) != mw) continue;
        const uint32_t vb = it % kDwAStages;
        umma::mbar_wait(&bar_vfull[vb], (it / kDwAStages) & 1);
        umma::fence_after_sync();
        if (umma::elect_one()) {
          const uint64_t dHh = dH0 + (uint64_t)((buf * L.buf_bytes) >> 4), dHl = dHh + (uint64_t)(L.h_bytes >> 5);
          const uint32_t aH = tmem + kDwACol + vb * 64, aL = aH + 32;      // a k-step = 16 persons = 8 words
          const uint32_t td = tmem + k * KH;
          uint32_t acc = fresh ? 0u : 1u;
#pragma unroll
          for (int ks = 0; ks < kDwTP / 16; ++ks) {
            const uint64_t kb = (uint64_t)ks * (uint64_t)((2 * RSH) >> 4);
            umma::mma_f16_ts(td, aH + 8 * ks, dHh + kb, idesc, acc);         // hi * hi
            umma::mma_f16_ts(td, aH + 8 * ks, dHl + kb, idesc, 1u);          // hi * lo
            umma::mma_f16_ts(td, aL + 8 * ks, dHh + kb, idesc, 1u);          // lo * hi
            acc = 1;
          }
          umma::commit(&bar_vfree[vb]);
        }
        __syncwarp();
      }
      fresh = false;
      const bool flush = (n + 1) % kDwFlushTiles == 0 || n + 1 == n_steps;
      if (umma::elect_one()) {
        umma::commit(&bar_tfree[buf]);                 // every MMA of this warp that reads this step's H tile has completed
        if (flush) umma::commit(&bar_done);            // flush: the drain warps add the accumulators to the partial rows
      }
      __syncwarp();
      if (flush) {
        umma::mbar_wait(&bar_drained, nflush & 1);
        umma::fence_after_sync();
        fresh = true;
        ++nflush;
      }
    }
  } else {
    // generators: thread = factor column c of the chunk (its TMEM lane) x half of the tile's persons.  The slab is
    // fixed, so the rows a thread multiplies are fixed offsets into the step buffer, found once: its G row (or, on
    // the diagonal of L, the prepared v_ii row against a row of ones; above the diagonal and beyond the factor a
    // row of zeros) and its column of the draws -- the loop itself has no branches.
    const int q4 = warp & 3, hp = warp >> 2;           // lane quadrant; persons kDwPPT hp .. of the tile
    const int half = q4 >> 1, r = 4 * (q4 & 1) + (lane >> 3), jj = lane & 7;
    const uint32_t ta = tmem + ((uint32_t)(32 * q4) << 16) + kDwACol + (uint32_t)(kDwPPT / 2 * hp);
    const uint32_t sbase = umma::smem_u32(smem);
    uint32_t goff[4], eoff[4];                         // byte offsets; >= offConst: constant rows (not per buffer)
    const uint32_t kConst = 0x80000000u;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int I = k < nblk ? utab[2 * k + half][0] : NB, Jc = k < nblk ? utab[2 * k + half][1] : 0;
      uint32_t go = kConst | (L.offConst + 256u), eo = kConst | L.offConst;          // zeros
      if (I < NB) {
        const uint32_t grow = L.offG + (uint32_t)((I - I_first) * 8 + r) * kDwGRec + (uint32_t)(2 * kDwPPT * hp);
        const uint32_t erow = L.offE + (uint32_t)(8 * Jc + jj) * kDwERec + (uint32_t)(2 * kDwPPT * hp);
        if (Jc != I || jj < r) { go = grow; eo = erow; }
        else if (jj == r) { go = grow + 256u; eo = kConst | L.offConst; }
        else if (P.w21) {                              // fc21 fused: a location slot is G of its row against the row of ones
          const int rl = mv_loc_row_of_slot(8 * r + jj);
          if (rl >= 0) {
            go = L.offG + (uint32_t)((I - I_first) * 8 + rl) * kDwGRec + (uint32_t)(2 * kDwPPT * hp);
            eo = kConst | L.offConst;
          }
        }
      }
      goff[k] = go; eoff[k] = eo;
    }
    auto lds128 = [](uint32_t a) {
      uint4 v;
      asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
      return v;
    };
    uint32_t it = 0, nflush = 0;
    for (int64_t n = 0; n < n_steps; ++n) {
      const uint32_t buf = (uint32_t)(n % NBUF);
      const uint32_t bb = sbase + buf * L.buf_bytes;
      umma::mbar_wait(&bar_efull[buf], (uint32_t)((n / NBUF) & 1));
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (k < nblk) {
          const uint32_t vb = it % kDwAStages;
          const uint32_t ga = (goff[k] & kConst) ? sbase + (goff[k] & ~kConst) : bb + goff[k];
          const uint32_t ea = (eoff[k] & kConst) ? sbase + (eoff[k] & ~kConst) : bb + eoff[k];
          uint32_t vh[kDwPPT / 2], vl[kDwPPT / 2];
#pragma unroll
          for (int p8 = 0; p8 < kDwPPT / 8; ++p8) {
            // the constant rows are 128 B of one value: any 16-byte piece will do, so no person offset there
            const uint32_t po = 16u * p8;
            const uint4 gh = lds128(ga + ((goff[k] & kConst) ? 0u : po)), gl = lds128(ga + 128u + ((goff[k] & kConst) ? 0u : po));
            const uint4 eh = lds128(ea + ((eoff[k] & kConst) ? 0u : po)), el = lds128(ea + 128u + ((eoff[k] & kConst) ? 0u : po));
            gen_pairs(gh, gl, eh, el, vh + 4 * p8, vl + 4 * p8);
          }
          umma::mbar_wait(&bar_vfree[vb], ((it / kDwAStages) & 1) ^ 1);
          umma::fence_after_sync();
          umma::tmem_st8(ta + vb * 64, vh);
          umma::tmem_st8(ta + vb * 64 + 32, vl);
          umma::tmem_st_wait();
          umma::fence_before_sync();
          warp_arrive(&bar_vfull[vb], lane);
          ++it;
        }
      }
      warp_arrive(&bar_tfree[buf], lane);              // done with this step's rows
      // Flush every kDwFlushTiles steps: the accumulators are added (round-to-nearest, fixed order) to this CTA's
      // partial rows, which bounds the chain of truncating tensor-core accumulations (see mvn_bwd_dh_kernel).
      // Thread = factor column of the chunk = TMEM lane; the two warps of a lane quadrant take alternate groups of
      // 16 columns, and the old partial sums are on their way before the accumulators are waited for.
      if ((n + 1) % kDwFlushTiles == 0 || n + 1 == n_steps) {
        const int t = 32 * q4 + lane;
        const uint32_t tl = tmem + ((uint32_t)(q4 * 32) << 16);
        bool waited = false;
        for (int k = 0; k < nblk; ++k) {
          // partial rows [qg][chunk][h][column]: for one h the warp's 32 columns are 128 contiguous bytes
          float* dst = P.dw_part + ((size_t)qg * n_chunks + (size_t)(c0 + k)) * KH * kMvCN + t;
          float old[kDwFG][16];
#pragma unroll
          for (int gi = 0; gi < kDwFG; ++gi) {
            const int q0 = 16 * (kDwPH * gi + hp);
#pragma unroll
            for (int q = 0; q < 16; ++q) old[gi][q] = (nflush > 0 && q0 < KH) ? dst[(size_t)(q0 + q) * kMvCN] : 0.f;
          }
          if (!waited) {
            umma::mbar_wait(&bar_done, nflush & 1);
            umma::fence_after_sync();
            waited = true;
          }
#pragma unroll
          for (int gi = 0; gi < kDwFG; ++gi) {
            const int q0 = 16 * (kDwPH * gi + hp);
            if (q0 < KH) {
              float v[16];
              umma::tmem_ld16(tl + k * KH + q0, v);
#pragma unroll
              for (int q = 0; q < 16; ++q) dst[(size_t)(q0 + q) * kMvCN] = fmaf(v[q], gfix, old[gi][q]);
            }
          }
        }
        umma::fence_before_sync();
        warp_arrive(&bar_drained, lane);
        ++nflush;
      }
    }
    if (n_steps == 0 && warp < 4) {                    // a CTA without persons still owns its partial rows
      for (int k = 0; k < nblk; ++k) {
        float* dst = P.dw_part + ((size_t)qg * n_chunks + (size_t)(c0 + k)) * KH * kMvCN + tid;
        for (int q = 0; q < KH; ++q) dst[(size_t)q * kMvCN] = 0.f;
      }
    }
  }
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  if (warp == kDwGenWarps + 1) umma::tmem_dealloc(tmem, 512);
}

// d_w22[p][h], d_b22[p] = sum over the person groups, in order, of the partial rows of padded column mv_col(i, j)
__global__ void mvn_dw_finalize_kernel(const float* __restrict__ part, int n_qg, int n_cols, int K, int H, int KH,
                                       float* __restrict__ d_w22, float* __restrict__ d_b22, float* __restrict__ d_w21,
                                       float* __restrict__ d_b21) {
  const int P = K * (K + 1) / 2;
  // d fc21: location row i sits at slot mv_loc_slot(i % 8) of the diagonal unit of block row i / 8
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < K * (H + 1) && (d_w21 || d_b21); e += gridDim.x * blockDim.x) {
    const int i = e / (H + 1), h = e - i * (H + 1);
    const int I = i >> 3;
    const size_t pc = (size_t)64 * (I * (I + 1) / 2 + I) + mv_loc_slot(i & 7);
    float t = 0.f;
    for (int q = 0; q < n_qg; ++q) t += part[((size_t)q * n_cols + (pc & ~size_t(kMvCN - 1))) * KH + (size_t)h * kMvCN + (pc & (kMvCN - 1))];
    if (h < H) { if (d_w21) d_w21[(size_t)i * H + h] = t; }
    else if (d_b21) d_b21[i] = t;
  }
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < P * (H + 1); e += gridDim.x * blockDim.x) {
    const int p = e / (H + 1), h = e - p * (H + 1);
    int i = (int)((sqrtf(8.f * (float)p + 1.f) - 1.f) * 0.5f);
    while (tri_off(i + 1) <= p) ++i;
    while (tri_off(i) > p) --i;
    const int j = p - tri_off(i);
    const size_t pc = (size_t)mv_col(i, j);
    float t = 0.f;
    for (int q = 0; q < n_qg; ++q) t += part[((size_t)q * n_cols + (pc & ~size_t(kMvCN - 1))) * KH + (size_t)h * kMvCN + (pc & (kMvCN - 1))];
    if (h < H) { if (d_w22) d_w22[(size_t)p * H + h] = t; }
    else if (d_b22) d_b22[p] = t;
  }
}

}  // namespace


size_t mvn_wimg_bytes(int K, int H) { return ((size_t)mv_chunks(K) * mv_chunk_bytes(H) + 255) & ~size_t(255); }
size_t mvn_wimgb_bytes(int K, int H) {
  return ((size_t)2 * mv_chunks(K) * (2 * 64 * mv_nh(H) * 2) + 255) & ~size_t(255);
}
static void dw_grid(int K, int H, int* n_pg, int* n_qg) {
  const int cpb = dw_cpb(H), n_chunks = mv_chunks(K);
  *n_pg = (n_chunks + cpb - 1) / cpb;
  int q = num_sms() / *n_pg;
  *n_qg = q < 1 ? 1 : q;
}
size_t mvn_himg_bytes(int64_t B, int H) {
  return ((size_t)((B + kDwTP - 1) / kDwTP) * mv_himg_tile_bytes(H) + 255) & ~size_t(255);
}
size_t mvn_eps_bytes(int64_t B, int K, int S) {       // the draws + one tile of zero padding (bulk copies of the last tile)
  return (((size_t)S * B + kMvTP) * K * 4 + 64 + 255) & ~size_t(255);
}
size_t mvn_dwpart_bytes(int K, int H) {
  int n_pg, n_qg;
  dw_grid(K, H, &n_pg, &n_qg);
  return ((size_t)n_qg * mv_chunks(K) * kMvCN * mv_kh(H) * 4 + 255) & ~size_t(255);
}
size_t mvn_vimg_bytes(int64_t B, int K, int S) {       // the generator's operand rows (mvn_dwprep_kernel)
  return ((size_t)((B + kDwTP - 1) / kDwTP) * S * mv_vimg_tile_bytes(K) + 255) & ~size_t(255);
}

int launch_mvn_tail(const MvnEncParams& P, int* loss_rows, cudaStream_t st) {
  int64_t g = (P.m.B * P.m.K + kTailThreads * 4 - 1) / (kTailThreads * 4);
  const int64_t cap = (int64_t)num_sms() * 8;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  VIPSY_CUDA_CHECK(cudaMemsetAsync(P.gmax, 0, 4, st));
  mvn_tail_kernel<<<(int)g, kTailThreads, 0, st>>>(P.m, P.gmax);
  VIPSY_CUDA_CHECK(cudaGetLastError());
  *loss_rows = (int)g;
  return VIPSY_OK;
}

int launch_mvn_bwd(const MvnEncParams& P, cudaStream_t st) {
  const int K = P.m.K, H = P.H;
  const int64_t n_tiles = (P.m.B + kMvTP - 1) / kMvTP;
  // the draws by bulk copy when they lie in our own (padded, pre-split) buffer: rows are whole 16-byte pieces
  const int bulk = P.eps_split != nullptr ? 1 : 0;
  if (bulk)       // the last tile's copy runs past the last person's row: finite (zero) padding behind the draws
    VIPSY_CUDA_CHECK(cudaMemsetAsync(P.eps_split + (size_t)P.m.S * P.m.B * K, 0, (size_t)kMvTP * K * 4 + 64, st));
  if (P.d_hidden) {
    const DhSmem L = dh_smem(K, H, bulk != 0);
    VIPSY_REQUIRE(L.total <= 227 * 1024, VIPSY_ERR_UNSUPPORTED, "mvn encoder backward: K=%d H=%d does not fit shared memory", K, H);
    VIPSY_CUDA_CHECK(cudaFuncSetAttribute(mvn_bwd_dh_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total));
    int64_t g = num_sms();
    if (g > n_tiles) g = n_tiles;
    mvn_bwd_dh_kernel<<<(int)g, kDhThreads, L.total, st>>>(P, bulk);
    VIPSY_CUDA_CHECK(cudaGetLastError());
  }
  if (P.d_w22 || P.d_b22 || P.d_w21 || P.d_b21) {
    const DwSmem L = dw_smem(K, H);
    VIPSY_REQUIRE(L.total <= 227 * 1024, VIPSY_ERR_UNSUPPORTED, "mvn encoder backward: K=%d H=%d does not fit shared memory", K, H);
    VIPSY_CUDA_CHECK(cudaFuncSetAttribute(mvn_bwd_dw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total));
    const int64_t n_tiles64 = (P.m.B + kDwTP - 1) / kDwTP;
    {
      const int64_t total = n_tiles64 * kDwTP * (mv_kh(H) / 8);
      int64_t g = (total + 255) / 256;
      if (g > (int64_t)num_sms() * 16) g = (int64_t)num_sms() * 16;
      mvn_himg_kernel<<<(int)g, 256, 0, st>>>(P.hidden, P.m.B, H, mv_kh(H), P.himg);
      VIPSY_CUDA_CHECK(cudaGetLastError());
      int64_t gp = n_tiles64 * P.m.S;
      if (gp > (int64_t)num_sms() * 8) gp = (int64_t)num_sms() * 8;
      mvn_dwprep_kernel<<<(int)gp, 256, 0, st>>>(P, P.m.eps ? P.m.eps : P.eps_out, P.vimg);
      VIPSY_CUDA_CHECK(cudaGetLastError());
    }
    int n_pg, n_qg;
    dw_grid(K, H, &n_pg, &n_qg);
    if (n_qg > n_tiles64) n_qg = (int)n_tiles64;
    mvn_bwd_dw_kernel<<<n_pg * n_qg, kDwThreads, L.total, st>>>(P, P.vimg, n_pg, n_qg);
    VIPSY_CUDA_CHECK(cudaGetLastError());
    const int total = K * (K + 1) / 2 * (H + 1);
    mvn_dw_finalize_kernel<<<(total + 255) / 256, 256, 0, st>>>(P.dw_part, n_qg, mv_chunks(K) * kMvCN, K, H, mv_kh(H),
                                                                P.d_w22, P.d_b22, P.d_w21, P.d_b21);
    VIPSY_CUDA_CHECK(cudaGetLastError());
  }
  return VIPSY_OK;
}

int launch_mvn_prep(const MvnEncParams& P, cudaStream_t st) {
  const int KH = mv_kh(P.H), n_chunks = mv_chunks(P.m.K);
  const int total = n_chunks * kMvCN * (KH / 8);
  mvn_prep_kernel<<<(total + 255) / 256, 256, 0, st>>>(P.w22, P.b22, P.w21, P.b21, P.m.K, P.H, KH, mv_nh(P.H), n_chunks, P.wimg,
                                                       P.wimg_b);
  VIPSY_CUDA_CHECK(cudaGetLastError());
  return VIPSY_OK;
}

int launch_mvn_fwd(const MvnEncParams& P, cudaStream_t st) {
  if (!P.m.eps) {                                      // the step's draws -> eps_split / eps_out
    VIPSY_REQUIRE(P.eps_split || P.eps_out, VIPSY_ERR_ARG, "mvn encoder: no buffer for the draws");
    if (P.m.K > 32) {
      int64_t g = ((int64_t)P.m.S * P.m.B + 7) / 8;
      if (g > (int64_t)num_sms() * 32) g = (int64_t)num_sms() * 32;
      mvn_noise_mid_kernel<<<(int)g, 256, 0, st>>>(P);
    } else {
      const size_t smem = (size_t)kNzThreads * (P.m.K + 1) * 4;
      VIPSY_CUDA_CHECK(cudaFuncSetAttribute(mvn_noise_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      int64_t g = (P.m.B + kNzThreads - 1) / kNzThreads * P.m.S;
      if (g > (int64_t)num_sms() * 4) g = (int64_t)num_sms() * 4;
      mvn_noise_kernel<<<(int)g, kNzThreads, smem, st>>>(P);
    }
    VIPSY_CUDA_CHECK(cudaGetLastError());
  }
  const FwSmem L = fw_smem(P.m.K, P.H);
  VIPSY_REQUIRE(L.nst >= 2, VIPSY_ERR_UNSUPPORTED, "mvn encoder: K=%d H=%d does not fit shared memory", P.m.K, P.H);
  VIPSY_CUDA_CHECK(cudaFuncSetAttribute(mvn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total));
  const int64_t n_tiles = (P.m.B + kMvTP - 1) / kMvTP;
  int64_t g = num_sms();
  if (g > n_tiles) g = n_tiles;
  mvn_fwd_kernel<<<(int)g, kFwThreads, L.total, st>>>(P);
  VIPSY_CUDA_CHECK(cudaGetLastError());
  return VIPSY_OK;
}


static int check_mvn(const VipsyMvnEncoderIrtArgs* a, const char* who) {
  VIPSY_REQUIRE(a != nullptr, VIPSY_ERR_ARG, "%s: args is NULL", who);
  const VipsyIrtArgs& r = a->irt;
  VIPSY_REQUIRE(r.batch >= 0 && r.n_dims >= 2 && r.n_dims <= 128 && r.n_particles > 0, VIPSY_ERR_ARG,
                "%s: bad sizes B=%lld K=%d S=%d (K in 2..128)", who, (long long)r.batch, r.n_dims, r.n_particles);
  VIPSY_REQUIRE(a->n_hidden >= 1 && a->n_hidden <= kMvMaxH, VIPSY_ERR_UNSUPPORTED,
                "%s: hidden_dim %d outside 1..%d", who, a->n_hidden, kMvMaxH);
  VIPSY_REQUIRE(a->hidden && a->w22 && a->b22 && (r.mu || a->w21), VIPSY_ERR_ARG,
                "%s: hidden, w22, b22 and irt.mu (or w21) are required", who);
  VIPSY_REQUIRE(!a->w21 || a->b21, VIPSY_ERR_ARG, "%s: w21 needs b21", who);
  VIPSY_REQUIRE(a->w21 || (!a->d_w21 && !a->d_b21), VIPSY_ERR_ARG, "%s: d_w21 / d_b21 need w21", who);
  VIPSY_REQUIRE((reinterpret_cast<uintptr_t>(a->workspace) & 127) == 0, VIPSY_ERR_ARG, "%s: workspace must be 128-byte aligned", who);
  return VIPSY_OK;
}

static MvnEncParams mvn_params(const VipsyMvnEncoderIrtArgs* a) {
  MvnEncParams P{};
  const VipsyIrtArgs& r = a->irt;
  P.m.B = r.batch; P.m.K = r.n_dims; P.m.S = r.n_particles; P.m.J = r.n_items;
  P.m.mu = r.mu; P.m.eps = r.eps; P.m.seed = r.seed; P.m.offset = r.offset; P.m.person_offset = r.person_offset;
  P.m.scale = r.scale;
  P.hidden = a->hidden; P.w22 = a->w22; P.b22 = a->b22; P.H = a->n_hidden;
  P.w21 = a->w21; P.b21 = a->b21;
  P.d_hidden = a->d_hidden; P.d_w22 = a->d_w22; P.d_b22 = a->d_b22; P.d_w21 = a->d_w21; P.d_b21 = a->d_b21;
  return P;
}

}  // namespace vipsy

using namespace vipsy;

extern "C" {

// workspace: [ mirt_layout of irt | fc22 operand images | the draws | partial d [W22 | b22] ]
struct MvnLayout { size_t mirt, wimg, wimgb, himg, vimg, eps, dwpart, gmax, total; };
static MvnLayout mvn_layout(const VipsyMvnEncoderIrtArgs* a) {
  const VipsyIrtArgs& r = a->irt;
  MvnLayout L;
  size_t o = 0;
  L.mirt = o; o += (mirt_workspace_bytes(&r) + 255) & ~size_t(255);
  L.wimg = o; o += mvn_wimg_bytes(r.n_dims, a->n_hidden);
  L.wimgb = o; o += mvn_wimgb_bytes(r.n_dims, a->n_hidden);
  L.himg = o; o += mvn_himg_bytes(r.batch, a->n_hidden);
  L.vimg = o; o += mvn_vimg_bytes(r.batch, r.n_dims, r.n_particles);
  L.eps = o; o += r.eps ? 0 : mvn_eps_bytes(r.batch, r.n_dims, r.n_particles);
  L.dwpart = o; o += mvn_dwpart_bytes(r.n_dims, a->n_hidden);
  L.gmax = o; o += 256;
  L.total = o;
  return L;
}

size_t vipsy_irt_mvn_encoder_workspace_bytes(const VipsyMvnEncoderIrtArgs* a) {
  if (!a || a->irt.batch < 0 || a->irt.n_dims < 2 || a->irt.n_dims > 128 || a->irt.n_items <= 0 || a->irt.n_particles <= 0 ||
      a->n_hidden < 1 || a->n_hidden > kMvMaxH)
    return 0;
  VipsyMvnEncoderIrtArgs t = *a;
  t.irt.cov_mode = VIPSY_COV_PERSON;
  return mvn_layout(&t).total;
}

int vipsy_irt_mvn_encoder_elbo_grad(const VipsyMvnEncoderIrtArgs* a, void* stream) {
  int rc = check_mvn(a, "vipsy_irt_mvn_encoder_elbo_grad");
  if (rc != VIPSY_OK) return rc;
  VipsyIrtArgs r = a->irt;
  r.cov_mode = VIPSY_COV_PERSON; r.chol = nullptr; r.d_chol = nullptr; r.sigma = nullptr; r.d_sigma = nullptr;
  r.chol_layout = VIPSY_CHOL_PACKED; r.chol_raw = 1;
  if (!r.mu && a->w21) r.mu = a->hidden;                 // fc21 fused: the location is computed here, irt.mu is never read
  rc = check_irt(&r);
  if (rc != VIPSY_OK) return rc;
  VIPSY_REQUIRE(r.model != VIPSY_IRT_1PL, VIPSY_ERR_ARG, "vipsy_irt_mvn_encoder_elbo_grad: irt_1pl is defined for K == 1 only");
  VIPSY_REQUIRE(r.estimator == VIPSY_EST_PATHWISE, VIPSY_ERR_UNSUPPORTED,
                "vipsy_irt_mvn_encoder_elbo_grad: the reference runs MultivariateNormal guides pathwise (SURVEY Finding 1)");
  VIPSY_REQUIRE((r.d_mu != nullptr) || (!a->d_hidden && !a->d_w22 && !a->d_b22), VIPSY_ERR_ARG,
                "vipsy_irt_mvn_encoder_elbo_grad: encoder gradients need irt.d_mu as well");
  cudaStream_t st = (cudaStream_t)stream;
  if (r.batch == 0) {
    VIPSY_CUDA_CHECK(cudaMemsetAsync(r.loss, 0, 8, st));
    return VIPSY_OK;
  }
  VipsyMvnEncoderIrtArgs t = *a;
  t.irt = r;
  const MvnLayout L = mvn_layout(&t);
  VIPSY_REQUIRE(a->workspace && a->workspace_bytes >= L.total, VIPSY_ERR_WORKSPACE,
                "vipsy_irt_mvn_encoder_elbo_grad: workspace %zu B < required %zu B", a->workspace_bytes, L.total);
  unsigned char* ws = reinterpret_cast<unsigned char*>(a->workspace);
  r.workspace = ws + L.mirt;
  r.workspace_bytes = L.wimg - L.mirt;
  MvnEncParams P = mvn_params(a);
  P.wimg = ws + L.wimg;
  P.wimg_b = ws + L.wimgb;
  P.himg = ws + L.himg;
  P.vimg = ws + L.vimg;
  // our own draws: pre-split fp16 pair words when a row is a whole number of 16-byte pieces (bulk copies), else fp32
  P.eps_out = (r.eps || (r.n_dims & 3) == 0) ? nullptr : reinterpret_cast<float*>(ws + L.eps);
  P.eps_split = (r.eps || (r.n_dims & 3) != 0) ? nullptr : reinterpret_cast<uint32_t*>(ws + L.eps);
  P.dw_part = reinterpret_cast<float*>(ws + L.dwpart);
  P.gmax = reinterpret_cast<float*>(ws + L.gmax);
  P.logdet_out = nullptr;
  return mirt_launch(&r, st, &P);
}

int vipsy_mvn_encoder_sample(const VipsyMvnEncoderIrtArgs* a, float* x, float* eps_out, float* diag, float* logdet,
                             void* stream) {
  int rc = check_mvn(a, "vipsy_mvn_encoder_sample");
  if (rc != VIPSY_OK) return rc;
  if (a->irt.batch == 0) return VIPSY_OK;
  const size_t need = mvn_wimg_bytes(a->irt.n_dims, a->n_hidden);
  VIPSY_REQUIRE(a->workspace && a->workspace_bytes >= need, VIPSY_ERR_WORKSPACE,
                "vipsy_mvn_encoder_sample: workspace %zu B < required %zu B", a->workspace_bytes, need);
  MvnEncParams P = mvn_params(a);
  P.wimg = reinterpret_cast<unsigned char*>(a->workspace);
  P.wimg_b = nullptr;
  P.m.x = x; P.eps_out = eps_out; P.m.diagL = diag; P.logdet_out = logdet; P.m.aux = nullptr;
  cudaStream_t st = (cudaStream_t)stream;
  VIPSY_REQUIRE(a->irt.eps || eps_out, VIPSY_ERR_ARG,
                "vipsy_mvn_encoder_sample: eps_out may be NULL only with injected noise (args->irt.eps)");
  if (a->irt.eps && eps_out)
    VIPSY_CUDA_CHECK(cudaMemcpyAsync(eps_out, a->irt.eps, (size_t)a->irt.n_particles * a->irt.batch * a->irt.n_dims * 4,
                                     cudaMemcpyDeviceToDevice, st));
  rc = launch_mvn_prep(P, st);
  if (rc != VIPSY_OK) return rc;
  return launch_mvn_fwd(P, st);
}

}  // extern "C"
