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
constexpr int kFwEpiWarps = 4;                 // warps 0-3: epilogue (TMEM lane quadrant = warp)
constexpr int kFwThreads = (kFwEpiWarps + 2) * 32;   // + warp 4: bulk-copy producer, warp 5: MMA issuer
constexpr int kFwTmemBufs = 4;                 // 4 x 128 accumulator columns

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

__global__ void mvn_prep_kernel(const float* __restrict__ w22, const float* __restrict__ b22, int K, int H, int KH,
                                int n_chunks, unsigned char* __restrict__ img) {
  const int n_oct = KH / 8;
  const int total = n_chunks * kMvCN * n_oct;
  const uint32_t RS = (uint32_t)n_oct * kCSb;
  const size_t tile_bytes = (size_t)kMvCN * KH * 2;
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
    }
    uint4 hi, lo;
    split8(v, hi, lo);
    const int c = pc / kMvCN, r = pc - c * kMvCN;
    unsigned char* base = img + (size_t)c * 2 * tile_bytes + (size_t)(r >> 3) * RS + (size_t)o * kCSb + (r & 7) * 16;
    *reinterpret_cast<uint4*>(base) = hi;
    *reinterpret_cast<uint4*>(base + tile_bytes) = lo;
  }
}

// ------------------------------------------------------------------------------------------------ forward
struct FwSmem {
  uint32_t offW, offHh, offHl, offEps, total;
  uint32_t chunk_bytes, tile_h_bytes;
  int nst, ES;
};
__host__ __device__ inline FwSmem fw_smem(int K, int H) {
  FwSmem s;
  const int KH = mv_kh(H);
  s.chunk_bytes = (uint32_t)mv_chunk_bytes(H);
  s.tile_h_bytes = (uint32_t)kMvTP * KH * 2;
  s.ES = (K + 7) / 8 * 8 + 4;
  const uint32_t fixed = 2 * s.tile_h_bytes + (uint32_t)kMvTP * s.ES * 4;
  const uint32_t cap = 227 * 1024 - 1024;
  int nst = (int)((cap - fixed) / s.chunk_bytes);
  if (nst > 4) nst = 4;
  s.nst = nst;
  uint32_t o = 0;
  s.offW = o; o += (uint32_t)(nst > 0 ? nst : 0) * s.chunk_bytes;
  s.offHh = o; o += s.tile_h_bytes;
  s.offHl = o; o += s.tile_h_bytes;
  s.offEps = o; o += (uint32_t)kMvTP * s.ES * 4;
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
  unsigned char* Hh = smem + L.offHh;
  unsigned char* Hl = smem + L.offHl;
  float* eps_s = reinterpret_cast<float*>(smem + L.offEps);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t n_tiles = (m.B + kMvTP - 1) / kMvTP;

  if (tid == 0) {
    for (int i = 0; i < 4; ++i) { umma::mbar_init(&bar_wfull[i], 1); umma::mbar_init(&bar_wempty[i], 1); }
    for (int i = 0; i < kFwTmemBufs; ++i) { umma::mbar_init(&bar_tfull[i], 1); umma::mbar_init(&bar_tempty[i], kFwEpiWarps * 32); }
    umma::mbar_init(&bar_hfull, kFwEpiWarps * 32);
    umma::fence_mbar_init();
  }
  if (warp == kFwEpiWarps + 1) umma::tmem_alloc(&tmem_base_sm, 512);
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem = tmem_base_sm;

  if (warp == kFwEpiWarps) {
    // =========================== bulk-copy producer: the operand image, chunk by chunk ========================
    if (lane == 0) {
      uint32_t it = 0;
      for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x)
        for (int s = 0; s < S; ++s)
          for (int c = 0; c < n_chunks; ++c, ++it) {
            const uint32_t st = it % NST, round = it / NST;
            umma::mbar_wait(&bar_wempty[st], (round & 1) ^ 1);
            umma::bulk_load(Wst + (size_t)st * L.chunk_bytes, P.wimg + (size_t)c * L.chunk_bytes, L.chunk_bytes,
                            &bar_wfull[st]);
          }
    }
  } else if (warp == kFwEpiWarps + 1) {
    // =========================== MMA issuer ====================================================================
    if (lane == 0) {
      const uint32_t idesc = umma::make_idesc_f16(kMvTP, kMvCN, umma::kMajorK, umma::kMajorK);
      const uint32_t sHh = umma::smem_u32(Hh), sHl = umma::smem_u32(Hl);
      uint32_t it = 0, ntile = 0;
      for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++ntile) {
        umma::mbar_wait(&bar_hfull, ntile & 1);
        umma::fence_after_sync();
        for (int s = 0; s < S; ++s)
          for (int c = 0; c < n_chunks; ++c, ++it) {
            const uint32_t st = it % NST, round = it / NST;
            const uint32_t buf = it % kFwTmemBufs, tround = it / kFwTmemBufs;
            umma::mbar_wait(&bar_wfull[st], round & 1);
            umma::mbar_wait(&bar_tempty[buf], (tround & 1) ^ 1);
            umma::fence_after_sync();
            const uint32_t sWh = umma::smem_u32(Wst + (size_t)st * L.chunk_bytes), sWl = sWh + L.chunk_bytes / 2;
            uint32_t acc = 0;
            for (int ks = 0; ks < KH / 16; ++ks) {
#pragma unroll
              for (int term = 0; term < 3; ++term) {            // hi*hi, hi*lo, lo*hi
                const uint64_t da = umma::make_desc((term == 2 ? sHl : sHh) + ks * 2 * kCSb, kCSb, RS);
                const uint64_t db = umma::make_desc((term == 1 ? sWl : sWh) + ks * 2 * kCSb, kCSb, RS);
                umma::mma_f16_ss(tmem + buf * kMvCN, da, db, idesc, acc);
                acc = 1;
              }
            }
            umma::commit(&bar_wempty[st]);
            umma::commit(&bar_tfull[buf]);
          }
      }
    }
  } else {
    // =========================== epilogue: thread = person = TMEM lane =========================================
    const int t = tid;                                   // 0..127
    float* er = eps_s + (size_t)t * ES;
    const uint32_t tlane = tmem + ((uint32_t)(warp * 32) << 16);
    const float inv_scale = 1.0f / kMvWScale;
    uint32_t it = 0;
    for (int k = K; k < ES; ++k) er[k] = 0.f;            // padding columns of a row's last group: finite, weight 0
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const int64_t b = tile * kMvTP + t;
      const bool live = b < m.B;
      // ---- [hidden | 1 | 0] row of this person -> fp16 hi / lo operand tiles (K-major) -------------------------
      {
        const float* hr = P.hidden + (live ? b : 0) * H;
        const bool vec = (H & 3) == 0;
        for (int o = 0; o < KH / 8; ++o) {
          float v[8];
          if (vec && live && o * 8 + 8 <= H) {
            const float4 a0 = *reinterpret_cast<const float4*>(hr + o * 8);
            const float4 a1 = *reinterpret_cast<const float4*>(hr + o * 8 + 4);
            v[0] = a0.x; v[1] = a0.y; v[2] = a0.z; v[3] = a0.w; v[4] = a1.x; v[5] = a1.y; v[6] = a1.z; v[7] = a1.w;
          } else {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const int k = o * 8 + q;
              v[q] = !live ? 0.f : (k < H ? hr[k] : (k == H ? 1.f : 0.f));
            }
          }
          uint4 hi, lo;
          split8(v, hi, lo);
          const uint32_t off = (uint32_t)(t >> 3) * RS + (uint32_t)o * kCSb + (uint32_t)(t & 7) * 16u;
          *reinterpret_cast<uint4*>(Hh + off) = hi;
          *reinterpret_cast<uint4*>(Hl + off) = lo;
        }
        umma::fence_proxy_async();
        umma::mbar_arrive(&bar_hfull);
      }
      float logdet = 0.f, auxp = 0.f;
      for (int s = 0; s < S; ++s) {
        const float ss = fill_noise_row(m, s, live ? b : 0, er, inv_scale);
        if (P.eps_out) {                      // the warp's 32 rows leave through coalesced stores
          __syncwarp();
          const int64_t b0 = tile * kMvTP + warp * 32;
          const float* e0 = eps_s + (size_t)(warp * 32) * ES;
          for (int pp = 0; pp < 32 && b0 + pp < m.B; ++pp) {
            float* dst = P.eps_out + ((int64_t)s * m.B + b0 + pp) * K;
            for (int k = lane; k < K; k += 32) dst[k] = e0[(size_t)pp * ES + k] * kMvWScale;
          }
          __syncwarp();
        }
        auxp = fmaf(0.5f, ss, auxp);
        const float* mur = m.mu + (live ? b : 0) * K;
        float* xr = m.x ? m.x + ((int64_t)s * m.B + (live ? b : 0)) * K : nullptr;
        int I = 0, Jc = 0;
        float acc[8], mu_r[8];                                   // mu of the current block row, fetched a block row ahead
#pragma unroll
        for (int r = 0; r < 8; ++r) { acc[r] = 0.f; mu_r[r] = r < K ? mur[r] : 0.f; }
        for (int c = 0; c < n_chunks; ++c, ++it) {
          const uint32_t buf = it % kFwTmemBufs, tround = it / kFwTmemBufs;
          umma::mbar_wait(&bar_tfull[buf], tround & 1);
          umma::fence_after_sync();
#pragma unroll 1
          for (int u = 0; u < kMvCN / 64; ++u) {
            if (I < NB) {                                        // warp-uniform
              float v[64];
              umma::tmem_ld64(tlane + buf * kMvCN + u * 64, v);
              const float4 e0 = *reinterpret_cast<const float4*>(er + 8 * Jc);
              const float4 e1 = *reinterpret_cast<const float4*>(er + 8 * Jc + 4);
#pragma unroll
              for (int r = 0; r < 8; ++r) {
                float d0 = v[8 * r] * e0.x, d1 = v[8 * r + 4] * e1.x;
                d0 = fmaf(v[8 * r + 1], e0.y, d0); d1 = fmaf(v[8 * r + 5], e1.y, d1);
                d0 = fmaf(v[8 * r + 2], e0.z, d0); d1 = fmaf(v[8 * r + 6], e1.z, d1);
                d0 = fmaf(v[8 * r + 3], e0.w, d0); d1 = fmaf(v[8 * r + 7], e1.w, d1);
                acc[r] += d0 + d1;
              }
              if (Jc == I) {                                      // diagonal block: rows 8I .. 8I+7 are complete
                const float ev[8] = {e0.x, e0.y, e0.z, e0.w, e1.x, e1.y, e1.z, e1.w};
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                  const int i = 8 * I + r;
                  if (i < K) {
                    const float dv = v[9 * r];                   // 64 raw_ii
                    const float dr = dv * inv_scale;
                    const float lii = __expf(dr);
                    const float xv = mu_r[r] + fmaf(lii * kMvWScale - dv, ev[r], acc[r]);  // exp(raw_ii) eps_i for raw_ii eps_i
                    if (live) {
                      if (xr) xr[i] = xv;
                      if (s == 0 && m.diagL) m.diagL[b * K + i] = lii;
                    }
                    auxp = fmaf(-0.5f * xv, xv, auxp);
                    if (s == 0) logdet += dr;
                  }
                  acc[r] = 0.f;
                  mu_r[r] = i + 8 < K ? mur[i + 8] : 0.f;
                }
                ++I; Jc = 0;
              } else {
                ++Jc;
              }
            }
          }
          umma::fence_before_sync();
          umma::mbar_arrive(&bar_tempty[buf]);
        }
      }
      if (live && m.aux) *reinterpret_cast<float4*>(m.aux + 4 * b) = make_float4(auxp + (float)S * logdet, 0.f, 0.f, 0.f);
      if (live && P.logdet_out) P.logdet_out[b] = logdet;
    }
  }
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  if (warp == kFwEpiWarps + 1) umma::tmem_dealloc(tmem, 512);
}

}  // namespace

int launch_mvn_prep(const MvnEncParams& P, cudaStream_t st) {
  const int KH = mv_kh(P.H), n_chunks = mv_chunks(P.m.K);
  const int total = n_chunks * kMvCN * (KH / 8);
  mvn_prep_kernel<<<(total + 255) / 256, 256, 0, st>>>(P.w22, P.b22, P.m.K, P.H, KH, n_chunks, P.wimg);
  VIPSY_CUDA_CHECK(cudaGetLastError());
  return VIPSY_OK;
}

int launch_mvn_fwd(const MvnEncParams& P, cudaStream_t st) {
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
  VIPSY_REQUIRE(a->hidden && a->w22 && a->b22 && r.mu, VIPSY_ERR_ARG, "%s: hidden, w22, b22 and irt.mu are required", who);
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
  P.d_hidden = a->d_hidden; P.d_w22 = a->d_w22; P.d_b22 = a->d_b22;
  return P;
}

size_t mvn_wimg_bytes(int K, int H) { return ((size_t)mv_chunks(K) * mv_chunk_bytes(H) + 255) & ~size_t(255); }

}  // namespace vipsy

using namespace vipsy;

extern "C" {

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
  P.m.x = x; P.eps_out = eps_out; P.m.diagL = diag; P.logdet_out = logdet; P.m.aux = nullptr;
  cudaStream_t st = (cudaStream_t)stream;
  rc = launch_mvn_prep(P, st);
  if (rc != VIPSY_OK) return rc;
  return launch_mvn_fwd(P, st);
}

}  // extern "C"
