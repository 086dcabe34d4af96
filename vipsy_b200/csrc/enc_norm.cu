// NormEncoder of the amortised K = 1 guide (vi.py:417-435, 677-684) on the stored response codes.
//
//   x      = code -> {0, 1, -1 (missing)}                       (vi.py:679-682)
//   pre    = fc1(x)            [B x H]   H <= 64
//   hidden = softplus(pre)     (torch: beta 1, threshold 20)
//   mu     = fc21(hidden),  sigma = exp(fc22(hidden))
//
// Two register-tiled SIMT products (B x J x H is 6.4 GFMA at c2, fp32 exact inputs; x never exists as a float
// matrix in HBM and the hidden activations are not stored, only `pre`):
//   forward : tile of 128 persons, thread = 8 persons (stride 16) x 4 hidden units, reduction over items from shared memory
//   backward: d fc1.weight[h][j] = sum_b dpre[b][h] * x[b][j], thread = 4 hidden units x 8 items held in
//             registers across all the tiles of a persistent CTA; dpre is rebuilt per tile from `pre`.
// Cross-CTA sums: one partial row per CTA, folded by finalize_kernel (fixed order, fp64) -> deterministic.
#include "enc_norm.cuh"

namespace vipsy {

constexpr int kEncTB = 128;             // persons per tile
constexpr int kEncH = 64;               // hidden slots (weights zero-padded up to it)
constexpr int kEncJC = 128;             // item columns per shared-memory chunk
constexpr int kEncLd = kEncJC + 4;      // xs row pitch (floats): 16-byte aligned rows
constexpr int kEncThreads = 256;

__device__ __forceinline__ float code_x(uint32_t c) { return c == 255u ? -1.0f : (c ? 1.0f : 0.0f); }

// codes of persons [tile0, tile0 + 128) x items [j0, j0 + 128): 16 words per thread, loaded into registers one
// tile ahead (the loads are in flight while the current tile is multiplied), then expanded to floats in xs[p][j]
// (0 outside B x J)
constexpr int kEncWords = kEncTB * (kEncJC / 4) / kEncThreads;      // 16
__device__ __forceinline__ void load_codes(uint32_t (&wd)[kEncWords], const EncParams& P, int64_t tile0, int j0,
                                           bool aligned) {
  const int w = threadIdx.x & 31, p0 = threadIdx.x >> 5;            // word w of persons p0 + 8 q
  const int j = j0 + 4 * w;
  if (aligned && P.idx == nullptr && tile0 + kEncTB <= P.B && (j + 3 < P.J || j >= P.J)) {
    // interior tile, whole (or absent) word: 16 independent loads, no per-word control flow
    const uint32_t* g = reinterpret_cast<const uint32_t*>(P.y + (tile0 + p0) * P.ld + j);
    const int64_t step = 2 * P.ld;                                   // 8 rows, in words
    const bool live = j < P.J;
#pragma unroll
    for (int q = 0; q < kEncWords; ++q) wd[q] = live ? __ldg(g + q * step) : 0u;
    return;
  }
#pragma unroll
  for (int q = 0; q < kEncWords; ++q) {
    const int64_t r = tile0 + p0 + 8 * q;
    uint32_t word = 0;
    if (r < P.B && j < P.J) {
      const uint8_t* g = P.y + (P.idx ? __ldg(P.idx + r) : r) * P.ld + j;
      if (aligned && j + 3 < P.J) {
        word = __ldg(reinterpret_cast<const uint32_t*>(g));
      } else {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (j + k < P.J) word |= (uint32_t)__ldg(g + k) << (8 * k);
      }
    }
    wd[q] = word;
  }
}
__device__ __forceinline__ void store_x(float* xs, const uint32_t (&wd)[kEncWords]) {
#pragma unroll
  for (int q = 0; q < kEncWords; ++q) {
    const int e = threadIdx.x + q * kEncThreads;
    const int p = e >> 5, w = e & 31;
    const uint32_t word = wd[q];
    *reinterpret_cast<float4*>(xs + p * kEncLd + 4 * w) =
        make_float4(code_x(word & 255u), code_x((word >> 8) & 255u), code_x((word >> 16) & 255u), code_x(word >> 24));
  }
}

// ------------------------------------------------------------------------------------------------ forward
// The products run as packed FFMA2 (two hidden units / two items per instruction): the FMA pipe stays the
// limiter and the shared-memory loads and broadcasts ride in the issue slots that frees.
__global__ void __launch_bounds__(kEncThreads, 2) enc_norm_fwd_kernel(const EncParams P) {
  extern __shared__ __align__(16) float sm[];
  float* xs = sm;                                   // [128][132]
  float* ws = xs + kEncTB * kEncLd;                 // [128][64]: ws[j][h] = w1[h][j0 + j]
  float* hv = ws + kEncJC * kEncH;                  // b1 | w21 | w22, 64 each
  const int t = threadIdx.x, hq = t & 15, pg = t >> 4;
  const bool aligned = ((reinterpret_cast<uintptr_t>(P.y) | (uintptr_t)P.ld) & 3) == 0;
  const int n_chunks = (P.J + kEncJC - 1) / kEncJC;
  const bool one_chunk = n_chunks == 1;
  if (t < 3 * kEncH) {
    const int h = t & 63, which = t >> 6;
    const float* src = which == 0 ? P.b1 : which == 1 ? P.w21 : P.w22;
    hv[t] = h < P.H ? src[h] : 0.0f;
  }
  auto stage_w = [&](int j0) {
    for (int e = t; e < kEncJC * kEncH; e += kEncThreads) {
      const int j = e >> 6, h = e & 63;             // conflict-free shared stores; w1 is L2-resident
      ws[j * kEncH + h] = (h < P.H && j0 + j < P.J) ? __ldg(P.w1 + (size_t)h * P.J + j0 + j) : 0.0f;
    }
  };
  if (one_chunk) stage_w(0);
  const float b21 = __ldg(P.b21), b22 = __ldg(P.b22);
  const int64_t n_tiles = (P.B + kEncTB - 1) / kEncTB;
  uint32_t wd[kEncWords];
  if ((int64_t)blockIdx.x < n_tiles) load_codes(wd, P, (int64_t)blockIdx.x * kEncTB, 0, aligned);
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t tile0 = tile * kEncTB;
    float2 acc[8][2];                               // person pg + 16 i, hidden units (4 hq, +1), (+2, +3)
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i][0] = acc[i][1] = make_float2(0.0f, 0.0f);
    for (int c = 0; c < n_chunks; ++c) {
      const int j0 = c * kEncJC;
      __syncthreads();                              // previous readers of xs / ws are done
      if (!one_chunk) stage_w(j0);
      store_x(xs, wd);
      __syncthreads();
      {                                             // next (tile, chunk): loads in flight during the product
        int nc = c + 1;
        int64_t nt = tile;
        if (nc == n_chunks) { nc = 0; nt += gridDim.x; }
        if (nt < n_tiles) load_codes(wd, P, nt * kEncTB, nc * kEncJC, aligned);
      }
      const int jn = min(kEncJC, (P.J - j0 + 3) & ~3);
      const float* xrow = xs + pg * kEncLd;        // persons pg + 16 i: a warp's two rows sit 4 banks apart
#pragma unroll 2
      for (int j = 0; j < jn; j += 4) {
        float2 w[4][2];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float4 v = *reinterpret_cast<const float4*>(ws + (j + k) * kEncH + hq * 4);
          w[k][0] = make_float2(v.x, v.y); w[k][1] = make_float2(v.z, v.w);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 x = *reinterpret_cast<const float4*>(xrow + i * 16 * kEncLd + j);
          const float xv[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float2 xx = make_float2(xv[k], xv[k]);
            acc[i][0] = __ffma2_rn(xx, w[k][0], acc[i][0]);
            acc[i][1] = __ffma2_rn(xx, w[k][1], acc[i][1]);
          }
        }
      }
    }
    // epilogue: bias, save pre, softplus, the two 64 -> 1 heads (16 threads of a half-warp share a person)
    const float4 bb = *reinterpret_cast<const float4*>(hv + hq * 4);
    const float4 v21 = *reinterpret_cast<const float4*>(hv + kEncH + hq * 4);
    const float4 v22 = *reinterpret_cast<const float4*>(hv + 2 * kEncH + hq * 4);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int64_t r = tile0 + pg + 16 * i;
      const float p0 = acc[i][0].x + bb.x, p1 = acc[i][0].y + bb.y, p2 = acc[i][1].x + bb.z, p3 = acc[i][1].y + bb.w;
      if (r < P.B) {
        float* dst = P.pre + (size_t)r * P.H + hq * 4;
        if ((P.H & 3) == 0) {
          if (hq * 4 < P.H) *reinterpret_cast<float4*>(dst) = make_float4(p0, p1, p2, p3);
        } else {
          if (hq * 4 + 0 < P.H) dst[0] = p0;
          if (hq * 4 + 1 < P.H) dst[1] = p1;
          if (hq * 4 + 2 < P.H) dst[2] = p2;
          if (hq * 4 + 3 < P.H) dst[3] = p3;
        }
      }
      float s0, s1, s2, s3, g;
      softplus_parts(p0, s0, g); softplus_parts(p1, s1, g); softplus_parts(p2, s2, g); softplus_parts(p3, s3, g);
      float m = fmaf(s3, v21.w, fmaf(s2, v21.z, fmaf(s1, v21.y, s0 * v21.x)));     // padded slots have weight 0
      float l = fmaf(s3, v22.w, fmaf(s2, v22.z, fmaf(s1, v22.y, s0 * v22.x)));
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) {
        m += __shfl_xor_sync(0xffffffffu, m, o);
        l += __shfl_xor_sync(0xffffffffu, l, o);
      }
      if (hq == 0 && r < P.B) {
        P.mu[r] = m + b21;
        P.sigma[r] = expf(l + b22);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ backward
// grid (x, y): x = persistent CTA over person tiles, y = block of 128 item columns of d fc1.weight
__global__ void __launch_bounds__(kEncThreads, 2) enc_norm_bwd_kernel(const EncParams P) {
  extern __shared__ __align__(16) float sm[];
  float* xs = sm;                                   // [128][132]
  float* dps = xs + kEncTB * kEncLd;                // [128][64]  d loss / d pre
  float* gm = dps + kEncTB * kEncH;                 // [128] d loss / d mu
  float* gl = gm + kEncTB;                          // [128] d loss / d log sigma
  float* red = gl + kEncTB;                         // [3][256] end-of-kernel reduction
  const int t = threadIdx.x, hq = t & 15, jg = t >> 4;
  const int h = t & 63, pq = t >> 6;                // role while dpre is rebuilt: hidden unit h, persons pq + 4 i
  const bool aligned = ((reinterpret_cast<uintptr_t>(P.y) | (uintptr_t)P.ld) & 3) == 0;
  const int j0 = blockIdx.y * kEncJC;
  const bool heads = blockIdx.y == 0;               // the small tensors are produced by the first column block
  const bool cols_live = j0 + jg * 8 < P.J;         // column groups beyond J have nothing to accumulate
  const float w21 = h < P.H ? __ldg(P.w21 + h) : 0.0f, w22 = h < P.H ? __ldg(P.w22 + h) : 0.0f;
  float2 acc[4][4];                                 // hidden unit 4 hq + a, item pairs (2 c, 2 c + 1) of 8 jg + ..
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[a][c] = make_float2(0.0f, 0.0f);
  float s_b1 = 0.0f, s_w21 = 0.0f, s_w22 = 0.0f, s_gm = 0.0f, s_gl = 0.0f;
  const int64_t n_tiles = (P.B + kEncTB - 1) / kEncTB;
  uint32_t wd[kEncWords];
  float n_dm = 0.0f, n_ds = 0.0f, n_sg = 0.0f;      // next tile's d_mu, d_sigma, sigma of person t
  auto load_heads = [&](int64_t tile0) {
    const int64_t r = tile0 + t;
    n_dm = n_ds = n_sg = 0.0f;
    if (t < kEncTB && r < P.B) { n_dm = __ldg(P.d_mu + r); n_ds = __ldg(P.d_sigma + r); n_sg = __ldg(P.sigma + r); }
  };
  if ((int64_t)blockIdx.x < n_tiles) {
    load_codes(wd, P, (int64_t)blockIdx.x * kEncTB, j0, aligned);
    load_heads((int64_t)blockIdx.x * kEncTB);
  }
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t tile0 = tile * kEncTB;
    __syncthreads();
    store_x(xs, wd);
    if (t < kEncTB) {
      const float a = n_dm * P.gscale, b = n_ds * n_sg * P.gscale;       // exp chain rule
      gm[t] = a; gl[t] = b;
      s_gm += a; s_gl += b;
    }
    __syncthreads();
    {
      float pv[kEncTB / 4];
#pragma unroll
      for (int i = 0; i < kEncTB / 4; ++i) {
        const int64_t r = tile0 + pq + 4 * i;
        pv[i] = (r < P.B && h < P.H) ? __ldg(P.pre + (size_t)r * P.H + h) : 0.0f;
      }
      asm volatile("" ::: "memory");                // all 32 loads are issued before the first use
#pragma unroll
      for (int i = 0; i < kEncTB / 4; ++i) {
        const int p = pq + 4 * i;
        float sp, sg;
        softplus_parts(pv[i], sp, sg);
        const float a = gm[p], b = gl[p];           // 0 for persons beyond B
        const float dp = fmaf(a, w21, b * w22) * sg;        // w21 = w22 = 0 for h >= H
        s_b1 += dp; s_w21 = fmaf(a, sp, s_w21); s_w22 = fmaf(b, sp, s_w22);
        dps[p * kEncH + h] = dp;
      }
    }
    __syncthreads();
    if (tile + gridDim.x < n_tiles) {               // next tile's inputs: in flight during the product
      load_codes(wd, P, (tile + gridDim.x) * kEncTB, j0, aligned);
      load_heads((tile + gridDim.x) * kEncTB);
    }
    if (cols_live) {
      const float* xcol = xs + jg * 8;
#pragma unroll 4
      for (int p = 0; p < kEncTB; ++p) {
        const float4 d = *reinterpret_cast<const float4*>(dps + p * kEncH + hq * 4);
        const float4 x0 = *reinterpret_cast<const float4*>(xcol + p * kEncLd);
        const float4 x1 = *reinterpret_cast<const float4*>(xcol + p * kEncLd + 4);
        const float2 xp[4] = {make_float2(x0.x, x0.y), make_float2(x0.z, x0.w), make_float2(x1.x, x1.y),
                              make_float2(x1.z, x1.w)};
        const float dv[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          const float2 dd = make_float2(dv[a], dv[a]);
#pragma unroll
          for (int c = 0; c < 4; ++c) acc[a][c] = __ffma2_rn(dd, xp[c], acc[a][c]);
        }
      }
    }
  }
  // this CTA's partial row: [d_w1 (H x J) | d_b1 (H) | d_w21 (H) | d_b21 | d_w22 (H) | d_b22]
  float* row = P.partial + (size_t)blockIdx.x * P.cols;
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int hh = hq * 4 + a;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int j = j0 + jg * 8 + 2 * c;
      if (hh < P.H && j < P.J) row[(size_t)hh * P.J + j] = acc[a][c].x;
      if (hh < P.H && j + 1 < P.J) row[(size_t)hh * P.J + j + 1] = acc[a][c].y;
    }
  }
  if (heads) {
    __syncthreads();
    red[t] = s_b1; red[256 + t] = s_w21; red[512 + t] = s_w22;
    __syncthreads();
    float* tail = row + (size_t)P.H * P.J;
    if (t < kEncH && t < P.H) {                     // four threads (pq) shared hidden unit t, fixed order
      tail[t] = (red[t] + red[64 + t]) + (red[128 + t] + red[192 + t]);
      tail[P.H + t] = (red[256 + t] + red[320 + t]) + (red[384 + t] + red[448 + t]);
      tail[2 * P.H + 1 + t] = (red[512 + t] + red[576 + t]) + (red[640 + t] + red[704 + t]);
    }
    __syncthreads();
    if (t < kEncTB) { red[t] = s_gm; red[256 + t] = s_gl; }
    __syncthreads();
    if (t < 32) {
      float a = (red[t] + red[32 + t]) + (red[64 + t] + red[96 + t]);
      float b = (red[256 + t] + red[288 + t]) + (red[320 + t] + red[352 + t]);
      a = warp_sum(a); b = warp_sum(b);
      if (t == 0) { tail[2 * P.H] = a; tail[3 * P.H + 1] = b; }
    }
  }
}

static int enc_cols(int H, int J) { return H * J + 3 * H + 2; }
static int enc_grid_x(int64_t B) {
  const int64_t tiles = (B + kEncTB - 1) / kEncTB;
  const int64_t cap = 2 * (int64_t)num_sms();
  return (int)(tiles < cap ? (tiles < 1 ? 1 : tiles) : cap);
}
constexpr size_t kEncFwdSmem = (size_t)(kEncTB * kEncLd + kEncJC * kEncH + 3 * kEncH) * 4;
constexpr size_t kEncBwdSmem = (size_t)(kEncTB * kEncLd + kEncTB * kEncH + 2 * kEncTB + 3 * 256) * 4;

static int check_enc(const VipsyNormEncoderArgs* a, bool bwd) {
  VIPSY_REQUIRE(a != nullptr, VIPSY_ERR_ARG, "vipsy_norm_encoder: args is NULL");
  VIPSY_REQUIRE(a->batch >= 0 && a->n_items > 0 && a->n_hidden > 0, VIPSY_ERR_ARG,
                "vipsy_norm_encoder: bad sizes B=%lld J=%d H=%d", (long long)a->batch, a->n_items, a->n_hidden);
  VIPSY_REQUIRE(a->n_hidden <= kEncH, VIPSY_ERR_UNSUPPORTED, "vipsy_norm_encoder: hidden_dim %d > %d", a->n_hidden,
                kEncH);
  VIPSY_REQUIRE(a->y && a->w1 && a->b1 && a->w21 && a->b21 && a->w22 && a->b22, VIPSY_ERR_ARG,
                "vipsy_norm_encoder: y and the six encoder tensors are required");
  VIPSY_REQUIRE(a->y_row_stride >= a->n_items, VIPSY_ERR_ARG, "vipsy_norm_encoder: y_row_stride < J");
  VIPSY_REQUIRE(a->idx != nullptr || a->batch <= a->n_rows, VIPSY_ERR_ARG, "vipsy_norm_encoder: batch > rows without idx");
  VIPSY_REQUIRE(a->sigma && a->pre, VIPSY_ERR_ARG, "vipsy_norm_encoder: sigma and pre are required");
  if (!bwd) {
    VIPSY_REQUIRE(a->mu, VIPSY_ERR_ARG, "vipsy_norm_encoder_forward: mu is required");
  } else {
    VIPSY_REQUIRE(a->d_mu && a->d_sigma, VIPSY_ERR_ARG, "vipsy_norm_encoder_backward: d_mu and d_sigma are required");
    VIPSY_REQUIRE(a->d_w1 && a->d_b1 && a->d_w21 && a->d_b21 && a->d_w22 && a->d_b22, VIPSY_ERR_ARG,
                  "vipsy_norm_encoder_backward: all six gradient outputs are required");
  }
  return VIPSY_OK;
}

static EncParams enc_params(const VipsyNormEncoderArgs* a) {
  EncParams P;
  P.y = a->y; P.ld = a->y_row_stride; P.idx = a->idx; P.B = a->batch; P.J = a->n_items; P.H = a->n_hidden;
  P.w1 = a->w1; P.b1 = a->b1; P.w21 = a->w21; P.b21 = a->b21; P.w22 = a->w22; P.b22 = a->b22;
  P.mu = a->mu; P.sigma = a->sigma; P.pre = a->pre; P.d_mu = a->d_mu; P.d_sigma = a->d_sigma;
  P.gscale = a->grad_scale; P.partial = nullptr; P.cols = enc_cols(a->n_hidden, a->n_items);
  P.xtiles = (a->x_tiles && !a->idx && a->batch == a->n_rows) ? reinterpret_cast<const unsigned char*>(a->x_tiles) : nullptr;
  return P;
}

}  // namespace vipsy

using namespace vipsy;

extern "C" {

size_t vipsy_norm_encoder_tiles_bytes(int64_t n_rows, int32_t n_items) {
  if (n_rows <= 0 || !enc_tc_resident_supported(1, n_items)) return 0;
  return (size_t)((n_rows + 127) / 128) * enc_tc_tile_bytes(n_items);
}

int vipsy_norm_encoder_pack_tiles(const uint8_t* y, int64_t n_rows, int32_t n_items, int64_t y_row_stride, void* tiles,
                                  void* stream) {
  VIPSY_REQUIRE(y && tiles && n_rows > 0 && y_row_stride >= n_items, VIPSY_ERR_ARG,
                "vipsy_norm_encoder_pack_tiles: bad arguments");
  VIPSY_REQUIRE(enc_tc_resident_supported(1, n_items), VIPSY_ERR_UNSUPPORTED,
                "vipsy_norm_encoder_pack_tiles: J = %d > 255 has no tensor-core path", n_items);
  VIPSY_REQUIRE((reinterpret_cast<uintptr_t>(tiles) & 127) == 0, VIPSY_ERR_ARG,
                "vipsy_norm_encoder_pack_tiles: tiles must be 128-byte aligned");
  return enc_tc_pack_launch(y, n_rows, n_items, y_row_stride, tiles, (cudaStream_t)stream);
}

size_t vipsy_norm_encoder_workspace_bytes(const VipsyNormEncoderArgs* a) {
  if (!a || a->n_items <= 0 || a->n_hidden <= 0 || a->n_hidden > kEncH || a->batch < 0) return 0;
  return (size_t)enc_grid_x(a->batch) * enc_cols(a->n_hidden, a->n_items) * sizeof(float);
}

int vipsy_norm_encoder_forward(const VipsyNormEncoderArgs* a, void* stream) {
  int rc = check_enc(a, false);
  if (rc != VIPSY_OK) return rc;
  if (a->batch == 0) return VIPSY_OK;
  // the attribute is per device: set on every launch (cheap), never cached process-wide
  VIPSY_CUDA_CHECK(cudaFuncSetAttribute(enc_norm_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)kEncFwdSmem));
  const EncParams P = enc_params(a);
  if ((P.xtiles ? enc_tc_resident_supported(P.H, P.J) : enc_tc_supported(P.H, P.J)) && !(a->flags & 1) &&
      (reinterpret_cast<uintptr_t>(a->pre) & 15) == 0)
    return enc_tc_fwd_launch(P, (cudaStream_t)stream);
  enc_norm_fwd_kernel<<<enc_grid_x(a->batch), kEncThreads, kEncFwdSmem, (cudaStream_t)stream>>>(P);
  VIPSY_CUDA_CHECK(cudaGetLastError());
  return VIPSY_OK;
}

int vipsy_norm_encoder_backward(const VipsyNormEncoderArgs* a, void* stream) {
  int rc = check_enc(a, true);
  if (rc != VIPSY_OK) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const int H = a->n_hidden, J = a->n_items;
  if (a->batch == 0) {
    VIPSY_CUDA_CHECK(cudaMemsetAsync(a->d_w1, 0, (size_t)H * J * 4, st));
    VIPSY_CUDA_CHECK(cudaMemsetAsync(a->d_b1, 0, (size_t)H * 4, st));
    VIPSY_CUDA_CHECK(cudaMemsetAsync(a->d_w21, 0, (size_t)H * 4, st));
    VIPSY_CUDA_CHECK(cudaMemsetAsync(a->d_w22, 0, (size_t)H * 4, st));
    VIPSY_CUDA_CHECK(cudaMemsetAsync(a->d_b21, 0, 4, st));
    VIPSY_CUDA_CHECK(cudaMemsetAsync(a->d_b22, 0, 4, st));
    return VIPSY_OK;
  }
  VIPSY_REQUIRE(a->workspace && a->workspace_bytes >= vipsy_norm_encoder_workspace_bytes(a), VIPSY_ERR_ARG,
                "vipsy_norm_encoder_backward: workspace too small (%zu < %zu)", a->workspace_bytes,
                vipsy_norm_encoder_workspace_bytes(a));
  VIPSY_CUDA_CHECK(cudaFuncSetAttribute(enc_norm_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)kEncBwdSmem));
  EncParams P = enc_params(a);
  P.partial = reinterpret_cast<float*>(a->workspace);
  int gx = enc_grid_x(a->batch);
  if ((P.xtiles ? enc_tc_resident_supported(H, J) : enc_tc_supported(H, J)) && !(a->flags & 1) &&
      (reinterpret_cast<uintptr_t>(a->pre) & 15) == 0) {
    if (gx > num_sms()) gx = num_sms();             // one 512-thread CTA per SM: a single wave, one TMEM drain each
    rc = enc_tc_bwd_launch(P, gx, st);
    if (rc != VIPSY_OK) return rc;
  } else {
    dim3 grid(gx, (J + kEncJC - 1) / kEncJC);
    enc_norm_bwd_kernel<<<grid, kEncThreads, kEncBwdSmem, st>>>(P);
    VIPSY_CUDA_CHECK(cudaGetLastError());
  }
  FinalizeSegs segs{};
  segs.n = 6;
  segs.out[0] = a->d_w1;  segs.cols[0] = H * J;
  segs.out[1] = a->d_b1;  segs.cols[1] = H;
  segs.out[2] = a->d_w21; segs.cols[2] = H;
  segs.out[3] = a->d_b21; segs.cols[3] = 1;
  segs.out[4] = a->d_w22; segs.cols[4] = H;
  segs.out[5] = a->d_b22; segs.cols[5] = 1;
  return launch_finalize(P.partial, gx, P.cols, 1.0f, segs, nullptr, 0, nullptr, st);
}

}  // extern "C"
