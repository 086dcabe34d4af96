// K = 2..8 latent dimensions with a per-person Cholesky factor: the whole evaluation in ONE kernel, on
// warp-shuffle FMAs (the north star's K <= 8 branch; vi.py:596-615, 686-693, 706-723).
//
// For small K the three-kernel path (sample -> contraction -> guide gradient, mirt.cu) is dominated by its
// generic 64 x 64 SIMT contraction, padded to 16 latent dimensions, and by the x / g_x round trips through
// HBM.  Here the K = 1 mapping (irt_uni.cu) is generalised instead: one lane = one person, a warp = 32
// persons; the person's factor L, its SC particles x_s = mu + L eps_s and the running d lb / d x_s sit in
// registers; items are walked in lock step, their K + 1 (+2) gradient sums cross the warp with the
// transposing butterfly into warp-private shared accumulators; d mu and d L are formed from registers at the
// end.  Nothing of size B x J, B x S x K or B x S x J touches HBM.  Deterministic: fixed butterflies, one
// partial row per CTA, fixed-order finalisation.
#include "mirt.cuh"

namespace vipsy {

constexpr int kFuWarps = 4;
constexpr int kFuThreads = kFuWarps * 32;

__host__ __device__ inline size_t fu_align16(size_t x) { return (x + 15) & ~size_t(15); }

__host__ __device__ inline int fused_nq(int K, int model) {
  return K + 1 + (model >= VIPSY_IRT_3PL ? 1 : 0) + (model >= VIPSY_IRT_4PL ? 1 : 0);
}
__host__ __device__ inline size_t fused_smem_bytes(int J, int K, int model) {
  size_t s = fu_align16((size_t)J * K * 4);                          // a[j][k]
  s += (size_t)J * 16;                                               // {-D log2e b, c, d-c, 1-d}
  s += (size_t)kFuWarps * fu_align16((size_t)fused_nq(K, model) * J * 4);   // warp-private item accumulators
  s += (size_t)kFuWarps * fu_align16((size_t)32 * J + 16);          // response tiles
  return s;
}

// SCORE: REINFORCE estimator for the guide parameters (north star (d); SURVEY A.2 "score mode"): x is a constant,
// each particle's weight w_s = scale (ll_s + log p(x_s) - log q(x_s)) - baseline multiplies the score of q,
// grad_mu log q = L^-T eps, grad_L log q = tril(L^-T eps eps^T) - diag(1 / L_ii); item gradients are unchanged.
// SC: particles held in registers at once (fused_sc(): 8 / 4 / 2 by K; 2 when the call has at most two particles)
__host__ __device__ constexpr int fused_sc(int K) { return K <= 2 ? 8 : (K <= 6 ? 4 : 2); }

template <int K, int MODEL, bool SCORE, int SC>
__global__ void __launch_bounds__(kFuThreads) mirt_fused_kernel(const MirtParams p) {
  constexpr int NQ = K + 1 + (MODEL >= 3 ? 1 : 0) + (MODEL >= 4 ? 1 : 0);
  constexpr int TRI = K * (K + 1) / 2;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int J = p.J, S = p.S;
  float* tabA = reinterpret_cast<float*>(smem_raw);
  float4* tabI = reinterpret_cast<float4*>(smem_raw + fu_align16((size_t)J * K * 4));
  unsigned char* q0 = reinterpret_cast<unsigned char*>(tabI) + (size_t)J * 16;
  const size_t acc_bytes = fu_align16((size_t)NQ * J * 4);
  float* acc_all = reinterpret_cast<float*>(q0);
  unsigned char* ytiles = q0 + (size_t)kFuWarps * acc_bytes;
  const size_t ytile_bytes = fu_align16((size_t)32 * J + 16);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* acc = reinterpret_cast<float*>(q0 + (size_t)warp * acc_bytes);       // [NQ][J]
  unsigned char* ytile = ytiles + (size_t)warp * ytile_bytes;
  const float negDl2e = -p.D * kLog2e;

  for (int i = threadIdx.x; i < J * K; i += kFuThreads) {
    const int j = i / K, k = i - j * K;
    tabA[i] = p.a[(size_t)k * J + j];
  }
  for (int j = threadIdx.x; j < J; j += kFuThreads) {
    const float cj = MODEL >= 3 ? p.c[j] : 0.f, dj = MODEL >= 4 ? p.d[j] : 1.f;
    tabI[j] = make_float4(negDl2e * p.b[j], cj, dj - cj, 1.f - dj);
  }
  for (int i = threadIdx.x; i < (int)(kFuWarps * acc_bytes / 4); i += kFuThreads) acc_all[i] = 0.f;
  __syncthreads();

  const float kf = -p.scale / (float)S;
  double loss_acc = 0.0;
  const int64_t ntiles = (p.B + 31) / 32;
  const int64_t gwarp = (int64_t)blockIdx.x * kFuWarps + warp, nwarps = (int64_t)gridDim.x * kFuWarps;
  const bool contiguous = p.idx == nullptr && p.ld_y == J;
  const bool packed = p.chol_layout == VIPSY_CHOL_PACKED;

  for (int64_t tile = gwarp; tile < ntiles; tile += nwarps) {
    const int64_t b0 = tile * 32, bi = b0 + lane;
    const bool valid = bi < p.B;
    const int rows = (int)min((int64_t)32, p.B - b0);
    // ---- stage the tile's response codes (as irt_uni.cu) ----------------------------------------------
    int mis = 0;
    if (contiguous) {
      const unsigned char* g0 = p.y + b0 * (int64_t)J;
      mis = (int)(reinterpret_cast<uintptr_t>(g0) & 15);
      const int4* gal = reinterpret_cast<const int4*>(g0 - mis);
      const int nvec = (mis + rows * J + 15) >> 4;
      int4* dst = reinterpret_cast<int4*>(ytile);
      for (int v = lane; v < nvec; v += 32) dst[v] = __ldg(gal + v);
    } else {
      for (int r = 0; r < rows; ++r) {
        const int64_t row = p.idx ? p.idx[b0 + r] : (b0 + r);
        const unsigned char* g = p.y + row * p.ld_y;
        for (int i = lane; i < J; i += 32) ytile[r * J + i] = g[i];
      }
    }
    __syncwarp();
    const unsigned char* yrow = ytile + mis + lane * J;

    // ---- this person's guide: mean and lower-triangular factor (row-major tril order) ------------------
    float mu[K], L[TRI];
#pragma unroll
    for (int k = 0; k < K; ++k) mu[k] = valid ? p.mu[bi * K + k] : 0.f;
    {
      const float* Lp = p.chol + bi * (packed ? TRI : K * K);
#pragma unroll
      for (int r = 0; r < K; ++r)
#pragma unroll
        for (int c = 0; c <= r; ++c) {
          float v = valid ? Lp[packed ? (r * (r + 1) / 2 + c) : (r * K + c)] : (r == c ? (p.chol_raw ? 0.f : 1.f) : 0.f);
          if (r == c && p.chol_raw) v = __expf(v);
          L[r * (r + 1) / 2 + c] = v;
        }
    }
    float dmu[K], dL[TRI];
#pragma unroll
    for (int k = 0; k < K; ++k) dmu[k] = 0.f;
#pragma unroll
    for (int i = 0; i < TRI; ++i) dL[i] = 0.f;
    float Lsum = 0.f, quad = 0.f;          // sum of log2 q;  0.5 sum_s (|eps_s|^2 - |x_s|^2)

    for (int s0 = 0; s0 < S; s0 += SC) {
      float eps[SC][K], x[SC][K], G[SC][K], ll2[SC];
#pragma unroll
      for (int s = 0; s < SC; ++s) ll2[s] = 0.f;
      if constexpr ((SC * K) % 4 == 0) {
        // in-kernel noise: the chunk's SC * K draws are consecutive elements n = s K + k of the person's stream
        // (noise_coord, K <= 32 layout), i.e. whole Philox blocks: one call per four draws
        if (p.eps == nullptr) {
          float flat[SC * K];
#pragma unroll
          for (int bl = 0; bl < SC * K / 4; ++bl) {
            const float4 z = normal4(p.seed, p.offset, (uint64_t)(p.person_offset + bi), (uint32_t)((s0 * K) >> 2) + bl);
            flat[4 * bl] = z.x; flat[4 * bl + 1] = z.y; flat[4 * bl + 2] = z.z; flat[4 * bl + 3] = z.w;
          }
#pragma unroll
          for (int s = 0; s < SC; ++s)
#pragma unroll
            for (int k = 0; k < K; ++k) eps[s][k] = (valid && s0 + s < S) ? flat[s * K + k] : 0.f;
        }
      }
#pragma unroll
      for (int s = 0; s < SC; ++s) {
        const bool live = valid && s0 + s < S;
        if (((SC * K) % 4 != 0) || p.eps != nullptr) {
#pragma unroll
          for (int k = 0; k < K; ++k) eps[s][k] = live ? noise_at(p, s0 + s, bi, k) : 0.f;
        }
#pragma unroll
        for (int r = 0; r < K; ++r) {
          float v = mu[r];
#pragma unroll
          for (int c = 0; c <= r; ++c) v = fmaf(L[r * (r + 1) / 2 + c], eps[s][c], v);
          x[s][r] = v;
          G[s][r] = 0.f;
          if (live) quad += 0.5f * (eps[s][r] * eps[s][r] - v * v);
        }
      }
      // ---- items ------------------------------------------------------------------------------------
      for (int j = 0; j < J; ++j) {
        const unsigned code = valid ? (unsigned)yrow[j] : 255u;
        const bool obs = code != 255u;
        if (!__any_sync(0xffffffffu, obs)) continue;
        float v[NQ];
#pragma unroll
        for (int q = 0; q < NQ; ++q) v[q] = 0.f;
        if (obs) {
          float a[K];
#pragma unroll
          for (int k = 0; k < K; ++k) a[k] = tabA[j * K + k];
          const float4 it = tabI[j];
          const bool y1 = code != 0u;
          const float base = MODEL >= 3 ? (y1 ? it.y : it.w) : 0.f;
          const float glf = MODEL >= 3 ? p.D * it.z : p.D;           // d logit-gradient factor D (d - c)
          float prod = 1.f;
#pragma unroll
          for (int s = 0; s < SC; ++s) {
            if (s0 + s < S) {                                         // warp-uniform
              float dot = 0.f;
#pragma unroll
              for (int k = 0; k < K; ++k) dot = fmaf(a[k], x[s][k], dot);
              float t = fmaf(negDl2e, dot, it.x);                     // -l log2 e
              t = y1 ? t : -t;
              const float e = ex2_approx(fminf(t, 126.f));
              const float sp = rcp_approx(1.f + e);                   // s (y = 1) | 1 - s (y = 0)
              const float es = e * sp;
              const float q = MODEL >= 3 ? fmaf(it.z, sp, base) : sp;
              const float qc = fminf(fmaxf(q, kEps32), kOneMinusEps32);
              if constexpr (SCORE) {
                ll2[s] += lg2_approx(qc);                             // per-particle log-likelihood (log2 units)
              } else {
                prod *= qc;
                if (SC == 8 && s == 3) { Lsum += lg2_approx(prod); prod = 1.f; }
              }
              if (q == qc) {                                          // the clamp blocks the gradient
                float w, a1 = 0.f, a2 = 0.f;
                if (MODEL >= 3) {
                  const float r = rcp_approx(qc);
                  a1 = r * sp;
                  a2 = r * es;
                  w = a1 * es;
                } else {
                  w = es;
                }
                const float ws = y1 ? w : -w;
                const float gl = ws * glf;
#pragma unroll
                for (int k = 0; k < K; ++k) {
                  if constexpr (!SCORE) G[s][k] = fmaf(gl, a[k], G[s][k]);
                  v[k] = fmaf(ws, x[s][k], v[k]);
                }
                v[K] += ws;
                if (MODEL >= 3) v[K + 1] += y1 ? a2 : -a1;
                if (MODEL >= 4) v[K + 2] += y1 ? a1 : -a2;
              }
            }
          }
          if constexpr (!SCORE) Lsum += lg2_approx(prod);
        }
        // K + 1 (+2) sums over the 32 persons, four quantities per butterfly
#pragma unroll
        for (int q0i = 0; q0i < NQ; q0i += 4) {
          float v4[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) v4[u] = q0i + u < NQ ? v[q0i + u] : 0.f;
          const float r = warp_sum_multi<4>(v4, lane);
          const int own = warp_sum_multi_owner<4>(lane);
          if (own >= 0 && q0i + own < NQ) acc[(q0i + own) * J + j] += r;
        }
      }
      // ---- this chunk's share of d mu, d L ---------------------------------------------------------------
#pragma unroll
      for (int s = 0; s < SC; ++s) {
        if (s0 + s < S) {
          if constexpr (SCORE) {
            float x2 = 0.f, e2 = 0.f, logdet = 0.f;
#pragma unroll
            for (int k = 0; k < K; ++k) {
              x2 = fmaf(x[s][k], x[s][k], x2);
              e2 = fmaf(eps[s][k], eps[s][k], e2);
              logdet += __logf(L[k * (k + 1) / 2 + k]);
            }
            Lsum += ll2[s];
            const float wgt = p.scale * (kLn2 * ll2[s] - 0.5f * x2 + 0.5f * e2 + logdet) - p.baseline;
            float u[K];                                     // L^T u = eps (back substitution)
#pragma unroll
            for (int r = K - 1; r >= 0; --r) {
              float t = eps[s][r];
#pragma unroll
              for (int c = r + 1; c < K; ++c) t = fmaf(-L[c * (c + 1) / 2 + r], u[c], t);
              u[r] = t / L[r * (r + 1) / 2 + r];
            }
#pragma unroll
            for (int r = 0; r < K; ++r) {
              dmu[r] = fmaf(wgt, u[r], dmu[r]);
#pragma unroll
              for (int c = 0; c <= r; ++c) dL[r * (r + 1) / 2 + c] = fmaf(wgt * u[r], eps[s][c], dL[r * (r + 1) / 2 + c]);
              dL[r * (r + 1) / 2 + r] -= wgt / L[r * (r + 1) / 2 + r];
            }
          } else {
#pragma unroll
            for (int r = 0; r < K; ++r) {
              const float gk = G[s][r] - x[s][r];           // g_x = (sum_j d lb/d x) - x
              dmu[r] += gk;
#pragma unroll
              for (int c = 0; c <= r; ++c) dL[r * (r + 1) / 2 + c] = fmaf(gk, eps[s][c], dL[r * (r + 1) / 2 + c]);
            }
          }
        }
      }
    }
    __syncwarp();      // everyone is done with ytile before the next tile overwrites it

    // ---- per-person outputs ----------------------------------------------------------------------------
    if (valid) {
      float logdet = 0.f;
#pragma unroll
      for (int k = 0; k < K; ++k) logdet += __logf(L[k * (k + 1) / 2 + k]);
      loss_acc += (double)(kf * (kLn2 * Lsum + quad + (float)S * logdet));
      const float gf = SCORE ? -1.f / (float)S : kf;      // score weights already carry the plate scale
      if (p.d_mu) {
#pragma unroll
        for (int k = 0; k < K; ++k) p.d_mu[bi * K + k] = gf * dmu[k];
      }
      if (p.d_chol) {
        float* out = p.d_chol + bi * (packed ? TRI : K * K);
#pragma unroll
        for (int r = 0; r < K; ++r)
#pragma unroll
          for (int c = 0; c <= r; ++c) {
            float v = gf * dL[r * (r + 1) / 2 + c];
            if (r == c) {
              const float lii = L[r * (r + 1) / 2 + r];
              if constexpr (SCORE) v = p.chol_raw ? v * lii : v;                  // exp chain rule only
              else v = p.chol_raw ? fmaf(v, lii, -p.scale) : v - p.scale / lii;   // - scale d(sum log L_ii); exp chain rule
            }
            out[packed ? (r * (r + 1) / 2 + c) : (r * K + c)] = v;
          }
      }
    }
  }

  // ---- CTA partials: [d_a (K x J) | d_b | d_c | d_d], loss ------------------------------------------------------
  __shared__ double loss_sm[kFuWarps];
  loss_acc = warp_sum(loss_acc);
  if (lane == 0) loss_sm[warp] = loss_acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < kFuWarps; ++w) t += loss_sm[w];
    p.loss_partial[blockIdx.x] = t;
  }
  float* part = p.item_partial + (size_t)blockIdx.x * (K * J + 3 * J);
  for (int i = threadIdx.x; i < (K + 3) * J; i += kFuThreads) {
    const int q = i / J, j = i - q * J;
    float t = 0.f;
    if (q < NQ) {
#pragma unroll
      for (int w = 0; w < kFuWarps; ++w) t += acc_all[(size_t)w * (acc_bytes / 4) + q * J + j];
      if (q <= K) t *= MODEL >= 3 ? p.D * tabI[j].z : p.D;           // d / d(a, b) carry D (d - c)
    }
    part[i] = t;
  }
}

// ------------------------------------------------------------------------------------------------------------------
bool mirt_fused_supported(const MirtParams& p) {
  // measured (scripts/dev_bench_mirt_fused.py): 2.4-8x faster than the three-kernel path for K = 2..7; at K = 8 the
  // tcgen05 contraction takes over (0.52 vs 0.68 ms).  The score estimator exists only here, so it keeps K = 8.
  const int kmax = p.score ? 8 : 7;
  return p.cov_mode == VIPSY_COV_PERSON && p.K >= 2 && p.K <= kmax && p.model >= VIPSY_IRT_2PL &&
         fused_smem_bytes(p.J, p.K, p.model) <= 200 * 1024;
}

template <int K, int MODEL, bool SCORE, int SC>
static int launch_fused_sc(const MirtParams& p, int max_grid, int* grid_out, cudaStream_t st) {
  auto kern = mirt_fused_kernel<K, MODEL, SCORE, SC>;
  const size_t smem = fused_smem_bytes(p.J, K, MODEL);
  static OccCache occ_cache;   // per instantiation
  int occ = 0;
  if (smem > 48 * 1024) VIPSY_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  VIPSY_CUDA_CHECK(cached_occupancy(occ_cache, kern, kFuThreads, smem, 8, &occ));
  const int64_t ntiles = (p.B + 31) / 32;
  int64_t g = (ntiles + kFuWarps - 1) / kFuWarps;
  const int64_t cap = (int64_t)num_sms() * occ;
  if (g > cap) g = cap;
  if (g > max_grid) g = max_grid;
  if (g < 1) g = 1;
  kern<<<(int)g, kFuThreads, smem, st>>>(p);
  VIPSY_CUDA_CHECK(cudaGetLastError());
  *grid_out = (int)g;
  return VIPSY_OK;
}

template <int K, int MODEL, bool SCORE>
static int launch_fused_one(const MirtParams& p, int max_grid, int* grid_out, cudaStream_t st) {
  // few particles: the wide chunk only costs registers (measured: K = 5, S = 1: 0.39 ms with SC = 2, 0.43 with 4)
  if constexpr (!SCORE && fused_sc(K) > 2)
    if (p.S <= 2) return launch_fused_sc<K, MODEL, SCORE, 2>(p, max_grid, grid_out, st);
  return launch_fused_sc<K, MODEL, SCORE, fused_sc(K)>(p, max_grid, grid_out, st);
}

template <int K>
static int launch_fused_model(const MirtParams& p, int max_grid, int* grid_out, cudaStream_t st) {
  if (p.score) {
    switch (p.model) {
      case VIPSY_IRT_2PL: return launch_fused_one<K, 2, true>(p, max_grid, grid_out, st);
      case VIPSY_IRT_3PL: return launch_fused_one<K, 3, true>(p, max_grid, grid_out, st);
      default: return launch_fused_one<K, 4, true>(p, max_grid, grid_out, st);
    }
  }
  switch (p.model) {
    case VIPSY_IRT_2PL: return launch_fused_one<K, 2, false>(p, max_grid, grid_out, st);
    case VIPSY_IRT_3PL: return launch_fused_one<K, 3, false>(p, max_grid, grid_out, st);
    default: return launch_fused_one<K, 4, false>(p, max_grid, grid_out, st);
  }
}

int launch_mirt_fused(const MirtParams& p, int max_grid, int* grid_out, cudaStream_t st) {
  switch (p.K) {
    case 2: return launch_fused_model<2>(p, max_grid, grid_out, st);
    case 3: return launch_fused_model<3>(p, max_grid, grid_out, st);
    case 4: return launch_fused_model<4>(p, max_grid, grid_out, st);
    case 5: return launch_fused_model<5>(p, max_grid, grid_out, st);
    case 6: return launch_fused_model<6>(p, max_grid, grid_out, st);
    case 7: return launch_fused_model<7>(p, max_grid, grid_out, st);
    default: return launch_fused_model<8>(p, max_grid, grid_out, st);
  }
}

}  // namespace vipsy
