// Multidimensional IRT (K > 1): Trace_ELBO of BaseIRT.model (vi.py:596-615) under the
// MultivariateNormal guides of VIRT / VaeIRT (vi.py:686-693, 706-723), pathwise.
//
// Three stream-ordered kernels, no logits / probabilities / per-cell gradients in HBM:
//   (1) mirt_sample    x_bs = mu_b + L_b eps_bs  (batched packed-tril matvec, L staged in
//                      shared memory once per person, LowerCholeskyTransform fused when raw);
//                      emits x (S*B*K), diag L (B*K) and the entropy/noise terms of the ELBO.
//   (2) mirt_lik       person-tile x item-tile register-blocked contraction x.a + b, the
//                      1..4PL link + masked Bernoulli log-lik fused into its epilogue, the
//                      per-cell d lb / d logit kept in shared memory and fed straight to the
//                      two backward contractions d x = G a^T and d a = x^T G.
//   (3) mirt_cholgrad  d mu, d L = -(scale/S) tril(sum_s g_x eps^T) - scale diag(1/L_ii)
//                      written coalesced in the caller's layout (packed/dense, raw chain rule).
// HBM traffic is the per-person Cholesky factor in and its gradient out.
#include <stdlib.h>

#include "mirt.cuh"
#include "enc_mvn.cuh"

namespace vipsy {

// ----------------------------------------------------------------------------------------
// (1) sampling.  RT threads per person (RT = 32/64/128 >= K), 128/RT persons per CTA pass.
// ----------------------------------------------------------------------------------------
constexpr int kSampThreads = 128;
constexpr int kSampChunk = 8;     // particles per pass over L

template <int RT>
__global__ void __launch_bounds__(kSampThreads) mirt_sample_kernel(const MirtParams p) {
  constexpr int SLOTS = kSampThreads / RT;
  extern __shared__ __align__(16) float smem_f[];
  const int K = p.K, tri = tri_off(K);
  float* Ls_all = smem_f;                              // [SLOTS or 1][tri]
  const bool shared_cov = p.cov_mode == VIPSY_COV_SHARED;
  float* es_all = Ls_all + (size_t)(shared_cov ? 1 : SLOTS) * tri;   // [SLOTS][kSampChunk][K]
  __shared__ float red_sm[kSampThreads / 32];
  const int slot = threadIdx.x / RT, t = threadIdx.x % RT;
  float* Ls = Ls_all + (shared_cov ? 0 : (size_t)slot * tri);
  float* es = es_all + (size_t)slot * kSampChunk * K;

  auto load_L = [&](float* dst, const float* src, int tid, int nthr) {
    if (p.chol_layout == VIPSY_CHOL_PACKED) {
      for (int e = tid; e < tri; e += nthr) dst[e] = src[e];
    } else {
      for (int i = 0; i < K; ++i)
        for (int m = tid; m <= i; m += nthr) dst[tri_off(i) + m] = src[(size_t)i * K + m];
    }
  };
  if (shared_cov) {
    load_L(Ls_all, p.chol, threadIdx.x, kSampThreads);
    __syncthreads();
    if (p.chol_raw) {
      for (int i = threadIdx.x; i < K; i += kSampThreads) Ls_all[tri_off(i) + i] = __expf(Ls_all[tri_off(i) + i]);
    }
    __syncthreads();
  }
  const int64_t stride_L = p.chol_layout == VIPSY_CHOL_PACKED ? tri : (int64_t)K * K;
  const int64_t npass = (p.B + SLOTS - 1) / SLOTS;
  for (int64_t pass = blockIdx.x; pass < npass; pass += gridDim.x) {
    const int64_t b = pass * SLOTS + slot;
    const bool valid = b < p.B;
    __syncthreads();
    if (!shared_cov && valid) load_L(Ls, p.chol + b * stride_L, t, RT);
    __syncthreads();
    float Lii = 1.f;
    if (valid && t < K) {
      float v = Ls[tri_off(t) + t];
      if (!shared_cov && p.chol_raw) v = __expf(v);
      Lii = v;
    }
    __syncthreads();
    if (!shared_cov && valid && t < K && p.chol_raw) Ls[tri_off(t) + t] = Lii;
    float auxp = (valid && t < K) ? (float)p.S * logf(Lii) : 0.f;
    if (valid && t < K && p.diagL) p.diagL[b * K + t] = Lii;
    const float mu_i = (valid && t < K) ? p.mu[b * K + t] : 0.f;
    for (int s0 = 0; s0 < p.S; s0 += kSampChunk) {
      const int ns = min(kSampChunk, p.S - s0);
      __syncthreads();
      if (valid)
        for (int e = t; e < ns * K; e += RT) {
          const float z = noise_at(p, s0 + e / K, b, e % K);
          es[e] = z;
        }
      __syncthreads();
      if (valid && t < K) {
        float acc[kSampChunk];
#pragma unroll
        for (int s = 0; s < kSampChunk; ++s) acc[s] = mu_i;
        const float* Lrow = Ls + tri_off(t);
        for (int m = 0; m <= t; ++m) {
          const float l = Lrow[m];
#pragma unroll
          for (int s = 0; s < kSampChunk; ++s)
            if (s < ns) acc[s] = fmaf(l, es[s * K + m], acc[s]);
        }
#pragma unroll
        for (int s = 0; s < kSampChunk; ++s)
          if (s < ns) {
            p.x[((int64_t)(s0 + s) * p.B + b) * K + t] = acc[s];
            const float e = es[s * K + t];
            auxp = fmaf(0.5f * e, e, auxp);
            auxp = fmaf(-0.5f * acc[s], acc[s], auxp);      // log N(x; 0, I) up to the constant
          }
      }
    }
    // aux_b: fixed-order reduction over the slot's warps
    auxp = warp_sum(auxp);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red_sm[threadIdx.x >> 5] = auxp;
    __syncthreads();
    if (valid && t == 0) {
      float tot = 0.f;
      for (int w = 0; w < RT / 32; ++w) tot += red_sm[slot * (RT / 32) + w];
      *reinterpret_cast<float4*>(p.aux + 4 * b) = make_float4(tot, 0.f, 0.f, 0.f);
    }
  }
}

// ----------------------------------------------------------------------------------------
// (2) likelihood contraction + epilogue + both backward contractions.
//     CTA = 16x16 threads, tile = 64 person-samples x 64 items, KI = ceil(K/16).
// ----------------------------------------------------------------------------------------
constexpr int kTM = 64, kTN = 64, kLikThreads = 256;
constexpr int kGS = kTN + 4;     // row stride of the G / A tiles (float4-aligned)

__host__ __device__ inline size_t lik_smem_floats(int K) {
  const int KP = K + 1;
  return (size_t)kTM * KP        // Xs [64][K+1]
       + (size_t)K * kGS          // As [K][68]
       + (size_t)kTN * KP         // At [64][K+1]
       + (size_t)kTM * kGS        // Gs [64][68]
       + 3 * 16 * kTN             // column-sum scratch
       + 4 * kTN;                 // item constants
}

template <int KI>
__global__ void __launch_bounds__(kLikThreads, 2) mirt_lik_kernel(const MirtParams p) {
  extern __shared__ __align__(16) float smem_f[];
  const int K = p.K, J = p.J, KP = K + 1;
  float* Xs = smem_f;
  float* As = Xs + (size_t)kTM * KP;
  float* At = As + (size_t)K * kGS;
  float* Gs = At + (size_t)kTN * KP;
  float* scr = Gs + (size_t)kTM * kGS;        // [3][16][64]
  float* itc = scr + 3 * 16 * kTN;            // [4][64]: B' = -D b log2e | c | d-c | 1-d
  __shared__ unsigned char ytile[kTM][kTN];

  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int64_t R = (int64_t)p.S * p.B;       // person-sample rows
  const int64_t ntile_m = (R + kTM - 1) / kTM;
  const int ntile_n = (J + kTN - 1) / kTN;
  const int ncol = K * J + 3 * J;
  float* part = p.item_partial + (size_t)blockIdx.x * ncol;
  for (int i = tid; i < ncol; i += kLikThreads) part[i] = 0.f;
  const float negDl2e = -p.D * kLog2e;
  const int model = p.model;

  for (int64_t tm = blockIdx.x; tm < ntile_m; tm += gridDim.x) {
    const int64_t r0 = tm * kTM;
    __syncthreads();
    // ---- X tile (rows contiguous in the workspace) ---------------------------
    for (int e = tid; e < kTM * K; e += kLikThreads) {
      const int pr = e / K, k = e - pr * K;
      Xs[pr * KP + k] = (r0 + pr < R) ? p.x[(r0 + pr) * K + k] : 0.f;
    }
    float acc2[4][KI];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int kk = 0; kk < KI; ++kk) acc2[i][kk] = 0.f;
    float llrow[4] = {0.f, 0.f, 0.f, 0.f};

    for (int tn = 0; tn < ntile_n; ++tn) {
      const int j0 = tn * kTN;
      __syncthreads();
      // ---- A tile in both orientations, item constants, response tile --------
      for (int e = tid; e < K * kTN; e += kLikThreads) {
        const int k = e / kTN, jj = e - k * kTN;
        const float v = (j0 + jj < J && model >= VIPSY_IRT_2PL) ? p.a[(size_t)k * J + j0 + jj]
                                                               : ((j0 + jj < J) ? 1.f : 0.f);
        As[k * kGS + jj] = v;
        At[jj * KP + k] = v;
      }
      if (tid < kTN) {
        const int j = j0 + tid;
        const bool ok = j < J;
        const float cj = (ok && model >= VIPSY_IRT_3PL) ? p.c[j] : 0.f;
        const float dj = (ok && model == VIPSY_IRT_4PL) ? p.d[j] : 1.f;
        itc[tid] = ok ? negDl2e * p.b[j] : 0.f;
        itc[kTN + tid] = cj;
        itc[2 * kTN + tid] = dj - cj;
        itc[3 * kTN + tid] = 1.f - dj;
      }
      for (int e = tid; e < kTM * kTN; e += kLikThreads) {
        const int pr = e / kTN, jj = e - pr * kTN;
        unsigned char code = 255;
        if (r0 + pr < R && j0 + jj < J) {
          const int64_t bb = (r0 + pr) % p.B;
          const int64_t row = p.idx ? p.idx[bb] : bb;
          code = p.y[row * p.ld_y + j0 + jj];
        }
        ytile[pr][jj] = code;
      }
      __syncthreads();

      // ---- GEMM 1: logits (4 rows x 4 items per thread) -----------------------
      float acc[4][4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) acc[i][jj] = 0.f;
      {
        const float* xr = Xs + (ty * 4) * KP;
        const float* ar = As + tx * 4;
#pragma unroll 4
        for (int k = 0; k < K; ++k) {
          const float4 av = *reinterpret_cast<const float4*>(ar + k * kGS);
          const float x0 = xr[k], x1 = xr[KP + k], x2 = xr[2 * KP + k], x3 = xr[3 * KP + k];
          acc[0][0] = fmaf(x0, av.x, acc[0][0]); acc[0][1] = fmaf(x0, av.y, acc[0][1]);
          acc[0][2] = fmaf(x0, av.z, acc[0][2]); acc[0][3] = fmaf(x0, av.w, acc[0][3]);
          acc[1][0] = fmaf(x1, av.x, acc[1][0]); acc[1][1] = fmaf(x1, av.y, acc[1][1]);
          acc[1][2] = fmaf(x1, av.z, acc[1][2]); acc[1][3] = fmaf(x1, av.w, acc[1][3]);
          acc[2][0] = fmaf(x2, av.x, acc[2][0]); acc[2][1] = fmaf(x2, av.y, acc[2][1]);
          acc[2][2] = fmaf(x2, av.z, acc[2][2]); acc[2][3] = fmaf(x2, av.w, acc[2][3]);
          acc[3][0] = fmaf(x3, av.x, acc[3][0]); acc[3][1] = fmaf(x3, av.y, acc[3][1]);
          acc[3][2] = fmaf(x3, av.z, acc[3][2]); acc[3][3] = fmaf(x3, av.w, acc[3][3]);
        }
      }
      // ---- epilogue: link, clamp, log-lik, d lb / d logit (same numerics as irt_uni.cu)
      float cs_c[4] = {0.f, 0.f, 0.f, 0.f}, cs_d[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int pr = ty * 4 + i;
        float gl4[4];
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          const int jl = tx * 4 + jj;
          const unsigned code = ytile[pr][jl];
          float gl = 0.f;
          if (code != 255u) {
            const bool y1 = code != 0u;
            float t = fmaf(negDl2e, acc[i][jj], itc[jl]);        // -l log2e
            t = y1 ? t : -t;
            t = fminf(t, 126.f);
            const float e = ex2_approx(t);
            const float sp = rcp_approx(1.f + e);                 // s (y=1) | 1-s (y=0)
            const float es = e * sp;
            const float dmc = itc[2 * kTN + jl];
            const float base = y1 ? itc[kTN + jl] : itc[3 * kTN + jl];
            const float q = model >= VIPSY_IRT_3PL ? fmaf(dmc, sp, base) : sp;
            const float qc = fminf(fmaxf(q, kEps32), kOneMinusEps32);
            llrow[i] += kLn2 * lg2_approx(qc);
            if (q == qc) {
              const float sgn = y1 ? 1.f : -1.f;
              if (model >= VIPSY_IRT_3PL) {
                const float r = rcp_approx(qc);
                const float us = r * sp, uc = r * es;
                gl = sgn * p.D * dmc * us * es;
                cs_c[jj] += y1 ? uc : -us;
                cs_d[jj] += y1 ? us : -uc;
              } else {
                gl = sgn * p.D * es;
              }
            }
          }
          gl4[jj] = gl;
        }
        *reinterpret_cast<float4*>(Gs + pr * kGS + tx * 4) = make_float4(gl4[0], gl4[1], gl4[2], gl4[3]);
      }
      if (model >= VIPSY_IRT_3PL) {
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          scr[(1 * 16 + ty) * kTN + tx * 4 + jj] = cs_c[jj];
          scr[(2 * 16 + ty) * kTN + tx * 4 + jj] = cs_d[jj];
        }
      }
      __syncthreads();

      // ---- GEMM 2: d x[p][k] += sum_j G[p][j] a[k][j]   (k = tx + 16 kk) ---------
      {
        const float* gr = Gs + (ty * 4) * kGS;
#pragma unroll 2
        for (int jj = 0; jj < kTN; ++jj) {
          const float g0 = gr[jj], g1 = gr[kGS + jj], g2 = gr[2 * kGS + jj], g3 = gr[3 * kGS + jj];
          const float* atr = At + jj * KP + tx;
#pragma unroll
          for (int kk = 0; kk < KI; ++kk) {
            const float av = (tx + 16 * kk < K) ? atr[16 * kk] : 0.f;
            acc2[0][kk] = fmaf(g0, av, acc2[0][kk]);
            acc2[1][kk] = fmaf(g1, av, acc2[1][kk]);
            acc2[2][kk] = fmaf(g2, av, acc2[2][kk]);
            acc2[3][kk] = fmaf(g3, av, acc2[3][kk]);
          }
        }
      }
      // ---- GEMM 3: d a[k][j] += sum_p x[p][k] G[p][j]   (k = ty + 16 kk, 4 items) -
      {
        float acc3[KI][4];
#pragma unroll
        for (int kk = 0; kk < KI; ++kk)
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) acc3[kk][jj] = 0.f;
#pragma unroll 2
        for (int pr = 0; pr < kTM; ++pr) {
          const float4 gv = *reinterpret_cast<const float4*>(Gs + pr * kGS + tx * 4);
          const float* xr = Xs + pr * KP + ty;
#pragma unroll
          for (int kk = 0; kk < KI; ++kk) {
            const float xv = (ty + 16 * kk < K) ? xr[16 * kk] : 0.f;
            acc3[kk][0] = fmaf(xv, gv.x, acc3[kk][0]);
            acc3[kk][1] = fmaf(xv, gv.y, acc3[kk][1]);
            acc3[kk][2] = fmaf(xv, gv.z, acc3[kk][2]);
            acc3[kk][3] = fmaf(xv, gv.w, acc3[kk][3]);
          }
        }
        if (model >= VIPSY_IRT_2PL) {
#pragma unroll
          for (int kk = 0; kk < KI; ++kk) {
            const int k = ty + 16 * kk;
            if (k < K) {
#pragma unroll
              for (int jj = 0; jj < 4; ++jj) {
                const int j = j0 + tx * 4 + jj;
                if (j < J) part[(size_t)k * J + j] += acc3[kk][jj];
              }
            }
          }
        }
      }
      // ---- item column sums: d b (from G), d c, d d (from the scratch rows) ------
      if (tid < kTN && j0 + tid < J) {
        float sb = 0.f;
        for (int pr = 0; pr < kTM; ++pr) sb += Gs[pr * kGS + tid];
        part[(size_t)K * J + j0 + tid] += sb;
        if (model >= VIPSY_IRT_3PL) {
          float sc = 0.f, sd = 0.f;
          for (int w = 0; w < 16; ++w) { sc += scr[(16 + w) * kTN + tid]; sd += scr[(32 + w) * kTN + tid]; }
          part[(size_t)K * J + J + j0 + tid] += sc;
          part[(size_t)K * J + 2 * J + j0 + tid] += sd;
        }
      }
    }
    // ---- per-row outputs: d x and the log-likelihood (prior terms live in aux / cholgrad) ----
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int pr = ty * 4 + i;
      float ll = llrow[i];
      ll += __shfl_xor_sync(0xffffffffu, ll, 8);
      ll += __shfl_xor_sync(0xffffffffu, ll, 4);
      ll += __shfl_xor_sync(0xffffffffu, ll, 2);
      ll += __shfl_xor_sync(0xffffffffu, ll, 1);
#pragma unroll
      for (int kk = 0; kk < KI; ++kk) {
        const int k = tx + 16 * kk;
        if (k < K && r0 + pr < R) p.gx[(r0 + pr) * K + k] = acc2[i][kk];
      }
      if (tx == 0 && r0 + pr < R) p.ell[r0 + pr] = ll;
    }
  }
}

// ----------------------------------------------------------------------------------------
// (3) guide gradients.  One person per CTA pass; warps own rows of L, lanes own columns.
// ----------------------------------------------------------------------------------------
constexpr int kCgThreads = 128;

__global__ void __launch_bounds__(kCgThreads) mirt_cholgrad_kernel(const MirtParams p) {
  extern __shared__ __align__(16) float smem_f[];
  const int K = p.K, S = p.S, tri = tri_off(K);
  float* gs = smem_f;                    // [S][K]  g_x scaled by -(scale/S)
  float* es = gs + (size_t)S * K;        // [S][K]
  float* accL = es + (size_t)S * K;      // [tri]   shared covariance only
  __shared__ double loss_sm[kCgThreads / 32];
  const bool shared_cov = p.cov_mode == VIPSY_COV_SHARED;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float kf = -p.scale / (float)S;
  if (shared_cov)
    for (int e = threadIdx.x; e < tri; e += kCgThreads) accL[e] = 0.f;
  double loss_acc = 0.0;
  const bool packed = p.chol_layout == VIPSY_CHOL_PACKED;
  const int64_t stride_L = packed ? tri : (int64_t)K * K;

  for (int64_t b = blockIdx.x; b < p.B; b += gridDim.x) {
    __syncthreads();
    float ellsum = 0.f;
    const int64_t Rrows = (int64_t)S * p.B;
    for (int e = threadIdx.x; e < S * K; e += kCgThreads) {
      const int s = e / K, k = e - s * K;
      const int64_t at = ((int64_t)s * p.B + b) * K + k;
      float g = -p.x[at];                                   // d log N(x; 0, I) / dx
      for (int part = 0; part < p.n_parts; ++part) g += p.gx[(int64_t)part * Rrows * K + at];
      gs[e] = kf * g;
      es[e] = noise_at(p, s, b, k);
    }
    if (threadIdx.x == 0) {
      for (int s = 0; s < S; ++s)
        for (int part = 0; part < p.n_parts; ++part) ellsum += p.ell[(int64_t)part * Rrows + (int64_t)s * p.B + b];
      const float4 ax = *reinterpret_cast<const float4*>(p.aux + 4 * b);
      loss_acc += (double)(kf * (ellsum + ((ax.x + ax.y) + (ax.z + ax.w))));
    }
    __syncthreads();
    if (p.d_mu)
      for (int k = threadIdx.x; k < K; k += kCgThreads) {
        float t = 0.f;
        for (int s = 0; s < S; ++s) t += gs[s * K + k];
        p.d_mu[b * K + k] = t;
      }
    if (p.d_chol) {
      float* out = shared_cov ? nullptr : p.d_chol + b * stride_L;
      for (int i = warp; i < K; i += kCgThreads / 32) {
        const float Lii = p.diagL[b * K + i];
        for (int m = lane; m <= i; m += 32) {
          float v = 0.f;
          for (int s = 0; s < S; ++s) v = fmaf(gs[s * K + i], es[s * K + m], v);
          if (m == i) {
            v -= p.scale / Lii;               // - scale d(sum log L_ii)/dL_ii
            if (p.chol_raw) v *= Lii;         // LowerCholeskyTransform chain rule
          }
          if (shared_cov) accL[tri_off(i) + m] += v;
          else out[packed ? tri_off(i) + m : (size_t)i * K + m] = v;
        }
      }
    }
  }
  loss_acc = warp_sum(loss_acc);
  if (lane == 0) loss_sm[warp] = loss_acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < kCgThreads / 32; ++w) t += loss_sm[w];
    p.loss_partial[blockIdx.x] = t;
  }
  if (shared_cov && p.d_chol)
    for (int e = threadIdx.x; e < tri; e += kCgThreads) p.chol_partial[(size_t)blockIdx.x * tri + e] = accL[e];
}

// scatter the packed shared-covariance gradient into the caller's layout
__global__ void mirt_shared_chol_out(const float* packed_in, float* out, int K, int dense) {
  const int tri = tri_off(K);
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < tri; e += gridDim.x * blockDim.x) {
    int i = (int)((sqrtf(8.f * (float)e + 1.f) - 1.f) * 0.5f);
    while (tri_off(i + 1) <= e) ++i;
    while (tri_off(i) > e) --i;
    const int m = e - tri_off(i);
    out[dense ? i * K + m : e] = packed_in[e];
  }
}

// ----------------------------------------------------------------------------------------
constexpr int kMirtMaxGrid = 148 * 8;

static size_t a16(size_t x) { return (x + 15) & ~size_t(15); }

struct MirtLayout {
  size_t x, gx, ell, aux, diag, item_part, chol_part, chol_tmp, loss_part, total;
};
static MirtLayout mirt_layout(const VipsyIrtArgs* a) {
  const size_t R = (size_t)a->n_particles * a->batch, K = a->n_dims, J = a->n_items, tri = K * (K + 1) / 2;
  MirtLayout L;
  size_t o = 0;
  const size_t parts = (size_t)mirt_tc_parts((int)J);      // >= 1: planes of d x / log-lik partials
  L.x = o; o += a16(R * K * 4);
  L.gx = o; o += a16(parts * R * K * 4);
  L.ell = o; o += a16(parts * R * 4);
  L.aux = o; o += a16((size_t)a->batch * 16);
  L.diag = o; o += a16((size_t)a->batch * K * 4);
  L.item_part = o; o += a16((size_t)kMirtMaxGrid * (K * J + 3 * J) * 4);
  L.chol_part = o; o += a16(a->cov_mode == VIPSY_COV_SHARED ? (size_t)kMirtMaxGrid * tri * 4 : 0);
  L.chol_tmp = o; o += a16(a->cov_mode == VIPSY_COV_SHARED ? tri * 4 : 0);
  L.loss_part = o; o += a16((size_t)(kMirtMaxGrid + kMaskRows) * 8);
  L.total = o;
  return L;
}

size_t mirt_workspace_bytes(const VipsyIrtArgs* a) { return mirt_layout(a).total; }

template <int KI>
static int launch_lik(const MirtParams& p, int* grid_out, cudaStream_t st) {
  auto k = mirt_lik_kernel<KI>;
  const size_t smem = lik_smem_floats(p.K) * 4;
  static OccCache occ_cache;   // per instantiation
  int occ = 0;
  if (smem > 48 * 1024) VIPSY_CUDA_CHECK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  VIPSY_CUDA_CHECK(cached_occupancy(occ_cache, k, kLikThreads, smem, 64, &occ));
  const int64_t ntile = ((int64_t)p.S * p.B + kTM - 1) / kTM;
  int64_t g = (int64_t)num_sms() * occ;
  if (g > kMirtMaxGrid) g = kMirtMaxGrid;
  if (g > ntile) g = ntile;
  if (g < 1) g = 1;
  *grid_out = (int)g;
  k<<<(int)g, kLikThreads, smem, st>>>(p);
  VIPSY_CUDA_CHECK(cudaGetLastError());
  return VIPSY_OK;
}

int mirt_launch(const VipsyIrtArgs* a, cudaStream_t st, MvnEncParams* enc) {
  const int K = a->n_dims, J = a->n_items, tri = K * (K + 1) / 2;
  VIPSY_REQUIRE(K >= 2 && K <= 128, VIPSY_ERR_UNSUPPORTED, "vipsy_irt_elbo_grad: K=%d outside 2..128", K);
  VIPSY_REQUIRE(a->estimator == VIPSY_EST_PATHWISE ||
                    (K <= 8 && a->cov_mode == VIPSY_COV_PERSON && !(a->flags & VIPSY_IRT_FLAG_NO_FUSED)),
                VIPSY_ERR_UNSUPPORTED,
                "vipsy_irt_elbo_grad: for K > 1 the score estimator is built in the fused kernel only (K <= 8, "
                "per-person factor); the reference never executes it for MultivariateNormal guides (SURVEY Finding 1)");
  VIPSY_REQUIRE(a->chol || enc, VIPSY_ERR_ARG, "vipsy_irt_elbo_grad: chol is required for K > 1");
  VIPSY_REQUIRE(a->cov_mode == VIPSY_COV_SHARED || a->cov_mode == VIPSY_COV_PERSON, VIPSY_ERR_ARG,
                "vipsy_irt_elbo_grad: cov_mode must be SHARED or PERSON for K > 1");
  VIPSY_REQUIRE(!enc || (a->cov_mode == VIPSY_COV_PERSON && a->estimator == VIPSY_EST_PATHWISE), VIPSY_ERR_UNSUPPORTED,
                "vipsy_irt_mvn_encoder_elbo_grad: per-person factor, pathwise estimator");
  VIPSY_REQUIRE((size_t)a->n_particles * K * 8 + (size_t)tri * 4 <= 200 * 1024, VIPSY_ERR_UNSUPPORTED,
                "vipsy_irt_elbo_grad: S*K too large for the guide-gradient kernel (S=%d K=%d)", a->n_particles, K);
  const MirtLayout L = mirt_layout(a);
  VIPSY_REQUIRE(a->workspace && a->workspace_bytes >= L.total, VIPSY_ERR_WORKSPACE,
                "vipsy_irt_elbo_grad: workspace %zu B < required %zu B", a->workspace_bytes, L.total);
  unsigned char* ws = reinterpret_cast<unsigned char*>(a->workspace);
  MirtParams p;
  p.y = a->y; p.ld_y = a->y_row_stride; p.idx = a->idx;
  p.mu = a->mu; p.chol = a->chol; p.eps = a->eps;
  p.a = a->a; p.b = a->b; p.c = a->c; p.d = a->d;
  p.B = a->batch; p.J = J; p.K = K; p.S = a->n_particles; p.model = a->model;
  p.cov_mode = a->cov_mode; p.chol_layout = a->chol_layout; p.chol_raw = a->chol_raw;
  p.D = a->D; p.scale = a->scale;
  p.score = a->estimator == VIPSY_EST_SCORE ? 1 : 0; p.baseline = a->baseline;
  p.seed = a->seed; p.offset = a->offset; p.person_offset = a->person_offset;
  p.x = reinterpret_cast<float*>(ws + L.x); p.gx = reinterpret_cast<float*>(ws + L.gx);
  p.ell = reinterpret_cast<float*>(ws + L.ell); p.aux = reinterpret_cast<float*>(ws + L.aux);
  p.diagL = reinterpret_cast<float*>(ws + L.diag);
  p.item_partial = reinterpret_cast<float*>(ws + L.item_part);
  p.chol_partial = reinterpret_cast<float*>(ws + L.chol_part);
  p.loss_partial = reinterpret_cast<double*>(ws + L.loss_part);
  p.d_mu = a->d_mu; p.d_chol = a->d_chol;
  { const char* e = getenv("VIPSY_DBG"); p.dbg = e ? atoi(e) : 0; }
  if (a->batch == 0) {
    VIPSY_CUDA_CHECK(cudaMemsetAsync(a->loss, 0, 8, st));
    return VIPSY_OK;
  }
  const bool shared_cov = a->cov_mode == VIPSY_COV_SHARED;

  int grid2 = 0, grid3 = 0;
  p.n_parts = 1;
  // K <= 8 with a per-person factor: one fused kernel (warp-shuffle FMAs, nothing staged through HBM)
  const bool fused = !enc && !(a->flags & VIPSY_IRT_FLAG_NO_FUSED) && mirt_fused_supported(p);
  if (fused) {
    int rc = launch_mirt_fused(p, kMirtMaxGrid, &grid2, st);
    if (rc != VIPSY_OK) return rc;
    grid3 = grid2;
  }
  const bool stream = !enc && !fused && mirt_stream_supported(p);
  const bool small = !enc && !fused && !stream && mirt_small_supported(p);
  if (!fused) {
  // (1) sample
  if (enc) {              // amortised guide: fc22 on the tensor cores inside the sampler (enc_mvn.cu)
    enc->m = p;
    int rc = launch_mvn_prep(*enc, st);
    if (rc != VIPSY_OK) return rc;
    rc = launch_mvn_fwd(*enc, st);
    if (rc != VIPSY_OK) return rc;
  } else if (small) {
    int rc = launch_sample_small(p, kMirtMaxGrid, st);
    if (rc != VIPSY_OK) return rc;
  } else if (stream) {
    int rc = launch_sample_stream(p, kMirtMaxGrid, st);
    if (rc != VIPSY_OK) return rc;
  } else {
    const int RT = K <= 32 ? 32 : (K <= 64 ? 64 : 128);
    const int slots = kSampThreads / RT;
    const size_t smem = ((size_t)(shared_cov ? 1 : slots) * tri + (size_t)slots * kSampChunk * K) * 4;
    int64_t g = (a->batch + slots - 1) / slots;
    if (g > kMirtMaxGrid) g = kMirtMaxGrid;
    auto launch = [&](auto kern) -> int {
      if (smem > 48 * 1024) VIPSY_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      kern<<<(int)g, kSampThreads, smem, st>>>(p);
      VIPSY_CUDA_CHECK(cudaGetLastError());
      return VIPSY_OK;
    };
    int rc = RT == 32 ? launch(mirt_sample_kernel<32>) : (RT == 64 ? launch(mirt_sample_kernel<64>) : launch(mirt_sample_kernel<128>));
    if (rc != VIPSY_OK) return rc;
  }
  // (2) likelihood + backward contractions: tcgen05 for K >= 8 (2PL), SIMT otherwise
  const bool want_tc = !(a->flags & VIPSY_IRT_FLAG_NO_TENSOR) && mirt_tc_supported(p);
  if (want_tc) {
    int rc = launch_lik_tc(p, &grid2, &p.n_parts, st);
    if (rc != VIPSY_OK) return rc;
  } else {
    const int KI = (K + 15) / 16;
    int rc;
    switch (KI) {
      case 1: rc = launch_lik<1>(p, &grid2, st); break;
      case 2: rc = launch_lik<2>(p, &grid2, st); break;
      case 3: rc = launch_lik<3>(p, &grid2, st); break;
      case 4: rc = launch_lik<4>(p, &grid2, st); break;
      case 5: rc = launch_lik<5>(p, &grid2, st); break;
      case 6: rc = launch_lik<6>(p, &grid2, st); break;
      case 7: rc = launch_lik<7>(p, &grid2, st); break;
      default: rc = launch_lik<8>(p, &grid2, st); break;
    }
    if (rc != VIPSY_OK) return rc;
  }
  // (3) guide gradients + loss partials
  if (enc) {              // d mu + loss here; d fc22.weight / bias and d hidden from on-chip outer products
    enc->m = p;           // n_parts is known now
    int rc = launch_mvn_tail(*enc, &grid3, st);
    if (rc != VIPSY_OK) return rc;
    if (a->d_mu) {
      rc = launch_mvn_bwd(*enc, st);
      if (rc != VIPSY_OK) return rc;
    }
  } else if (small) {
    int rc = launch_cholgrad_small(p, kMirtMaxGrid, &grid3, st);
    if (rc != VIPSY_OK) return rc;
  } else if (stream) {
    int rc = launch_cholgrad_stream(p, kMirtMaxGrid, &grid3, st);
    if (rc != VIPSY_OK) return rc;
  } else {
    const size_t smem = ((size_t)2 * a->n_particles * K + (shared_cov ? tri : 0)) * 4;
    int64_t g = a->batch;
    const int64_t cap = shared_cov ? (int64_t)num_sms() * 2 : kMirtMaxGrid;
    if (g > cap) g = cap;
    grid3 = (int)g;
    if (smem > 48 * 1024) VIPSY_CUDA_CHECK(cudaFuncSetAttribute(mirt_cholgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    mirt_cholgrad_kernel<<<grid3, kCgThreads, smem, st>>>(p);
    VIPSY_CUDA_CHECK(cudaGetLastError());
  }
  }   // !fused
  // masked cells: log(1 - eps) each (common.cuh), as extra loss partial rows
  int mask_rows = 0;
  {
    int rc = launch_masked_term(a->y, a->y_row_stride, a->idx, a->batch, J, a->scale,
                                (a->flags & VIPSY_IRT_FLAG_MASKED_KNOWN) ? a->n_masked_cells : -1,
                                p.loss_partial + grid3, &mask_rows, st);
    if (rc != VIPSY_OK) return rc;
  }
  // (4) fixed-order reductions: item gradients (+ loss), shared-covariance gradient
  {
    FinalizeSegs segs{};
    segs.n = 4;
    segs.out[0] = a->model >= VIPSY_IRT_2PL ? a->d_a : nullptr; segs.cols[0] = K * J; segs.mask[0] = a->a_free;
    segs.out[1] = a->d_b; segs.cols[1] = J;
    segs.out[2] = a->model >= VIPSY_IRT_3PL ? a->d_c : nullptr; segs.cols[2] = J;
    segs.out[3] = a->model >= VIPSY_IRT_4PL ? a->d_d : nullptr; segs.cols[3] = J;
    int rc = launch_finalize(p.item_partial, grid2, K * J + 3 * J, -a->scale / (float)a->n_particles, segs,
                             p.loss_partial, grid3 + mask_rows, a->loss, st);
    if (rc != VIPSY_OK) return rc;
  }
  if (shared_cov && a->d_chol) {
    float* tmp = reinterpret_cast<float*>(ws + L.chol_tmp);
    FinalizeSegs segs{};
    segs.n = 1; segs.out[0] = tmp; segs.cols[0] = tri;
    int rc = launch_finalize(p.chol_partial, grid3, tri, 1.0f, segs, nullptr, 0, nullptr, st);
    if (rc != VIPSY_OK) return rc;
    mirt_shared_chol_out<<<(tri + 255) / 256, 256, 0, st>>>(tmp, a->d_chol, K, a->chol_layout == VIPSY_CHOL_DENSE);
    VIPSY_CUDA_CHECK(cudaGetLastError());
  }
  return VIPSY_OK;
}

}  // namespace vipsy
