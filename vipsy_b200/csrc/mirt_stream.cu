// Streaming forms of the two guide-side kernels of the K > 1 path for a per-person Cholesky factor
// and S <= 2 particles (the reference's setting: Trace_ELBO(num_particles=1), test.py:361):
//
//   mirt_sample_stream   x_bs = mu_b + L_b eps_bs         reads  L once   (K(K+1)/2 floats / person)
//   mirt_cholgrad_stream d L_b = -(scale/S) tril(sum_s g_bs eps_bs^T) - scale diag(1/L_ii)
//                                                          writes d L once
//
// Both are pure HBM streams (vi.py:686-693, 717-723: a 20 KB factor per person at K = 100), so the
// design goal is bytes in flight, not math: one warp per person, lanes stride the columns of a row
// (coalesced 128-byte row segments straight from / to global memory, no shared-memory staging, no
// CTA barriers), 32 rows of partial dot products per lane reduced with one transposing butterfly
// (31 shuffles per 32 rows), noise regenerated from the Philox counter in both kernels.
#include <algorithm>
#include <cstdlib>
#include <type_traits>

#include "mirt.cuh"
#include "umma.cuh"

namespace vipsy {

constexpr int kStreamWarps = 8;
constexpr int kStreamThreads = kStreamWarps * 32;

// v[r] (r = 0..31) per lane -> returns on lane l the sum over all lanes of v[l]. Fixed order: deterministic.
__device__ __forceinline__ float transpose_reduce32(float (&v)[32], int lane) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
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

template <int S, bool PACKED>
__global__ void __launch_bounds__(kStreamThreads) mirt_sample_stream_kernel(const MirtParams p) {
  const int K = p.K, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t stride_L = PACKED ? tri_off(K) : (int64_t)K * K;
  const int nrb = (K + 31) >> 5;

  for (int64_t b = (int64_t)blockIdx.x * kStreamWarps + warp; b < p.B; b += (int64_t)gridDim.x * kStreamWarps) {
    const float* Lp = p.chol + b * stride_L;
    // per-lane state for columns m = lane + 32 c: noise, mean, own diagonal element L_mm
    float e[S][4], mu_r[4], lii[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int m = lane + 32 * c;
      const bool ok = c < nrb && m < K;
      mu_r[c] = ok ? p.mu[b * K + m] : 0.f;
      float dg = ok ? Lp[PACKED ? tri_off(m) + m : m * K + m] : 0.f;
      if (p.chol_raw) dg = __expf(dg);
      lii[c] = ok ? dg : 1.f;
    }
#pragma unroll
    for (int s = 0; s < S; ++s) noise_lane4(p, s, b, lane, e[s]);
    float auxp = 0.f;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int m = lane + 32 * c;
      if (c < nrb && m < K) {
        p.diagL[b * K + m] = lii[c];
        auxp += (float)S * __logf(lii[c]);
      }
#pragma unroll
      for (int s = 0; s < S; ++s) auxp = fmaf(0.5f * e[s][c], e[s][c], auxp);
    }

#pragma unroll
    for (int rb = 0; rb < 4; ++rb) {
      if (rb < nrb) {
        float part[S][32];
#pragma unroll
        for (int sg = 0; sg < 4; ++sg) {                   // 8 rows at a time: all their loads in flight together
          float l[8][4];
#pragma unroll
          for (int r8 = 0; r8 < 8; ++r8) {
            const int r = sg * 8 + r8, i = 32 * rb + r;
            const float* Lrow = Lp + (PACKED ? tri_off(i) : i * K);
            const bool row_ok = i < K;                     // warp-uniform
#pragma unroll
            for (int c = 0; c < rb; ++c) l[r8][c] = row_ok ? Lrow[lane + 32 * c] : 0.f;
            l[r8][rb] = (row_ok && lane < r) ? Lrow[lane + 32 * rb] : 0.f;
          }
#pragma unroll
          for (int r8 = 0; r8 < 8; ++r8) {
            const int r = sg * 8 + r8;
            const float ld = (lane == r) ? lii[rb] : l[r8][rb];   // diagonal element goes through exp (raw)
#pragma unroll
            for (int s = 0; s < S; ++s) {
              float acc = ld * e[s][rb];
#pragma unroll
              for (int c = 0; c < rb; ++c) acc = fmaf(l[r8][c], e[s][c], acc);
              part[s][r] = acc;
            }
          }
        }
        const int i = 32 * rb + lane;
#pragma unroll
        for (int s = 0; s < S; ++s) {
          const float tot = transpose_reduce32(part[s], lane);
          if (i < K) {
            const float xv = mu_r[rb] + tot;
            p.x[((int64_t)s * p.B + b) * K + i] = xv;
            auxp = fmaf(-0.5f * xv, xv, auxp);
          }
        }
      }
    }
    auxp = warp_sum(auxp);
    if (lane == 0) *reinterpret_cast<float4*>(p.aux + 4 * b) = make_float4(auxp, 0.f, 0.f, 0.f);
  }
}

// PS particles per pass over the factor (1 or 2); S > PS takes ceil(S / PS) passes per person, the later ones
// re-reading the person's 20 KB factor from L1 / L2 (it was just streamed in), so HBM still sees L once.
template <int PS, bool PACKED>
__global__ void __launch_bounds__(kStreamThreads) mirt_sample_stream_multi_kernel(const MirtParams p) {
  const int K = p.K, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t stride_L = PACKED ? tri_off(K) : (int64_t)K * K;
  const int nrb = (K + 31) >> 5;

  for (int64_t b = (int64_t)blockIdx.x * kStreamWarps + warp; b < p.B; b += (int64_t)gridDim.x * kStreamWarps) {
    const float* Lp = p.chol + b * stride_L;
    // per-lane state for columns m = lane + 32 c: mean, own diagonal element L_mm
    float mu_r[4], lii[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int m = lane + 32 * c;
      const bool ok = c < nrb && m < K;
      mu_r[c] = ok ? p.mu[b * K + m] : 0.f;
      float dg = ok ? Lp[PACKED ? tri_off(m) + m : m * K + m] : 0.f;
      if (p.chol_raw) dg = __expf(dg);
      lii[c] = ok ? dg : 1.f;
    }
    float auxp = 0.f;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int m = lane + 32 * c;
      if (c < nrb && m < K) {
        p.diagL[b * K + m] = lii[c];
        auxp += (float)p.S * __logf(lii[c]);
      }
    }

    for (int s0 = 0; s0 < p.S; s0 += PS) {
      float e[PS][4];
#pragma unroll
      for (int s = 0; s < PS; ++s) {
        if (s0 + s < p.S) noise_lane4(p, s0 + s, b, lane, e[s]);
        else { e[s][0] = e[s][1] = e[s][2] = e[s][3] = 0.f; }
      }
#pragma unroll
      for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int s = 0; s < PS; ++s) auxp = fmaf(0.5f * e[s][c], e[s][c], auxp);

#pragma unroll
      for (int rb = 0; rb < 4; ++rb) {
        if (rb < nrb) {
          float part[PS][32];
#pragma unroll
          for (int sg = 0; sg < 4; ++sg) {                   // 8 rows at a time: all their loads in flight together
            float l[8][4];
#pragma unroll
            for (int r8 = 0; r8 < 8; ++r8) {
              const int r = sg * 8 + r8, i = 32 * rb + r;
              const float* Lrow = Lp + (PACKED ? tri_off(i) : i * K);
              const bool row_ok = i < K;                     // warp-uniform
#pragma unroll
              for (int c = 0; c < rb; ++c) l[r8][c] = row_ok ? Lrow[lane + 32 * c] : 0.f;
              l[r8][rb] = (row_ok && lane < r) ? Lrow[lane + 32 * rb] : 0.f;
            }
#pragma unroll
            for (int r8 = 0; r8 < 8; ++r8) {
              const int r = sg * 8 + r8;
              const float ld = (lane == r) ? lii[rb] : l[r8][rb];   // diagonal element goes through exp (raw)
#pragma unroll
              for (int s = 0; s < PS; ++s) {
                float acc = ld * e[s][rb];
#pragma unroll
                for (int c = 0; c < rb; ++c) acc = fmaf(l[r8][c], e[s][c], acc);
                part[s][r] = acc;
              }
            }
          }
          const int i = 32 * rb + lane;
#pragma unroll
          for (int s = 0; s < PS; ++s) {
            const float tot = transpose_reduce32(part[s], lane);
            if (i < K && s0 + s < p.S) {
              const float xv = mu_r[rb] + tot;
              p.x[((int64_t)(s0 + s) * p.B + b) * K + i] = xv;
              auxp = fmaf(-0.5f * xv, xv, auxp);
            }
          }
        }
      }
    }
    auxp = warp_sum(auxp);
    if (lane == 0) *reinterpret_cast<float4*>(p.aux + 4 * b) = make_float4(auxp, 0.f, 0.f, 0.f);
  }
}

// Raw inputs of one person for the guide-gradient stream, fetched one person AHEAD of their use so the
// global-load latency hides behind the previous person's row stores.
template <int S>
struct CgInputs {
  float x[S][4], g[S][4][4], lii[4], ev;
};

template <int S>
__device__ __forceinline__ void cg_fetch(const MirtParams& p, int64_t b, int lane, int nrb, int64_t Rrows, CgInputs<S>& in) {
  const int K = p.K;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const int m = lane + 32 * c;
    const bool ok = c < nrb && m < K;
    in.lii[c] = ok ? p.diagL[b * K + m] : 1.f;
#pragma unroll
    for (int s = 0; s < S; ++s) {
      const int64_t at = ((int64_t)s * p.B + b) * K + m;
      in.x[s][c] = ok ? p.x[at] : 0.f;
#pragma unroll
      for (int u = 0; u < 4; ++u) in.g[s][c][u] = (ok && u < p.n_parts) ? p.gx[(int64_t)u * Rrows * K + at] : 0.f;
    }
  }
  // loss inputs: lanes fetch the 4 aux words and the S * n_parts log-likelihood partials in parallel
  float ev = lane < 4 ? p.aux[4 * b + lane] : 0.f;
  for (int q = lane; q < S * p.n_parts; q += 32) {
    const int s = q / p.n_parts, part = q - s * p.n_parts;
    ev += p.ell[(int64_t)part * Rrows + (int64_t)s * p.B + b];
  }
  in.ev = ev;
}

template <int S, bool PACKED>
__global__ void __launch_bounds__(kStreamThreads) mirt_cholgrad_stream_kernel(const MirtParams p) {
  __shared__ double loss_sm[kStreamWarps];
  const int K = p.K, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t stride_L = PACKED ? tri_off(K) : (int64_t)K * K;
  const int nrb = (K + 31) >> 5;
  const float kf = -p.scale / (float)S;
  const int64_t Rrows = (int64_t)S * p.B;
  const int64_t step = (int64_t)gridDim.x * kStreamWarps;
  double loss_acc = 0.0;

  int64_t b = (int64_t)blockIdx.x * kStreamWarps + warp;
  CgInputs<S> cur, nxt;
  if (b < p.B) cg_fetch<S>(p, b, lane, nrb, Rrows, cur);
  for (; b < p.B; b += step) {
    if (b + step < p.B) cg_fetch<S>(p, b + step, lane, nrb, Rrows, nxt);
    // per-lane state for columns m = lane + 32 c; the diagonal entry of row m is v * dmul + dadd
    float e[S][4], gk[S][4], dmul[4], dadd[4];
#pragma unroll
    for (int s = 0; s < S; ++s) noise_lane4(p, s, b, lane, e[s]);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int m = lane + 32 * c;
      const bool ok = c < nrb && m < K;
      dmul[c] = p.chol_raw ? cur.lii[c] : 1.f;                  // LowerCholeskyTransform chain rule
      dadd[c] = p.chol_raw ? -p.scale : -p.scale / cur.lii[c];  // - scale d(sum log L_ii)/dL_ii
      float dmu = 0.f;
#pragma unroll
      for (int s = 0; s < S; ++s) {
        float g = -cur.x[s][c] + ((cur.g[s][c][0] + cur.g[s][c][1]) + (cur.g[s][c][2] + cur.g[s][c][3]));
        if (ok)
          for (int part = 4; part < p.n_parts; ++part)          // item tiles beyond 512 items (rare)
            g += p.gx[(int64_t)part * Rrows * K + ((int64_t)s * p.B + b) * K + m];
        g *= kf;
        gk[s][c] = g;
        dmu += g;
      }
      if (ok && p.d_mu) p.d_mu[b * K + m] = dmu;
    }
    {
      const float ev = warp_sum(cur.ev);
      if (lane == 0) loss_acc += (double)(kf * ev);
    }
    if (p.d_chol) {
      float* out = p.d_chol + b * stride_L;
#pragma unroll
      for (int rb = 0; rb < 4; ++rb) {
        if (rb < nrb) {
#pragma unroll
          for (int r = 0; r < 32; ++r) {
            const int i = 32 * rb + r;
            if (i < K) {                                    // warp-uniform
              float gi[S];
#pragma unroll
              for (int s = 0; s < S; ++s) gi[s] = __shfl_sync(0xffffffffu, gk[s][rb], r);
              float* orow = out + (PACKED ? tri_off(i) : i * K);
#pragma unroll
              for (int c = 0; c < rb; ++c) {
                float v = gi[0] * e[0][c];
#pragma unroll
                for (int s = 1; s < S; ++s) v = fmaf(gi[s], e[s][c], v);
                orow[lane + 32 * c] = v;
              }
              if (lane <= r) {
                float v = gi[0] * e[0][rb];
#pragma unroll
                for (int s = 1; s < S; ++s) v = fmaf(gi[s], e[s][rb], v);
                if (lane == r) v = fmaf(v, dmul[rb], dadd[rb]);
                orow[lane + 32 * rb] = v;
              }
            }
          }
        }
      }
    }
    cur = nxt;
  }
  loss_acc = warp_sum(loss_acc);
  if (lane == 0) loss_sm[warp] = loss_acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < kStreamWarps; ++w) t += loss_sm[w];
    p.loss_partial[blockIdx.x] = t;
  }
}

// Guide gradient for 2 < S <= kCgMaxS particles: same mapping (warp per person, lanes over the columns of a row,
// coalesced row stores), particle loop inside the row loop; g and eps of all particles stay in registers.
constexpr int kCgMaxS = 8;

template <bool PACKED>
__global__ void __launch_bounds__(kStreamThreads) mirt_cholgrad_stream_multi_kernel(const MirtParams p) {
  __shared__ double loss_sm[kStreamWarps];
  const int K = p.K, S = p.S, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t stride_L = PACKED ? tri_off(K) : (int64_t)K * K;
  const int nrb = (K + 31) >> 5;
  const float kf = -p.scale / (float)S;
  const int64_t Rrows = (int64_t)S * p.B;
  double loss_acc = 0.0;

  for (int64_t b = (int64_t)blockIdx.x * kStreamWarps + warp; b < p.B; b += (int64_t)gridDim.x * kStreamWarps) {
    float e[kCgMaxS][4], gk[kCgMaxS][4], dmul[4], dadd[4], dmu[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int s = 0; s < kCgMaxS; ++s) {
      if (s < S) {
        noise_lane4(p, s, b, lane, e[s]);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int m = lane + 32 * c;
          float g = 0.f;
          if (c < nrb && m < K) {
            const int64_t at = ((int64_t)s * p.B + b) * K + m;
            // the first four item-tile planes are fetched together (independent loads), the rest (J > 512) in a loop
            float g4[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) g4[u] = u < p.n_parts ? p.gx[(int64_t)u * Rrows * K + at] : 0.f;
            g = -p.x[at] + ((g4[0] + g4[1]) + (g4[2] + g4[3]));
            for (int part = 4; part < p.n_parts; ++part) g += p.gx[(int64_t)part * Rrows * K + at];
          }
          g *= kf;
          gk[s][c] = g;
          dmu[c] += g;
        }
      } else {
#pragma unroll
        for (int c = 0; c < 4; ++c) { e[s][c] = 0.f; gk[s][c] = 0.f; }
      }
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int m = lane + 32 * c;
      const bool ok = c < nrb && m < K;
      const float lii = ok ? p.diagL[b * K + m] : 1.f;
      dmul[c] = p.chol_raw ? lii : 1.f;                  // LowerCholeskyTransform chain rule
      dadd[c] = p.chol_raw ? -p.scale : -p.scale / lii;  // - scale d(sum log L_ii)/dL_ii
      if (ok && p.d_mu) p.d_mu[b * K + m] = dmu[c];
    }
    {
      float ev = lane < 4 ? p.aux[4 * b + lane] : 0.f;
      for (int q = lane; q < S * p.n_parts; q += 32) {
        const int s = q / p.n_parts, part = q - s * p.n_parts;
        ev += p.ell[(int64_t)part * Rrows + (int64_t)s * p.B + b];
      }
      ev = warp_sum(ev);
      if (lane == 0) loss_acc += (double)(kf * ev);
    }
    if (p.d_chol) {
      float* out = p.d_chol + b * stride_L;
#pragma unroll
      for (int rb = 0; rb < 4; ++rb) {
        if (rb < nrb) {
#pragma unroll 4
          for (int r = 0; r < 32; ++r) {
            const int i = 32 * rb + r;
            if (i < K) {                                    // warp-uniform
              float gi[kCgMaxS];
#pragma unroll
              for (int s = 0; s < kCgMaxS; ++s) gi[s] = __shfl_sync(0xffffffffu, gk[s][rb], r);
              float* orow = out + (PACKED ? tri_off(i) : i * K);
#pragma unroll
              for (int c = 0; c < rb; ++c) {
                float v = gi[0] * e[0][c];
#pragma unroll
                for (int s = 1; s < kCgMaxS; ++s) v = fmaf(gi[s], e[s][c], v);
                orow[lane + 32 * c] = v;
              }
              if (lane <= r) {
                float v = gi[0] * e[0][rb];
#pragma unroll
                for (int s = 1; s < kCgMaxS; ++s) v = fmaf(gi[s], e[s][rb], v);
                if (lane == r) v = fmaf(v, dmul[rb], dadd[rb]);
                orow[lane + 32 * rb] = v;
              }
            }
          }
        }
      }
    }
  }
  loss_acc = warp_sum(loss_acc);
  if (lane == 0) loss_sm[warp] = loss_acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < kStreamWarps; ++w) t += loss_sm[w];
    p.loss_partial[blockIdx.x] = t;
  }
}

// ---------------------------------------------------------------------------------------------------
// Sampler fed by bulk copies (packed per-person factor, S <= 2): the factor stream is pulled through shared
// memory by one producer thread with cp.async.bulk into a ring of stages that complete on mbarriers, so the
// bytes in flight (nst x 32-40 KB per SM) no longer depend on how many load registers the math warps can spare,
// and the math reads rows with conflict-free LDS instead of predicated global loads.
//   stage  = G consecutive persons, G = 1 / 2 / 4 so that G K(K+1)/2 floats are a multiple of 16 bytes (what a
//            bulk copy needs; K = 100: two persons, 40 400 B)
//   group  = 4 warps working on one stage; a person belongs to WPP = 4 / G of them, warp h of a person owns the
//            rows i = WPP ri + h: the column chunk that holds the diagonal, (WPP ri) / 32, is a compile-time
//            constant of the unrolled row loop, rows at or beyond K cost nothing but a uniform test
//   ring   = item n of the CTA -> stage n % nst, group n % 4 (nst = 5 at K = 100: four stages being read, one in
//            flight); noise, means and the noise stash are prepared BEFORE the warp waits for its stage.
//            Successive uses of a stage belong to DIFFERENT groups, and an mbarrier phase is one bit: a group that
//            runs a whole ring ahead of a slower one would find the parity of use u - 1 still pending and take it
//            for "use u has landed" (seen as wrong draws and, once the arrival counts were off, a hang -- with
//            injected noise, whose loads skew the groups).  The producer therefore publishes how many items it
//            has issued (`issued`, release / acquire through shared memory) and a group looks at a stage's `full`
//            barrier only after its own item has been issued: every earlier use of the stage has then completed
//            and the parity test is unambiguous.
// The diagonal term L_ii eps_i (through exp for a raw factor), diag L and the log-determinant are finished by
// the lane that owns row i after the transposing butterfly.  aux slot h of a person = its warp h's share.
constexpr int kBulkGroups = 4, kBulkGroupWarps = 4;
constexpr int kBulkConsumerWarps = kBulkGroups * kBulkGroupWarps;
constexpr int kBulkThreads = (kBulkConsumerWarps + 1) * 32;
constexpr int kBulkMaxStages = 6;

template <int S, int WPP>
__global__ void __launch_bounds__(kBulkThreads, 1) mirt_sample_bulk_kernel(const MirtParams p, const int nst,
                                                                         const uint32_t stage_bytes) {
  extern __shared__ __align__(128) unsigned char bulk_smem[];
  constexpr int G = 4 / WPP;          // persons per stage
  constexpr int NBLK = 4 / WPP;       // 32-row blocks a warp can own (K <= 128)
  const int K = p.K, P = tri_off(K);
  float* stash_all = reinterpret_cast<float*>(bulk_smem + (size_t)nst * stage_bytes);       // [warp][S][128]
  uint64_t* full = reinterpret_cast<uint64_t*>(stash_all + kBulkConsumerWarps * S * 128);
  uint64_t* empty = full + kBulkMaxStages;
  volatile int* issued = reinterpret_cast<volatile int*>(empty + kBulkMaxStages);   // items the producer has issued
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < nst; ++i) {
      umma::mbar_init(full + i, 1);
      umma::mbar_init(empty + i, kBulkGroupWarps);
    }
    *issued = 0;
    umma::fence_mbar_init();
  }
  __syncthreads();
  const int64_t ngroups = (p.B + G - 1) / G;
  const int64_t nitems = blockIdx.x < ngroups ? (ngroups - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

  if (warp == kBulkConsumerWarps) {                       // ---- producer
    if (lane == 0) {
      for (int64_t n = 0; n < nitems; ++n) {
        const int st = (int)(n % nst);
        const uint32_t use = (uint32_t)(n / nst);
        if (use > 0) umma::mbar_wait(empty + st, (use - 1) & 1);
        const int64_t b0 = (blockIdx.x + n * gridDim.x) * G;
        const int nv = (int)min((int64_t)G, p.B - b0);
        float* dst = reinterpret_cast<float*>(bulk_smem + (size_t)st * stage_bytes);
        const float* src = p.chol + b0 * P;
        const uint32_t bytes = (uint32_t)nv * (uint32_t)P * 4u, bulk = bytes & ~15u;
        for (uint32_t o = bulk; o < bytes; o += 4) dst[o >> 2] = src[o >> 2];   // last, partial group only
        umma::bulk_load(dst, src, bulk, full + st);
        __threadfence_block();
        *issued = (int)n + 1;
      }
    }
    return;
  }

  // ---- consumers
  const int g = warp / kBulkGroupWarps, wg = warp % kBulkGroupWarps, pi = wg / WPP, h = wg % WPP;
  float* es = stash_all + (size_t)warp * S * 128;
  for (int64_t n = g; n < nitems; n += kBulkGroups) {
    const int st = (int)(n % nst);
    const uint32_t use = (uint32_t)(n / nst);
    const int64_t b_raw = (blockIdx.x + n * gridDim.x) * G + pi;
    const bool valid = b_raw < p.B;                      // warp-uniform; a person beyond the batch (last, partial
    const int64_t b = valid ? b_raw : p.B - 1;           // group) runs as the last person and stores nothing
    float e[S][4], mu_l[NBLK];
#pragma unroll
    for (int s = 0; s < S; ++s) noise_lane4(p, s, b, lane, e[s]);
#pragma unroll
    for (int rblk = 0; rblk < NBLK; ++rblk) {
      const int il = WPP * (32 * rblk + lane) + h;
      mu_l[rblk] = il < K ? p.mu[b * K + il] : 0.f;
    }
#pragma unroll
    for (int s = 0; s < S; ++s)
#pragma unroll
      for (int c = 0; c < 4; ++c) es[s * 128 + lane + 32 * c] = e[s][c];
    __syncwarp();
    while (*issued <= (int)n) {}                         // see the header: parity is only meaningful from here on
    __threadfence_block();
    umma::mbar_wait(full + st, use & 1);
    const float* Ls = reinterpret_cast<const float*>(bulk_smem + (size_t)st * stage_bytes) + (size_t)(valid ? pi : 0) * P;
    const float* Ll = Ls + lane;
    float auxp = 0.f;
#pragma unroll
    for (int rblk = 0; rblk < NBLK; ++rblk) {
      float part[S][32];
      // rows of the block, h as a compile-time constant: offsets, chunk counts and diagonal positions are
      // immediates.  CHECK = false: every row of the group of 8 is below K.
      auto rows8 = [&](auto hc, auto r8c, auto checkc) {
        constexpr int H = decltype(hc)::value, R8 = decltype(r8c)::value;
        constexpr bool CHECK = decltype(checkc)::value;
#pragma unroll
        for (int rr = 0; rr < 8; ++rr) {
          constexpr int kW = WPP;
          const int r = R8 + rr, i = kW * (32 * rblk + r) + H;           // constants after unrolling
          const int nch = i >> 5, lim = i - 32 * nch;                    // lanes left of the diagonal in chunk nch
          // a row at or beyond K reads row 0 instead (in bounds); its sum is never used
          const float* Lrow = Ll + ((!CHECK || i < K) ? tri_off(i) : 0);
          float acc[S];
#pragma unroll
          for (int s = 0; s < S; ++s) acc[s] = 0.f;
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            if (c < nch) {
              const float v = Lrow[32 * c];
#pragma unroll
              for (int s = 0; s < S; ++s) acc[s] = fmaf(v, e[s][c], acc[s]);
            } else if (c == nch && lim > 0) {
              const float v = lane < lim ? Lrow[32 * c] : 0.f;
#pragma unroll
              for (int s = 0; s < S; ++s) acc[s] = fmaf(v, e[s][c], acc[s]);
            }
          }
#pragma unroll
          for (int s = 0; s < S; ++s) part[s][r] = acc[s];
        }
      };
      auto fill = [&](auto hc) {
        constexpr int H = decltype(hc)::value;
        auto group = [&](auto r8c) {
          constexpr int R8 = decltype(r8c)::value;
          if (WPP * (32 * rblk + R8 + 7) + H < K) {
            rows8(hc, r8c, std::false_type{});
          } else if (WPP * (32 * rblk + R8) + H < K) {
            rows8(hc, r8c, std::true_type{});
          } else {
#pragma unroll
            for (int rr = 0; rr < 8; ++rr)
#pragma unroll
              for (int s = 0; s < S; ++s) part[s][R8 + rr] = 0.f;
          }
        };
        group(std::integral_constant<int, 0>{});
        group(std::integral_constant<int, 8>{});
        group(std::integral_constant<int, 16>{});
        group(std::integral_constant<int, 24>{});
      };
      if constexpr (WPP == 1) {
        fill(std::integral_constant<int, 0>{});
      } else if constexpr (WPP == 2) {
        if (h == 0) fill(std::integral_constant<int, 0>{});
        else fill(std::integral_constant<int, 1>{});
      } else {
        if (h == 0) fill(std::integral_constant<int, 0>{});
        else if (h == 1) fill(std::integral_constant<int, 1>{});
        else if (h == 2) fill(std::integral_constant<int, 2>{});
        else fill(std::integral_constant<int, 3>{});
      }
      // butterfly and the lane's row outside every branch: no convergence bookkeeping around the shuffles
      const int il = WPP * (32 * rblk + lane) + h;         // the row this lane finishes
      const bool own = il < K;
      float lii = 1.f;
      if (own) {
        lii = Ls[tri_off(il) + il];
        if (p.chol_raw) lii = __expf(lii);
        auxp += (float)S * __logf(lii);
        if (valid) p.diagL[b * K + il] = lii;
      }
#pragma unroll
      for (int s = 0; s < S; ++s) {
        const float tot = transpose_reduce32(part[s], lane);
        if (own) {
          const float ei = es[s * 128 + il];
          const float xv = mu_l[rblk] + fmaf(lii, ei, tot);
          if (valid) p.x[((int64_t)s * p.B + b) * K + il] = xv;
          auxp = fmaf(0.5f * ei, ei, auxp);
          auxp = fmaf(-0.5f * xv, xv, auxp);
        }
      }
    }
    auxp = warp_sum(auxp);
    if (valid) {
      if (lane == 0) p.aux[4 * b + h] = auxp;
      if (h == 0 && lane >= WPP && lane < 4) p.aux[4 * b + lane] = 0.f;
    }
    __syncwarp();
    if (lane == 0) umma::mbar_arrive(empty + st);
  }
}

template <int S, int WPP>
static int launch_sample_bulk_t(const MirtParams& p, int nst, uint32_t stage_bytes, size_t smem, cudaStream_t st) {
  auto k = mirt_sample_bulk_kernel<S, WPP>;
  VIPSY_CUDA_CHECK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  constexpr int G = 4 / WPP;
  const int64_t ngroups = (p.B + G - 1) / G;
  const int grid = (int)std::min<int64_t>(num_sms(), std::max<int64_t>(ngroups, 1));
  k<<<grid, kBulkThreads, smem, st>>>(p, nst, stage_bytes);
  VIPSY_CUDA_CHECK(cudaGetLastError());
  return VIPSY_OK;
}

// VIPSY_NO_BULK=1 keeps the register-fed sampler reachable (A/B measurements).
static bool bulk_disabled() {
  static const bool v = [] { const char* e = getenv("VIPSY_NO_BULK"); return e && e[0] == '1'; }();
  return v;
}

static bool sample_bulk_plan(const MirtParams& p, int* wpp, int* nst, uint32_t* stage_bytes, size_t* smem) {
  if (bulk_disabled() || p.chol_layout != VIPSY_CHOL_PACKED || p.S > 2 || p.K <= 32 || p.K > 128) return false;
  if (reinterpret_cast<uintptr_t>(p.chol) & 15) return false;
  const int P = p.K * (p.K + 1) / 2;
  *wpp = (P % 4 == 0) ? 4 : ((P % 2 == 0) ? 2 : 1);
  *stage_bytes = (uint32_t)(4 / *wpp) * (uint32_t)P * 4u;
  const size_t fixed = (size_t)kBulkConsumerWarps * p.S * 128 * 4 + 2 * kBulkMaxStages * 8 + 16;
  const int fit = (int)((227 * 1024 - fixed) / *stage_bytes);
  *nst = std::min(kBulkMaxStages, fit);
  *smem = (size_t)*nst * *stage_bytes + fixed;
  // One thread issues a bulk copy every ~600 cycles whatever its size: stages below 32 KB cannot carry an SM's share
  // of the stream (measured: K = 100 and 128 gain 7 %, K = 96 / 64 / 50 / 40 lose 7-70 % with 18 / 8 / 20 / 3 KB stages).
  return *nst >= 4 && *stage_bytes >= 32 * 1024;
}

bool mirt_stream_supported(const MirtParams& p) {
  return p.cov_mode == VIPSY_COV_PERSON && p.S <= kCgMaxS && p.K <= 128 && p.K > 32;   // noise_lane4 needs the K > 32 layout
}

static int stream_grid(const MirtParams& p, int max_grid) {
  int64_t g = (p.B + kStreamWarps - 1) / kStreamWarps;
  if (g > max_grid) g = max_grid;
  return (int)(g < 1 ? 1 : g);
}

int launch_sample_stream(const MirtParams& p, int max_grid, cudaStream_t st) {
  {
    int wpp = 0, nst = 0;
    uint32_t sb = 0;
    size_t smem = 0;
    if (sample_bulk_plan(p, &wpp, &nst, &sb, &smem)) {
      if (p.S == 1) {
        if (wpp == 4) return launch_sample_bulk_t<1, 4>(p, nst, sb, smem, st);
        if (wpp == 2) return launch_sample_bulk_t<1, 2>(p, nst, sb, smem, st);
        return launch_sample_bulk_t<1, 1>(p, nst, sb, smem, st);
      }
      if (wpp == 4) return launch_sample_bulk_t<2, 4>(p, nst, sb, smem, st);
      if (wpp == 2) return launch_sample_bulk_t<2, 2>(p, nst, sb, smem, st);
      return launch_sample_bulk_t<2, 1>(p, nst, sb, smem, st);
    }
  }
  const int g = stream_grid(p, max_grid);
  const bool packed = p.chol_layout == VIPSY_CHOL_PACKED;
  if (p.S == 1) {
    if (packed) mirt_sample_stream_kernel<1, true><<<g, kStreamThreads, 0, st>>>(p);
    else mirt_sample_stream_kernel<1, false><<<g, kStreamThreads, 0, st>>>(p);
  } else if (p.S == 2) {
    if (packed) mirt_sample_stream_kernel<2, true><<<g, kStreamThreads, 0, st>>>(p);
    else mirt_sample_stream_kernel<2, false><<<g, kStreamThreads, 0, st>>>(p);
  } else if (p.S & 1) {                // odd particle counts: one particle per pass
    if (packed) mirt_sample_stream_multi_kernel<1, true><<<g, kStreamThreads, 0, st>>>(p);
    else mirt_sample_stream_multi_kernel<1, false><<<g, kStreamThreads, 0, st>>>(p);
  } else {
    if (packed) mirt_sample_stream_multi_kernel<2, true><<<g, kStreamThreads, 0, st>>>(p);
    else mirt_sample_stream_multi_kernel<2, false><<<g, kStreamThreads, 0, st>>>(p);
  }
  VIPSY_CUDA_CHECK(cudaGetLastError());
  return VIPSY_OK;
}

int launch_cholgrad_stream(const MirtParams& p, int max_grid, int* grid_out, cudaStream_t st) {
  const int g = stream_grid(p, max_grid);
  const bool packed = p.chol_layout == VIPSY_CHOL_PACKED;
  if (p.S == 1) {
    if (packed) mirt_cholgrad_stream_kernel<1, true><<<g, kStreamThreads, 0, st>>>(p);
    else mirt_cholgrad_stream_kernel<1, false><<<g, kStreamThreads, 0, st>>>(p);
  } else if (p.S == 2) {
    if (packed) mirt_cholgrad_stream_kernel<2, true><<<g, kStreamThreads, 0, st>>>(p);
    else mirt_cholgrad_stream_kernel<2, false><<<g, kStreamThreads, 0, st>>>(p);
  } else {
    if (packed) mirt_cholgrad_stream_multi_kernel<true><<<g, kStreamThreads, 0, st>>>(p);
    else mirt_cholgrad_stream_multi_kernel<false><<<g, kStreamThreads, 0, st>>>(p);
  }
  VIPSY_CUDA_CHECK(cudaGetLastError());
  *grid_out = g;
  return VIPSY_OK;
}

}  // namespace vipsy
