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
#include "mirt.cuh"

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

bool mirt_stream_supported(const MirtParams& p) {
  return p.cov_mode == VIPSY_COV_PERSON && p.S <= kCgMaxS && p.K <= 128 && p.K > 32;   // noise_lane4 needs the K > 32 layout
}

static int stream_grid(const MirtParams& p, int max_grid) {
  int64_t g = (p.B + kStreamWarps - 1) / kStreamWarps;
  if (g > max_grid) g = max_grid;
  return (int)(g < 1 ? 1 : g);
}

int launch_sample_stream(const MirtParams& p, int max_grid, cudaStream_t st) {
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
