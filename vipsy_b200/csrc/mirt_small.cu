// Guide-side kernels of the K > 1 path for SMALL latent dimension (2 <= K <= 32), any particle count S,
// per-person Cholesky factors (VIRT per-person covariance vi.py:717-723, MvnEncoder output vi.py:444-454;
// the reference's 2-, 3-, 5- and 10-dimensional settings, test.py:363-453):
//
//   mirt_sample_small    one THREAD per (person, particle): x = mu + L eps with the person's factor
//                        (<= 528 floats) in shared memory, noise straight from the Philox counter
//   mirt_cholgrad_small  one THREAD per output element of d L: sum_s g_s[i] eps_s[m], g and eps of the
//                        CTA's persons staged in shared memory; d mu, loss partials
//
// The large-K kernels (mirt_stream.cu) give a whole warp to one person; at K = 10 that leaves two thirds
// of the lanes idle and re-reads the factor per particle.  Here every lane has a full unit of work.
#include "mirt.cuh"

namespace vipsy {

constexpr int kSmThreads = 128;

// flattened noise layout of noise_coord() for K <= 32: element n = s*K + k lives in block n >> 2
template <int KMAX>
__device__ __forceinline__ void noise_row(const MirtParams& p, int s, int64_t b, float (&e)[KMAX]) {
  const int K = p.K;
  if (p.eps) {
    const float* src = p.eps + ((int64_t)s * p.B + b) * K;
#pragma unroll
    for (int k = 0; k < KMAX; ++k) e[k] = k < K ? src[k] : 0.f;
    return;
  }
  const uint32_t n0 = (uint32_t)s * (uint32_t)K;
  const uint32_t blk0 = n0 >> 2;
  const int sh = (int)(n0 & 3u);                 // element k sits at slot sh + k of the concatenated blocks
#pragma unroll
  for (int q = 0; q < KMAX / 4 + 1; ++q) {
    if (q * 4 < sh + K) {
      const float4 z = normal4(p.seed, p.offset, (uint64_t)(p.person_offset + b), blk0 + q);
      const float zz[4] = {z.x, z.y, z.z, z.w};
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        // slot q*4 + c  ->  k = q*4 + c - sh
#pragma unroll
        for (int d = 0; d < 4; ++d) {
          const int k = q * 4 + c - d;
          if (d == sh && k >= 0 && k < KMAX) e[k] = zz[c];
        }
      }
    }
  }
#pragma unroll
  for (int k = 0; k < KMAX; ++k)
    if (k >= K) e[k] = 0.f;
}

template <int KMAX>
__global__ void __launch_bounds__(kSmThreads) mirt_sample_small_kernel(const MirtParams p, const int PB) {
  extern __shared__ __align__(16) float smf[];
  const int K = p.K, S = p.S, tri = tri_off(K), trp = tri | 1;      // odd stride: persons land on distinct banks
  float* Ls = smf;                                                  // [PB][trp]
  float* red = Ls + (size_t)PB * trp;                               // [PB][S]   per-(person, particle) aux terms
  float* lsum = red + (size_t)PB * S;                               // [PB]      sum_i log L_ii
  const bool packed = p.chol_layout == VIPSY_CHOL_PACKED;
  const int64_t stride_L = packed ? tri : (int64_t)K * K;
  const int tid = threadIdx.x;
  const int64_t npass = (p.B + PB - 1) / PB;

  for (int64_t pass = blockIdx.x; pass < npass; pass += gridDim.x) {
    const int64_t b0 = pass * PB;
    const int cnt = (int)min((int64_t)PB, p.B - b0);
    __syncthreads();
    if (packed) {
      const float* src = p.chol + b0 * tri;                         // cnt persons are contiguous: coalesced
      for (int e = tid; e < cnt * tri; e += kSmThreads) {
        const int pb = e / tri, off = e - pb * tri;
        Ls[pb * trp + off] = src[e];
      }
    } else {
      for (int e = tid; e < cnt * K * K; e += kSmThreads) {
        const int pb = e / (K * K), r = e - pb * K * K, i = r / K, m = r - i * K;
        if (m <= i) Ls[pb * trp + tri_off(i) + m] = p.chol[(b0 + pb) * stride_L + r];
      }
    }
    __syncthreads();
    if (tid < cnt) {                                                // diagonal: exp (raw), diag L out, log-det
      float ls = 0.f;
      for (int i = 0; i < K; ++i) {
        float v = Ls[tid * trp + tri_off(i) + i];
        if (p.chol_raw) v = __expf(v);
        Ls[tid * trp + tri_off(i) + i] = v;
        p.diagL[(b0 + tid) * K + i] = v;
        ls += __logf(v);
      }
      lsum[tid] = ls;
    }
    __syncthreads();
    for (int w = tid; w < cnt * S; w += kSmThreads) {               // one (person, particle) per thread
      const int pb = w / S, s = w - pb * S;
      const int64_t b = b0 + pb;
      float e[KMAX];
      noise_row<KMAX>(p, s, b, e);
      const float* L = Ls + pb * trp;
      float auxp = 0.f;
      float* xo = p.x + ((int64_t)s * p.B + b) * K;
      const float* mu = p.mu + b * K;
#pragma unroll
      for (int i = 0; i < KMAX; ++i) {
        if (i < K) {
          float acc = mu[i];
#pragma unroll
          for (int m = 0; m <= i; ++m) acc = fmaf(L[tri_off(i) + m], e[m], acc);
          xo[i] = acc;
          auxp = fmaf(0.5f * e[i], e[i], auxp);
          auxp = fmaf(-0.5f * acc, acc, auxp);
        }
      }
      red[pb * S + s] = auxp;
    }
    __syncthreads();
    if (tid < cnt) {                                                // fixed-order sum over particles
      float t = (float)S * lsum[tid];
      for (int s = 0; s < S; ++s) t += red[tid * S + s];
      *reinterpret_cast<float4*>(p.aux + 4 * (b0 + tid)) = make_float4(t, 0.f, 0.f, 0.f);
    }
  }
}

__global__ void __launch_bounds__(kSmThreads) mirt_cholgrad_small_kernel(const MirtParams p, const int PB) {
  extern __shared__ __align__(16) float smf[];
  __shared__ double loss_sm[kSmThreads / 32];
  __shared__ unsigned char row_of[528], col_of[528];
  const int K = p.K, S = p.S, tri = tri_off(K);
  const int SK = S * K, SKp = SK | 1;
  float* gs = smf;                                  // [PB][SKp]  kf * (d log-lik / dx - x)
  float* es = gs + (size_t)PB * SKp;                // [PB][SKp]  noise
  const bool packed = p.chol_layout == VIPSY_CHOL_PACKED;
  const int64_t stride_L = packed ? tri : (int64_t)K * K;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float kf = -p.scale / (float)S;
  const int64_t Rrows = (int64_t)S * p.B;
  const int64_t npass = (p.B + PB - 1) / PB;
  double loss_acc = 0.0;
  for (int e = tid; e < tri; e += kSmThreads) {
    int i = (int)((sqrtf(8.f * (float)e + 1.f) - 1.f) * 0.5f);
    while (tri_off(i + 1) <= e) ++i;
    while (tri_off(i) > e) --i;
    row_of[e] = (unsigned char)i;
    col_of[e] = (unsigned char)(e - tri_off(i));
  }

  for (int64_t pass = blockIdx.x; pass < npass; pass += gridDim.x) {
    const int64_t b0 = pass * PB;
    const int cnt = (int)min((int64_t)PB, p.B - b0);
    __syncthreads();
    // ---- stage g and eps of the pass's persons --------------------------------------------------------
    for (int e = tid; e < cnt * SK; e += kSmThreads) {
      const int pb = e / SK, r = e - pb * SK, s = r / K, k = r - s * K;
      const int64_t at = ((int64_t)s * p.B + b0 + pb) * K + k;
      float g = -p.x[at];                                           // d log N(x; 0, I) / dx
      for (int part = 0; part < p.n_parts; ++part) g += p.gx[(int64_t)part * Rrows * K + at];
      gs[pb * SKp + r] = kf * g;
      if (p.eps) es[pb * SKp + r] = p.eps[at];
    }
    if (!p.eps) {
      const int nblk = (SK + 3) >> 2;
      for (int q = tid; q < cnt * nblk; q += kSmThreads) {
        const int pb = q / nblk, blk = q - pb * nblk;
        const float4 z = normal4(p.seed, p.offset, (uint64_t)(p.person_offset + b0 + pb), (uint32_t)blk);
        const float zz[4] = {z.x, z.y, z.z, z.w};
#pragma unroll
        for (int c = 0; c < 4; ++c)
          if (blk * 4 + c < SK) es[pb * SKp + blk * 4 + c] = zz[c];
      }
    }
    __syncthreads();
    // ---- d mu ------------------------------------------------------------------------------------------
    if (p.d_mu)
      for (int e = tid; e < cnt * K; e += kSmThreads) {
        const int pb = e / K, k = e - pb * K;
        float t = 0.f;
        for (int s = 0; s < S; ++s) t += gs[pb * SKp + s * K + k];
        p.d_mu[(b0 + pb) * K + k] = t;
      }
    // ---- d L: one output element per thread --------------------------------------------------------------
    if (p.d_chol)
      for (int e = tid; e < cnt * tri; e += kSmThreads) {
        const int pb = e / tri, off = e - pb * tri;
        const int i = row_of[off], m = col_of[off];
        const float* g = gs + pb * SKp + i;
        const float* z = es + pb * SKp + m;
        float v = 0.f;
        for (int s = 0; s < S; ++s) v = fmaf(g[s * K], z[s * K], v);
        if (m == i) {
          const float lii = p.diagL[(b0 + pb) * K + i];
          v -= p.scale / lii;                                       // - scale d(sum log L_ii)/dL_ii
          if (p.chol_raw) v *= lii;                                 // LowerCholeskyTransform chain rule
        }
        p.d_chol[(b0 + pb) * stride_L + (packed ? off : i * K + m)] = v;
      }
    // ---- loss terms ----------------------------------------------------------------------------------------
    if (tid < cnt) {
      const int64_t b = b0 + tid;
      float ellsum = 0.f;
      for (int s = 0; s < S; ++s)
        for (int part = 0; part < p.n_parts; ++part) ellsum += p.ell[(int64_t)part * Rrows + (int64_t)s * p.B + b];
      const float4 ax = *reinterpret_cast<const float4*>(p.aux + 4 * b);
      loss_acc += (double)(kf * (ellsum + ((ax.x + ax.y) + (ax.z + ax.w))));
    }
  }
  loss_acc = warp_sum(loss_acc);
  if (lane == 0) loss_sm[warp] = loss_acc;
  __syncthreads();
  if (tid == 0) {
    double t = 0.0;
    for (int w = 0; w < kSmThreads / 32; ++w) t += loss_sm[w];
    p.loss_partial[blockIdx.x] = t;
  }
}

// ----------------------------------------------------------------------------------------------------------
bool mirt_small_supported(const MirtParams& p) {
  return p.cov_mode == VIPSY_COV_PERSON && p.K >= 2 && p.K <= 32 && p.S <= 4096 && (int64_t)p.S * p.K <= 24000;
}

static int small_pb_sample(const MirtParams& p) {
  int pb = kSmThreads / p.S;
  if (pb < 1) pb = 1;
  if (pb > 64) pb = 64;
  if ((int64_t)pb > p.B) pb = (int)p.B;
  return pb < 1 ? 1 : pb;
}
static int small_pb_cholgrad(const MirtParams& p) {
  int pb = 6144 / (2 * p.S * p.K);                  // <= 48 KB of staged g + eps
  if (pb < 1) pb = 1;
  if (pb > 64) pb = 64;
  if ((int64_t)pb > p.B) pb = (int)p.B;
  return pb < 1 ? 1 : pb;
}

int launch_sample_small(const MirtParams& p, int max_grid, cudaStream_t st) {
  const int PB = small_pb_sample(p), tri = p.K * (p.K + 1) / 2;
  const size_t smem = ((size_t)PB * (tri | 1) + (size_t)PB * p.S + PB) * 4;
  int64_t g = (p.B + PB - 1) / PB;
  if (g > max_grid) g = max_grid;
  auto launch = [&](auto kern) -> int {
    if (smem > 48 * 1024) VIPSY_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<(int)g, kSmThreads, smem, st>>>(p, PB);
    VIPSY_CUDA_CHECK(cudaGetLastError());
    return VIPSY_OK;
  };
  if (p.K <= 4) return launch(mirt_sample_small_kernel<4>);
  if (p.K <= 8) return launch(mirt_sample_small_kernel<8>);
  if (p.K <= 16) return launch(mirt_sample_small_kernel<16>);
  return launch(mirt_sample_small_kernel<32>);
}

int launch_cholgrad_small(const MirtParams& p, int max_grid, int* grid_out, cudaStream_t st) {
  const int PB = small_pb_cholgrad(p);
  const size_t smem = (size_t)2 * PB * ((p.S * p.K) | 1) * 4;
  int64_t g = (p.B + PB - 1) / PB;
  if (g > max_grid) g = max_grid;
  if (smem > 48 * 1024)
    VIPSY_CUDA_CHECK(cudaFuncSetAttribute(mirt_cholgrad_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  mirt_cholgrad_small_kernel<<<(int)g, kSmThreads, smem, st>>>(p, PB);
  VIPSY_CUDA_CHECK(cudaGetLastError());
  *grid_out = (int)g;
  return VIPSY_OK;
}

}  // namespace vipsy
