// REINFORCE cognitive diagnosis: Trace_ELBO over BaseCDM.model (vi.py:745-756) with the
// Bernoulli guide of VCDM / VaeCDM (vi.py:796-816) -- the score-function path the
// reference actually executes (SURVEY Finding 1, Appendix A.5).
//
// One warp per person.  The sampled attribute vector is a K-bit mask obtained with a
// single ballot (lane k draws u_k < attr_p_k); eta_j is then the AND/compare
// [mask & q_j == q_j] (DINA) or [|q_j| >= 2][mask & q_j != 0] (DINO as executed),
// lanes stride the items.  Item gradients only depend on how often each (eta, y)
// combination occurred, so they are accumulated as exact integer counts per lane and
// turned into d_g, d_s once per CTA: no floating-point reduction order to worry about.
// The prior is Bernoulli(probs=1.5) (sic, vi.py:753): log p(1) = log(1-eps), log p(0) = log(eps).
// No missing-data handling in the reference: a missing cell yields NaN, here too.
#include "common.cuh"

namespace vipsy {

constexpr int kScWarps = 4;
constexpr int kScThreads = kScWarps * 32;
constexpr int kScSlots = 8;          // J <= 256

struct CdmScoreParams {
  const uint8_t* y; int64_t ld_y; const int64_t* idx;
  const uint32_t* qmask;
  const float *g, *s, *attr_p, *u;
  int64_t B;
  int J, K, S, model;
  float scale, baseline;
  uint64_t seed, offset; int64_t person_offset;
  float* d_attr_p;
  float* partial;        // [grid][2J]
  double* loss_partial;
};

__global__ void __launch_bounds__(kScThreads) cdm_score_kernel(const CdmScoreParams p) {
  __shared__ double loss_sm[kScWarps];
  __shared__ int cnt_sm[kScWarps][4];
  extern __shared__ __align__(16) int smem_i[];      // [4 warps][4][J] counts at the end
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int J = p.J, K = p.K, nslots = (J + 31) >> 5;
  const bool dino = p.model == VIPSY_CDM_DINO;
  // per-lane item constants
  uint32_t qm[kScSlots];
  float l11[kScSlots], l10[kScSlots], l01[kScSlots], l00[kScSlots];   // lb(eta, y)
  int c11[kScSlots], c10[kScSlots], c01[kScSlots], c00[kScSlots];
#pragma unroll
  for (int k = 0; k < kScSlots; ++k) {
    const int j = k * 32 + lane;
    c11[k] = c10[k] = c01[k] = c00[k] = 0;
    qm[k] = 0; l11[k] = l10[k] = l01[k] = l00[k] = 0.f;
    if (k < nslots && j < J) {
      qm[k] = p.qmask[j] & (uint32_t)((K >= 32) ? 0xffffffffu : ((1u << K) - 1u));
      const float p1c = fminf(fmaxf(1.f - p.s[j], kEps32), kOneMinusEps32);
      const float p0c = fminf(fmaxf(p.g[j], kEps32), kOneMinusEps32);
      l11[k] = logf(p1c); l10[k] = log1pf(-p1c); l01[k] = logf(p0c); l00[k] = log1pf(-p0c);
    }
  }
  const float lp1 = log1pf(-kEps32), lp0 = kLogEps32;     // prior Bernoulli(probs=1.5) clamped
  const float invS = 1.f / (float)p.S;
  double loss_acc = 0.0;
  const int64_t gwarp = (int64_t)blockIdx.x * kScWarps + warp, nwarps = (int64_t)gridDim.x * kScWarps;
  const float qnan = __int_as_float(0x7fc00000);

  for (int64_t b = gwarp; b < p.B; b += nwarps) {
    const int64_t row = p.idx ? p.idx[b] : b;
    unsigned codes[kScSlots];
#pragma unroll
    for (int k = 0; k < kScSlots; ++k) {
      const int j = k * 32 + lane;
      codes[k] = (k < nslots && j < J) ? p.y[row * p.ld_y + j] : 254u;    // 254: padding lane
    }
    float phi = 0.5f, phic = 0.5f;
    if (lane < K) { phi = p.attr_p[b * K + lane]; phic = fminf(fmaxf(phi, kEps32), kOneMinusEps32); }
    const float lq1 = logf(phic), lq0 = log1pf(-phic);
    const float dq1 = (phi == phic) ? 1.f / phic : 0.f, dq0 = (phi == phic) ? -1.f / (1.f - phic) : 0.f;
    float dphi = 0.f, elbo_sum = 0.f;
    for (int sidx = 0; sidx < p.S; ++sidx) {
      float u = 2.f;
      if (lane < K) {
        if (p.u) u = p.u[((int64_t)sidx * p.B + b) * K + lane];
        else {
          const uint32_t n = (uint32_t)(sidx * K + lane);
          const uint4 r = philox_draw(p.seed, p.offset, (uint64_t)(p.person_offset + b), n >> 2);
          const uint32_t rr[4] = {r.x, r.y, r.z, r.w};
          u = u01(rr[n & 3]);
        }
      }
      const bool ak = (lane < K) && (u < phi);                  // torch.bernoulli: rand < p
      const uint32_t mask = __ballot_sync(0xffffffffu, ak);
      float part = (lane < K) ? ((ak ? lp1 : lp0) - (ak ? lq1 : lq0)) : 0.f;   // log p - log q
      float lqk = (lane < K) ? (ak ? lq1 : lq0) : 0.f;
      (void)lqk;
#pragma unroll
      for (int k = 0; k < kScSlots; ++k) {
        if (k < nslots) {
          const unsigned cd = codes[k];
          if (cd != 254u) {
            const uint32_t q = qm[k];
            const bool eta = dino ? (__popc(q) >= 2 && (mask & q) != 0u) : ((mask & q) == q);
            if (cd == 255u) part = qnan;
            else if (cd != 0u) { part += eta ? l11[k] : l01[k]; if (eta) ++c11[k]; else ++c01[k]; }
            else { part += eta ? l10[k] : l00[k]; if (eta) ++c10[k]; else ++c00[k]; }
          }
        }
      }
      const float elbo = p.scale * warp_sum(part);
      elbo_sum += elbo;
      const float w = elbo - p.baseline;
      dphi = fmaf(w, ak ? dq1 : dq0, dphi);
    }
    if (lane == 0) loss_acc -= (double)(elbo_sum * invS);
    if (p.d_attr_p && lane < K) p.d_attr_p[b * K + lane] = -invS * dphi;
  }
  loss_acc = warp_sum(loss_acc);
  if (lane == 0) loss_sm[warp] = loss_acc;
  // counts -> shared -> per-CTA float partial
  int* cw = smem_i + (size_t)warp * 4 * J;
#pragma unroll
  for (int k = 0; k < kScSlots; ++k) {
    const int j = k * 32 + lane;
    if (k < nslots && j < J) { cw[j] = c11[k]; cw[J + j] = c10[k]; cw[2 * J + j] = c01[k]; cw[3 * J + j] = c00[k]; }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < kScWarps; ++w) t += loss_sm[w];
    p.loss_partial[blockIdx.x] = t;
  }
  (void)cnt_sm;
  for (int j = threadIdx.x; j < J; j += kScThreads) {
    int n11 = 0, n10 = 0, n01 = 0, n00 = 0;
    for (int w = 0; w < kScWarps; ++w) {
      const int* c = smem_i + (size_t)w * 4 * J;
      n11 += c[j]; n10 += c[J + j]; n01 += c[2 * J + j]; n00 += c[3 * J + j];
    }
    const float p1 = 1.f - p.s[j], p0 = p.g[j];
    const float p1c = fminf(fmaxf(p1, kEps32), kOneMinusEps32), p0c = fminf(fmaxf(p0, kEps32), kOneMinusEps32);
    const float k1 = (p1 == p1c) ? 1.f : 0.f, k0 = (p0 == p0c) ? 1.f : 0.f;
    // d ll / d g = n01 / g - n00 / (1-g);   d ll / d s = -n11 / (1-s) + n10 / s
    p.partial[(size_t)blockIdx.x * 2 * J + j] = k0 * ((float)n01 / p0c - (float)n00 / (1.f - p0c));
    p.partial[(size_t)blockIdx.x * 2 * J + J + j] = k1 * (-(float)n11 / p1c + (float)n10 / (1.f - p1c));
  }
}

constexpr int kScMaxCtasPerSM = 8;

size_t cdm_score_workspace_bytes(const VipsyCdmScoreArgs* a) {
  const size_t rows = (size_t)num_sms() * kScMaxCtasPerSM;
  return ((rows * 8 + 15) & ~size_t(15)) + rows * 2 * (size_t)a->n_items * 4;
}

int cdm_score_launch(const VipsyCdmScoreArgs* a, cudaStream_t st) {
  CdmScoreParams p;
  p.y = a->y; p.ld_y = a->y_row_stride; p.idx = a->idx; p.qmask = a->qmask;
  p.g = a->g; p.s = a->s; p.attr_p = a->attr_p; p.u = a->u;
  p.B = a->batch; p.J = a->n_items; p.K = a->n_attrs; p.S = a->n_particles; p.model = a->model;
  p.scale = a->scale; p.baseline = a->baseline;
  p.seed = a->seed; p.offset = a->offset; p.person_offset = a->person_offset;
  p.d_attr_p = a->d_attr_p;
  const size_t rows = (size_t)num_sms() * kScMaxCtasPerSM;
  p.loss_partial = reinterpret_cast<double*>(a->workspace);
  p.partial = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(a->workspace) + ((rows * 8 + 15) & ~size_t(15)));
  int64_t ctas = (a->batch + kScWarps - 1) / kScWarps;
  if (ctas > (int64_t)rows) ctas = rows;
  if (ctas < 1) ctas = 1;
  const size_t smem = (size_t)kScWarps * 4 * p.J * 4;
  cdm_score_kernel<<<(int)ctas, kScThreads, smem, st>>>(p);
  VIPSY_CUDA_CHECK(cudaGetLastError());
  FinalizeSegs segs{};
  segs.n = 2;
  segs.out[0] = a->d_g; segs.cols[0] = p.J;
  segs.out[1] = a->d_s; segs.cols[1] = p.J;
  return launch_finalize(p.partial, (int)ctas, 2 * p.J, -a->scale / (float)p.S, segs, p.loss_partial, (int)ctas,
                         a->loss, st);
}

}  // namespace vipsy

using namespace vipsy;

extern "C" {

size_t vipsy_cdm_score_workspace_bytes(const VipsyCdmScoreArgs* a) {
  if (!a || a->n_items <= 0) return 0;
  return cdm_score_workspace_bytes(a);
}

int vipsy_cdm_score_elbo_grad(const VipsyCdmScoreArgs* a, void* stream) {
  VIPSY_REQUIRE(a != nullptr, VIPSY_ERR_ARG, "vipsy_cdm_score: args is NULL");
  VIPSY_REQUIRE(a->model == VIPSY_CDM_DINA || a->model == VIPSY_CDM_DINO, VIPSY_ERR_ARG,
                "vipsy_cdm_score: unknown model %d (CDM_FUN has dina, dino)", a->model);
  VIPSY_REQUIRE(a->batch >= 0 && a->n_items > 0 && a->n_attrs > 0 && a->n_particles > 0, VIPSY_ERR_ARG,
                "vipsy_cdm_score: bad sizes");
  if (a->batch == 0) {                       // nothing to evaluate: loss 0, gradients 0
    VIPSY_REQUIRE(a->loss, VIPSY_ERR_ARG, "vipsy_cdm_score: loss is required");
    cudaStream_t st = (cudaStream_t)stream;
    VIPSY_CUDA_CHECK(cudaMemsetAsync(a->loss, 0, 8, st));
    if (a->d_g) VIPSY_CUDA_CHECK(cudaMemsetAsync(a->d_g, 0, (size_t)a->n_items * 4, st));
    if (a->d_s) VIPSY_CUDA_CHECK(cudaMemsetAsync(a->d_s, 0, (size_t)a->n_items * 4, st));
    return VIPSY_OK;
  }
  VIPSY_REQUIRE(a->n_attrs <= 32 && a->n_items <= 32 * kScSlots, VIPSY_ERR_UNSUPPORTED,
                "vipsy_cdm_score: K=%d (<=32) J=%d (<=%d)", a->n_attrs, a->n_items, 32 * kScSlots);
  VIPSY_REQUIRE(a->y && a->qmask && a->g && a->s && a->attr_p && a->loss, VIPSY_ERR_ARG,
                "vipsy_cdm_score: y, qmask, g, s, attr_p, loss are required");
  VIPSY_REQUIRE(a->y_row_stride >= a->n_items, VIPSY_ERR_ARG, "vipsy_cdm_score: y_row_stride < J");
  VIPSY_REQUIRE(a->idx != nullptr || a->batch <= a->n_rows, VIPSY_ERR_ARG, "vipsy_cdm_score: batch > rows without idx");
  VIPSY_REQUIRE(a->workspace && a->workspace_bytes >= cdm_score_workspace_bytes(a), VIPSY_ERR_WORKSPACE,
                "vipsy_cdm_score: workspace too small");
  return cdm_score_launch(a, (cudaStream_t)stream);
}

}  // extern "C"
