// Shared declarations of the K > 1 kernels (mirt.cu: sampling, SIMT contraction, guide gradients;
// mirt_tc.cu: tcgen05 contraction).
#pragma once
#include "common.cuh"

namespace vipsy {

// ----------------------------------------------------------------------------------------
struct MirtParams {
  const uint8_t* y; int64_t ld_y; const int64_t* idx;
  const float *mu, *chol, *eps;
  const float *a, *b, *c, *d;
  int64_t B;
  int J, K, S, model, cov_mode, chol_layout, chol_raw;
  float D, scale;
  int score;            // 1: score-function (REINFORCE) estimator for the guide parameters (fused K <= 8 kernel only)
  float baseline;       // its fixed control variate
  uint64_t seed, offset; int64_t person_offset;
  // workspace
  float* x;        // [S*B][K]
  float* gx;       // [n_parts][S*B][K]  partial d(sum_j lb)/dx, one plane per item tile of the contraction
  float* ell;      // [n_parts][S*B]     partial sum_j lb
  float* aux;      // [B][4]     S sum_i log L_ii + 0.5 sum_s (|eps_s|^2 - |x_s|^2), split over up to 4 row blocks
  int dbg;         // development switches (VIPSY_DBG), 0 in production
  int n_parts;     // planes of gx / ell the contraction kernel wrote (1 for the SIMT kernel)
  float* diagL;    // [B][K]
  float* item_partial;   // [grid2][K*J + 3J]   (d_a | d_b | d_c | d_d)
  float* chol_partial;   // [grid3][tri]        (shared covariance only)
  double* loss_partial;  // [grid3]
  // outputs
  float *d_mu, *d_chol;
};

__device__ __forceinline__ int tri_off(int i) { return (i * (i + 1)) >> 1; }

__device__ __forceinline__ float noise_at(const MirtParams& p, int s, int64_t b, int k) {
  if (p.eps) return p.eps[((int64_t)s * p.B + b) * p.K + k];
  uint32_t block;
  int comp;
  noise_coord(s, k, p.K, block, comp);
  const float4 z = normal4(p.seed, p.offset, (uint64_t)(p.person_offset + b), block);
  return comp == 0 ? z.x : (comp == 1 ? z.y : (comp == 2 ? z.z : z.w));
}

// N(0,1) draws of the dimensions lane, lane+32, lane+64, lane+96 of person b, particle s (0 beyond K):
// one Philox call per lane thanks to noise_coord()'s K > 1 layout.
__device__ __forceinline__ void noise_lane4(const MirtParams& p, int s, int64_t b, int lane, float (&z)[4]) {
  if (p.eps) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int m = lane + 32 * c;
      z[c] = m < p.K ? p.eps[((int64_t)s * p.B + b) * p.K + m] : 0.f;
    }
  } else {
    const float4 v = normal4(p.seed, p.offset, (uint64_t)(p.person_offset + b), (uint32_t)s * 32u + (uint32_t)lane);
    z[0] = lane < p.K ? v.x : 0.f;
    z[1] = lane + 32 < p.K ? v.y : 0.f;
    z[2] = lane + 64 < p.K ? v.z : 0.f;
    z[3] = lane + 96 < p.K ? v.w : 0.f;
  }
}

// K <= 8, per-person factor: the whole evaluation in one kernel (mirt_fused.cu); writes item_partial /
// loss_partial rows [0, *grid_out), d_mu, d_chol.
bool mirt_fused_supported(const MirtParams& p);
int launch_mirt_fused(const MirtParams& p, int max_grid, int* grid_out, cudaStream_t st);

// tcgen05 likelihood contraction (mirt_tc.cu).  rows_out = partial rows written to item_partial,
// parts_out = number of d x / log-lik partial planes (one per 128-item tile).
bool mirt_tc_supported(const MirtParams& p);
int mirt_tc_parts(int J);
int launch_lik_tc(const MirtParams& p, int* rows_out, int* parts_out, cudaStream_t st);


// streaming sampler / guide-gradient kernels for per-person factors, S <= 2 (mirt_stream.cu)
bool mirt_stream_supported(const MirtParams& p);
int launch_sample_stream(const MirtParams& p, int max_grid, cudaStream_t st);
int launch_cholgrad_stream(const MirtParams& p, int max_grid, int* grid_out, cudaStream_t st);

// The whole K > 1 evaluation (mirt.cu).  enc != null: the guide is the amortised one of enc_mvn.cuh -- the sampler
// and the guide-gradient kernels are replaced by the fused fc22 kernels, a->chol / d_chol are not used.
struct MvnEncParams;
size_t mirt_workspace_bytes(const VipsyIrtArgs* a);
int mirt_launch(const VipsyIrtArgs* a, cudaStream_t st, MvnEncParams* enc);

// thread-per-(person, particle) / thread-per-element kernels for K <= 32, any S (mirt_small.cu)
bool mirt_small_supported(const MirtParams& p);
int launch_sample_small(const MirtParams& p, int max_grid, cudaStream_t st);
int launch_cholgrad_small(const MirtParams& p, int max_grid, int* grid_out, cudaStream_t st);

}  // namespace vipsy
