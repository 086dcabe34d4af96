// Shared pieces of the NormEncoder kernels (enc_norm.cu: SIMT, any shape; enc_norm_tc.cu: tcgen05, J <= 127).
#pragma once
#include "common.cuh"

namespace vipsy {

struct EncParams {
  const uint8_t* y; int64_t ld; const int64_t* idx; int64_t B; int J, H;
  const float *w1, *b1, *w21, *b21, *w22, *b22;
  float *mu, *sigma, *pre;
  const float *d_mu, *d_sigma; float gscale;
  float* partial; int cols;             // backward: [gridDim.x x cols]
  const unsigned char* xtiles;          // resident bf16 operand tiles or nullptr
};

// softplus (torch: beta 1, threshold 20) and its derivative from one exponential: three MUFU ops, absolute error
// ~1e-7 (beyond the threshold log2(1 + e) rounds to 0 and the value is x itself, as in torch)
__device__ __forceinline__ void softplus_parts(float x, float& sp, float& sg) {
  const float e = ex2_approx(-fabsf(x) * kLog2e);
  const float t = 1.0f + e;
  const float r = rcp_approx(t);
  sp = fmaf(lg2_approx(t), kLn2, fmaxf(x, 0.0f));
  sg = x >= 0.0f ? r : e * r;                      // softplus'(x) = sigmoid(x)
}

// tensor-core variants (enc_norm_tc.cu); the partial-row layout and the finalize step are shared
bool enc_tc_supported(int H, int J);
bool enc_tc_resident_supported(int H, int J);
size_t enc_tc_tile_bytes(int J);
int enc_tc_fwd_launch(const EncParams& P, cudaStream_t st);
int enc_tc_bwd_launch(const EncParams& P, int grid_x, cudaStream_t st);
int enc_tc_pack_launch(const uint8_t* y, int64_t n_rows, int J, int64_t ld, void* tiles, cudaStream_t st);

}  // namespace vipsy
