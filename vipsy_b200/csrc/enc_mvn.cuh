// Amortised multivariate-normal guide of VaeIRT for K > 1 (vi.py:685-693) with the tail of MvnEncoder
// (fc22 + the scatter into a lower-triangular matrix, vi.py:444, 449-455) fused into the sampler and its
// backward pass (enc_mvn.cu): the B x K(K+1)/2 factor matrix -- 4 GB for 200 000 persons at K = 100 -- and its
// gradient never exist in HBM.
#pragma once
#include "mirt.cuh"

namespace vipsy {

constexpr int kMvTP = 128;            // persons per tile (UMMA M of the forward product)
constexpr int kMvCN = 128;            // padded factor columns per chunk (UMMA N of the forward product)
constexpr int kMvMaxH = 64;           // hidden units (the reference's default hidden_dim, vi.py:659)
constexpr float kMvWScale = 64.f;     // power-of-two prescale of fc22.weight / bias before the fp16 hi/lo split

// Padded column order of the factor.  L is cut into 8 x 8 blocks; block row I (rows 8I .. 8I+7) owns the units
// (I, Jc), Jc = 0 .. I, of 64 columns each, and inside a unit the order is (row r, column jj):
//     (i, j) -> 64 (I (I + 1) / 2 + Jc) + 8 r + jj,   I = i / 8, r = i % 8, Jc = j / 8, jj = j % 8.
// Eight consecutive columns are eight consecutive j of one row; a unit is one 8-vector of the noise against eight
// rows (8 independent accumulators per person, the noise read once per 64 products); the diagonal of L sits at
// the fixed slots 9 r of the units with Jc == I.  Entries with j > i (or i >= K) are padding with weight 0.
__host__ __device__ inline int mv_nb(int K) { return (K + 7) >> 3; }
__host__ __device__ inline int mv_pp(int K) { return 32 * mv_nb(K) * (mv_nb(K) + 1); }
__host__ __device__ inline int mv_col(int i, int j) {
  const int I = i >> 3, Jc = j >> 3;
  return 64 * (I * (I + 1) / 2 + Jc) + 8 * (i & 7) + (j & 7);
}
// fc21 (the location head) rides in the padding above the diagonal of the diagonal units: row r of block row I
// (location 8 I + r) sits at local column mv_loc_slot(r) of unit (I, I) -- (0, r) for r >= 1, (1, 2) for r = 0 --
// as a factor column that meets a constant draw of 1.
__host__ __device__ inline int mv_loc_slot(int r) { return r == 0 ? 10 : r; }
__host__ __device__ inline int mv_loc_row_of_slot(int idx) { return idx == 10 ? 0 : (idx >= 1 && idx <= 7 ? idx : -1); }
__host__ __device__ inline int mv_kh(int H) { return (H + 1 + 15) / 16 * 16; }   // + the ones column (fc22.bias)
__host__ __device__ inline int mv_nh(int H) { return (H + 15) / 16 * 16; }
__host__ __device__ inline int mv_chunks(int K) { return (mv_pp(K) + kMvCN - 1) / kMvCN; }
// one chunk of the operand image: fp16 "hi" tile then "lo" tile, each 128 padded columns x KH in core-matrix layout
__host__ __device__ inline size_t mv_chunk_bytes(int H) { return (size_t)2 * kMvCN * mv_kh(H) * 2; }

struct MvnEncParams {
  MirtParams m;              // B, K, S, mu, eps (injected) / seed.., x, aux, diagL, gx, ell, loss_partial, d_mu
  const float* hidden;       // [B][H]
  const float* w22;          // [P][H]
  const float* b22;          // [P]
  const float* w21;          // [K][H] fc21.weight, or null: the location comes in as m.mu
  const float* b21;          // [K]
  int H;
  unsigned char* wimg;       // [chunks][hi | lo] operand image of 64 * [fc22.weight | fc22.bias | 0]
  unsigned char* wimg_b;     // [units][hi | lo] 64 padded columns x NH of 64 * fc22.weight (backward: d hidden), or null
  unsigned char* himg;       // [ceil(B / 64)][hi | lo] 64 persons x KH of [hidden | 1 | 0] (backward: d fc22), or null
  unsigned char* vimg;       // [S][ceil(B / 64)] operand rows of the d fc22 generator (mvn_dwprep_kernel), or null
  uint32_t* eps_split;       // [S*B + pad][K] the draws as fp16 (hi, lo) pair words (K % 4 == 0; see split_pair), or null
  float* eps_out;            // [S*B][K] the draws (for the backward pass and for parity tests), or null
  float* logdet_out;         // [B] sum_i log L_ii, or null
  float* dw_part;            // [n_qg][chunk][KH][128 columns of the chunk]
  float* gmax;               // device scalar: max |G| of the step (power-of-two normalisation of the fp16 operand V)
  float* d_hidden;           // outputs
  float* d_w22;
  float* d_b22;
  float* d_w21;              // [K][H], or null
  float* d_b21;              // [K], or null
};

size_t mvn_enc_extra_bytes(const VipsyMvnEncoderIrtArgs* a);
int launch_mvn_prep(const MvnEncParams& P, cudaStream_t st);
int launch_mvn_fwd(const MvnEncParams& P, cudaStream_t st);
int launch_mvn_tail(const MvnEncParams& P, int* loss_rows, cudaStream_t st);
int launch_mvn_bwd(const MvnEncParams& P, cudaStream_t st);
size_t mvn_wimg_bytes(int K, int H);
size_t mvn_wimgb_bytes(int K, int H);
size_t mvn_dwpart_bytes(int K, int H);
size_t mvn_himg_bytes(int64_t B, int H);
size_t mvn_vimg_bytes(int64_t B, int K, int S);
size_t mvn_eps_bytes(int64_t B, int K, int S);

}  // namespace vipsy
