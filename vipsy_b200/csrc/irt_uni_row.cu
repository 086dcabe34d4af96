// K == 1, ONE particle (the reference's default Trace_ELBO(num_particles=1), vi.py:627): row kernels.
//
// With S particles per person the lane-per-person kernel (irt_uni.cu) amortises its per-item warp
// butterfly over S evaluations; at S == 1 that butterfly and the shared-memory traffic are most of
// the instruction stream and the kernel runs at a few percent of the HBM roof.  Here the mapping is
// turned by 90 degrees: a WARP owns a person at a time, LANES own items.
//
//   dense  (irt_uni_row_kernel):  lane l owns the items of response words l, l+32, ... (4 items per
//          32-bit word).  Its item parameters AND its item-gradient accumulators live in registers for
//          the whole kernel: no shuffles, no shared memory and no atomics on the item side.  The only
//          cross-lane step is the per-person sum of d lb / d x, taken for four persons at once with the
//          transposing butterfly (7 shuffles per 4 persons).  Rows are read straight from global memory
//          (coalesced 128-byte segments), four persons ahead of the arithmetic.
//   sparse (irt_uni_rowsp_kernel, VIPSY_IRT_FLAG_SPARSE; BASELINE config 4, 90 % NaN): the warp first
//          compacts the person's observed cells with ballots into a small shared-memory list and then
//          evaluates the list one entry per lane: work is proportional to the OBSERVED cells while the
//          byte matrix is still streamed once at full width.  Within one person all entries are distinct
//          items, so the warp-private shared-memory item accumulators are updated with plain
//          read-modify-writes in a fixed order: deterministic, no atomics.
//
// Arithmetic per cell is the same formula class as irt_uni.cuh::item_evals (signed logit, clamp,
// zero gradient where the clamp fired; rational form for a 4PL that cannot reach the clamp).
// Replaces the same reference code as irt_uni.cu: vi.py:22-66, 574-595, 617-625, 677-684, 701-705.
#include "irt_uni.cuh"

namespace vipsy {

constexpr int kRowGroup = 4;          // persons whose d lb / d x sums are reduced together

// per-person variational inputs of a 32-person tile, one person per lane
struct RowTileLane {
  float mu, sg, eps, x;
  bool valid;
};
__device__ __forceinline__ RowTileLane row_tile_lane(const IrtUniParams& p, int64_t bi) {
  RowTileLane t;
  t.valid = bi < p.B;
  t.mu = t.valid ? p.mu[bi] : 0.f;
  t.sg = t.valid ? p.sigma[bi] : 1.f;
  if (p.eps) {
    t.eps = t.valid ? __ldg(p.eps + bi) : 0.f;
  } else {
    t.eps = normal4(p.seed, p.offset, (uint64_t)(p.person_offset + bi), 0u).x;   // particle 0 of irt_uni.cu's stream
  }
  t.x = fmaf(t.sg, t.eps, t.mu);
  return t;
}
// per-person epilogue (S == 1): loss terms other than the likelihood, d mu, d sigma
// n_masked: this lane's share of the tile's masked cells (each is Bernoulli(probs=0).log_prob(0) = log(1-eps))
__device__ __forceinline__ void row_tile_finish(const IrtUniParams& p, const RowTileLane& t, int64_t bi, float G,
                                                float Ltile, int n_masked, double& loss_acc) {
  float prior = (float)n_masked * kMaskedLogProb;
  if (t.valid) {
    prior += -0.5f * t.x * t.x + 0.5f * t.eps * t.eps + logf(t.sg);
    if (p.d_mu) {
      p.d_mu[bi] = -p.scale * (G - t.x);
      p.d_sigma[bi] = -p.scale * (G - t.x) * t.eps - p.scale / t.sg;
    }
  }
  loss_acc -= (double)(p.scale * fmaf(kLn2, Ltile, prior));
}

// Cells [4w, 4w+4) of a response row as one 32-bit word, whatever the row's byte alignment and length: an
// unaligned row takes two aligned loads and a funnel shift (the second only if it still holds bytes of this
// row); cells at or beyond J come back as 0xff = missing.  Requires 4 w < J.
__device__ __forceinline__ uint32_t row_word(const uint8_t* base, int w, int J) {
  const int nb = J - 4 * w;                                    // cells of this word inside the row
  const uintptr_t addr = reinterpret_cast<uintptr_t>(base) + 4 * (uintptr_t)w;
  const int m = (int)(addr & 3);
  const uint32_t* ap = reinterpret_cast<const uint32_t*>(addr - m);
  uint32_t word = __ldg(ap);
  if (m != 0) {
    const uint32_t hi = (m + min(nb, 4) > 4) ? __ldg(ap + 1) : 0xffffffffu;
    word = __funnelshift_r(word, hi, 8 * m);
  }
  if (nb < 4) word |= 0xffffffffu << (8 * nb);
  return word;
}

// every row of the matrix starts on a 4-byte boundary and holds whole words: plain 32-bit loads
__host__ __device__ inline bool rows_word_aligned(const IrtUniParams& p) {
  return ((reinterpret_cast<uintptr_t>(p.y) | (uintptr_t)p.ld_y | (uintptr_t)p.J) & 3) == 0;
}

template <int NP, bool ALIGNED>
__device__ __forceinline__ void row_load_group(const IrtUniParams& p, int64_t b_first, int lane, int nwords,
                                               uint32_t (&W)[kRowGroup][NP]) {
  if constexpr (ALIGNED) {                         // straight-line loads the compiler can hoist
#pragma unroll
    for (int i = 0; i < kRowGroup; ++i) {
      const int64_t b = b_first + i;
      const bool ok = b < p.B;
      int64_t row = b;
      if (ok && p.idx) row = __ldg(p.idx + b);
      const uint32_t* g = reinterpret_cast<const uint32_t*>(p.y + row * p.ld_y);
#pragma unroll
      for (int k = 0; k < NP; ++k) {
        const int w = lane + 32 * k;
        W[i][k] = (ok && w < nwords) ? __ldg(g + w) : 0xffffffffu;
      }
    }
    return;
  }
#pragma unroll
  for (int i = 0; i < kRowGroup; ++i) {
    const int64_t b = b_first + i;
    const bool ok = b < p.B;
    int64_t row = b;
    if (ok && p.idx) row = __ldg(p.idx + b);
    const uint8_t* g = p.y + row * p.ld_y;
#pragma unroll
    for (int k = 0; k < NP; ++k) {
      const int w = lane + 32 * k;
      W[i][k] = (ok && w < nwords) ? row_word(g, w, p.J) : 0xffffffffu;
    }
  }
}

// sum Gl[i] over the warp for the kRowGroup persons of group g; lane 4g+i keeps person i's total
__device__ __forceinline__ void row_group_reduce(const float (&Gl)[kRowGroup], int lane, int g, float& Gmine) {
  const float r = warp_sum_multi<kRowGroup>(Gl, lane);            // lane 8*i holds the total of person i
  const float v = __shfl_sync(0xffffffffu, r, (lane & 3) * 8);
  if ((lane >> 2) == g) Gmine = v;
}

// ---------------------------------------------------------------------------------------------------
// dense
// ---------------------------------------------------------------------------------------------------
// register budget: measured on B200, 5-6 CTAs/SM beat both the unconstrained allocation (3-4 CTAs) and the
// spilling 6-8 CTA one for the packed 1PL / 2PL path
constexpr int row_min_ctas(int MODEL, int NP) {
  if (MODEL > 2) return 1;
  return NP == 1 ? 6 : NP == 2 ? 5 : 3;
}
// ALIGNED: every row starts on a 4-byte boundary and holds whole words (plain 32-bit loads); otherwise row_word()
template <int MODEL, int NP, bool FAST, bool ALIGNED>
__global__ void __launch_bounds__(kUniThreads, row_min_ctas(MODEL, NP)) irt_uni_row_kernel(const IrtUniParams p) {
  constexpr int NQ = MODEL;
  constexpr int JP = NP * 128;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* acc_sm = reinterpret_cast<float*>(smem_raw);             // [warp][NQ][JP], used after the main loop
  const int J = p.J, nwords = (J + 3) >> 2;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

  // item parameters of this lane's 4*NP items
  float Ap[NP][4], Bp[NP][4], Cc[NP][4], Dm[NP][4], Om[NP][4], Gc[NP][4];
  int ok = 1;
#pragma unroll
  for (int k = 0; k < NP; ++k)
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int j = 4 * (lane + 32 * k) + c;
      const bool in = j < J;
      const float aj = (MODEL >= 2 && in) ? p.a[j] : (in ? 1.f : 0.f);
      const float bj = in ? p.b[j] : 0.f;
      const float cj = (MODEL >= 3 && in) ? p.c[j] : 0.f;
      const float dj = (MODEL == 4 && in) ? p.d[j] : 1.f;
      Ap[k][c] = -p.D * kLog2e * aj;
      Bp[k][c] = -p.D * kLog2e * bj;
      Cc[k][c] = cj;
      Dm[k][c] = dj - cj;
      Om[k][c] = 1.f - dj;
      Gc[k][c] = p.D * aj * (dj - cj);
      if (FAST && in) ok &= (cj >= 1e-3f) && (dj <= 1.f - 1e-3f) && (dj - cj >= 1e-3f);
    }
  const bool use_fast = FAST && __all_sync(0xffffffffu, ok);
  // packed (1PL / 2PL) path: halves of A', B', D a for the cell pairs (0,1), (2,3) of each word
  float2 Ah[NP][2], Bh[NP][2], Gh[NP][2], acc2[NP][2][NQ >= 2 ? 2 : 1];
#pragma unroll
  for (int k = 0; k < NP; ++k)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      Ah[k][h] = make_float2(0.5f * Ap[k][2 * h], 0.5f * Ap[k][2 * h + 1]);
      Bh[k][h] = make_float2(0.5f * Bp[k][2 * h], 0.5f * Bp[k][2 * h + 1]);
      Gh[k][h] = make_float2(0.5f * Gc[k][2 * h], 0.5f * Gc[k][2 * h + 1]);
      acc2[k][h][0] = make_float2(0.f, 0.f);
      if constexpr (NQ >= 2) acc2[k][h][1] = make_float2(0.f, 0.f);
    }

  float acc[NP][4][NQ];
#pragma unroll
  for (int k = 0; k < NP; ++k)
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
      for (int q = 0; q < NQ; ++q) acc[k][c][q] = 0.f;

  double loss_acc = 0.0;
  const int64_t ntiles = (p.B + 31) / 32;
  const int64_t gwarp = (int64_t)blockIdx.x * kUniWarps + warp;
  const int64_t nwarps = (int64_t)gridDim.x * kUniWarps;

  uint32_t Wc[kRowGroup][NP], Wn[kRowGroup][NP];
  if (gwarp < ntiles) row_load_group<NP, ALIGNED>(p, gwarp * 32, lane, nwords, Wc);

  for (int64_t tile = gwarp; tile < ntiles; tile += nwarps) {
    const int64_t b0 = tile * 32;
    const RowTileLane tl = row_tile_lane(p, b0 + lane);
    float Gmine = 0.f, Ltile = 0.f;
    // codes are 0 / 1 / 255: bit 1 of a byte marks a missing cell.  Padding words and out-of-range persons
    // read as missing too; lane 0 takes them off again.
    const int rows = (int)min((int64_t)32, p.B - b0);
    int n_half = 0;       // cells this lane saw as missing (incl. padding words and out-of-range persons)
#pragma unroll 1
    for (int g = 0; g < 32 / kRowGroup; ++g) {
      // next group's rows (next tile's first group after the last one); out-of-range persons read as missing
      const int64_t nb = g + 1 < 32 / kRowGroup ? b0 + (g + 1) * kRowGroup : (tile + nwarps) * 32;
      row_load_group<NP, ALIGNED>(p, nb, lane, nwords, Wn);
      float Gl[kRowGroup];
#pragma unroll
      for (int i = 0; i < kRowGroup; ++i) {
        const float x = __shfl_sync(0xffffffffu, tl.x, g * kRowGroup + i);
        float Gi = 0.f;
        if constexpr (MODEL <= 2) {
          // ---- 1PL / 2PL: two cells per instruction (FFMA2 / FMUL2 / FADD2), no per-cell selects ----------
          // sg2 = +2 (y = 1), -2 (y = 0), 0 (missing) is assembled from the response byte with integer ops:
          // the top byte of the float is 0x40 / 0xc0 / 0x00.  t = sg2 * (A'/2 x + B'/2) is the logit signed by
          // the response, and 0 for a missing cell, whose sp = 1/2 then multiplies the likelihood product by
          // exactly 2^-1 (added back through n_half) while its gradient sg2 * (1 - sp) vanishes by itself.
          const float2 xx = make_float2(x, x);
          float2 G2 = make_float2(0.f, 0.f);
#pragma unroll
          for (int k = 0; k < NP; ++k) {
            const uint32_t word = Wc[i][k];
            n_half += __popc(word & 0x02020202u);
            const uint32_t nw = ~word;
            const uint32_t T = ((nw << 7) & 0x80808080u) | (nw & 0x40404040u);
            float2 sg[2], t[2], sp[2], es[2];
            sg[0] = make_float2(__uint_as_float(__byte_perm(T, 0u, 0x0444u)), __uint_as_float(__byte_perm(T, 0u, 0x1444u)));
            sg[1] = make_float2(__uint_as_float(__byte_perm(T, 0u, 0x2444u)), __uint_as_float(__byte_perm(T, 0u, 0x3444u)));
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              t[h] = __fmul2_rn(sg[h], __ffma2_rn(Ah[k][h], xx, Bh[k][h]));
              const float2 e = make_float2(ex2_approx(fminf(t[h].x, 23.f)), ex2_approx(fminf(t[h].y, 23.f)));
              const float2 ope = __fadd2_rn(e, make_float2(1.f, 1.f));
              sp[h] = make_float2(rcp_approx(ope.x), rcp_approx(ope.y));
              es[h] = __fmul2_rn(e, sp[h]);
            }
            // probability clamp (sp < eps <=> t >= 23): value log(eps) comes from the min above, the gradient
            // is switched off here.  Rare, so one test per word.
            if (fmaxf(fmaxf(t[0].x, t[0].y), fmaxf(t[1].x, t[1].y)) >= 23.f) {
              if (t[0].x >= 23.f) es[0].x = 0.f;
              if (t[0].y >= 23.f) es[0].y = 0.f;
              if (t[1].x >= 23.f) es[1].x = 0.f;
              if (t[1].y >= 23.f) es[1].y = 0.f;
            }
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const float2 ws = __fmul2_rn(sg[h], es[h]);             // 2 * d lb / d logit (signed)
              G2 = __ffma2_rn(Gh[k][h], ws, G2);
              acc2[k][h][0] = __fadd2_rn(acc2[k][h][0], ws);
              if constexpr (NQ >= 2) acc2[k][h][1] = __ffma2_rn(ws, xx, acc2[k][h][1]);
            }
            const float2 pr = __fmul2_rn(sp[0], sp[1]);
            Ltile += lg2_approx(pr.x * pr.y);
          }
          Gi = G2.x + G2.y;
        } else {
#pragma unroll
        for (int k = 0; k < NP; ++k) {
          const uint32_t word = Wc[i][k];
          n_half += __popc(word & 0x02020202u);
          float prod = 1.f;
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const uint32_t code = (word >> (8 * c)) & 0xffu;
            const bool obs = code != 255u, y1 = code != 0u;
            const float t0 = fmaf(Ap[k][c], x, Bp[k][c]);
            const float t = y1 ? t0 : -t0;
            const float base = MODEL >= 3 ? (y1 ? Cc[k][c] : Om[k][c]) : 0.f;
            float w, a1, a2, qc;
            if (FAST && use_fast) cell_eval1<MODEL, true>(t, Dm[k][c], base, obs, w, a1, a2, qc);
            else cell_eval1<MODEL, false>(t, Dm[k][c], base, obs, w, a1, a2, qc);
            const float ws = y1 ? w : -w;
            Gi = fmaf(Gc[k][c], ws, Gi);
            acc[k][c][0] += ws;
            if constexpr (NQ >= 2) acc[k][c][1] = fmaf(ws, x, acc[k][c][1]);
            if constexpr (NQ >= 3) acc[k][c][2] += y1 ? a2 : -a1;
            if constexpr (NQ >= 4) acc[k][c][3] += y1 ? a1 : -a2;
            prod *= qc;
          }
          Ltile += lg2_approx(prod);
        }
        }
        Gl[i] = Gi;
      }
      row_group_reduce(Gl, lane, g, Gmine);
#pragma unroll
      for (int i = 0; i < kRowGroup; ++i)
#pragma unroll
        for (int k = 0; k < NP; ++k) Wc[i][k] = Wn[i][k];
    }
    if constexpr (MODEL <= 2) Ltile += (float)n_half;     // every missing cell put a factor 1/2 into the product
    row_tile_finish(p, tl, b0 + lane, Gmine, Ltile, n_half + (lane == 0 ? rows * J - 32 * JP : 0), loss_acc);
  }

  // ---- CTA partials -----------------------------------------------------------------------------------
  __shared__ double loss_sm[kUniWarps];
  loss_acc = warp_sum(loss_acc);
  if (lane == 0) loss_sm[warp] = loss_acc;
  if constexpr (MODEL <= 2) {
#pragma unroll
    for (int k = 0; k < NP; ++k)
#pragma unroll
      for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
          acc[k][2 * h][q] = 0.5f * acc2[k][h][q].x;
          acc[k][2 * h + 1][q] = 0.5f * acc2[k][h][q].y;
        }
  }
  if (p.want_grad) {
#pragma unroll
    for (int k = 0; k < NP; ++k)
#pragma unroll
      for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int q = 0; q < NQ; ++q) acc_sm[((size_t)warp * NQ + q) * JP + 4 * (lane + 32 * k) + c] = acc[k][c][q];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < kUniWarps; ++w) t += loss_sm[w];
    p.loss_partial[blockIdx.x] = t;
  }
  if (p.want_grad) {
    for (int i = threadIdx.x; i < NQ * J; i += kUniThreads) {
      const int q = i / J, j = i - q * J;
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < kUniWarps; ++w) t += acc_sm[((size_t)w * NQ + q) * JP + j];
      if (q < 2) {
        const float cj = MODEL >= 3 ? p.c[j] : 0.f, dj = MODEL == 4 ? p.d[j] : 1.f;
        t *= p.D * (dj - cj);
      }
      p.item_partial[(size_t)blockIdx.x * NQ * J + i] = t;
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// dense, J <= 64: a row is WP = 2..16 words, so 32 / WP persons share a warp step
// ---------------------------------------------------------------------------------------------------
// Same mapping as irt_uni_row_kernel with the lanes split into 32 / WP groups of WP lanes: lane l owns word
// l % WP (4 items, parameters and gradient sums in registers) and, in step t, works on person t * (32 / WP) +
// l / WP of the tile.  The per-person sum of d lb / d x needs log2(WP) shuffles inside the group; the item sums
// of the groups are combined once, at the end of the kernel.  Typical test lengths (J = 20..60; the reference's
// own data sets have J = 5 and 6) keep all 32 lanes busy this way.
template <int MODEL, int WP, bool FAST>
__global__ void __launch_bounds__(kUniThreads) irt_uni_rowsm_kernel(const IrtUniParams p) {
  constexpr int NQ = MODEL;
  constexpr int PPS = 32 / WP;          // persons per step
  constexpr int JP = 4 * WP;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* acc_sm = reinterpret_cast<float*>(smem_raw);             // [warp][NQ][JP], used after the main loop
  const int J = p.J, nwords = (J + 3) >> 2;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int wl = lane % WP, grp = lane / WP;

  float Ap[4], Bp[4], Cc[4], Dm[4], Om[4], Gc[4];
  int ok = 1;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const int j = 4 * wl + c;
    const bool in = j < J;
    const float aj = (MODEL >= 2 && in) ? p.a[j] : (in ? 1.f : 0.f);
    const float bj = in ? p.b[j] : 0.f;
    const float cj = (MODEL >= 3 && in) ? p.c[j] : 0.f;
    const float dj = (MODEL == 4 && in) ? p.d[j] : 1.f;
    Ap[c] = -p.D * kLog2e * aj;
    Bp[c] = -p.D * kLog2e * bj;
    Cc[c] = cj;
    Dm[c] = dj - cj;
    Om[c] = 1.f - dj;
    Gc[c] = p.D * aj * (dj - cj);
    if (FAST && in) ok &= (cj >= 1e-3f) && (dj <= 1.f - 1e-3f) && (dj - cj >= 1e-3f);
  }
  const bool use_fast = FAST && __all_sync(0xffffffffu, ok);
  float2 Ah[2], Bh[2], Gh[2], acc2[2][NQ >= 2 ? 2 : 1];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    Ah[h] = make_float2(0.5f * Ap[2 * h], 0.5f * Ap[2 * h + 1]);
    Bh[h] = make_float2(0.5f * Bp[2 * h], 0.5f * Bp[2 * h + 1]);
    Gh[h] = make_float2(0.5f * Gc[2 * h], 0.5f * Gc[2 * h + 1]);
    acc2[h][0] = make_float2(0.f, 0.f);
    if constexpr (NQ >= 2) acc2[h][1] = make_float2(0.f, 0.f);
  }
  float acc[4][NQ];
#pragma unroll
  for (int c = 0; c < 4; ++c)
#pragma unroll
    for (int q = 0; q < NQ; ++q) acc[c][q] = 0.f;

  double loss_acc = 0.0;
  const int64_t ntiles = (p.B + 31) / 32;
  const int64_t gwarp = (int64_t)blockIdx.x * kUniWarps + warp;
  const int64_t nwarps = (int64_t)gridDim.x * kUniWarps;

  const bool aligned = rows_word_aligned(p);
  for (int64_t tile = gwarp; tile < ntiles; tile += nwarps) {
    const int64_t b0 = tile * 32;
    const RowTileLane tl = row_tile_lane(p, b0 + lane);
    const int rows = (int)min((int64_t)32, p.B - b0);
    uint32_t W[WP];
#pragma unroll
    for (int t = 0; t < WP; ++t) {
      const int64_t b = b0 + t * PPS + grp;
      const bool okb = b < p.B && wl < nwords;
      int64_t row = b;
      if (okb && p.idx) row = __ldg(p.idx + b);
      const uint8_t* g = p.y + row * p.ld_y;
      W[t] = !okb ? 0xffffffffu : (aligned ? __ldg(reinterpret_cast<const uint32_t*>(g) + wl) : row_word(g, wl, J));
    }
    float Gmine = 0.f, Ltile = 0.f;
    int n_half = 0;
#pragma unroll
    for (int t = 0; t < WP; ++t) {
      const float x = __shfl_sync(0xffffffffu, tl.x, t * PPS + grp);
      const uint32_t word = W[t];
      n_half += __popc(word & 0x02020202u);
      float Gi = 0.f;
      if constexpr (MODEL <= 2) {
        // packed 1PL / 2PL cells, see irt_uni_row_kernel
        const float2 xx = make_float2(x, x);
        const uint32_t nw = ~word;
        const uint32_t T = ((nw << 7) & 0x80808080u) | (nw & 0x40404040u);
        float2 sg[2], tt[2], sp[2], es[2];
        sg[0] = make_float2(__uint_as_float(__byte_perm(T, 0u, 0x0444u)), __uint_as_float(__byte_perm(T, 0u, 0x1444u)));
        sg[1] = make_float2(__uint_as_float(__byte_perm(T, 0u, 0x2444u)), __uint_as_float(__byte_perm(T, 0u, 0x3444u)));
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          tt[h] = __fmul2_rn(sg[h], __ffma2_rn(Ah[h], xx, Bh[h]));
          const float2 e = make_float2(ex2_approx(fminf(tt[h].x, 23.f)), ex2_approx(fminf(tt[h].y, 23.f)));
          const float2 ope = __fadd2_rn(e, make_float2(1.f, 1.f));
          sp[h] = make_float2(rcp_approx(ope.x), rcp_approx(ope.y));
          es[h] = __fmul2_rn(e, sp[h]);
        }
        if (fmaxf(fmaxf(tt[0].x, tt[0].y), fmaxf(tt[1].x, tt[1].y)) >= 23.f) {
          if (tt[0].x >= 23.f) es[0].x = 0.f;
          if (tt[0].y >= 23.f) es[0].y = 0.f;
          if (tt[1].x >= 23.f) es[1].x = 0.f;
          if (tt[1].y >= 23.f) es[1].y = 0.f;
        }
        float2 G2 = make_float2(0.f, 0.f);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const float2 ws = __fmul2_rn(sg[h], es[h]);
          G2 = __ffma2_rn(Gh[h], ws, G2);
          acc2[h][0] = __fadd2_rn(acc2[h][0], ws);
          if constexpr (NQ >= 2) acc2[h][1] = __ffma2_rn(ws, xx, acc2[h][1]);
        }
        const float2 pr = __fmul2_rn(sp[0], sp[1]);
        Ltile += lg2_approx(pr.x * pr.y);
        Gi = G2.x + G2.y;
      } else {
        float prod = 1.f;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const uint32_t code = (word >> (8 * c)) & 0xffu;
          const bool obs = code != 255u, y1 = code != 0u;
          const float t0 = fmaf(Ap[c], x, Bp[c]);
          const float tv = y1 ? t0 : -t0;
          const float base = y1 ? Cc[c] : Om[c];
          float w, a1, a2, qc;
          if (FAST && use_fast) cell_eval1<MODEL, true>(tv, Dm[c], base, obs, w, a1, a2, qc);
          else cell_eval1<MODEL, false>(tv, Dm[c], base, obs, w, a1, a2, qc);
          const float ws = y1 ? w : -w;
          Gi = fmaf(Gc[c], ws, Gi);
          acc[c][0] += ws;
          acc[c][1] = fmaf(ws, x, acc[c][1]);
          if constexpr (NQ >= 3) acc[c][2] += y1 ? a2 : -a1;
          if constexpr (NQ >= 4) acc[c][3] += y1 ? a1 : -a2;
          prod *= qc;
        }
        Ltile += lg2_approx(prod);
      }
      // the person's sum over its WP lanes, then handed to the lane that owns the person (lane index = slot in tile)
#pragma unroll
      for (int o = WP / 2; o > 0; o >>= 1) Gi += __shfl_xor_sync(0xffffffffu, Gi, o);
      const float v = __shfl_sync(0xffffffffu, Gi, (lane % PPS) * WP);
      if (lane / PPS == t) Gmine = v;
    }
    if constexpr (MODEL <= 2) Ltile += (float)n_half;     // every missing cell put a factor 1/2 into the product
    row_tile_finish(p, tl, b0 + lane, Gmine, Ltile, n_half + (lane == 0 ? rows * J - 32 * JP : 0), loss_acc);
  }

  // ---- item sums: groups -> lane group 0, warps -> CTA partial row ----------------------------------------------
  if constexpr (MODEL <= 2) {
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
      for (int q = 0; q < NQ; ++q) {
        acc[2 * h][q] = 0.5f * acc2[h][q].x;
        acc[2 * h + 1][q] = 0.5f * acc2[h][q].y;
      }
  }
#pragma unroll
  for (int c = 0; c < 4; ++c)
#pragma unroll
    for (int q = 0; q < NQ; ++q)
#pragma unroll
      for (int o = WP; o < 32; o <<= 1) acc[c][q] += __shfl_xor_sync(0xffffffffu, acc[c][q], o);
  __shared__ double loss_sm[kUniWarps];
  loss_acc = warp_sum(loss_acc);
  if (lane == 0) loss_sm[warp] = loss_acc;
  if (p.want_grad && grp == 0) {
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
      for (int q = 0; q < NQ; ++q) acc_sm[((size_t)warp * NQ + q) * JP + 4 * wl + c] = acc[c][q];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < kUniWarps; ++w) t += loss_sm[w];
    p.loss_partial[blockIdx.x] = t;
  }
  if (p.want_grad) {
    for (int i = threadIdx.x; i < NQ * J; i += kUniThreads) {
      const int q = i / J, j = i - q * J;
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < kUniWarps; ++w) t += acc_sm[((size_t)w * NQ + q) * JP + j];
      if (q < 2) {
        const float cj = MODEL >= 3 ? p.c[j] : 0.f, dj = MODEL == 4 ? p.d[j] : 1.f;
        t *= p.D * (dj - cj);
      }
      p.item_partial[(size_t)blockIdx.x * NQ * J + i] = t;
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// sparse
// ---------------------------------------------------------------------------------------------------
__host__ __device__ inline size_t rowsp_smem_bytes(int J, int NQ) {
  size_t s = (size_t)J * 16 + align16((size_t)J * 8);
  s += (size_t)kUniWarps * align16((size_t)NQ * J * 4);
  s += (size_t)kUniWarps * align16((size_t)J * 2);
  return s;
}

template <int MODEL, int NP, bool FAST, bool ALIGNED>
__global__ void __launch_bounds__(kUniThreads) irt_uni_rowsp_kernel(const IrtUniParams p) {
  constexpr int NQ = MODEL;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int J = p.J, nwords = (J + 3) >> 2;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float4* tab = reinterpret_cast<float4*>(smem_raw);                                   // {A', B', c, d-c}
  float2* tab2 = reinterpret_cast<float2*>(smem_raw + (size_t)J * 16);                 // {D a (d-c), 1-d}
  unsigned char* q0 = reinterpret_cast<unsigned char*>(tab2) + align16((size_t)J * 8);
  const size_t acc_bytes = align16((size_t)NQ * J * 4);
  float* acc_all = reinterpret_cast<float*>(q0);
  float* acc = reinterpret_cast<float*>(q0 + (size_t)warp * acc_bytes);               // [NQ][J], warp-private
  uint16_t* list = reinterpret_cast<uint16_t*>(q0 + (size_t)kUniWarps * acc_bytes + (size_t)warp * align16((size_t)J * 2));

  int ok = 1;
  for (int j = threadIdx.x; j < J; j += kUniThreads) {
    const float aj = MODEL >= 2 ? p.a[j] : 1.f;
    const float cj = MODEL >= 3 ? p.c[j] : 0.f;
    const float dj = MODEL == 4 ? p.d[j] : 1.f;
    tab[j] = make_float4(-p.D * kLog2e * aj, -p.D * kLog2e * p.b[j], cj, dj - cj);
    tab2[j] = make_float2(p.D * aj * (dj - cj), 1.f - dj);
    if (FAST) ok &= (cj >= 1e-3f) && (dj <= 1.f - 1e-3f) && (dj - cj >= 1e-3f);
  }
  for (int i = threadIdx.x; i < (int)(kUniWarps * acc_bytes / 4); i += kUniThreads) acc_all[i] = 0.f;
  const bool use_fast = FAST ? (__syncthreads_and(ok) != 0) : (__syncthreads(), false);

  double loss_acc = 0.0;
  const int64_t ntiles = (p.B + 31) / 32;
  const int64_t gwarp = (int64_t)blockIdx.x * kUniWarps + warp;
  const int64_t nwarps = (int64_t)gridDim.x * kUniWarps;
  const uint32_t lt_mask = (1u << lane) - 1u;

  uint32_t Wc[kRowGroup][NP], Wn[kRowGroup][NP];
  if (gwarp < ntiles) row_load_group<NP, ALIGNED>(p, gwarp * 32, lane, nwords, Wc);

  for (int64_t tile = gwarp; tile < ntiles; tile += nwarps) {
    const int64_t b0 = tile * 32;
    const RowTileLane tl = row_tile_lane(p, b0 + lane);
    float Gmine = 0.f, Ltile = 0.f;
    const int rows = (int)min((int64_t)32, p.B - b0);
    int n_obs = 0;
#pragma unroll 1
    for (int g = 0; g < 32 / kRowGroup; ++g) {
      const int64_t nb = g + 1 < 32 / kRowGroup ? b0 + (g + 1) * kRowGroup : (tile + nwarps) * 32;
      row_load_group<NP, ALIGNED>(p, nb, lane, nwords, Wn);
      float Gl[kRowGroup];
#pragma unroll
      for (int i = 0; i < kRowGroup; ++i) {
        const float x = __shfl_sync(0xffffffffu, tl.x, g * kRowGroup + i);
        // ---- compact the observed cells of this person: entry = item | response << 15 ----------------
        int cnt = 0;
#pragma unroll
        for (int k = 0; k < NP; ++k) {
          const uint32_t word = Wc[i][k];
          if (!__any_sync(0xffffffffu, word != 0xffffffffu)) continue;
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const uint32_t code = (word >> (8 * c)) & 0xffu;
            const bool obs = code != 255u;
            const uint32_t m = __ballot_sync(0xffffffffu, obs);
            if (obs) list[cnt + __popc(m & lt_mask)] = (uint16_t)((4 * (lane + 32 * k) + c) | (code != 0u ? 0x8000u : 0u));
            cnt += __popc(m);
          }
        }
        n_obs += cnt;
        __syncwarp();
        // ---- one entry per lane ---------------------------------------------------------------------
        float Gi = 0.f;
        for (int t0 = 0; t0 < cnt; t0 += 32) {
          const bool obs = t0 + lane < cnt;
          const uint32_t ent = obs ? list[t0 + lane] : 0u;
          const int j = (int)(ent & 0x7fffu);
          const bool y1 = (ent & 0x8000u) != 0u;
          const float4 it = tab[j];
          const float2 it2 = tab2[j];
          const float tt0 = fmaf(it.x, x, it.y);
          const float t = y1 ? tt0 : -tt0;
          const float base = MODEL >= 3 ? (y1 ? it.z : it2.y) : 0.f;
          float w, a1, a2, qc;
          if (FAST && use_fast) cell_eval1<MODEL, true>(t, it.w, base, obs, w, a1, a2, qc);
          else cell_eval1<MODEL, false>(t, it.w, base, obs, w, a1, a2, qc);
          const float ws = y1 ? w : -w;
          Gi = fmaf(it2.x, ws, Gi);
          Ltile += lg2_approx(qc);
          if (obs && p.want_grad) {
            acc[j] += ws;
            if constexpr (NQ >= 2) acc[J + j] = fmaf(ws, x, acc[J + j]);
            if constexpr (NQ >= 3) acc[2 * J + j] += y1 ? a2 : -a1;
            if constexpr (NQ >= 4) acc[3 * J + j] += y1 ? a1 : -a2;
          }
        }
        __syncwarp();       // list is rewritten by the next person; acc updates ordered between persons
        Gl[i] = Gi;
      }
      row_group_reduce(Gl, lane, g, Gmine);
#pragma unroll
      for (int i = 0; i < kRowGroup; ++i)
#pragma unroll
        for (int k = 0; k < NP; ++k) Wc[i][k] = Wn[i][k];
    }
    row_tile_finish(p, tl, b0 + lane, Gmine, Ltile, lane == 0 ? rows * J - n_obs : 0, loss_acc);
  }

  __shared__ double loss_sm[kUniWarps];
  loss_acc = warp_sum(loss_acc);
  if (lane == 0) loss_sm[warp] = loss_acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < kUniWarps; ++w) t += loss_sm[w];
    p.loss_partial[blockIdx.x] = t;
  }
  if (p.want_grad) {
    for (int i = threadIdx.x; i < NQ * J; i += kUniThreads) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < kUniWarps; ++w) t += acc_all[(size_t)w * (acc_bytes / 4) + i];
      const int q = i / J, j = i - q * J;
      if (q < 2) t *= p.D * tab[j].w;
      p.item_partial[(size_t)blockIdx.x * NQ * J + i] = t;
    }
  }
}

// ---------------------------------------------------------------------------------------------------
template <typename Kern>
static int launch_row(Kern k, IrtUniParams& p, size_t smem, cudaStream_t st, int* grid_out, OccCache* occ_cache) {
  if (smem > 48 * 1024) VIPSY_CUDA_CHECK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int occ = 0;                          // capped at 8: workspace rows = num_sms * 8 (irt_uni_workspace_bytes)
  VIPSY_CUDA_CHECK(cached_occupancy(*occ_cache, k, kUniThreads, smem, 8, &occ));
  const int64_t ntiles = (p.B + 31) / 32;
  int64_t ctas = (ntiles + kUniWarps - 1) / kUniWarps;
  const int64_t cap = (int64_t)num_sms() * occ;
  if (ctas > cap) ctas = cap;
  if (ctas < 1) ctas = 1;
  *grid_out = (int)ctas;
  k<<<(int)ctas, kUniThreads, smem, st>>>(p);
  VIPSY_CUDA_CHECK(cudaGetLastError());
  return VIPSY_OK;
}

template <int MODEL, int NP, bool FAST>
static int launch_row_dense(IrtUniParams& p, cudaStream_t st, int* grid) {
  static OccCache occ, occ_u;
  const size_t smem = (size_t)kUniWarps * MODEL * NP * 128 * 4;
  if (!rows_word_aligned(p))
    return launch_row(irt_uni_row_kernel<MODEL, NP, FAST, false>, p, smem, st, grid, &occ_u);
  return launch_row(irt_uni_row_kernel<MODEL, NP, FAST, true>, p, smem, st, grid, &occ);
}
template <int MODEL, int NP, bool FAST>
static int launch_row_sparse(IrtUniParams& p, cudaStream_t st, int* grid) {
  static OccCache occ, occ_u;
  if (!rows_word_aligned(p))
    return launch_row(irt_uni_rowsp_kernel<MODEL, NP, FAST, false>, p, rowsp_smem_bytes(p.J, MODEL), st, grid, &occ_u);
  return launch_row(irt_uni_rowsp_kernel<MODEL, NP, FAST, true>, p, rowsp_smem_bytes(p.J, MODEL), st, grid, &occ);
}

template <int MODEL, int WP, bool FAST>
static int launch_row_small(IrtUniParams& p, cudaStream_t st, int* grid) {
  static OccCache occ;
  const size_t smem = (size_t)kUniWarps * MODEL * 4 * WP * 4;
  return launch_row(irt_uni_rowsm_kernel<MODEL, WP, FAST>, p, smem, st, grid, &occ);
}

template <int MODEL, bool FAST>
static int launch_row_np(bool sparse, int NP, IrtUniParams& p, cudaStream_t st, int* grid) {
  if (!sparse && p.J <= 64) {                       // several persons per warp step
    const int nw = (p.J + 3) >> 2;
    if (nw <= 2) return launch_row_small<MODEL, 2, FAST>(p, st, grid);
    if (nw <= 4) return launch_row_small<MODEL, 4, FAST>(p, st, grid);
    if (nw <= 8) return launch_row_small<MODEL, 8, FAST>(p, st, grid);
    return launch_row_small<MODEL, 16, FAST>(p, st, grid);
  }
  if (sparse) {
    switch (NP) {
      case 1: return launch_row_sparse<MODEL, 1, FAST>(p, st, grid);
      case 2: return launch_row_sparse<MODEL, 2, FAST>(p, st, grid);
      case 3: return launch_row_sparse<MODEL, 3, FAST>(p, st, grid);
      default: return launch_row_sparse<MODEL, 4, FAST>(p, st, grid);
    }
  }
  if constexpr (MODEL <= 2) {
    switch (NP) {
      case 1: return launch_row_dense<MODEL, 1, FAST>(p, st, grid);
      case 2: return launch_row_dense<MODEL, 2, FAST>(p, st, grid);
      case 3: return launch_row_dense<MODEL, 3, FAST>(p, st, grid);
      default: return launch_row_dense<MODEL, 4, FAST>(p, st, grid);
    }
  } else {
    return NP == 1 ? launch_row_dense<MODEL, 1, FAST>(p, st, grid) : launch_row_dense<MODEL, 2, FAST>(p, st, grid);
  }
}

// Does the row mapping cover this call?  S == 1 pathwise; any row alignment and length (row_word); few enough items
// for the register-resident (dense) / 4-word-prefetch (sparse) layouts.  Short rows (J <= 64) take the
// several-persons-per-step kernel; the compaction form needs rows long enough to be worth compacting.
bool irt_uni_row_supported(const VipsyIrtArgs* a, bool sparse) {
  if (a->n_particles != 1 || a->estimator != VIPSY_EST_PATHWISE) return false;
  const int J = a->n_items;
  if (sparse && J > 64) return J <= 512 && rowsp_smem_bytes(J, a->model) <= 200 * 1024;
  return J <= (a->model <= 2 ? 512 : 256);
}

int irt_uni_row_launch(int model, bool fast, bool sparse, IrtUniParams& p, cudaStream_t st, int* grid_out) {
  const int NP = (p.J + 127) / 128;
  if (p.J <= 64) sparse = false;
  switch (model) {
    case VIPSY_IRT_1PL: return launch_row_np<1, false>(sparse, NP, p, st, grid_out);
    case VIPSY_IRT_2PL: return launch_row_np<2, false>(sparse, NP, p, st, grid_out);
    case VIPSY_IRT_3PL: return launch_row_np<3, false>(sparse, NP, p, st, grid_out);
    default:
      return fast ? launch_row_np<4, true>(sparse, NP, p, st, grid_out)
                  : launch_row_np<4, false>(sparse, NP, p, st, grid_out);
  }
}

}  // namespace vipsy
