// Enumerated cognitive-diagnosis models: exact marginal over the 2^K attribute
// patterns (TraceEnum_ELBO over VCCDM.model vi.py:839-857; VCHoDina.model/guide
// vi.py:897-934), loss and every gradient in one pass, one warp per person.
//
// The reference materialises a C x B x J Bernoulli log-prob tensor.  Here the
// pattern axis is handled as a lattice transform instead of a contraction:
//
//   DINA  eta_cj = [q_j subset-of c]                      (bit masks, vi.py:78-81)
//   ll_b(c) = sum_j lb(y_bj | eta_cj) = base_b + sum_j delta_bj [q_j subset-of c]
//           = base_b + (zeta f_b)(c),   f_b(m) = sum_{j: q_j = m} delta_bj
//
// so all 2^K pattern log-likelihoods of a person are the subset-sum (zeta)
// transform of a sparse 2^K vector: K 2^(K-1) adds instead of J 2^K MACs, exact
// fp32 adds, eta never materialised (it is the AND/compare [q & c == q]).
// The backward pass needs R_bj = sum_c post_bc eta_cj = (superset-zeta post_b)(q_j),
// the transposed transform.  DINO as the reference executes it (vi.py:95-98,
// SURVEY App. B.1) is eta_cj = [|q_j| >= 2] (1 - [q_j subset-of ~c]): same
// machinery on the complemented pattern index.
//
// Register/lane layout for the 1024-entry lattice (K is embedded in 10 bits;
// unused attribute bits replicate patterns, which leaves a uniform prior's
// marginal unchanged and is compensated exactly for the HO-DINA prior):
//   layout A: c = q*128 + lane*4 + e   -> registers hold bits {0,1,7,8,9}
//   layout B: c = hi*128 + t*4 + lo    -> registers hold bits {2..6}, lane = hi*4+lo
// A zeta pass over the 5 register bits is 80 in-register adds; the switch A<->B
// is one padded shared-memory transpose (4 KB out, 4 KB back, conflict-free).
#include "common.cuh"

namespace vipsy {

constexpr int kCdmWarps = 4;
constexpr int kCdmThreads = kCdmWarps * 32;
constexpr int kLat = 1024;              // lattice size handled in registers
constexpr int kLatPad = kLat + 32;      // +4 floats per 128
constexpr int kMaxSlots = 8;            // J <= 256

struct CdmEnumParams {
  const uint8_t* y;
  int64_t ld_y;
  const int64_t* idx;
  const uint32_t* qmask;
  const float *g, *s, *lam0, *lam1, *mu, *sigma, *eps;
  int64_t B;
  int J, K, S, model, hodina;
  float scale;
  uint64_t seed, offset;
  int64_t person_offset;
  float *d_mu, *d_sigma, *posterior;
  const float* log_prior;   // B x 2^K per-person log pattern prior (VaeCCDM, vi.py:880-886) or null = uniform
  int score_missing;
  float* partial;        // [grid][2J + 2K]
  double* loss_partial;  // [grid]
  int want_grad;
};

__device__ __forceinline__ int pad_addr(int c) { return c + ((c >> 7) << 2); }

// The 32 lattice values of a lane live in 16 float2 registers, value t = pair t/2, half t%2, so that four of
// the five stages of each transform (and the log-sum-exp arithmetic) run as sm_100 packed FADD2 / FFMA2 / FMUL2.
#define LAT(v, t) (((t) & 1) ? (v)[(t) >> 1].y : (v)[(t) >> 1].x)

// subset-sum over the 5 register-index bits: v[i] += v[i ^ bit] for i with bit set
__device__ __forceinline__ void zeta_regs(float2 (&v)[16]) {
#pragma unroll
  for (int k = 0; k < 16; ++k) v[k].y += v[k].x;                           // bit 0: inside the pair
#pragma unroll
  for (int bit = 1; bit < 16; bit <<= 1) {
#pragma unroll
    for (int k = 0; k < 16; ++k)
      if (k & bit) v[k] = __fadd2_rn(v[k], v[k ^ bit]);
  }
}
// superset-sum (transpose of the above): v[i] += v[i | bit] for i with bit clear
__device__ __forceinline__ void zeta_regs_T(float2 (&v)[16]) {
#pragma unroll
  for (int k = 0; k < 16; ++k) v[k].x += v[k].y;
#pragma unroll
  for (int bit = 1; bit < 16; bit <<= 1) {
#pragma unroll
    for (int k = 0; k < 16; ++k)
      if (!(k & bit)) v[k] = __fadd2_rn(v[k], v[k | bit]);
  }
}

// registers in layout A (index i = q*4 + e) <-> padded shared memory
__device__ __forceinline__ void store_A(float* X, const float2 (&v)[16], int lane) {
#pragma unroll
  for (int q = 0; q < 8; ++q)
    *reinterpret_cast<float4*>(X + q * 132 + lane * 4) = make_float4(v[2 * q].x, v[2 * q].y, v[2 * q + 1].x, v[2 * q + 1].y);
}
__device__ __forceinline__ void load_A(const float* X, float2 (&v)[16], int lane) {
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const float4 t = *reinterpret_cast<const float4*>(X + q * 132 + lane * 4);
    v[2 * q] = make_float2(t.x, t.y);
    v[2 * q + 1] = make_float2(t.z, t.w);
  }
}
// registers in layout B (index t) <-> padded shared memory; bank = (hi*4 + t*4 + lo) % 32: conflict-free
__device__ __forceinline__ void store_B(float* X, const float2 (&v)[16], int lane) {
  const int base = (lane >> 2) * 132 + (lane & 3);
#pragma unroll
  for (int t = 0; t < 32; ++t) X[base + t * 4] = LAT(v, t);
}
__device__ __forceinline__ void load_B(const float* X, float2 (&v)[16], int lane) {
  const int base = (lane >> 2) * 132 + (lane & 3);
#pragma unroll
  for (int t = 0; t < 32; ++t) LAT(v, t) = X[base + t * 4];
}

// Sum 16 per-lane values across the warp with a transposing butterfly (16 shuffles):
// afterwards every lane holds the total of v[(lane >> 1) & 15].
__device__ __forceinline__ float warp_sum16(float (&v)[16], int lane) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const bool hi = lane & 16;
    const float keep = hi ? v[i + 8] : v[i], send = hi ? v[i] : v[i + 8];
    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const bool hi = lane & 8;
    const float keep = hi ? v[i + 4] : v[i], send = hi ? v[i] : v[i + 4];
    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const bool hi = lane & 4;
    const float keep = hi ? v[i + 2] : v[i], send = hi ? v[i] : v[i + 2];
    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
  {
    const bool hi = lane & 2;
    const float keep = hi ? v[1] : v[0], send = hi ? v[0] : v[1];
    v[0] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
  }
  return v[0] + __shfl_xor_sync(0xffffffffu, v[0], 1);
}

struct ItemConst {      // per item, built once per CTA
  float base1, base0;   // lb(y | eta = 0): log clamp(g), log(1 - clamp(g))
  float delta1, delta0; // lb(y | eta = 1) - lb(y | eta = 0)
  float da1, da0;       // d lb(y|eta=1) / d s   for y = 1, 0
  float db1, db0;       // d lb(y|eta=0) / d g   for y = 1, 0
  // VaeCCDM scores a missing cell as Bernoulli(p).log_prob(-1) = 2 log(1-p) - log p (vi.py:881-891: y = -1, no mask)
  float baseM, deltaM, daM, dbM;
};

// NS: 32-item slots per lane (J <= 32 NS); the per-lane item accumulators are 4 NS registers
// SCOREM (the VaeCCDM variant): per-person log prior input and / or missing cells scored as a third response
// category instead of being skipped.  A separate instantiation: the plain kernel keeps its register budget.
template <bool HODINA, int NS, bool SCOREM = false>
__global__ void __launch_bounds__(kCdmThreads, 3) cdm_enum_kernel(const CdmEnumParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int J = p.J, K = p.K;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // shared layout: ItemConst[J] | mask[J] | round[J] | flags[J] | per-warp F[kLatPad] X[kLatPad]
  ItemConst* ic = reinterpret_cast<ItemConst*>(smem_raw);
  int* imask = reinterpret_cast<int*>(ic + J);
  int* iround = imask + J;
  int* iflag = iround + J;
  float* wbuf = reinterpret_cast<float*>(smem_raw + (((size_t)J * (sizeof(ItemConst) + 12) + 15) & ~size_t(15)));
  float* F = wbuf + (size_t)warp * 2 * kLatPad;
  float* X = F + kLatPad;
  __shared__ int max_round_sm;
  __shared__ double loss_sm[kCdmWarps];

  if (threadIdx.x == 0) max_round_sm = 0;
  __syncthreads();
  const bool dino = p.model == VIPSY_CDM_DINO;
  for (int j = threadIdx.x; j < J; j += kCdmThreads) {
    const float gj = p.g[j], sj = p.s[j];
    const float p1 = 1.f - sj, p0 = gj;                       // P(y=1 | eta=1), P(y=1 | eta=0)  (vi.py:82)
    const float p1c = fminf(fmaxf(p1, kEps32), kOneMinusEps32), p0c = fminf(fmaxf(p0, kEps32), kOneMinusEps32);
    const float a1 = logf(p1c), a0 = log1pf(-p1c), b1 = logf(p0c), b0 = log1pf(-p0c);
    ItemConst c;
    c.base1 = b1; c.base0 = b0; c.delta1 = a1 - b1; c.delta0 = a0 - b0;
    const float k1 = (p1 == p1c) ? 1.f : 0.f, k0 = (p0 == p0c) ? 1.f : 0.f;   // clamp => zero gradient
    c.da1 = -k1 / p1c; c.da0 = k1 / (1.f - p1c);              // d/ds of log(1-s), log(s)
    c.db1 = k0 / p0c;  c.db0 = -k0 / (1.f - p0c);             // d/dg of log(g), log(1-g)
    c.baseM = 2.f * b0 - b1; c.deltaM = (2.f * a0 - a1) - (2.f * b0 - b1);
    c.daM = k1 * (2.f / (1.f - p1c) + 1.f / p1c);             // d/ds of 2 log(s) - log(1-s)
    c.dbM = k0 * (-2.f / (1.f - p0c) - 1.f / p0c);            // d/dg of 2 log(1-g) - log(g)
    const uint32_t m = p.qmask[j] & (uint32_t)((1u << K) - 1u);
    int flag = 0;                                             // bit0: inverted (DINO), bit1: dead (eta == 0)
    if (dino) flag = (__popc(m) >= 2) ? 1 : 2;
    if (flag & 1) { c.delta1 = -c.delta1; c.delta0 = -c.delta0; c.deltaM = -c.deltaM; }   // ll = base + delta - delta [q subset ~c]
    ic[j] = c;
    imask[j] = (int)m;
    iflag[j] = flag;
    int r = 0;
    if (!(flag & 2))
      for (int i = 0; i < j; ++i) {
        const uint32_t mi = p.qmask[i] & (uint32_t)((1u << K) - 1u);
        const bool dead_i = dino && __popc(mi) < 2;
        r += (!dead_i && mi == m);
      }
    iround[j] = r;
    atomicMax(&max_round_sm, r);
  }
  for (int i = threadIdx.x; i < kCdmWarps * 2 * kLatPad; i += kCdmThreads) wbuf[i] = 0.f;
  __syncthreads();
  const int max_round = max_round_sm;
  constexpr int nslots = NS;

  // per-lane item accumulators: slot k <-> item k*32 + lane
  float accS1[NS], accS0[NS], accN1[NS], accN0[NS], accSM[SCOREM ? NS : 1], accNM[SCOREM ? NS : 1];
#pragma unroll
  for (int k = 0; k < NS; ++k) accS1[k] = accS0[k] = accN1[k] = accN0[k] = 0.f;
#pragma unroll
  for (int k = 0; k < (SCOREM ? NS : 1); ++k) accSM[k] = accNM[k] = 0.f;
  float accL0 = 0.f, accL1 = 0.f;     // HO-DINA: lane (2k, 2k+1) accumulates d lam0_k, d lam1_k
  double loss_acc = 0.0;
  // this lane's items (slot k <-> item k*32 + lane), hoisted out of the person loop: lattice address of the
  // item's mask (-1: no lattice term), duplicate-mask round, flags, and the y = 1 / 0 constants
  int it_addr[NS], it_round[NS], it_flag[NS];
  float it_b1[NS], it_b0[NS], it_d1[NS], it_d0[NS], it_bM[SCOREM ? NS : 1], it_dM[SCOREM ? NS : 1];
#pragma unroll
  for (int k = 0; k < NS; ++k) {
    const int j = k * 32 + lane;
    it_addr[k] = -1; it_round[k] = 0; it_flag[k] = 2;
    it_b1[k] = it_b0[k] = it_d1[k] = it_d0[k] = 0.f;
    if (j < J) {
      const ItemConst c = ic[j];
      const int fl = iflag[j];
      it_flag[k] = fl;
      it_round[k] = iround[j];
      it_addr[k] = (fl & 2) ? -1 : pad_addr(imask[j]);
      it_d1[k] = c.delta1; it_d0[k] = c.delta0;
      it_b1[k] = c.base1 + ((fl & 1) ? -c.delta1 : 0.f);
      it_b0[k] = c.base0 + ((fl & 1) ? -c.delta0 : 0.f);
      if constexpr (SCOREM) {
        it_dM[k] = c.deltaM;
        it_bM[k] = c.baseM + ((fl & 1) ? -c.deltaM : 0.f);
      }
    }
  }

  // HO-DINA per-lane constants: attribute owned by this lane for the logit stage
  const int kk = (lane >> 1) & 15;
  float lam0k = 0.f, lam1k = 0.f;
  if (HODINA && kk < K) { lam0k = p.lam0[kk]; lam1k = p.lam1[kk]; }
  const float shift = (float)(10 - K) * kLn2;      // phantom-bit replication of each real pattern
  const float S_f = (float)p.S, invS = 1.f / S_f;

  const int64_t gwarp = (int64_t)blockIdx.x * kCdmWarps + warp;
  const int64_t nwarps = (int64_t)gridDim.x * kCdmWarps;

  // the response codes of a person are fetched one person ahead: ~10 % of the stall samples sat on their first use
  auto load_codes = [&](int64_t b, unsigned (&cd)[NS]) {
#pragma unroll
    for (int k = 0; k < NS; ++k) cd[k] = 255u;
    if (b < p.B) {
      const int64_t row = p.idx ? p.idx[b] : b;
      const uint8_t* yrow = p.y + row * p.ld_y;
#pragma unroll
      for (int k = 0; k < NS; ++k) {
        if (k < nslots) {
          const int j = k * 32 + lane;
          if (j < J) cd[k] = yrow[j];
        }
      }
    }
  };
  unsigned codes_next[NS];
  load_codes(gwarp, codes_next);
  for (int64_t b = gwarp; b < p.B; b += nwarps) {
    // ---- responses -> base, sparse f --------------------------------------
    unsigned codes[NS];
    float base_part = 0.f;
#pragma unroll
    for (int k = 0; k < NS; ++k) codes[k] = codes_next[k];
    load_codes(b + nwarps, codes_next);
    // SCOREM: a missing cell of a real item becomes code 2 (third category), padding slots stay 255
    if constexpr (SCOREM) {
      if (p.score_missing) {
#pragma unroll
        for (int k = 0; k < NS; ++k)
          if (codes[k] == 255u && k * 32 + lane < J) codes[k] = 2u;
      }
    }
#pragma unroll
    for (int k = 0; k < NS; ++k)
      if (codes[k] != 255u)
        base_part += (SCOREM && codes[k] == 2u) ? it_bM[SCOREM ? k : 0] : (codes[k] != 0u ? it_b1[k] : it_b0[k]);
    for (int rd = 0; rd <= max_round; ++rd) {          // items sharing a mask take turns (plain read-modify-write)
#pragma unroll
      for (int k = 0; k < NS; ++k)
        if (codes[k] != 255u && it_addr[k] >= 0 && it_round[k] == rd)
          F[it_addr[k]] += (SCOREM && codes[k] == 2u) ? it_dM[SCOREM ? k : 0] : (codes[k] != 0u ? it_d1[k] : it_d0[k]);
      __syncwarp();
    }
    const float base = warp_sum(base_part);
    float2 v[16];
    load_A(F, v, lane);
    __syncwarp();
#pragma unroll
    for (int k = 0; k < NS; ++k)        // keep F sparse-zero for the next person
      if (codes[k] != 255u && it_addr[k] >= 0) F[it_addr[k]] = 0.f;
    // ---- forward lattice transform: ll(c) - base ----------------------------
    zeta_regs(v);
    store_A(X, v, lane);
    __syncwarp();
    load_B(X, v, lane);
    zeta_regs(v);
    __syncwarp();          // X is rewritten below

    // ---- posterior over patterns (per particle for HO-DINA) -----------------
    float2 post[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) post[k] = make_float2(0.f, 0.f);
    float mu = 0.f, sg = 1.f, se = 0.f, see = 0.f, gth_sum = 0.f, gth_eps = 0.f, marg_sum = 0.f;
    if (HODINA) { mu = p.mu[b]; sg = p.sigma[b]; }
    const int n_part = HODINA ? p.S : 1;
    for (int sidx = 0; sidx < n_part; ++sidx) {
      float eps_s = 0.f, theta = 0.f, pik = 0.5f;
      if (HODINA) {
        if (p.eps) eps_s = p.eps[(int64_t)sidx * p.B + b];
        else {
          const float4 z = normal4(p.seed, p.offset, (uint64_t)(p.person_offset + b), (uint32_t)(sidx >> 2));
          const float zz[4] = {z.x, z.y, z.z, z.w};
          eps_s = zz[sidx & 3];
        }
        theta = fmaf(sg, eps_s, mu);
        se += eps_s; see = fmaf(eps_s, eps_s, see);
        // attribute kk: log pi, log(1-pi) without cancellation; phantom attributes split evenly
        float lp1 = -kLn2, lp0 = -kLn2;
        if (kk < K) {
          const float z = fmaf(theta, lam1k, lam0k);
          const float az = fabsf(z);
          const float sp = log1pf(__expf(-az));            // softplus(-|z|)
          lp1 = -(z < 0.f ? az + sp : sp);                 // -softplus(-z)
          lp0 = -(z < 0.f ? sp : az + sp);                 // -softplus(z)
          pik = 1.f / (1.f + __expf(-z));
        }
        // gather the 10 (lp1, lp0) pairs every lane needs
        float l1[10], l0[10];
#pragma unroll
        for (int a = 0; a < 10; ++a) {
          l1[a] = __shfl_sync(0xffffffffu, lp1, 2 * a);
          l0[a] = __shfl_sync(0xffffffffu, lp0, 2 * a);
        }
        // lane bits: lo = attrs 0,1; hi = attrs 7,8,9.  register bits t = attrs 2..6
        float lw_lane = ((lane & 1) ? l1[0] : l0[0]) + ((lane & 2) ? l1[1] : l0[1]) +
                        ((lane & 4) ? l1[7] : l0[7]) + ((lane & 8) ? l1[8] : l0[8]) + ((lane & 16) ? l1[9] : l0[9]);
        // log prior of the lane's 32 patterns, built by doubling over the register bits (pairs: FADD2)
        float2 lj[16];
        lj[0].x = lw_lane + l0[2] + l0[3] + l0[4] + l0[5] + l0[6];
        lj[0].y = lj[0].x + (l1[2] - l0[2]);
#pragma unroll
        for (int bit = 1; bit < 5; ++bit) {
          const float dlt = l1[2 + bit] - l0[2 + bit];
          const float2 d2 = make_float2(dlt, dlt);
#pragma unroll
          for (int k = (1 << (bit - 1)); k < (1 << bit); ++k) lj[k] = __fadd2_rn(lj[k - (1 << (bit - 1))], d2);
        }
        // Categorical(probs=w): p = w / sum w (= w up to rounding), floor at eps;
        // bit t of `live` = pattern t is not floored
        uint32_t live = 0u;
        float m = -3.0e38f;
#pragma unroll
        for (int t = 0; t < 32; ++t) {
          const float lwt = LAT(lj, t);
          if (!((lwt + shift) < kLogEps32)) live |= 1u << t;
          const float lc = fminf(fmaxf(lwt, kLogEps32 - shift), -kEps32 - shift);   // log clamp(p, eps, 1-eps)
          const float joint = LAT(v, t) + lc;
          LAT(lj, t) = joint;
          m = fmaxf(m, joint);
        }
        m = warp_max(m);
        const float2 l2e = make_float2(kLog2e, kLog2e), nm = make_float2(-m * kLog2e, -m * kLog2e);
        float2 ss = make_float2(0.f, 0.f);
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          const float2 t2 = __ffma2_rn(lj[k], l2e, nm);
          lj[k] = make_float2(ex2_approx(t2.x), ex2_approx(t2.y));
          ss = __fadd2_rn(ss, lj[k]);
        }
        const float ssum = warp_sum(ss.x + ss.y);
        const float inv = 1.f / ssum;
        marg_sum += m + logf(ssum) + base;
        // posterior; attribute marginals of the non-floored mass by a halving tree over the register bits:
        // after level b the survivors are sums over bits 0..b, and the odd partner of each pair carries bit b
        float red[16];
#pragma unroll
        for (int a = 0; a < 16; ++a) red[a] = 0.f;
        const float2 inv2 = make_float2(inv, inv), invS2 = make_float2(invS, invS);
        float2 rn[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          const float2 r = __fmul2_rn(lj[k], inv2);
          post[k] = __ffma2_rn(r, invS2, post[k]);
          rn[k] = make_float2((live >> (2 * k)) & 1u ? r.x : 0.f, (live >> (2 * k + 1)) & 1u ? r.y : 0.f);
        }
        float lvl[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) { red[2] += rn[k].y; lvl[k] = rn[k].x + rn[k].y; }      // register bit 0 = attribute 2
#pragma unroll
        for (int bit = 1; bit < 5; ++bit) {
          const int n = 16 >> bit;                                                        // survivors after this level
#pragma unroll
          for (int k = 0; k < n; ++k) {
            red[2 + bit] += lvl[2 * k + 1];
            lvl[k] = lvl[2 * k] + lvl[2 * k + 1];
          }
        }
        const float tot = lvl[0];
        red[0] = (lane & 1) ? tot : 0.f; red[1] = (lane & 2) ? tot : 0.f;
        red[7] = (lane & 4) ? tot : 0.f; red[8] = (lane & 8) ? tot : 0.f; red[9] = (lane & 16) ? tot : 0.f;
        red[10] = tot;
        const float mine = warp_sum16(red, lane);                     // T_kk (kk < 10) or F (kk == 10)
        const float Fn = __shfl_sync(0xffffffffu, mine, 20);          // lane 20 -> index 10
        // d marg / d logit_k = T_k - pi_k F   (SURVEY A.3 derivation in DESIGN.md)
        const float Mk = (kk < K) ? (mine - pik * Fn) : 0.f;
        float gth = (lane & 1) ? 0.f : Mk * lam1k;                    // each attribute once (even lanes)
        gth = warp_sum(gth) - theta;                                  // + d log N(theta;0,1)/d theta
        gth_sum += gth;
        gth_eps = fmaf(gth, eps_s, gth_eps);
        if (!(lane & 1)) { accL0 += Mk; accL1 = fmaf(Mk, theta, accL1); }
      } else {
        // uniform Categorical(1/C): log prior = -K ln2 per real pattern, exact under embedding;
        // or the person's own log prior (VaeCCDM), staged through X so that the global reads are coalesced
        float prior_const = -10.f * kLn2;
        if (SCOREM && p.log_prior) {
          const float* lp = p.log_prior + b * ((int64_t)1 << K);
          const int cm = (1 << K) - 1;
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            float4 t4;
            const int c0 = q * 128 + lane * 4;
            t4.x = lp[(dino ? ~(c0 + 0) : (c0 + 0)) & cm];
            t4.y = lp[(dino ? ~(c0 + 1) : (c0 + 1)) & cm];
            t4.z = lp[(dino ? ~(c0 + 2) : (c0 + 2)) & cm];
            t4.w = lp[(dino ? ~(c0 + 3) : (c0 + 3)) & cm];
            *reinterpret_cast<float4*>(X + q * 132 + lane * 4) = t4;
          }
          __syncwarp();
          float2 lpv[16];
          load_B(X, lpv, lane);
          __syncwarp();
#pragma unroll
          for (int k = 0; k < 16; ++k) v[k] = __fadd2_rn(v[k], lpv[k]);
          prior_const = -shift;                    // each real pattern is replicated 2^(10-K) times
        }
        float m = -3.0e38f;
#pragma unroll
        for (int k = 0; k < 16; ++k) m = fmaxf(m, fmaxf(v[k].x, v[k].y));
        m = warp_max(m);
        // exp(v - m) = 2^(v log2e - m log2e): one packed FFMA per pair, then the two MUFU.EX2
        const float2 l2e = make_float2(kLog2e, kLog2e), nm = make_float2(-m * kLog2e, -m * kLog2e);
        float2 ss = make_float2(0.f, 0.f);
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          const float2 t2 = __ffma2_rn(v[k], l2e, nm);
          post[k] = make_float2(ex2_approx(t2.x), ex2_approx(t2.y));
          ss = __fadd2_rn(ss, post[k]);
        }
        const float ssum = warp_sum(ss.x + ss.y);
        const float inv = 1.f / ssum;
        marg_sum += m + logf(ssum) + base + prior_const;
        const float2 inv2 = make_float2(inv, inv);
#pragma unroll
        for (int k = 0; k < 16; ++k) post[k] = __fmul2_rn(post[k], inv2);
      }
    }

    // ---- loss + guide gradients ---------------------------------------------
    if (HODINA) {
      // + sum_s [log N(theta_s;0,1) - log q(theta_s)] = sum_s(-theta^2/2 + eps^2/2) + S log sigma
      const float sumx2 = fmaf(S_f * mu, mu, fmaf(2.f * mu * sg, se, sg * sg * see));
      const float elbo = marg_sum - 0.5f * sumx2 + 0.5f * see + S_f * logf(sg);
      if (lane == 0) {
        loss_acc -= (double)(p.scale * invS * elbo);
        if (p.d_mu) {
          p.d_mu[b] = -p.scale * invS * gth_sum;
          p.d_sigma[b] = -p.scale * invS * gth_eps - p.scale / sg;
        }
      }
    } else if (lane == 0) {
      loss_acc -= (double)(p.scale * marg_sum);
    }
    if (p.posterior) {
      // real patterns only: embedded copies carry equal mass
      const float rep = (float)(1 << (10 - K));
#pragma unroll
      for (int t = 0; t < 32; ++t) {
        int c = (lane >> 2) * 128 + t * 4 + (lane & 3);
        if (c < (1 << K)) {
          if (dino) c = (~c) & ((1 << K) - 1);
          p.posterior[b * ((int64_t)1 << K) + c] = LAT(post, t) * rep;
        }
      }
    }

    // ---- backward lattice transform: R(m) = sum_{c superset m} post(c) -----
    if (p.want_grad) {
      zeta_regs_T(post);
      store_B(X, post, lane);
      __syncwarp();
      load_A(X, post, lane);
      zeta_regs_T(post);
      __syncwarp();
      store_A(X, post, lane);
      __syncwarp();
#pragma unroll
      for (int k = 0; k < NS; ++k) {
        if (codes[k] != 255u) {
          float R = it_addr[k] >= 0 ? X[it_addr[k]] : 0.f;
          R = (it_flag[k] & 2) ? 0.f : ((it_flag[k] & 1) ? 1.f - R : R);     // posterior mass with eta_cj = 1
          if (SCOREM && codes[k] == 2u) { accSM[SCOREM ? k : 0] += R; accNM[SCOREM ? k : 0] += 1.f; }
          else if (codes[k] != 0u) { accS1[k] += R; accN1[k] += 1.f; }
          else { accS0[k] += R; accN0[k] += 1.f; }
        }
      }
      __syncwarp();
    }
  }

  // ---- CTA partials -----------------------------------------------------------
  loss_acc = warp_sum(loss_acc);
  if (lane == 0) loss_sm[warp] = loss_acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w2 = 0; w2 < kCdmWarps; ++w2) t += loss_sm[w2];
    p.loss_partial[blockIdx.x] = t;
  }
  if (p.want_grad) {
    // combine the 4 warps through shared memory (reuse wbuf): rows [warp][2J + 2K]
    const int ncol = 2 * J + 2 * K;
    __syncthreads();
    float* red = wbuf;
#pragma unroll
    for (int k = 0; k < NS; ++k) {
      if (k < nslots) {
        const int j = k * 32 + lane;
        if (j < J) {
          const ItemConst c = ic[j];
          // d/dg: eta = 0 mass is (n - S);  d/ds: eta = 1 mass is S
          float dg = (accN1[k] - accS1[k]) * c.db1 + (accN0[k] - accS0[k]) * c.db0;
          float ds = accS1[k] * c.da1 + accS0[k] * c.da0;
          if constexpr (SCOREM) {
            dg += (accNM[k] - accSM[k]) * c.dbM;
            ds += accSM[k] * c.daM;
          }
          red[warp * ncol + j] = dg;
          red[warp * ncol + J + j] = ds;
        }
      }
    }
    if (!(lane & 1) && kk < K) {
      red[warp * ncol + 2 * J + kk] = HODINA ? accL0 : 0.f;
      red[warp * ncol + 2 * J + K + kk] = HODINA ? accL1 : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < ncol; i += kCdmThreads) {
      float t = 0.f;
      for (int w2 = 0; w2 < kCdmWarps; ++w2) t += red[w2 * ncol + i];
      p.partial[(size_t)blockIdx.x * ncol + i] = t;
    }
  }
}

static size_t cdm_smem_bytes(int J) {
  size_t s = ((size_t)J * (sizeof(ItemConst) + 12) + 15) & ~size_t(15);
  return s + (size_t)kCdmWarps * 2 * kLatPad * 4;
}
constexpr int kCdmMaxCtasPerSM = 8;

size_t cdm_enum_workspace_bytes(const VipsyCdmEnumArgs* a) {
  const size_t rows = (size_t)num_sms() * kCdmMaxCtasPerSM;
  return (((rows + kMaskRows) * 8 + 15) & ~size_t(15)) + rows * (2 * (size_t)a->n_items + 2 * (size_t)a->n_attrs) * 4;
}

int cdm_enum_launch(const VipsyCdmEnumArgs* a, cudaStream_t st) {
  CdmEnumParams p;
  p.y = a->y; p.ld_y = a->y_row_stride; p.idx = a->idx; p.qmask = a->qmask;
  p.g = a->g; p.s = a->s; p.lam0 = a->lam0; p.lam1 = a->lam1; p.mu = a->mu; p.sigma = a->sigma; p.eps = a->eps;
  p.B = a->batch; p.J = a->n_items; p.K = a->n_attrs; p.S = a->hodina ? a->n_particles : 1;
  p.model = a->hodina ? VIPSY_CDM_DINA : a->model; p.hodina = a->hodina;
  p.scale = a->scale; p.seed = a->seed; p.offset = a->offset; p.person_offset = a->person_offset;
  p.d_mu = a->d_mu; p.d_sigma = a->d_sigma; p.posterior = a->posterior;
  p.log_prior = a->hodina ? nullptr : a->log_prior;
  p.score_missing = a->score_missing != 0 && !a->hodina;
  const bool scorem = (p.score_missing || p.log_prior) && !a->hodina;
  p.want_grad = (a->d_g || a->d_s || a->d_lam0 || a->d_lam1) ? 1 : 0;
  const size_t rows_max = (size_t)num_sms() * kCdmMaxCtasPerSM;
  p.loss_partial = reinterpret_cast<double*>(a->workspace);
  p.partial = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(a->workspace) + (((rows_max + kMaskRows) * 8 + 15) & ~size_t(15)));
  const size_t smem = cdm_smem_bytes(p.J);
  const int ns = p.J <= 64 ? 2 : (p.J <= 128 ? 4 : 8);
  void (*kern)(const CdmEnumParams);
  if (a->hodina) kern = ns == 2 ? cdm_enum_kernel<true, 2> : (ns == 4 ? cdm_enum_kernel<true, 4> : cdm_enum_kernel<true, 8>);
  else if (scorem) kern = ns == 2 ? cdm_enum_kernel<false, 2, true> : (ns == 4 ? cdm_enum_kernel<false, 4, true> : cdm_enum_kernel<false, 8, true>);
  else kern = ns == 2 ? cdm_enum_kernel<false, 2> : (ns == 4 ? cdm_enum_kernel<false, 4> : cdm_enum_kernel<false, 8>);
  static OccCache occ_cache[9];
  int occ = 0;
  const int which = (a->hodina ? 3 : (scorem ? 6 : 0)) + (ns == 2 ? 0 : (ns == 4 ? 1 : 2));
  if (smem > 48 * 1024) VIPSY_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  VIPSY_CUDA_CHECK(cached_occupancy(occ_cache[which], kern, kCdmThreads, smem, kCdmMaxCtasPerSM, &occ));
  int64_t ctas = (a->batch + kCdmWarps - 1) / kCdmWarps;
  const int64_t cap = (int64_t)num_sms() * occ;
  if (ctas > cap) ctas = cap;
  if (ctas < 1) ctas = 1;
  kern<<<(int)ctas, kCdmThreads, smem, st>>>(p);
  VIPSY_CUDA_CHECK(cudaGetLastError());
  // masked cells (VCCDM masks them, vi.py:852-856): log(1 - eps) each in every pattern's log-likelihood, i.e. a
  // constant of the person's marginal
  int mask_rows = 0;
  if (!p.score_missing) {
    int rc = launch_masked_term(a->y, a->y_row_stride, a->idx, a->batch, p.J, a->scale, -1, p.loss_partial + ctas,
                                &mask_rows, st);
    if (rc != VIPSY_OK) return rc;
  }
  FinalizeSegs segs{};
  segs.n = 4;
  segs.out[0] = a->d_g; segs.cols[0] = p.J;
  segs.out[1] = a->d_s; segs.cols[1] = p.J;
  segs.out[2] = a->d_lam0; segs.cols[2] = p.K;
  segs.out[3] = a->d_lam1; segs.cols[3] = p.K;
  const float alpha = -a->scale / (float)p.S;
  // d_g, d_s were accumulated from the particle-averaged posterior: alpha_items = -scale.  The lam
  // columns are sums over particles: alpha = -scale / S.  Two finalize launches keep it simple.
  if (p.want_grad) {
    FinalizeSegs s1 = segs; s1.out[2] = nullptr; s1.out[3] = nullptr;
    int rc = launch_finalize(p.partial, (int)ctas, 2 * p.J + 2 * p.K, -a->scale, s1, p.loss_partial, (int)ctas + mask_rows, a->loss, st);
    if (rc != VIPSY_OK) return rc;
    if (a->hodina && (a->d_lam0 || a->d_lam1)) {
      FinalizeSegs s2 = segs; s2.out[0] = nullptr; s2.out[1] = nullptr;
      return launch_finalize(p.partial, (int)ctas, 2 * p.J + 2 * p.K, alpha, s2, nullptr, 0, nullptr, st);
    }
    return VIPSY_OK;
  }
  return launch_finalize(nullptr, 0, 0, alpha, segs, p.loss_partial, (int)ctas + mask_rows, a->loss, st);
}

}  // namespace vipsy

using namespace vipsy;

extern "C" {

size_t vipsy_cdm_enum_workspace_bytes(const VipsyCdmEnumArgs* a) {
  if (!a || a->n_items <= 0 || a->n_attrs <= 0) return 0;
  return cdm_enum_workspace_bytes(a);
}

int vipsy_cdm_enum_elbo_grad(const VipsyCdmEnumArgs* a, void* stream) {
  VIPSY_REQUIRE(a != nullptr, VIPSY_ERR_ARG, "vipsy_cdm_enum: args is NULL");
  VIPSY_REQUIRE(a->model == VIPSY_CDM_DINA || a->model == VIPSY_CDM_DINO, VIPSY_ERR_ARG,
                "vipsy_cdm_enum: unknown model %d (CDM_FUN has dina, dino)", a->model);
  VIPSY_REQUIRE(a->batch >= 0 && a->n_items > 0 && a->n_attrs > 0 && a->n_particles > 0, VIPSY_ERR_ARG,
                "vipsy_cdm_enum: bad sizes B=%lld J=%d K=%d S=%d", (long long)a->batch, a->n_items, a->n_attrs,
                a->n_particles);
  if (a->batch == 0) {                       // nothing to evaluate: loss 0, gradients 0
    VIPSY_REQUIRE(a->loss, VIPSY_ERR_ARG, "vipsy_cdm_enum: loss is required");
    cudaStream_t st = (cudaStream_t)stream;
    VIPSY_CUDA_CHECK(cudaMemsetAsync(a->loss, 0, 8, st));
    if (a->d_g) VIPSY_CUDA_CHECK(cudaMemsetAsync(a->d_g, 0, (size_t)a->n_items * 4, st));
    if (a->d_s) VIPSY_CUDA_CHECK(cudaMemsetAsync(a->d_s, 0, (size_t)a->n_items * 4, st));
    if (a->d_lam0) VIPSY_CUDA_CHECK(cudaMemsetAsync(a->d_lam0, 0, (size_t)a->n_attrs * 4, st));
    if (a->d_lam1) VIPSY_CUDA_CHECK(cudaMemsetAsync(a->d_lam1, 0, (size_t)a->n_attrs * 4, st));
    return VIPSY_OK;
  }
  VIPSY_REQUIRE(a->n_attrs <= 10, VIPSY_ERR_UNSUPPORTED,
                "vipsy_cdm_enum: K=%d attributes (2^K patterns) > 10 is not built", a->n_attrs);
  VIPSY_REQUIRE(a->n_items <= 32 * kMaxSlots, VIPSY_ERR_UNSUPPORTED, "vipsy_cdm_enum: J=%d > %d", a->n_items,
                32 * kMaxSlots);
  VIPSY_REQUIRE(a->y && a->qmask && a->g && a->s && a->loss, VIPSY_ERR_ARG,
                "vipsy_cdm_enum: y, qmask, g, s, loss are required");
  VIPSY_REQUIRE(a->y_row_stride >= a->n_items, VIPSY_ERR_ARG, "vipsy_cdm_enum: y_row_stride < J");
  VIPSY_REQUIRE(a->idx != nullptr || a->batch <= a->n_rows, VIPSY_ERR_ARG, "vipsy_cdm_enum: batch > rows without idx");
  if (a->hodina) {
    VIPSY_REQUIRE(a->lam0 && a->lam1 && a->mu && a->sigma, VIPSY_ERR_ARG,
                  "vipsy_cdm_enum: HO-DINA needs lam0, lam1, mu, sigma");
    VIPSY_REQUIRE((a->d_mu == nullptr) == (a->d_sigma == nullptr), VIPSY_ERR_ARG,
                  "vipsy_cdm_enum: d_mu and d_sigma go together");
  }
  VIPSY_REQUIRE(a->workspace && a->workspace_bytes >= cdm_enum_workspace_bytes(a), VIPSY_ERR_WORKSPACE,
                "vipsy_cdm_enum: workspace %zu B < required %zu B", a->workspace_bytes, cdm_enum_workspace_bytes(a));
  return cdm_enum_launch(a, (cudaStream_t)stream);
}

}  // extern "C"
