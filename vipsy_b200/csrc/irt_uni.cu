// Unidimensional IRT (K == 1): fused sampling + 1..4PL link + masked Bernoulli
// log-likelihood + every gradient, S Monte-Carlo particles per person, one pass.
//
// Replaces, per SVI.step (vi.py:503-516), the whole Trace_ELBO evaluation of
// BaseIRT.model (vi.py:574-595, _get_p_data vi.py:617-625, irt_1pl..4pl
// vi.py:22-66) under the VIRT / VaeIRT Normal guide (vi.py:701-705, 677-684).
//
// Mapping: one lane = one person, a warp = 32 consecutive persons (a "tile").
// The tile's response codes are staged in shared memory with 16-byte loads;
// item parameters sit in shared memory as one float4 per item (broadcast reads).
// For each item the lane runs the S particles out of registers (eps[SC]) -- the
// logits never leave the register file.  Per-item gradient sums are reduced
// across the warp with a transposing butterfly (6 shuffles for 4 quantities)
// into warp-private shared-memory accumulators; CTAs write one partial row each
// and a second tiny kernel sums the rows in a fixed order => bit-reproducible.
//
// Arithmetic follows the reference's formulas: probability, clamp to
// [eps, 1-eps], log, zero gradient where the clamp fires -- see item_evals().
// Roof: 3-4 MUFU per (person, item, particle) evaluation => SFU-bound, not HBM.
#include <algorithm>
#include <cstdlib>

#include "irt_uni.cuh"

namespace vipsy {

// Item loop of the pathwise estimator as ONE basic block per item (PIPE instantiation of the kernel below).
// The lane-per-person loop used to spend a quarter of each item in a serial chain that feeds nothing to the
// MUFU pipe: response byte -> vote -> branch -> table loads -> evaluations -> six dependent shuffles ->
// shared-memory add.  Here (i) the cross-warp reduction of item j-1 is issued inside the evaluations of item j
// (the sums wait in registers; the first round adds zeros to row 0), (ii) missing cells are evaluated like
// y = 1 and masked by selects instead of a divergent branch, (iii) the response byte is fetched one item
// ahead, and (iv) the rational-form / clamp choice is made per tile outside the loop (MODE), so nothing splits
// the block and the scheduler interleaves the shuffle chain with the MUFU stream.
// MODE 0: probability-clamp form; 1: rational form, exponent clamped; 2: rational form, exponent bounded
// by the caller (no FMNMX).  Same sums in the same order as the branchy loop: bit-identical results.
template <int MODEL, int SC, int MODE>
struct UniItem {
  static constexpr int NQ = MODEL;
  // evaluates item j for the lane's person; v[] = the lane's terms of the item-gradient sums (masked)
  template <typename EpsT>
  static __device__ __forceinline__ void eval(const int j, const unsigned code, const float4* __restrict__ tab,
                                              const float2* __restrict__ tab2, const float mu, const float sg,
                                              const EpsT& eps,
                                              float (&v)[NQ], float& Lsum, float& G1, float& G2) {
    const bool obs = code != 255u;
    const float4 it = tab[j];
    const float2 it2 = tab2[j];
    const bool y1 = code != 0u;                                     // a missing cell runs as y = 1 and is masked below
    const float A = y1 ? it.x : -it.x, Bc = y1 ? it.y : -it.y;
    const float T0 = fmaf(A, mu, Bc), T1 = A * sg;
    const float base = MODEL >= 3 ? (y1 ? it.z : it2.y) : 0.f;
    float W1, W2, A1, A2, Lj = 0.f;
    float dummy[1];
    if constexpr (MODE == 0) item_evals<MODEL, false, SC, false>(it.w, base, T0, T1, eps, dummy, W1, W2, A1, A2, Lj);
    else if constexpr (MODE == 1) item_evals<MODEL, false, SC, true, true>(it.w, base, T0, T1, eps, dummy, W1, W2, A1, A2, Lj);
    else item_evals<MODEL, false, SC, true, false>(it.w, base, T0, T1, eps, dummy, W1, W2, A1, A2, Lj);
    const float sgn = obs ? (y1 ? 1.f : -1.f) : 0.f;
    W1 *= sgn; W2 *= sgn;
    Lsum += obs ? Lj : 0.f;
    G1 = fmaf(it2.x, W1, G1);
    G2 = fmaf(it2.x, W2, G2);
    v[0] = W1;
    if constexpr (NQ >= 2) v[1] = fmaf(sg, W2, mu * W1);
    if constexpr (NQ >= 3) v[2] = obs ? (y1 ? A2 : -A1) : 0.f;
    if constexpr (NQ >= 4) v[3] = obs ? (y1 ? A1 : -A2) : 0.f;
  }
};

template <int MODEL, int SC, int MODE, int UNR, typename EpsT>
__device__ __forceinline__ void uni_item_loop(const int J, const float4* __restrict__ tab, const float2* __restrict__ tab2,
                                              const unsigned char* __restrict__ yrow, const bool valid, const int lane,
                                              float* __restrict__ acc, const float mu, const float sg,
                                              const EpsT& eps,
                                              float& Lsum, float& G1, float& G2) {
  constexpr int NQ = MODEL;
  using It = UniItem<MODEL, SC, MODE>;
  const int own = warp_sum_multi_owner<NQ>(lane);
  float vp[UNR][NQ];
  int jp[UNR];
#pragma unroll
  for (int u = 0; u < UNR; ++u) {
    jp[u] = 0;
#pragma unroll
    for (int q = 0; q < NQ; ++q) vp[u][q] = 0.f;
  }
  // response bytes one step ahead; reads past the row stay inside the warp's tile buffer (16 bytes of slack)
  unsigned next[UNR];
#pragma unroll
  for (int u = 0; u < UNR; ++u) next[u] = yrow[u < J ? u : 0];
  int j = 0;
  for (; j + UNR <= J; j += UNR) {
    unsigned code[UNR];
    bool any_obs = false;
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      code[u] = valid ? next[u] : 255u;
      any_obs |= code[u] != 255u;
    }
#pragma unroll
    for (int u = 0; u < UNR; ++u) next[u] = yrow[(j + UNR + u < J ? j + UNR + u : j)];
    if (!__any_sync(0xffffffffu, any_obs)) continue;
    // the previous step's items cross the warp while this step's are evaluated
    float v[UNR][NQ];
#pragma unroll
    for (int u = 0; u < UNR; ++u) It::eval(j + u, code[u], tab, tab2, mu, sg, eps, v[u], Lsum, G1, G2);
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const float r = warp_sum_multi<NQ>(vp[u], lane);
      if (own >= 0) acc[own * J + jp[u]] += r;
    }
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      jp[u] = j + u;
#pragma unroll
      for (int q = 0; q < NQ; ++q) vp[u][q] = v[u][q];
    }
  }
#pragma unroll
  for (int u = 0; u < UNR; ++u) {
    const float r = warp_sum_multi<NQ>(vp[u], lane);
    if (own >= 0) acc[own * J + jp[u]] += r;
  }
  if constexpr (UNR > 1) {
    for (; j < J; ++j) {                                            // J % UNR items, unpipelined
      const unsigned code = valid ? (unsigned)yrow[j] : 255u;
      if (!__any_sync(0xffffffffu, code != 255u)) continue;
      float v[NQ];
      It::eval(j, code, tab, tab2, mu, sg, eps, v, Lsum, G1, G2);
      const float r = warp_sum_multi<NQ>(v, lane);
      if (own >= 0) acc[own * J + j] += r;
    }
  }
}

// MODEL: 1..4 (number of item parameters).  SCORE: REINFORCE estimator for the
// guide parameters.  SC: particles kept in registers at once (S % SC == 0).
// FAST (4PL only): compile in the rational-form path; each CTA checks the item
// table once and uses it only if no item can reach the probability clamp.
// PIPE (pathwise only): the item loop is uni_item_loop above.
// ESM: the lane's noise pairs live in shared memory ([pair][thread], one conflict-free LDS.64 per pair) instead of 32
// registers, so that five CTAs fit an SM (96 registers): a fifth warp per scheduler for the MUFU pipe.
template <int MODEL, bool SCORE, int SC, bool FAST, bool PIPE, int kUniUnroll = 1, bool ESM = false>
__global__ void __launch_bounds__(kUniThreads, ESM ? 5 : 4) irt_uni_kernel(const IrtUniParams p) {
  static_assert(!(PIPE && SCORE), "the pipelined item loop serves the pathwise estimator");
  constexpr int NQ = MODEL;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int J = p.J;
  float4* tab = reinterpret_cast<float4*>(smem_raw);
  float2* tab2 = reinterpret_cast<float2*>(smem_raw + (size_t)J * 16);
  float* acc_all = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(tab2) + align16((size_t)J * 8));
  unsigned char* ytiles = reinterpret_cast<unsigned char*>(acc_all) + align16((size_t)kUniWarps * NQ * J * 4);
  const size_t ytile_bytes = align16((size_t)32 * J + 16);
  typename NoiseVec<SC>::type* eps_sm =
      reinterpret_cast<typename NoiseVec<SC>::type*>(ytiles + (size_t)kUniWarps * ytile_bytes) + threadIdx.x;

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* acc = acc_all + (size_t)warp * NQ * J;
  unsigned char* ytile = ytiles + (size_t)warp * ytile_bytes;

  for (int j = threadIdx.x; j < J; j += kUniThreads) {
    const float aj = MODEL >= 2 ? p.a[j] : 1.f;
    const float cj = MODEL >= 3 ? p.c[j] : 0.f;
    const float dj = MODEL == 4 ? p.d[j] : 1.f;
    tab[j] = make_float4(-p.D * kLog2e * aj, -p.D * kLog2e * p.b[j], cj, dj - cj);
    tab2[j] = make_float2(p.D * aj * (dj - cj), 1.f - dj);
  }
  for (int i = threadIdx.x; i < kUniWarps * NQ * J; i += kUniThreads) acc_all[i] = 0.f;
  bool use_fast = false;
  float amax = 0.f, bmax = 0.f;   // max |A'_j|, |B'_j| over the item table (PIPE && FAST: exponent bound per tile)
  if (FAST) {   // every item bounded away from the clamp => rational form is safe
    int ok = 1;
    for (int j = threadIdx.x; j < J; j += kUniThreads) {
      const float cj = p.c[j], dj = p.d[j];
      ok &= (cj >= 1e-3f) && (dj <= 1.f - 1e-3f) && (dj - cj >= 1e-3f);
    }
    use_fast = __syncthreads_and(ok) != 0;
    if (PIPE) {
      for (int j = lane; j < J; j += 32) {
        amax = fmaxf(amax, fabsf(tab[j].x));
        bmax = fmaxf(bmax, fabsf(tab[j].y));
      }
#pragma unroll
      for (int o = 16; o; o >>= 1) {
        amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
        bmax = fmaxf(bmax, __shfl_xor_sync(0xffffffffu, bmax, o));
      }
    }
  } else {
    __syncthreads();
  }

  const float S_f = (float)p.S;
  const float k_scale = p.scale / S_f;
  double loss_acc = 0.0;

  const int64_t ntiles = (p.B + 31) / 32;
  const int64_t gwarp = (int64_t)blockIdx.x * kUniWarps + warp;
  const int64_t nwarps = (int64_t)gridDim.x * kUniWarps;

  for (int64_t tile = gwarp; tile < ntiles; tile += nwarps) {
    const int64_t b0 = tile * 32;
    const int64_t bi = b0 + lane;
    const bool valid = bi < p.B;
    const int rows = (int)min((int64_t)32, p.B - b0);

    // ---- stage the tile's response codes ---------------------------------
    int mis = 0;
    if (p.packed2) {
      // 2-bit transport rows (4 cells per byte, 0 / 1 / 3 = missing; J % 4 == 0, so the tile is one contiguous run
      // of rows * J / 4 bytes that expands to rows * J contiguous codes): a 32-bit word becomes one 16-byte store
      const unsigned char* g0 = p.y + b0 * (int64_t)(J >> 2);
      const int nbytes = rows * (J >> 2), nw = nbytes >> 2;
      const uint32_t* gw = reinterpret_cast<const uint32_t*>(g0);          // b0 is a multiple of 32: 4-byte aligned
      int4* dst = reinterpret_cast<int4*>(ytile);
      auto expand = [](const uint32_t b) {                                  // one packed byte -> four code bytes
        const uint32_t t = (b | (b << 6) | (b << 12) | (b << 18)) & 0x03030303u;
        const uint32_t m = t & (t >> 1) & 0x01010101u;                      // 3 -> 255
        return t | ((m << 8) - m);
      };
      for (int v = lane; v < nw; v += 32) {
        const uint32_t w = __ldg(gw + v);
        dst[v] = make_int4((int)expand(w & 0xffu), (int)expand((w >> 8) & 0xffu), (int)expand((w >> 16) & 0xffu),
                           (int)expand(w >> 24));
      }
      for (int k = (nw << 2) + lane; k < nbytes; k += 32)                   // last, partial tile: up to 3 bytes
        reinterpret_cast<uint32_t*>(ytile)[k] = expand(g0[k]);
    } else if (p.contiguous) {
      const unsigned char* g0 = p.y + b0 * (int64_t)J;
      mis = (int)(reinterpret_cast<uintptr_t>(g0) & 15);
      const int4* gal = reinterpret_cast<const int4*>(g0 - mis);
      const int nvec = (mis + rows * J + 15) >> 4;
      int4* dst = reinterpret_cast<int4*>(ytile);
      int v = lane;
      for (; v + 96 < nvec; v += 128) {
        const int4 r0 = __ldg(gal + v), r1 = __ldg(gal + v + 32), r2 = __ldg(gal + v + 64), r3 = __ldg(gal + v + 96);
        dst[v] = r0; dst[v + 32] = r1; dst[v + 64] = r2; dst[v + 96] = r3;
      }
      for (; v < nvec; v += 32) dst[v] = __ldg(gal + v);
    } else {
      for (int r = 0; r < rows; ++r) {
        const int64_t row = p.idx ? p.idx[b0 + r] : (b0 + r);
        const unsigned char* g = p.y + row * p.ld_y;
        for (int i = lane; i < J; i += 32) ytile[r * J + i] = g[i];
      }
    }
    __syncwarp();
    const unsigned char* yrow = ytile + mis + lane * J;

    const float mu = valid ? p.mu[bi] : 0.f;
    const float sg = valid ? p.sigma[bi] : 1.f;
    float se = 0.f, see = 0.f, G1 = 0.f, G2 = 0.f, Lsum = 0.f, sw1 = 0.f, sw2 = 0.f;

    for (int s0 = 0; s0 < p.S; s0 += SC) {
      // ---- noise for SC particles ----------------------------------------
      using V = typename NoiseVec<SC>::type;
      constexpr int VW = NoiseVec<SC>::W;
      V eps[SC / VW];
      {
        float ef[SC];
        if (p.eps) {
#pragma unroll
          for (int i = 0; i < SC; ++i) ef[i] = valid ? __ldg(p.eps + (int64_t)(s0 + i) * p.B + bi) : 0.f;
        } else {
          const uint64_t person = (uint64_t)(p.person_offset + bi);
          if (SC >= 4) {
#pragma unroll
            for (int i = 0; i < SC; i += 4) {
              const float4 z = normal4(p.seed, p.offset, person, (uint32_t)((s0 + i) >> 2));
              ef[i] = z.x;
              if (i + 1 < SC) ef[i + 1] = z.y;
              if (i + 2 < SC) ef[i + 2] = z.z;
              if (i + 3 < SC) ef[i + 3] = z.w;
            }
          } else {
            const float4 z = normal4(p.seed, p.offset, person, (uint32_t)(s0 >> 2));
            const float zz[4] = {z.x, z.y, z.z, z.w};
#pragma unroll
            for (int i = 0; i < SC; ++i) ef[i] = zz[(s0 + i) & 3];
          }
        }
#pragma unroll
        for (int i = 0; i < SC; ++i) {
          se += ef[i];
          see = fmaf(ef[i], ef[i], see);
        }
        if constexpr (VW == 2) {
#pragma unroll
          for (int i = 0; i < SC / 2; ++i) eps[i] = make_float2(ef[2 * i], ef[2 * i + 1]);
        } else {
          eps[0] = ef[0];
        }
      }
      float lls[SCORE ? SC : 1];
      if (SCORE) {
#pragma unroll
        for (int i = 0; i < SC; ++i) lls[i] = 0.f;
      }

      // ---- items ----------------------------------------------------------
      if constexpr (PIPE) {
        if constexpr (ESM) {
          __syncwarp();                                  // the previous chunk's readers are done
#pragma unroll
          for (int i = 0; i < SC / VW; ++i) eps_sm[i * kUniThreads] = eps[i];
          __syncwarp();
        }
#define UNI_LOOP(MODE)                                                                                              \
  do {                                                                                                              \
    if constexpr (ESM)                                                                                              \
      uni_item_loop<MODEL, SC, MODE, kUniUnroll>(J, tab, tab2, yrow, valid, lane, acc, mu, sg, eps_acc, Lsum, G1, G2);  \
    else                                                                                                            \
      uni_item_loop<MODEL, SC, MODE, kUniUnroll>(J, tab, tab2, yrow, valid, lane, acc, mu, sg, eps, Lsum, G1, G2);  \
  } while (0)
        EpsShared<SC> eps_acc{eps_sm, kUniThreads};
        if (FAST && use_fast) {
          // |tt| <= max|A'| (|mu| + sigma max|eps|) + max|B'| for every item and particle of the person
          float emax = 0.f;
#pragma unroll
          for (int i = 0; i < SC / VW; ++i) {
            if constexpr (VW == 2) emax = fmaxf(emax, fmaxf(fabsf(eps[i].x), fabsf(eps[i].y)));
            else emax = fmaxf(emax, fabsf(eps[i]));
          }
          const float bound = fmaf(amax, fmaf(fabsf(sg), emax, fabsf(mu)), bmax);
          if (__all_sync(0xffffffffu, bound < 60.f))      // false for NaN / inf inputs as well
            UNI_LOOP(2);
          else
            UNI_LOOP(1);
        } else {
          UNI_LOOP(0);
        }
#undef UNI_LOOP
      } else
      for (int j = 0; j < J; ++j) {
        const unsigned code = valid ? (unsigned)yrow[j] : 255u;
        const bool obs = code != 255u;
        if (!__any_sync(0xffffffffu, obs)) continue;
        float v[NQ];
#pragma unroll
        for (int q = 0; q < NQ; ++q) v[q] = 0.f;
        if (obs) {
          const float4 it = tab[j];
          const float2 it2 = tab2[j];
          const bool y1 = code != 0u;
          // sign the logit by the response: sp = s (y=1) | 1-s (y=0)
          const float A = y1 ? it.x : -it.x, Bc = y1 ? it.y : -it.y;
          const float T0 = fmaf(A, mu, Bc), T1 = A * sg;
          const float base = MODEL >= 3 ? (y1 ? it.z : it2.y) : 0.f;   // c | 1-d (exact)
          float W1, W2, A1, A2, Lj = 0.f;
          if (FAST && use_fast)
            item_evals<MODEL, SCORE, SC, true>(it.w, base, T0, T1, eps, lls, W1, W2, A1, A2, Lj);
          else
            item_evals<MODEL, SCORE, SC, false>(it.w, base, T0, T1, eps, lls, W1, W2, A1, A2, Lj);
          // undo the signing: dlb/dl = +w (y=1) | -w (y=0); W2 pairs with the true eps
          const float sgn = y1 ? 1.f : -1.f;
          W1 *= sgn; W2 *= sgn;
          Lsum += Lj;
          G1 = fmaf(it2.x, W1, G1);
          G2 = fmaf(it2.x, W2, G2);
          v[0] = W1;
          if constexpr (NQ >= 2) v[1] = fmaf(sg, W2, mu * W1);   // sum_s dlb/dl * x_s
          // dlb/dc = (1-s)/q (y=1) | -s/q (y=0);  dlb/dd = s/q (y=1) | -(1-s)/q (y=0)
          if constexpr (NQ >= 3) v[2] = y1 ? A2 : -A1;
          if constexpr (NQ >= 4) v[3] = y1 ? A1 : -A2;
        }
        if (p.want_grad) {
          const float r = warp_sum_multi<NQ>(v, lane);
          const int own = warp_sum_multi_owner<NQ>(lane);
          if (own >= 0) acc[own * J + j] += r;
        }
      }

      if (SCORE) {
        const float lsg = __logf(sg);
#pragma unroll
        for (int i = 0; i < SC; ++i) {
          float ei;
          if constexpr (VW == 2) ei = (i & 1) ? eps[i >> 1].y : eps[i >> 1].x; else ei = eps[0];
          const float x = fmaf(sg, ei, mu);
          const float e2 = ei * ei;
          const float wgt = p.scale * (kLn2 * lls[i] - 0.5f * x * x + 0.5f * e2 + lsg) - p.baseline;
          sw1 = fmaf(wgt, ei, sw1);
          sw2 = fmaf(wgt, e2 - 1.f, sw2);
          Lsum += lls[i];
        }
      }
    }
    __syncwarp();   // everyone is done with ytile before the next tile overwrites it

    // ---- per-person epilogue ---------------------------------------------
    if (valid) {
      const float sumx = fmaf(sg, se, S_f * mu);
      const float sumxe = fmaf(sg, see, mu * se);
      const float sumx2 = fmaf(S_f * mu, mu, fmaf(2.f * mu * sg, se, sg * sg * see));
      const float elbo = kLn2 * Lsum - 0.5f * sumx2 + 0.5f * see + S_f * logf(sg);
      loss_acc -= (double)(k_scale * elbo);
      if (p.d_mu) {
        if (SCORE) {
          const float inv = 1.f / (S_f * sg);
          p.d_mu[bi] = -sw1 * inv;
          p.d_sigma[bi] = -sw2 * inv;
        } else {
          p.d_mu[bi] = -k_scale * (G1 - sumx);
          p.d_sigma[bi] = -k_scale * (G2 - sumxe) - p.scale / sg;
        }
      }
    }
  }

  // ---- CTA partials ---------------------------------------------------------
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
      for (int w = 0; w < kUniWarps; ++w) t += acc_all[(size_t)w * NQ * J + i];
      const int q = i / J, j = i - q * J;
      if (q < 2) t *= p.D * tab[j].w;     // d l / d(b, a) carry D (d - c)
      p.item_partial[(size_t)blockIdx.x * NQ * J + i] = t;
    }
  }
}

// ---------------------------------------------------------------------------
constexpr int kUniMaxCtasPerSM = 8;

template <int MODEL, bool SCORE, int SC, bool FAST, bool PIPE, int UNR = 1, bool ESM = false>
static int launch_one_k(IrtUniParams& p, size_t smem, cudaStream_t st, int* grid_out) {
  auto k = irt_uni_kernel<MODEL, SCORE, SC, FAST, PIPE, UNR, ESM>;
  if (ESM) smem += (size_t)kUniThreads * SC * 4;
  static OccCache occ_cache;   // per instantiation
  int occ = 0;
  if (smem > 48 * 1024) VIPSY_CUDA_CHECK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  VIPSY_CUDA_CHECK(cached_occupancy(occ_cache, k, kUniThreads, smem, kUniMaxCtasPerSM, &occ));
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

// Pathwise estimator: pipelined item loop, two items per block when at most 8 particles sit in registers
// (measured, 1M x 100 4PL: S=32 2.105 -> 2.015 ms with one item per block, 2.047 with two; S=8 0.742 -> 0.700 /
// 0.635 ms; S=4 0.510 -> 0.489 / 0.449 ms; profiles/r02_uni_pipelined_ab.txt).  The REINFORCE estimator keeps
// the branchy loop.
template <int MODEL, bool SCORE, int SC, bool FAST>
static int launch_one(IrtUniParams& p, size_t smem, cudaStream_t st, int* grid_out) {
  if constexpr (SCORE) return launch_one_k<MODEL, SCORE, SC, FAST, false>(p, smem, st, grid_out);
  else {
    if constexpr (SC == 32) {
      // 32 particles: the noise pairs in shared memory leave room for a fifth CTA per SM (96 registers).  Worth it for
      // launches of at least four rounds of that grid (1M x 100 4PL: 2.007 -> 1.946 ms); shorter launches -- the chunks of
      // the host-buffer pipeline -- measured better with four CTAs (e2e 2.26 vs 2.32 ms), VIPSY_UNI_ESM=0 / 1 forces.
      static const int esm_env = [] { const char* e = getenv("VIPSY_UNI_ESM"); return e ? atoi(e) : -1; }();
      const bool fits5 = 5 * (smem + (size_t)kUniThreads * SC * 4 + 1024) <= (size_t)228 * 1024;   // short tests only
      const bool esm = fits5 && (esm_env >= 0 ? esm_env != 0 : p.B >= (int64_t)4 * num_sms() * 5 * kUniWarps * 32);
      if (esm) return launch_one_k<MODEL, SCORE, SC, FAST, true, 1, true>(p, smem, st, grid_out);
    }
    return launch_one_k<MODEL, SCORE, SC, FAST, true, (SC <= 8 ? 2 : 1)>(p, smem, st, grid_out);
  }
}

template <int MODEL, bool SCORE, bool FAST>
static int launch_sc(int SC, IrtUniParams& p, size_t smem, cudaStream_t st, int* grid) {
  switch (SC) {
    case 32: return launch_one<MODEL, SCORE, 32, FAST>(p, smem, st, grid);
    case 16: return launch_one<MODEL, SCORE, 16, FAST>(p, smem, st, grid);
    case 8: return launch_one<MODEL, SCORE, 8, FAST>(p, smem, st, grid);
    case 4: return launch_one<MODEL, SCORE, 4, FAST>(p, smem, st, grid);
    case 2: return launch_one<MODEL, SCORE, 2, FAST>(p, smem, st, grid);
    default: return launch_one<MODEL, SCORE, 1, FAST>(p, smem, st, grid);
  }
}

template <bool SCORE>
static int launch_model(int model, bool fast, int SC, IrtUniParams& p, size_t smem, cudaStream_t st, int* grid) {
  switch (model) {
    case VIPSY_IRT_1PL: return launch_sc<1, SCORE, false>(SC, p, smem, st, grid);
    case VIPSY_IRT_2PL: return launch_sc<2, SCORE, false>(SC, p, smem, st, grid);
    case VIPSY_IRT_3PL: return launch_sc<3, SCORE, false>(SC, p, smem, st, grid);
    default:
      return fast ? launch_sc<4, SCORE, true>(SC, p, smem, st, grid)
                  : launch_sc<4, SCORE, false>(SC, p, smem, st, grid);
  }
}

static size_t uni_loss_rows() { return (size_t)num_sms() * kUniMaxCtasPerSM + kMaskRows; }

size_t irt_uni_workspace_bytes(const VipsyIrtArgs* a) {
  const int rows = num_sms() * kUniMaxCtasPerSM;
  return align16(uni_loss_rows() * 8) + (size_t)rows * a->model * a->n_items * 4;
}

int irt_uni_launch(const VipsyIrtArgs* a, cudaStream_t st) {
  const int J = a->n_items, NQ = a->model;
  const size_t smem = uni_smem_bytes(J, NQ);
  VIPSY_REQUIRE(smem <= 227 * 1024, VIPSY_ERR_UNSUPPORTED,
                "vipsy_irt_elbo_grad(K=1): J=%d needs %zu B of shared memory (max 232448)", J, smem);
  VIPSY_REQUIRE(a->workspace && a->workspace_bytes >= irt_uni_workspace_bytes(a), VIPSY_ERR_WORKSPACE,
                "vipsy_irt_elbo_grad: workspace %zu B < required %zu B", a->workspace_bytes,
                irt_uni_workspace_bytes(a));
  IrtUniParams p;
  p.y = a->y; p.ld_y = a->y_row_stride; p.idx = a->idx;
  p.mu = a->mu; p.sigma = a->sigma; p.eps = a->eps;
  p.a = a->a; p.b = a->b; p.c = a->c; p.d = a->d;
  p.B = a->batch; p.J = J; p.S = a->n_particles;
  p.D = a->D; p.scale = a->scale; p.baseline = a->baseline;
  p.seed = a->seed; p.offset = a->offset; p.person_offset = a->person_offset;
  p.d_mu = a->d_mu; p.d_sigma = a->d_sigma;
  p.loss_partial = reinterpret_cast<double*>(a->workspace);
  p.item_partial = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(a->workspace) +
                                            align16(uni_loss_rows() * 8));
  p.contiguous = (a->idx == nullptr && a->y_row_stride == J) ? 1 : 0;
  p.packed2 = (a->flags & VIPSY_IRT_FLAG_Y_PACKED2) ? 1 : 0;
  p.want_grad = (a->d_b || a->d_a || a->d_c || a->d_d) ? 1 : 0;
  p.row_ptr = a->obs_row_ptr; p.entries = a->obs_entries;

  int SC = 32;
  while (SC > 1 && (a->n_particles % SC) != 0) SC >>= 1;
  const bool fast = a->model == VIPSY_IRT_4PL && !(a->flags & VIPSY_IRT_FLAG_STRICT);
  int grid = 0;
  int rc;
  const bool sparse = (a->flags & VIPSY_IRT_FLAG_SPARSE) && a->estimator == VIPSY_EST_PATHWISE && J < 32768 &&
                      irt_uni_sparse_smem_bytes(J, NQ) <= 227 * 1024;
  const bool want_sparse = (a->flags & VIPSY_IRT_FLAG_SPARSE) && a->estimator == VIPSY_EST_PATHWISE;
  // S > 1 with the observed-cell CSR at hand: the sparse kernel walks it directly (no staging, no compaction)
  const bool csr_multi = a->obs_row_ptr && a->obs_entries && a->n_particles > 1 &&
                         a->estimator == VIPSY_EST_PATHWISE && J < 32768 && !p.packed2;
  if (!csr_multi && !irt_uni_csr_supported(a)) {         // a CSR no kernel of this call can use (e.g. the score estimator)
    p.row_ptr = nullptr;
    p.entries = nullptr;
  }
  int mask_rows = 0;        // extra loss partials written by masked_term_kernel (lane-per-person dense kernel only)
  if (p.packed2) {                                                           // 2-bit rows: the pipelined particle kernel
    rc = launch_model<false>(a->model, fast, SC, p, smem, st, &grid);
    mask_rows = 1;
  } else if (irt_uni_csr_supported(a))
    rc = irt_uni_csr_launch(a, fast, p, st, &grid);                          // S == 1 over the observed-cell CSR
  else if (!(a->flags & VIPSY_IRT_FLAG_NO_ROW) && irt_uni_row_supported(a, want_sparse))
    rc = irt_uni_row_launch(a->model, fast, want_sparse, p, st, &grid);     // S == 1: warp per person
  else if (sparse || csr_multi)
    rc = irt_uni_sparse_launch(a->model, fast, SC, p, st, &grid);   // CSR given: its entries are the lists
  else {
    rc = a->estimator == VIPSY_EST_SCORE ? launch_model<true>(a->model, fast, SC, p, smem, st, &grid)
                                         : launch_model<false>(a->model, fast, SC, p, smem, st, &grid);
    mask_rows = 1;
  }
  if (rc != VIPSY_OK) return rc;
  if (mask_rows) {
    const int64_t known = (a->flags & VIPSY_IRT_FLAG_MASKED_KNOWN) ? a->n_masked_cells : -1;
    if (p.packed2)
      rc = launch_masked_term_2bit(a->y, a->batch * (int64_t)(J >> 2), a->scale, known, p.loss_partial + grid,
                                   &mask_rows, st);
    else
      rc = launch_masked_term(a->y, a->y_row_stride, a->idx, a->batch, J, a->scale, known, p.loss_partial + grid,
                              &mask_rows, st);
    if (rc != VIPSY_OK) return rc;
  }

  FinalizeSegs segs{};
  segs.n = NQ;
  float* outs[4] = {a->d_b, a->d_a, a->d_c, a->d_d};
  for (int q = 0; q < NQ; ++q) { segs.out[q] = outs[q]; segs.cols[q] = J; segs.mask[q] = nullptr; }
  if (NQ >= 2) segs.mask[1] = a->a_free;
  return launch_finalize(p.want_grad ? p.item_partial : nullptr, p.want_grad ? grid : 0, p.want_grad ? NQ * J : 0,
                         -a->scale / (float)a->n_particles, segs, p.loss_partial, grid + mask_rows, a->loss, st);
}

}  // namespace vipsy
