// K == 1, mostly-missing response matrices (BASELINE config 4: 2M persons x 200 items, 90 % NaN;
// the reference's masking rule vi.py:621-624 makes a missing cell contribute exactly nothing).
//
// The dense kernel (irt_uni.cu) walks all J items of a 32-person warp tile in lock step: with 10 %
// observed cells nearly every item still has SOME observed lane, so the warp pays the full evaluation
// for ~3 useful lanes.  Here each lane first compacts its own row -- (item, response) pairs of the
// observed cells, 16 bits each, in shared memory -- and the warp then runs max_lane(count) dense
// iterations: ~30 instead of 200 at 10 % density.  Lanes are on different items in the same
// iteration, so the per-item gradient sums cannot use the warp butterfly of the dense kernel; they
// are accumulated in 64-bit FIXED POINT (2^-32 resolution) with native 32-bit shared-memory atomics.
// Integer addition is associative: the result is bit-reproducible whatever the arrival order, and
// more accurate than an fp32 running sum.  Per-person sums (d mu, d sigma, log-lik) stay in the
// lane's registers in list order.
//
// CSR instantiation (S > 1 with the observed-cell CSR of irt_uni_csr.cu given): the lists already exist in global
// memory -- no tile staging, no compaction, and above all no 19 KB of shared memory per warp, so four CTAs fit an SM
// instead of two (the compaction variant runs two warps per scheduler).  A lane walks its person's entries straight
// from global memory, one entry ahead of its use (40 bytes per person: one or two cache lines, L1 hits after the first).
#include "irt_uni.cuh"

namespace vipsy {

constexpr float kFix = 4294967296.0f;          // 2^32
constexpr float kInvFix = 1.0f / 4294967296.0f;

size_t irt_uni_sparse_smem_bytes(int J, int NQ) {
  size_t s = (size_t)J * 16;                               // float4 item table
  s += align16((size_t)J * 8);                             // {D*a*(d-c), 1-d}
  s += align16((size_t)NQ * J * 8);                        // int64 fixed-point item accumulators (per CTA)
  s += (size_t)kUniWarps * align16((size_t)32 * J + 16);   // response tiles
  s += (size_t)kUniWarps * align16((size_t)32 * J * 2);    // compacted (item | response << 15) lists, [t][lane]
  return s;
}
static size_t irt_uni_sparse_csr_smem_bytes(int J, int NQ) {
  return (size_t)J * 16 + align16((size_t)J * 8) + align16((size_t)NQ * J * 8);
}

// 64-bit fixed-point accumulate built from native 32-bit shared-memory atomics: the low word wraps,
// every wrap (carry) and the addend's own high word go to the high word.  The pair (hi, lo) ends up
// as the exact 64-bit sum whatever the order of arrival (the number of wraps of a sum of unsigned
// words does not depend on the order), hence bit-reproducible.  |v| < 1 needs the second atomic only
// when the carry and the sign extension do not cancel.
__device__ __forceinline__ void fix_add(uint32_t* lo, int32_t* hi, float v) {
  const long long x = __float2ll_rn(v * kFix);
  const uint32_t xl = (uint32_t)x;
  int32_t xh = (int32_t)(x >> 32);
  const uint32_t old = atomicAdd(lo, xl);
  xh += (uint32_t)(old + xl) < old ? 1 : 0;
  if (xh != 0) atomicAdd(hi, xh);
}

template <int MODEL, int SC, bool FAST, bool CSR>
__global__ void __launch_bounds__(kUniThreads, CSR ? 4 : 3) irt_uni_sparse_kernel(const IrtUniParams p) {
  constexpr int NQ = MODEL;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int J = p.J;
  float4* tab = reinterpret_cast<float4*>(smem_raw);
  float2* tab2 = reinterpret_cast<float2*>(smem_raw + (size_t)J * 16);
  uint32_t* acc_lo = reinterpret_cast<uint32_t*>(reinterpret_cast<unsigned char*>(tab2) + align16((size_t)J * 8));
  int32_t* acc_hi = reinterpret_cast<int32_t*>(acc_lo + (size_t)NQ * J);
  unsigned char* ytiles = reinterpret_cast<unsigned char*>(acc_lo) + align16((size_t)NQ * J * 8);
  const size_t ytile_bytes = align16((size_t)32 * J + 16);
  unsigned char* lists = ytiles + (size_t)kUniWarps * ytile_bytes;
  const size_t list_bytes = align16((size_t)32 * J * 2);

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned char* ytile = ytiles + (size_t)warp * ytile_bytes;
  uint16_t* list = reinterpret_cast<uint16_t*>(lists + (size_t)warp * list_bytes);   // entry t of lane l: list[t*32 + l]

  for (int j = threadIdx.x; j < J; j += kUniThreads) {
    const float aj = MODEL >= 2 ? p.a[j] : 1.f;
    const float cj = MODEL >= 3 ? p.c[j] : 0.f;
    const float dj = MODEL == 4 ? p.d[j] : 1.f;
    tab[j] = make_float4(-p.D * kLog2e * aj, -p.D * kLog2e * p.b[j], cj, dj - cj);
    tab2[j] = make_float2(p.D * aj * (dj - cj), 1.f - dj);
  }
  for (int i = threadIdx.x; i < NQ * J; i += kUniThreads) { acc_lo[i] = 0u; acc_hi[i] = 0; }
  bool use_fast = false;
  if (FAST) {
    int ok = 1;
    for (int j = threadIdx.x; j < J; j += kUniThreads) {
      const float cj = p.c[j], dj = p.d[j];
      ok &= (cj >= 1e-3f) && (dj <= 1.f - 1e-3f) && (dj - cj >= 1e-3f);
    }
    use_fast = __syncthreads_and(ok) != 0;
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

    int cnt = 0;
    const uint16_t* ebase = nullptr;                     // CSR: the lane's entries in global memory
    if constexpr (CSR) {
      if (valid) {
        const int64_t row = p.idx ? __ldg(p.idx + bi) : bi;       // minibatch: the CSR is that of the whole matrix
        const int64_t r0 = p.row_ptr[row];
        cnt = (int)(p.row_ptr[row + 1] - r0);
        ebase = p.entries + r0;
      }
    } else {
    // ---- stage the tile's response codes (same as the dense kernel) ----------------------------------
    int mis = 0;
    if (p.contiguous) {
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

    // ---- compact this lane's row: (item | response << 15) of the observed cells, in item order --------
    if (valid) {
      const unsigned char* yrow = ytile + mis + lane * J;
      if (((mis + lane * J) & 3) == 0 && (J & 3) == 0) {
        const uint32_t* yw = reinterpret_cast<const uint32_t*>(yrow);
        for (int w = 0; w < (J >> 2); ++w) {
          const uint32_t word = yw[w];
          if (word != 0xffffffffu) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint32_t code = (word >> (8 * k)) & 0xffu;
              if (code != 255u) {
                list[cnt * 32 + lane] = (uint16_t)((4 * w + k) | (code != 0u ? 0x8000u : 0u));
                ++cnt;
              }
            }
          }
        }
      } else {
        for (int j = 0; j < J; ++j) {
          const uint32_t code = yrow[j];
          if (code != 255u) {
            list[cnt * 32 + lane] = (uint16_t)(j | (code != 0u ? 0x8000u : 0u));
            ++cnt;
          }
        }
      }
    }
    }   // !CSR
    int maxcnt = cnt;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) maxcnt = max(maxcnt, __shfl_xor_sync(0xffffffffu, maxcnt, o));

    const float mu = valid ? p.mu[bi] : 0.f;
    const float sg = valid ? p.sigma[bi] : 1.f;
    float se = 0.f, see = 0.f, G1 = 0.f, G2 = 0.f, Lsum = 0.f;

    for (int s0 = 0; s0 < p.S; s0 += SC) {
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
      float lls[1];

      // ---- observed cells only; U independent cells per iteration so that their dependency chains
      //      (table look-up -> ex2 -> rcp -> ...) overlap when there are few particles ---------------------
      constexpr int U = SC <= 2 ? 4 : (SC <= 8 ? 2 : 1);
      uint32_t ahead[U];                                   // CSR: entries of the next step, already in flight
      if constexpr (CSR) {
#pragma unroll
        for (int u = 0; u < U; ++u) ahead[u] = u < cnt ? (uint32_t)__ldg(ebase + u) : 0u;
      }
      for (int t0 = 0; t0 < maxcnt; t0 += U) {
        int jj[U];
        bool yy[U], act[U];
        float w1[U], w2[U], a1[U], a2[U], lj[U], dm[U];
        uint32_t cur[U];
        if constexpr (CSR) {
#pragma unroll
          for (int u = 0; u < U; ++u) {
            cur[u] = ahead[u];
            ahead[u] = t0 + U + u < cnt ? (uint32_t)__ldg(ebase + t0 + U + u) : 0u;
          }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          act[u] = t0 + u < cnt;
          uint32_t ent;
          if constexpr (CSR) ent = cur[u];
          else ent = act[u] ? list[(t0 + u) * 32 + lane] : 0u;              // inactive slots evaluate item 0, discarded
          const int j = (int)(ent & 0x7fffu);
          const bool y1 = (ent & 0x8000u) != 0u;
          const float4 it = tab[j];
          const float2 it2 = tab2[j];
          const float A = y1 ? it.x : -it.x, Bc = y1 ? it.y : -it.y;
          const float T0 = fmaf(A, mu, Bc), T1 = A * sg;
          const float base = MODEL >= 3 ? (y1 ? it.z : it2.y) : 0.f;
          float W1, W2, A1, A2, Lj = 0.f;
          if (FAST && use_fast)
            item_evals<MODEL, false, SC, true>(it.w, base, T0, T1, eps, lls, W1, W2, A1, A2, Lj);
          else
            item_evals<MODEL, false, SC, false>(it.w, base, T0, T1, eps, lls, W1, W2, A1, A2, Lj);
          const float sgn = y1 ? 1.f : -1.f;
          jj[u] = j; yy[u] = y1; w1[u] = W1 * sgn; w2[u] = W2 * sgn; a1[u] = A1; a2[u] = A2; lj[u] = Lj; dm[u] = it2.x;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          if (act[u]) {
            const int j = jj[u];
            Lsum += lj[u];
            G1 = fmaf(dm[u], w1[u], G1);
            G2 = fmaf(dm[u], w2[u], G2);
            if (p.want_grad) {
              fix_add(acc_lo + j, acc_hi + j, w1[u]);
              if constexpr (NQ >= 2) fix_add(acc_lo + J + j, acc_hi + J + j, fmaf(sg, w2[u], mu * w1[u]));
              if constexpr (NQ >= 3) fix_add(acc_lo + 2 * J + j, acc_hi + 2 * J + j, yy[u] ? a2[u] : -a1[u]);
              if constexpr (NQ >= 4) fix_add(acc_lo + 3 * J + j, acc_hi + 3 * J + j, yy[u] ? a1[u] : -a2[u]);
            }
          }
        }
      }
    }
    __syncwarp();

    if (valid) {
      const float sumx = fmaf(sg, se, S_f * mu);
      const float sumxe = fmaf(sg, see, mu * se);
      const float sumx2 = fmaf(S_f * mu, mu, fmaf(2.f * mu * sg, se, sg * sg * see));
      const float elbo = kLn2 * Lsum - 0.5f * sumx2 + 0.5f * see + S_f * (logf(sg) + (float)(J - cnt) * kMaskedLogProb);
      loss_acc -= (double)(k_scale * elbo);
      if (p.d_mu) {
        p.d_mu[bi] = -k_scale * (G1 - sumx);
        p.d_sigma[bi] = -k_scale * (G2 - sumxe) - p.scale / sg;
      }
    }
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
      // fixed point -> float: integer part + fraction
      float t = (float)acc_hi[i] + (float)acc_lo[i] * kInvFix;
      const int q = i / J, j = i - q * J;
      if (q < 2) t *= p.D * tab[j].w;     // d l / d(b, a) carry D (d - c)
      p.item_partial[(size_t)blockIdx.x * NQ * J + i] = t;
    }
  }
}

template <int MODEL, int SC, bool FAST, bool CSR>
static int launch_sparse_k(IrtUniParams& p, size_t smem, cudaStream_t st, int* grid_out) {
  auto k = irt_uni_sparse_kernel<MODEL, SC, FAST, CSR>;
  static OccCache occ_cache;   // per instantiation
  int occ = 0;
  if (smem > 48 * 1024) VIPSY_CUDA_CHECK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  VIPSY_CUDA_CHECK(cached_occupancy(occ_cache, k, kUniThreads, smem, 8, &occ));
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

template <int MODEL, int SC, bool FAST>
static int launch_sparse_one(IrtUniParams& p, size_t smem, cudaStream_t st, int* grid_out) {
  if (p.row_ptr && p.entries)
    return launch_sparse_k<MODEL, SC, FAST, true>(p, irt_uni_sparse_csr_smem_bytes(p.J, MODEL), st, grid_out);
  return launch_sparse_k<MODEL, SC, FAST, false>(p, smem, st, grid_out);
}

template <int MODEL, bool FAST>
static int launch_sparse_sc(int SC, IrtUniParams& p, size_t smem, cudaStream_t st, int* grid) {
  switch (SC) {
    case 32: return launch_sparse_one<MODEL, 32, FAST>(p, smem, st, grid);
    case 16: return launch_sparse_one<MODEL, 16, FAST>(p, smem, st, grid);
    case 8: return launch_sparse_one<MODEL, 8, FAST>(p, smem, st, grid);
    case 4: return launch_sparse_one<MODEL, 4, FAST>(p, smem, st, grid);
    case 2: return launch_sparse_one<MODEL, 2, FAST>(p, smem, st, grid);
    default: return launch_sparse_one<MODEL, 1, FAST>(p, smem, st, grid);
  }
}

int irt_uni_sparse_launch(int model, bool fast, int SC, IrtUniParams& p, cudaStream_t st, int* grid_out) {
  const size_t smem = irt_uni_sparse_smem_bytes(p.J, model);
  switch (model) {
    case VIPSY_IRT_1PL: return launch_sparse_sc<1, false>(SC, p, smem, st, grid_out);
    case VIPSY_IRT_2PL: return launch_sparse_sc<2, false>(SC, p, smem, st, grid_out);
    case VIPSY_IRT_3PL: return launch_sparse_sc<3, false>(SC, p, smem, st, grid_out);
    default:
      return fast ? launch_sparse_sc<4, true>(SC, p, smem, st, grid_out)
                  : launch_sparse_sc<4, false>(SC, p, smem, st, grid_out);
  }
}

}  // namespace vipsy
