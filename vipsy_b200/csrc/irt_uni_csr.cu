// K == 1, one particle, responses given as a CSR of the OBSERVED cells (BASELINE config 4: 2M persons x
// 200 items, 90 % NaN).  The reference keeps the NaN-coded float matrix and masks it on every step
// (vi.py:617-625); the response matrix never changes between steps, so the compaction is done once
// (vipsy_observed_row_ptr / vipsy_observed_fill below, the "CSR-of-observed" layout of SURVEY 8f-3) and the
// per-step kernel streams 2 bytes per observed cell instead of 1 byte per cell: c4 moves 80 + 16 MB of
// response data per step instead of 400 MB, and does arithmetic only where there is a response.
//
// Mapping: a sub-warp of SW lanes (8, 16 or 32, chosen from the mean row length) owns a person; its
// lanes take that person's entries SW at a time.  32 / SW persons are in flight per warp instruction, so
// short rows still fill the lanes.  Entries of one person are distinct items, and every sub-warp has its
// own shared-memory item accumulators: plain read-modify-writes in a fixed order, deterministic, no atomics.
#include <algorithm>

#include "irt_uni.cuh"

namespace vipsy {

// shared-memory accesses by 32-bit shared-window address: keeps the compiler from re-deriving the window base
// (S2UR SR_CgaCtaId + 3 uniform ops) in front of every access of the hot loop
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ float2 lds_f2(uint32_t a) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void sts_f2(uint32_t a, float2 v) {
  asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(a), "f"(v.x), "f"(v.y) : "memory");
}
__device__ __forceinline__ float lds_f1(uint32_t a) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void sts_f1(uint32_t a, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory"); }

// accumulator row of one item: NQ sums padded to 1, 2 or 4 floats so that one LDS/STS moves the row
__host__ __device__ constexpr int csr_acc_width(int NQ) { return NQ <= 2 ? NQ : 4; }
// item table: {A', B'} (1PL / 2PL; D a = -A' / log2 e) or {A', B', c, d-c} + {D a (d-c), 1-d}
__host__ __device__ inline size_t csr_tab_bytes(int J, int NQ) {
  return NQ <= 2 ? align16((size_t)J * 8) : (size_t)J * 16 + align16((size_t)J * 8);
}
__host__ __device__ inline size_t csr_smem_bytes(int J, int NQ, int SW) {
  return csr_tab_bytes(J, NQ) + (size_t)kUniWarps * (32 / SW) * align16((size_t)csr_acc_width(NQ) * J * 4);
}

// per-lane view of one person of a 32-person tile
struct CsrLane {
  float mu, sg, eps, x;
  int64_t beg;
  int cnt;
  bool valid;
};
__device__ __forceinline__ CsrLane csr_lane(const IrtUniParams& p, int64_t bi) {
  CsrLane t;
  t.valid = bi < p.B;
  t.mu = 0.f; t.sg = 1.f; t.beg = 0; t.cnt = 0;
  if (t.valid) {
    t.mu = p.mu[bi];
    t.sg = p.sigma[bi];
    const int64_t row = p.idx ? __ldg(p.idx + bi) : bi;
    t.beg = __ldg(p.row_ptr + row);
    t.cnt = (int)(__ldg(p.row_ptr + row + 1) - t.beg);
  }
  if (p.eps) t.eps = t.valid ? __ldg(p.eps + bi) : 0.f;
  else t.eps = t.valid ? normal4(p.seed, p.offset, (uint64_t)(p.person_offset + bi), 0u).x : 0.f;
  t.x = fmaf(t.sg, t.eps, t.mu);
  return t;
}

template <int MODEL, int SW, bool FAST>
__global__ void __launch_bounds__(kUniThreads) irt_uni_csr_kernel(const IrtUniParams p) {
  constexpr int NQ = MODEL;
  constexpr int AW = csr_acc_width(NQ);
  constexpr int NSUB = 32 / SW;
  constexpr int NBP = SW == 8 ? 3 : 2;          // batches of the next round whose entries are loaded ahead
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int J = p.J;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sub = lane / SW, sl = lane % SW;
  float2* tabAB = reinterpret_cast<float2*>(smem_raw);                                 // MODEL <= 2
  float4* tab = reinterpret_cast<float4*>(smem_raw);                                   // MODEL >= 3
  float2* tab2 = reinterpret_cast<float2*>(smem_raw + (size_t)J * 16);
  unsigned char* q0 = smem_raw + csr_tab_bytes(J, NQ);
  const size_t acc_bytes = align16((size_t)AW * J * 4);
  float* acc_all = reinterpret_cast<float*>(q0);
  float* acc = reinterpret_cast<float*>(q0 + (size_t)(warp * NSUB + sub) * acc_bytes);   // [J][AW], sub-warp-private

  int ok = 1;
  for (int j = threadIdx.x; j < J; j += kUniThreads) {
    const float aj = MODEL >= 2 ? p.a[j] : 1.f;
    const float cj = MODEL >= 3 ? p.c[j] : 0.f;
    const float dj = MODEL == 4 ? p.d[j] : 1.f;
    if constexpr (MODEL <= 2) {
      tabAB[j] = make_float2(-p.D * kLog2e * aj, -p.D * kLog2e * p.b[j]);
    } else {
      tab[j] = make_float4(-p.D * kLog2e * aj, -p.D * kLog2e * p.b[j], cj, dj - cj);
      tab2[j] = make_float2(p.D * aj * (dj - cj), 1.f - dj);
    }
    if (FAST) ok &= (cj >= 1e-3f) && (dj <= 1.f - 1e-3f) && (dj - cj >= 1e-3f);
  }
  for (int i = threadIdx.x; i < (int)(kUniWarps * NSUB * acc_bytes / 4); i += kUniThreads) acc_all[i] = 0.f;
  const bool use_fast = FAST ? (__syncthreads_and(ok) != 0) : (__syncthreads(), false);

  double loss_acc = 0.0;
  const int64_t ntiles = (p.B + 31) / 32;
  const int64_t gwarp = (int64_t)blockIdx.x * kUniWarps + warp;
  const int64_t nwarps = (int64_t)gridDim.x * kUniWarps;

  // one batch: this lane's entry `ent` (valid iff obs) of the person whose latent is x
  float Gi, Lr;
  const uint32_t tab_s = smem_addr(tabAB), acc_s = smem_addr(acc);
  const bool want_grad = p.want_grad != 0;
  auto batch = [&](uint32_t ent, bool obs, float x) {
    const int j = (int)(ent & 0x7fffu);
    if constexpr (MODEL <= 2) {
      // lean 1PL / 2PL form: sign flip by XOR, one select puts an idle lane at t = -126 (sp = 1, 1 - sp = 0),
      // the probability clamp is the min at t = 23 (sp = eps) with the gradient switched off beyond it
      const float2 ab = lds_f2(tab_s + 8u * (uint32_t)j);
      const uint32_t flip = (~ent << 16) & 0x80000000u;                  // y == 0: negate
      float t = __uint_as_float(__float_as_uint(fmaf(ab.x, x, ab.y)) ^ flip);
      t = obs ? t : -126.f;
      const float e = ex2_approx(fminf(t, 23.f));
      const float sp = rcp_approx(e + 1.f);
      const float w = t < 23.f ? e * sp : 0.f;
      const float ws = __uint_as_float(__float_as_uint(w) ^ flip);
      Gi = fmaf(ab.x, ws, Gi);
      Lr *= sp;
      if (obs && want_grad) {
        if constexpr (NQ == 2) {
          const uint32_t ap = acc_s + 8u * (uint32_t)j;
          float2 v = lds_f2(ap);
          v.x += ws;
          v.y = fmaf(ws, x, v.y);
          sts_f2(ap, v);
        } else {
          const uint32_t ap = acc_s + 4u * (uint32_t)j;
          sts_f1(ap, lds_f1(ap) + ws);
        }
      }
    } else {
      const bool y1 = (ent & 0x8000u) != 0u;
      const float4 it = tab[j];
      const float2 it2 = tab2[j];
      const float tt0 = fmaf(it.x, x, it.y);
      const float t = y1 ? tt0 : -tt0;
      const float base = y1 ? it.z : it2.y;
      float w, a1, a2, qc;
      if (FAST && use_fast) cell_eval1<MODEL, true>(t, it.w, base, obs, w, a1, a2, qc);
      else cell_eval1<MODEL, false>(t, it.w, base, obs, w, a1, a2, qc);
      const float ws = y1 ? w : -w;
      Gi = fmaf(it2.x, ws, Gi);
      Lr *= qc;
      if (obs && want_grad) {
        float4* ap = reinterpret_cast<float4*>(acc) + j;
        float4 v = *ap;
        v.x += ws;
        v.y = fmaf(ws, x, v.y);
        v.z += y1 ? a2 : -a1;
        v.w += y1 ? a1 : -a2;
        *ap = v;
      }
    }
  };

  CsrLane cur = csr_lane(p, gwarp * 32 + lane);
  for (int64_t tile = gwarp; tile < ntiles; tile += nwarps) {
    const int64_t bi = tile * 32 + lane;
    const CsrLane nxt = csr_lane(p, (tile + nwarps) * 32 + lane);     // next tile's inputs, in flight during this one
    float Gmine = 0.f, Ltile = 0.f;

    // ---- SW rounds, NSUB persons per round; the first NBP batches of the next round are loaded ahead ------
    float x = __shfl_sync(0xffffffffu, cur.x, sub);
    int cnt = __shfl_sync(0xffffffffu, cur.cnt, sub);
    const uint16_t* pe = p.entries + __shfl_sync(0xffffffffu, cur.beg, sub) + sl;
    uint32_t ent[NBP];
#pragma unroll
    for (int b = 0; b < NBP; ++b) ent[b] = __ldg(b * SW + sl < cnt ? pe + b * SW : p.entries);
#pragma unroll 1
    for (int r = 0; r < SW; ++r) {
      const int pn = (r + 1 < SW ? (r + 1) * NSUB : 0) + sub;
      const float xn = __shfl_sync(0xffffffffu, cur.x, pn);
      const int cntn = r + 1 < SW ? __shfl_sync(0xffffffffu, cur.cnt, pn) : 0;
      const uint16_t* pen = p.entries + __shfl_sync(0xffffffffu, cur.beg, pn) + sl;
      uint32_t entn[NBP];
#pragma unroll
      // unconditional loads (an idle lane re-reads entry 0, a valid entry it then ignores): the loaded register
      // is not touched until the next round, so the load latency hides behind this round's batches
      for (int b = 0; b < NBP; ++b) entn[b] = __ldg(b * SW + sl < cntn ? pen + b * SW : p.entries);

      int cmax = cnt;                                  // longest row among this round's NSUB persons
#pragma unroll
      for (int o = SW; o < 32; o <<= 1) cmax = max(cmax, __shfl_xor_sync(0xffffffffu, cmax, o));
      Gi = 0.f;
      Lr = 1.f;
#pragma unroll
      for (int b = 0; b < NBP; ++b)
        if (b * SW < cmax) batch(ent[b], b * SW + sl < cnt, x);
      Ltile += lg2_approx(Lr);                         // <= 3 factors >= eps each: no underflow
      for (int t0 = NBP * SW; t0 < cmax; t0 += SW) {   // rows longer than the look-ahead
        const bool obs = t0 + sl < cnt;
        Lr = 1.f;
        batch(obs ? (uint32_t)__ldg(pe + t0) : 0u, obs, x);
        Ltile += lg2_approx(Lr);
      }
      __syncwarp();                                    // orders the accumulator updates of consecutive persons
#pragma unroll
      for (int o = SW / 2; o > 0; o >>= 1) Gi += __shfl_xor_sync(0xffffffffu, Gi, o);
      const float v = __shfl_sync(0xffffffffu, Gi, (lane % NSUB) * SW);    // person r*NSUB + s was done by sub-warp s
      if (lane / NSUB == r) Gmine = v;
      x = xn; cnt = cntn; pe = pen;
#pragma unroll
      for (int b = 0; b < NBP; ++b) ent[b] = entn[b];
    }
    if constexpr (MODEL <= 2) Gmine *= -kLn2;          // D a = -A' / log2(e)

    float prior = 0.f;
    if (cur.valid) {
      prior = -0.5f * cur.x * cur.x + 0.5f * cur.eps * cur.eps + logf(cur.sg) + (float)(J - cur.cnt) * kMaskedLogProb;
      if (p.d_mu) {
        p.d_mu[bi] = -p.scale * (Gmine - cur.x);
        p.d_sigma[bi] = -p.scale * (Gmine - cur.x) * cur.eps - p.scale / cur.sg;
      }
    }
    loss_acc -= (double)(p.scale * fmaf(kLn2, Ltile, prior));
    cur = nxt;
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
      const int q = i / J, j = i - q * J;
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < kUniWarps * NSUB; ++w) t += acc_all[(size_t)w * (acc_bytes / 4) + j * AW + q];
      if (q < 2) {
        const float cj = MODEL >= 3 ? p.c[j] : 0.f, dj = MODEL == 4 ? p.d[j] : 1.f;
        t *= p.D * (dj - cj);
      }
      p.item_partial[(size_t)blockIdx.x * NQ * J + i] = t;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
template <int MODEL, int SW, bool FAST>
static int launch_csr_one(IrtUniParams& p, cudaStream_t st, int* grid_out) {
  auto k = irt_uni_csr_kernel<MODEL, SW, FAST>;
  static OccCache occ_cache;   // per instantiation
  int occ = 0;
  const size_t smem = csr_smem_bytes(p.J, MODEL, SW);
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

template <int MODEL, bool FAST>
static int launch_csr_sw(int SW, IrtUniParams& p, cudaStream_t st, int* grid) {
  switch (SW) {
    case 8: return launch_csr_one<MODEL, 8, FAST>(p, st, grid);
    case 16: return launch_csr_one<MODEL, 16, FAST>(p, st, grid);
    default: return launch_csr_one<MODEL, 32, FAST>(p, st, grid);
  }
}

bool irt_uni_csr_supported(const VipsyIrtArgs* a) {
  return a->obs_row_ptr && a->obs_entries && a->n_particles == 1 && a->estimator == VIPSY_EST_PATHWISE &&
         a->n_items <= 32768 && csr_smem_bytes(a->n_items, a->model, 32) <= 200 * 1024;
}

int irt_uni_csr_launch(const VipsyIrtArgs* a, bool fast, IrtUniParams& p, cudaStream_t st, int* grid_out) {
  // sub-warp width from the mean row length (obs_nnz / n_rows); fall back to wider sub-warps when the
  // per-sub-warp accumulators would not fit
  const double mean = a->n_rows > 0 && a->obs_nnz > 0 ? (double)a->obs_nnz / (double)a->n_rows : (double)a->n_items;
  int SW = mean <= 28.0 ? 8 : (mean <= 80.0 ? 16 : 32);
  while (SW < 32 && csr_smem_bytes(a->n_items, a->model, SW) > 48 * 1024) SW *= 2;
  switch (a->model) {
    case VIPSY_IRT_1PL: return launch_csr_sw<1, false>(SW, p, st, grid_out);
    case VIPSY_IRT_2PL: return launch_csr_sw<2, false>(SW, p, st, grid_out);
    case VIPSY_IRT_3PL: return launch_csr_sw<3, false>(SW, p, st, grid_out);
    default:
      return fast ? launch_csr_sw<4, true>(SW, p, st, grid_out) : launch_csr_sw<4, false>(SW, p, st, grid_out);
  }
}

// ---------------------------------------------------------------------------------------------------------
// one-time construction of the layout: counts -> exclusive scan -> fill
// ---------------------------------------------------------------------------------------------------------
__global__ void observed_count_kernel(const uint8_t* __restrict__ y, int64_t n_rows, int J, int64_t ld,
                                      int32_t* __restrict__ counts) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = warp; r < n_rows; r += nw) {
    const uint8_t* g = y + r * ld;
    int c = 0;
    for (int j = lane; j < J; j += 32) c += g[j] != 255;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if (lane == 0) counts[r] = c;
  }
}

// row_ptr[i] = sum_{r < i} counts[r], i = 0..n; one CTA, each thread owns a contiguous chunk
__global__ void __launch_bounds__(1024) exclusive_scan_kernel(const int32_t* __restrict__ counts, int64_t n,
                                                              int64_t* __restrict__ row_ptr) {
  __shared__ int64_t part[1024];
  const int t = threadIdx.x;
  const int64_t chunk = (n + 1023) / 1024;
  const int64_t lo = min(n, (int64_t)t * chunk), hi = min(n, lo + chunk);
  int64_t s = 0;
  for (int64_t i = lo; i < hi; ++i) s += counts[i];
  part[t] = s;
  __syncthreads();
  for (int o = 1; o < 1024; o <<= 1) {           // Hillis-Steele inclusive scan of the 1024 chunk sums
    const int64_t v = t >= o ? part[t - o] : 0;
    __syncthreads();
    part[t] += v;
    __syncthreads();
  }
  int64_t run = part[t] - s;
  for (int64_t i = lo; i < hi; ++i) { row_ptr[i] = run; run += counts[i]; }
  if (t == 1023) row_ptr[n] = part[1023];
}

__global__ void observed_fill_kernel(const uint8_t* __restrict__ y, int64_t n_rows, int J, int64_t ld,
                                     const int64_t* __restrict__ row_ptr, uint16_t* __restrict__ entries) {
  const int lane = threadIdx.x & 31;
  const uint32_t lt = (1u << lane) - 1u;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = warp; r < n_rows; r += nw) {
    const uint8_t* g = y + r * ld;
    int64_t o = row_ptr[r];
    for (int j0 = 0; j0 < J; j0 += 32) {
      const int j = j0 + lane;
      const uint32_t code = j < J ? g[j] : 255u;
      const bool obs = code != 255u;
      const uint32_t m = __ballot_sync(0xffffffffu, obs);
      if (obs) entries[o + __popc(m & lt)] = (uint16_t)(j | (code != 0u ? 0x8000u : 0u));
      o += __popc(m);
    }
  }
}

}  // namespace vipsy

using namespace vipsy;

extern "C" {

int vipsy_observed_row_ptr(const uint8_t* y, int64_t n_rows, int32_t n_items, int64_t y_row_stride,
                           int32_t* counts_scratch, int64_t* row_ptr, void* stream) {
  VIPSY_REQUIRE(y && counts_scratch && row_ptr && n_rows >= 0 && n_items > 0 && n_items <= 32768 &&
                    y_row_stride >= n_items,
                VIPSY_ERR_ARG, "vipsy_observed_row_ptr: bad arguments (J must be <= 32768)");
  cudaStream_t st = (cudaStream_t)stream;
  if (n_rows > 0) {
    const int blocks = (int)std::min<int64_t>((n_rows + 7) / 8, (int64_t)num_sms() * 16);
    observed_count_kernel<<<blocks, 256, 0, st>>>(y, n_rows, n_items, y_row_stride, counts_scratch);
  }
  exclusive_scan_kernel<<<1, 1024, 0, st>>>(counts_scratch, n_rows, row_ptr);
  VIPSY_CUDA_CHECK(cudaGetLastError());
  return VIPSY_OK;
}

int vipsy_observed_fill(const uint8_t* y, int64_t n_rows, int32_t n_items, int64_t y_row_stride,
                        const int64_t* row_ptr, uint16_t* entries, void* stream) {
  VIPSY_REQUIRE(y && row_ptr && entries && n_rows >= 0 && n_items > 0 && n_items <= 32768 && y_row_stride >= n_items,
                VIPSY_ERR_ARG, "vipsy_observed_fill: bad arguments");
  if (n_rows == 0) return VIPSY_OK;
  const int blocks = (int)std::min<int64_t>((n_rows + 7) / 8, (int64_t)num_sms() * 16);
  observed_fill_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(y, n_rows, n_items, y_row_stride, row_ptr, entries);
  VIPSY_CUDA_CHECK(cudaGetLastError());
  return VIPSY_OK;
}

}  // extern "C"
