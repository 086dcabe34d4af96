// Error reporting, deterministic partial-sum finalisation and the data-format
// helpers either side of the ELBO path (response packing, eta table, noise fill).
#include <stdarg.h>
#include <string.h>

#include <algorithm>

#include "common.cuh"

namespace vipsy {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// ---------------------------------------------------------------------------
// out[c] = alpha * sum_r partial[r][c], rows summed in a fixed order with fp64
// accumulation.  One CTA per 32 columns; warp w walks rows w, w+8, ...
// (coalesced 128-byte row segments), the 8 warp totals are combined in order.
// ---------------------------------------------------------------------------
struct FinalizeArgs {
  const float* partial;
  int rows, cols;
  float alpha;
  FinalizeSegs segs;
  const double* loss_partial;
  int loss_rows;
  double* loss_out;
};

constexpr int kFinWarps = 32;   // at most; 8 below 256 partial rows (launch_finalize)

__global__ void __launch_bounds__(kFinWarps * 32) finalize_kernel(const FinalizeArgs a) {
  __shared__ double sm[kFinWarps][33];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nwarps = blockDim.x >> 5;                  // 8 or 32: many partial rows want more loads in flight
  const int col = blockIdx.x * 32 + lane;
  if (blockIdx.x * 32 < a.cols) {
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    bool wanted = false;                               // columns of a segment without an output are never written
    {                                                  // by the producers either: do not read them
      int c = col;
      for (int i = 0; i < a.segs.n && col < a.cols; ++i) {
        if (c < a.segs.cols[i]) { wanted = a.segs.out[i] != nullptr; break; }
        c -= a.segs.cols[i];
      }
    }
    if (col < a.cols && wanted) {
      int r = warp;
      for (; r + 3 * nwarps < a.rows; r += 4 * nwarps) {
        const float v0 = a.partial[(size_t)r * a.cols + col];
        const float v1 = a.partial[(size_t)(r + nwarps) * a.cols + col];
        const float v2 = a.partial[(size_t)(r + 2 * nwarps) * a.cols + col];
        const float v3 = a.partial[(size_t)(r + 3 * nwarps) * a.cols + col];
        s0 += v0; s1 += v1; s2 += v2; s3 += v3;
      }
      for (; r < a.rows; r += nwarps) s0 += a.partial[(size_t)r * a.cols + col];
    }
    sm[warp][lane] = (s0 + s1) + (s2 + s3);
    __syncthreads();
    if (warp == 0 && col < a.cols) {
      double t = 0.0;
      for (int w = 0; w < nwarps; ++w) t += sm[w][lane];
      int c = col;
      for (int i = 0; i < a.segs.n; ++i) {
        if (c < a.segs.cols[i]) {
          if (a.segs.out[i]) {
            float v = a.alpha * (float)t;
            if (a.segs.mask[i] && !a.segs.mask[i][c]) v = 0.f;
            a.segs.out[i][c] = v;
          }
          break;
        }
        c -= a.segs.cols[i];
      }
    }
  }
  if (blockIdx.x == gridDim.x - 1 && a.loss_out) {   // last CTA also folds the loss partials
    __syncthreads();
    double t = 0.0;
    for (int r = threadIdx.x; r < a.loss_rows; r += blockDim.x) t += a.loss_partial[r];
    t = warp_sum(t);
    if (lane == 0) sm[warp][32] = t;
    __syncthreads();
    if (threadIdx.x == 0) {
      double tot = 0.0;
      for (int w = 0; w < nwarps; ++w) tot += sm[w][32];
      *a.loss_out = tot;
    }
  }
}

int launch_finalize(const float* partial, int rows, int cols, float alpha, const FinalizeSegs& segs,
                    const double* loss_partial, int loss_rows, double* loss_out, cudaStream_t stream) {
  FinalizeArgs a{partial, rows, cols, alpha, segs, loss_partial, loss_rows, loss_out};
  int grid = (cols + 31) / 32;
  if (grid < 1) grid = 1;
  finalize_kernel<<<grid, (rows >= 256 ? kFinWarps : 8) * 32, 0, stream>>>(a);
  VIPSY_CUDA_CHECK(cudaGetLastError());
  return VIPSY_OK;
}

// ---------------------------------------------------------------------------
// Masked cells score log(1 - eps) each (kMaskedLogProb) and nothing else.  The K = 1 particle kernel sits at its
// register cap (any extra work in it costs 3-5 % through re-planned register allocation, measured) and the
// K > 1 / enumeration kernels never look at a masked cell at all, so this constant of the batch is produced by a
// side kernel: CTA b writes -scale * log(1-eps) * (masked cells of its rows) as one more fp64 loss partial;
// integer counts, fixed order => deterministic.  With the count known to the caller (the response matrix never
// changes between steps) one thread just writes it.
// ---------------------------------------------------------------------------
__device__ __forceinline__ int missing_in(const int4& v) {      // codes are 0 / 1 / 255: bit 1 marks a missing cell
  return __popc((uint32_t)v.x & 0x02020202u) + __popc((uint32_t)v.y & 0x02020202u) +
         __popc((uint32_t)v.z & 0x02020202u) + __popc((uint32_t)v.w & 0x02020202u);
}
__global__ void __launch_bounds__(256) masked_term_kernel(const uint8_t* __restrict__ y, int64_t ld,
                                                          const int64_t* __restrict__ idx, int64_t B, int J,
                                                          float scale, int64_t known, double* __restrict__ out) {
  if (known >= 0) {
    if (threadIdx.x == 0) out[blockIdx.x] = blockIdx.x == 0 ? -(double)scale * (double)kMaskedLogProb * (double)known : 0.0;
    return;
  }
  __shared__ long long part[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  long long n = 0;
  if (idx == nullptr && ld == J) {
    // contiguous batch: one flat byte range, 16-byte loads, four in flight per thread
    const int64_t total = B * (int64_t)J;
    const uintptr_t base = reinterpret_cast<uintptr_t>(y);
    int64_t head = (int64_t)((16 - (base & 15)) & 15);
    if (head > total) head = total;
    const int64_t nvec = (total - head) >> 4;
    const int4* v = reinterpret_cast<const int4*>(y + head);
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (int64_t)gridDim.x * blockDim.x;
    int64_t i = tid;
    for (; i + 3 * nth < nvec; i += 4 * nth) {
      const int4 a = __ldg(v + i), b = __ldg(v + i + nth), c = __ldg(v + i + 2 * nth), d = __ldg(v + i + 3 * nth);
      n += missing_in(a) + missing_in(b) + missing_in(c) + missing_in(d);
    }
    for (; i < nvec; i += nth) n += missing_in(__ldg(v + i));
    if (blockIdx.x == 0) {
      for (int64_t k = threadIdx.x; k < head; k += blockDim.x) n += y[k] == 255;
      for (int64_t k = head + (nvec << 4) + threadIdx.x; k < total; k += blockDim.x) n += y[k] == 255;
    }
  } else {
    for (int64_t r = (int64_t)blockIdx.x * 8 + warp; r < B; r += (int64_t)gridDim.x * 8) {
      const uint8_t* g = y + (idx ? idx[r] : r) * ld;
      if (((reinterpret_cast<uintptr_t>(g) | (uintptr_t)J) & 3) == 0) {
        const uint32_t* gw = reinterpret_cast<const uint32_t*>(g);
        for (int w = lane; w < (J >> 2); w += 32) n += __popc(__ldg(gw + w) & 0x02020202u);
      } else {
        for (int j = lane; j < J; j += 32) n += g[j] == 255;
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) n += __shfl_xor_sync(0xffffffffu, n, o);
  if (lane == 0) part[warp] = n;
  __syncthreads();
  if (threadIdx.x == 0) {
    long long t = 0;
    for (int w = 0; w < 8; ++w) t += part[w];
    out[blockIdx.x] = -(double)scale * (double)kMaskedLogProb * (double)t;
  }
}

int launch_masked_term(const uint8_t* y, int64_t ld, const int64_t* idx, int64_t B, int J, float scale, int64_t known,
                       double* out_rows, int* n_rows, cudaStream_t st) {
  int rows = known >= 0 ? 1 : (int)((B + 63) / 64 < kMaskRows ? (B + 63) / 64 : kMaskRows);
  if (rows < 1) rows = 1;
  masked_term_kernel<<<rows, 256, 0, st>>>(y, ld, idx, B, J, scale, known, out_rows);
  VIPSY_CUDA_CHECK(cudaGetLastError());
  *n_rows = rows;
  return VIPSY_OK;
}

__global__ void __launch_bounds__(256) masked_term_2bit_kernel(const uint8_t* __restrict__ y2, int64_t n_bytes, float scale,
                                                               double* __restrict__ out) {
  __shared__ long long part[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  long long n = 0;
  const int64_t nw = n_bytes >> 2;                       // y2 is 4-byte aligned (checked by the caller)
  const uint32_t* w = reinterpret_cast<const uint32_t*>(y2);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nw; i += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t v = __ldg(w + i);
    n += __popc(v & (v >> 1) & 0x55555555u);             // cells coded 3
  }
  if (blockIdx.x == 0)
    for (int64_t k = (nw << 2) + threadIdx.x; k < n_bytes; k += blockDim.x) {
      const uint32_t v = y2[k];
      n += __popc(v & (v >> 1) & 0x55u);
    }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) n += __shfl_xor_sync(0xffffffffu, n, o);
  if (lane == 0) part[warp] = n;
  __syncthreads();
  if (threadIdx.x == 0) {
    long long t = 0;
    for (int k = 0; k < 8; ++k) t += part[k];
    out[blockIdx.x] = -(double)scale * (double)kMaskedLogProb * (double)t;
  }
}

int launch_masked_term_2bit(const uint8_t* y2, int64_t n_bytes, float scale, int64_t known, double* out_rows, int* n_rows,
                            cudaStream_t st) {
  if (known >= 0) return launch_masked_term(y2, 0, nullptr, 0, 0, scale, known, out_rows, n_rows, st);
  int rows = (int)std::min<int64_t>(std::max<int64_t>((n_bytes + 16383) / 16384, 1), kMaskRows);
  masked_term_2bit_kernel<<<rows, 256, 0, st>>>(y2, n_bytes, scale, out_rows);
  VIPSY_CUDA_CHECK(cudaGetLastError());
  *n_rows = rows;
  return VIPSY_OK;
}

// ---------------------------------------------------------------------------
// float responses with NaN = missing (vi.py:621) -> one byte per cell
// ---------------------------------------------------------------------------
__global__ void pack_responses_kernel(const float* __restrict__ y, int64_t n, uint8_t* __restrict__ out) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x * 4;
  for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n; i += stride) {
    if (i + 3 < n && ((reinterpret_cast<uintptr_t>(y + i) & 15) == 0) &&
        ((reinterpret_cast<uintptr_t>(out + i) & 3) == 0)) {
      const float4 v = *reinterpret_cast<const float4*>(y + i);
      uchar4 c;
      c.x = v.x != v.x ? 255 : (v.x != 0.f);
      c.y = v.y != v.y ? 255 : (v.y != 0.f);
      c.z = v.z != v.z ? 255 : (v.z != 0.f);
      c.w = v.w != v.w ? 255 : (v.w != 0.f);
      *reinterpret_cast<uchar4*>(out + i) = c;
    } else {
      for (int64_t k = i; k < n && k < i + 4; ++k) {
        const float v = y[k];
        out[k] = v != v ? 255 : (v != 0.f);
      }
    }
  }
}

// 2-bit host transport format of the response codes: byte i of a row holds cells 4i..4i+3, two bits each
// (0, 1, 3 = missing), rows padded to whole bytes.  Four times fewer bytes over PCIe than the code matrix,
// sixteen times fewer than the reference's NaN-coded float matrix.  One thread expands one packed byte.
__global__ void unpack_2bit_kernel(const uint8_t* __restrict__ packed, int64_t n_rows, int J, uint8_t* __restrict__ codes,
                                   int64_t ld) {
  const int PB = (J + 3) >> 2;
  const int64_t total = n_rows * PB;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / PB;
    const int w = (int)(i - r * PB);
    const uint32_t v = packed[i];
    uint8_t* o = codes + r * ld + 4 * w;
    uint32_t out = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint32_t b = (v >> (2 * k)) & 3u;
      out |= (b == 3u ? 255u : b) << (8 * k);
    }
    if (4 * w + 3 < J && ((reinterpret_cast<uintptr_t>(o) & 3) == 0)) {
      *reinterpret_cast<uint32_t*>(o) = out;
    } else {
      for (int k = 0; k < 4 && 4 * w + k < J; ++k) o[k] = (uint8_t)(out >> (8 * k));
    }
  }
}

// eta table, exactly 0/1.  DINA: every required attribute mastered.  DINO as the
// reference executes it (vi.py:95-98): 1 iff the item needs at
// least two attributes and at least one of them is mastered.
__global__ void eta_table_kernel(const uint32_t* __restrict__ qmask, int J, int K, int model,
                                 uint8_t* __restrict__ eta) {
  const int64_t n = (int64_t)J << K;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t c = (uint32_t)(i / J);
    const uint32_t q = qmask[i % J];
    uint8_t e;
    if (model == VIPSY_CDM_DINA) e = (c & q) == q;
    else e = (__popc(q) >= 2) && ((c & q) != 0);
    eta[i] = e;
  }
}

template <bool NORMAL>
__global__ void fill_noise_kernel(float* __restrict__ out, int S, int64_t B, int K, uint64_t seed,
                                  uint64_t offset, int64_t person_offset, const uint64_t* __restrict__ offset_dev) {
  if (offset_dev) offset += *offset_dev;      // CUDA-graph replays: the call counter lives in device memory
  if (NORMAL && K > 32 && K <= 128) {         // the 32 < K <= 128 layout of noise_coord(): one element per thread
    const int64_t total = (int64_t)S * B * K;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
      const int k = (int)(i % K);
      const int64_t b = (i / K) % B;
      const int s = (int)(i / ((int64_t)K * B));
      uint32_t blk;
      int comp;
      noise_coord(s, k, K, blk, comp);
      const float4 z = normal4(seed, offset, (uint64_t)(person_offset + b), blk);
      out[i] = comp == 0 ? z.x : (comp == 1 ? z.y : (comp == 2 ? z.z : z.w));
    }
    return;
  }
  const int64_t per_person = (int64_t)S * K;
  const int64_t blocks_pp = (per_person + 3) / 4;
  const int64_t total = B * blocks_pp;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = i / blocks_pp;
    const uint32_t blk = (uint32_t)(i % blocks_pp);
    const uint4 r = philox_draw(seed, offset, (uint64_t)(person_offset + b), blk);
    float4 z;
    if (NORMAL) z = box_muller4(r);
    else z = make_float4(u01(r.x), u01(r.y), u01(r.z), u01(r.w));
    const float zz[4] = {z.x, z.y, z.z, z.w};
    for (int e = 0; e < 4; ++e) {
      const int64_t n = (int64_t)blk * 4 + e;
      if (n < per_person) {
        const int64_t s = n / K, k = n % K;
        out[(s * B + b) * K + k] = zz[e];
      }
    }
  }
}

// ---------------------------------------------------------------------------
// Multi-tensor Adam with the patched SVI.step's extras folded in (vi.py:508-515): the `free` mask multiplies the
// gradient (vi.py:511-512), torch.optim.Adam's default update follows (no weight decay, no amsgrad:
// pyro.optim.Adam), the gradient is zeroed (vi.py:515).  One launch for every parameter of the
// step; the step count and the learning rates may live in device memory so that a CUDA graph of the whole
// iteration can be replayed.
// ---------------------------------------------------------------------------
constexpr int kAdamMaxTensors = 24;
struct AdamBatch {
  VipsyAdamTensor t[kAdamMaxTensors];
  int n;
};

__global__ void __launch_bounds__(256) adam_kernel(const AdamBatch b, const int64_t* __restrict__ step_dev, int zero_grad) {
  const VipsyAdamTensor t = b.t[blockIdx.y];
  const int64_t step = *step_dev - t.step_offset;       // this tensor's own step count, >= 1
  if (step < 1) return;
  const float lr = t.lr_dev ? *t.lr_dev : t.lr;
  // torch.optim.Adam (single-tensor form): step_size = lr / (1 - b1^t); denom = sqrt(v) / sqrt(1 - b2^t) + eps
  const double bc1 = 1.0 - pow((double)t.beta1, (double)step);
  const double bc2 = 1.0 - pow((double)t.beta2, (double)step);
  const float step_size = (float)((double)lr / bc1);
  const float inv_sqrt_bc2 = (float)(1.0 / sqrt(bc2));
  const float b1 = t.beta1, b2 = t.beta2, omb1 = 1.f - t.beta1, omb2 = 1.f - t.beta2;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < t.n; i += (int64_t)gridDim.x * blockDim.x) {
    float g = t.grad[i];
    if (t.free_mask && !t.free_mask[i]) g = 0.f;
    const float m = fmaf(b1, t.exp_avg[i], omb1 * g);             // lerp(m, g, 1 - b1)
    const float v = fmaf(b2, t.exp_avg_sq[i], omb2 * g * g);
    t.exp_avg[i] = m;
    t.exp_avg_sq[i] = v;
    const float denom = sqrtf(v) * inv_sqrt_bc2 + t.eps;
    t.param[i] -= step_size * (m / denom);
    if (zero_grad) t.grad[i] = 0.f;
  }
}

// ---------------------------------------------------------------------------
// One-shot all-reduce over NVLink peer memory for the path's only exchange: the packed item-gradient
// buffer + the loss, 0.4-200 KB, latency-bound.  Every rank owns a symmetric buffer (same layout on all ranks,
// mapped into every process): [flags: world x uint32 | pad][slot 0 payload][slot 1 payload]; a payload is
// n_float floats followed by n_double doubles.  The producer kernels of this rank (finalize_kernel) write the
// payload of slot epoch & 1 directly; this kernel then
//   1. publishes: system-scope fence, writes `epoch` into flag[rank] of every peer,
//   2. waits until all world flags of its own buffer have reached `epoch`,
//   3. reads every peer's payload over NVLink and sums in RANK ORDER (all ranks get bit-identical results),
// one launch, no NCCL call, no host round trip.  Slots alternate, so a peer still reading slot e while this rank
// already produces e + 1 is safe; slot e is rewritten only at e + 2, after the barrier of e + 1.
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ float ld_relaxed_sys_f32(const float* p) {
  float v;
  asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ double ld_relaxed_sys_f64(const double* p) {
  double v;
  asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ float4 ld_relaxed_sys_f32x4(const float* p) {
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}

constexpr int kPeerMaxWorld = 16;      // peers whose loads are kept in flight together

__global__ void __launch_bounds__(256) peer_allreduce_kernel(const uint64_t* __restrict__ peer_bufs, int rank, int world,
                                                             int64_t flag_bytes, int64_t slot_bytes, int n_float,
                                                             int n_double, float* __restrict__ out_float,
                                                             double* __restrict__ out_double, uint32_t epoch) {
  const int tid = threadIdx.x;
  if (blockIdx.x == 0) {                 // publish once: everything this rank wrote before is visible system-wide
    __threadfence_system();
    if (tid < world) st_release_sys(reinterpret_cast<uint32_t*>(peer_bufs[tid]) + rank, epoch);
  }
  if (tid < world) {                     // every CTA waits for every peer's flag in its own buffer
    const uint32_t* mine = reinterpret_cast<const uint32_t*>(peer_bufs[rank]) + tid;
    while ((int32_t)(ld_acquire_sys(mine) - epoch) < 0) {}
  }
  __syncthreads();
  const int64_t off = flag_bytes + (int64_t)(epoch & 1u) * slot_bytes;
  const int64_t doff = off + (((int64_t)n_float * 4 + 7) & ~int64_t(7));
  // floats, four at a time; all peers' loads of a chunk are issued before the first add (one NVLink round trip)
  const int nvec = n_float >> 2;
  for (int i = blockIdx.x * blockDim.x + tid; i < nvec; i += gridDim.x * blockDim.x) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int r0 = 0; r0 < world; r0 += kPeerMaxWorld) {
      float4 v[kPeerMaxWorld];
#pragma unroll
      for (int k = 0; k < kPeerMaxWorld; ++k)
        if (r0 + k < world) v[k] = ld_relaxed_sys_f32x4(reinterpret_cast<const float*>(peer_bufs[r0 + k] + off) + 4 * i);
#pragma unroll
      for (int k = 0; k < kPeerMaxWorld; ++k)
        if (r0 + k < world) { acc.x += v[k].x; acc.y += v[k].y; acc.z += v[k].z; acc.w += v[k].w; }
    }
    *reinterpret_cast<float4*>(out_float + 4 * i) = acc;
  }
  if (blockIdx.x == 0) {
    for (int i = (nvec << 2) + tid; i < n_float; i += blockDim.x) {
      float s = 0.f;
      for (int r = 0; r < world; ++r) s += ld_relaxed_sys_f32(reinterpret_cast<const float*>(peer_bufs[r] + off) + i);
      out_float[i] = s;
    }
    for (int i = tid; i < n_double; i += blockDim.x) {
      double s = 0.0;
      for (int r = 0; r < world; ++r) s += ld_relaxed_sys_f64(reinterpret_cast<const double*>(peer_bufs[r] + doff) + i);
      out_double[i] = s;
    }
  }
}

}  // namespace vipsy

using namespace vipsy;

// ---------------------------------------------------------------------------
// Minibatch draw of the `data` plate: B distinct indices of [0, N), one CTA.  Every position owns a Philox stream
// (position, round); a position whose value already occurs at a LOWER position redraws in the next round, until
// all are distinct.  The result is invariant under relabelling of the values, hence uniform over ordered
// B-subsets (what randperm(N)[:B] gives); fixed priority => deterministic.  B <= 2048: the scan is O(B^2 / 1024).
// ---------------------------------------------------------------------------
constexpr int kSampleThreads = 1024;
__global__ void __launch_bounds__(kSampleThreads) sample_indices_kernel(int64_t* __restrict__ out, int B, int64_t N,
                                                                        uint64_t seed, uint64_t* counter) {
  __shared__ int64_t cand[VIPSY_SAMPLE_MAX_BATCH];
  const uint64_t off = *counter;
  const int t = threadIdx.x;
  bool redraw[2] = {true, true};
  for (uint32_t round = 0;; ++round) {
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int i = t + q * kSampleThreads;
      if (i < B && redraw[q]) {
        const uint4 r = philox_draw(seed ^ 0x9E3779B97F4A7C15ull, off, (uint64_t)i, round);
        const uint64_t u = ((uint64_t)r.x << 32) | r.y;
        cand[i] = (int64_t)__umul64hi(u, (uint64_t)N);               // floor(u N / 2^64)
      }
    }
    __syncthreads();
    int any = 0;
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int i = t + q * kSampleThreads;
      redraw[q] = false;
      if (i < B) {
        const int64_t v = cand[i];
        for (int j = 0; j < i; ++j)
          if (cand[j] == v) { redraw[q] = true; break; }
        any |= redraw[q];
      }
    }
    if (!__syncthreads_or(any)) break;                               // also orders the scan before the next draws
  }
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    const int i = t + q * kSampleThreads;
    if (i < B) out[i] = cand[i];
  }
  if (t == 0) *counter = off + 1;
}
static int launch_sample_indices(int64_t* out, int B, int64_t N, uint64_t seed, uint64_t* counter, cudaStream_t st) {
  if (B == 0) return VIPSY_OK;
  sample_indices_kernel<<<1, kSampleThreads, 0, st>>>(out, B, N, seed, counter);
  VIPSY_CUDA_CHECK(cudaGetLastError());
  return VIPSY_OK;
}

extern "C" {

int vipsy_abi_version(void) { return VIPSY_ABI_VERSION; }
const char* vipsy_last_error(void) { return g_err; }

int vipsy_pack_responses(const float* y_f32, int64_t n_cells, uint8_t* codes, void* stream) {
  VIPSY_REQUIRE(y_f32 && codes && n_cells >= 0, VIPSY_ERR_ARG, "vipsy_pack_responses: null argument");
  if (n_cells == 0) return VIPSY_OK;
  int64_t blocks = (n_cells / 4 + 255) / 256;
  const int64_t cap = (int64_t)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  pack_responses_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(y_f32, n_cells, codes);
  VIPSY_CUDA_CHECK(cudaGetLastError());
  return VIPSY_OK;
}

int vipsy_pack_responses_2bit_host(const uint8_t* codes_host, int64_t n_rows, int32_t n_items, int64_t row_stride,
                                   uint8_t* packed_host) {
  VIPSY_REQUIRE(codes_host && packed_host && n_rows >= 0 && n_items > 0 && row_stride >= n_items, VIPSY_ERR_ARG,
                "vipsy_pack_responses_2bit_host: bad argument");
  const int PB = (n_items + 3) >> 2;
  for (int64_t r = 0; r < n_rows; ++r) {
    const uint8_t* src = codes_host + r * row_stride;
    uint8_t* dst = packed_host + r * PB;
    for (int w = 0; w < PB; ++w) {
      uint32_t v = 0;
      for (int k = 0; k < 4; ++k) {
        const int j = 4 * w + k;
        const uint32_t c = j < n_items ? src[j] : 255u;
        VIPSY_REQUIRE(c <= 1u || c == 255u, VIPSY_ERR_ARG,
                      "vipsy_pack_responses_2bit_host: code %u at (%lld, %d) is not 0 / 1 / 255", c, (long long)r, j);
        v |= (c & 3u) << (2 * k);
      }
      dst[w] = (uint8_t)v;
    }
  }
  return VIPSY_OK;
}

int vipsy_unpack_responses_2bit(const uint8_t* packed, int64_t n_rows, int32_t n_items, uint8_t* codes,
                                int64_t codes_row_stride, void* stream) {
  VIPSY_REQUIRE(packed && codes && n_rows >= 0 && n_items > 0 && codes_row_stride >= n_items, VIPSY_ERR_ARG,
                "vipsy_unpack_responses_2bit: bad argument");
  if (n_rows == 0) return VIPSY_OK;
  const int64_t total = n_rows * ((n_items + 3) >> 2);
  int64_t blocks = (total + 255) / 256;
  if (blocks > (int64_t)num_sms() * 16) blocks = (int64_t)num_sms() * 16;
  unpack_2bit_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(packed, n_rows, n_items, codes, codes_row_stride);
  VIPSY_CUDA_CHECK(cudaGetLastError());
  return VIPSY_OK;
}

int vipsy_cdm_eta_table(const uint32_t* qmask, int32_t n_items, int32_t n_attrs, int32_t model,
                        uint8_t* eta, void* stream) {
  VIPSY_REQUIRE(qmask && eta, VIPSY_ERR_ARG, "vipsy_cdm_eta_table: null argument");
  VIPSY_REQUIRE(n_items > 0 && n_attrs > 0 && n_attrs <= 20, VIPSY_ERR_UNSUPPORTED,
                "vipsy_cdm_eta_table: need 0 < J, 0 < K <= 20 (got J=%d K=%d)", n_items, n_attrs);
  VIPSY_REQUIRE(model == VIPSY_CDM_DINA || model == VIPSY_CDM_DINO, VIPSY_ERR_ARG,
                "vipsy_cdm_eta_table: unknown model %d", model);
  const int64_t n = (int64_t)n_items << n_attrs;
  int64_t blocks = (n + 255) / 256;
  if (blocks > num_sms() * 8) blocks = num_sms() * 8;
  eta_table_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(qmask, n_items, n_attrs, model, eta);
  VIPSY_CUDA_CHECK(cudaGetLastError());
  return VIPSY_OK;
}

static int fill_noise(bool normal, float* out, int32_t S, int64_t B, int32_t K, uint64_t seed,
                      uint64_t offset, int64_t person_offset, void* stream, const uint64_t* offset_dev = nullptr) {
  VIPSY_REQUIRE(out && S > 0 && B >= 0 && K > 0, VIPSY_ERR_ARG, "vipsy_fill_*: bad argument");
  if (B == 0) return VIPSY_OK;
  const int64_t total = B * (((int64_t)S * K + 3) / 4);
  int64_t blocks = (total + 255) / 256;
  if (blocks > num_sms() * 8) blocks = num_sms() * 8;
  if (normal)
    fill_noise_kernel<true><<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(out, S, B, K, seed, offset, person_offset, offset_dev);
  else
    fill_noise_kernel<false><<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(out, S, B, K, seed, offset, person_offset, offset_dev);
  VIPSY_CUDA_CHECK(cudaGetLastError());
  return VIPSY_OK;
}

int vipsy_fill_normal(float* out, int32_t S, int64_t B, int32_t K, uint64_t seed, uint64_t offset,
                      int64_t person_offset, void* stream) {
  return fill_noise(true, out, S, B, K, seed, offset, person_offset, stream);
}
int vipsy_fill_uniform(float* out, int32_t S, int64_t B, int32_t K, uint64_t seed, uint64_t offset,
                       int64_t person_offset, void* stream) {
  return fill_noise(false, out, S, B, K, seed, offset, person_offset, stream);
}

int vipsy_fill_normal_dev(float* out, int32_t S, int64_t B, int32_t K, uint64_t seed, const uint64_t* offset_dev,
                          int64_t person_offset, void* stream) {
  VIPSY_REQUIRE(offset_dev, VIPSY_ERR_ARG, "vipsy_fill_normal_dev: offset_dev is NULL");
  return fill_noise(true, out, S, B, K, seed, 0, person_offset, stream, offset_dev);
}
int vipsy_fill_uniform_dev(float* out, int32_t S, int64_t B, int32_t K, uint64_t seed, const uint64_t* offset_dev,
                           int64_t person_offset, void* stream) {
  VIPSY_REQUIRE(offset_dev, VIPSY_ERR_ARG, "vipsy_fill_uniform_dev: offset_dev is NULL");
  return fill_noise(false, out, S, B, K, seed, 0, person_offset, stream, offset_dev);
}

int vipsy_sample_indices_dev(int64_t* idx_out, int32_t batch, int64_t n_rows, uint64_t seed, uint64_t* counter_dev,
                             void* stream) {
  VIPSY_REQUIRE(idx_out && counter_dev && batch >= 0 && n_rows > 0 && batch <= n_rows, VIPSY_ERR_ARG,
                "vipsy_sample_indices_dev: bad arguments (batch %d, rows %lld)", batch, (long long)n_rows);
  VIPSY_REQUIRE(batch <= VIPSY_SAMPLE_MAX_BATCH, VIPSY_ERR_UNSUPPORTED, "vipsy_sample_indices_dev: batch %d > %d", batch,
                (int)VIPSY_SAMPLE_MAX_BATCH);
  return launch_sample_indices(idx_out, batch, n_rows, seed, counter_dev, (cudaStream_t)stream);
}

int vipsy_peer_allreduce(const uint64_t* peer_bufs_dev, int32_t rank, int32_t world, int64_t flag_bytes,
                         int64_t slot_bytes, int32_t n_float, int32_t n_double, float* out_float, double* out_double,
                         uint32_t epoch, void* stream) {
  VIPSY_REQUIRE(peer_bufs_dev && world >= 1 && world <= 256 && rank >= 0 && rank < world && n_float >= 0 && n_double >= 0,
                VIPSY_ERR_ARG, "vipsy_peer_allreduce: bad arguments");
  VIPSY_REQUIRE(flag_bytes >= (int64_t)world * 4 && (flag_bytes & 7) == 0 && (slot_bytes & 7) == 0 &&
                    slot_bytes >= (((int64_t)n_float * 4 + 7) & ~int64_t(7)) + (int64_t)n_double * 8,
                VIPSY_ERR_ARG, "vipsy_peer_allreduce: layout does not fit (flag_bytes %lld, slot_bytes %lld)",
                (long long)flag_bytes, (long long)slot_bytes);
  VIPSY_REQUIRE((n_float == 0 || out_float) && (n_double == 0 || out_double), VIPSY_ERR_ARG,
                "vipsy_peer_allreduce: output pointer is NULL");
  VIPSY_REQUIRE((reinterpret_cast<uintptr_t>(out_float) & 15) == 0 && (flag_bytes & 15) == 0 && (slot_bytes & 15) == 0,
                VIPSY_ERR_ARG, "vipsy_peer_allreduce: out_float, flag_bytes and slot_bytes must be 16-byte aligned");
  int grid = (n_float / 4 + 255) / 256;
  if (grid < 1) grid = 1;
  if (grid > 32) grid = 32;
  peer_allreduce_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(peer_bufs_dev, rank, world, flag_bytes, slot_bytes, n_float,
                                                             n_double, out_float, out_double, epoch);
  VIPSY_CUDA_CHECK(cudaGetLastError());
  return VIPSY_OK;
}

int vipsy_adam_step(const VipsyAdamTensor* tensors, int32_t n_tensors, const int64_t* step_dev, int32_t zero_grad,
                    void* stream) {
  VIPSY_REQUIRE(tensors && n_tensors >= 0 && step_dev, VIPSY_ERR_ARG, "vipsy_adam_step: null argument");
  for (int32_t t0 = 0; t0 < n_tensors; t0 += kAdamMaxTensors) {
    AdamBatch b;
    b.n = n_tensors - t0 < kAdamMaxTensors ? n_tensors - t0 : kAdamMaxTensors;
    int64_t max_n = 0;
    for (int i = 0; i < b.n; ++i) {
      b.t[i] = tensors[t0 + i];
      VIPSY_REQUIRE(b.t[i].param && b.t[i].grad && b.t[i].exp_avg && b.t[i].exp_avg_sq && b.t[i].n >= 0, VIPSY_ERR_ARG,
                    "vipsy_adam_step: tensor %d has a null pointer", t0 + i);
      if (b.t[i].n > max_n) max_n = b.t[i].n;
    }
    int64_t bx = (max_n + 1023) / 1024;                 // 256 threads x 4 elements per pass
    if (bx > (int64_t)num_sms() * 4) bx = (int64_t)num_sms() * 4;
    if (bx < 1) bx = 1;
    adam_kernel<<<dim3((unsigned)bx, (unsigned)b.n), 256, 0, (cudaStream_t)stream>>>(b, step_dev, zero_grad);
    VIPSY_CUDA_CHECK(cudaGetLastError());
  }
  return VIPSY_OK;
}

}  // extern "C"
