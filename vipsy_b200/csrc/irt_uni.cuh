// Shared pieces of the K == 1 kernels (irt_uni.cu: dense item loop; irt_uni_sparse.cu: compacted
// item loop for mostly-missing response matrices).
#pragma once
#include "common.cuh"

namespace vipsy {

constexpr int kUniWarps = 4;
constexpr int kUniThreads = kUniWarps * 32;

struct IrtUniParams {
  const uint8_t* y;
  int64_t ld_y;
  const int64_t* idx;
  const float* mu;
  const float* sigma;
  const float* eps;
  const float *a, *b, *c, *d;
  int64_t B;
  int J, S;
  float D, scale, baseline;
  uint64_t seed, offset;
  int64_t person_offset;
  float* d_mu;
  float* d_sigma;
  float* item_partial;    // [grid][NQ*J]
  double* loss_partial;   // [grid]
  int contiguous;         // rows of the batch are contiguous in y (no idx, ld == J)
  int packed2;            // y holds 2-bit rows (J / 4 bytes each, contiguous): expanded while a tile is staged
  int want_grad;
  const int64_t* row_ptr;   // CSR of the observed cells (irt_uni_csr.cu), or null
  const uint16_t* entries;
};

__host__ __device__ inline size_t align16(size_t x) { return (x + 15) & ~size_t(15); }

__host__ __device__ inline size_t uni_smem_bytes(int J, int NQ) {
  size_t s = (size_t)J * 16;                        // float4 item table
  s += align16((size_t)J * 8);                      // {D*a*(d-c), 1-d}
  s += align16((size_t)kUniWarps * NQ * J * 4);     // warp-private item accumulators
  s += (size_t)kUniWarps * align16((size_t)32 * J + 16);  // response tiles
  return s;
}

// ---- 1- and 2-wide fp32 helpers: sm_100's packed FFMA2/FADD2/FMUL2 halve the
// issue slots of the non-SFU math, which is what this kernel runs out of first.
__device__ __forceinline__ float vfma(float a, float b, float c) { return fmaf(a, b, c); }
__device__ __forceinline__ float2 vfma(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ float vmul(float a, float b) { return a * b; }
__device__ __forceinline__ float2 vmul(float2 a, float2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ float vadd(float a, float b) { return a + b; }
__device__ __forceinline__ float2 vadd(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float vex2(float a) { return ex2_approx(a); }
__device__ __forceinline__ float2 vex2(float2 a) { return make_float2(ex2_approx(a.x), ex2_approx(a.y)); }
__device__ __forceinline__ float vrcp(float a) { return rcp_approx(a); }
__device__ __forceinline__ float2 vrcp(float2 a) { return make_float2(rcp_approx(a.x), rcp_approx(a.y)); }
__device__ __forceinline__ float vmin(float a, float m) { return fminf(a, m); }
__device__ __forceinline__ float2 vmin(float2 a, float m) { return make_float2(fminf(a.x, m), fminf(a.y, m)); }
__device__ __forceinline__ float vclampp(float a) { return fminf(fmaxf(a, kEps32), kOneMinusEps32); }
__device__ __forceinline__ float2 vclampp(float2 a) { return make_float2(vclampp(a.x), vclampp(a.y)); }
// keep v where the probability clamp did not fire (q == qc), else 0
__device__ __forceinline__ float vkeep(float v, float q, float qc) { return q == qc ? v : 0.f; }
__device__ __forceinline__ float2 vkeep(float2 v, float2 q, float2 qc) {
  return make_float2(q.x == qc.x ? v.x : 0.f, q.y == qc.y ? v.y : 0.f);
}
__device__ __forceinline__ float vhsum(float a) { return a; }
__device__ __forceinline__ float vhsum(float2 a) { return a.x + a.y; }
__device__ __forceinline__ float vhprod(float a) { return a; }
__device__ __forceinline__ float vhprod(float2 a) { return a.x * a.y; }
template <typename V> __device__ __forceinline__ V vbcast(float a);
template <> __device__ __forceinline__ float vbcast<float>(float a) { return a; }
template <> __device__ __forceinline__ float2 vbcast<float2>(float a) { return make_float2(a, a); }
__device__ __forceinline__ void vlg2_add(float q, float* lls) { lls[0] += lg2_approx(q); }
__device__ __forceinline__ void vlg2_add(float2 q, float* lls) { lls[0] += lg2_approx(q.x); lls[1] += lg2_approx(q.y); }

template <int SC> struct NoiseVec { using type = float2; static constexpr int W = 2; };
template <> struct NoiseVec<1> { using type = float; static constexpr int W = 1; };

// All SC particles of one (person, item) cell, two particles per instruction.
// it = {A', B', c, d-c}, omd = 1-d, A' = -D a log2(e), B' = -D b log2(e).
// The caller passes T0, T1 already signed by the response: tt = T1*eps + T0 is
// -l*log2(e) for y=1 and +l*log2(e) for y=0, so with e' = 2^tt
//     sp = 1/(1+e')  is  s   for y=1  and  1-s  for y=0       (no cancellation)
//     q  = base + (d-c) sp,   base = c (y=1) | 1-d (y=0)       (= P | 1-P)
// i.e. the same instruction stream serves both responses.  Then torch's
// Bernoulli(probs) rule: clamp q to [eps, 1-eps], log, zero gradient
// where the clamp fired.  Sums over the SC particles (all non-negative; the caller
// applies the signs):  W1 = sum w, W2 = sum w*eps with w = s(1-s)/q;
// A1 = sum sp/q; A2 = sum (1-sp)/q; Lj = sum log2 q (SCORE: per particle into lls).
//
// FASTPATH (4PL, every item 1e-3 <= c, d <= 1-1e-3, d-c >= 1e-3: the clamp is
// unreachable): q = (base e' + base + (d-c)) / (1+e') =: num/ope, and the single
// reciprocal R = 1/(num ope) yields sp = num R, sp/q = ope R, w = e' R:
// 2.125 MUFU per evaluation instead of 3.25.
//
// CLAMP = false (FASTPATH only): the caller has bounded tt below 60 for every particle, so the two FMNMX per
// pair are dropped.  Sums start from the first pair's terms instead of 0 / 1 (the compiler does not fold
// packed operations on constants).
// item_evals_v leaves the sums as pairs and the last group's probability product unlogged; item_evals is the
// complete form.
// eps: anything indexable by pair (the register array, or EpsShared below)
template <int SC>
struct EpsShared {                 // the lane's noise pairs in shared memory, [pair][thread]: conflict-free LDS.64
  const typename NoiseVec<SC>::type* base;
  int stride;
  __device__ __forceinline__ typename NoiseVec<SC>::type operator[](int i) const { return base[i * stride]; }
};
template <int MODEL, bool SCORE, int SC, bool FASTPATH, bool CLAMP = true, typename EpsT>
__device__ __forceinline__ void item_evals_v(const float dmc, const float base, const float T0, const float T1,
                                             const EpsT& eps,
                                             float (&lls)[SCORE ? SC : 1], typename NoiseVec<SC>::type& W1o,
                                             typename NoiseVec<SC>::type& W2o, typename NoiseVec<SC>::type& A1o,
                                             typename NoiseVec<SC>::type& A2o, typename NoiseVec<SC>::type& prodo,
                                             float& Lj) {
  using V = typename NoiseVec<SC>::type;
  constexpr int W = NoiseVec<SC>::W;
  constexpr int NV = SC / W;
  constexpr int GROUP = FASTPATH ? 8 / W : 4 / W;     // factors multiplied before one lg2 (>= 1e-24 / 2e-28)
  const V T0v = vbcast<V>(T0), T1v = vbcast<V>(T1), one = vbcast<V>(1.f);
  V W1 = vbcast<V>(0.f), W2 = W1, A1 = W1, A2 = W1, prod = one;
  if constexpr (FASTPATH) {
    const V basev = vbcast<V>(base), bdv = vbcast<V>(base + dmc);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const V ep = eps[i];
      V tt = vfma(T1v, ep, T0v);
      if constexpr (CLAMP) tt = vmin(tt, 60.f);
      const V e = vex2(tt);
      const V ope = vadd(e, one);
      const V num = vfma(basev, e, bdv);
      const V R = vrcp(vmul(num, ope));
      const V sp = vmul(num, R);
      const V q = vmul(num, sp);
      const V us = vmul(ope, R);                                 // sp / q = 1 / num
      const V w = vmul(e, R);                                    // sp (1-sp) / q
      if (i == 0) {
        A1 = us; A2 = vmul(us, e); W1 = w; W2 = vmul(w, ep);
      } else {
        A1 = vadd(A1, us);
        A2 = vfma(us, e, A2);                                    // (1-sp) / q = e' / num
        W1 = vadd(W1, w);
        W2 = vfma(w, ep, W2);
      }
      if constexpr (SCORE) {
        vlg2_add(q, &lls[i * W]);
      } else {
        prod = (i % GROUP) == 0 ? q : vmul(prod, q);
        if ((i % GROUP) == GROUP - 1 && i != NV - 1) Lj += lg2_approx(vhprod(prod));
      }
    }
  } else {
    const V dmcv = vbcast<V>(dmc), basev = vbcast<V>(base);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const V ep = eps[i];
      const V tt = vmin(vfma(T1v, ep, T0v), 126.f);              // e' finite: sp, e' sp never NaN
      const V e = vex2(tt);
      const V sp = vrcp(vadd(e, one));
      const V es = vmul(e, sp);                                  // 1 - sp
      V q;
      if constexpr (MODEL >= 3) q = vfma(dmcv, sp, basev); else q = sp;
      const V qc = vclampp(q);
      V w;
      if constexpr (MODEL >= 3) {
        const V r = vkeep(vrcp(qc), q, qc);
        const V us = vmul(r, sp);
        A1 = vadd(A1, us);
        A2 = vfma(r, es, A2);
        w = vmul(us, es);
      } else {
        w = vkeep(es, q, qc);                                    // |y - s|
      }
      W1 = vadd(W1, w);
      W2 = vfma(w, ep, W2);
      if constexpr (SCORE) {
        vlg2_add(qc, &lls[i * W]);
      } else {
        prod = (i % GROUP) == 0 ? qc : vmul(prod, qc);
        if ((i % GROUP) == GROUP - 1 && i != NV - 1) Lj += lg2_approx(vhprod(prod));
      }
    }
  }
  W1o = W1; W2o = W2; A1o = A1; A2o = A2; prodo = prod;
}

template <int MODEL, bool SCORE, int SC, bool FASTPATH, bool CLAMP = true, typename EpsT>
__device__ __forceinline__ void item_evals(const float dmc, const float base, const float T0, const float T1,
                                           const EpsT& eps,
                                           float (&lls)[SCORE ? SC : 1], float& W1o, float& W2o, float& A1o,
                                           float& A2o, float& Lj) {
  typename NoiseVec<SC>::type W1, W2, A1, A2, prod;
  item_evals_v<MODEL, SCORE, SC, FASTPATH, CLAMP>(dmc, base, T0, T1, eps, lls, W1, W2, A1, A2, prod, Lj);
  if constexpr (!SCORE) Lj += lg2_approx(vhprod(prod));
  W1o = vhsum(W1); W2o = vhsum(W2); A1o = vhsum(A1); A2o = vhsum(A2);
}


// One (person, item) cell, one particle.  t = signed logit in log2 units (see item_evals).
// Returns w = s(1-s)/q, a1 = sp/q, a2 = (1-sp)/q (all 0 when !obs) and the clamped probability (1 when !obs).
template <int MODEL, bool FASTPATH>
__device__ __forceinline__ void cell_eval1(const float t, const float dmc, const float base, const bool obs,
                                           float& w, float& a1, float& a2, float& qc) {
  if constexpr (FASTPATH) {
    const float e = ex2_approx(fminf(t, 60.f));
    const float ope = e + 1.f;
    const float num = fmaf(base, e, base + dmc);
    float R = rcp_approx(num * ope);
    R = obs ? R : 0.f;
    const float sp = num * R;
    const float q = num * sp;
    a1 = ope * R;
    a2 = a1 * e;
    w = e * R;
    qc = obs ? q : 1.f;
  } else {
    const float e = ex2_approx(fminf(t, 126.f));
    const float sp = rcp_approx(e + 1.f);
    const float es = e * sp;
    float q;
    if constexpr (MODEL >= 3) q = fmaf(dmc, sp, base); else q = sp;
    const float qcl = vclampp(q);
    const bool keep = obs && (q == qcl);
    if constexpr (MODEL >= 3) {
      const float r = keep ? rcp_approx(qcl) : 0.f;
      a1 = r * sp;
      a2 = r * es;
      w = a1 * es;
    } else {
      a1 = 0.f; a2 = 0.f;
      w = keep ? es : 0.f;
    }
    qc = obs ? qcl : 1.f;
  }
}

// sparse (compacted) variant, irt_uni_sparse.cu
size_t irt_uni_sparse_smem_bytes(int J, int NQ);
int irt_uni_sparse_launch(int model, bool fast, int SC, IrtUniParams& p, cudaStream_t st, int* grid_out);


// S == 1 row kernels (warp per person, lanes over items), irt_uni_row.cu
bool irt_uni_row_supported(const VipsyIrtArgs* a, bool sparse);
int irt_uni_row_launch(int model, bool fast, bool sparse, IrtUniParams& p, cudaStream_t st, int* grid_out);


// S == 1 over a CSR of the observed cells, irt_uni_csr.cu
bool irt_uni_csr_supported(const VipsyIrtArgs* a);
int irt_uni_csr_launch(const VipsyIrtArgs* a, bool fast, IrtUniParams& p, cudaStream_t st, int* grid_out);

}  // namespace vipsy
