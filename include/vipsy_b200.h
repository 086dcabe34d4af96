/* vipsy_b200.h -- C ABI of the B200-native ELBO + gradient path.
 *
 * The reference (inuyasha2012/vipsy) has no native ABI: its hot path is reached
 * through Pyro's ELBO-object protocol
 *     loss_and_grads(model, guide, data)      vi.py:507  (SVI.step, vi.py:503-516)
 *     loss(model, guide, data)                vi.py:770, vi.py:942 (evaluate_loss)
 * selected by the `loss=` kwarg of fit() (vi.py:627, 758, 862, 936).  These
 * entry points are what a binding of that protocol calls:
 * one call = one ELBO + gradient evaluation over a batch of persons.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer unless the name ends in `_host`;
 *  - parameters are the CONSTRAINED values pyro.param() returns (a, b, c, d,
 *    sigma, g, s, lam1 ...); gradients are w.r.t. those values (the chain rule
 *    through sigmoid / exp stays with the caller) -- except `chol` with
 *    chol_raw=1, which is consumed raw (LowerCholeskyTransform fused, A.4);
 *  - `stream` is a cudaStream_t passed as void*; calls are stream-ordered, never
 *    synchronise, never allocate; the caller owns every buffer incl. workspace;
 *  - return value: 0 or a negative VIPSY_ERR_*; vipsy_last_error() gives text;
 *  - responses are one byte per cell: 0, 1, or VIPSY_Y_MISSING (NaN in the
 *    reference's float matrix, vi.py:621-624).
 */
#ifndef VIPSY_B200_H
#define VIPSY_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VIPSY_ABI_VERSION 3

enum {
  VIPSY_OK = 0,
  VIPSY_ERR_ARG = -1,         /* null / inconsistent argument            */
  VIPSY_ERR_UNSUPPORTED = -2, /* size or mode this build does not cover   */
  VIPSY_ERR_WORKSPACE = -3,   /* workspace too small                      */
  VIPSY_ERR_CUDA = -4         /* a CUDA runtime call failed               */
};

enum { VIPSY_Y_MISSING = 255 };

/* BaseIRT.IRT_FUN keys, vi.py:538-543 */
enum { VIPSY_IRT_1PL = 1, VIPSY_IRT_2PL = 2, VIPSY_IRT_3PL = 3, VIPSY_IRT_4PL = 4 };
/* BaseCDM.CDM_FUN keys, vi.py:728-731 */
enum { VIPSY_CDM_DINA = 0, VIPSY_CDM_DINO = 1 };
/* gradient estimator for the Gaussian guide (SURVEY Finding 1 / Appendix A.2) */
enum { VIPSY_EST_PATHWISE = 0, VIPSY_EST_SCORE = 1 };
/* guide covariance, VIRT.guide vi.py:701-723 */
enum { VIPSY_COV_DIAG1 = 0,   /* K == 1: sigma[B]                                     */
       VIPSY_COV_SHARED = 1,  /* share_cov=True: one K x K factor                     */
       VIPSY_COV_PERSON = 2 };/* one factor per person                                */
/* flags: STRICT keeps the reference's formula class (sigmoid -> c+(d-c)s -> clamp)
 * everywhere; without it the 4PL kernel switches, per launch and on the device,
 * to an algebraically equal rational form when every item has 1e-3 <= c, d <= 1-1e-3
 * (where the probability clamp cannot fire). */
enum { VIPSY_IRT_FLAG_STRICT = 1,
       VIPSY_IRT_FLAG_NO_TENSOR = 2,    /* K > 1: force the SIMT contraction (default: tcgen05 when K >= 8) */
       VIPSY_IRT_FLAG_SPARSE = 4,       /* K == 1 hint: most cells are missing -> per-person compaction of the
                                          observed cells (same results; pays off below ~50 % density)        */
       VIPSY_IRT_FLAG_NO_ROW = 8,       /* K == 1, S == 1: force the lane-per-person kernels (default: the
                                          warp-per-person row kernels, J <= 512 (256 for 3PL / 4PL))          */
       VIPSY_IRT_FLAG_MASKED_KNOWN = 16,/* n_masked_cells is valid: saves one counting pass over the batch rows
                                          in the K == 1, S > 1 path (the matrix is constant between steps)   */
       VIPSY_IRT_FLAG_NO_FUSED = 32,    /* 2 <= K <= 8, per-person factor: force the three-kernel path (sample,
                                          contraction, guide gradient) instead of the single fused kernel    */
       VIPSY_IRT_FLAG_Y_PACKED2 = 64 }; /* K == 1, S > 1, pathwise, J % 4 == 0, no idx: `y` holds the rows in the 2-bit
                                          transport format (vipsy_pack_responses_2bit_host: 4 cells per byte, 0 / 1 /
                                          3 = missing), y_row_stride = J / 4 bytes, y 4-byte aligned; the particle
                                          kernel expands a tile while staging it -- no byte matrix in HBM (what
                                          vipsy_irt_host_step uses for its chunks)                              */
enum { VIPSY_CHOL_DENSE = 0,  /* row-major K x K, strict upper triangle ignored        */
       VIPSY_CHOL_PACKED = 1 };/* torch.tril_indices row-major order, K(K+1)/2 (vi.py:453) */

/* ------------------------------------------------------------------------
 * IRT: Trace_ELBO over BaseIRT.model (vi.py:574-625) with the VIRT / VaeIRT
 * Gaussian guide (vi.py:673-723), S particles, on a batch of B persons.
 *
 *   loss = -(scale/S) * sum_s sum_b [ sum_j m_bj lb(y_bj, P_bjs) + log p(x_bs) - log q(x_bs) ]
 *
 * x_bs = mu_b + sigma_b eps_bs (K==1) or mu_b + L_b eps_bs.  All S particles of
 * one call share the batch (idx); a caller that re-draws the minibatch per
 * particle (Pyro does when subsample_size < N) issues one call per particle
 * or stacks the particles along the batch axis with S=1.
 * ---------------------------------------------------------------------- */
typedef struct VipsyIrtArgs {
  /* sizes */
  int64_t n_rows;          /* rows of y                                         */
  int64_t batch;           /* B persons evaluated                               */
  int32_t n_items;         /* J                                                 */
  int32_t n_dims;          /* K                                                 */
  int32_t n_particles;     /* S                                                 */
  int32_t model;           /* VIPSY_IRT_*                                       */
  int32_t estimator;       /* VIPSY_EST_*                                       */
  int32_t cov_mode;        /* VIPSY_COV_*                                       */
  int32_t chol_layout;     /* VIPSY_CHOL_* (K > 1)                              */
  int32_t chol_raw;        /* 1: diag of chol is log-scale (exp fused), grads raw */
  float D;                 /* scaling constant, vi.py:549                        */
  float scale;             /* N / B of the `data` plate                         */
  float baseline;          /* score estimator control variate (0 = reference)   */
  int32_t flags;           /* VIPSY_IRT_FLAG_*                                  */
  /* inputs */
  const uint8_t* y;        /* n_rows x J response codes                          */
  int64_t y_row_stride;    /* bytes between rows of y (>= J)                     */
  const int64_t* idx;      /* B row indices into y, or NULL for rows 0..B-1      */
  const float* mu;         /* B x K variational means (already gathered)         */
  const float* sigma;      /* K==1: B std-devs                                   */
  const float* chol;       /* K>1: K x K | B x K x K | B x K(K+1)/2              */
  const float* eps;        /* S x B x K injected N(0,1) noise, or NULL = Philox  */
  uint64_t seed;           /* Philox key   (eps == NULL)                         */
  uint64_t offset;         /* Philox counter offset, bump once per call          */
  int64_t person_offset;   /* global index of batch row 0 (multi-GPU shards)     */
  const float* a;          /* K x J slopes (ignored for 1PL)                     */
  const float* b;          /* J intercepts                                       */
  const float* c;          /* J lower asymptotes (3PL, 4PL)                      */
  const float* d;          /* J upper asymptotes (4PL)                           */
  const uint8_t* a_free;   /* K x J 0/1 mask applied to d_a (vi.py:511-512) or NULL */
  /* outputs */
  double* loss;            /* 1                                                 */
  float* d_mu;             /* B x K                                             */
  float* d_sigma;          /* K==1: B                                           */
  float* d_chol;           /* same layout as chol                               */
  float* d_a;              /* K x J                                             */
  float* d_b;              /* J                                                 */
  float* d_c;              /* J                                                 */
  float* d_d;              /* J                                                 */
  /* scratch */
  void* workspace;
  size_t workspace_bytes;
  /* optional (ABI 2): the observed cells of y as a CSR, built once per response matrix with
   * vipsy_observed_row_ptr + vipsy_observed_fill.  Row r of y owns entries[row_ptr[r] .. row_ptr[r+1]);
   * an entry is item | response << 15.  Used when K == 1 with the pathwise estimator (S == 1: a sub-warp per
   * person; S > 1: a lane per person walking its entries): work and traffic scale with the observed cells, not
   * with n_rows x J; with idx the CSR is that of the whole matrix.  y must still be valid (the score estimator
   * reads it). */
  const int64_t* obs_row_ptr;   /* n_rows + 1, or NULL                               */
  const uint16_t* obs_entries;  /* obs_nnz entries, or NULL                          */
  int64_t obs_nnz;              /* host copy of row_ptr[n_rows] (picks the lane mapping; 0 = unknown) */
  int64_t n_masked_cells;       /* missing cells among the B batch rows (with VIPSY_IRT_FLAG_MASKED_KNOWN) */
} VipsyIrtArgs;

/* Bytes of workspace vipsy_irt_elbo_grad needs for these sizes (0 on bad args). */
size_t vipsy_irt_workspace_bytes(const VipsyIrtArgs* args);
/* loss + every gradient (SVI.step's loss_and_grads, vi.py:507).  Pass NULL
 * gradient pointers + estimator PATHWISE for the value only (evaluate_loss). */
int vipsy_irt_elbo_grad(const VipsyIrtArgs* args, void* stream);

/* ------------------------------------------------------------------------
 * The same evaluation (K == 1) with HOST buffers: what a binding of loss_and_grads(model, guide, data)
 * (vi.py:507) calls when `data` and the variational parameters live in host memory.  The persons are cut into
 * n_chunks chunks; copies in, kernels and copies out overlap on the three caller-owned streams; everything is
 * enqueued before the call returns and the host outputs are valid once `s_out` has been synchronised.  Pinned
 * host memory is needed for the copies to be asynchronous.  `dev` supplies sizes, model, scalars, flags,
 * seed / offset / person_offset; its pointer fields are ignored.
 * ---------------------------------------------------------------------- */
typedef struct VipsyIrtHostArgs {
  VipsyIrtArgs dev;
  const uint8_t* y_host;     /* batch x J codes, or batch x (J+3)/4 rows of the 2-bit transport format */
  int32_t y_packed2;         /* 1: y_host is in the 2-bit transport format                             */
  int32_t n_chunks;          /* >= 1                                                                   */
  const float* mu_host;      /* batch                                                                  */
  const float* sigma_host;   /* batch                                                                  */
  const float* a_host;       /* J (2PL+)                                                               */
  const float* b_host;       /* J                                                                      */
  const float* c_host;       /* J (3PL+)                                                               */
  const float* d_host;       /* J (4PL)                                                                */
  double* loss_host;         /* 1                                                                      */
  float* d_mu_host;          /* batch, or NULL                                                         */
  float* d_sigma_host;       /* batch; NULL exactly when d_mu_host is NULL                             */
  float* item_grads_host;    /* [d_b | d_a | d_c | d_d], J x model floats, or NULL                     */
  void* dev_scratch;         /* device scratch, vipsy_irt_host_scratch_bytes()                         */
  size_t dev_scratch_bytes;
  void* s_in;                /* cudaStream_t: host -> device copies                                    */
  void* s_run;               /* cudaStream_t: kernels                                                  */
  void* s_out;               /* cudaStream_t: device -> host copies                                    */
  const uint8_t* a_free_host; /* J 0/1 bytes: the `free` mask of site 'a' (vi.py:511-512) applied to d_a, or NULL */
} VipsyIrtHostArgs;
size_t vipsy_irt_host_scratch_bytes(const VipsyIrtHostArgs* args);
int vipsy_irt_host_step(const VipsyIrtHostArgs* args);

/* ------------------------------------------------------------------------
 * Amortised multivariate-normal guide: VaeIRT.guide for K > 1 (vi.py:685-693) with the tail of MvnEncoder
 * (vi.py:438-455) fused in.  The reference computes, per step and for the B batch rows,
 *     hidden = softplus(fc1(x));  loc = fc21(hidden);  raw = fc22(hidden)            (vi.py:446-451)
 *     L = tril(scatter(raw), -1) + diag(exp(diag))   (vi.py:452-454 + LowerCholeskyTransform, vi.py:693)
 *     x = loc + L eps  ~  MultivariateNormal(loc, scale_tril = L)                    (vi.py:693, rsample)
 * and backpropagates through all of it.  raw is B x K(K+1)/2 floats (4 GB for 200 000 persons at K = 100) that
 * are written once and read once, and so is its gradient.  This entry takes `hidden` and `loc` (B x H and
 * B x K -- the caller keeps fc1 / fc21 in its framework, as north_star asks for the encoder MLP; or hands fc21 over
 * as well, see w21 below) and evaluates
 * fc22 on the tensor cores inside the sampler: factor rows are consumed straight from tensor memory, then the
 * likelihood contraction of vipsy_irt_elbo_grad runs, and the backward pass forms d fc22.weight, d fc22.bias
 * and d hidden from the outer products (d loss / d x)(eps)^T on chip.  Neither raw nor its gradient ever
 * exists in HBM.  Packed order of raw: torch.tril_indices(K, K), row-major (vi.py:453).
 *
 * `irt` supplies sizes, data, item parameters, mu (= loc), noise, the loss and d_mu / item-gradient outputs exactly
 * as for vipsy_irt_elbo_grad; its sigma / chol / d_sigma / d_chol / workspace fields are ignored
 * (cov_mode PERSON, pathwise estimator, K >= 2).  H <= 64.
 * ---------------------------------------------------------------------- */
typedef struct VipsyMvnEncoderIrtArgs {
  VipsyIrtArgs irt;
  int32_t n_hidden;        /* H                                                                         */
  int32_t reserved;
  const float* hidden;     /* B x H   softplus(fc1(x)) of the batch rows (in batch order)                */
  const float* w22;        /* K(K+1)/2 x H   fc22.weight                                                 */
  const float* b22;        /* K(K+1)/2       fc22.bias                                                   */
  float* d_hidden;         /* B x H   d loss / d hidden through fc22 (fc21's share follows from d_mu)    */
  float* d_w22;            /* K(K+1)/2 x H                                                               */
  float* d_b22;            /* K(K+1)/2                                                                   */
  void* workspace;
  size_t workspace_bytes;  /* >= vipsy_irt_mvn_encoder_workspace_bytes()                                 */
  /* Optional: the location head fc21 as well (vi.py:443, 450: loc = fc21(hidden)).  With w21 != NULL the call
   * computes loc itself (irt.mu is ignored and may be NULL) -- each row of fc21 rides in a padding slot of the
   * diagonal 8 x 8 block of its factor row, as one more column against a constant draw of 1 -- and returns
   * d fc21.weight / d fc21.bias; d_hidden then holds the gradient through BOTH heads.  irt.d_mu (= d loc) is
   * still written when non-NULL. */
  const float* w21;        /* K x H   fc21.weight, or NULL                                               */
  const float* b21;        /* K       fc21.bias (required with w21)                                      */
  float* d_w21;            /* K x H, or NULL                                                             */
  float* d_b21;            /* K, or NULL                                                                 */
} VipsyMvnEncoderIrtArgs;
size_t vipsy_irt_mvn_encoder_workspace_bytes(const VipsyMvnEncoderIrtArgs* args);
/* loss + every gradient (NULL gradient pointers: the value only, evaluate_loss) */
int vipsy_irt_mvn_encoder_elbo_grad(const VipsyMvnEncoderIrtArgs* args, void* stream);
/* The guide's sample site alone (vi.py:693): x [S][B][K] = loc + L eps, the draws eps [S][B][K], diag L [B][K]
 * and logdet [B] = sum_i log L_ii.  x, diag and logdet may be NULL; eps_out only with injected noise (the library's own
 * draws need a home).  Uses args->workspace like the full entry. */
int vipsy_mvn_encoder_sample(const VipsyMvnEncoderIrtArgs* args, float* x, float* eps_out, float* diag, float* logdet,
                             void* stream);

/* ------------------------------------------------------------------------
 * Enumerated CDMs: TraceEnum_ELBO over VCCDM.model (vi.py:839-857), optionally
 * with the HO-DINA pattern prior and Gaussian guide of VCHoDina (vi.py:897-934).
 *
 *   loss = -(scale/S) sum_s sum_b [ logsumexp_c( log prior_bc + sum_j m_bj lb(y_bj, p_cj) )
 *                                   (+ log N(theta_bs;0,1) - log q(theta_bs)) ]
 * Pattern c has attribute k mastered iff bit k of c is set (vi.py:825-837).
 * ---------------------------------------------------------------------- */
typedef struct VipsyCdmEnumArgs {
  int64_t n_rows;
  int64_t batch;
  int32_t n_items;         /* J                                                 */
  int32_t n_attrs;         /* K  (C = 2^K patterns)                             */
  int32_t n_particles;     /* S (HO-DINA noise draws; 1 otherwise)              */
  int32_t model;           /* VIPSY_CDM_*  (HO-DINA is always DINA, vi.py:905)  */
  int32_t hodina;          /* 1: VCHoDina prior + theta guide                   */
  int32_t reserved0;
  float scale;
  float reserved1;
  const uint8_t* y;
  int64_t y_row_stride;
  const int64_t* idx;
  const uint32_t* qmask;   /* J bit masks: bit k = Q[k][j] (vi.py:741)          */
  const float* g;          /* J guess                                            */
  const float* s;          /* J slip                                             */
  const float* lam0;       /* K  (hodina)                                        */
  const float* lam1;       /* K  (hodina, positive)                              */
  const float* mu;         /* B  (hodina) theta_local                            */
  const float* sigma;      /* B  (hodina) theta_scale                            */
  const float* eps;        /* S x B injected noise or NULL = Philox              */
  uint64_t seed, offset;
  int64_t person_offset;
  double* loss;
  float* d_g;              /* J                                                  */
  float* d_s;              /* J                                                  */
  float* d_lam0;           /* K                                                  */
  float* d_lam1;           /* K                                                  */
  float* d_mu;             /* B                                                  */
  float* d_sigma;          /* B                                                  */
  float* posterior;        /* optional B x C pattern posterior (read-out), or NULL */
  void* workspace;
  size_t workspace_bytes;
  /* VaeCCDM (vi.py:866-891), ignored with hodina: */
  const float* log_prior;  /* B x C log Categorical(attr_p).logits of the batch persons, or NULL = uniform prior;
                              d loss / d log_prior = -scale * posterior                                        */
  int32_t score_missing;   /* 1: a missing cell is scored as Bernoulli(p).log_prob(-1) = 2 log(1-p) - log p
                              (the reference substitutes -1 and does not mask, vi.py:881-891); 0: masked     */
  int32_t reserved2;
} VipsyCdmEnumArgs;

size_t vipsy_cdm_enum_workspace_bytes(const VipsyCdmEnumArgs* args);
int vipsy_cdm_enum_elbo_grad(const VipsyCdmEnumArgs* args, void* stream);

/* ------------------------------------------------------------------------
 * REINFORCE CDM: Trace_ELBO over BaseCDM.model (vi.py:745-756) with the
 * Bernoulli guide of VCDM / VaeCDM (vi.py:796-816): the score-function path
 * the reference actually executes.  attr_bk = [u_bk < attr_p_bk].
 * ---------------------------------------------------------------------- */
typedef struct VipsyCdmScoreArgs {
  int64_t n_rows;
  int64_t batch;
  int32_t n_items, n_attrs, n_particles, model;
  float scale, baseline;
  const uint8_t* y;
  int64_t y_row_stride;
  const int64_t* idx;
  const uint32_t* qmask;
  const float* g;
  const float* s;
  const float* attr_p;     /* B x K guide probabilities (gathered)               */
  const float* u;          /* S x B x K injected U(0,1) or NULL = Philox          */
  uint64_t seed, offset;
  int64_t person_offset;
  double* loss;
  float* d_g;
  float* d_s;
  float* d_attr_p;         /* B x K                                              */
  void* workspace;
  size_t workspace_bytes;
} VipsyCdmScoreArgs;

size_t vipsy_cdm_score_workspace_bytes(const VipsyCdmScoreArgs* args);
int vipsy_cdm_score_elbo_grad(const VipsyCdmScoreArgs* args, void* stream);

/* ------------------------------------------------------------------------
 * Data-format helpers either side of the path
 * ---------------------------------------------------------------------- */
/* float response matrix with NaN = missing (the reference's `data`) -> codes */
int vipsy_pack_responses(const float* y_f32, int64_t n_cells, uint8_t* codes, void* stream);
/* 2-bit host transport format: byte i of a row = cells 4i..4i+3, two bits each (0, 1, 3 = missing), rows
 * padded to (J+3)/4 bytes.  The packer runs on the HOST (both pointers are host memory; done once when the
 * response matrix is loaded); the expander runs on the device, stream-ordered, after the packed rows were
 * copied in -- 4x fewer PCIe bytes than the code matrix, 16x fewer than the NaN-coded float matrix. */
int vipsy_pack_responses_2bit_host(const uint8_t* codes_host, int64_t n_rows, int32_t n_items, int64_t row_stride,
                                   uint8_t* packed_host);
int vipsy_unpack_responses_2bit(const uint8_t* packed, int64_t n_rows, int32_t n_items, uint8_t* codes,
                                int64_t codes_row_stride, void* stream);
/* CSR of the observed cells (the "2-byte per response" layout for mostly-missing matrices, J <= 32768).
 * Step 1 writes row_ptr[0..n_rows] (counts_scratch: n_rows int32 of scratch); the caller reads
 * row_ptr[n_rows] to size `entries`; step 2 writes entry = item | (response != 0) << 15 in item order. */
int vipsy_observed_row_ptr(const uint8_t* y, int64_t n_rows, int32_t n_items, int64_t y_row_stride,
                           int32_t* counts_scratch, int64_t* row_ptr, void* stream);
int vipsy_observed_fill(const uint8_t* y, int64_t n_rows, int32_t n_items, int64_t y_row_stride,
                        const int64_t* row_ptr, uint16_t* entries, void* stream);
/* eta table, bit-exact: out[c*J + j] = eta_cj in {0,1} (dina/dino, vi.py:78-81/95-98) */
int vipsy_cdm_eta_table(const uint32_t* qmask, int32_t n_items, int32_t n_attrs, int32_t model,
                        uint8_t* eta, void* stream);
/* the in-kernel noise generator, exposed: out[s*B*K + b*K + k] (Philox4x32-10 + Box-Muller) */
int vipsy_fill_normal(float* out, int32_t n_particles, int64_t batch, int32_t n_dims,
                      uint64_t seed, uint64_t offset, int64_t person_offset, void* stream);
int vipsy_fill_uniform(float* out, int32_t n_particles, int64_t batch, int32_t n_dims,
                       uint64_t seed, uint64_t offset, int64_t person_offset, void* stream);

/* Same generators with the call counter read from device memory (offset = *offset_dev at execution time): a
 * captured CUDA graph of the whole SVI iteration draws fresh noise on every replay. */
int vipsy_fill_normal_dev(float* out, int32_t n_particles, int64_t batch, int32_t n_dims, uint64_t seed,
                          const uint64_t* offset_dev, int64_t person_offset, void* stream);
int vipsy_fill_uniform_dev(float* out, int32_t n_particles, int64_t batch, int32_t n_dims, uint64_t seed,
                           const uint64_t* offset_dev, int64_t person_offset, void* stream);

/* ------------------------------------------------------------------------
 * The path's one exchange in a person-sharded job (item / shared-factor / encoder gradients + loss, summed over
 * ranks) as a single kernel over NVLink peer memory instead of a collective-library call.  peer_bufs_dev: device
 * array of `world` addresses, entry r = base of rank r's symmetric buffer as mapped into THIS process.  Buffer
 * layout (identical on all ranks): flag_bytes of flags (world uint32, zero before first use), then two slots of
 * slot_bytes; a slot holds n_float floats, padding to 8 bytes, n_double doubles.  The caller has this rank's
 * payload written to slot (epoch & 1) by earlier work on `stream`; epochs start at 1 and grow by 1 per call, the
 * same on every rank.  Sums are taken in rank order: bit-identical on all ranks.
 * ---------------------------------------------------------------------- */
int vipsy_peer_allreduce(const uint64_t* peer_bufs_dev, int32_t rank, int32_t world, int64_t flag_bytes,
                         int64_t slot_bytes, int32_t n_float, int32_t n_double, float* out_float, double* out_double,
                         uint32_t epoch, void* stream);

/* ------------------------------------------------------------------------
 * The tail of the patched SVI.step (vi.py:508-515) on the device: multiply the gradient by the site's `free`
 * mask (vi.py:511-512), torch.optim.Adam's default update (one state per parameter tensor, as pyro.optim.Adam
 * keeps them), zero the gradient (vi.py:515) -- one launch for all parameter tensors of the step.
 * *step_dev is the optimiser's global step count AFTER this step's increment; a tensor that joined later has
 * step_offset > 0.  lr_dev (optional) overrides lr so that schedulers work under CUDA-graph replay.
 * ---------------------------------------------------------------------- */
typedef struct VipsyAdamTensor {
  float* param;              /* unconstrained values, updated in place     */
  float* grad;               /* gradient, zeroed when zero_grad != 0       */
  float* exp_avg;            /* first moment                               */
  float* exp_avg_sq;         /* second moment                              */
  const uint8_t* free_mask;  /* 0/1 per element or NULL                    */
  const float* lr_dev;       /* device learning rate or NULL               */
  int64_t n;                 /* elements                                   */
  int64_t step_offset;       /* global steps taken before this tensor's first */
  float lr, beta1, beta2, eps;
} VipsyAdamTensor;
int vipsy_adam_step(const VipsyAdamTensor* tensors_host, int32_t n_tensors, const int64_t* step_dev,
                    int32_t zero_grad, void* stream);

/* pyro.plate('data', N, subsample_size=B) (idx = randperm(N)[:B]): B distinct row indices of
 * [0, n_rows), uniformly distributed over the ordered B-subsets, in ONE launch with the call counter in device
 * memory (*counter_dev is read, then incremented, by the kernel) -- graph-safe, so a captured SVI iteration draws a
 * fresh minibatch on every replay.  Deterministic in (seed, counter).  batch <= VIPSY_SAMPLE_MAX_BATCH. */
enum { VIPSY_SAMPLE_MAX_BATCH = 2048 };
int vipsy_sample_indices_dev(int64_t* idx_out, int32_t batch, int64_t n_rows, uint64_t seed, uint64_t* counter_dev,
                             void* stream);

/* ------------------------------------------------------------------------
 * Amortised guide of VaeIRT, K = 1: NormEncoder (vi.py:417-435) evaluated on the stored response codes with the
 * guide's input rule NaN -> -1 (vi.py:679-682):
 *     hidden = softplus(fc1(x)),  mu = fc21(hidden),  sigma = exp(fc22(hidden)).
 * forward writes mu, sigma (the kernel inputs of vipsy_irt_elbo_grad) and the fc1 pre-activations; backward
 * takes the kernel's d_mu / d_sigma and produces the gradients of all six encoder tensors (d x is never
 * needed), deterministic: per-CTA partial rows summed in a fixed order with fp64 accumulation.  The fp32
 * response matrix and the hidden activations of the torch formulation are never materialised.
 * n_hidden <= 64 (the reference's default hidden_dim, vi.py:661); any n_items.  J <= 127 (J <= 255 with x_tiles):
 * both products run on the tensor cores (tcgen05, 3-term bf16 split of the fp32 factor, fp32 accumulation in
 * TMEM); else FFMA2 tiles.
 * ---------------------------------------------------------------------- */
typedef struct VipsyNormEncoderArgs {
  int64_t n_rows;            /* rows of y                                   */
  int64_t batch;             /* B                                           */
  int32_t n_items;           /* J                                           */
  int32_t n_hidden;          /* H <= 64                                     */
  const uint8_t* y;          /* response codes 0 / 1 / 255, n_rows x y_row_stride */
  int64_t y_row_stride;
  const int64_t* idx;        /* [B] or NULL                                 */
  const float* w1;           /* fc1.weight  [H x J] row-major               */
  const float* b1;           /* fc1.bias    [H]                             */
  const float* w21;          /* fc21.weight [H]                             */
  const float* b21;          /* fc21.bias   [1]                             */
  const float* w22;          /* fc22.weight [H]                             */
  const float* b22;          /* fc22.bias   [1]                             */
  float* mu;                 /* [B]   forward out                           */
  float* sigma;              /* [B]   forward out, backward in              */
  float* pre;                /* [B x H] fc1 pre-activations: forward out, backward in */
  const float* d_mu;         /* [B]   backward in                           */
  const float* d_sigma;      /* [B]   backward in (w.r.t. sigma, not log sigma) */
  float grad_scale;          /* multiplies every output gradient (1 / particle groups) */
  int32_t flags;             /* bit 0: SIMT kernels only (default: tcgen05 kernels when J <= 127) */
  float* d_w1;               /* [H x J] backward out (overwritten)          */
  float* d_b1;               /* [H]                                         */
  float* d_w21;              /* [H]                                         */
  float* d_b21;              /* [1]                                         */
  float* d_w22;              /* [H]                                         */
  float* d_b22;              /* [1]                                         */
  void* workspace;           /* backward: vipsy_norm_encoder_workspace_bytes */
  size_t workspace_bytes;
  const void* x_tiles;       /* optional (idx == NULL, batch == n_rows, J <= 255): the response matrix as resident
                                bf16 operand tiles from vipsy_norm_encoder_pack_tiles -- fetched with one bulk copy
                                per 128-person tile instead of being expanded from the codes in every call */
} VipsyNormEncoderArgs;
/* bytes of the tile buffer for n_rows persons (32 KB per 128 persons, 64 KB when J > 127); 0 if J > 255 */
size_t vipsy_norm_encoder_tiles_bytes(int64_t n_rows, int32_t n_items);
/* y (n_rows x y_row_stride codes) -> tiles: per 128 persons one 32 KB block in the shared-memory operand layout
 * (8 x 8 bf16 core matrices, 128 or 256 item slots, slot J = 1 for existing persons).  Done once: the matrix is
 * constant. */
int vipsy_norm_encoder_pack_tiles(const uint8_t* y, int64_t n_rows, int32_t n_items, int64_t y_row_stride,
                                  void* tiles, void* stream);
size_t vipsy_norm_encoder_workspace_bytes(const VipsyNormEncoderArgs* args);
int vipsy_norm_encoder_forward(const VipsyNormEncoderArgs* args, void* stream);
int vipsy_norm_encoder_backward(const VipsyNormEncoderArgs* args, void* stream);

/* ------------------------------------------------------------------------
 * Diagnostics: one 128-row tcgen05 tile product per operand orientation of the
 * tensor-core MIRT kernel (fp16 hi/lo split, fp32 accumulation in TMEM).
 *   variant 0: out[128 x c2] = t1[r1 x c1] . t2[r2 x c2]        (c1 == r2)
 *   variant 1: out[128 x r2] = t1[r1 x c1] . t2[r2 x c2]^T      (c1 == c2)
 *   variant 2: out[128 x c2] = t1[r1 x c1]^T . t2[r2 x c2]      (r1 == r2; rows >= c1 of out undefined)
 *   variant 3 / 4: as 0 / 1 with the A operand in tensor memory (tcgen05.st, then tcgen05.mma [d], [a], b-desc)
 * n_split = 1 (hi*hi only) .. 3 (+ hi*lo + lo*hi).  Used by tests/test_umma_selftest.py.
 * ---------------------------------------------------------------------- */
int vipsy_selftest_umma(const float* t1, int32_t r1, int32_t c1, const float* t2, int32_t r2, int32_t c2,
                        float* out, int32_t variant, int32_t swap_mn_strides, int32_t n_split, void* stream);

int vipsy_abi_version(void);
const char* vipsy_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* VIPSY_B200_H */
