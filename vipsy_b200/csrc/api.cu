// extern "C" entry points: argument validation and dispatch to the kernel families.
#include "common.cuh"

namespace vipsy {
size_t irt_uni_workspace_bytes(const VipsyIrtArgs* a);
int irt_uni_launch(const VipsyIrtArgs* a, cudaStream_t st);
size_t mirt_workspace_bytes(const VipsyIrtArgs* a);
struct MvnEncParams;
int mirt_launch(const VipsyIrtArgs* a, cudaStream_t st, MvnEncParams* enc);
int umma_selftest_launch(const float* t1, int R1, int C1, const float* t2, int R2, int C2, float* out, int variant,
                         int swap, int n_split, cudaStream_t st);
}  // namespace vipsy

namespace vipsy {
int check_irt(const VipsyIrtArgs* a) {
  VIPSY_REQUIRE(a != nullptr, VIPSY_ERR_ARG, "vipsy_irt: args is NULL");
  VIPSY_REQUIRE(a->model >= VIPSY_IRT_1PL && a->model <= VIPSY_IRT_4PL, VIPSY_ERR_ARG,
                "vipsy_irt: unknown model %d (IRT_FUN has irt_1pl..irt_4pl)", a->model);
  VIPSY_REQUIRE(a->estimator == VIPSY_EST_PATHWISE || a->estimator == VIPSY_EST_SCORE, VIPSY_ERR_ARG,
                "vipsy_irt: unknown estimator %d", a->estimator);
  VIPSY_REQUIRE(a->batch >= 0 && a->n_items > 0 && a->n_dims > 0 && a->n_particles > 0, VIPSY_ERR_ARG,
                "vipsy_irt: bad sizes B=%lld J=%d K=%d S=%d", (long long)a->batch, a->n_items, a->n_dims,
                a->n_particles);
  VIPSY_REQUIRE(a->idx != nullptr || a->batch <= a->n_rows, VIPSY_ERR_ARG,
                "vipsy_irt: batch %lld > rows %lld without idx", (long long)a->batch, (long long)a->n_rows);
  if (a->flags & VIPSY_IRT_FLAG_Y_PACKED2)
    VIPSY_REQUIRE(a->n_dims == 1 && a->n_particles > 1 && a->estimator == VIPSY_EST_PATHWISE && (a->n_items & 3) == 0 &&
                      a->idx == nullptr && a->y_row_stride == a->n_items / 4 && a->obs_row_ptr == nullptr &&
                      !(a->flags & VIPSY_IRT_FLAG_SPARSE) && (reinterpret_cast<uintptr_t>(a->y) & 3) == 0,
                  VIPSY_ERR_UNSUPPORTED,
                  "vipsy_irt: VIPSY_IRT_FLAG_Y_PACKED2 needs K == 1, S > 1, the pathwise estimator, J %% 4 == 0, no idx / "
                  "CSR / sparse hint, y_row_stride == J / 4 and a 4-byte aligned y");
  else
    VIPSY_REQUIRE(a->y_row_stride >= a->n_items, VIPSY_ERR_ARG, "vipsy_irt: y_row_stride < J");
  VIPSY_REQUIRE(a->y && a->mu && a->b && a->loss, VIPSY_ERR_ARG, "vipsy_irt: y, mu, b, loss are required");
  VIPSY_REQUIRE(a->model < VIPSY_IRT_2PL || a->a, VIPSY_ERR_ARG, "vipsy_irt: a is required for 2PL+");
  VIPSY_REQUIRE(a->model < VIPSY_IRT_3PL || a->c, VIPSY_ERR_ARG, "vipsy_irt: c is required for 3PL+");
  VIPSY_REQUIRE(a->model < VIPSY_IRT_4PL || a->d, VIPSY_ERR_ARG, "vipsy_irt: d is required for 4PL");
  VIPSY_REQUIRE(a->n_particles <= (1 << 20), VIPSY_ERR_UNSUPPORTED, "vipsy_irt: S too large");
  if (a->n_dims == 1) {
    VIPSY_REQUIRE(a->sigma, VIPSY_ERR_ARG, "vipsy_irt: sigma is required for K == 1");
    VIPSY_REQUIRE((a->d_mu == nullptr) == (a->d_sigma == nullptr), VIPSY_ERR_ARG,
                  "vipsy_irt: d_mu and d_sigma go together");
  }
  return VIPSY_OK;
}
}  // namespace vipsy

using namespace vipsy;

extern "C" {

size_t vipsy_irt_workspace_bytes(const VipsyIrtArgs* a) {
  if (!a || a->n_items <= 0 || a->n_dims <= 0 || a->model < 1 || a->model > 4) return 0;
  if (a->n_dims == 1) return irt_uni_workspace_bytes(a);
  if (a->batch < 0 || a->n_particles <= 0) return 0;
  return mirt_workspace_bytes(a);
}

// B == 0: nothing to evaluate; the outputs are defined (loss 0, item gradients 0) and no kernel needs a pointer
static int irt_empty_batch(const VipsyIrtArgs* a, cudaStream_t st) {
  const size_t J = (size_t)a->n_items, K = (size_t)a->n_dims;
  VIPSY_REQUIRE(a->loss, VIPSY_ERR_ARG, "vipsy_irt: loss is required");
  VIPSY_CUDA_CHECK(cudaMemsetAsync(a->loss, 0, 8, st));
  if (a->d_b) VIPSY_CUDA_CHECK(cudaMemsetAsync(a->d_b, 0, J * 4, st));
  if (a->d_a && a->model >= VIPSY_IRT_2PL) VIPSY_CUDA_CHECK(cudaMemsetAsync(a->d_a, 0, K * J * 4, st));
  if (a->d_c && a->model >= VIPSY_IRT_3PL) VIPSY_CUDA_CHECK(cudaMemsetAsync(a->d_c, 0, J * 4, st));
  if (a->d_d && a->model >= VIPSY_IRT_4PL) VIPSY_CUDA_CHECK(cudaMemsetAsync(a->d_d, 0, J * 4, st));
  if (a->d_chol && a->cov_mode == VIPSY_COV_SHARED) {
    const size_t n = a->chol_layout == VIPSY_CHOL_PACKED ? K * (K + 1) / 2 : K * K;
    VIPSY_CUDA_CHECK(cudaMemsetAsync(a->d_chol, 0, n * 4, st));
  }
  return VIPSY_OK;
}

int vipsy_irt_elbo_grad(const VipsyIrtArgs* a, void* stream) {
  if (a && a->batch == 0 && a->n_items > 0 && a->n_dims > 0 && a->model >= VIPSY_IRT_1PL && a->model <= VIPSY_IRT_4PL)
    return irt_empty_batch(a, (cudaStream_t)stream);
  int rc = check_irt(a);
  if (rc != VIPSY_OK) return rc;
  if (a->n_dims == 1) return irt_uni_launch(a, (cudaStream_t)stream);
  VIPSY_REQUIRE(a->model != VIPSY_IRT_1PL, VIPSY_ERR_ARG,
                "vipsy_irt_elbo_grad: irt_1pl is D*(x+b) (vi.py:29): defined for K == 1 only");
  return mirt_launch(a, (cudaStream_t)stream, nullptr);
}

int vipsy_selftest_umma(const float* t1, int32_t r1, int32_t c1, const float* t2, int32_t r2, int32_t c2, float* out,
                        int32_t variant, int32_t swap_mn_strides, int32_t n_split, void* stream) {
  return umma_selftest_launch(t1, r1, c1, t2, r2, c2, out, variant, swap_mn_strides, n_split, (cudaStream_t)stream);
}

}  // extern "C"
