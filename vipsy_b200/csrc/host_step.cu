// Host-buffer entry to the K == 1 path: what a binding of loss_and_grads(model, guide, data) (vi.py:507)
// calls when the response matrix and the variational parameters live in (pinned) HOST memory.
//
// The persons are cut into chunks; chunk i+1's host->device copy and chunk i-1's device->host copy of the
// per-person gradients overlap chunk i's kernel on three caller-owned streams.  The response rows may arrive in
// the 2-bit transport format (vipsy_pack_responses_2bit_host) and are expanded on the device right after their
// copy.  Item-level gradients and the loss are written per chunk and summed in chunk order by one small
// kernel (deterministic), then copied out.  Everything is enqueued from C: no interpreter between launches.
#include <vector>

#include "common.cuh"

namespace vipsy {

static inline size_t al256(size_t x) { return (x + 255) & ~size_t(255); }

struct HostLayout {
  size_t y, y2, mu, sigma, d_mu, d_sigma, items, a_free, grads, loss, gsum, lsum, ws, total;
  size_t ws_bytes;
  int ncol;
};

static int item_cols(int model, int J) { return J * model; }   // [d_b | d_a | d_c | d_d], K == 1

static HostLayout host_layout(const VipsyIrtHostArgs* h) {
  const VipsyIrtArgs& a = h->dev;
  const size_t N = (size_t)a.batch, J = (size_t)a.n_items;
  HostLayout L;
  size_t o = 0;
  L.ncol = item_cols(a.model, a.n_items);
  L.y = o; o += al256(N * J);
  L.y2 = o; o += h->y_packed2 ? al256(N * ((J + 3) / 4)) : 0;
  L.mu = o; o += al256(N * 4);
  L.sigma = o; o += al256(N * 4);
  L.d_mu = o; o += al256(N * 4);
  L.d_sigma = o; o += al256(N * 4);
  L.items = o; o += al256(4 * J * 4);
  L.a_free = o; o += h->a_free_host ? al256(J) : 0;
  L.grads = o; o += al256((size_t)h->n_chunks * L.ncol * 4);
  L.loss = o; o += al256((size_t)h->n_chunks * 8);
  L.gsum = o; o += al256((size_t)L.ncol * 4);
  L.lsum = o; o += 256;
  VipsyIrtArgs t = a;
  t.n_dims = 1;
  L.ws_bytes = vipsy_irt_workspace_bytes(&t);
  L.ws = o; o += al256(L.ws_bytes);
  L.total = o;
  return L;
}

__global__ void sum_chunks_kernel(const float* __restrict__ grads, const double* __restrict__ loss, int n_chunks, int ncol,
                                  float* __restrict__ gsum, double* __restrict__ lsum) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < ncol) {
    float t = 0.f;
    for (int c = 0; c < n_chunks; ++c) t += grads[(size_t)c * ncol + i];
    gsum[i] = t;
  }
  if (i == 0) {
    double t = 0.0;
    for (int c = 0; c < n_chunks; ++c) t += loss[c];
    *lsum = t;
  }
}

struct EventPool {
  std::vector<cudaEvent_t> ev;
  cudaEvent_t get(size_t i) {
    while (ev.size() <= i) {
      cudaEvent_t e;
      if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) return nullptr;
      ev.push_back(e);
    }
    return ev[i];
  }
};

}  // namespace vipsy

using namespace vipsy;

extern "C" {

size_t vipsy_irt_host_scratch_bytes(const VipsyIrtHostArgs* h) {
  if (!h || h->dev.batch < 0 || h->dev.n_items <= 0 || h->dev.model < 1 || h->dev.model > 4 || h->n_chunks < 1) return 0;
  return host_layout(h).total;
}

int vipsy_irt_host_step(const VipsyIrtHostArgs* h) {
  VIPSY_REQUIRE(h != nullptr, VIPSY_ERR_ARG, "vipsy_irt_host_step: args is NULL");
  const VipsyIrtArgs& a = h->dev;
  VIPSY_REQUIRE(a.n_dims == 1, VIPSY_ERR_UNSUPPORTED, "vipsy_irt_host_step: built for K == 1");
  VIPSY_REQUIRE(a.batch >= 0 && a.n_items > 0 && a.model >= 1 && a.model <= 4 && a.n_particles > 0 && h->n_chunks >= 1,
                VIPSY_ERR_ARG, "vipsy_irt_host_step: bad sizes");
  VIPSY_REQUIRE(h->y_host && h->mu_host && h->sigma_host && h->b_host && h->loss_host, VIPSY_ERR_ARG,
                "vipsy_irt_host_step: y, mu, sigma, b, loss host pointers are required");
  VIPSY_REQUIRE(a.model < 2 || h->a_host, VIPSY_ERR_ARG, "vipsy_irt_host_step: a is required for 2PL+");
  VIPSY_REQUIRE(a.model < 3 || h->c_host, VIPSY_ERR_ARG, "vipsy_irt_host_step: c is required for 3PL+");
  VIPSY_REQUIRE(a.model < 4 || h->d_host, VIPSY_ERR_ARG, "vipsy_irt_host_step: d is required for 4PL");
  VIPSY_REQUIRE((h->d_mu_host == nullptr) == (h->d_sigma_host == nullptr), VIPSY_ERR_ARG,
                "vipsy_irt_host_step: d_mu_host and d_sigma_host go together (both or neither)");
  VIPSY_REQUIRE(h->s_in && h->s_run && h->s_out, VIPSY_ERR_ARG, "vipsy_irt_host_step: three streams are required");
  const HostLayout L = host_layout(h);
  VIPSY_REQUIRE(h->dev_scratch && h->dev_scratch_bytes >= L.total, VIPSY_ERR_WORKSPACE,
                "vipsy_irt_host_step: scratch %zu B < required %zu B", h->dev_scratch_bytes, L.total);
  cudaStream_t s_in = (cudaStream_t)h->s_in, s_run = (cudaStream_t)h->s_run, s_out = (cudaStream_t)h->s_out;
  unsigned char* base = reinterpret_cast<unsigned char*>(h->dev_scratch);
  const int64_t N = a.batch;
  const int J = a.n_items, PB = (J + 3) / 4;
  const bool want_grad = h->item_grads_host != nullptr || h->d_mu_host != nullptr;
  // 2-bit rows go straight into the particle kernel (expanded while a tile is staged) when it is the kernel that will
  // run: S > 1, pathwise, J % 4 == 0, no sparse hint; otherwise they are expanded to a byte matrix first
  const bool direct2 = h->y_packed2 && a.n_particles > 1 && a.estimator == VIPSY_EST_PATHWISE && (J & 3) == 0 &&
                       !(a.flags & VIPSY_IRT_FLAG_SPARSE);
  static thread_local EventPool pool;

  // item parameters: rows b, a, c, d of a 4 x J device table
  float* items = reinterpret_cast<float*>(base + L.items);
  VIPSY_CUDA_CHECK(cudaMemcpyAsync(items, h->b_host, (size_t)J * 4, cudaMemcpyHostToDevice, s_in));
  if (a.model >= 2) VIPSY_CUDA_CHECK(cudaMemcpyAsync(items + J, h->a_host, (size_t)J * 4, cudaMemcpyHostToDevice, s_in));
  if (a.model >= 3) VIPSY_CUDA_CHECK(cudaMemcpyAsync(items + 2 * J, h->c_host, (size_t)J * 4, cudaMemcpyHostToDevice, s_in));
  if (a.model >= 4) VIPSY_CUDA_CHECK(cudaMemcpyAsync(items + 3 * J, h->d_host, (size_t)J * 4, cudaMemcpyHostToDevice, s_in));
  // the `free` mask of site 'a' (vi.py:511-512): applied to d_a per chunk by the kernel path, as on the device entry
  uint8_t* a_free_dev = nullptr;
  if (h->a_free_host && a.model >= 2) {
    a_free_dev = base + L.a_free;
    VIPSY_CUDA_CHECK(cudaMemcpyAsync(a_free_dev, h->a_free_host, (size_t)J, cudaMemcpyHostToDevice, s_in));
  }

  // Chunk schedule: sizes in whole "waves" of the persistent particle kernel (SMs x 4 CTAs x 4 warps x 32
  // persons per round; a chunk of 3.3 waves costs 4 rounds) when the batch is that large, else whole warps.
  // The first chunk is a single wave so that the first kernel starts after a ~2 MB copy, not a quarter of
  // the matrix.
  const int64_t wave = (int64_t)num_sms() * 4 * 4 * 32;
  int64_t first = 0, per;
  if (h->n_chunks >= 3 && N >= 6 * wave) {
    first = wave;
    per = (N - first + h->n_chunks - 2) / (h->n_chunks - 1);
    per = (per + wave - 1) / wave * wave;
  } else {
    per = (N + h->n_chunks - 1) / h->n_chunks;
    if (per >= 2 * wave) per = (per + wave - 1) / wave * wave;
  }
  per = (per + 31) / 32 * 32;
  if (per < 32) per = 32;
  int used = 0;
  for (int64_t s0 = 0; s0 < N; ++used) {
    const int64_t want = (used == 0 && first > 0) ? first : per;
    const int64_t n = (N - s0 < want) ? N - s0 : want;
    VIPSY_REQUIRE(used < h->n_chunks, VIPSY_ERR_UNSUPPORTED, "vipsy_irt_host_step: chunk schedule overflow");
    cudaEvent_t ev_in = pool.get(2 * used), ev_run = pool.get(2 * used + 1);
    VIPSY_REQUIRE(ev_in && ev_run, VIPSY_ERR_CUDA, "vipsy_irt_host_step: cudaEventCreate failed");
    uint8_t* y = base + L.y + (size_t)s0 * J;
    float* mu = reinterpret_cast<float*>(base + L.mu) + s0;
    float* sg = reinterpret_cast<float*>(base + L.sigma) + s0;
    float* dmu = reinterpret_cast<float*>(base + L.d_mu) + s0;
    float* dsg = reinterpret_cast<float*>(base + L.d_sigma) + s0;
    if (h->y_packed2)
      VIPSY_CUDA_CHECK(cudaMemcpyAsync(base + L.y2 + (size_t)s0 * PB, h->y_host + (size_t)s0 * PB, (size_t)n * PB,
                                       cudaMemcpyHostToDevice, s_in));
    else
      VIPSY_CUDA_CHECK(cudaMemcpyAsync(y, h->y_host + (size_t)s0 * J, (size_t)n * J, cudaMemcpyHostToDevice, s_in));
    VIPSY_CUDA_CHECK(cudaMemcpyAsync(mu, h->mu_host + s0, (size_t)n * 4, cudaMemcpyHostToDevice, s_in));
    VIPSY_CUDA_CHECK(cudaMemcpyAsync(sg, h->sigma_host + s0, (size_t)n * 4, cudaMemcpyHostToDevice, s_in));
    VIPSY_CUDA_CHECK(cudaEventRecord(ev_in, s_in));

    VIPSY_CUDA_CHECK(cudaStreamWaitEvent(s_run, ev_in, 0));
    if (h->y_packed2 && !direct2) {
      int rc = vipsy_unpack_responses_2bit(base + L.y2 + (size_t)s0 * PB, n, J, y, J, s_run);
      if (rc != VIPSY_OK) return rc;
    }
    VipsyIrtArgs c = a;
    c.n_rows = n; c.batch = n; c.n_dims = 1; c.cov_mode = VIPSY_COV_DIAG1;
    c.y = direct2 ? base + L.y2 + (size_t)s0 * PB : y;
    c.y_row_stride = direct2 ? PB : J;
    c.idx = nullptr; c.mu = mu; c.sigma = sg; c.chol = nullptr; c.eps = nullptr;
    c.person_offset = a.person_offset + s0;
    c.b = items; c.a = a.model >= 2 ? items + J : nullptr; c.c = a.model >= 3 ? items + 2 * J : nullptr;
    c.d = a.model >= 4 ? items + 3 * J : nullptr;
    c.a_free = a_free_dev;
    c.loss = reinterpret_cast<double*>(base + L.loss) + used;
    float* g = reinterpret_cast<float*>(base + L.grads) + (size_t)used * L.ncol;
    c.d_mu = want_grad ? dmu : nullptr; c.d_sigma = want_grad ? dsg : nullptr; c.d_chol = nullptr;
    c.d_b = want_grad ? g : nullptr;
    c.d_a = (want_grad && a.model >= 2) ? g + J : nullptr;
    c.d_c = (want_grad && a.model >= 3) ? g + 2 * J : nullptr;
    c.d_d = (want_grad && a.model >= 4) ? g + 3 * J : nullptr;
    c.workspace = base + L.ws; c.workspace_bytes = L.ws_bytes;
    c.obs_row_ptr = nullptr; c.obs_entries = nullptr; c.obs_nnz = 0;
    // the caller's count of missing cells covers the whole batch: the first chunk carries it, the others carry 0
    c.flags = (a.flags & ~VIPSY_IRT_FLAG_Y_PACKED2) | (direct2 ? VIPSY_IRT_FLAG_Y_PACKED2 : 0);
    if (a.flags & VIPSY_IRT_FLAG_MASKED_KNOWN) c.n_masked_cells = used == 0 ? a.n_masked_cells : 0;
    int rc = vipsy_irt_elbo_grad(&c, s_run);
    if (rc != VIPSY_OK) return rc;
    VIPSY_CUDA_CHECK(cudaEventRecord(ev_run, s_run));

    if (want_grad && h->d_mu_host) {
      VIPSY_CUDA_CHECK(cudaStreamWaitEvent(s_out, ev_run, 0));
      VIPSY_CUDA_CHECK(cudaMemcpyAsync(h->d_mu_host + s0, dmu, (size_t)n * 4, cudaMemcpyDeviceToHost, s_out));
      VIPSY_CUDA_CHECK(cudaMemcpyAsync(h->d_sigma_host + s0, dsg, (size_t)n * 4, cudaMemcpyDeviceToHost, s_out));
    }
    s0 += n;
  }
  // chunk sums in chunk order, then the small read-back
  float* gsum = reinterpret_cast<float*>(base + L.gsum);
  double* lsum = reinterpret_cast<double*>(base + L.lsum);
  if (used == 0) {
    VIPSY_CUDA_CHECK(cudaMemsetAsync(gsum, 0, (size_t)L.ncol * 4, s_run));
    VIPSY_CUDA_CHECK(cudaMemsetAsync(lsum, 0, 8, s_run));
  } else {
    sum_chunks_kernel<<<(L.ncol + 255) / 256, 256, 0, s_run>>>(reinterpret_cast<float*>(base + L.grads),
                                                               reinterpret_cast<double*>(base + L.loss), used,
                                                               want_grad ? L.ncol : 0, gsum, lsum);
    VIPSY_CUDA_CHECK(cudaGetLastError());
  }
  cudaEvent_t ev_done = pool.get(2 * (size_t)h->n_chunks + 2);
  VIPSY_REQUIRE(ev_done, VIPSY_ERR_CUDA, "vipsy_irt_host_step: cudaEventCreate failed");
  VIPSY_CUDA_CHECK(cudaEventRecord(ev_done, s_run));
  VIPSY_CUDA_CHECK(cudaStreamWaitEvent(s_out, ev_done, 0));
  if (want_grad && h->item_grads_host)
    VIPSY_CUDA_CHECK(cudaMemcpyAsync(h->item_grads_host, gsum, (size_t)L.ncol * 4, cudaMemcpyDeviceToHost, s_out));
  VIPSY_CUDA_CHECK(cudaMemcpyAsync(h->loss_host, lsum, 8, cudaMemcpyDeviceToHost, s_out));
  return VIPSY_OK;
}

}  // extern "C"
