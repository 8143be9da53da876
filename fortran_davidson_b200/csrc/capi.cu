// extern "C" surface declared in include/davidson_b200.h.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <memory>
#include <mutex>
#include <vector>

#include "solver.cuh"

using namespace dav;

namespace dav {
static thread_local std::string g_last_error;
void set_last_error(const std::string& s) { g_last_error = s; }
}  // namespace dav

#define API_BEGIN try {
#define API_END                                   \
  }                                               \
  catch (const dav::Error& e) {                   \
    dav::set_last_error(e.msg);                   \
    (void)cudaGetLastError();                     \
    return e.code;                                \
  }                                               \
  catch (const std::exception& e) {               \
    dav::set_last_error(e.what());                \
    return DAV_ERR_INVALID;                       \
  }                                               \
  return DAV_OK;

namespace {

int parse_method(const char* method) {
  if (!method) DAV_THROW(DAV_ERR_INVALID, "method is NULL");
  std::string m(method);
  while (!m.empty() && (m.back() == ' ' || m.back() == '\0')) m.pop_back();
  if (m == "DPR") return DAV_METHOD_DPR;
  if (m == "GJD") return DAV_METHOD_GJD;
  // the reference's `select case` has no default branch (davidson.f90:656-669): the correction
  // would be left undefined; reject instead
  DAV_THROW(DAV_ERR_INVALID, "unknown correction method '%s' (expected DPR or GJD)", method);
}

void need(bool cond, const char* what) {
  if (!cond) DAV_THROW(DAV_ERR_INVALID, "%s", what);
}

// small RAII context for the utility entry points: device 0 (or the current one), own stream
// Device of the calls that take no handle (the drop-in solver calls and the lapack_wrapper / array_utils mirrors):
// dav_set_default_device(), else the environment variable DAV_DEVICE, else device 0.
int g_default_device = -1;
std::mutex g_dropin_mu;
// cached handle of dav_generalized_eigensolver_dense; a raw pointer on purpose: it is never destroyed by a static
// destructor (CUDA calls after the runtime has shut down), only by dav_release_cache() or an error
dav_solver* g_dropin = nullptr;
void drop_cached_handle() {
  delete g_dropin;
  g_dropin = nullptr;
}
int default_device() {
  if (g_default_device >= 0) return g_default_device;
  const char* e = std::getenv("DAV_DEVICE");
  return e ? std::max(0, std::atoi(e)) : 0;
}

struct Ctx {
  cudaStream_t s = nullptr;
  Ctx() {
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count <= 0) {
      (void)cudaGetLastError();
      DAV_THROW(DAV_ERR_CUDA, "no CUDA device available; this library has no CPU fallback");
    }
    const int dev = default_device();
    if (dev >= count) DAV_THROW(DAV_ERR_INVALID, "default device %d out of range (%d devices)", dev, count);
    CK(cudaSetDevice(dev));
    CK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
  }
  ~Ctx() {
    if (s) cudaStreamDestroy(s);
  }
  void sync() { CK(cudaStreamSynchronize(s)); }
};

void h2d(double* d, const double* h, size_t n, cudaStream_t s) {
  CK(cudaMemcpyAsync(d, h, n * 8, cudaMemcpyHostToDevice, s));
}
void d2h(double* h, const double* d, size_t n, cudaStream_t s) {
  CK(cudaMemcpyAsync(h, d, n * 8, cudaMemcpyDeviceToHost, s));
}

void check_status_dev(Ctx& c, int* status, const char* where) {
  int h = 0;
  CK(cudaMemcpyAsync(&h, status, sizeof(int), cudaMemcpyDeviceToHost, c.s));
  c.sync();
  if (h & 2) DAV_THROW(DAV_ERR_NOT_POSDEF, "%s: matrix is not positive definite", where);
  if (h & 1) DAV_THROW(DAV_ERR_NO_CONVERGENCE, "%s: eigensolver failed (NaN input or no convergence)", where);
}

// all eigenpairs of (mtx, stx) on device; w ascending, vec k x k
void eigensolve_dev(Ctx& c, int k, const double* mtx_h, const double* stx_h, double* w_h, double* vec_h, int ncols) {
  const size_t kk = (size_t)k * k;
  DevBuf<double> S1, S2, U, sv, Tm, Z, Y, w, scratch;
  DevBuf<int> status;
  S1.alloc(kk); Y.alloc(kk); w.alloc(k); scratch.alloc(sym_eigh_scratch_doubles(k)); status.alloc(1);
  CK(cudaMemsetAsync(status.p, 0, sizeof(int), c.s));
  h2d(S1.p, mtx_h, kk, c.s);
  if (!stx_h) {
    sym_eigh(c.s, k, S1.p, Y.p, w.p, scratch.p, status.p);
  } else {
    S2.alloc(kk); U.alloc(kk); sv.alloc(k); Tm.alloc(kk); Z.alloc(kk);
    h2d(S2.p, stx_h, kk, c.s);
    sym_eigh(c.s, k, S2.p, U.p, sv.p, scratch.p, status.p);
    scale_cols_rsqrt_checked(c.s, k, U.p, sv.p, Tm.p, status.p);
    symmetrize_from_upper(c.s, k, S1.p, k);
    gemm(c.s, false, k, k, k, 1.0, S1.p, k, Tm.p, k, 0.0, Z.p, k, nullptr, 0);
    gemm(c.s, true, k, k, k, 1.0, Tm.p, k, Z.p, k, 0.0, S1.p, k, nullptr, 0);
    sym_eigh(c.s, k, S1.p, Z.p, w.p, scratch.p, status.p);
    gemm(c.s, false, k, k, k, 1.0, Tm.p, k, Z.p, k, 0.0, Y.p, k, nullptr, 0);
  }
  check_status_dev(c, status.p, "lapack_generalized_eigensolver");
  d2h(w_h, w.p, ncols, c.s);
  d2h(vec_h, Y.p, (size_t)k * ncols, c.s);
  c.sync();
}

}  // namespace

extern "C" {

const char* dav_last_error(void) { return dav::g_last_error.c_str(); }
int dav_version(void) { return DAV_VERSION; }

int dav_device_count(void) {
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess) {
    (void)cudaGetLastError();
    return 0;
  }
  return count;
}

int dav_set_default_device(int device) {
  API_BEGIN
  need(device >= -1, "device must be >= 0 (or -1: back to DAV_DEVICE / 0)");
  if (device >= 0) {
    const int count = dav_device_count();
    need(count > 0, "no CUDA device available; this library has no CPU fallback");
    need(device < count, "device out of range");
  }
  g_default_device = device;
  API_END
}

int dav_alloc_pinned(size_t bytes, void** ptr) {
  API_BEGIN
  need(ptr != nullptr && bytes > 0, "bad arguments");
  CK(cudaMallocHost(ptr, bytes));
  API_END
}

int dav_free_pinned(void* ptr) {
  API_BEGIN
  if (ptr) CK(cudaFreeHost(ptr));
  API_END
}

int dav_partition_rows(int64_t n, int world_size, int rank, int64_t* row_begin, int64_t* row_end) {
  if (n < 0 || world_size < 1 || rank < 0 || rank >= world_size || !row_begin || !row_end) {
    dav::set_last_error("dav_partition_rows: bad arguments");
    return DAV_ERR_INVALID;
  }
  int64_t chunk = n;
  if (world_size > 1) chunk = ((n + world_size - 1) / world_size + 127) / 128 * 128;
  int64_t b = std::min<int64_t>(n, (int64_t)rank * chunk);
  int64_t e = std::min<int64_t>(n, b + chunk);
  *row_begin = b;
  *row_end = e;
  return DAV_OK;
}

int dav_bench_fp64_pipe(dav_solver_t* h, int reps, double* dmma_tflops) {
  API_BEGIN
  need(h && dmma_tflops && reps >= 1, "bad arguments");
  CK(cudaSetDevice(h->device));
  *dmma_tflops = dmma_peak_tflops(h->stream, reps);
  API_END
}

int dav_debug_matvec_rect(int device, int64_t m, int64_t k, int b, double* max_abs_diff, double* scale) {
  API_BEGIN
  need(m >= 1 && k >= 1 && b >= 1 && b <= 128 && max_abs_diff && scale, "bad arguments");
  CK(cudaSetDevice(device));
  Ctx c;
  const int64_t lda = round_up(m, 16);
  DevBuf<double> A, X, W1, W2;
  A.alloc((size_t)lda * k); X.alloc((size_t)k * b); W1.alloc((size_t)m * b); W2.alloc((size_t)m * b);
  fill_random(c.s, A.p, (size_t)lda * k, 0xA11CEULL);
  fill_random(c.s, X.p, (size_t)k * b, 0xB0BULL);
  MatvecPlan* plan = matvec_plan_create(A.p, m, k, lda, b);
  try {
    matvec_dmma(c.s, plan, b, X.p, k, W1.p, m);
    gemm(c.s, false, m, b, k, 1.0, A.p, lda, X.p, k, 0.0, W2.p, m, nullptr, 0);
    c.sync();
  } catch (...) {
    matvec_plan_destroy(plan);
    throw;
  }
  matvec_plan_destroy(plan);
  std::vector<double> h1((size_t)m * b), h2((size_t)m * b);
  d2h(h1.data(), W1.p, h1.size(), c.s);
  d2h(h2.data(), W2.p, h2.size(), c.s);
  c.sync();
  double d = 0.0, sc = 0.0;
  for (size_t i = 0; i < h1.size(); ++i) {
    const double e = std::fabs(h1[i] - h2[i]);
    if (!(e <= d)) d = e;  // NaN-propagating maximum
    sc = std::max(sc, std::fabs(h2[i]));
  }
  *max_abs_diff = d;
  *scale = sc;
  API_END
}

int dav_debug_matvec_schedule(int64_t m, int64_t k, int b, int num_sms, int schedule, long long* info) {
  API_BEGIN
  const int rc = matvec_schedule_selftest(m, k, b, num_sms, schedule, info);
  if (rc != 0) DAV_THROW(DAV_ERR_INVALID, "matvec schedule self-test failed: check %d", rc);
  API_END
}

// ---------------------------------------------------------------------------------------------
// handle API
// ---------------------------------------------------------------------------------------------
int dav_get_unique_id(void* id128) {
  API_BEGIN
  need(id128 != nullptr, "id buffer is NULL");
  Comm::get_unique_id(id128);
  API_END
}

int dav_create(dav_solver_t** h, int device) {
  API_BEGIN
  need(h != nullptr, "handle pointer is NULL");
  *h = new dav_solver(device, 0, 1, nullptr);
  API_END
}

int dav_create_distributed(dav_solver_t** h, int device, int rank, int world_size, const void* id128) {
  API_BEGIN
  need(h != nullptr, "handle pointer is NULL");
  need(world_size >= 1 && rank >= 0 && rank < world_size, "bad rank / world_size");
  need(world_size == 1 || id128 != nullptr, "NCCL id is NULL");
  *h = new dav_solver(device, rank, world_size, id128);
  API_END
}

int dav_destroy(dav_solver_t* h) {
  API_BEGIN
  delete h;
  API_END
}

int dav_matrix_generate_diagonal_dominant(dav_solver_t* h, int which, int64_t n, double sparsity, int has_diag_val,
                                          double diag_val, uint64_t seed) {
  API_BEGIN
  need(h && (which == 0 || which == 1), "bad handle / slot");
  h->generate_diagonal_dominant(which, n, sparsity, has_diag_val, diag_val, seed);
  API_END
}

int dav_matrix_upload(dav_solver_t* h, int which, int64_t n, const double* host_matrix, int64_t ld) {
  API_BEGIN
  need(h && (which == 0 || which == 1), "bad handle / slot");
  h->upload(which, n, host_matrix, ld);
  API_END
}

int dav_matrix_upload_rows(dav_solver_t* h, int which, int64_t n, const double* host_rows, int64_t ld) {
  API_BEGIN
  need(h && (which == 0 || which == 1), "bad handle / slot");
  h->upload_rows(which, n, host_rows, ld);
  API_END
}

int dav_matrix_set_operator(dav_solver_t* h, int which, int64_t n, int op) {
  API_BEGIN
  need(h && (which == 0 || which == 1), "bad handle / slot");
  h->set_operator(which, n, op);
  API_END
}

int dav_matrix_set_callback(dav_solver_t* h, int which, int64_t n, dav_gemv_fn fn, void* ctx, const double* diag) {
  API_BEGIN
  need(h && (which == 0 || which == 1), "bad handle / slot");
  h->set_callback(which, n, fn, ctx, diag);
  API_END
}

int dav_matrix_set_device_callback(dav_solver_t* h, int which, int64_t n, dav_device_gemv_fn fn, void* ctx,
                                   const double* diag) {
  API_BEGIN
  need(h && (which == 0 || which == 1), "bad handle / slot");
  h->set_device_callback(which, n, fn, ctx, diag);
  API_END
}

int dav_set_profiling(dav_solver_t* h, int per_phase_spans) {
  API_BEGIN
  need(h, "bad handle");
  h->profile_spans = per_phase_spans != 0;
  API_END
}

int dav_matrix_clear(dav_solver_t* h, int which) {
  API_BEGIN
  need(h && (which == 0 || which == 1), "bad handle / slot");
  h->clear_matrix(which);
  API_END
}

int dav_matrix_download(dav_solver_t* h, int which, double* host_rows, int64_t ld) {
  API_BEGIN
  need(h && (which == 0 || which == 1) && host_rows, "bad handle / slot / buffer");
  h->download(which, host_rows, ld);
  API_END
}

int dav_solve(dav_solver_t* h, int lowest, int method, int max_iterations, double tolerance, int max_dim_sub,
              double* eigenvalues, double* eigenvectors, int64_t ldv, int* iters) {
  API_BEGIN
  need(h && eigenvalues && iters, "bad handle / output pointers");
  h->solve(lowest, method, max_iterations, tolerance, max_dim_sub, eigenvalues, eigenvectors, ldv, iters);
  API_END
}

int dav_solve_local(dav_solver_t* h, int lowest, int method, int max_iterations, double tolerance, int max_dim_sub,
                    double* eigenvalues, double* eigenvectors_local, int64_t ldv_local, int* iters) {
  API_BEGIN
  need(h && eigenvalues && iters, "bad handle / output pointers");
  h->local_vectors = true;
  try {
    h->solve(lowest, method, max_iterations, tolerance, max_dim_sub, eigenvalues, eigenvectors_local, ldv_local, iters);
  } catch (...) {
    h->local_vectors = false;
    throw;
  }
  h->local_vectors = false;
  API_END
}

int dav_get_stats(dav_solver_t* h, dav_stats_t* out) {
  API_BEGIN
  need(h && out, "bad handle / output");
  *out = h->stats;
  API_END
}

int dav_set_matvec_impl(dav_solver_t* h, int impl) {
  API_BEGIN
  need(h && impl >= DAV_MATVEC_AUTO && impl <= DAV_MATVEC_TMA_DMMA, "bad handle / impl");
  h->matvec_impl = impl;
  API_END
}

int dav_block_matvec(dav_solver_t* h, int which, int64_t b, const double* x, int64_t ldx, double* w, int64_t ldw) {
  API_BEGIN
  need(h && (which == 0 || which == 1) && x && w && b >= 1, "bad arguments");
  CK(cudaSetDevice(h->device));
  need(h->mat[which].kind != dav_solver::NONE, "no matrix in that slot");
  const int64_t n = h->n, nl = h->nl;
  need(ldx >= n && ldw >= std::max<int64_t>(nl, 1), "leading dimension too small");
  DevBuf<double> X, W;
  X.alloc((size_t)n * b);
  W.alloc((size_t)std::max<int64_t>(nl, 1) * b);
  CK(cudaMemcpy2DAsync(X.p, (size_t)n * 8, x, (size_t)ldx * 8, (size_t)n * 8, (size_t)b, cudaMemcpyHostToDevice,
                       h->stream));
  h->spans.clear();
  h->ev_used = 0;
  h->apply_full(which, X.p, n, (int)b, W.p, std::max<int64_t>(nl, 1));
  if (nl > 0)
    CK(cudaMemcpy2DAsync(w, (size_t)ldw * 8, W.p, (size_t)nl * 8, (size_t)nl * 8, (size_t)b, cudaMemcpyDeviceToHost,
                         h->stream));
  CK(cudaStreamSynchronize(h->stream));
  API_END
}

int dav_bench_block_matvec(dav_solver_t* h, int which, int64_t b, int reps, float* ms_out) {
  API_BEGIN
  need(h && (which == 0 || which == 1) && ms_out && b >= 1 && reps >= 1, "bad arguments");
  CK(cudaSetDevice(h->device));
  need(h->mat[which].kind != dav_solver::NONE, "no matrix in that slot");
  const int64_t n = h->n, nl = std::max<int64_t>(h->nl, 1);
  DevBuf<double> X, W;
  X.alloc((size_t)n * b);
  W.alloc((size_t)nl * b);
  fill_random(h->stream, X.p, (size_t)n * b, 0xB10CULL);
  std::vector<cudaEvent_t> ev(reps + 1);
  for (auto& e : ev) CK(cudaEventCreate(&e));
  h->spans.clear();
  h->ev_used = 0;
  h->apply_full(which, X.p, n, (int)b, W.p, nl);  // untimed: builds the plan
  CK(cudaStreamSynchronize(h->stream));
  for (int r = 0; r < reps; ++r) {
    h->spans.clear();
    h->ev_used = 0;
    CK(cudaEventRecord(ev[r], h->stream));
    h->apply_full(which, X.p, n, (int)b, W.p, nl);
  }
  CK(cudaEventRecord(ev[reps], h->stream));
  CK(cudaStreamSynchronize(h->stream));
  for (int r = 0; r < reps; ++r) CK(cudaEventElapsedTime(&ms_out[r], ev[r], ev[r + 1]));
  for (auto& e : ev) cudaEventDestroy(e);
  API_END
}

int dav_debug_collective(dav_solver_t* h, int kind, int64_t count, int reps, double* out2) {
  API_BEGIN
  need(h && out2, "bad handle / output");
  CK(cudaSetDevice(h->device));
  need(h->comm.active(), "dav_debug_collective needs a distributed handle");
  need(kind <= 1 || h->n > 0, "set a matrix first (the gathers use the handle's row partition)");
  h->comm.debug_exchange(kind, count, reps, h->n, h->nl, h->row0, h->stream, out2);
  API_END
}

int dav_debug_chol_inv(int b, const double* g, double* t, double* flag, float* ms) {
  API_BEGIN
  need(b >= 1 && g && t && flag, "bad arguments");
  Ctx c;
  DevBuf<double> G, T, F, W;
  const size_t bb = (size_t)b * b;
  G.alloc(bb); T.alloc(bb); F.alloc(1); W.alloc(bb + b);
  h2d(G.p, g, bb, c.s);
  need(chol_inv_upper(c.s, b, G.p, T.p, F.p, W.p, W.n), "chol_inv_upper: no kernel for this width");  // warm-up
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  CK(cudaEventRecord(e0, c.s));
  chol_inv_upper(c.s, b, G.p, T.p, F.p, W.p, W.n);
  CK(cudaEventRecord(e1, c.s));
  d2h(t, T.p, bb, c.s);
  d2h(flag, F.p, 1, c.s);
  c.sync();
  float dt = 0.f;
  CK(cudaEventElapsedTime(&dt, e0, e1));
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  if (ms) *ms = dt;
  API_END
}

int dav_debug_pip_small(int mode, int kold, int b, const double* gall, double* z, double* metrics, int* launched) {
  API_BEGIN
  need((mode == 0 || mode == 1) && kold >= 1 && b >= 1 && gall && z && metrics && launched, "bad arguments");
  Ctx c;
  const size_t cnt = (size_t)(kold + b) * b;
  DevBuf<double> G, Z, M;
  G.alloc(cnt); Z.alloc(cnt); M.alloc(4);
  h2d(G.p, gall, cnt, c.s);
  CK(cudaMemsetAsync(Z.p, 0, cnt * 8, c.s));
  CK(cudaMemsetAsync(M.p, 0, 4 * 8, c.s));
  *launched = pip_small(c.s, mode, kold, b, G.p, Z.p, M.p) ? 1 : 0;
  d2h(z, Z.p, cnt, c.s);
  d2h(metrics, M.p, 4, c.s);
  c.sync();
  API_END
}

int dav_debug_gemm_bench(char transA, int64_t m, int64_t n, int64_t k, int reps, int to_partials, float* ms_out,
                         double* max_err) {
  API_BEGIN
  need((transA == 'N' || transA == 'T') && m >= 1 && n >= 1 && k >= 1 && reps >= 1 && ms_out, "bad arguments");
  Ctx c;
  const bool ta = transA == 'T';
  // operands as the solver has them: tall blocks with a padded leading dimension, small factors dense
  const int64_t lda = ta ? round_up(k, 16) : round_up(m, 16), ldb = ta ? round_up(k, 16) : k;
  DevBuf<double> A, B, Cm, Cr, ws;
  A.alloc((size_t)lda * (ta ? m : k));
  B.alloc((size_t)ldb * n);
  Cm.alloc((size_t)m * n);
  Cr.alloc((size_t)m * n);
  ws.alloc(std::max<size_t>((size_t)m * n * 600, (size_t)1 << 22));
  fill_random(c.s, A.p, A.n, 11);
  fill_random(c.s, B.p, B.n, 12);
  int parts = 0;
  auto run = [&]() {
    if (to_partials && ta) gemm(c.s, ta, m, n, k, 1.0, A.p, lda, B.p, ldb, 0.0, nullptr, 0, ws.p, ws.n, &parts);
    else gemm(c.s, ta, m, n, k, 1.0, A.p, lda, B.p, ldb, 0.0, Cm.p, m, ws.p, ws.n);
  };
  run();
  std::vector<cudaEvent_t> ev(2 * (size_t)reps);
  for (auto& e : ev) CK(cudaEventCreate(&e));
  for (int r = 0; r < reps; ++r) {
    CK(cudaEventRecord(ev[2 * r], c.s));
    run();
    CK(cudaEventRecord(ev[2 * r + 1], c.s));
  }
  c.sync();
  for (int r = 0; r < reps; ++r) CK(cudaEventElapsedTime(&ms_out[r], ev[2 * r], ev[2 * r + 1]));
  for (auto& e : ev) cudaEventDestroy(e);
  if (max_err) {  // against the SIMT kernel (the in-library reference), only for the direct (non-partial) form
    *max_err = -1.0;
    if (!(to_partials && ta)) {
      setenv("DAV_GEMM_IMPL", "0", 1);
      gemm(c.s, ta, m, n, k, 1.0, A.p, lda, B.p, ldb, 0.0, Cr.p, m, ws.p, ws.n);
      unsetenv("DAV_GEMM_IMPL");
      std::vector<double> h1((size_t)m * n), h2((size_t)m * n);
      d2h(h1.data(), Cm.p, h1.size(), c.s);
      d2h(h2.data(), Cr.p, h2.size(), c.s);
      c.sync();
      double e = 0.0;
      for (size_t i = 0; i < h1.size(); ++i) e = std::max(e, std::fabs(h1[i] - h2[i]));
      *max_err = e;
    }
  }
  API_END
}

int dav_comm_info(dav_solver_t* h, int* peer_transport, long long* peer_calls, long long* nccl_calls) {
  API_BEGIN
  need(h, "bad handle");
  if (peer_transport) *peer_transport = h->comm.peer() ? 1 : 0;
  if (peer_calls) *peer_calls = h->comm.peer_calls;
  if (nccl_calls) *nccl_calls = h->comm.nccl_calls;
  API_END
}

// ---------------------------------------------------------------------------------------------
// drop-in solver calls
// ---------------------------------------------------------------------------------------------
int dav_generalized_eigensolver_dense(int64_t n, const double* matrix, int64_t lda, const double* second_matrix,
                                      int64_t ldb, int lowest, const char* method, int max_iterations,
                                      double tolerance, int max_dim_sub, double* eigenvalues, double* eigenvectors,
                                      int64_t ldv, int* iters) {
  API_BEGIN
  need(matrix && eigenvalues && eigenvectors && iters, "NULL argument");
  const int m = parse_method(method);
  // One cached handle per process serves the drop-in calls: a caller that diagonalises a matrix of the same size again
  // and again (the reference's use inside SCF / BSE loops) reuses the device block of the matrix, its TMA plan, the
  // workspace and the page-locked staging instead of cudaMalloc + cudaFree of n^2 doubles per call.
  // dav_release_cache() frees it; DAV_DROPIN_CACHE=0 restores one handle per call.
  static const bool use_cache = [] { const char* e = std::getenv("DAV_DROPIN_CACHE"); return !(e && std::atoi(e) == 0); }();
  std::unique_lock<std::mutex> lock(g_dropin_mu);
  std::unique_ptr<dav_solver> local;
  dav_solver* s = nullptr;
  if (use_cache) {
    if (g_dropin && g_dropin->device != default_device()) drop_cached_handle();
    if (!g_dropin) g_dropin = new dav_solver(default_device(), 0, 1, nullptr);
    s = g_dropin;
  } else {
    local.reset(new dav_solver(default_device(), 0, 1, nullptr));
    s = local.get();
  }
  try {
    if (s->n != n) {  // another problem size: start from a clean handle state
      s->clear_matrix(0);
      s->clear_matrix(1);
    }
    s->upload(0, n, matrix, lda);
    if (second_matrix) s->upload(1, n, second_matrix, ldb);
    else s->clear_matrix(1);
    s->solve(lowest, m, max_iterations, tolerance, max_dim_sub, eigenvalues, eigenvectors, ldv, iters);
  } catch (...) {
    if (use_cache) drop_cached_handle();  // unknown state after an error
    throw;
  }
  API_END
}

int dav_upload_bytes(dav_solver_t* h, double* bytes) {
  API_BEGIN
  need(bytes != nullptr, "bad arguments");
  std::unique_lock<std::mutex> lock(g_dropin_mu);
  const dav_solver* s = h ? h : g_dropin;
  *bytes = s ? s->last_upload_bytes : 0.0;
  API_END
}

int dav_release_cache(void) {
  API_BEGIN
  std::unique_lock<std::mutex> lock(g_dropin_mu);
  drop_cached_handle();
  API_END
}

int dav_generalized_eigensolver_free(int64_t n, dav_gemv_fn fun_matrix_gemv, void* ctx_matrix,
                                     dav_gemv_fn fun_second_matrix_gemv, void* ctx_second, const double* diag_matrix,
                                     const double* diag_second_matrix, int lowest, const char* method,
                                     int max_iterations, double tolerance, int max_dim_sub, double* eigenvalues,
                                     double* ritz_vectors, int64_t ldv, int* iters) {
  API_BEGIN
  need(fun_matrix_gemv && fun_second_matrix_gemv && eigenvalues && ritz_vectors && iters, "NULL argument");
  const int m = parse_method(method);
  std::unique_ptr<dav_solver> s(new dav_solver(default_device(), 0, 1, nullptr));
  s->set_callback(0, n, fun_matrix_gemv, ctx_matrix, diag_matrix);
  s->set_callback(1, n, fun_second_matrix_gemv, ctx_second, diag_second_matrix);
  s->solve(lowest, m, max_iterations, tolerance, max_dim_sub, eigenvalues, ritz_vectors, ldv, iters);
  API_END
}

int dav_generalized_eigensolver_free_builtin(int64_t n, int op_matrix, int op_second_matrix, int lowest,
                                             const char* method, int max_iterations, double tolerance,
                                             int max_dim_sub, double* eigenvalues, double* ritz_vectors, int64_t ldv,
                                             int* iters) {
  API_BEGIN
  need(eigenvalues && ritz_vectors && iters, "NULL argument");
  const int m = parse_method(method);
  std::unique_ptr<dav_solver> s(new dav_solver(default_device(), 0, 1, nullptr));
  s->set_operator(0, n, op_matrix);
  s->set_operator(1, n, op_second_matrix);
  s->solve(lowest, m, max_iterations, tolerance, max_dim_sub, eigenvalues, ritz_vectors, ldv, iters);
  API_END
}

// ---------------------------------------------------------------------------------------------
// array_utils / lapack_wrapper mirrors
// ---------------------------------------------------------------------------------------------
int dav_generate_diagonal_dominant(int64_t m, double sparsity, const double* diag_val, uint64_t seed, double* arr,
                                   int64_t ld) {
  API_BEGIN
  need(m >= 1 && arr && ld >= m, "bad arguments");
  Ctx c;
  DevBuf<double> A;
  A.alloc((size_t)m * m);
  gen_diag_dominant(c.s, A.p, m, m, m, 0, sparsity, diag_val ? 1 : 0, diag_val ? *diag_val : 0.0, seed);
  CK(cudaMemcpy2DAsync(arr, (size_t)ld * 8, A.p, (size_t)m * 8, (size_t)m * 8, (size_t)m, cudaMemcpyDeviceToHost,
                       c.s));
  c.sync();
  API_END
}

int dav_generate_preconditioner(int64_t n, const double* diag, int dim_sub, double* precond, int64_t ld) {
  API_BEGIN
  need(n >= 1 && diag && precond && dim_sub >= 1 && dim_sub <= n && ld >= n, "bad arguments");
  Ctx c;
  need(dim_sub <= 1024, "generate_preconditioner: dim_sub <= 1024 supported");
  DevBuf<double> d, val, sv;
  DevBuf<int64_t> idx, si;
  DevBuf<int> status;
  d.alloc(n); val.alloc(dim_sub); idx.alloc(dim_sub); status.alloc(1);
  sv.alloc(topk_scratch_entries(n, dim_sub)); si.alloc(topk_scratch_entries(n, dim_sub));
  CK(cudaMemsetAsync(status.p, 0, sizeof(int), c.s));
  h2d(d.p, diag, n, c.s);
  topk_smallest(c.s, d.p, nullptr, n, 0, dim_sub, val.p, idx.p, status.p, sv.p, si.p);
  std::vector<int64_t> hidx(dim_sub);
  CK(cudaMemcpyAsync(hidx.data(), idx.p, (size_t)dim_sub * 8, cudaMemcpyDeviceToHost, c.s));
  check_status_dev(c, status.p, "generate_preconditioner");
  for (int j = 0; j < dim_sub; ++j) {
    std::fill(precond + (size_t)j * ld, precond + (size_t)j * ld + n, 0.0);
    precond[(size_t)j * ld + hidx[j]] = 1.0;
  }
  API_END
}

double dav_norm_value(int64_t n, const double* vector) {
  double r = 0.0;
  if (dav_norm(n, vector, &r) != DAV_OK) return std::nan("");
  return r;
}

int dav_norm(int64_t n, const double* vector, double* result) {
  API_BEGIN
  need(n >= 0 && result && (vector || n == 0), "bad arguments");
  Ctx c;
  DevBuf<double> v, part, out;
  v.alloc(std::max<int64_t>(n, 1)); part.alloc(64); out.alloc(1);
  if (n) h2d(v.p, vector, n, c.s);
  col_norms2(c.s, n, 1, v.p, std::max<int64_t>(n, 1), part.p, out.p);
  double h = 0.0;
  d2h(&h, out.p, 1, c.s);
  c.sync();
  *result = std::sqrt(h);
  API_END
}

int dav_lapack_generalized_eigensolver(int dim, const double* mtx, const double* stx, double* eigenvalues,
                                       double* eigenvectors) {
  API_BEGIN
  need(dim >= 1 && mtx && eigenvalues && eigenvectors, "bad arguments");
  Ctx c;
  eigensolve_dev(c, dim, mtx, stx, eigenvalues, eigenvectors, dim);
  API_END
}

int dav_sym_eigh_info(int dim, const double* mtx, double* eigenvalues, double* eigenvectors, double* info, int reps,
                      float* ms_out) {
  API_BEGIN
  need(dim >= 1 && mtx && eigenvalues && eigenvectors && info && reps >= 0 && (reps == 0 || ms_out), "bad arguments");
  Ctx c;
  const int k = dim;
  const size_t kk = (size_t)k * k;
  DevBuf<double> S1, Y, w, scratch;
  DevBuf<int> status;
  S1.alloc(kk); Y.alloc(kk); w.alloc(k); scratch.alloc(sym_eigh_scratch_doubles(k)); status.alloc(1);
  CK(cudaMemsetAsync(status.p, 0, sizeof(int), c.s));
  h2d(S1.p, mtx, kk, c.s);
  std::vector<cudaEvent_t> ev(reps + 1);
  for (auto& e : ev) CK(cudaEventCreate(&e));
  sym_eigh(c.s, k, S1.p, Y.p, w.p, scratch.p, status.p);
  for (int r = 0; r < reps; ++r) {
    CK(cudaEventRecord(ev[r], c.s));
    sym_eigh(c.s, k, S1.p, Y.p, w.p, scratch.p, status.p);
  }
  CK(cudaEventRecord(ev[reps], c.s));
  check_status_dev(c, status.p, "sym_eigh");
  for (int r = 0; r < reps; ++r) CK(cudaEventElapsedTime(&ms_out[r], ev[r], ev[r + 1]));
  for (auto& e : ev) cudaEventDestroy(e);
  d2h(eigenvalues, w.p, k, c.s);
  d2h(eigenvectors, Y.p, kk, c.s);
  for (int q = 0; q < 8; ++q) info[q] = 0.0;
  info[0] = -1.0;
  if (sym_eigh_uses_tridiag(k)) {
    double f[9];
    CK(cudaMemcpyAsync(f, sym_eigh_flags(scratch.p, k), sizeof(f), cudaMemcpyDeviceToHost, c.s));
    c.sync();
    int acc = 0;
    std::memcpy(&acc, &f[8], sizeof(int));
    info[0] = acc; info[1] = f[0]; info[2] = f[1]; info[3] = f[2];
    for (int q = 0; q < 4; ++q) info[4 + q] = f[3 + q];
  }
  c.sync();
  API_END
}

int dav_lapack_generalized_eigensolver_lowest(int dim, const double* mtx, const double* stx, int lowest,
                                              double* eigenvalues, double* eigenvectors) {
  API_BEGIN
  need(dim >= 1 && mtx && stx && eigenvalues && eigenvectors && lowest >= 1 && lowest <= dim, "bad arguments");
  Ctx c;
  eigensolve_dev(c, dim, mtx, stx, eigenvalues, eigenvectors, lowest);
  API_END
}

// CholeskyQR2: G = A^T A = R^T R, A <- A R^-1, twice.  Q spans the columns like DGEQRF+DORGQR's Q
// and equals it up to column signs (R has a positive diagonal here).
int dav_lapack_qr(int64_t m, int n, double* basis, int64_t ld) {
  API_BEGIN
  need(m >= 1 && n >= 1 && basis && ld >= m && n <= m, "bad arguments (needs m >= n)");
  Ctx c;
  DevBuf<double> A, B, G, Rinv, ws;
  DevBuf<int> status;
  const size_t nn = (size_t)n * n;
  A.alloc((size_t)m * n); B.alloc((size_t)m * n); G.alloc(nn); Rinv.alloc(nn); status.alloc(1);
  ws.alloc(std::max<size_t>(nn * 64, (size_t)1 << 20));
  CK(cudaMemsetAsync(status.p, 0, sizeof(int), c.s));
  CK(cudaMemcpy2DAsync(A.p, (size_t)m * 8, basis, (size_t)ld * 8, (size_t)m * 8, (size_t)n, cudaMemcpyHostToDevice,
                       c.s));
  double* cur = A.p;
  double* other = B.p;
  for (int pass = 0; pass < 2; ++pass) {
    gemm(c.s, true, n, n, m, 1.0, cur, m, cur, m, 0.0, G.p, n, ws.p, ws.n);
    cholesky_upper(c.s, n, G.p, n, status.p);
    invert_upper(c.s, n, G.p, n, Rinv.p);
    gemm(c.s, false, m, n, n, 1.0, cur, m, Rinv.p, n, 0.0, other, m, nullptr, 0);
    std::swap(cur, other);
  }
  int hst = 0;
  CK(cudaMemcpyAsync(&hst, status.p, sizeof(int), cudaMemcpyDeviceToHost, c.s));
  c.sync();
  if (hst != 0) {
    // rank-deficient / ill-conditioned basis: DGEQRF + DORGQR (lapack_wrapper.f90:176-236) never fail, they return an
    // orthonormal basis that contains the span of the input.  Same contract through the solver's SVQB loop
    // (Jacobi on the scaled Gram matrix, deficient directions refilled, repeated until orthonormal).
    std::unique_ptr<dav_solver> sv(new dav_solver(default_device(), 0, 1, nullptr));
    sv->set_dims(m);
    sv->alloc_work(1, std::max(n, 2));
    CK(cudaMemcpy2DAsync(sv->C.p, (size_t)sv->ldv * 8, basis, (size_t)ld * 8, (size_t)m * 8, (size_t)n,
                         cudaMemcpyHostToDevice, sv->stream));
    sv->orthonormalize_block(sv->C.p, n, 0, sv->C.p);
    CK(cudaMemcpy2DAsync(basis, (size_t)ld * 8, sv->C.p, (size_t)sv->ldv * 8, (size_t)m * 8, (size_t)n,
                         cudaMemcpyDeviceToHost, sv->stream));
    CK(cudaStreamSynchronize(sv->stream));
    return DAV_OK;
  }
  CK(cudaMemcpy2DAsync(basis, (size_t)ld * 8, cur, (size_t)m * 8, (size_t)m * 8, (size_t)n, cudaMemcpyDeviceToHost,
                       c.s));
  c.sync();
  API_END
}

// DSYSV('U') (lapack_wrapper.f90:238-277): arr x = brr for a symmetric matrix of which only the upper triangle is
// read; brr <- x.  Entirely on the device: symmetrise, blocked LU with partial pivoting (csrc/densesolve.cu), back
// substitution.  (r01 went through an eigendecomposition and host loops, n <= 1024.)
int dav_lapack_solver(int n, const double* arr, double* brr) {
  API_BEGIN
  need(n >= 1 && arr && brr, "bad arguments");
  Ctx c;
  DevBuf<double> A;
  DevBuf<int> piv, status;
  A.alloc((size_t)n * (n + 1));
  piv.alloc(n);
  status.alloc(1);
  CK(cudaMemsetAsync(status.p, 0, sizeof(int), c.s));
  h2d(A.p, arr, (size_t)n * n, c.s);
  h2d(A.p + (size_t)n * n, brr, n, c.s);
  symmetrize_from_upper(c.s, n, A.p, n);
  lu_solve(c.s, n, 1, A.p, n, piv.p, status.p);
  int hst = 0;
  CK(cudaMemcpyAsync(&hst, status.p, sizeof(int), cudaMemcpyDeviceToHost, c.s));
  d2h(brr, A.p + (size_t)n * n, n, c.s);
  c.sync();
  if (hst != 0) DAV_THROW(DAV_ERR_NOT_POSDEF, "lapack_solver: singular matrix (or NaN input)");
  API_END
}

int dav_lapack_matmul(char transA, char transB, int64_t rows_a, int64_t cols_a, const double* arr, int64_t rows_b,
                      int64_t cols_b, const double* brr, double alpha, double* mtx) {
  API_BEGIN
  need(arr && brr && mtx && rows_a >= 1 && cols_a >= 1 && rows_b >= 1 && cols_b >= 1, "bad arguments");
  need((transA == 'N' || transA == 'T') && (transB == 'N' || transB == 'T'), "trans must be 'N' or 'T'");
  const int64_t m = transA == 'T' ? cols_a : rows_a, k = transA == 'T' ? rows_a : cols_a;
  const int64_t kb = transB == 'T' ? cols_b : rows_b, n = transB == 'T' ? rows_b : cols_b;
  need(k == kb, "inner dimensions differ");
  Ctx c;
  DevBuf<double> A, B, Cm, ws;
  A.alloc((size_t)rows_a * cols_a); B.alloc((size_t)k * n); Cm.alloc((size_t)m * n);
  ws.alloc(std::max<size_t>((size_t)m * n * 16, (size_t)1 << 20));
  h2d(A.p, arr, (size_t)rows_a * cols_a, c.s);
  if (transB == 'T') {  // op(B) = B^T: the device GEMM takes B as stored K x N
    DevBuf<double> Bt;
    Bt.alloc((size_t)rows_b * cols_b);
    h2d(Bt.p, brr, (size_t)rows_b * cols_b, c.s);
    transpose(c.s, rows_b, cols_b, Bt.p, rows_b, B.p, k);
    c.sync();  // Bt goes out of scope
  } else {
    h2d(B.p, brr, (size_t)k * n, c.s);
  }
  gemm(c.s, transA == 'T', m, n, k, alpha, A.p, rows_a, B.p, k, 0.0, Cm.p, m, ws.p, ws.n);
  d2h(mtx, Cm.p, (size_t)m * n, c.s);
  c.sync();
  API_END
}

int dav_lapack_matrix_vector(char transA, int64_t m, int64_t n, const double* mtx, const double* vector, double alpha,
                             double* rs) {
  API_BEGIN
  need(mtx && vector && rs && m >= 1 && n >= 1, "bad arguments");
  need(transA == 'N' || transA == 'T', "trans must be 'N' or 'T'");
  const int64_t mo = transA == 'T' ? n : m, k = transA == 'T' ? m : n;
  Ctx c;
  DevBuf<double> A, x, y, ws;
  A.alloc((size_t)m * n); x.alloc(k); y.alloc(mo);
  ws.alloc(std::max<size_t>((size_t)mo * 64, (size_t)1 << 20));
  h2d(A.p, mtx, (size_t)m * n, c.s);
  h2d(x.p, vector, k, c.s);
  gemm(c.s, transA == 'T', mo, 1, k, alpha, A.p, m, x.p, k, 0.0, y.p, mo, ws.p, ws.n);
  d2h(rs, y.p, mo, c.s);
  c.sync();
  API_END
}

int dav_lapack_sort(char id, int64_t n, double* vector, int32_t* keys) {
  API_BEGIN
  need(vector && keys && n >= 1 && (id == 'I' || id == 'D'), "bad arguments");
  Ctx c;
  // DLASRT + the permutation (lapack_wrapper.f90:367-392): one bitonic sort of (value, position) pairs on the device;
  // keys[original position] = 1-based rank, ties in the original order
  const int64_t np = sort_pairs_padded(n);
  DevBuf<double> d, key;
  DevBuf<int64_t> idx;
  DevBuf<int> status;
  d.alloc(n); key.alloc(np); idx.alloc(np); status.alloc(1);
  CK(cudaMemsetAsync(status.p, 0, sizeof(int), c.s));
  h2d(d.p, vector, n, c.s);
  sort_pairs(c.s, n, d.p, id == 'D', key.p, idx.p, status.p);
  std::vector<double> hk(n);
  std::vector<int64_t> hi(n);
  d2h(hk.data(), key.p, n, c.s);
  CK(cudaMemcpyAsync(hi.data(), idx.p, (size_t)n * 8, cudaMemcpyDeviceToHost, c.s));
  check_status_dev(c, status.p, "lapack_sort");
  for (int64_t t = 0; t < n; ++t) {
    vector[t] = (id == 'D') ? -hk[t] : hk[t];
    keys[hi[t]] = (int32_t)(t + 1);
  }
  API_END
}

int dav_free_matmul(int op, int64_t n, int64_t b, const double* array, double* out) {
  API_BEGIN
  need(array && out && n >= 1 && b >= 1, "bad arguments");
  need(op >= DAV_OP_BENCHMARK_MTX && op <= DAV_OP_TEST_STX, "unknown built-in operator");
  std::unique_ptr<dav_solver> s(new dav_solver(default_device(), 0, 1, nullptr));
  s->set_operator(0, n, op);
  DevBuf<double> X, W;
  X.alloc((size_t)n * b); W.alloc((size_t)n * b);
  h2d(X.p, array, (size_t)n * b, s->stream);
  s->apply_full(0, X.p, n, (int)b, W.p, n);
  d2h(out, W.p, (size_t)n * b, s->stream);
  CK(cudaStreamSynchronize(s->stream));
  API_END
}

int dav_compute_on_the_fly(int op, int64_t i, int64_t dim, double* vector) {
  API_BEGIN
  need(vector && dim >= 1 && i >= 1 && i <= dim, "bad arguments");
  need(op >= DAV_OP_BENCHMARK_MTX && op <= DAV_OP_TEST_STX, "unknown built-in operator");
  std::unique_ptr<dav_solver> s(new dav_solver(default_device(), 0, 1, nullptr));
  s->set_operator(0, dim, op);
  s->ensure_etab();
  DevBuf<double> col;
  col.alloc(dim);
  free_column_builtin(s->stream, op, dim, i - 1, s->etab.p, col.p);
  d2h(vector, col.p, dim, s->stream);
  CK(cudaStreamSynchronize(s->stream));
  API_END
}

}  // extern "C"
