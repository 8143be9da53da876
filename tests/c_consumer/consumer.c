/* A C99 consumer of include/davidson_b200.h, written like INTEGRATION.md section 2 (test infrastructure): proves that
 * the header is plain C (no C++-isms, no torch types), that every call form the document shows compiles against the
 * declared prototypes, and that the shared library links from C.  Run without a GPU it may only make non-computing
 * calls; every computing entry point must then fail loudly with DAV_ERR_CUDA (there is no CPU fallback). */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "davidson_b200.h"

static void host_op(const double* x, double* y, int64_t n, int64_t b, void* ctx) {
  (void)ctx;
  memcpy(y, x, sizeof(double) * (size_t)(n * b));
}

static void device_op(const double* d_x, int64_t ldx, double* d_y, int64_t ldy, int64_t n, int64_t b, int64_t row_begin,
                      int64_t nrows, void* stream, void* ctx) {
  (void)d_x; (void)ldx; (void)d_y; (void)ldy; (void)n; (void)b; (void)row_begin; (void)nrows; (void)stream; (void)ctx;
}

/* the call forms of INTEGRATION.md section 2; compiled always, executed only with a GPU (argv[1] = "run") */
static int integration_calls(int64_t n, const double* A, double* ev, double* vec, int with_callbacks) {
  int iters = 0;
  int rc = dav_generalized_eigensolver_dense(n, A, n, NULL, n, 4, "DPR", 1000, 1e-8, 0, ev, vec, n, &iters);
  if (rc) return rc;
  dav_solver_t* h = NULL;
  rc = dav_create(&h, 0);
  if (rc) return rc;
  rc = dav_matrix_generate_diagonal_dominant(h, 0, n, 1e-4, 0, 0.0, 0);
  if (!rc) rc = dav_set_profiling(h, 1);
  if (!rc) rc = dav_solve(h, 4, DAV_METHOD_DPR, 1000, 1e-8, 0, ev, vec, n, &iters);
  dav_stats_t st;
  if (!rc) rc = dav_get_stats(h, &st);
  if (!rc && with_callbacks) rc = dav_matrix_set_device_callback(h, 0, n, device_op, NULL, NULL);
  if (!rc && with_callbacks) rc = dav_matrix_set_callback(h, 1, n, host_op, NULL, NULL);
  dav_destroy(h);
  return rc;
}

int main(int argc, char** argv) {
  printf("version %d\n", dav_version());
  int64_t b0 = -1, e0 = -1, b1 = -1, e1 = -1;
  if (dav_partition_rows(100000, 8, 0, &b0, &e0) || dav_partition_rows(100000, 8, 7, &b1, &e1)) return 2;
  printf("rows %lld %lld %lld %lld\n", (long long)b0, (long long)e0, (long long)b1, (long long)e1);
  const int64_t n = 64;
  double* A = calloc((size_t)(n * n), sizeof(double));
  double ev[4], *vec = calloc((size_t)(n * 4), sizeof(double));
  for (int64_t i = 0; i < n; ++i) A[i + i * n] = (double)(i + 1);
  if (argc > 1 && strcmp(argv[1], "run") == 0) {
    const int rc = integration_calls(n, A, ev, vec, argc > 2);
    printf("rc %d ev0 %.12f\n", rc, ev[0]);
    return rc ? 1 : 0;
  }
  /* no GPU: the drop-in call must refuse, with a message */
  int iters = -7;
  const int rc = dav_generalized_eigensolver_dense(n, A, n, NULL, n, 4, "DPR", 10, 1e-8, 0, ev, vec, n, &iters);
  printf("rc %d (%s) msg: %s\n", rc, rc == DAV_ERR_CUDA ? "DAV_ERR_CUDA" : "other", dav_last_error());
  free(A);
  free(vec);
  return 0;
}
