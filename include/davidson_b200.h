/*
 * davidson_b200.h -- C ABI of the B200-native Davidson eigensolver.
 *
 * This is the drop-in boundary: the entry points below are exactly what a Fortran
 * `iso_c_binding` shim (fortran/davidson.f90 in this repo), a ctypes stub, or any other FFI for
 * the reference's `davidson` / `array_utils` / `lapack_wrapper` modules would bind.  Citations
 * are into NLESC-JCER/Fortran_Davidson (`src/...`).
 *
 * Conventions
 *   - every matrix is column-major (Fortran order) IEEE binary64; `ld*` are leading dimensions
 *   - every function returns 0 on success or a DAV_ERR_* code; dav_last_error() gives the text
 *     (the reference prints + `error stop`s, lapack_wrapper.f90:395-408 -- the shim does that
 *     with this text)
 *   - host pointers are never retained after a call returns; a dav_solver_t owns device memory only
 *   - there is NO CPU fallback: every entry point that computes fails with DAV_ERR_CUDA when no
 *     sm_100 device is usable
 */
#ifndef DAVIDSON_B200_H
#define DAVIDSON_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default) /* the library is built with -fvisibility=hidden */
#endif

#define DAV_VERSION 100

/* error codes */
enum {
  DAV_OK = 0,
  DAV_ERR_INVALID = 1,        /* bad argument (incl. unknown method; the reference leaves the
                                 correction undefined there, davidson.f90:656-669) */
  DAV_ERR_CUDA = 2,           /* CUDA runtime / no device */
  DAV_ERR_NOT_POSDEF = 3,     /* projected second_matrix not positive definite (DSYGV info > n) */
  DAV_ERR_NO_CONVERGENCE = 4, /* small dense solver did not converge (DSYEV info > 0) */
  DAV_ERR_BASIS_TOO_LARGE = 5,/* expansion would exceed n columns (DORGQR info < 0 in the reference) */
  DAV_ERR_COMM = 6,           /* NCCL */
  DAV_ERR_STATE = 7           /* call order (e.g. solve before a matrix was set) */
};

/* correction methods (davidson.f90:61-64, :656-669) */
enum { DAV_METHOD_DPR = 0, DAV_METHOD_GJD = 1 };

/* built-in on-the-fly operators for the matrix-free path */
enum {
  DAV_OP_BENCHMARK_MTX = 0, /* benchmark_free.f90:38-63  compute_matrix_on_the_fly */
  DAV_OP_IDENTITY = 1,      /* benchmark_free.f90:65-76  compute_stx_on_the_fly (identity) */
  DAV_OP_TEST_MTX = 2,      /* tests/test_utils.f90:37-51 (same entries as 0) */
  DAV_OP_TEST_STX = 3       /* tests/test_utils.f90:54-68,97-116 (sin variant, unit diagonal) */
};

/* block matvec implementations (diagnostics / parity tests) */
enum { DAV_MATVEC_AUTO = 0, DAV_MATVEC_SIMT = 1, DAV_MATVEC_TMA_DMMA = 2 };

const char* dav_last_error(void);
int dav_version(void);
/* number of usable CUDA devices (0 without a GPU; never fails) */
int dav_device_count(void);
/* GPU used by every call that takes no handle (the drop-in solver calls of section 1 and the array_utils /
 * lapack_wrapper mirrors of section 3): this call, else the environment variable DAV_DEVICE, else device 0.
 * device = -1 returns to the environment / default.  (The reference has no notion of a device: davidson.f90:51-83.) */
int dav_set_default_device(int device);

/* =============================================================================================
 * 1. Drop-in solver calls (host pointers in, host pointers out).
 * ============================================================================================= */

/* generalized_eigensolver_dense (davidson.f90:51-246), also `eigensolver` of README.md:20 when
 * second_matrix == NULL.
 *   matrix, second_matrix : n x n, column-major, intent(in); second_matrix may be NULL (optional)
 *   method                : "DPR" or "GJD" (NUL terminated)
 *   max_dim_sub           : <= 0 means "not present" (default 10*lowest, davidson.f90:115-119)
 *   eigenvalues[lowest], eigenvectors[ldv x lowest] : outputs (first `lowest` Ritz pairs of the
 *                           last Rayleigh-Ritz step, davidson.f90:186-187)
 *   iters                 : iteration index at convergence, or max_iterations+1 plus the warning
 *                           " Warning: Algorithm did not converge!!" (davidson.f90:232-235)       */
int dav_generalized_eigensolver_dense(int64_t n, const double* matrix, int64_t lda, const double* second_matrix,
                                      int64_t ldb, int lowest, const char* method, int max_iterations,
                                      double tolerance, int max_dim_sub, double* eigenvalues, double* eigenvectors,
                                      int64_t ldv, int* iters);
/* dav_generalized_eigensolver_dense keeps ONE handle per process between calls (device block of the matrix, TMA plan,
 * workspace, page-locked staging): a second call with a matrix of the same size only uploads and solves.  This frees
 * it (the next call allocates again).  Environment DAV_DROPIN_CACHE=0: one handle per call, nothing kept. */
int dav_release_cache(void);

/* Host callback of the matrix-free path: Y(n x b) = Op * X(n x b), both column-major with leading
 * dimension n (the shape of `fun_matrix_gemv`, davidson.f90:317-325). */
typedef void (*dav_gemv_fn)(const double* x, double* y, int64_t n, int64_t b, void* ctx);

/* generalized_eigensolver_free (davidson.f90:277-460) with host callbacks.  Always generalized
 * (fun_second_matrix_gemv is not optional, :327-335), `method` accepted and ignored (:428),
 * convergence non-sticky (:416), *iters left untouched when not converged (:417).
 * diag_matrix / diag_second_matrix may be NULL: they are then extracted by n operator
 * applications like extract_diagonal_free (davidson.f90:490-523). */
int dav_generalized_eigensolver_free(int64_t n, dav_gemv_fn fun_matrix_gemv, void* ctx_matrix,
                                     dav_gemv_fn fun_second_matrix_gemv, void* ctx_second,
                                     const double* diag_matrix, const double* diag_second_matrix, int lowest,
                                     const char* method, int max_iterations, double tolerance, int max_dim_sub,
                                     double* eigenvalues, double* ritz_vectors, int64_t ldv, int* iters);

/* Same, with built-in device generators (DAV_OP_*) instead of callbacks: the operator entries are
 * generated on the fly in registers, nothing n x n is ever stored. */
int dav_generalized_eigensolver_free_builtin(int64_t n, int op_matrix, int op_second_matrix, int lowest,
                                             const char* method, int max_iterations, double tolerance,
                                             int max_dim_sub, double* eigenvalues, double* ritz_vectors,
                                             int64_t ldv, int* iters);

/* =============================================================================================
 * 2. Device-resident handle API (the timed path): matrices live in HBM, row-block sharded over
 *    the ranks of one node; only eigenpairs, flags and norms come back.
 * ============================================================================================= */
typedef struct dav_solver dav_solver_t;

/* NCCL bootstrap: rank 0 fills a 128-byte id, the caller broadcasts it (e.g. torch.distributed). */
int dav_get_unique_id(void* id128);
int dav_create(dav_solver_t** h, int device);
int dav_create_distributed(dav_solver_t** h, int device, int rank, int world_size, const void* id128);
int dav_destroy(dav_solver_t* h);

/* page-locked host memory for result arrays: dav_solve / the drop-in calls write eigenvectors into such a block by
 * DMA directly (a pageable destination is staged and copied on the host instead) */
int dav_alloc_pinned(size_t bytes, void** ptr);
int dav_free_pinned(void* ptr);

/* rows [row_begin, row_end) of an n-row matrix owned by `rank` of `world_size` (contiguous blocks,
 * multiples of 128 rows except the last) -- pure host arithmetic, usable without a GPU. */
int dav_partition_rows(int64_t n, int world_size, int rank, int64_t* row_begin, int64_t* row_end);

/* which: 0 = matrix, 1 = second_matrix */
/* generate_diagonal_dominant (array_utils.f90:86-113) on device with the counter-based stream
 * shared with the oracle: a(i,j)=a(j,i)=sparsity*U(seed,min,max), a(i,i)=diag_val or i (1-based). */
int dav_matrix_generate_diagonal_dominant(dav_solver_t* h, int which, int64_t n, double sparsity, int has_diag_val,
                                          double diag_val, uint64_t seed);
/* upload a host matrix (each rank copies its own row block) */
int dav_matrix_upload(dav_solver_t* h, int which, int64_t n, const double* host_matrix, int64_t ld);
/* upload only this rank's row block: host_rows is (row_end-row_begin) x n column-major with leading dimension
 * ld >= row_end-row_begin (dav_partition_rows); a rank then never needs the whole matrix in host memory */
int dav_matrix_upload_rows(dav_solver_t* h, int which, int64_t n, const double* host_rows, int64_t ld);
/* matrix-free: built-in generator or host callback (diag may be NULL) */
int dav_matrix_set_operator(dav_solver_t* h, int which, int64_t n, int op);
int dav_matrix_set_callback(dav_solver_t* h, int which, int64_t n, dav_gemv_fn fn, void* ctx, const double* diag);
/* matrix-free with a DEVICE functor: the generalisation of the reference's fun(i, dim) / fun_matrix_gemv operators
 * (davidson.f90:317-325, :526-569) that keeps the block on the GPU.  The library calls
 *     fn(d_x, ldx, d_y, ldy, n, b, row_begin, nrows, cuda_stream, ctx)
 * with DEVICE pointers: d_x = the complete block X (n x b, column-major, leading dimension ldx), d_y = where the rows
 * [row_begin, row_begin + nrows) of Op * X go (nrows x b, leading dimension ldy; nrows = this rank's share, = n on
 * one GPU).  fn must only ENQUEUE work on `cuda_stream` (a cudaStream_t) and return; no host copy, no
 * synchronisation, any rank count.  diag: the operator's diagonal (host pointer, n entries) or NULL, in which case it
 * is extracted by applying the functor to unit vectors, 64 at a time, like extract_diagonal_free
 * (davidson.f90:490-523). */
typedef void (*dav_device_gemv_fn)(const double* d_x, int64_t ldx, double* d_y, int64_t ldy, int64_t n, int64_t b,
                                   int64_t row_begin, int64_t nrows, void* cuda_stream, void* ctx);
int dav_matrix_set_device_callback(dav_solver_t* h, int which, int64_t n, dav_device_gemv_fn fn, void* ctx,
                                   const double* diag);
int dav_matrix_clear(dav_solver_t* h, int which);
/* copy the local row block back (row_end-row_begin rows x n columns, ld >= rows) -- parity tests */
int dav_matrix_download(dav_solver_t* h, int which, double* host_rows, int64_t ld);

/* solve with the matrices currently set.  Dense matrices -> dense semantics (sticky convergence,
 * iters = max_iterations+1 when not converged); operators/callbacks -> free semantics.
 * eigenvectors: full n x lowest on every rank (may be NULL to skip the device->host copy). */
int dav_solve(dav_solver_t* h, int lowest, int method, int max_iterations, double tolerance, int max_dim_sub,
              double* eigenvalues, double* eigenvectors, int64_t ldv, int* iters);
/* row-sharded result: like dav_solve, but every rank receives only ITS rows of the Ritz vectors
 * (rows dav_partition_rows(n, world, rank) x lowest, leading dimension ldv_local) -- no all-gather of the result.
 * On a single-rank handle it is identical to dav_solve. */
int dav_solve_local(dav_solver_t* h, int lowest, int method, int max_iterations, double tolerance, int max_dim_sub,
                    double* eigenvalues, double* eigenvectors_local, int64_t ldv_local, int* iters);

typedef struct {
  double solve_ms;          /* CUDA-event time of the last dav_solve (device work of the whole loop) */
  double matvec_ms;         /* sum of block-matvec kernel time inside it (events on the solver stream) */
  double matvec_bytes;      /* algorithmic bytes of those launches: 8*nl*n + 8*n*b + 8*nl*b each */
  double matvec_flops;      /* 2*nl*n*b each */
  int matvec_launches;      /* block-matvec launches (A and B count separately) */
  int kernel_launches;      /* every kernel this library launched inside the solve */
  int iterations;           /* outer iterations executed */
  int trace_len;            /* entries valid in trace_* */
  int trace_k[64];          /* basis width at each Rayleigh-Ritz step */
  double trace_err[64];     /* largest residual norm among the `lowest` pairs at that step */
  int last_matvec_b;        /* width of the last block */
  double last_matvec_ms;    /* and its kernel time */
  double rr_ms, orth_ms, resid_ms, proj_ms, init_ms; /* phase times (events) */
  int gjd_inner_iterations; /* block-MINRES iterations of the GJD correction (each = one block matvec per matrix) */
  double gather_ms;         /* all-gather of the new basis block over the ranks (+ staging kernels) */
  double output_ms;         /* Ritz vectors of the result: V*y, gather, copy to the host */
  double comm_ms;           /* time inside NCCL collectives (nested in the phases above; includes waiting for peers) */
  int collectives;          /* inter-GPU exchanges of the solve (own peer-memory kernels or NCCL calls) */
  int spans_dropped;        /* phase spans not timed because the event pool was exhausted (0: the *_ms are complete) */
  int peer_transport;       /* 1: the exchanges were the library's peer-memory kernels, 0: NCCL / single GPU */
  int pip_fallbacks;        /* expansion blocks whose fast orthonormalisation was rejected and redone by the SVQB loop */
} dav_stats_t;
int dav_get_stats(dav_solver_t* h, dav_stats_t* out);
/* per_phase_spans != 0: the following solves also time every phase with CUDA events (matvec_ms, rr_ms, orth_ms, ...,
 * comm_ms of dav_stats_t; ~80 event records per solve, ~0.1 ms).  Default 0: only solve_ms is timed, the per-phase
 * fields stay 0.  The environment variable DAV_SPANS=1 switches the spans on for every handle. */
int dav_set_profiling(dav_solver_t* h, int per_phase_spans);
/* Bytes the last matrix upload of handle h moved over PCIe (h == NULL: the cached handle of the drop-in call).  On one
 * GPU a host matrix that passes a sampled symmetry check (2^17 random pairs compared exactly) is uploaded as its
 * upper triangle only and mirrored on the device -- DSYEV 'U' semantics, half the bytes; an input that fails the
 * check, several ranks, n < 2048 or DAV_SYMMETRIC_UPLOAD=0 take the full upload. */
int dav_upload_bytes(dav_solver_t* h, double* bytes);

/* knobs: DAV_MATVEC_* implementation of the block matvec */
int dav_set_matvec_impl(dav_solver_t* h, int impl);

/* Block matvec on the resident matrix: W(local rows x b) = M(local rows, :) * X(n x b).
 * Host X in (ldx >= n), host W out (ldw >= local rows).  Parity-test entry. */
int dav_block_matvec(dav_solver_t* h, int which, int64_t b, const double* x, int64_t ldx, double* w, int64_t ldw);
/* Time `reps` launches of the block matvec kernel on resident data with CUDA events on the solver
 * stream (X = deterministic pseudo-random block kept on device).  ms_out[reps]. */
int dav_bench_block_matvec(dav_solver_t* h, int which, int64_t b, int reps, float* ms_out);

/* Measured FP64 tensor-pipe peak of the handle's device: TFLOP/s of back-to-back DMMA.8x8x4 (mma.sync.m8n8k4.f64) from
 * registers on every SM, best of `reps` launches of ~2 ms.  Denominator of the FP64 roofline in bench.py. */
int dav_bench_fp64_pipe(dav_solver_t* h, int reps, double* dmma_tflops);

/* Device self-check of the block matvec on a RECTANGULAR m x k block (the shape of a rank's row block in the sharded
 * solve: m = n / ranks rows, k = n columns): pseudo-random A and X generated on the device, W = A X by the TMA/DMMA
 * kernel (schedule / stage depth as DAV_MATVEC_SCHEDULE / DAV_MATVEC_BK or the defaults select) against the
 * tall-skinny GEMM kernel of the library; returns max |W_matvec - W_gemm| and max |W_gemm|.  Lets one GPU validate the
 * multi-GPU tile shapes. */
int dav_debug_matvec_rect(int device, int64_t m, int64_t k, int b, double* max_abs_diff, double* scale);

/* Test entry of the b x b Cholesky + triangular inverse of the block orthonormalisation (host pointers): g = b x b
 * symmetric positive definite, column-major, upper triangle read; t <- R^-1 with g = R^T R (upper triangular, zeros
 * below); *flag = 1.0 when a pivot was not safely positive (t is then undefined); *ms (may be NULL) = kernel time. */
int dav_debug_chol_inv(int b, const double* g, double* t, double* flag, float* ms);

/* Test entry of the fused small-matrix kernel of the block orthonormalisation (csrc/smalldense.cu pip_small_kernel; host
 * pointers): gall = (kold + b) x b column-major, rows 0..kold = H = V^T C, rows kold.. = C^T C; z <- [-H Tm; Tm] with
 * Tm^T (C^T C - H^T H) Tm = I; mode 0 = first pass (scaled Cholesky), 1 = second pass (series); metrics[4] as in the
 * kernel; *launched = 0 when the shape does not fit the fused kernel (the solver then uses the separate kernels). */
int dav_debug_pip_small(int mode, int kold, int b, const double* gall, double* z, double* metrics, int* launched);

/* Timing entry of the tall-skinny products around the block matvec (csrc/dgemm.cu), operands generated on the
 * device: C (m x n) = op(A) * B with op(A) m x k, B k x n; transA 'T' = the projection shape (k = local rows, split-K).
 * ms_out[reps] = event time of each call; to_partials != 0 (transA 'T' only): the product is left as split-K
 * partials, as the solver's fused reduce consumes it.  *max_err (may be NULL) = max |C - C_simt| against the SIMT
 * kernel of the same library (-1 when not compared). */
int dav_debug_gemm_bench(char transA, int64_t m, int64_t n, int64_t k, int reps, int to_partials, float* ms_out,
                         double* max_err);

/* Self-check and timing of the inter-GPU exchanges of a distributed handle (collective: every rank calls it with the
 * same arguments).  kind 0 = all-reduce on the transport in use (peer-memory kernel when the GPUs map each other,
 * else NCCL), 1 = all-reduce through NCCL, 2 = gather of an n x count block (every rank stores its rows into every
 * peer's copy), 3 = the same into the packed MMA-fragment order of the block matvec.  count = doubles (kinds 0, 1)
 * or columns (2, 3; a matrix must be set: the handle's row partition is used).
 * out2[0] = microseconds per call (CUDA events over `reps` back-to-back calls), out2[1] = max abs error against the
 * analytically known result. */
int dav_debug_collective(dav_solver_t* h, int kind, int64_t count, int reps, double* out2);
/* transport of a distributed handle: *peer_transport = 1 when the exchanges are the library's own peer-memory
 * kernels (0: NCCL); calls issued so far on either transport. */
int dav_comm_info(dav_solver_t* h, int* peer_transport, long long* peer_calls, long long* nccl_calls);

/* Host-only check of the block-matvec work schedule (full waves + stream-K remainder) for an M x K local block and
 * a b-column block on a device with num_sms SMs: runs the very functions the kernel and its fixup pass use and
 * verifies that every (row tile, k step) unit is computed exactly once and every partial tile is summed exactly
 * once.  schedule: 0 = pure stream-K, 1 = waves + stream-K remainder, 2 = waves + aligned split-K remainder,
 * < 0 = what DAV_MATVEC_SCHEDULE / the default selects.
 * info[10] (may be NULL) = {grid, waves, first remainder tile, stream-K quota, tiles, k steps, partial segments,
 * tile rows, split-K pieces per remainder tile (0 = stream-K), k steps per piece}.  Needs no GPU.  Returns DAV_OK or DAV_ERR_INVALID (dav_last_error() names the failed check). */
int dav_debug_matvec_schedule(int64_t m, int64_t k, int b, int num_sms, int schedule, long long* info);

/* =============================================================================================
 * 3. array_utils / lapack_wrapper mirrors on device (host pointers in/out).
 * ============================================================================================= */
/* generate_diagonal_dominant (array_utils.f90:86-113); diag_val may be NULL ("not present") */
int dav_generate_diagonal_dominant(int64_t m, double sparsity, const double* diag_val, uint64_t seed, double* arr,
                                   int64_t ld);
/* generate_preconditioner (array_utils.f90:136-160): n x dim_sub one-hot columns at the dim_sub
 * smallest entries of diag, ascending, ties by index.  diag is NOT modified (the reference sorts
 * it in place, a side effect no caller relies on). */
int dav_generate_preconditioner(int64_t n, const double* diag, int dim_sub, double* precond, int64_t ld);
/* norm (array_utils.f90:46-53) */
int dav_norm(int64_t n, const double* vector, double* result);
/* the same as a value (NaN on any error): lets the Fortran shim keep `norm` PURE like array_utils.f90:46 */
double dav_norm_value(int64_t n, const double* vector);
/* lapack_generalized_eigensolver (lapack_wrapper.f90:14-91): all eigenpairs ascending, upper
 * triangle read; stx may be NULL.  Device Jacobi. */
int dav_lapack_generalized_eigensolver(int dim, const double* mtx, const double* stx, double* eigenvalues,
                                       double* eigenvectors);
/* diagnostic / micro-benchmark of the Rayleigh-Ritz eigensolver on one dim x dim symmetric matrix (upper triangle
 * read): runs it 1 + reps times, ms_out[reps] = CUDA-event time of each timed run (may be NULL with reps = 0),
 * info[8] = {tridiagonal fast path accepted by the guard (1/0; -1 when dim uses Jacobi directly), max|S|,
 * max|Y^T Y - I| before the Newton-Schulz step, max|S y - theta y|, 4 reserved (profiling builds)}. */
int dav_sym_eigh_info(int dim, const double* mtx, double* eigenvalues, double* eigenvectors, double* info, int reps,
                      float* ms_out);
/* lapack_generalized_eigensolver_lowest (lapack_wrapper.f90:93-174) */
int dav_lapack_generalized_eigensolver_lowest(int dim, const double* mtx, const double* stx, int lowest,
                                              double* eigenvalues, double* eigenvectors);
/* lapack_qr (lapack_wrapper.f90:176-236): orthonormal basis of the columns, in place (m >= n) */
int dav_lapack_qr(int64_t m, int n, double* basis, int64_t ld);
/* lapack_solver (lapack_wrapper.f90:238-277): symmetric solve arr * x = brr, x overwrites brr */
int dav_lapack_solver(int n, const double* arr, double* brr);
/* lapack_matmul (lapack_wrapper.f90:279-328): mtx = alpha * op(arr) * op(brr); shapes as stored */
int dav_lapack_matmul(char transA, char transB, int64_t rows_a, int64_t cols_a, const double* arr, int64_t rows_b,
                      int64_t cols_b, const double* brr, double alpha, double* mtx);
/* lapack_matrix_vector (lapack_wrapper.f90:330-364) */
int dav_lapack_matrix_vector(char transA, int64_t m, int64_t n, const double* mtx, const double* vector,
                             double alpha, double* rs);
/* lapack_sort (lapack_wrapper.f90:367-392): sorts vector in place ('I'/'D'), keys[i] = 1-based
 * position of original element i in the sorted vector (ties: stable) */
int dav_lapack_sort(char id, int64_t n, double* vector, int32_t* keys);
/* free_matmul (davidson.f90:526-569) with a built-in generator: out = Op * array */
int dav_free_matmul(int op, int64_t n, int64_t b, const double* array, double* out);
/* column i (1-based) of a built-in operator: compute_matrix_on_the_fly(i, dim) */
int dav_compute_on_the_fly(int op, int64_t i, int64_t dim, double* vector);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* DAVIDSON_B200_H */
