// Launchers of every device kernel of the library (definitions in the .cu files next to this one).
#pragma once
#include "common.cuh"

namespace dav {

// Every launcher bumps this counter (bench.py reports it as `gpu_launches`).
extern thread_local long long g_kernel_launches;

// ---- dgemm.cu : generic SIMT FP64 GEMM (tall-skinny shapes, deterministic split-K) -------------
// C(M x N, ldc) = alpha * op(A) * B(K x N, ldb) + beta * C
//   transA == false : A is M x K (lda)            -- "NN": V*G updates, AV*Y, C*T
//   transA == true  : A is stored K x M (lda)     -- "TN": projections V^T W (K = local rows)
// ws: workspace for split-K partials (TN with long K); ws_doubles its capacity.
// partials_out != nullptr: the product (alpha = 1, beta = 0 implied) is LEFT as *partials_out partial blocks of
// M x N doubles in ws (ws[z*M*N + m + j*M]) and C is not touched: the caller sums them -- Comm::reduce_sum does that
// together with the sum over the ranks and the output layout in one kernel.
// split != nullptr: op(A) is given as two blocks -- TN: rows [0, at) of the result come from A, [at, M) from A2
// (both K x . with their own leading dimension); NN: columns [0, at) of A from A, [at, K) from A2 (at % 4 == 0).
struct GemmSplit {
  const double* A2;
  int64_t lda2;
  int64_t at;
};
void gemm(cudaStream_t s, bool transA, int64_t M, int64_t N, int64_t K, double alpha, const double* A,
          int64_t lda, const double* B, int64_t ldb, double beta, double* C, int64_t ldc, double* ws,
          size_t ws_doubles, int* partials_out = nullptr, const GemmSplit* split = nullptr);

// Residuals + corrections of nc Ritz pairs in one pass (davidson.f90:163-170 via the stored products, :688-696):
//   R(:, j) = AV y_j - theta_j (BV|V) y_j;   C(:, j) = R(:, j) / (theta_j dB - dA)  (write_correction; dB == nullptr
//   -> 1), else C(:, j) = (BV|V) y_j (the GJD solver wants the B-product);   n2out[j] = sum of R(:, j)^2.
// AV, BV: nl x k (ldv); Y: k x nc (ldy).  partial: >= residual_fused_partials(nl, nc) doubles.
size_t residual_fused_partials(int64_t nl, int nc);
void residual_fused(cudaStream_t s, int64_t nl, int nc, int k, const double* AV, const double* BV, int64_t ldv,
                    const double* Y, int64_t ldy, const double* theta, const double* dA, const double* dB,
                    bool write_correction, double* R, int64_t ldr, double* C, int64_t ldc, double* partial,
                    double* n2out);

// ---- matvec_dmma.cu : the hot kernel.  W(M x b) = A(M x K, lda) * X(K x b)  ---------------------
// TMA (2D tensor map, 128B swizzle) -> mbarrier pipeline -> FP64 DMMA, persistent stream-K grid.
struct MatvecPlan;  // opaque: tensor map + workspace for one resident matrix
MatvecPlan* matvec_plan_create(const double* A, int64_t M, int64_t K, int64_t lda, int max_b);
void matvec_plan_destroy(MatvecPlan* p);
bool matvec_dmma_supported();
// X is K x b column-major (ldx); Xp is scratch for the packed copy of X (>= round_up(K,64)*round_up(b,8)).
void matvec_dmma(cudaStream_t s, MatvecPlan* plan, int b, const double* X, int64_t ldx, double* W, int64_t ldw);
// the same with X already in the packed fragment order (layout: comm.cuh, Comm::gather_rows_packed)
int64_t matvec_kpad(int64_t K);                  // padded row count of the packed block
size_t matvec_packed_doubles(int64_t K, int b);  // size of the packed block of b columns
void matvec_dmma_packed(cudaStream_t s, MatvecPlan* plan, int b, const double* Xpacked, double* W, int64_t ldw);
// host model of the (waves + stream-K) schedule the kernel executes; see matvec_dmma.cu.  0 = consistent.
int matvec_schedule_selftest(int64_t M, int64_t K, int b, int num_sms, int schedule, long long* info);

// ---- microbench.cu : measured FP64 tensor-pipe peak (denominator of the roofline in bench.py) ----------------
double dmma_peak_tflops(cudaStream_t s, int reps);

// ---- freeops.cu : on-the-fly operators (benchmark_free.f90:38-76, tests/test_utils.f90:37-116) --
// W(rows row0..row0+nl) = Op * X(n x b); etab[n] = (double)expf(i/n) table (built on host with glibc expf).
void free_matmul_builtin(cudaStream_t s, int op, int64_t n, int64_t row0, int64_t nl, int b, const double* etab,
                         const double* X, int64_t ldx, double* W, int64_t ldw);
void free_diag_builtin(cudaStream_t s, int op, int64_t n, int64_t row0, int64_t nl, const double* etab,
                       double* diag);
void free_column_builtin(cudaStream_t s, int op, int64_t n, int64_t col /*0-based*/, const double* etab,
                         double* out);

// out(:, j) = Op(rows row0..row0+nl, idx[j]) : A*V for a one-hot V without generating the whole operator
void free_gather_columns_builtin(cudaStream_t s, int op, int64_t n, int64_t row0, int64_t nl, const double* etab,
                                 const int64_t* idx, int k, double* out, int64_t ldo);
// ---- freeops_dmma.cu : the same operators on the FP64 tensor pipe (entries from a piecewise polynomial) ----------
struct FreeTables;  // (e_t, 1/e_t) table + polynomial coefficients + packed-X scratch of one operator
FreeTables* free_tables_create(int op, int64_t n, const double* etab_host);
void free_tables_destroy(FreeTables* t);
bool free_tables_usable(const FreeTables* t);   // false: the fitted table failed its accuracy check
double free_tables_max_err(const FreeTables* t);
void free_matmul_dmma(cudaStream_t s, FreeTables* t, int64_t row0, int64_t nl, int b, const double* X, int64_t ldx,
                      double* W, int64_t ldw);

// ---- smalldense.cu : k x k problems on one CTA ----------------------------------------------------
// Two-sided Jacobi, round-robin parallel ordering.  S: k x k (ld k), upper triangle read, destroyed.
// Y: k x k eigenvectors sorted by ascending eigenvalue w.  status: device int, set nonzero on failure
// (1 = no convergence / NaN).  scratch: >= jacobi_scratch_doubles(k) doubles of global memory (rotation log of up
// to 40 sweeps, or the S/V copies of the large-k path).
inline size_t jacobi_scratch_doubles(int k) { return 44 * (size_t)(k + 2) * (size_t)(k + 2); }
// skip (nullable): device flag; when *skip != 0 at kernel start the call is a no-op (see sym_eigh).
void jacobi_eigh(cudaStream_t s, int k, double* S, double* Y, double* w, double* scratch, int* status,
                 const int* skip = nullptr);
// ---- trideig.cu : the Rayleigh-Ritz eigensolver.  Same contract as jacobi_eigh (S: upper triangle read, NOT
// modified).  k >= 16: Householder tridiagonalisation + one warp per eigenpair (multisection, twisted
// factorisation, back-transformation) + a-posteriori guard; Jacobi when the guard rejects or k is small.
// scratch: >= sym_eigh_scratch_doubles(k).
size_t sym_eigh_scratch_doubles(int k);
// lds: leading dimension of S (0 = k); a strided S is accepted for k >= 16 (the fast path copies it anyway)
void sym_eigh(cudaStream_t s, int k, double* S, double* Y, double* w, double* scratch, int* status, int64_t lds = 0);
bool sym_eigh_uses_tridiag(int k);
// guard outputs of the last sym_eigh call on this scratch: double[8] {max|S|, max|G-I|, max residual} + int accept
double* sym_eigh_flags(double* scratch, int k);
// T(:,j) = U(:,j) / sqrt(sv[j]); status |= 2 if some sv[j] <= 0 (not positive definite)
void scale_cols_rsqrt_checked(cudaStream_t s, int k, const double* U, const double* sv, double* T, int* status);
// D = diag(G)^-1/2 (0 where diag <= 0); G <- D G D
void gram_prescale(cudaStream_t s, int k, double* G, double* D);
// SVQB transform: T(i,j) = D[i] * U(i,j) / sqrt(sv[j]) for sv[j] > thr*max(sv), else column flagged
// deficient (flags[j] = 1, T(:,j) = 0).
void svqb_make_T(cudaStream_t s, int k, const double* U, const double* sv, const double* D, double* T,
                 int* flags);
// lower triangle <- upper triangle of the k x k matrix (ld)
void symmetrize_from_upper(cudaStream_t s, int k, double* S, int64_t ld);
// max |G - (minus_identity ? I : 0)| over the rows x cols matrix -> out[0] (single CTA; NaN -> 1e300)
void max_abs_dev(cudaStream_t s, int rows, int cols, const double* G, int64_t ld, bool minus_identity, double* out);
// Cholesky (upper, G = R^T R) in place on one CTA; status |= 2 when not positive definite
void cholesky_upper(cudaStream_t s, int k, double* G, int64_t ld, int* status);
// G (b x b, ld b, only read) = R^T R; T <- R^-1 (dense, upper).  flag[0] (device double) <- 1.0 if a pivot is not
// safely positive, else 0.0.  Blocks too wide for shared memory (b >= ~170) use gwork (>= b*(b+1) doubles of global
// scratch); returns false (nothing launched) only when that is needed and missing.
bool chol_inv_upper(cudaStream_t s, int b, const double* G, double* T, double* flag, double* gwork = nullptr,
                    size_t gwork_doubles = 0);
// fused block orthonormalisation helpers (see solver.cu::orthonormalize_block_pip)
void pip_prepare(cudaStream_t s, int k, int b, const double* Gall, const double* P, double* Gs, double* D,
                 double* metrics);
void pip_finish(cudaStream_t s, int k, int b, const double* Tinv, const double* D, double* Tm, double* M);
// the five calls above (small GEMM H^T H, pip_prepare, chol_inv_upper, pip_finish, small GEMM -H Tm) in ONE kernel:
// Gall ((kold + b) x b) -> Z = [-H Tm; Tm], metrics[0..3] (see smalldense.cu).  mode 0: first pass (scaled Cholesky),
// mode 1: second pass (Tm = (I + E)^-1/2 by its series; metrics[2] raised when |E| >= 1e-5).  Returns false (nothing
// launched) when the block does not fit one CTA's shared memory: the caller then uses the separate kernels.
bool pip_small(cudaStream_t s, int mode, int kold, int b, const double* Gall, double* Z, double* metrics);
// Rinv = inverse of the upper triangular R (k x k)
void invert_upper(cudaStream_t s, int k, const double* R, int64_t ld, double* Rinv);

// ---- densesolve.cu : the remaining lapack_wrapper helpers on the device -----------------------------------------
// Aaug: n x (n + nrhs) column-major (lda), the right-hand sides behind the matrix; on return the nrhs columns hold the
// solutions.  Blocked LU with partial pivoting; piv: n ints of scratch; status |= 2 for a singular matrix / NaN.
void lu_solve(cudaStream_t s, int n, int nrhs, double* Aaug, int64_t lda, int* piv, int* status);
void transpose(cudaStream_t s, int64_t rows, int64_t cols, const double* in, int64_t ldi, double* out, int64_t ldo);
// key[0..n) = in sorted ascending (descending: by -in, i.e. descending in), idx = original positions, ties by position
// (the stable order); key / idx hold sort_pairs_padded(n) entries.  status |= 1 on NaN.
int64_t sort_pairs_padded(int64_t n);
void sort_pairs(cudaStream_t s, int64_t n, const double* in, bool descending, double* key, int64_t* idx, int* status);

// ---- vecops.cu : n x k vector-block kernels and generators --------------------------------------
void fill_zero(cudaStream_t s, double* p, size_t count);
void copy_matrix(cudaStream_t s, int64_t rows, int64_t cols, const double* src, int64_t lds, double* dst,
                 int64_t ldd);
// dst(k x k, ldd)[r0.., c0..] block copies used for the projected matrices
void gen_diag_dominant(cudaStream_t s, double* A, int64_t lda, int64_t nl, int64_t n, int64_t row0, double sparsity,
                       int has_diag, double diag_val, uint64_t seed);
void extract_diag(cudaStream_t s, const double* A, int64_t lda, int64_t nl, int64_t row0, double* out);
// (value,index) of the `k` smallest entries of diag[0..count) (global index = gidx[i] or row0 + i), ascending,
// ties by index (entries with a negative gidx are padding).  Multi-CTA rank counting.  status |= 1 on NaN.
// scratch_val / scratch_idx: >= topk_scratch_entries(count, k) entries each.
inline size_t topk_scratch_entries(int64_t count, int k) {
  return 2 * ((size_t)((count + 1023) / 1024) * (size_t)k + (size_t)k);
}
void topk_smallest(cudaStream_t s, const double* diag, const int64_t* gidx /*nullable*/, int64_t count,
                   int64_t row0, int k, double* out_val, int64_t* out_idx, int* status, double* scratch_val,
                   int64_t* scratch_idx);
// A(i, j) <- A(j, i) for i > j (n x n, lda): completes a matrix of which only the upper triangle was uploaded
// (rows [row_begin, row_end) of the lower triangle only when given: multiples of 32, row_end < 0 = n)
void mirror_upper_to_lower(cudaStream_t s, double* A, int64_t lda, int64_t n, int64_t row_begin = 0,
                           int64_t row_end = -1);
// X (n x w, ld n) = columns c0 .. c0 + w of the identity; diag[i - row0] = Y(i - row0, i - c0) for the owned rows of
// that column range (extract_diagonal_free, davidson.f90:490-523, with the block on the device)
void unit_block(cudaStream_t s, double* X, int64_t n, int64_t c0, int w);
void take_diagonal(cudaStream_t s, const double* Y, int64_t ldy, int64_t nl, int64_t row0, int64_t c0, int w,
                   double* diag);
// V(nl x k) one-hot: V(idx[j]-row0, j) = 1 when idx[j] is a local row
void set_onehot(cudaStream_t s, double* V, int64_t ldv, int64_t nl, int64_t row0, const int64_t* idx, int k);
// out(:, j) = A(:, idx[j])   (A is nl x n local row block)
void gather_columns(cudaStream_t s, const double* A, int64_t lda, int64_t nl, const int64_t* idx, int k,
                    double* out, int64_t ldo);
// out(nl x k) = diag(d) applied to one-hot columns: out(i,j) = (row0+i == idx[j]) ? d[i] : 0  (unused for dense)
// squared column norms, deterministic two-stage; partial: >= k*256 doubles
void col_norms2(cudaStream_t s, int64_t nl, int k, const double* X, int64_t ldx, double* partial, double* out);
// X(:,j) *= 1/sqrt(n2[j]) when n2[j] > 0
void scale_cols_rsqrt(cudaStream_t s, int64_t nl, int k, double* X, int64_t ldx, const double* n2);
// residual + DPR (davidson.f90:163-170 via stored products, :688-696; free :401-410, :484):
//   r = R - theta_j * C ; R <- r ; C <- r / (theta_j * dB_i - dA_i)   (dB == nullptr -> 1)
//   n2part: partial squared norms of r (deterministic two-stage, as col_norms2)
void residual_dpr(cudaStream_t s, int64_t nl, int k, double* R, int64_t ldr, double* C, int64_t ldc,
                  const double* theta, const double* dA, const double* dB, bool write_correction, double* partial,
                  double* n2out);
// pseudo-random refill of flagged columns (rank-deficient directions of an expansion block)
void fill_random_cols(cudaStream_t s, double* X, int64_t ldx, int64_t nl, int64_t row0, const int* flags, int k,
                      uint64_t salt);
// deterministic pseudo-random block in [-1,1) (bench input)
void fill_random(cudaStream_t s, double* X, size_t count, uint64_t salt);
// staged all-gather layout [rank][chunk x b] -> column-major n x b
void unstage_allgather(cudaStream_t s, const double* stage, int world, int64_t chunk, int64_t n, int b, double* X,
                       int64_t ldx);
// column-major nl x b (ldx) -> contiguous chunk x b (zero padded rows)
void stage_block(cudaStream_t s, const double* X, int64_t ldx, int64_t nl, int64_t chunk, int b, double* stage);

}  // namespace dav
