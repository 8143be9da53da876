// =====================================================================================
// TEST INFRASTRUCTURE ONLY -- CPU oracle for the Davidson hot path.
//
// A statement-by-statement C++ restatement of the reference's Fortran algorithm
// (NLESC-JCER/Fortran_Davidson: src/davidson.f90, src/array_utils.f90,
// src/lapack_wrapper.f90, src/benchmark_free.f90, src/tests/test_utils.f90), linked
// against the *real* LAPACK/BLAS that ships with scipy (OpenBLAS, LP64, symbols
// scipy_dsyev_ ...).  The reference itself cannot be compiled in this image (there is
// no Fortran compiler), so this file plus real LAPACK is the parity oracle.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// leg may load this library.  Nothing in fortran_davidson_b200/ (the product) may.
//
// Parity pinning (see DESIGN.md "Oracle"): the reference's own tests only pin
//   (i) lowest eigenvalues == scipy.linalg.eigh on the same matrix (test_davidson.py:36-40,67-69)
//  (ii) lapack wrapper eigenpairs == eigh (test_lapack.py:47-51)
// (iii) residual norms < 1e-8 (test_dense_properties.f90:31-39)
// and tests/test_oracle.py checks all three.  Iteration counts and eigenvectors beyond the
// residual are NOT pinned by any reference test ("parity unpinned" there); they are pinned
// only by this restatement.
//
// Every array is column-major (Fortran order), indices in comments are 1-based like the
// reference, C loops are 0-based.
// =====================================================================================
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

typedef int lapack_int;  // LP64 build of scipy's OpenBLAS

extern "C" {
// Fortran calling convention: everything by reference, hidden string lengths appended.
void scipy_dsyev_(const char* jobz, const char* uplo, const lapack_int* n, double* a, const lapack_int* lda,
                  double* w, double* work, const lapack_int* lwork, lapack_int* info, size_t, size_t);
void scipy_dsygv_(const lapack_int* itype, const char* jobz, const char* uplo, const lapack_int* n, double* a,
                  const lapack_int* lda, double* b, const lapack_int* ldb, double* w, double* work,
                  const lapack_int* lwork, lapack_int* info, size_t, size_t);
void scipy_dsygvx_(const lapack_int* itype, const char* jobz, const char* range, const char* uplo,
                   const lapack_int* n, double* a, const lapack_int* lda, double* b, const lapack_int* ldb,
                   const double* vl, const double* vu, const lapack_int* il, const lapack_int* iu,
                   const double* abstol, lapack_int* m, double* w, double* z, const lapack_int* ldz, double* work,
                   const lapack_int* lwork, lapack_int* iwork, lapack_int* ifail, lapack_int* info, size_t, size_t,
                   size_t);
void scipy_dgeqrf_(const lapack_int* m, const lapack_int* n, double* a, const lapack_int* lda, double* tau,
                   double* work, const lapack_int* lwork, lapack_int* info);
void scipy_dorgqr_(const lapack_int* m, const lapack_int* n, const lapack_int* k, double* a,
                   const lapack_int* lda, const double* tau, double* work, const lapack_int* lwork,
                   lapack_int* info);
void scipy_dsysv_(const char* uplo, const lapack_int* n, const lapack_int* nrhs, double* a, const lapack_int* lda,
                  lapack_int* ipiv, double* b, const lapack_int* ldb, double* work, const lapack_int* lwork,
                  lapack_int* info, size_t);
void scipy_dgemm_(const char* ta, const char* tb, const lapack_int* m, const lapack_int* n, const lapack_int* k,
                  const double* alpha, const double* a, const lapack_int* lda, const double* b,
                  const lapack_int* ldb, const double* beta, double* c, const lapack_int* ldc, size_t, size_t);
void scipy_dgemv_(const char* ta, const lapack_int* m, const lapack_int* n, const double* alpha, const double* a,
                  const lapack_int* lda, const double* x, const lapack_int* incx, const double* beta, double* y,
                  const lapack_int* incy, size_t);
void scipy_dlasrt_(const char* id, const lapack_int* n, double* d, lapack_int* info, size_t);
void scipy_openblas_set_num_threads(int);
int scipy_openblas_get_num_threads(void);
}

namespace {

typedef std::vector<double> dvec;

// Error convention of the reference: check_lapack_call prints and `error stop`s
// (lapack_wrapper.f90:395-408).  The oracle returns the LAPACK info instead so that tests can
// look at it; the message text is kept.
int check_lapack_call(lapack_int info, const char* name) {
  if (info != 0) {
    std::fprintf(stderr, " call to subroutine: %s has failed!\n info: %d\n", name, (int)info);
    return (int)info;
  }
  return 0;
}

inline size_t idx(size_t i, size_t j, size_t ld) { return i + j * ld; }

}  // namespace

extern "C" {

void orc_set_num_threads(int n) { scipy_openblas_set_num_threads(n); }
int orc_get_num_threads(void) { return scipy_openblas_get_num_threads(); }

// -------------------------------------------------------------------------------------
// lapack_wrapper.f90
// -------------------------------------------------------------------------------------

// lapack_generalized_eigensolver (lapack_wrapper.f90:14-91): DSYEV('V','U') or
// DSYGV(itype=1,'V','U') on local copies; all eigenpairs, ascending.
int orc_lapack_generalized_eigensolver(int dim, const double* mtx, const double* stx /* may be NULL */,
                                       double* eigenvalues, double* eigenvectors) {
  const bool gev = stx != nullptr;
  const lapack_int n = dim, itype = 1;
  dvec mtx_copy(mtx, mtx + (size_t)dim * dim);  // :47-48
  dvec stx_copy;
  if (gev) stx_copy.assign(stx, stx + (size_t)dim * dim);  // :50-53
  dvec w(std::max(dim, 1));
  lapack_int info = 0, lwork = -1;
  double wq = 0.0;
  if (gev) {  // workspace query, :58-64
    scipy_dsygv_(&itype, "V", "U", &n, mtx_copy.data(), &n, stx_copy.data(), &n, w.data(), &wq, &lwork, &info, 1,
                 1);
    if (int e = check_lapack_call(info, "DSYGV")) return e;
  } else {
    scipy_dsyev_("V", "U", &n, mtx_copy.data(), &n, w.data(), &wq, &lwork, &info, 1, 1);
    if (int e = check_lapack_call(info, "DSYEV")) return e;
  }
  lwork = std::max(1, (int)wq);  // :67
  dvec work(lwork);
  if (gev) {  // :72-78
    scipy_dsygv_(&itype, "V", "U", &n, mtx_copy.data(), &n, stx_copy.data(), &n, w.data(), work.data(), &lwork,
                 &info, 1, 1);
    if (int e = check_lapack_call(info, "DSYGV")) return e;
  } else {
    scipy_dsyev_("V", "U", &n, mtx_copy.data(), &n, w.data(), work.data(), &lwork, &info, 1, 1);
    if (int e = check_lapack_call(info, "DSYEV")) return e;
  }
  std::copy(w.begin(), w.begin() + dim, eigenvalues);             // :81
  std::copy(mtx_copy.begin(), mtx_copy.end(), eigenvectors);      // :82
  return 0;
}

// lapack_generalized_eigensolver_lowest (lapack_wrapper.f90:93-174): DSYGVX range 'I' 1..lowest.
// Never called by the solver; restated for the wrapper parity tests.  The reference leaves
// `abstol` uninitialised (:117,146); the oracle uses 0 (= LAPACK default tolerance).
int orc_lapack_generalized_eigensolver_lowest(int dim, const double* mtx, const double* stx, int lowest,
                                              double* eigenvalues, double* eigenvectors) {
  const lapack_int n = dim, itype = 1, il = 1, iu = lowest;
  dvec mtx_copy(mtx, mtx + (size_t)dim * dim), stx_copy(stx, stx + (size_t)dim * dim);
  const double vl = 0.0, vu = 0.0, abstol = 0.0;
  lapack_int m = 0, info = 0, lwork = -1, iwq = 0;
  std::vector<lapack_int> ifail(dim);
  dvec w(dim), z((size_t)dim * lowest);
  double wq = 0.0;
  scipy_dsygvx_(&itype, "V", "I", "U", &n, mtx_copy.data(), &n, stx_copy.data(), &n, &vl, &vu, &il, &iu, &abstol,
                &m, w.data(), z.data(), &n, &wq, &lwork, &iwq, ifail.data(), &info, 1, 1, 1);
  if (int e = check_lapack_call(info, "DSYGVX")) return e;
  lwork = std::max(1, (int)wq);
  dvec work(lwork);
  std::vector<lapack_int> iwork(std::max((int)lwork, 5 * dim));
  scipy_dsygvx_(&itype, "V", "I", "U", &n, mtx_copy.data(), &n, stx_copy.data(), &n, &vl, &vu, &il, &iu, &abstol,
                &m, w.data(), z.data(), &n, work.data(), &lwork, iwork.data(), ifail.data(), &info, 1, 1, 1);
  if (int e = check_lapack_call(info, "DSYGVX")) return e;
  std::copy(w.begin(), w.begin() + lowest, eigenvalues);
  std::copy(z.begin(), z.end(), eigenvectors);
  return 0;
}

// lapack_qr (lapack_wrapper.f90:176-236): DGEQRF then DORGQR(m, n, min(m,n)); in place.
int orc_lapack_qr(int m_, int n_, double* basis) {
  const lapack_int m = m_, n = n_, k = std::min(m_, n_);
  dvec tau(std::max(n_, 1));
  lapack_int info = 0, lwork = -1;
  double wq = 0.0;
  scipy_dgeqrf_(&m, &n, basis, &m, tau.data(), &wq, &lwork, &info);  // :205
  if (int e = check_lapack_call(info, "DGEQRF")) return e;
  lwork = std::max(1, (int)wq);
  dvec work(lwork);
  scipy_dgeqrf_(&m, &n, basis, &m, tau.data(), work.data(), &lwork, &info);  // :214
  if (int e = check_lapack_call(info, "DGEQRF")) return e;
  lwork = -1;
  scipy_dorgqr_(&m, &n, &k, basis, &m, tau.data(), &wq, &lwork, &info);  // :221
  if (int e = check_lapack_call(info, "DORGQR")) return e;
  lwork = std::max(1, (int)wq);
  work.assign(lwork, 0.0);
  scipy_dorgqr_(&m, &n, &k, basis, &m, tau.data(), work.data(), &lwork, &info);  // :230
  if (int e = check_lapack_call(info, "DORGQR")) return e;
  return 0;
}

// lapack_solver (lapack_wrapper.f90:238-277): DSYSV('U', n, 1) in place; on info>0 the reference
// pokes tiny() into arr(info,info) of the *already factorised* array and calls DSYSV again
// (:269-273).  Restated literally.
int orc_lapack_solver(int n_, double* arr, double* brr) {
  const lapack_int n = n_, nrhs = 1;
  std::vector<lapack_int> ipiv(std::max(n_, 1));
  lapack_int info = 0, lwork = -1;
  double wq = 0.0;
  scipy_dsysv_("U", &n, &nrhs, arr, &n, ipiv.data(), brr, &n, &wq, &lwork, &info, 1);
  if (int e = check_lapack_call(info, "DSYSV")) return e;
  lwork = std::max(1, (int)wq);
  dvec work(lwork);
  scipy_dsysv_("U", &n, &nrhs, arr, &n, ipiv.data(), brr, &n, work.data(), &lwork, &info, 1);
  if (info > 0) {
    arr[idx(info - 1, info - 1, n_)] = 2.2250738585072014e-308;  // tiny(1d0)
    scipy_dsysv_("U", &n, &nrhs, arr, &n, ipiv.data(), brr, &n, work.data(), &lwork, &info, 1);
    if (int e = check_lapack_call(info, "DSYSV")) return e;
  }
  return info < 0 ? (int)info : 0;
}

// lapack_matmul (lapack_wrapper.f90:279-328): mtx = alpha * op(arr) * op(brr) via DGEMM.
// (rows_a, cols_a), (rows_b, cols_b) are the stored shapes; out is m x n, returned through *m_out,*n_out.
void orc_lapack_matmul(char transA, char transB, int rows_a, int cols_a, const double* arr, int rows_b,
                       int cols_b, const double* brr, double alpha, double* mtx) {
  lapack_int m, n, k, lda, ldb;
  if (transA == 'T') { k = rows_a; m = cols_a; lda = k; } else { k = cols_a; m = rows_a; lda = m; }
  if (transB == 'T') { n = rows_b; ldb = n; } else { n = cols_b; ldb = k; }
  const double zero = 0.0;
  std::fill(mtx, mtx + (size_t)m * n, 0.0);  // :324
  scipy_dgemm_(&transA, &transB, &m, &n, &k, &alpha, arr, &lda, brr, &ldb, &zero, mtx, &m, 1, 1);
  (void)cols_b;
}

// lapack_matrix_vector (lapack_wrapper.f90:330-364): rs = alpha * op(mtx) * vector via DGEMV.
void orc_lapack_matrix_vector(char transA, int m_, int n_, const double* mtx, const double* vector, double alpha,
                              double* rs) {
  const lapack_int m = m_, n = n_, one = 1;
  const double zero = 0.0;
  std::fill(rs, rs + m_, 0.0);  // the reference allocates rs(m) whatever transA is (:359)
  scipy_dgemv_(&transA, &m, &n, &alpha, mtx, &m, vector, &one, &zero, rs, &one, 1);
}

// lapack_sort (lapack_wrapper.f90:367-392): sorts `vector` IN PLACE with DLASRT and returns
// keys(i) = position j of the original element i in the sorted vector, found by an O(n^2) exact
// match search `abs(vector(j) - xs(i)) < 1e-16` (a single precision literal) with NO early
// exit, so for duplicated values the last matching j wins.  keys is 1-based like the reference.
int orc_lapack_sort(char id, int n_, double* vector, int* keys) {
  const lapack_int n = n_;
  dvec xs(vector, vector + n_);
  lapack_int info = 0;
  scipy_dlasrt_(&id, &n, vector, &info, 1);
  if (int e = check_lapack_call(info, "DLASRT")) return e;
  const double eps = (double)1e-16f;
  for (int i = 0; i < n_; ++i) {
    keys[i] = 0;  // the reference leaves unmatched keys undefined; 0 = "no key"
    for (int j = 0; j < n_; ++j) {
      if (std::fabs(vector[j] - xs[i]) < eps) keys[i] = j + 1;
    }
  }
  return 0;
}

// -------------------------------------------------------------------------------------
// array_utils.f90
// -------------------------------------------------------------------------------------

// norm (array_utils.f90:46-53): sqrt(sum(vector**2)).
double orc_norm(int n, const double* v) {
  double s = 0.0;
  for (int i = 0; i < n; ++i) s += v[i] * v[i];
  return std::sqrt(s);
}

// Counter-based uniform generator shared bit-for-bit with the CUDA library
// (fortran_davidson_b200/csrc/generators.cuh): the reference draws from the compiler PRNG
// (`call random_number(arr)`, array_utils.f90:96) with no seed anywhere in the tree, so any
// fixed stream is an equally valid instance; this one is keyed on (seed, min(i,j), max(i,j))
// so that host, device and every row shard regenerate identical entries.
static inline uint64_t mix64(uint64_t z) {
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}
double orc_uniform01(uint64_t seed, uint64_t lo, uint64_t hi) {
  uint64_t h = mix64(seed + 0x9E3779B97F4A7C15ULL * (lo + 1));
  h = mix64(h ^ (0xD6E8FEB86659FD93ULL * (hi + 1)));
  return (double)(h >> 11) * (1.0 / 9007199254740992.0);  // [0,1), 53 bits
}

// generate_diagonal_dominant (array_utils.f90:86-113): random_number * sparsity, upper triangle
// mirrored onto the lower (arr(i,j) = arr(j,i) for i > j), diagonal = diag_val or the 1-based
// row index.  diag_val == NULL means "not present".
void orc_generate_diagonal_dominant(int m, double sparsity, const double* diag_val, uint64_t seed, double* arr) {
#pragma omp parallel for schedule(static)
  for (int j = 0; j < m; ++j) {
    for (int i = 0; i < m; ++i) {
      if (i == j) {
        arr[idx(i, j, m)] = diag_val ? *diag_val : (double)(i + 1);
      } else {
        const uint64_t lo = (uint64_t)std::min(i, j), hi = (uint64_t)std::max(i, j);
        arr[idx(i, j, m)] = orc_uniform01(seed, lo, hi) * sparsity;
      }
    }
  }
}

// diagonal (array_utils.f90:115-134): the O(m^2) double loop is kept (it is part of the
// reference's cost).
void orc_diagonal(int m, const double* matrix, double* diagonal) {
  for (int i = 0; i < m; ++i)
    for (int j = 0; j < m; ++j)
      if (i == j) diagonal[i] = matrix[idx(i, j, m)];
}

// search_key (array_utils.f90:162-179): first j with keys(j) == i; undefined if absent.  The
// oracle returns -1 when absent (caller documents what it does then).
static int search_key(const std::vector<int>& keys, int i) {
  for (size_t j = 0; j < keys.size(); ++j)
    if (keys[j] == i) return (int)j;
  return -1;
}

// generate_preconditioner (array_utils.f90:136-160): sorts diag in place (!) and returns the
// n x dim_sub matrix of unit vectors e_{p_i}, p_i = position of the i-th smallest diagonal entry.
// When the reference's key search is undefined (duplicated diagonal values skip a rank) the
// oracle falls back to the stable order by index, which is the documented behaviour of the CUDA path.
void orc_generate_preconditioner(int n, double* diag, int dim_sub, double* precond) {
  dvec original(diag, diag + n);
  std::vector<int> keys(n);
  orc_lapack_sort('I', n, diag, keys.data());
  std::fill(precond, precond + (size_t)n * dim_sub, 0.0);
  std::vector<int> pos(dim_sub);
  bool defined = true;
  for (int i = 1; i <= dim_sub; ++i) {
    pos[i - 1] = search_key(keys, i);
    if (pos[i - 1] < 0) defined = false;
  }
  if (!defined) {  // reference behaviour undefined (a rank is skipped): stable order by index for ALL columns
    std::vector<int> stable(n);
    for (int t = 0; t < n; ++t) stable[t] = t;
    std::stable_sort(stable.begin(), stable.end(), [&](int a, int b) { return original[a] < original[b]; });
    for (int i = 0; i < dim_sub; ++i) pos[i] = stable[i];
  }
  for (int i = 0; i < dim_sub; ++i) precond[idx(pos[i], i, n)] = 1.0;
}

// -------------------------------------------------------------------------------------
// Matrix-free operators (benchmark_free.f90:38-76 and tests/test_utils.f90:37-116)
// -------------------------------------------------------------------------------------
enum {
  ORC_OP_BENCHMARK_MTX = 0,  // benchmark_free.f90:38-63
  ORC_OP_IDENTITY = 1,       // benchmark_free.f90:65-76 (stx = identity)
  ORC_OP_TEST_MTX = 2,       // test_utils.f90:37-51 + expensive_function_1 (identical to 0)
  ORC_OP_TEST_STX = 3,       // test_utils.f90:54-68 + expensive_function_2 (sin variant, diag := 1)
  ORC_OP_DENSE = 4           // apply a stored dense matrix (used for host-callback style tests)
};

// compute_matrix_on_the_fly(i, dim) -> column i (1-based).  `exp(real(i)/real(dim))` is single
// precision (benchmark_free.f90:50,53), `1e-4` is a single precision literal (:55,57), the
// diagonal gets `real(i)` (single, exact below 2^24) added (:61).
void orc_compute_on_the_fly(int op, int i, int dim, double* vector) {
  if (op == ORC_OP_IDENTITY) {
    std::fill(vector, vector + dim, 0.0);
    vector[i - 1] = 1.0;
    return;
  }
  const double scale = (double)1e-4f;
  const double x = (double)expf((float)i / (float)dim);
  for (int j = 1; j <= dim; ++j) {
    const double y = (double)expf((float)j / (float)dim);
    const double a = (j >= i) ? std::atan2(x, y) : std::atan2(y, x);
    const double l = std::log(std::sqrt(a));
    vector[j - 1] = (op == ORC_OP_TEST_STX ? std::sin(l) : std::cos(l)) * scale;
  }
  if (op == ORC_OP_TEST_STX)
    vector[i - 1] = 1.0;
  else
    vector[i - 1] = vector[i - 1] + (double)(float)i;
}

struct FreeOp {
  int op;
  int dim;
  const double* dense;  // ORC_OP_DENSE only
};

// free_matmul (davidson.f90:526-569): matrix(i,j) = dot_product(fun(i,dim), array(:,j)),
// OpenMP-parallel over i.  (Uses column i as row i: relies on symmetry.)
void orc_free_matmul(int op, int dim1, int dim2, const double* array, double* matrix, const double* dense) {
#pragma omp parallel
  {
    dvec vec(dim1);
#pragma omp for schedule(static)
    for (int i = 1; i <= dim1; ++i) {
      if (op == ORC_OP_DENSE) {
        std::copy(dense + (size_t)(i - 1) * dim1, dense + (size_t)i * dim1, vec.begin());
      } else {
        orc_compute_on_the_fly(op, i, dim1, vec.data());
      }
      for (int j = 0; j < dim2; ++j) {
        double s = 0.0;
        const double* col = array + (size_t)j * dim1;
        for (int l = 0; l < dim1; ++l) s += vec[l] * col[l];
        matrix[idx(i - 1, j, dim1)] = s;
      }
    }
  }
}

// -------------------------------------------------------------------------------------
// davidson.f90 : corrections (submodule correction_methods_generalized_dense, :630-752)
// -------------------------------------------------------------------------------------

// compute_DPR_generalized_dense (:673-698)
static void compute_DPR_generalized_dense(int m, int k, const double* matrix, const double* eigenvalues,
                                          const double* residues, const double* second_matrix, double* correction) {
  const bool gev = second_matrix != nullptr;
  for (int j = 0; j < k; ++j)
    for (int ii = 0; ii < m; ++ii) {
      if (gev)
        correction[idx(ii, j, m)] =
            residues[idx(ii, j, m)] / (eigenvalues[j] * second_matrix[idx(ii, ii, m)] - matrix[idx(ii, ii, m)]);
      else
        correction[idx(ii, j, m)] = residues[idx(ii, j, m)] / (eigenvalues[j] - matrix[idx(ii, ii, m)]);
    }
}

// compute_GJD_generalized_dense (:700-734): per column k:
//   xs = I - u u^T ; ys = A - theta B (or A - theta I) ; arr = xs * (ys * xs) ; DSYSV arr t = -r.
static int compute_GJD_generalized_dense(int m, int ncols, const double* matrix, const double* eigenvalues,
                                         const double* ritz_vectors, const double* residues,
                                         const double* second_matrix, double* correction) {
  const bool gev = second_matrix != nullptr;
  const size_t mm = (size_t)m * m;
  dvec arr(mm), xs(mm), ys(mm), tmp(mm), brr(m);
  for (int k = 0; k < ncols; ++k) {
    const double* rs = ritz_vectors + (size_t)k * m;  // :720
    // xs = eye(m,m) - rs rs^T (:721, lapack_matmul('N','T', rs, rs))
    orc_lapack_matmul('N', 'T', m, 1, rs, m, 1, rs, 1.0, xs.data());
    for (size_t t = 0; t < mm; ++t) xs[t] = -xs[t];
    for (int i = 0; i < m; ++i) xs[idx(i, i, m)] += 1.0;
    if (gev) {  // :722-726
      for (size_t t = 0; t < mm; ++t) ys[t] = matrix[t] - eigenvalues[k] * second_matrix[t];
    } else {
      std::copy(matrix, matrix + mm, ys.begin());  // substract_from_diagonal (:736-750)
      for (int i = 0; i < m; ++i) ys[idx(i, i, m)] -= eigenvalues[k];
    }
    orc_lapack_matmul('N', 'N', m, m, ys.data(), m, m, xs.data(), 1.0, tmp.data());  // ys * xs
    orc_lapack_matmul('N', 'N', m, m, xs.data(), m, m, tmp.data(), 1.0, arr.data()); // xs * (ys*xs), :727
    for (int i = 0; i < m; ++i) brr[i] = -residues[idx(i, k, m)];                    // :728
    if (int e = orc_lapack_solver(m, arr.data(), brr.data())) return e;              // :730
    std::copy(brr.begin(), brr.end(), correction + (size_t)k * m);                   // :731
  }
  return 0;
}

enum { ORC_METHOD_DPR = 0, ORC_METHOD_GJD = 1 };

// Per-iteration trace (what the parity harness compares): basis width at the RR step and the
// largest residual norm among the `lowest` pairs.
struct Trace {
  int* k;
  double* max_err;
  int cap;
  void put(int it, int kk, double e) {
    if (k && it < cap) { k[it] = kk; max_err[it] = e; }
  }
};

// -------------------------------------------------------------------------------------
// generalized_eigensolver_dense (davidson.f90:51-246)
// max_dim_sub <= 0 means "not present"; second_matrix == NULL means "not present".
// Returns 0, or a LAPACK info (the reference would `error stop`).
// -------------------------------------------------------------------------------------
int orc_generalized_eigensolver_dense(int n, const double* matrix, const double* second_matrix, int lowest,
                                      int method, int max_iterations, double tolerance, int max_dim_sub,
                                      double* eigenvalues, double* eigenvectors, int* iters, int* trace_k,
                                      double* trace_err, int trace_cap) {
  Trace trace{trace_k, trace_err, trace_cap};
  const int initial_dimension = lowest * 2;                              // :108
  std::vector<char> has_converged(lowest, 0);                            // :112
  const int max_dim = max_dim_sub > 0 ? max_dim_sub : lowest * 10;       // :115-119
  const bool gev = second_matrix != nullptr;                             // :122

  dvec d(n);
  orc_diagonal(n, matrix, d.data());                                     // :127
  int kcur = initial_dimension;
  dvec V((size_t)n * kcur);
  orc_generate_preconditioner(n, d.data(), initial_dimension, V.data()); // :128

  dvec tmp, matrix_proj, second_matrix_proj;
  auto project = [&](const double* M, dvec& out) {                       // V^T (M V), :131,134,223,226
    tmp.assign((size_t)n * kcur, 0.0);
    orc_lapack_matmul('N', 'N', n, n, M, n, kcur, V.data(), 1.0, tmp.data());
    out.assign((size_t)kcur * kcur, 0.0);
    orc_lapack_matmul('T', 'N', n, kcur, V.data(), n, kcur, tmp.data(), 1.0, out.data());
  };
  project(matrix, matrix_proj);
  if (gev) project(second_matrix, second_matrix_proj);

  dvec eigenvalues_sub, eigenvectors_sub, ritz_vectors, residues, correction, guess(n), av(n), errors(lowest);
  int i;
  for (i = 1; i <= max_iterations; ++i) {                                // :138
    eigenvalues_sub.assign(kcur, 0.0);
    eigenvectors_sub.assign((size_t)kcur * kcur, 0.0);
    if (int e = orc_lapack_generalized_eigensolver(kcur, matrix_proj.data(),
                                                   gev ? second_matrix_proj.data() : nullptr,
                                                   eigenvalues_sub.data(), eigenvectors_sub.data()))
      return e;                                                          // :152-156
    ritz_vectors.assign((size_t)n * kcur, 0.0);
    orc_lapack_matmul('N', 'N', n, kcur, V.data(), kcur, kcur, eigenvectors_sub.data(), 1.0,
                      ritz_vectors.data());                              // :159
    residues.assign((size_t)n * kcur, 0.0);
    for (int j = 0; j < kcur; ++j) {                                     // :163-170, one DGEMV (two if gev) per column
      const double* rv = ritz_vectors.data() + (size_t)j * n;
      if (gev) {
        orc_lapack_matrix_vector('N', n, n, second_matrix, rv, 1.0, guess.data());
        for (int t = 0; t < n; ++t) guess[t] = eigenvalues_sub[j] * guess[t];
      } else {
        for (int t = 0; t < n; ++t) guess[t] = eigenvalues_sub[j] * rv[t];
      }
      orc_lapack_matrix_vector('N', n, n, matrix, rv, 1.0, av.data());
      double* r = residues.data() + (size_t)j * n;
      for (int t = 0; t < n; ++t) r[t] = av[t] - guess[t];
    }
    double max_err = 0.0;
    for (int j = 0; j < lowest; ++j) {                                   // :173-178
      errors[j] = orc_norm(n, residues.data() + (size_t)j * n);
      if (errors[j] < tolerance) has_converged[j] = 1;
      max_err = std::max(max_err, errors[j]);
    }
    trace.put(i - 1, kcur, max_err);
    std::copy(eigenvalues_sub.begin(), eigenvalues_sub.begin() + lowest, eigenvalues);            // :186
    std::copy(ritz_vectors.begin(), ritz_vectors.begin() + (size_t)n * lowest, eigenvectors);     // :187
    bool all_conv = true;
    for (int j = 0; j < lowest; ++j) all_conv = all_conv && has_converged[j];
    if (all_conv) {                                                      // :189-192
      *iters = i;
      break;
    }
    if (kcur <= max_dim) {                                               // :195
      correction.assign((size_t)n * kcur, 0.0);
      if (method == ORC_METHOD_DPR) {                                    // :656-669
        compute_DPR_generalized_dense(n, kcur, matrix, eigenvalues_sub.data(), residues.data(), second_matrix,
                                      correction.data());
      } else if (method == ORC_METHOD_GJD) {
        if (int e = compute_GJD_generalized_dense(n, kcur, matrix, eigenvalues_sub.data(), ritz_vectors.data(),
                                                  residues.data(), second_matrix, correction.data()))
          return e;
      } else {
        return -1000;  // the reference leaves `correction` undefined for an unknown method
      }
      V.insert(V.end(), correction.begin(), correction.end());           // concatenate, :210
      kcur *= 2;
      if (int e = orc_lapack_qr(n, kcur, V.data())) return e;            // :213
    } else {
      dvec Vnew((size_t)n * initial_dimension, 0.0);                     // :218
      orc_lapack_matmul('N', 'N', n, kcur, V.data(), kcur, initial_dimension, eigenvectors_sub.data(), 1.0,
                        Vnew.data());
      V.swap(Vnew);
      kcur = initial_dimension;
    }
    project(matrix, matrix_proj);                                        // :223
    if (gev) project(second_matrix, second_matrix_proj);                 // :225-227
  }
  if (i > max_iterations) {                                              // :232-235
    *iters = i;
    std::printf(" Warning: Algorithm did not converge!!\n");
  }
  return 0;
}

// compute_DPR_free (davidson.f90:463-488)
static void compute_DPR_free(int m, int k, const double* eigenvalues, const double* residues,
                             const double* diag_matrix, const double* diag_second_matrix, double* correction) {
  for (int j = 0; j < k; ++j)
    for (int ii = 0; ii < m; ++ii)
      correction[idx(ii, j, m)] =
          residues[idx(ii, j, m)] / (eigenvalues[j] * diag_second_matrix[ii] - diag_matrix[ii]);
}

// extract_diagonal_free (davidson.f90:490-523): n applications of the operator to unit vectors.
static void extract_diagonal_free(const FreeOp& f, int dim, double* out) {
  dvec tmp_array(dim), res(dim);
  for (int ii = 0; ii < dim; ++ii) {
    std::fill(tmp_array.begin(), tmp_array.end(), 0.0);
    tmp_array[ii] = 1.0;
    orc_free_matmul(f.op, dim, 1, tmp_array.data(), res.data(), f.dense);
    out[ii] = res[ii];
  }
}

// -------------------------------------------------------------------------------------
// generalized_eigensolver_free (davidson.f90:277-460).  `method` is accepted and ignored
// (always DPR, :428); convergence is non-sticky (:416); iters is only assigned on convergence
// (:417) -- the oracle initialises it to -1 so "unassigned" is observable.
// op_a / op_b: ORC_OP_* ; dense_a / dense_b only for ORC_OP_DENSE.
// -------------------------------------------------------------------------------------
int orc_generalized_eigensolver_free(int dim_matrix, int op_a, const double* dense_a, int op_b,
                                     const double* dense_b, int lowest, int method, int max_iterations,
                                     double tolerance, int max_dim_sub, double* eigenvalues, double* ritz_vectors,
                                     int* iters, int* trace_k, double* trace_err, int trace_cap) {
  (void)method;
  Trace trace{trace_k, trace_err, trace_cap};
  const FreeOp fa{op_a, dim_matrix, dense_a}, fb{op_b, dim_matrix, dense_b};
  const int n = dim_matrix;
  const int initial_dimension = lowest * 2;                              // :352
  const int max_dim = max_dim_sub > 0 ? max_dim_sub : lowest * 10;       // :355-359
  *iters = -1;

  dvec diag_matrix(n), diag_second_matrix(n), copy_d;
  extract_diagonal_free(fa, n, diag_matrix.data());                      // :365
  extract_diagonal_free(fb, n, diag_second_matrix.data());               // :366
  copy_d = diag_matrix;                                                  // :371
  int kcur = initial_dimension;
  dvec V((size_t)n * kcur);
  orc_generate_preconditioner(n, copy_d.data(), initial_dimension, V.data());  // :372

  dvec matrixV, second_matrixV, matrix_proj, second_matrix_proj, eigenvalues_sub, eigenvectors_sub, residues, guess,
      correction, errors(lowest);
  int i;
  for (i = 1; i <= max_iterations; ++i) {                                // :375
    matrixV.assign((size_t)n * kcur, 0.0);
    second_matrixV.assign((size_t)n * kcur, 0.0);
    orc_free_matmul(fa.op, n, kcur, V.data(), matrixV.data(), fa.dense);         // :378
    orc_free_matmul(fb.op, n, kcur, V.data(), second_matrixV.data(), fb.dense);  // :379
    matrix_proj.assign((size_t)kcur * kcur, 0.0);
    second_matrix_proj.assign((size_t)kcur * kcur, 0.0);
    orc_lapack_matmul('T', 'N', n, kcur, V.data(), n, kcur, matrixV.data(), 1.0, matrix_proj.data());  // :380
    orc_lapack_matmul('T', 'N', n, kcur, V.data(), n, kcur, second_matrixV.data(), 1.0,
                      second_matrix_proj.data());                                                      // :381
    eigenvalues_sub.assign(kcur, 0.0);
    eigenvectors_sub.assign((size_t)kcur * kcur, 0.0);
    if (int e = orc_lapack_generalized_eigensolver(kcur, matrix_proj.data(), second_matrix_proj.data(),
                                                   eigenvalues_sub.data(), eigenvectors_sub.data()))
      return e;                                                          // :394
    // ritz_vectors = V * eigenvectors_sub(:, :lowest)  (:397)
    orc_lapack_matmul('N', 'N', n, kcur, V.data(), kcur, lowest, eigenvectors_sub.data(), 1.0, ritz_vectors);
    // residues = matrixV*y - (second_matrixV*y)*lambda  (:401-410)
    residues.assign((size_t)n * kcur, 0.0);
    guess.assign((size_t)n * kcur, 0.0);
    orc_lapack_matmul('N', 'N', n, kcur, second_matrixV.data(), kcur, kcur, eigenvectors_sub.data(), 1.0,
                      residues.data());
    for (int j = 0; j < kcur; ++j)
      for (int t = 0; t < n; ++t) guess[idx(t, j, n)] = residues[idx(t, j, n)] * eigenvalues_sub[j];
    orc_lapack_matmul('N', 'N', n, kcur, matrixV.data(), kcur, kcur, eigenvectors_sub.data(), 1.0,
                      residues.data());
    for (size_t t = 0; t < residues.size(); ++t) residues[t] -= guess[t];
    double max_err = 0.0;
    bool all_conv = true;
    for (int j = 0; j < lowest; ++j) {                                   // :412-416
      errors[j] = orc_norm(n, residues.data() + (size_t)j * n);
      max_err = std::max(max_err, errors[j]);
      all_conv = all_conv && (errors[j] < tolerance);
    }
    trace.put(i - 1, kcur, max_err);
    if (all_conv) {
      *iters = i;
      break;
    }
    if (kcur <= max_dim) {                                               // :422
      correction.assign((size_t)n * kcur, 0.0);
      compute_DPR_free(n, kcur, eigenvalues_sub.data(), residues.data(), diag_matrix.data(),
                       diag_second_matrix.data(), correction.data());    // :428
      V.insert(V.end(), correction.begin(), correction.end());           // :431
      kcur *= 2;
      if (int e = orc_lapack_qr(n, kcur, V.data())) return e;            // :434
    } else {
      dvec Vnew((size_t)n * initial_dimension, 0.0);                     // :438
      orc_lapack_matmul('N', 'N', n, kcur, V.data(), kcur, initial_dimension, eigenvectors_sub.data(), 1.0,
                        Vnew.data());
      V.swap(Vnew);
      kcur = initial_dimension;
    }
  }
  if (i > max_iterations / initial_dimension) {                          // :444 (sic)
    std::printf(" Warning: Algorithm did not converge!!\n");
  }
  std::copy(eigenvalues_sub.begin(), eigenvalues_sub.begin() + lowest, eigenvalues);  // :451
  return 0;
}

}  // extern "C"
