// Device-resident block Davidson driver (replaces generalized_eigensolver_dense / _free,
// davidson.f90:51-246 and :277-460).  See solver.cu for the algorithm notes.
#pragma once
#include <vector>

#include "comm.cuh"
#include "kernels.cuh"

struct dav_solver {
  enum Kind { NONE = 0, DENSE = 1, BUILTIN = 2, CALLBACK = 3, DEVCALLBACK = 4 };
  struct Matrix {
    Kind kind = NONE;
    int64_t n = 0;
    dav::DevBuf<double> A;  // DENSE: local row block, nl x n column-major, leading dimension lda
    int64_t lda = 0;
    dav::MatvecPlan* plan = nullptr;
    int op = 0;                      // BUILTIN
    dav::FreeTables* ftab = nullptr; // BUILTIN: tables of the tensor-pipe generator (freeops_dmma.cu)
    dav_gemv_fn fn = nullptr;        // CALLBACK
    dav_device_gemv_fn dfn = nullptr;  // DEVCALLBACK: enqueues on the solver's stream, device pointers
    void* ctx = nullptr;
    dav::DevBuf<double> diag;        // local diagonal entries
    bool diag_valid = false;
  };

  int device = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t stream2 = nullptr;  // mirrors the uploaded panels of a symmetric matrix while the next ones arrive
  dav::Comm comm;
  int64_t n = 0, nl = 0, row0 = 0, chunk = 0;  // global size, local rows, first local row, rows per rank
  Matrix mat[2];
  dav::DevBuf<double> etab;  // e_t table of the built-in operators
  std::vector<double> etab_host;
  int matvec_impl = DAV_MATVEC_AUTO;
  bool local_vectors = false;  // dav_solve_local: Ritz vectors returned row-sharded
  bool profile_spans = false;  // dav_set_profiling: per-phase event spans (the *_ms fields of dav_stats_t)
  dav_stats_t stats;

  // ---- work space of a solve
  int64_t ldv = 0;
  int kcap = 0;
  dav::DevBuf<double> V, AV, BV, R, C, T, Xfull, stage_s, stage_r;
  // peer transport (comm.cuh): every rank's copy of the gathered block, written by all ranks directly
  dav::SymBuf xsym;  // n x b column-major (generic consumers: matrix-free operators, output, SIMT kernels)
  dav::SymBuf xpk;   // packed MMA-fragment order (the dense TMA/DMMA matvec)
  dav::DevBuf<double> Ap, Bp, Y, theta, G, U, sv, D, Tm, S1, S2, Z, jscratch, norms2, partial, gemm_ws, gemm_ws2, small;
  dav::DevBuf<int> status, flags, gjd_active;
  dav::DevBuf<double> gjd_buf, gjd_st;
  dav::DevBuf<int64_t> idx, topk_idx;
  dav::DevBuf<double> cand_val, topk_val;
  std::vector<double> host_x, host_y;  // callback staging
  // page-locked staging of the eigenvectors on their way to the caller's (pageable) array: a direct device ->
  // pageable copy runs at ~6 GB/s (2 ms for the 12.8 MB of n = 100,000 x 16), through here at PCIe speed
  double* pinned_out = nullptr;
  size_t pinned_out_n = 0;
  double* pinned(size_t count);
  double* pip_flags_host = nullptr;  // page-locked landing zone of the orthonormalisation flags
  cudaEvent_t pip_flags_ev = nullptr;
  std::vector<cudaEvent_t> ev_pool;
  struct Span { int a, b, kind; };
  std::vector<Span> spans;
  int ev_used = 0;

  dav_solver(int device, int rank, int world, const void* id128);
  ~dav_solver();

  void set_dims(int64_t n);
  void clear_matrix(int which);
  void generate_diagonal_dominant(int which, int64_t n, double sparsity, int has_diag, double diag_val,
                                  uint64_t seed);
  void upload(int which, int64_t n, const double* host, int64_t ld);
  bool host_looks_symmetric(const double* host, int64_t ld, int64_t n);
  double last_upload_bytes = 0;  // bytes the last upload moved over PCIe (the symmetric upload moves half)
  void upload_rows(int which, int64_t n, const double* host_rows, int64_t ld);
  void set_operator(int which, int64_t n, int op);
  void set_callback(int which, int64_t n, dav_gemv_fn fn, void* ctx, const double* diag);
  void set_device_callback(int which, int64_t n, dav_device_gemv_fn fn, void* ctx, const double* diag);
  void download(int which, double* host_rows, int64_t ld);
  void ensure_etab();
  void ensure_diag(int which);
  void ensure_plan(int which, int max_b);

  // W(local rows x b) = M_which * X, X given as the local row block (nl x b, ldx) of a row-sharded block
  void apply(int which, const double* Xlocal, int64_t ldx, int b, double* W, int64_t ldw);
  // same with X already complete (n x b) on this rank
  void apply_full(int which, const double* Xfull_, int64_t ldx, int b, double* W, int64_t ldw);
  const double* gather_rows(const double* Xlocal, int64_t ldx, int b, int64_t* ld_out);
  // true when the new block can go straight into every peer's packed copy (all matrices dense on the DMMA path)
  bool packed_gather_usable() const;
  // W = M_which * X for the block last gathered with comm.gather_rows_packed into xpk
  void apply_packed(int which, int b, double* W, int64_t ldw);
  void count_matvec(int b);
  // WA = M_0 * X and (WB != nullptr) WB = M_1 * X for a row-sharded block X (nl x b, ldv): one exchange of X
  void apply_both(const double* Xlocal, int b, double* WA, double* WB);

  int solve(int lowest, int method, int max_iterations, double tolerance, int max_dim_sub, double* eigenvalues,
            double* eigenvectors, int64_t ldvec, int* iters);

  // pieces of the iteration
  void alloc_work(int lowest, int kcap_);
  void rayleigh_ritz(int k, bool gev);
  void orthonormalize_block(double* Cblk, int b, int kold, double* dest);
  bool orthonormalize_block_pip(int b, int kold);  // enqueues the fast path; false = shape unsupported, nothing enqueued
  bool pip_confirm();                              // true when the flags of the enqueued fast path accept it
  void gjd_correction(int k, bool gev, double outer_tolerance);
  void project_new_block(int which, int kold, int b);
  void tn_reduce(int M, int N, const double* A, const double* B, dav::DevBuf<double>& ws, const dav::ReduceOut& out,
                 const dav::GemmSplit* split = nullptr);
  void full_projection(int which, int k);
  void allreduce(double* buf, size_t count);
  void allgather(const void* send, void* recv, size_t bytes_per_rank);
  int begin_span(int kind);
  void end_span(int id);
  void check_status(const char* where);
};
