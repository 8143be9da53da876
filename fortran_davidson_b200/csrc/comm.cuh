// Inter-GPU exchanges of the row-block sharded solver (one process per GPU, one node).
//
// Two transports behind one interface:
//   * PEER (default on a node whose GPUs see each other over NVLink / NVSwitch): every rank exports a few cudaMalloc
//     segments with cudaIpc, maps its peers' segments, and the exchanges are OUR kernels storing straight into peer
//     memory ("push"), followed by a flag with release semantics at system scope.
//       - all-reduce (k x k projection / Gram partials, residual norms: <= a few 100 KB, latency-bound): one kernel,
//         one-shot, fused with the split-K reduction that produces the partials and with the layout of the result:
//         every value travels with its flag in one 16-byte store into line[rank] of every peer; the receiver polls
//         the lines and adds them in rank order -> bit-identical on every rank, independent of arrival order, one
//         NVLink latency instead of a ~40 us NCCL call.
//       - all-gather of the new basis block: every rank stores its rows directly into every peer's copy of the full
//         n x b block -- either column-major or already in the MMA-fragment order the block matvec consumes -- so
//         the stage / all-gather / unstage / pack passes of the NCCL path disappear.
//   * NCCL (fallback when peer mapping is unavailable, and the bootstrap of the peer transport: the IPC handles
//     are exchanged with one ncclAllGather).  Resolved at run time with dlopen: no link-time dependency, single-GPU
//     use never touches it.
#pragma once
#include <vector>

#include "common.cuh"

namespace dav {

constexpr int COMM_MAX_RANKS = 16;

// symmetric control block (one per rank, peers write their own entry of each array)
struct PeerCtl {
  unsigned long long ar_flag[COMM_MAX_RANKS];   // all-reduce: epoch of the last contribution of rank r
  unsigned long long ag_ready[COMM_MAX_RANKS];  // gather: rank r has reached gather #epoch (its readers are done)
  unsigned long long ag_done[COMM_MAX_RANKS];   // gather: rank r's rows of gather #epoch have landed here
  unsigned int counter[4];                      // local CTA arrival counters
  int error;                                    // local: a wait timed out
  int pad[3];
};

// what the exchange kernels need: the peers' views of one symmetric segment + the control blocks
struct PeerArgs {
  double* data[COMM_MAX_RANKS];
  PeerCtl* ctl[COMM_MAX_RANKS];
  int rank, world;
};

// where a reduced k x b block goes (Comm::reduce_sum):
//   mode 0: C(m, j) at C[m + j*ldc]
//   mode 1: block column `kold..` of a symmetric projected matrix (leading dimension ldc): entry (m, kold+j) is stored
//           when it lies in the upper triangle and mirrored below the diagonal (what copy + symmetrize_from_upper did)
struct ReduceOut {
  int mode;
  double* C;
  int64_t ldc;
  int kold;
};

// a symmetric segment: the same number of bytes on every rank, every rank holds a mapping of every peer's copy
struct SymBuf {
  void* local = nullptr;
  void* peer[COMM_MAX_RANKS] = {};
  size_t bytes = 0;
  double* p() const { return (double*)local; }
};

class Comm {
 public:
  Comm() {}
  ~Comm();
  Comm(const Comm&) = delete;
  Comm& operator=(const Comm&) = delete;

  static void get_unique_id(void* id128);
  void init(int rank, int world, const void* id128);
  int rank() const { return rank_; }
  int world() const { return world_; }
  bool active() const { return world_ > 1; }
  bool peer() const { return peer_; }  // peer-memory transport in use

  // in-place sum over ranks (bit-identical on every rank)
  void allreduce_sum(double* buf, size_t count, cudaStream_t s);
  // ONE kernel: element (m, j) = sum_z src[z*M*N + m + j*M] (the split-K partials of a tall-skinny product, fixed
  // order), summed over the ranks, stored per `out`.  Works on a single rank too (then it is the split-K reduction
  // with the output layout fused in).
  void reduce_sum(const double* src, int splits, int64_t M, int64_t N, const ReduceOut& out, cudaStream_t s);
  // recv[r*bytes .. (r+1)*bytes) = send of rank r (small payloads; NCCL)
  void allgather(const void* send, void* recv, size_t bytes_per_rank, cudaStream_t s);
  // two segments per rank, send = [a | b]:  recv = [a of rank 0 .. a of rank P-1 | b of rank 0 .. b of rank P-1]
  void allgather2(const void* send, void* recv, size_t bytes_a, size_t bytes_b, cudaStream_t s);

  // ---- peer transport only -----------------------------------------------------------------------------------
  // collective: (re)allocate a symmetric segment of at least `bytes` (same value on every rank)
  void sym_reserve(SymBuf& b, size_t bytes, cudaStream_t s);
  void sym_release(SymBuf& b);
  // largest all-reduce the one-shot kernel has to carry (collective; grows the slot segment)
  void reserve_allreduce(size_t max_count, cudaStream_t s);
  // dst (symmetric, n x b column-major, leading dimension n) <- rows [row0, row0+nl) of every rank
  void gather_rows(const double* Xlocal, int64_t ldx, int64_t nl, int64_t row0, int64_t n, int b, SymBuf& dst,
                   cudaStream_t s);
  // same, stored in the packed MMA-fragment order of the block matvec: element (k, j) of the chunk of `bpad`
  // columns starting at column c0 lives at  c0*Kpad + ((k/8 * (bpad/8) + (j-c0)/8) * 64 + ((j-c0)%8)*8 + k%8);
  // rows K..Kpad and columns beyond b are zero filled.  chunk width 128 (the matvec's widest tile).
  void gather_rows_packed(const double* Xlocal, int64_t ldx, int64_t nl, int64_t row0, int64_t n, int64_t Kpad, int b,
                          SymBuf& dst, cudaStream_t s);
  // nonzero when a wait inside an exchange kernel timed out (a peer died); read with a stream sync by the caller
  const int* error_flag() const { return ctl_.local ? &((PeerCtl*)ctl_.local)->error : nullptr; }
  long long peer_calls = 0, nccl_calls = 0;
  // self-check + timing of one exchange (dav_debug_collective): kind 0 all-reduce (transport in use), 1 all-reduce
  // through NCCL, 2 gather_rows, 3 gather_rows_packed; count = doubles (0, 1) or columns (2, 3).
  // out[0] = microseconds per call (events, `reps` back-to-back calls), out[1] = max abs error of the result.
  void debug_exchange(int kind, int64_t count, int reps, int64_t n, int64_t nl, int64_t row0, cudaStream_t s,
                      double* out);

 private:
  void setup_peer(cudaStream_t s);
  void nccl_allreduce(double* buf, size_t count, cudaStream_t s);
  void launch_ll(const double* src, int splits, int64_t M, int64_t total, const ReduceOut& out, int64_t gather_seg,
                 bool exchange, cudaStream_t s);
  DevBuf<double> nccl_tmp_;  // contiguous staging of reduce_sum on the NCCL transport
  PeerArgs args_for(const SymBuf& b) const;
  int rank_ = 0, world_ = 1;
  void* comm_ = nullptr;
  bool peer_ = false, peer_tried_ = false;
  SymBuf ctl_, slots_;
  size_t slot_cap_ = 0;  // 16-byte lines per (parity, source rank)
  unsigned long long ar_epoch_ = 0, ag_epoch_ = 0;
  void* hbuf_ = nullptr;  // device scratch for the handle exchange
};

// packed layout of the block matvec (see matvec_dmma.cu): padded column count of a chunk of bc <= 128 columns
int matvec_bpad(int bc);

}  // namespace dav
