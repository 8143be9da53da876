// Thin NCCL binding resolved at run time with dlopen: the library has no link-time dependency on
// NCCL (single-GPU use never touches it) and shares whichever libnccl.so.2 the process already
// loaded (torch bundles one).  Used only for the small per-iteration exchanges of the row-block
// sharded solver: all-gather of the new basis block, all-reduce of the k x k projection / Gram
// partials and of the residual norms.
#pragma once
#include "common.cuh"

namespace dav {

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

  // in-place sum over ranks
  void allreduce_sum(double* buf, size_t count, cudaStream_t s);
  // recv[r*count .. (r+1)*count) = send of rank r
  void allgather(const void* send, void* recv, size_t bytes_per_rank, cudaStream_t s);

 private:
  int rank_ = 0, world_ = 1;
  void* comm_ = nullptr;
};

}  // namespace dav
