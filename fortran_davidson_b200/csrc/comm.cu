#include "comm.cuh"

#include <dlfcn.h>

#include <cstring>
#include <mutex>

namespace dav {
namespace {

struct NcclUniqueId { char internal[128]; };
typedef int ncclResult_t;
typedef void* ncclComm_t;
enum { NCCL_UINT8 = 1, NCCL_FLOAT64 = 8, NCCL_SUM = 0 };

struct Api {
  ncclResult_t (*GetUniqueId)(NcclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, NcclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  bool ok = false;
  std::string why;
};

Api& api() {
  static Api a;
  static std::once_flag once;
  std::call_once(once, [] {
    const char* names[] = {"libnccl.so.2", "libnccl.so", "/usr/lib/x86_64-linux-gnu/libnccl.so.2"};
    void* h = nullptr;
    for (const char* nm : names) {
      h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
      if (h) break;
    }
    if (!h) {
      a.why = "libnccl.so.2 not found";
      return;
    }
#define SYM(field, name)                                   \
  *(void**)(&a.field) = dlsym(h, name);                    \
  if (!a.field) {                                          \
    a.why = std::string("missing NCCL symbol ") + name;    \
    return;                                                \
  }
    SYM(GetUniqueId, "ncclGetUniqueId");
    SYM(CommInitRank, "ncclCommInitRank");
    SYM(CommDestroy, "ncclCommDestroy");
    SYM(AllReduce, "ncclAllReduce");
    SYM(AllGather, "ncclAllGather");
    SYM(GetErrorString, "ncclGetErrorString");
#undef SYM
    a.ok = true;
  });
  return a;
}

void check(ncclResult_t r, const char* what) {
  if (r != 0) DAV_THROW(DAV_ERR_COMM, "NCCL %s failed: %s", what, api().GetErrorString ? api().GetErrorString(r) : "?");
}

}  // namespace

void Comm::get_unique_id(void* id128) {
  Api& a = api();
  if (!a.ok) DAV_THROW(DAV_ERR_COMM, "NCCL unavailable: %s", a.why.c_str());
  NcclUniqueId id;
  check(a.GetUniqueId(&id), "ncclGetUniqueId");
  std::memcpy(id128, &id, 128);
}

void Comm::init(int rank, int world, const void* id128) {
  rank_ = rank;
  world_ = world;
  if (world <= 1) return;
  Api& a = api();
  if (!a.ok) DAV_THROW(DAV_ERR_COMM, "NCCL unavailable: %s", a.why.c_str());
  NcclUniqueId id;
  std::memcpy(&id, id128, 128);
  ncclComm_t c = nullptr;
  check(a.CommInitRank(&c, world, id, rank), "ncclCommInitRank");
  comm_ = c;
}

Comm::~Comm() {
  if (comm_) api().CommDestroy((ncclComm_t)comm_);
}

void Comm::allreduce_sum(double* buf, size_t count, cudaStream_t s) {
  if (world_ <= 1 || count == 0) return;
  check(api().AllReduce(buf, buf, count, NCCL_FLOAT64, NCCL_SUM, (ncclComm_t)comm_, s), "ncclAllReduce");
}

void Comm::allgather(const void* send, void* recv, size_t bytes_per_rank, cudaStream_t s) {
  if (world_ <= 1) {
    if (send != recv) CK(cudaMemcpyAsync(recv, send, bytes_per_rank, cudaMemcpyDeviceToDevice, s));
    return;
  }
  check(api().AllGather(send, recv, bytes_per_rank, NCCL_UINT8, (ncclComm_t)comm_, s), "ncclAllGather");
}

}  // namespace dav
