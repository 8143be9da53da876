#include "comm.cuh"

#include <dlfcn.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "kernels.cuh"

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

// ---- device side of the peer transport ------------------------------------------------------------------------
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// waits until *flag >= epoch; gives up after ~4 s (a peer died) and records it, so that a broken job ends with an
// error instead of hanging the GPU
__device__ __forceinline__ void wait_flag(const unsigned long long* flag, unsigned long long epoch, int* error) {
  if (ld_acquire_sys(flag) >= epoch) return;
  const unsigned long long t0 = globaltimer_ns();
  unsigned spins = 0;
  while (ld_acquire_sys(flag) < epoch) {
    if ((++spins & 1023u) == 0 && globaltimer_ns() - t0 > 4000000000ULL) {
      *error = 1;
      return;
    }
  }
}

// all CTAs have issued (and fenced) their peer stores -> the last one to arrive returns true
__device__ __forceinline__ bool last_cta_arrives(unsigned int* counter) {
  __shared__ int is_last;
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned prev = atomicAdd(counter, 1u);
    is_last = (prev == gridDim.x - 1);
    if (is_last) *counter = 0;  // every other CTA has already passed its atomicAdd
    __threadfence();
  }
  __syncthreads();
  return is_last != 0;
}

// ---- low-latency one-shot reduction (the small exchanges: projection / Gram partials, norms, candidates) -------
// Every 8-byte word travels in a 16-byte line {lo32, flag, hi32, flag} written with ONE 16-byte store; the receiver
// polls the line itself until both flags carry this call's number -- no fence, no separate flag, no counter: one
// NVLink one-way latency per exchange (the LL protocol of NCCL, which relies only on 8-byte store atomicity).
// Line layout on every rank: [parity][source rank][cap].  Two parities: a rank can start exchange e+1 (parity
// (e+1)&1) while a slower peer still polls the lines of e, but not e+2 -- that needs the peer's lines of e+1, which
// the peer stores only after its kernel of e has finished (stream order).  Flags never repeat within a parity
// (the call number itself, 32 bit), lines start zeroed, call numbers start at 1.
struct LLArgs {
  uint4* ll[COMM_MAX_RANKS];
  int rank, world;
};

__device__ __forceinline__ void ll_store(uint4* p, double v, unsigned flag) {
  const unsigned long long u = (unsigned long long)__double_as_longlong(v);
  asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"((unsigned)u), "r"(flag),
               "r"((unsigned)(u >> 32)), "r"(flag)
               : "memory");
}
__device__ __forceinline__ double ll_wait(const uint4* p, unsigned flag, int* error) {
  unsigned lo, f1, hi, f2;
  unsigned spins = 0;
  unsigned long long t0 = 0;
  for (;;) {
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(lo), "=r"(f1), "=r"(hi), "=r"(f2) : "l"(p)
                 : "memory");
    if (f1 == flag && f2 == flag) break;
    if ((++spins & 4095u) == 0) {  // a peer died: end with an error instead of hanging the GPU
      if (t0 == 0) t0 = globaltimer_ns();
      else if (globaltimer_ns() - t0 > 4000000000ULL) { *error = 1; break; }
    }
  }
  return __longlong_as_double((long long)(((unsigned long long)hi << 32) | lo));
}

// out modes of the reduction (ReduceOut in comm.cuh)
__device__ __forceinline__ void reduce_store(const ReduceOut& o, int64_t M, int64_t e, double v) {
  const int64_t m = e % M, j = e / M;
  if (o.mode == 0) {
    o.C[m + j * o.ldc] = v;
  } else if (o.mode == 1) {
    // block column kold.. of a projected matrix: the upper part is stored, the strictly upper part mirrored
    const int64_t c = o.kold + j;
    if (m <= c) o.C[m + c * o.ldc] = v;
    if (m < c) o.C[c + m * o.ldc] = v;
  } else {
    // two-segment gather (no sum is formed by the caller in this mode): handled in the kernel
  }
}

// value e = sum over z of src[z * total + e] (split-K partials, fixed order), summed over the ranks in rank order
// (bit-identical on every rank), stored according to `out`.  world == 1: no exchange.
// gather_seg >= 0: no rank sum; out.C[...] receives every rank's words (two-segment all-gather, see allgather2).
__global__ void __launch_bounds__(256) ll_reduce_kernel(LLArgs a, const double* __restrict__ src, int splits,
                                                        int64_t M, int64_t total, size_t cap, unsigned flag, int par,
                                                        ReduceOut out, int64_t gather_seg, int* error, int sg) {
  const int P = a.world, r = a.rank;
  // r02: `sg` (1, 2, 4 or 8) neighbouring lanes share one output word: each sums every sg-th split-K partial (two
  // chains), the group adds up by a fixed butterfly (same bits on every lane and rank), lane 0 of the group goes on.
  // With ~130 partials per word one thread per word was a chain of ~33 dependent L2 round trips (17 us).
  const int sub = threadIdx.x & (sg - 1);
  const int64_t gid0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / sg;
  const int64_t gstride = (int64_t)gridDim.x * blockDim.x / sg;
  const int64_t rounds = (total + gstride - 1) / gstride;  // every lane of a warp runs the same number of rounds
  for (int64_t it = 0; it < rounds; ++it) {
    const int64_t e = gid0 + it * gstride;
    const bool live = e < total;
    double s0 = 0.0, s1 = 0.0;
    if (live) {
      int z = sub;
      for (; z + sg < splits; z += 2 * sg) {
        s0 += src[(size_t)z * total + e];
        s1 += src[(size_t)(z + sg) * total + e];
      }
      if (z < splits) s0 += src[(size_t)z * total + e];
    }
    double v = s0 + s1;
    for (int o = sg >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (!live || sub != 0) continue;
    if (P > 1) {
      const size_t line = ((size_t)par * P + r) * cap + (size_t)e;
#pragma unroll 4
      for (int q = 1; q < P; ++q) ll_store(a.ll[(r + q) % P] + line, v, flag);  // peers staggered
      const uint4* mine = a.ll[r] + (size_t)par * P * cap + (size_t)e;
      if (gather_seg >= 0) {
        const int64_t count = total;
        for (int p = 0; p < P; ++p) {
          const double w = (p == r) ? v : ll_wait(mine + (size_t)p * cap, flag, error);
          const int64_t o = e < gather_seg ? (int64_t)p * gather_seg + e
                                           : (int64_t)P * gather_seg + (int64_t)p * (count - gather_seg) + (e - gather_seg);
          out.C[o] = w;
        }
        continue;
      }
      double t = 0.0;
      for (int p = 0; p < P; ++p) t += (p == r) ? v : ll_wait(mine + (size_t)p * cap, flag, error);
      v = t;
    }
    reduce_store(out, M, e, v);
  }
}

// gather, shared prologue / epilogue.  Prologue = barrier: nobody stores into a peer's block before that peer's
// stream has reached its own gather call (all its readers of the previous block are complete by stream order).
__device__ __forceinline__ void gather_prologue(const PeerArgs& a, unsigned long long epoch) {
  const int P = a.world, r = a.rank;
  PeerCtl* me = a.ctl[r];
  if (blockIdx.x == 0 && (int)threadIdx.x < P) st_release_sys(&a.ctl[threadIdx.x]->ag_ready[r], epoch);
  if ((int)threadIdx.x < P) wait_flag(&me->ag_ready[threadIdx.x], epoch, &me->error);
  __syncthreads();
}
__device__ __forceinline__ void gather_epilogue(const PeerArgs& a, unsigned long long epoch) {
  const int P = a.world, r = a.rank;
  PeerCtl* me = a.ctl[r];
  if (last_cta_arrives(&me->counter[1])) {
    if ((int)threadIdx.x < P) {
      __threadfence_system();
      st_release_sys(&a.ctl[threadIdx.x]->ag_done[r], epoch);
      wait_flag(&me->ag_done[threadIdx.x], epoch, &me->error);
    }
  }
}

// dst(row0 + i, j) on every rank <- X(i, j); dst column-major with leading dimension n
__global__ void __launch_bounds__(256) peer_gather_rows_kernel(PeerArgs a, const double* __restrict__ X, int64_t ldx,
                                                               int64_t nl, int64_t row0, int64_t n, int b,
                                                               unsigned long long epoch) {
  gather_prologue(a, epoch);
  const int P = a.world;
  const int64_t total = nl * b;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t j = e / nl, i = e - j * nl;
    const double v = X[i + j * ldx];
    const int64_t o = row0 + i + j * n;
#pragma unroll 4
    for (int p = 0; p < P; ++p) a.data[(a.rank + p) % P][o] = v;  // start with myself, peers staggered
  }
  gather_epilogue(a, epoch);
}

// the same into the packed fragment order (see comm.cuh); this rank owns rows [row0, row1) of the padded range
// [0, Kpad): row1 = row0 + nl, except that the last rank also zero-fills K..Kpad.  One thread per 16-byte pair
// (k, k+1) of the packed layout: coalesced 16-byte peer stores.
__global__ void __launch_bounds__(256)
    peer_gather_packed_kernel(PeerArgs a, const double* __restrict__ X, int64_t ldx, int64_t nl, int64_t row0,
                              int64_t row_end /* row0 + nl, or Kpad on the last rank */, int64_t Kpad, int b,
                              int nchunks, unsigned long long epoch) {
  gather_prologue(a, epoch);
  const int P = a.world;
  const int64_t kq0 = row0 >> 3, nkq = (row_end - row0 + 7) >> 3;  // row0 is a multiple of 8
  for (int c = 0; c < nchunks; ++c) {
    const int c0 = c * 128;
    const int bc = min(128, b - c0);
    const int bp = ((bc + 7) / 8) * 8;  // same rule as matvec_bpad()
    const int warps_n = bp <= 32 ? 1 : (bp <= 64 ? 2 : 4);
    int nt = (bp + 8 * warps_n - 1) / (8 * warps_n);
    if (warps_n > 1 && nt < 3) nt = 3;
    const int ntt = nt * warps_n;  // 8-column tiles of the chunk (bpad = 8 * ntt)
    const int64_t base = (int64_t)c0 * Kpad;
    // pairs: per k-group of 8 rows: ntt tiles x 8 columns x 4 row pairs
    const int64_t per_kq = (int64_t)ntt * 32;
    const int64_t total = nkq * per_kq;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total;
         e += (int64_t)gridDim.x * blockDim.x) {
      const int64_t kq = e / per_kq;
      const int rem = (int)(e - kq * per_kq);
      const int jt = rem >> 5, g = (rem >> 2) & 7, kp = rem & 3;
      const int j = c0 + jt * 8 + g;
      const int64_t k = (kq0 + kq) * 8 + 2 * kp;
      double2 v = make_double2(0.0, 0.0);
      if (j < b) {
        const int64_t i = k - row0;
        if (i < nl) v.x = X[i + (int64_t)j * ldx];
        if (i + 1 < nl) v.y = X[i + 1 + (int64_t)j * ldx];
      }
      const int64_t o = base + ((kq0 + kq) * ntt + jt) * 64 + g * 8 + 2 * kp;
#pragma unroll 4
      for (int p = 0; p < P; ++p) *reinterpret_cast<double2*>(a.data[(a.rank + p) % P] + o) = v;
    }
  }
  gather_epilogue(a, epoch);
}

// ---- self-checks (dav_debug_collective) --------------------------------------------------------------------------
__device__ __forceinline__ double dbg_value(int64_t row, int64_t col) {
  return (double)((row * 131 + col * 7919) % 1000003) * 1e-3 + 1.0;
}
__global__ void dbg_fill_rows_kernel(double* X, int64_t ld, int64_t nl, int64_t row0, int b) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < nl * b; e += (int64_t)gridDim.x * blockDim.x)
    X[e % nl + (e / nl) * ld] = dbg_value(row0 + e % nl, e / nl);
}
__global__ void dbg_fill_vec_kernel(double* x, int64_t count, int rank) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < count; e += (int64_t)gridDim.x * blockDim.x)
    x[e] = (double)(rank + 1) + 1e-3 * (double)(e % 4096);
}
// err[0] = max |got - expected| (atomicMax on the bit pattern of a non-negative double)
__device__ __forceinline__ void dbg_max(double* err, double d) {
  atomicMax((unsigned long long*)err, (unsigned long long)__double_as_longlong(fabs(d)));
}
__global__ void dbg_check_vec_kernel(const double* x, int64_t count, int world, double* err) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < count; e += (int64_t)gridDim.x * blockDim.x) {
    double s = 0.0;
    for (int r = 0; r < world; ++r) s += (double)(r + 1) + 1e-3 * (double)(e % 4096);
    dbg_max(err, x[e] - s);
  }
}
__global__ void dbg_check_full_kernel(const double* X, int64_t n, int b, double* err) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n * b; e += (int64_t)gridDim.x * blockDim.x)
    dbg_max(err, X[e] - dbg_value(e % n, e / n));
}
__global__ void dbg_check_packed_kernel(const double* Xp, int64_t n, int64_t Kpad, int b, double* err) {
  const int nchunks = (b + 127) / 128;
  for (int c = 0; c < nchunks; ++c) {
    const int c0 = c * 128, bc = min(128, b - c0);
    const int bp = ((bc + 7) / 8) * 8;
    const int warps_n = bp <= 32 ? 1 : (bp <= 64 ? 2 : 4);
    int nt = (bp + 8 * warps_n - 1) / (8 * warps_n);
    if (warps_n > 1 && nt < 3) nt = 3;
    const int ntt = nt * warps_n, bpad = ntt * 8;
    const int64_t total = Kpad * bpad;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
      const int kk = (int)(e & 7), g = (int)((e >> 3) & 7);
      const int64_t blk = e >> 6;
      const int jt = (int)(blk % ntt);
      const int64_t k = (blk / ntt) * 8 + kk;
      const int j = c0 + jt * 8 + g;
      const double expect = (k < n && j < b) ? dbg_value(k, j) : 0.0;
      dbg_max(err, Xp[(int64_t)c0 * Kpad + e] - expect);
    }
  }
}

}  // namespace

void Comm::debug_exchange(int kind, int64_t count, int reps, int64_t n, int64_t nl, int64_t row0, cudaStream_t s,
                          double* out) {
  if (world_ <= 1) DAV_THROW(DAV_ERR_STATE, "debug_exchange needs more than one rank");
  if (kind < 0 || kind > 3 || count < 1 || reps < 1) DAV_THROW(DAV_ERR_INVALID, "debug_exchange: bad arguments");
  DevBuf<double> x, err;
  err.alloc(1);
  CK(cudaMemsetAsync(err.p, 0, 8, s));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  SymBuf dst;
  if (kind <= 1) {
    reserve_allreduce((size_t)count, s);
    x.alloc((size_t)count);
    for (int r = 0; r <= reps; ++r) {  // r == 0: warm-up + correctness
      if (r == 1) CK(cudaEventRecord(e0, s));
      if (r <= 1) dbg_fill_vec_kernel<<<64, 256, 0, s>>>(x.p, count, rank_);
      if (kind == 0) allreduce_sum(x.p, (size_t)count, s);
      else nccl_allreduce(x.p, (size_t)count, s);
      if (r == 0) dbg_check_vec_kernel<<<64, 256, 0, s>>>(x.p, count, world_, err.p);
    }
  } else {
    if (!peer_) reserve_allreduce(1, s);
    if (!peer_) DAV_THROW(DAV_ERR_STATE, "debug_exchange: the peer transport is not available");
    const int b = (int)count;
    const int64_t Kpad = matvec_kpad(n);
    x.alloc((size_t)std::max<int64_t>(nl, 1) * b);
    sym_reserve(dst, std::max((size_t)n * b, matvec_packed_doubles(n, b)) * 8, s);
    dbg_fill_rows_kernel<<<128, 256, 0, s>>>(x.p, std::max<int64_t>(nl, 1), nl, row0, b);
    for (int r = 0; r <= reps; ++r) {
      if (r == 1) CK(cudaEventRecord(e0, s));
      if (kind == 2) gather_rows(x.p, std::max<int64_t>(nl, 1), nl, row0, n, b, dst, s);
      else gather_rows_packed(x.p, std::max<int64_t>(nl, 1), nl, row0, n, Kpad, b, dst, s);
      if (r == 0) {
        if (kind == 2) dbg_check_full_kernel<<<256, 256, 0, s>>>(dst.p(), n, b, err.p);
        else dbg_check_packed_kernel<<<256, 256, 0, s>>>(dst.p(), n, Kpad, b, err.p);
      }
    }
  }
  CK(cudaEventRecord(e1, s));
  CK(cudaStreamSynchronize(s));
  CK_LAUNCH();
  float ms = 0.f;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  double herr = 0.0;
  CK(cudaMemcpy(&herr, err.p, 8, cudaMemcpyDeviceToHost));
  int perr = 0;
  if (error_flag()) CK(cudaMemcpy(&perr, error_flag(), sizeof(int), cudaMemcpyDeviceToHost));
  if (dst.local) sym_release(dst);
  if (perr) DAV_THROW(DAV_ERR_COMM, "a wait inside a peer exchange kernel timed out");
  out[0] = 1e3 * (double)ms / reps;
  out[1] = herr;
}

int matvec_bpad(int bc) {
  int bp = (int)round_up(bc, 8);
  const int warps_n = bp <= 32 ? 1 : (bp <= 64 ? 2 : 4);
  int nt = (bp + 8 * warps_n - 1) / (8 * warps_n);
  if (warps_n > 1 && nt < 3) nt = 3;
  return nt * warps_n * 8;
}

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
  if (world > COMM_MAX_RANKS) DAV_THROW(DAV_ERR_INVALID, "at most %d ranks are supported", COMM_MAX_RANKS);
  Api& a = api();
  if (!a.ok) DAV_THROW(DAV_ERR_COMM, "NCCL unavailable: %s", a.why.c_str());
  NcclUniqueId id;
  std::memcpy(&id, id128, 128);
  ncclComm_t c = nullptr;
  check(a.CommInitRank(&c, world, id, rank), "ncclCommInitRank");
  comm_ = c;
}

Comm::~Comm() {
  if (peer_ || ctl_.local || slots_.local) {
    sym_release(slots_);
    sym_release(ctl_);
  }
  if (hbuf_) cudaFree(hbuf_);
  if (comm_) api().CommDestroy((ncclComm_t)comm_);
}

// ---- symmetric segments ----------------------------------------------------------------------------------------
// Collective.  Exchanges (handle, ok) of every rank with one ncclAllGather; the peer transport is used only if the
// mapping worked on EVERY rank (the decision is part of the exchanged data, so all ranks take the same branch).
void Comm::sym_reserve(SymBuf& b, size_t bytes, cudaStream_t s) {
  if (world_ <= 1) return;
  bytes = (size_t)round_up((int64_t)std::max<size_t>(bytes, 256), 256);
  if (b.local && b.bytes >= bytes) return;
  sym_release(b);
  struct Rec { cudaIpcMemHandle_t h; int ok; int pad[15]; };
  static_assert(sizeof(Rec) == 128, "record size");
  Rec mine;
  std::memset(&mine, 0, sizeof(mine));
  void* ptr = nullptr;
  cudaError_t e = cudaMalloc(&ptr, bytes);
  if (e != cudaSuccess) DAV_THROW(DAV_ERR_CUDA, "cudaMalloc of %zu bytes failed: %s", bytes, cudaGetErrorString(e));
  CK(cudaMemsetAsync(ptr, 0, bytes, s));
  mine.ok = cudaIpcGetMemHandle(&mine.h, ptr) == cudaSuccess;
  (void)cudaGetLastError();
  if (!hbuf_) CK(cudaMalloc(&hbuf_, sizeof(Rec) * (COMM_MAX_RANKS + 1)));
  Rec* dsend = (Rec*)hbuf_;
  Rec* drecv = dsend + 1;
  std::vector<Rec> all((size_t)world_);
  auto exchange = [&]() {
    CK(cudaMemcpyAsync(dsend, &mine, sizeof(Rec), cudaMemcpyHostToDevice, s));
    check(api().AllGather(dsend, drecv, sizeof(Rec), NCCL_UINT8, (ncclComm_t)comm_, s), "ncclAllGather");
    CK(cudaMemcpyAsync(all.data(), drecv, sizeof(Rec) * world_, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
  };
  exchange();
  bool ok = true;
  for (int r = 0; r < world_; ++r) ok = ok && all[(size_t)r].ok;
  b.local = ptr;
  b.bytes = bytes;
  b.peer[rank_] = ptr;
  if (ok) {
    for (int r = 0; r < world_ && ok; ++r) {
      if (r == rank_) continue;
      void* q = nullptr;
      if (cudaIpcOpenMemHandle(&q, all[(size_t)r].h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
        (void)cudaGetLastError();
        ok = false;
      } else {
        b.peer[r] = q;
      }
    }
  }
  // second round: did every rank manage to map every peer?
  mine.ok = ok;
  exchange();
  bool all_ok = true;
  for (int r = 0; r < world_; ++r) all_ok = all_ok && all[(size_t)r].ok;
  if (!all_ok) {
    for (int r = 0; r < world_; ++r)
      if (r != rank_ && b.peer[r]) {
        cudaIpcCloseMemHandle(b.peer[r]);
        b.peer[r] = nullptr;
      }
    DAV_THROW(DAV_ERR_COMM, "peer mapping (cudaIpc) of a %zu-byte segment failed on some rank", bytes);
  }
}

void Comm::sym_release(SymBuf& b) {
  if (!b.local) return;
  cudaDeviceSynchronize();
  for (int r = 0; r < world_; ++r)
    if (r != rank_ && b.peer[r]) cudaIpcCloseMemHandle(b.peer[r]);
  // nobody may unmap-and-free while a peer still has stores in flight towards this segment: every rank has
  // synchronised its device above; one small NCCL all-reduce is the barrier between "all closed" and "free"
  if (comm_ && hbuf_) {
    if (api().AllReduce(hbuf_, hbuf_, 1, NCCL_FLOAT64, NCCL_SUM, (ncclComm_t)comm_, (cudaStream_t)0) == 0)
      cudaDeviceSynchronize();
  }
  cudaFree(b.local);
  b = SymBuf();
}

void Comm::setup_peer(cudaStream_t s) {
  if (peer_tried_) return;
  peer_tried_ = true;
  peer_ = false;
  if (world_ <= 1) return;
  const char* env = std::getenv("DAV_PEER_COLLECTIVES");
  if (env && std::atoi(env) == 0) return;
  try {
    sym_reserve(ctl_, sizeof(PeerCtl), s);
    peer_ = true;
  } catch (const Error&) {
    if (env && std::atoi(env) == 1) throw;  // forced: fail loudly
    peer_ = false;
  }
}

PeerArgs Comm::args_for(const SymBuf& b) const {
  PeerArgs a;
  std::memset(&a, 0, sizeof(a));
  for (int r = 0; r < world_; ++r) {
    a.data[r] = (double*)b.peer[r];
    a.ctl[r] = (PeerCtl*)ctl_.peer[r];
  }
  a.rank = rank_;
  a.world = world_;
  return a;
}

void Comm::reserve_allreduce(size_t max_count, cudaStream_t s) {
  if (world_ <= 1) return;
  setup_peer(s);
  if (!peer_) return;
  if (max_count <= slot_cap_) return;
  const size_t cap = (size_t)round_up((int64_t)max_count, 32);
  sym_reserve(slots_, 2 * (size_t)world_ * cap * 16, s);  // 16-byte lines, zeroed: no line carries a live flag
  slot_cap_ = cap;
}

void Comm::nccl_allreduce(double* buf, size_t count, cudaStream_t s) {
  check(api().AllReduce(buf, buf, count, NCCL_FLOAT64, NCCL_SUM, (ncclComm_t)comm_, s), "ncclAllReduce");
  ++nccl_calls;
}

void Comm::launch_ll(const double* src, int splits, int64_t M, int64_t total, const ReduceOut& out, int64_t gather_seg,
                     bool exchange, cudaStream_t s) {
  LLArgs a;
  std::memset(&a, 0, sizeof(a));
  a.rank = rank_;
  a.world = (exchange && peer_) ? world_ : 1;
  unsigned flag = 0;
  int par = 0;
  if (a.world > 1) {
    for (int r = 0; r < world_; ++r) a.ll[r] = (uint4*)slots_.peer[r];
    ++ar_epoch_;
    if ((ar_epoch_ & 0xffffffffULL) == 0) ar_epoch_ += 2;  // flag 0 is the "empty line" pattern; keep the parity
    flag = (unsigned)(ar_epoch_ & 0xffffffffULL);
    par = (int)(ar_epoch_ & 1ULL);
    ++peer_calls;
  }
  const int sg = splits >= 32 ? 8 : (splits >= 8 ? 4 : (splits >= 2 ? 2 : 1));
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div(total * sg, 256), 1184));
  ll_reduce_kernel<<<grid, 256, 0, s>>>(a, src, splits, M, total, slot_cap_, flag, par, out, gather_seg,
                                        a.world > 1 ? &((PeerCtl*)ctl_.local)->error : nullptr, sg);
  CK_LAUNCH();
  ++g_kernel_launches;
}

void Comm::reduce_sum(const double* src, int splits, int64_t M, int64_t N, const ReduceOut& out, cudaStream_t s) {
  const int64_t total = M * N;
  if (total <= 0) return;
  if (world_ > 1 && (!peer_ || (size_t)total > slot_cap_)) {
    // NCCL transport: local split-K reduction into a contiguous block, ncclAllReduce, then the output layout
    nccl_tmp_.alloc((size_t)total);
    ReduceOut flat;
    flat.mode = 0;
    flat.C = nccl_tmp_.p;
    flat.ldc = M;
    flat.kold = 0;
    launch_ll(src, splits, M, total, flat, -1, false, s);
    nccl_allreduce(nccl_tmp_.p, (size_t)total, s);
    launch_ll(nccl_tmp_.p, 1, M, total, out, -1, false, s);
    return;
  }
  launch_ll(src, splits, M, total, out, -1, true, s);
}

void Comm::allreduce_sum(double* buf, size_t count, cudaStream_t s) {
  if (world_ <= 1 || count == 0) return;
  if (!peer_ || count > slot_cap_) {
    nccl_allreduce(buf, count, s);
    return;
  }
  ReduceOut out;
  out.mode = 0;
  out.C = buf;
  out.ldc = (int64_t)count;
  out.kold = 0;
  launch_ll(buf, 1, (int64_t)count, (int64_t)count, out, -1, true, s);  // in place: a thread reads its word, then writes it
}

void Comm::allgather(const void* send, void* recv, size_t bytes_per_rank, cudaStream_t s) {
  allgather2(send, recv, bytes_per_rank, 0, s);
}

void Comm::allgather2(const void* send, void* recv, size_t bytes_a, size_t bytes_b, cudaStream_t s) {
  const size_t bytes = bytes_a + bytes_b;
  if (world_ <= 1) {
    if (send != recv) CK(cudaMemcpyAsync(recv, send, bytes, cudaMemcpyDeviceToDevice, s));
    return;
  }
  if (peer_ && bytes_a % 8 == 0 && bytes_b % 8 == 0 && bytes / 8 <= slot_cap_ && ((uintptr_t)send & 7) == 0 &&
      ((uintptr_t)recv & 7) == 0) {
    // small payloads ride on the one-shot exchange (8-byte words, no arithmetic on them)
    ReduceOut out;
    out.mode = 2;
    out.C = (double*)recv;
    out.ldc = 0;
    out.kold = 0;
    launch_ll((const double*)send, 1, (int64_t)(bytes / 8), (int64_t)(bytes / 8), out, (int64_t)(bytes_a / 8), true, s);
    return;
  }
  check(api().AllGather(send, recv, bytes_a, NCCL_UINT8, (ncclComm_t)comm_, s), "ncclAllGather");
  ++nccl_calls;
  if (bytes_b) {
    check(api().AllGather((const char*)send + bytes_a, (char*)recv + bytes_a * world_, bytes_b, NCCL_UINT8,
                          (ncclComm_t)comm_, s), "ncclAllGather");
    ++nccl_calls;
  }
}

void Comm::gather_rows(const double* Xlocal, int64_t ldx, int64_t nl, int64_t row0, int64_t n, int b, SymBuf& dst,
                       cudaStream_t s) {
  if (!peer_ || dst.bytes < (size_t)n * b * 8) DAV_THROW(DAV_ERR_STATE, "gather_rows: symmetric block not reserved");
  ++ag_epoch_;
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div(nl * b, 1024), 128));
  peer_gather_rows_kernel<<<grid, 256, 0, s>>>(args_for(dst), Xlocal, ldx, nl, row0, n, b, ag_epoch_);
  CK_LAUNCH();
  ++g_kernel_launches;
  ++peer_calls;
}

void Comm::gather_rows_packed(const double* Xlocal, int64_t ldx, int64_t nl, int64_t row0, int64_t n, int64_t Kpad,
                              int b, SymBuf& dst, cudaStream_t s) {
  const int nchunks = (b + 127) / 128;
  const size_t need = ((size_t)(nchunks - 1) * 128 + (size_t)matvec_bpad(b - (nchunks - 1) * 128)) * (size_t)Kpad * 8;
  if (!peer_ || dst.bytes < need) DAV_THROW(DAV_ERR_STATE, "gather_rows_packed: symmetric block not reserved");
  if (row0 % 8 != 0) DAV_THROW(DAV_ERR_STATE, "gather_rows_packed: row blocks must start at a multiple of 8");
  ++ag_epoch_;
  const bool last = row0 + nl >= n;
  const int64_t row_end = last ? Kpad : row0 + nl;
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div((row_end - row0) * matvec_bpad(std::min(b, 128)),
                                                                         2048), 128));
  peer_gather_packed_kernel<<<grid, 256, 0, s>>>(args_for(dst), Xlocal, ldx, nl, row0, row_end, Kpad, b, nchunks,
                                                 ag_epoch_);
  CK_LAUNCH();
  ++g_kernel_launches;
  ++peer_calls;
}

}  // namespace dav
