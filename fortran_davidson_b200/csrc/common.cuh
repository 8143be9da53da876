// Shared host/device helpers of the B200 Davidson library.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>
#include <string>

#include "../../include/davidson_b200.h"

namespace dav {

struct Error {
  int code;
  std::string msg;
};

void set_last_error(const std::string& s);

#define DAV_THROW(code_, ...)                                   \
  do {                                                          \
    char _b[512];                                               \
    std::snprintf(_b, sizeof(_b), __VA_ARGS__);                 \
    throw ::dav::Error{(code_), std::string(_b)};               \
  } while (0)

#define CK(call)                                                                                           \
  do {                                                                                                     \
    cudaError_t _e = (call);                                                                               \
    if (_e != cudaSuccess)                                                                                 \
      DAV_THROW(DAV_ERR_CUDA, "CUDA error %s at %s:%d: %s", cudaGetErrorName(_e), __FILE__, __LINE__,      \
                cudaGetErrorString(_e));                                                                   \
  } while (0)

#define CK_LAUNCH() CK(cudaGetLastError())

// Per-DEVICE facts and one-time settings (a process may drive several GPUs: dav_create(h, device)).
// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is a per-device setting of a kernel, so the "already done" mark
// is kept per (kernel, device); both helpers are thread safe.
int device_max_smem_optin();  // cudaDevAttrMaxSharedMemoryPerBlockOptin of the current device (cached per device)
int device_num_sms();         // SM count of the current device (cached per device)
void ensure_dyn_smem_impl(const void* func, int bytes);
template <typename F>
inline void ensure_dyn_smem(F* kernel, int bytes) { ensure_dyn_smem_impl((const void*)kernel, bytes); }

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
inline int64_t round_up(int64_t a, int64_t b) { return ceil_div(a, b) * b; }

// RAII device buffer of doubles (or raw bytes)
template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  DevBuf() {}
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  ~DevBuf() { release(); }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    n = 0;
  }
  void alloc(size_t count) {
    if (count <= n && p) return;
    release();
    if (count == 0) return;
    cudaError_t e = cudaMalloc((void**)&p, count * sizeof(T));
    if (e != cudaSuccess) {
      p = nullptr;
      DAV_THROW(DAV_ERR_CUDA, "cudaMalloc of %zu bytes failed: %s", count * sizeof(T), cudaGetErrorString(e));
    }
    n = count;
  }
};

#ifdef __CUDACC__
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Counter-based uniform stream shared bit-for-bit with the oracle (oracle/davidson_oracle.cpp
// orc_uniform01): keyed on (seed, min(i,j), max(i,j)) so every shard regenerates identical entries.
__host__ __device__ __forceinline__ uint64_t mix64(uint64_t z) {
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}
__host__ __device__ __forceinline__ double uniform01(uint64_t seed, uint64_t lo, uint64_t hi) {
  uint64_t h = mix64(seed + 0x9E3779B97F4A7C15ULL * (lo + 1));
  h = mix64(h ^ (0xD6E8FEB86659FD93ULL * (hi + 1)));
  return (double)(h >> 11) * (1.0 / 9007199254740992.0);
}
#endif

}  // namespace dav
