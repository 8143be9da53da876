// VERDICT r1 item 9, GPU half of the gate: what does an Ozaki-style INT8 emulation of the FP64 block matvec cost on a
// B200?  (1) IMMA throughput through mma.sync m16n8k32 s8 (the tensor path this library's kernels already use for
// FP64; tcgen05 kind::i8 would be faster still), register only; (2) the cost of splitting FP64 entries into five
// signed 7-bit slices with integer instructions (what the A tile needs once per matvec), with and without the
// write-back; (3) a host check that the slices reproduce the truncated fixed-point value exactly.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ozaki_lab scripts/ozaki_lab.cu && /tmp/ozaki_lab
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

#define NSL 5
#define WBITS 7

__global__ void __launch_bounds__(256) imma_peak(int iters, int* out) {
  int acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = threadIdx.x + i + j;
  unsigned a0 = threadIdx.x * 0x01010101u, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, b0 = a0 ^ 0x55555555u, b1 = ~a0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+r"(acc[i][0]), "+r"(acc[i][1]), "+r"(acc[i][2]), "+r"(acc[i][3])
                   : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  }
  int s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) s += acc[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// five signed 7-bit digits of trunc(a * 2^(35 - e)), most significant first, |a| < 2^(e-1)
__device__ __host__ inline void split5(double a, int e, int q[NSL]) {
  union { double d; unsigned long long u; } v;
  v.d = a;
  const unsigned long long bits = v.u;
  const int ex = (int)((bits >> 52) & 0x7ff);
  unsigned long long mant = bits & 0xFFFFFFFFFFFFFull;
  if (ex != 0) mant |= 1ull << 52;
  const int sh = (1075 + e - NSL * WBITS) - (ex == 0 ? 1 : ex);
  const unsigned long long F = sh >= 64 ? 0ull : (mant >> sh);  // sh >= 19 by the bound on |a|
  const int sgn = (bits >> 63) ? -1 : 1;
#pragma unroll
  for (int s = 0; s < NSL; ++s) q[s] = sgn * (int)((F >> (WBITS * (NSL - 1 - s))) & 127);
}

// four consecutive entries of a row -> one packed word (4 x s8) per slice
template <bool STORE>
__global__ void __launch_bounds__(256) split_rate(const double* __restrict__ A, size_t n4, const int* __restrict__ rowexp,
                                                  unsigned* __restrict__ out, unsigned* __restrict__ sink) {
  unsigned x = 0;
  for (size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x; g < n4; g += (size_t)gridDim.x * blockDim.x) {
    const double2 p0 = reinterpret_cast<const double2*>(A)[2 * g], p1 = reinterpret_cast<const double2*>(A)[2 * g + 1];
    const int e = rowexp[g & 4095];
    int q[4][NSL];
    split5(p0.x, e, q[0]); split5(p0.y, e, q[1]); split5(p1.x, e, q[2]); split5(p1.y, e, q[3]);
#pragma unroll
    for (int s = 0; s < NSL; ++s) {
      const unsigned w = (unsigned)(q[0][s] & 0xff) | ((unsigned)(q[1][s] & 0xff) << 8) | ((unsigned)(q[2][s] & 0xff) << 16) |
                         ((unsigned)(q[3][s] & 0xff) << 24);
      if (STORE) out[(size_t)s * n4 + g] = w;
      else x ^= w;
    }
  }
  if (!STORE) sink[blockIdx.x * blockDim.x + threadIdx.x] = x;
}

int main() {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  float ms;
  {  // (1) IMMA peak
    int* out; cudaMalloc(&out, 148 * 8 * 256 * 4);
    const int iters = 20000, grid = 148 * 8;
    imma_peak<<<grid, 256>>>(100, out);
    cudaEventRecord(e0); imma_peak<<<grid, 256>>>(iters, out); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    const double ops = (double)grid * 8 * iters * 8 * (2.0 * 16 * 8 * 32);
    printf("IMMA mma.sync m16n8k32 s8: %.3f ms  %.1f TOPS (dense INT8 peak of the chip through tcgen05: 4500)\n", ms, ops / ms * 1e-9);
    cudaFree(out);
  }
  {  // (2) split rate on 2 GiB of doubles
    const size_t n = (size_t)1 << 28, n4 = n / 4;
    double* A; cudaMalloc(&A, n * 8);
    unsigned *out, *sink; cudaMalloc(&out, NSL * n4 * 4); cudaMalloc(&sink, 148 * 8 * 256 * 4);
    std::vector<double> h(1 << 20);
    srand(1);
    for (auto& v : h) v = (rand() / (double)RAND_MAX - 0.3) * 1e-4;
    for (size_t o = 0; o < n; o += h.size()) cudaMemcpy(A + o, h.data(), h.size() * 8, cudaMemcpyHostToDevice);
    std::vector<int> re(4096, -12);  // |a| <= 0.7e-4 < 2^(-13): e = -12
    int* rowexp; cudaMalloc(&rowexp, 4096 * 4); cudaMemcpy(rowexp, re.data(), 4096 * 4, cudaMemcpyHostToDevice);
    const int grid = 148 * 8;
    for (int store = 0; store < 2; ++store) {
      for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        if (store) split_rate<true><<<grid, 256>>>(A, n4, rowexp, out, sink);
        else split_rate<false><<<grid, 256>>>(A, n4, rowexp, out, sink);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
      }
      cudaEventElapsedTime(&ms, e0, e1);
      printf("split 5 x s8 slices (%s): %.3f ms for %zu entries = %.1f G entries/s, %.0f GB/s of FP64 read%s\n",
             store ? "slices written back" : "slices consumed in registers", ms, n, n / ms * 1e-6, n * 8.0 / ms * 1e-6,
             store ? " + 5/8 of it written" : "");
    }
    // (3) exactness of the digits
    double worst = 0;
    for (int i = 0; i < 100000; ++i) {
      int q[NSL];
      split5(h[i], -12, q);
      double rec = 0;
      for (int s = 0; s < NSL; ++s) rec += q[s] * std::ldexp(1.0, -WBITS * (s + 1));
      rec = std::ldexp(rec, -12);
      worst = std::fmax(worst, std::fabs(rec - h[i]) / std::ldexp(1.0, -12));
      for (int s = 0; s < NSL; ++s) if (q[s] < -127 || q[s] > 127) { printf("digit out of range\n"); return 1; }
    }
    printf("reconstruction: max |sum_s q_s 2^(e - 7(s+1)) - a| / 2^e = %.3e (truncation bound 2^-35 = %.3e)\n", worst, std::ldexp(1.0, -35));
  }
  return 0;
}
