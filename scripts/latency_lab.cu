// r02 latency lab: what one barrier-separated step of a single-CTA FP64 kernel really costs on B200, by pattern.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/latency_lab scripts/latency_lab.cu && /tmp/latency_lab
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 2000

template <int PATTERN>
__global__ void __launch_bounds__(1024) lab(double* out, long long* cyc, double seed) {
  __shared__ __align__(16) double s[1024];
  __shared__ double piv;
  const int tid = threadIdx.x, nt = blockDim.x;
  s[tid] = seed + tid * 1e-3;
  if (tid == 0) piv = seed;
  __syncthreads();
  double x = seed * 0.5, y = 1.0 + tid * 1e-6;
  long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
    if (PATTERN == 0) {  // barrier only
      __syncthreads();
    } else if (PATTERN == 1) {  // barrier + LDS + FMA + STS
      __syncthreads();
      x = fma(s[(tid + it) & (nt - 1)], y, x);
      __syncthreads();
      s[tid] = x;
    } else if (PATTERN == 2) {  // barrier, everyone reads what ONE thread (moving owner) computed with rsqrt
      __syncthreads();
      const double p = piv;
      x = fma(p, y, x);
      __syncthreads();
      if (tid == (it & (nt - 1))) piv = rsqrt(fabs(x) + 1.5);
    } else if (PATTERN == 3) {  // same with a plain FMA instead of rsqrt
      __syncthreads();
      const double p = piv;
      x = fma(p, y, x);
      __syncthreads();
      if (tid == (it & (nt - 1))) piv = fma(x, 0.5, 1.5);
    } else if (PATTERN == 4) {  // 16-lane all-reduce (4 double shuffles), no barrier
      double v = x;
#pragma unroll
      for (int m = 8; m >= 1; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
      x = v * 0.0625;
    } else if (PATTERN == 5) {  // dependent rsqrt + drcp chain, no barrier
      x = rsqrt(fabs(x) + 1.5);
      x = __drcp_rn(x + 2.0);
    } else if (PATTERN == 6) {  // Cholesky-like step: barrier; LDS x2; 16 FMAs (2 deep); owner: rsqrt + STS
      __syncthreads();
      const double rb = s[(tid * 4 + it) & (nt - 1)], ri = piv;
      double acc[16];
#pragma unroll
      for (int q = 0; q < 16; ++q) acc[q] = fma(rb * ri, y + q, x);
#pragma unroll
      for (int q = 0; q < 16; ++q) x += acc[q] * 1e-30;
      if (tid == (it & (nt - 1))) piv = rsqrt(fabs(x) + 1.5);
      if ((tid >> 4) == (it & 15)) s[(tid + it) & (nt - 1)] = x;
    } else if (PATTERN == 7) {  // one barrier per step, owner WARP does a 16-lane all-reduce + rsqrt, others 32 FMAs
      __syncthreads();
      const double p = piv;
      if ((tid >> 5) == (it & ((nt >> 5) - 1))) {
        double v = x * p;
#pragma unroll
        for (int m = 8; m >= 1; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
        if ((tid & 31) == 0) piv = rsqrt(fabs(v) + 1.5);
      } else {
        double a0 = x, a1 = y;
#pragma unroll
        for (int q = 0; q < 16; ++q) { a0 = fma(a0, p, 1e-3); a1 = fma(a1, p, 1e-3); }
        x = a0 + a1 * 1e-30;
      }
    }
  }
  long long t1 = clock64();
  if (tid == 0) cyc[0] = (t1 - t0) / ITERS;
  out[tid] = x;
}

template <int P>
void run(const char* name) {
  double* out; long long* cyc;
  cudaMalloc(&out, 1024 * 8); cudaMalloc(&cyc, 8);
  for (int nt : {64, 256, 1024}) {
    lab<P><<<1, nt>>>(out, cyc, 1.25);
    lab<P><<<1, nt>>>(out, cyc, 1.25);
    cudaDeviceSynchronize();
    long long h = 0; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-78s threads %4d: %5lld cycles / step\n", name, nt, h);
  }
  cudaFree(out); cudaFree(cyc);
}

int main() {
  run<0>("0 barrier only");
  run<1>("1 barrier + LDS + FMA, barrier + STS");
  run<2>("2 barrier, all read piv + FMA, barrier, ONE moving owner: rsqrt -> piv");
  run<3>("3 same with FMA instead of rsqrt");
  run<4>("4 16-lane double all-reduce (4 shuffle stages), no barrier");
  run<5>("5 rsqrt + drcp dependent chain, no barrier");
  run<6>("6 Cholesky-like step: barrier, 2 LDS, 16 FMAs, owner rsqrt + STS");
  run<7>("7 one barrier, owner warp: 16-lane all-reduce + rsqrt; others 32 dependent FMAs");
  return 0;
}
