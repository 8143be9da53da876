// Microbenchmark: are DFMA (FP64 SIMT pipe) and DMMA (mma.sync.m8n8k4.f64) separate pipes on sm_100a?
// Runs DMMA-only, DFMA-only and interleaved mixes, reports TFLOP/s each.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_pipes fp64_pipes.cu && ./fp64_pipes
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

template <int NM, int NF>
__global__ void __launch_bounds__(256) k(int iters, double* out, double a, double b) {
  double m[8][2], f[16];
#pragma unroll
  for (int i = 0; i < 8; ++i) m[i][0] = m[i][1] = threadIdx.x;
#pragma unroll
  for (int i = 0; i < 16; ++i) f[i] = threadIdx.x + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (i < NM) dmma(m[i][0], m[i][1], a, b);
#pragma unroll
      for (int j = 0; j < 2; ++j)
        if (i * 2 + j < NF) asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(f[i * 2 + j]) : "d"(a), "d"(b));
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += m[i][0] + m[i][1];
#pragma unroll
  for (int i = 0; i < 16; ++i) s += f[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NM, int NF>
void run(const char* name) {
  double* out;
  cudaMalloc(&out, 148 * 8 * 256 * 8);
  const int iters = 20000, grid = 148 * 4;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<NM, NF><<<grid, 256>>>(100, out, 1.0, 1e-9);
  cudaEventRecord(e0);
  k<NM, NF><<<grid, 256>>>(iters, out, 1.0, 1e-9);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double warps = (double)grid * 8;
  const double fl_m = warps * iters * NM * 512.0, fl_f = warps * 32 * iters * NF * 2.0;
  printf("%-28s %8.3f ms  dmma %6.2f TF  dfma %6.2f TF  total %6.2f TF\n", name, ms, fl_m / ms * 1e-9, fl_f / ms * 1e-9,
         (fl_m + fl_f) / ms * 1e-9);
  cudaFree(out);
}

int main() {
  run<8, 0>("dmma only (8/iter)");
  run<0, 16>("dfma only (16/iter)");
  run<8, 16>("dmma 8 + dfma 16");
  run<8, 8>("dmma 8 + dfma 8");
  run<8, 4>("dmma 8 + dfma 4");
  run<4, 16>("dmma 4 + dfma 16");
  return 0;
}
