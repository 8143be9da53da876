// Microbenchmark: dependent-issue latency (cycles) of FP64 ops on one warp / one thread on sm_100a.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_latency fp64_latency.cu && ./fp64_latency
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double* out, long long* cyc, double a, double b, int n) {
  __shared__ double sm[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = 1.0 + 1e-9 * i;
  __syncthreads();
  double x = a + threadIdx.x;
  long long t0, t1;
  int slot = 0;
#define MEAS(body)                                  \
  t0 = clock64();                                   \
  for (int i = 0; i < n; ++i) { body; }             \
  t1 = clock64();                                   \
  if (threadIdx.x == 0) cyc[slot] = (t1 - t0) / n;  \
  ++slot;
  MEAS(x = fma(x, b, a))                                    // 0 DFMA
  MEAS(x = x + a)                                           // 1 DADD
  MEAS(x = __shfl_xor_sync(0xffffffffu, x, 1) + a)          // 2 SHFL64 + DADD
  MEAS(x = rsqrt(x * x + 1.0))                              // 3 rsqrt (+fma)
  MEAS(x = __drcp_rn(x + 1.5))                              // 4 drcp (+add)
  MEAS(x = a / (x + 1.5))                                   // 5 ddiv (+add)
  MEAS(x = sqrt(x * x + 1.5))                               // 6 sqrt (+fma)
  MEAS(x = sm[((int)x) & 1023] + a)                         // 7 LDS + convert + DADD
  MEAS(__syncthreads(); x += 1.0)                           // 8 bar + DADD
  float y = (float)x;
  MEAS(y = fmaf(y, 1.0001f, 0.5f))                          // 9 FFMA
  out[threadIdx.x] = x + y;
}
int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 1024 * 8); cudaMallocManaged(&cyc, 16 * 8);
  const char* names[] = {"DFMA", "DADD", "SHFL64+DADD", "rsqrt+DFMA", "drcp+DADD", "ddiv+DADD", "sqrt+DFMA", "LDS+cvt+DADD", "bar.sync+DADD", "FFMA"};
  for (int threads : {32, 256, 1024}) {
    k<<<1, threads>>>(out, cyc, 1.0, 0.999, 2000);
    cudaDeviceSynchronize();
    printf("threads %4d:", threads);
    for (int i = 0; i < 10; ++i) printf("  %s %lld", names[i], cyc[i]);
    printf("\n");
  }
  return 0;
}
