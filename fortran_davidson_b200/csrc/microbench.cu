// Live denominator of the FP64 roofline: the DMMA (mma.sync.m8n8k4.f64) issue rate of the whole chip, measured with a
// register-only kernel (no memory traffic), so bench.py can state the block matvec as a fraction of what the FP64
// tensor pipe can do on THIS device instead of only against cuBLAS DGEMM.  (DMMA and DFMA share one pipe on B200:
// profiles/r01_fp64_pipes_microbench.txt, scripts/fp64_pipes.cu.)
#include "kernels.cuh"

namespace dav {
namespace {

__global__ void __launch_bounds__(256) dmma_peak_kernel(int iters, double* out, double a, double b) {
  double m[8][2];
#pragma unroll
  for (int i = 0; i < 8; ++i) m[i][0] = m[i][1] = threadIdx.x;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i)  // 8 independent accumulator chains per warp
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(m[i][0]), "+d"(m[i][1])
                   : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += m[i][0] + m[i][1];
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

}  // namespace

// TFLOP/s of back-to-back DMMA.8x8x4 on every SM (4 CTAs of 8 warps per SM), best of `reps` timed launches.
double dmma_peak_tflops(cudaStream_t s, int reps) {
  int dev = 0, sms = 0;
  CK(cudaGetDevice(&dev));
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  if (sms <= 0) sms = 148;
  const int grid = sms * 4, iters = 8000;
  DevBuf<double> out;
  out.alloc((size_t)grid * 256);
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  dmma_peak_kernel<<<grid, 256, 0, s>>>(200, out.p, 1.0, 1e-9);  // warm-up
  CK_LAUNCH();
  double best = 0.0;
  for (int r = 0; r < reps; ++r) {
    CK(cudaEventRecord(e0, s));
    dmma_peak_kernel<<<grid, 256, 0, s>>>(iters, out.p, 1.0, 1e-9);
    CK_LAUNCH();
    CK(cudaEventRecord(e1, s));
    CK(cudaEventSynchronize(e1));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    const double flops = (double)grid * 8 /*warps*/ * iters * 8 /*DMMA per iteration*/ * 512.0;
    if (ms > 0.f) best = std::max(best, flops / ms * 1e-9);
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  return best;
}

}  // namespace dav
