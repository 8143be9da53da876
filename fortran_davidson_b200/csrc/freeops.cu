// Matrix-free path: the reference's on-the-fly operators (benchmark_free.f90:38-76,
// tests/test_utils.f90:37-116) applied as free_matmul does (davidson.f90:526-569):
//   out(i, j) = sum_l a(l, i) * X(l, j),  a(l, i) = f(atan2(e_min(i,l), e_max(i,l))) * 1e-4 (+ diagonal)
// with e_t = dble(expf(real(t)/real(dim))) taken from a host-built table (glibc expf, identical
// bits to the oracle), f = cos(log(sqrt(.))) or sin(log(sqrt(.))).
// Entries are generated in registers, staged as a 64x64 tile in shared memory and consumed by a
// register-tiled mini GEMM; nothing n x n ever exists.  Bound: FP64 transcendental throughput.
#include <algorithm>

#include "kernels.cuh"

namespace dav {
namespace {

constexpr int FT = 64;  // tile edge

__device__ __forceinline__ double op_entry(int op, int64_t gi, int64_t gl, const double* __restrict__ etab) {
  // gi, gl 0-based
  const double scale = (double)1e-4f;  // single precision literal of the reference
  const double ei = etab[gi], el = etab[gl];
  const double lo = gi <= gl ? ei : el, hi = gi <= gl ? el : ei;
  const double a = atan2(lo, hi);
  const double l = log(sqrt(a));
  if (op == DAV_OP_TEST_STX) return (gi == gl) ? 1.0 : sin(l) * scale;
  double v = cos(l) * scale;
  if (gi == gl) v += (double)(float)(gi + 1);
  return v;
}

// grid.x = row tiles of 64; 256 threads.  acc[c]: row r = tid % 64, columns jq + 4c.
template <int NACC>
__global__ void __launch_bounds__(256) free_matmul_kernel(int op, int64_t n, int64_t row0, int64_t nl, int b, int j0,
                                                          const double* __restrict__ etab,
                                                          const double* __restrict__ X, int64_t ldx,
                                                          double* __restrict__ W, int64_t ldw) {
  __shared__ double tile[FT][FT + 1];  // tile[l][r]
  const int tid = threadIdx.x;
  const int r = tid % FT, jq = tid / FT;
  const int64_t i0 = (int64_t)blockIdx.x * FT;
  double acc[NACC];
#pragma unroll
  for (int c = 0; c < NACC; ++c) acc[c] = 0.0;
  for (int64_t l0 = 0; l0 < n; l0 += FT) {
    // generate: thread -> rows r, l = jq + 4*t
#pragma unroll 4
    for (int t = 0; t < FT / 4; ++t) {
      const int l = jq + 4 * t;
      const int64_t gi = row0 + i0 + r, gl = l0 + l;
      tile[l][r] = (i0 + r < nl && gl < n) ? op_entry(op, gi, gl, etab) : 0.0;
    }
    __syncthreads();
    const int lmax = (int)min((int64_t)FT, n - l0);
    for (int l = 0; l < lmax; ++l) {
      const double a = tile[l][r];
      const double* xrow = X + l0 + l;
#pragma unroll
      for (int c = 0; c < NACC; ++c) {
        const int j = j0 + jq + 4 * c;
        if (j < b) acc[c] = fma(a, __ldg(xrow + (int64_t)j * ldx), acc[c]);
      }
    }
    __syncthreads();
  }
  if (i0 + r < nl) {
#pragma unroll
    for (int c = 0; c < NACC; ++c) {
      const int j = j0 + jq + 4 * c;
      if (j < b) W[i0 + r + (int64_t)j * ldw] = acc[c];
    }
  }
}

__global__ void copy_rows_kernel(int64_t row0, int64_t nl, int b, const double* __restrict__ X, int64_t ldx,
                                 double* __restrict__ W, int64_t ldw) {
  for (int j = blockIdx.y; j < b; j += gridDim.y)
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nl; i += (int64_t)gridDim.x * blockDim.x)
      W[i + (int64_t)j * ldw] = X[row0 + i + (int64_t)j * ldx];
}

__global__ void free_diag_kernel(int op, int64_t n, int64_t row0, int64_t nl, const double* __restrict__ etab,
                                 double* __restrict__ diag) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nl; i += (int64_t)gridDim.x * blockDim.x)
    diag[i] = (op == DAV_OP_IDENTITY) ? 1.0 : op_entry(op, row0 + i, row0 + i, etab);
}

__global__ void free_column_kernel(int op, int64_t n, int64_t col, const double* __restrict__ etab,
                                   double* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = (op == DAV_OP_IDENTITY) ? (i == col ? 1.0 : 0.0) : op_entry(op, i, col, etab);
}

__global__ void free_gather_columns_kernel(int op, int64_t n, int64_t row0, int64_t nl,
                                           const double* __restrict__ etab, const int64_t* __restrict__ idx,
                                           double* __restrict__ out, int64_t ldo) {
  const int64_t col = idx[blockIdx.y];
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nl; i += (int64_t)gridDim.x * blockDim.x)
    out[i + (int64_t)blockIdx.y * ldo] =
        (op == DAV_OP_IDENTITY) ? (row0 + i == col ? 1.0 : 0.0) : op_entry(op, row0 + i, col, etab);
}

}  // namespace

void free_gather_columns_builtin(cudaStream_t s, int op, int64_t n, int64_t row0, int64_t nl, const double* etab,
                                 const int64_t* idx, int k, double* out, int64_t ldo) {
  if (nl <= 0 || k <= 0) return;
  dim3 grid((unsigned)std::max<int64_t>(1, std::min<int64_t>(ceil_div(nl, 256), 592)), (unsigned)k);
  free_gather_columns_kernel<<<grid, 256, 0, s>>>(op, n, row0, nl, etab, idx, out, ldo);
  CK_LAUNCH();
  ++g_kernel_launches;
}

void free_matmul_builtin(cudaStream_t s, int op, int64_t n, int64_t row0, int64_t nl, int b, const double* etab,
                         const double* X, int64_t ldx, double* W, int64_t ldw) {
  if (nl <= 0 || b <= 0) return;
  if (op == DAV_OP_IDENTITY) {
    dim3 grid((unsigned)std::max<int64_t>(1, std::min<int64_t>(ceil_div(nl, 256), 592)), (unsigned)std::min(b, 65535));
    copy_rows_kernel<<<grid, 256, 0, s>>>(row0, nl, b, X, ldx, W, ldw);
    CK_LAUNCH();
    ++g_kernel_launches;
    return;
  }
  const unsigned gx = (unsigned)ceil_div(nl, FT);
  // column chunks of up to 64 (16 accumulators x 4 column groups); entries are regenerated per chunk
  for (int j0 = 0; j0 < b; j0 += 64) {
    const int w = std::min(64, b - j0);
    if (w <= 16)
      free_matmul_kernel<4><<<gx, 256, 0, s>>>(op, n, row0, nl, b, j0, etab, X, ldx, W, ldw);
    else if (w <= 32)
      free_matmul_kernel<8><<<gx, 256, 0, s>>>(op, n, row0, nl, b, j0, etab, X, ldx, W, ldw);
    else
      free_matmul_kernel<16><<<gx, 256, 0, s>>>(op, n, row0, nl, b, j0, etab, X, ldx, W, ldw);
    CK_LAUNCH();
    ++g_kernel_launches;
  }
}

void free_diag_builtin(cudaStream_t s, int op, int64_t n, int64_t row0, int64_t nl, const double* etab,
                       double* diag) {
  if (nl <= 0) return;
  free_diag_kernel<<<(unsigned)std::max<int64_t>(1, std::min<int64_t>(ceil_div(nl, 256), 1184)), 256, 0, s>>>(
      op, n, row0, nl, etab, diag);
  CK_LAUNCH();
  ++g_kernel_launches;
}

void free_column_builtin(cudaStream_t s, int op, int64_t n, int64_t col, const double* etab, double* out) {
  free_column_kernel<<<(unsigned)std::max<int64_t>(1, std::min<int64_t>(ceil_div(n, 256), 1184)), 256, 0, s>>>(
      op, n, col, etab, out);
  CK_LAUNCH();
  ++g_kernel_launches;
}

}  // namespace dav
