// Device versions of the remaining lapack_wrapper helpers (lapack_wrapper.f90:238-277 DSYSV, :367-392 DLASRT + keys,
// :279-328 DGEMM with op(B) = B^T): r01 ran parts of them in host loops.
//   lu_solve       n x n system(s) by blocked LU with partial pivoting (panel of 32 columns on one CTA, row swaps,
//                  32 x 32 unit-lower solves, trailing update on the tensor-pipe GEMM), then back substitution
//   sort_pairs     full ascending sort of (key, index) pairs: bitonic network in global memory
//   transpose      out(c, r) = in(r, c)
#include <algorithm>

#include "kernels.cuh"

namespace dav {
namespace {

constexpr int LU_NB = 32;

// panel columns [j0, j0 + jb) of A (n rows): unblocked LU with partial pivoting; piv[c] = pivot row of column c;
// status |= 2 when a pivot column is exactly zero (singular matrix) or holds a NaN
__global__ void __launch_bounds__(1024) lu_panel_kernel(int n, int j0, int jb, double* __restrict__ A, int64_t lda,
                                                        int* __restrict__ piv, int* status) {
  __shared__ double rv[32];
  __shared__ int ri[32];
  __shared__ int s_piv;
  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5;
  for (int jj = 0; jj < jb; ++jj) {
    const int col = j0 + jj;
    double best = -1.0;
    int bi = col;
    bool nan = false;
    for (int i = col + tid; i < n; i += nt) {
      const double v = fabs(A[i + (int64_t)col * lda]);
      if (!(v == v)) nan = true;
      if (v > best) { best = v; bi = i; }
    }
    for (int o = 16; o > 0; o >>= 1) {
      const double ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
    }
    if (__any_sync(0xffffffffu, nan) && lane == 0) atomicOr(status, 2);
    if (lane == 0) { rv[warp] = best; ri[warp] = bi; }
    __syncthreads();
    if (tid == 0) {
      double b = rv[0];
      int p = ri[0];
      for (int w = 1; w < (nt >> 5); ++w)
        if (rv[w] > b || (rv[w] == b && ri[w] < p)) { b = rv[w]; p = ri[w]; }
      if (!(b > 0.0)) { atomicOr(status, 2); p = col; }
      s_piv = p;
      piv[col] = p;
    }
    __syncthreads();
    const int p = s_piv;
    if (p != col)
      for (int c = tid; c < jb; c += nt) {
        const double a = A[col + (int64_t)(j0 + c) * lda], b = A[p + (int64_t)(j0 + c) * lda];
        A[col + (int64_t)(j0 + c) * lda] = b;
        A[p + (int64_t)(j0 + c) * lda] = a;
      }
    __syncthreads();
    const double d = A[col + (int64_t)col * lda];
    const double rinv = d != 0.0 ? 1.0 / d : 0.0;
    for (int i = col + 1 + tid; i < n; i += nt) A[i + (int64_t)col * lda] *= rinv;
    __syncthreads();
    const int rows = n - col - 1, cols = jb - jj - 1;
    for (int64_t e = tid; e < (int64_t)rows * cols; e += nt) {
      const int i = col + 1 + (int)(e % rows), c = col + 1 + (int)(e / rows);
      A[i + (int64_t)c * lda] -= A[i + (int64_t)col * lda] * A[col + (int64_t)c * lda];
    }
    __syncthreads();
  }
}

// the panel's row interchanges applied to the columns outside it (one thread per column)
__global__ void lu_swap_kernel(int j0, int jb, int ncols, double* __restrict__ A, int64_t lda,
                               const int* __restrict__ piv) {
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < ncols; c += gridDim.x * blockDim.x) {
    if (c >= j0 && c < j0 + jb) continue;
    double* col = A + (int64_t)c * lda;
    for (int jj = 0; jj < jb; ++jj) {
      const int r1 = j0 + jj, r2 = piv[r1];
      if (r2 != r1) {
        const double a = col[r1];
        col[r1] = col[r2];
        col[r2] = a;
      }
    }
  }
}

// U12 = L11^-1 A12 for the columns right of the panel (unit lower triangular L11 in shared memory)
__global__ void __launch_bounds__(128) lu_trsm_kernel(int j0, int jb, int ncols, double* __restrict__ A, int64_t lda) {
  __shared__ double L[LU_NB][LU_NB + 1];
  for (int e = threadIdx.x; e < jb * jb; e += blockDim.x) L[e % jb][e / jb] = A[j0 + e % jb + (int64_t)(j0 + e / jb) * lda];
  __syncthreads();
  const int c = j0 + jb + blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ncols) return;
  double x[LU_NB];
  double* col = A + (int64_t)c * lda + j0;
#pragma unroll
  for (int i = 0; i < LU_NB; ++i) x[i] = i < jb ? col[i] : 0.0;
#pragma unroll
  for (int jj = 0; jj < LU_NB; ++jj)
#pragma unroll
    for (int ii = jj + 1; ii < LU_NB; ++ii)
      if (ii < jb) x[ii] = fma(-L[ii][jj], x[jj], x[ii]);
#pragma unroll
  for (int i = 0; i < LU_NB; ++i)
    if (i < jb) col[i] = x[i];
}

// back substitution U x = y for the nrhs columns right of the n x n factor (one CTA per right-hand side)
__global__ void __launch_bounds__(1024) lu_backsolve_kernel(int n, double* __restrict__ A, int64_t lda) {
  double* y = A + (int64_t)(n + blockIdx.x) * lda;
  __shared__ double xi;
  for (int i = n - 1; i >= 0; --i) {
    if (threadIdx.x == 0) {
      const double v = y[i] / A[i + (int64_t)i * lda];
      y[i] = v;
      xi = v;
    }
    __syncthreads();
    const double v = xi;
    for (int r = threadIdx.x; r < i; r += blockDim.x) y[r] = fma(-A[r + (int64_t)i * lda], v, y[r]);
    __syncthreads();
  }
}

__global__ void transpose_kernel(int64_t rows, int64_t cols, const double* __restrict__ in, int64_t ldi,
                                 double* __restrict__ out, int64_t ldo) {
  __shared__ double tile[32][33];
  const int64_t r0 = (int64_t)blockIdx.x * 32, c0 = (int64_t)blockIdx.y * 32;
  for (int q = threadIdx.y; q < 32; q += blockDim.y) {
    const int64_t r = r0 + threadIdx.x, c = c0 + q;
    tile[q][threadIdx.x] = (r < rows && c < cols) ? in[r + c * ldi] : 0.0;
  }
  __syncthreads();
  for (int q = threadIdx.y; q < 32; q += blockDim.y) {
    const int64_t c = c0 + threadIdx.x, r = r0 + q;
    if (r < rows && c < cols) out[c + r * ldo] = tile[threadIdx.x][q];
  }
}

// one compare-exchange stage of the bitonic network over np (power of two) pairs; (key, index) is a total order
__global__ void bitonic_stage_kernel(int64_t np, int64_t size, int64_t stride, double* __restrict__ key,
                                     int64_t* __restrict__ idx) {
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < np / 2; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = 2 * t - (t & (stride - 1)), j = i + stride;
    const double vi = key[i], vj = key[j];
    const int64_t gi = idx[i], gj = idx[j];
    const bool j_first = (vj < vi) || (vj == vi && gj < gi);
    const bool ascending = (i & size) == 0;
    if (j_first == ascending) {
      key[i] = vj; key[j] = vi;
      idx[i] = gj; idx[j] = gi;
    }
  }
}

__global__ void sort_init_kernel(int64_t n, int64_t np, const double* __restrict__ in, int negate,
                                 double* __restrict__ key, int64_t* __restrict__ idx, int* status) {
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < np; t += (int64_t)gridDim.x * blockDim.x) {
    double v = INFINITY;
    int64_t g = INT64_MAX;
    if (t < n) {
      v = negate ? -in[t] : in[t];
      g = t;
      if (!(v == v)) { atomicOr(status, 1); v = INFINITY; }
    }
    key[t] = v;
    idx[t] = g;
  }
}

}  // namespace

void lu_solve(cudaStream_t s, int n, int nrhs, double* Aaug, int64_t lda, int* piv, int* status) {
  const int ncols = n + nrhs;
  for (int j0 = 0; j0 < n; j0 += LU_NB) {
    const int jb = std::min(LU_NB, n - j0);
    lu_panel_kernel<<<1, 1024, 0, s>>>(n, j0, jb, Aaug, lda, piv, status);
    CK_LAUNCH();
    ++g_kernel_launches;
    lu_swap_kernel<<<std::max(1, std::min(64, (ncols + 127) / 128)), 128, 0, s>>>(j0, jb, ncols, Aaug, lda, piv);
    CK_LAUNCH();
    ++g_kernel_launches;
    const int right = ncols - j0 - jb;
    if (right > 0) {
      lu_trsm_kernel<<<(right + 127) / 128, 128, 0, s>>>(j0, jb, ncols, Aaug, lda);
      CK_LAUNCH();
      ++g_kernel_launches;
      const int below = n - j0 - jb;
      if (below > 0)
        gemm(s, false, below, right, jb, -1.0, Aaug + (j0 + jb) + (int64_t)j0 * lda, lda,
             Aaug + j0 + (int64_t)(j0 + jb) * lda, lda, 1.0, Aaug + (j0 + jb) + (int64_t)(j0 + jb) * lda, lda, nullptr,
             0);
    }
  }
  lu_backsolve_kernel<<<nrhs, 1024, 0, s>>>(n, Aaug, lda);
  CK_LAUNCH();
  ++g_kernel_launches;
}

void transpose(cudaStream_t s, int64_t rows, int64_t cols, const double* in, int64_t ldi, double* out, int64_t ldo) {
  if (rows <= 0 || cols <= 0) return;
  const dim3 grid((unsigned)ceil_div(rows, 32), (unsigned)ceil_div(cols, 32));
  transpose_kernel<<<grid, dim3(32, 8), 0, s>>>(rows, cols, in, ldi, out, ldo);
  CK_LAUNCH();
  ++g_kernel_launches;
}

int64_t sort_pairs_padded(int64_t n) {
  int64_t np = 2;
  while (np < n) np <<= 1;
  return np;
}

void sort_pairs(cudaStream_t s, int64_t n, const double* in, bool descending, double* key, int64_t* idx, int* status) {
  const int64_t np = sort_pairs_padded(n);
  const int blocks = (int)std::min<int64_t>(ceil_div(np, 256), 1184);
  sort_init_kernel<<<blocks, 256, 0, s>>>(n, np, in, descending ? 1 : 0, key, idx, status);
  CK_LAUNCH();
  ++g_kernel_launches;
  const int blocks2 = (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div(np / 2, 256), 1184));
  for (int64_t size = 2; size <= np; size <<= 1)
    for (int64_t stride = size >> 1; stride > 0; stride >>= 1) {
      bitonic_stage_kernel<<<blocks2, 256, 0, s>>>(np, size, stride, key, idx);
      CK_LAUNCH();
      ++g_kernel_launches;
    }
}

}  // namespace dav
