// n x k vector-block kernels, generators and layout helpers.  All HBM-bound streaming kernels:
// coalesced along the row index (column-major blocks), grids sized in multiples of the SM count.
#include <algorithm>
#include <map>
#include <mutex>
#include <utility>

#include "kernels.cuh"

namespace dav {

// ---- per-device facts / one-time kernel settings (common.cuh) ---------------------------------------------------
namespace {
constexpr int MAX_DEVICES = 64;
struct DevFacts { int max_smem = -1, sms = -1; };
std::mutex g_dev_mu;
DevFacts g_dev[MAX_DEVICES];
std::map<std::pair<const void*, int>, int> g_smem_set;  // (kernel, device) -> bytes already granted

const DevFacts& dev_facts() {
  int dev = 0;
  CK(cudaGetDevice(&dev));
  if (dev < 0 || dev >= MAX_DEVICES) DAV_THROW(DAV_ERR_CUDA, "device ordinal %d out of range", dev);
  std::lock_guard<std::mutex> lk(g_dev_mu);
  DevFacts& f = g_dev[dev];
  if (f.max_smem < 0) {
    CK(cudaDeviceGetAttribute(&f.max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    CK(cudaDeviceGetAttribute(&f.sms, cudaDevAttrMultiProcessorCount, dev));
    if (f.sms <= 0) f.sms = 148;
  }
  return f;
}
}  // namespace

int device_max_smem_optin() { return dev_facts().max_smem; }
int device_num_sms() { return dev_facts().sms; }

void ensure_dyn_smem_impl(const void* func, int bytes) {
  int dev = 0;
  CK(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lk(g_dev_mu);
  int& have = g_smem_set[std::make_pair(func, dev)];
  if (have >= bytes) return;
  CK(cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  have = bytes;
}

namespace {

// grid sizing: a few CTAs per SM of the B200 (148 SMs); a different SM count only changes the number of grid-stride
// iterations, never the result
constexpr int SMS = 148;

__global__ void fill_zero_kernel(double* p, size_t count) {
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < count; e += (size_t)gridDim.x * blockDim.x)
    p[e] = 0.0;
}

__global__ void copy_matrix_kernel(int64_t rows, int64_t cols, const double* __restrict__ src, int64_t lds,
                                   double* __restrict__ dst, int64_t ldd) {
  const int64_t total = rows * cols;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = e % rows, j = e / rows;
    dst[i + j * ldd] = src[i + j * lds];
  }
}

// generate_diagonal_dominant (array_utils.f90:86-113) for the local row block [row0, row0+nl)
__global__ void gen_diag_dominant_kernel(double* __restrict__ A, int64_t lda, int64_t nl, int64_t n, int64_t row0,
                                         double sparsity, int has_diag, double diag_val, uint64_t seed) {
  // grid.y walks columns, x walks rows: coalesced stores along the column
  for (int64_t j = blockIdx.y; j < n; j += gridDim.y) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nl; i += (int64_t)gridDim.x * blockDim.x) {
      const int64_t gi = row0 + i;
      double v;
      if (gi == j) v = has_diag ? diag_val : (double)(gi + 1);
      else {
        const uint64_t lo = (uint64_t)(gi < j ? gi : j), hi = (uint64_t)(gi < j ? j : gi);
        v = uniform01(seed, lo, hi) * sparsity;
      }
      A[i + j * lda] = v;
    }
  }
}

__global__ void extract_diag_kernel(const double* __restrict__ A, int64_t lda, int64_t nl, int64_t row0,
                                    double* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nl; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = A[i + (row0 + i) * lda];
}

// k smallest (value, global index) pairs in ascending lexicographic order: every CTA takes a
// chunk of 1024 candidates into shared memory, sorts it (bitonic network on the pairs) and writes its first k
// pairs.  Applied repeatedly (chunk winners become the next round's
// candidates) until one chunk is left; ties are broken by index, so the selection is the stable order.
constexpr int TOPK_CHUNK = 1024;
constexpr long long TOPK_PAD = 0x7fffffffffffffffLL;
__global__ void __launch_bounds__(TOPK_CHUNK) rank_select_kernel(const double* __restrict__ val,
                                                                 const int64_t* __restrict__ gidx, int64_t count,
                                                                 int64_t row0, int k, double* __restrict__ out_val,
                                                                 int64_t* __restrict__ out_idx, int final_round,
                                                                 int* status) {
  __shared__ double sv[TOPK_CHUNK];
  __shared__ long long sg[TOPK_CHUNK];
  const int t = threadIdx.x;
  const int64_t e = (int64_t)blockIdx.x * TOPK_CHUNK + t;
  double v = INFINITY;
  long long g = TOPK_PAD;
  if (e < count) {
    v = val[e];
    g = gidx ? (long long)gidx[e] : (long long)(row0 + e);
    if (g < 0) { g = TOPK_PAD; v = INFINITY; }
    if (!(v == v)) { atomicOr(status, 1); v = INFINITY; }
  }
  sv[t] = v;
  sg[t] = g;
  // r02: bitonic sort of the chunk's (value, index) pairs in shared memory -- 55 compare-exchange stages of 512 pairs
  // (~5 us) instead of every thread counting the keys that precede its own (1024 x 1024 comparisons, 55 us per
  // round, three rounds at n = 100,000).  (value, index) is a total order, so the result is the same stable order.
  for (int size = 2; size <= TOPK_CHUNK; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      __syncthreads();
      if (t < TOPK_CHUNK / 2) {
        const int i = 2 * t - (t & (stride - 1)), j = i + stride;
        const double vi = sv[i], vj = sv[j];
        const long long gi = sg[i], gj = sg[j];
        const bool j_first = (vj < vi) || (vj == vi && gj < gi);
        const bool ascending = (i & size) == 0;
        if (j_first == ascending) {
          sv[i] = vj; sv[j] = vi;
          sg[i] = gj; sg[j] = gi;
        }
      }
    }
  }
  __syncthreads();
  if (t < k) {  // (fewer than k candidates in this chunk: the padding sorts last and is written as empty slots)
    const long long gs = sg[t];
    out_val[(size_t)blockIdx.x * k + t] = gs == TOPK_PAD ? INFINITY : sv[t];
    out_idx[(size_t)blockIdx.x * k + t] = gs == TOPK_PAD ? -1 : (int64_t)gs;
  }
  (void)final_round;
}

// lower triangle <- upper triangle of the n x n matrix A (lda), 32 x 32 tiles through shared memory; block (ti, tj)
// with ti >= tj writes tile (rows of ti, columns of tj) from the transposed upper tile (rows of tj, columns of ti)
__global__ void __launch_bounds__(256) mirror_upper_kernel(double* __restrict__ A, int64_t lda, int64_t n,
                                                           int64_t tile_row0) {
  const int64_t ti = tile_row0 + blockIdx.y, tj = blockIdx.x;
  if (tj > ti) return;
  __shared__ double tile[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int q = ty; q < 32; q += 8) {  // upper tile: rows tj*32 + tx, columns ti*32 + q
    const int64_t r = tj * 32 + tx, c = ti * 32 + q;
    tile[q][tx] = (r < n && c < n) ? A[r + c * lda] : 0.0;
  }
  __syncthreads();
  for (int q = ty; q < 32; q += 8) {  // lower tile: rows ti*32 + tx, columns tj*32 + q  <-  upper (tj*32 + q, ti*32 + tx)
    const int64_t r = ti * 32 + tx, c = tj * 32 + q;
    if (r < n && c < n && r > c) A[r + c * lda] = tile[tx][q];
  }
}

// X (n x w) = columns c0 .. c0 + w of the identity
__global__ void unit_block_kernel(double* __restrict__ X, int64_t n, int64_t c0, int w) {
  const int64_t total = n * w;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t j = e / n, i = e - j * n;
    X[e] = (i == c0 + j) ? 1.0 : 0.0;
  }
}

// diag[i - row0] = Y(i - row0, i - c0) for the global rows i in [c0, c0 + w) that this rank owns
__global__ void take_diagonal_kernel(const double* __restrict__ Y, int64_t ldy, int64_t nl, int64_t row0, int64_t c0,
                                     int w, double* __restrict__ diag) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= w) return;
  const int64_t i = c0 + j - row0;
  if (i >= 0 && i < nl) diag[i] = Y[i + (int64_t)j * ldy];
}

__global__ void set_onehot_kernel(double* V, int64_t ldv, int64_t nl, int64_t row0, const int64_t* idx, int k) {
  for (int j = threadIdx.x; j < k; j += blockDim.x) {
    const int64_t r = idx[j] - row0;
    if (r >= 0 && r < nl) V[r + (int64_t)j * ldv] = 1.0;
  }
}

__global__ void gather_columns_kernel(const double* __restrict__ A, int64_t lda, int64_t nl,
                                      const int64_t* __restrict__ idx, int k, double* __restrict__ out, int64_t ldo) {
  for (int j = blockIdx.y; j < k; j += gridDim.y) {
    const double* src = A + idx[j] * lda;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nl; i += (int64_t)gridDim.x * blockDim.x)
      out[i + (int64_t)j * ldo] = src[i];
  }
}

constexpr int NCHUNK = 64;  // row chunks of the two-stage column reductions

// partial[j * NCHUNK + c] = sum over chunk c of X(i,j)^2
__global__ void col_norms2_stage1(int64_t nl, const double* __restrict__ X, int64_t ldx, double* __restrict__ partial) {
  __shared__ double red[32];
  const int j = blockIdx.y, c = blockIdx.x;
  const int64_t per = (nl + NCHUNK - 1) / NCHUNK;
  const int64_t beg = (int64_t)c * per, end = min(nl, beg + per);
  double s = 0.0;
  for (int64_t i = beg + threadIdx.x; i < end; i += blockDim.x) {
    const double v = X[i + (int64_t)j * ldx];
    s = fma(v, v, s);
  }
  s = warp_sum(s);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) red[wid] = s;
  __syncthreads();
  if (wid == 0) {
    double x = lane < (blockDim.x >> 5) ? red[lane] : 0.0;
    x = warp_sum(x);
    if (lane == 0) partial[(size_t)j * NCHUNK + c] = x;
  }
}

__global__ void col_reduce_stage2(int k, const double* __restrict__ partial, double* __restrict__ out) {
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < k; j += gridDim.x * blockDim.x) {
    double s = 0.0;
    for (int c = 0; c < NCHUNK; ++c) s += partial[(size_t)j * NCHUNK + c];
    out[j] = s;
  }
}

__global__ void scale_cols_rsqrt_kernel(int64_t nl, int k, double* __restrict__ X, int64_t ldx,
                                        const double* __restrict__ n2) {
  for (int j = blockIdx.y; j < k; j += gridDim.y) {
    const double d = n2[j];
    if (!(d > 0.0)) continue;
    const double f = 1.0 / sqrt(d);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nl; i += (int64_t)gridDim.x * blockDim.x)
      X[i + (int64_t)j * ldx] *= f;
  }
}

// fused residual + norm partial + DPR: grid (NCHUNK, k)
__global__ void residual_dpr_kernel(int64_t nl, double* __restrict__ R, int64_t ldr, double* __restrict__ C,
                                    int64_t ldc, const double* __restrict__ theta, const double* __restrict__ dA,
                                    const double* __restrict__ dB, int write_correction,
                                    double* __restrict__ partial) {
  __shared__ double red[32];
  const int j = blockIdx.y, c = blockIdx.x;
  const int64_t per = (nl + NCHUNK - 1) / NCHUNK;
  const int64_t beg = (int64_t)c * per, end = min(nl, beg + per);
  const double th = theta[j];
  double s = 0.0;
  for (int64_t i = beg + threadIdx.x; i < end; i += blockDim.x) {
    const double r = R[i + (int64_t)j * ldr] - th * C[i + (int64_t)j * ldc];
    R[i + (int64_t)j * ldr] = r;
    s = fma(r, r, s);
    if (write_correction) {
      const double den = th * (dB ? dB[i] : 1.0) - dA[i];  // davidson.f90:691,693 / :484, unguarded
      C[i + (int64_t)j * ldc] = r / den;
    }
  }
  s = warp_sum(s);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) red[wid] = s;
  __syncthreads();
  if (wid == 0) {
    double x = lane < (blockDim.x >> 5) ? red[lane] : 0.0;
    x = warp_sum(x);
    if (lane == 0) partial[(size_t)j * NCHUNK + c] = x;
  }
}

__global__ void fill_random_cols_kernel(double* X, int64_t ldx, int64_t nl, int64_t row0, const int* flags, int k,
                                        uint64_t salt) {
  for (int j = blockIdx.y; j < k; j += gridDim.y) {
    if (!flags[j]) continue;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nl; i += (int64_t)gridDim.x * blockDim.x)
      X[i + (int64_t)j * ldx] = 2.0 * uniform01(salt, (uint64_t)(row0 + i), (uint64_t)j) - 1.0;
  }
}

__global__ void fill_random_kernel(double* X, size_t count, uint64_t salt) {
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < count; e += (size_t)gridDim.x * blockDim.x)
    X[e] = 2.0 * uniform01(salt, e, 0x5bd1e995ULL) - 1.0;
}

__global__ void unstage_allgather_kernel(const double* __restrict__ stage, int world, int64_t chunk, int64_t n, int b,
                                         double* __restrict__ X, int64_t ldx) {
  for (int j = blockIdx.y; j < b; j += gridDim.y)
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
      const int64_t r = i / chunk, li = i - r * chunk;
      X[i + (int64_t)j * ldx] = stage[(size_t)r * chunk * b + (size_t)j * chunk + li];
    }
}

__global__ void stage_block_kernel(const double* __restrict__ X, int64_t ldx, int64_t nl, int64_t chunk, int b,
                                   double* __restrict__ stage) {
  for (int j = blockIdx.y; j < b; j += gridDim.y)
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < chunk; i += (int64_t)gridDim.x * blockDim.x)
      stage[(size_t)j * chunk + i] = (i < nl) ? X[i + (int64_t)j * ldx] : 0.0;
}

inline dim3 grid2(int64_t rows, int cols) {
  const unsigned gx = (unsigned)std::max<int64_t>(1, std::min<int64_t>(ceil_div(rows, 256), 4 * SMS));
  const unsigned gy = (unsigned)std::max(1, std::min(cols, 65535));
  return dim3(gx, gy);
}
inline int grid1(size_t count) {
  return (int)std::max<size_t>(1, std::min<size_t>((count + 255) / 256, (size_t)8 * SMS));
}

}  // namespace

#define LAUNCHED()  \
  do {              \
    CK_LAUNCH();    \
    ++g_kernel_launches; \
  } while (0)

void fill_zero(cudaStream_t s, double* p, size_t count) {
  if (!count) return;
  fill_zero_kernel<<<grid1(count), 256, 0, s>>>(p, count);
  LAUNCHED();
}

void copy_matrix(cudaStream_t s, int64_t rows, int64_t cols, const double* src, int64_t lds, double* dst,
                 int64_t ldd) {
  if (rows <= 0 || cols <= 0) return;
  copy_matrix_kernel<<<grid1((size_t)rows * cols), 256, 0, s>>>(rows, cols, src, lds, dst, ldd);
  LAUNCHED();
}

void gen_diag_dominant(cudaStream_t s, double* A, int64_t lda, int64_t nl, int64_t n, int64_t row0, double sparsity,
                       int has_diag, double diag_val, uint64_t seed) {
  if (nl <= 0) return;
  dim3 grid((unsigned)std::max<int64_t>(1, std::min<int64_t>(ceil_div(nl, 256), 64)),
            (unsigned)std::min<int64_t>(n, 8 * SMS));
  gen_diag_dominant_kernel<<<grid, 256, 0, s>>>(A, lda, nl, n, row0, sparsity, has_diag, diag_val, seed);
  LAUNCHED();
}

void extract_diag(cudaStream_t s, const double* A, int64_t lda, int64_t nl, int64_t row0, double* out) {
  if (nl <= 0) return;
  extract_diag_kernel<<<grid1((size_t)nl), 256, 0, s>>>(A, lda, nl, row0, out);
  LAUNCHED();
}

void topk_smallest(cudaStream_t s, const double* diag, const int64_t* gidx, int64_t count, int64_t row0, int k,
                   double* out_val, int64_t* out_idx, int* status, double* scratch_val, int64_t* scratch_idx) {
  if (k > TOPK_CHUNK) DAV_THROW(DAV_ERR_INVALID, "top-k selection supports k <= %d", TOPK_CHUNK);
  // rounds ping-pong between the two halves of the scratch arrays; the last round writes the outputs
  const size_t half = topk_scratch_entries(count, k) / 2;
  const double* cv = diag;
  const int64_t* ci = gidx;
  int64_t m = count, r0 = row0;
  int side = 0;
  for (;;) {
    const int64_t nch = std::max<int64_t>(1, ceil_div(m, TOPK_CHUNK));
    double* ov = (nch == 1) ? out_val : scratch_val + side * half;
    int64_t* oi = (nch == 1) ? out_idx : scratch_idx + side * half;
    rank_select_kernel<<<(unsigned)nch, TOPK_CHUNK, 0, s>>>(cv, ci, m, r0, k, ov, oi, nch == 1, status);
    LAUNCHED();
    if (nch == 1) break;
    cv = ov; ci = oi; m = nch * k; r0 = 0;
    side ^= 1;
  }
}

void mirror_upper_to_lower(cudaStream_t s, double* A, int64_t lda, int64_t n, int64_t row_begin, int64_t row_end) {
  // rows [row_begin, row_end) of the lower triangle (both multiples of 32, or row_end == n) from columns of the same
  // range of the upper triangle: lets the caller mirror a column panel as soon as it has arrived
  if (row_end < 0) row_end = n;
  if (n <= 1 || row_end <= row_begin) return;
  const unsigned t0 = (unsigned)(row_begin / 32), t1 = (unsigned)((row_end + 31) / 32);
  mirror_upper_kernel<<<dim3(t1, t1 - t0), 256, 0, s>>>(A, lda, n, (int64_t)t0);
  LAUNCHED();
}

void unit_block(cudaStream_t s, double* X, int64_t n, int64_t c0, int w) {
  if (n <= 0 || w <= 0) return;
  unit_block_kernel<<<grid1((size_t)n * w), 256, 0, s>>>(X, n, c0, w);
  LAUNCHED();
}

void take_diagonal(cudaStream_t s, const double* Y, int64_t ldy, int64_t nl, int64_t row0, int64_t c0, int w,
                   double* diag) {
  if (w <= 0) return;
  take_diagonal_kernel<<<(w + 63) / 64, 64, 0, s>>>(Y, ldy, nl, row0, c0, w, diag);
  LAUNCHED();
}

void set_onehot(cudaStream_t s, double* V, int64_t ldv, int64_t nl, int64_t row0, const int64_t* idx, int k) {
  set_onehot_kernel<<<1, 256, 0, s>>>(V, ldv, nl, row0, idx, k);
  LAUNCHED();
}

void gather_columns(cudaStream_t s, const double* A, int64_t lda, int64_t nl, const int64_t* idx, int k, double* out,
                    int64_t ldo) {
  if (nl <= 0 || k <= 0) return;
  gather_columns_kernel<<<grid2(nl, k), 256, 0, s>>>(A, lda, nl, idx, k, out, ldo);
  LAUNCHED();
}

void col_norms2(cudaStream_t s, int64_t nl, int k, const double* X, int64_t ldx, double* partial, double* out) {
  if (k <= 0) return;
  col_norms2_stage1<<<dim3(NCHUNK, k), 256, 0, s>>>(nl, X, ldx, partial);
  LAUNCHED();
  col_reduce_stage2<<<(k + 127) / 128, 128, 0, s>>>(k, partial, out);
  LAUNCHED();
}

void scale_cols_rsqrt(cudaStream_t s, int64_t nl, int k, double* X, int64_t ldx, const double* n2) {
  if (nl <= 0 || k <= 0) return;
  scale_cols_rsqrt_kernel<<<grid2(nl, k), 256, 0, s>>>(nl, k, X, ldx, n2);
  LAUNCHED();
}

void residual_dpr(cudaStream_t s, int64_t nl, int k, double* R, int64_t ldr, double* C, int64_t ldc,
                  const double* theta, const double* dA, const double* dB, bool write_correction, double* partial,
                  double* n2out) {
  if (k <= 0) return;
  residual_dpr_kernel<<<dim3(NCHUNK, k), 256, 0, s>>>(nl, R, ldr, C, ldc, theta, dA, dB, write_correction ? 1 : 0,
                                                      partial);
  LAUNCHED();
  col_reduce_stage2<<<(k + 127) / 128, 128, 0, s>>>(k, partial, n2out);
  LAUNCHED();
}

void fill_random_cols(cudaStream_t s, double* X, int64_t ldx, int64_t nl, int64_t row0, const int* flags, int k,
                      uint64_t salt) {
  if (nl <= 0 || k <= 0) return;
  fill_random_cols_kernel<<<grid2(nl, k), 256, 0, s>>>(X, ldx, nl, row0, flags, k, salt);
  LAUNCHED();
}

void fill_random(cudaStream_t s, double* X, size_t count, uint64_t salt) {
  if (!count) return;
  fill_random_kernel<<<grid1(count), 256, 0, s>>>(X, count, salt);
  LAUNCHED();
}

void unstage_allgather(cudaStream_t s, const double* stage, int world, int64_t chunk, int64_t n, int b, double* X,
                       int64_t ldx) {
  unstage_allgather_kernel<<<grid2(n, b), 256, 0, s>>>(stage, world, chunk, n, b, X, ldx);
  LAUNCHED();
}

void stage_block(cudaStream_t s, const double* X, int64_t ldx, int64_t nl, int64_t chunk, int b, double* stage) {
  stage_block_kernel<<<grid2(chunk, b), 256, 0, s>>>(X, ldx, nl, chunk, b, stage);
  LAUNCHED();
}

}  // namespace dav
