// FP64 GEMM for the tall-skinny shapes around the block matvec (SIMT kernel + tensor-pipe kernel):
//   NN  C(M x N) = alpha * A(M x K) * B(K x N) + beta * C      M = local rows (huge), K,N <= few hundred
//   TN  C(M x N) = alpha * A(K x M)^T * B(K x N) + beta * C    K = local rows (huge) -> split-K
// Replaces the reference's lapack_matmul / DGEMM call sites other than the A*V stream
// (davidson.f90:131,159,218,223,380-381,397,407-410,438; lapack_wrapper.f90:279-328).
// 64x64x16 tiles, 256 threads, 4x4 register microtile; split-K partials are summed in a fixed
// order by a second kernel, so results are bit-reproducible run to run.
// Default since r01 v6: the tensor-pipe variant of both shapes (gemm_dmma_kernel below): the same 64x64 output
// tiles and split-K decomposition, but each of 4 warps owns a 32x32 block of the tile and feeds DMMA.8x8x4
// (mma.sync.m8n8k4.f64) straight from global memory / L1 -- the operands of these products are tall-skinny blocks that
// are read exactly once, so there is nothing to stage in shared memory and no barrier in the loop.  Measured in the
// n = 100,000 solve: orthonormalisation 1.69 -> 1.26 ms, residuals 0.88 -> 0.58, projections 0.45 -> 0.29.
// DAV_GEMM_IMPL=0 selects the SIMT kernel (kept as the in-library reference of the parity test).
#include <algorithm>
#include <cstdlib>

#include "kernels.cuh"

namespace dav {

thread_local long long g_kernel_launches = 0;

namespace {

constexpr int BM = 64, BN = 64, BK = 16, LDS_ = 66, NT = 256;
constexpr int GEMM_SPLIT_TARGET_DEFAULT = 296;  // CTAs a split-K launch aims for (DAV_GEMM_SPLIT_TARGET overrides)
constexpr int GEMM_IMPL_DEFAULT = 1;  // 1: tensor-pipe kernel, 0: SIMT kernel (DAV_GEMM_IMPL overrides)

template <bool TA>
__global__ void __launch_bounds__(NT) gemm_kernel(int64_t M, int64_t N, int64_t K, int64_t Kchunk, double alpha,
                                                  const double* __restrict__ A, int64_t lda,
                                                  const double* __restrict__ B, int64_t ldb, double beta,
                                                  double* __restrict__ C, int64_t ldc, double* __restrict__ ws,
                                                  int splits, int to_ws) {
  __shared__ __align__(16) double As[BK][LDS_];
  __shared__ __align__(16) double Bs[BK][LDS_];
  const int tid = threadIdx.x;
  const int tm = tid % 16, tn = tid / 16;
  const int64_t m0 = (int64_t)blockIdx.x * BM, n0 = (int64_t)blockIdx.y * BN;
  const int z = blockIdx.z;
  const int64_t kbeg = (int64_t)z * Kchunk;
  const int64_t kend = min(K, kbeg + Kchunk);

  double acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;

  for (int64_t k0 = kbeg; k0 < kend; k0 += BK) {
    // ---- load A tile into As[k][m]
    if (!TA) {
      const int m = tid % 64, kq = tid / 64;
      const int64_t gm = m0 + m;
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int k = kq + 4 * r;
        const int64_t gk = k0 + k;
        As[k][m] = (gm < M && gk < kend) ? A[gm + gk * lda] : 0.0;
      }
    } else {
      const int k = tid % 16, mq = tid / 16;
      const int64_t gk = k0 + k;
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int m = mq + 16 * r;
        const int64_t gm = m0 + m;
        As[k][m] = (gm < M && gk < kend) ? A[gk + gm * lda] : 0.0;
      }
    }
    // ---- load B tile into Bs[k][j]
    {
      const int k = tid % 16, jq = tid / 16;
      const int64_t gk = k0 + k;
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int j = jq + 16 * r;
        const int64_t gj = n0 + j;
        Bs[k][j] = (gj < N && gk < kend) ? B[gk + gj * ldb] : 0.0;
      }
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const double2 a01 = *reinterpret_cast<const double2*>(&As[kk][tm * 4]);
      const double2 a23 = *reinterpret_cast<const double2*>(&As[kk][tm * 4 + 2]);
      const double2 b01 = *reinterpret_cast<const double2*>(&Bs[kk][tn * 4]);
      const double2 b23 = *reinterpret_cast<const double2*>(&Bs[kk][tn * 4 + 2]);
      const double a[4] = {a01.x, a01.y, a23.x, a23.y};
      const double b[4] = {b01.x, b01.y, b23.x, b23.y};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }

  if (splits == 1 && !to_ws) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t gj = n0 + tn * 4 + j;
      if (gj >= N) continue;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int64_t gm = m0 + tm * 4 + i;
        if (gm >= M) continue;
        double* c = C + gm + gj * ldc;
        *c = (beta == 0.0) ? alpha * acc[i][j] : alpha * acc[i][j] + beta * (*c);
      }
    }
  } else {
    double* w = ws + (size_t)z * (size_t)M * (size_t)N;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t gj = n0 + tn * 4 + j;
      if (gj >= N) continue;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int64_t gm = m0 + tm * 4 + i;
        if (gm >= M) continue;
        w[gm + gj * M] = acc[i][j];
      }
    }
  }
}

// Tensor-pipe variant.  Fragment layout of mma.sync.m8n8k4.f64 (g = lane >> 2, t = lane & 3): a = op(A)[m0+g][k+t],
// b = B[k+t][n0+g], d0/d1 = C[m0+g][n0+2t], C[m0+g][n0+2t+1].  TA: A is stored K x M (the projections, K = local rows).
template <bool TA>
__global__ void __launch_bounds__(128) gemm_dmma_kernel(int64_t M, int64_t N, int64_t K, int64_t Kchunk, double alpha,
                                                        const double* __restrict__ A, int64_t lda,
                                                        const double* __restrict__ B, int64_t ldb, double beta,
                                                        double* __restrict__ C, int64_t ldc, double* __restrict__ ws,
                                                        int splits, int to_ws) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int64_t m0 = (int64_t)blockIdx.x * BM + (warp & 1) * 32;
  const int64_t n0 = (int64_t)blockIdx.y * BN + (warp >> 1) * 32;
  if (m0 >= M || n0 >= N) return;  // warp-uniform: the whole warp leaves together
  const int z = blockIdx.z;
  const int64_t kbeg = (int64_t)z * Kchunk;
  const int64_t kend = min(K, kbeg + Kchunk);

  double acc[4][4][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  // per-lane base pointers and bounds of the four m-tiles / n-tiles
  const double* ap[4];
  const double* bp[4];
  bool am[4], bn[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t m = m0 + 8 * i + g;
    am[i] = m < M;
    ap[i] = TA ? A + (am[i] ? m : 0) * lda + t : A + (am[i] ? m : 0) + (int64_t)t * lda;
    const int64_t n = n0 + 8 * i + g;
    bn[i] = n < N;
    bp[i] = B + (bn[i] ? n : 0) * ldb + t;
  }
  const int64_t astep = TA ? 1 : lda;  // distance between consecutive k in op(A)

#pragma unroll 2
  for (int64_t kk = kbeg; kk < kend; kk += 4) {
    const bool kin = kk + t < kend;
    double a[4], b[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      a[i] = (kin && am[i]) ? ap[i][kk * astep] : 0.0;
      b[i] = (kin && bn[i]) ? bp[i][kk] : 0.0;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j)
        asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
            : "+d"(acc[i][j][0]), "+d"(acc[i][j][1])
            : "d"(a[i]), "d"(b[j]));
  }

#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t m = m0 + 8 * i + g;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const int64_t n = n0 + 8 * j + 2 * t + c;
        if (n >= N) continue;
        if (splits == 1 && !to_ws) {
          double* q = C + m + n * ldc;
          *q = (beta == 0.0) ? alpha * acc[i][j][c] : alpha * acc[i][j][c] + beta * (*q);
        } else {
          ws[(size_t)z * (size_t)M * (size_t)N + (size_t)m + (size_t)n * (size_t)M] = acc[i][j][c];
        }
      }
  }
}

// Sums the split-K partials in a fixed order (bit-reproducible).  A CTA owns 64 consecutive output elements; its 4
// thread groups each add every 4th partial (coalesced 512-byte rows of the workspace) and the 4 group sums are
// combined in shared memory in group order.
__global__ void __launch_bounds__(256) splitk_reduce_kernel(int64_t M, int64_t N, int splits, double alpha,
                                                            const double* __restrict__ ws, double beta,
                                                            double* __restrict__ C, int64_t ldc) {
  __shared__ double part[4][64];
  const int64_t total = M * N;
  const int el = threadIdx.x & 63, g = threadIdx.x >> 6;
  const int64_t e = (int64_t)blockIdx.x * 64 + el;
  double s = 0.0;
  if (e < total)
    for (int z = g; z < splits; z += 4) s += ws[(size_t)z * total + e];
  part[g][el] = s;
  __syncthreads();
  if (g == 0 && e < total) {
    s = ((part[0][el] + part[1][el]) + part[2][el]) + part[3][el];
    const int64_t m = e % M, j = e / M;
    double* c = C + m + j * ldc;
    *c = (beta == 0.0) ? alpha * s : alpha * s + beta * (*c);
  }
}

}  // namespace

void gemm(cudaStream_t s, bool transA, int64_t M, int64_t N, int64_t K, double alpha, const double* A, int64_t lda,
          const double* B, int64_t ldb, double beta, double* C, int64_t ldc, double* ws, size_t ws_doubles,
          int* partials_out) {
  if (M <= 0 || N <= 0) {
    if (partials_out) *partials_out = 0;
    return;
  }
  if (partials_out && (ws == nullptr || ws_doubles < (size_t)M * (size_t)N))
    DAV_THROW(DAV_ERR_STATE, "gemm: workspace too small for the partials of a %lld x %lld product", (long long)M,
              (long long)N);
  const int to_ws = partials_out ? 1 : 0;
  const int64_t gx = ceil_div(M, BM), gy = ceil_div(N, BN);
  int splits = 1;
  // read per call so one process can compare the two implementations (tests, bench A/B)
  const char* impl_env = std::getenv("DAV_GEMM_IMPL");
  const int impl = impl_env ? std::atoi(impl_env) : GEMM_IMPL_DEFAULT;
  if (K > 1024 && gx * gy < 592) {  // long reduction, few output tiles: split K over the grid
    // about two CTAs per SM in total: more partials only lengthen the reduction (each is M x N doubles of traffic)
    const char* tgt_env = std::getenv("DAV_GEMM_SPLIT_TARGET");
    const int64_t target = tgt_env ? std::max(1, std::atoi(tgt_env)) : GEMM_SPLIT_TARGET_DEFAULT;
    splits = (int)std::min<int64_t>(ceil_div(target, gx * gy), ceil_div(K, 256));
    const size_t need = (size_t)M * (size_t)N;
    if (ws == nullptr || need == 0) splits = 1;
    else splits = (int)std::min<size_t>((size_t)splits, ws_doubles / need);
    if (splits < 1) splits = 1;
  }
  const int64_t Kchunk = round_up(ceil_div(std::max<int64_t>(K, 1), splits), BK);
  splits = (int)ceil_div(std::max<int64_t>(K, 1), Kchunk);
  if (gy > 65535 || splits > 65535) DAV_THROW(DAV_ERR_INVALID, "gemm grid too large");
  dim3 grid((unsigned)gx, (unsigned)gy, (unsigned)splits);
  if (impl == 1 && transA)
    gemm_dmma_kernel<true><<<grid, 128, 0, s>>>(M, N, K, Kchunk, alpha, A, lda, B, ldb, beta, C, ldc, ws, splits, to_ws);
  else if (impl == 1)
    gemm_dmma_kernel<false><<<grid, 128, 0, s>>>(M, N, K, Kchunk, alpha, A, lda, B, ldb, beta, C, ldc, ws, splits, to_ws);
  else if (transA)
    gemm_kernel<true><<<grid, NT, 0, s>>>(M, N, K, Kchunk, alpha, A, lda, B, ldb, beta, C, ldc, ws, splits, to_ws);
  else
    gemm_kernel<false><<<grid, NT, 0, s>>>(M, N, K, Kchunk, alpha, A, lda, B, ldb, beta, C, ldc, ws, splits, to_ws);
  CK_LAUNCH();
  ++g_kernel_launches;
  if (partials_out) {  // the caller reduces the partials itself (Comm::reduce_sum: split-K + ranks + layout)
    *partials_out = splits;
    return;
  }
  if (splits > 1) {
    const int64_t total = M * N;
    const int blocks = (int)ceil_div(total, 64);
    splitk_reduce_kernel<<<blocks, 256, 0, s>>>(M, N, splits, alpha, ws, beta, C, ldc);
    CK_LAUNCH();
    ++g_kernel_launches;
  }
}

}  // namespace dav
