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
//
// r02: these products are latency-bound, not flop-bound, whenever a rank holds few rows (8 GPUs: 12,500): the r01
// kernel walked K with at most two k-steps of loads in flight per warp and 4 warps per CTA -- 40 us for a 50 MFLOP
// projection.  Now (i) the k-loop is processed in chunks of UN = 4 k-steps whose 32 loads per lane are all issued
// before the 64 DMMAs that consume them, (ii) KG warp groups per CTA take interleaved quarters of the CTA's K range
// and are summed through shared memory in a fixed tree order (bit-reproducible) -- the dependent chain per warp is
// KG times shorter and the number of split-K partials does not grow, (iii) blocks of <= 32 columns use a 128 x 32
// CTA tile (WNS = 1) so that no warp idles on columns that do not exist.  Out-of-range rows / columns of a sub-tile
// read clamped (valid) addresses and accumulate into registers that are never stored.
constexpr int UN_MAX = 4;

template <bool TA, int KG, int WNS, bool VEC>
__global__ void __launch_bounds__(128 * KG) gemm_dmma_kernel(int64_t M, int64_t N, int64_t K, int64_t Kchunk,
                                                             double alpha, const double* __restrict__ A, int64_t lda,
                                                             const double* __restrict__ B, int64_t ldb, double beta,
                                                             double* __restrict__ C, int64_t ldc,
                                                             double* __restrict__ ws, int splits, int to_ws,
                                                             const double* __restrict__ A2, int64_t lda2,
                                                             int64_t asplit) {
  extern __shared__ __align__(16) double red_sm[];
  constexpr int WMS = 4 / WNS;
  constexpr int UN = KG == 4 ? 2 : UN_MAX;  // 512 threads leave 128 registers per thread: two k-steps in flight
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int sub = warp & 3, kg = warp >> 2;
  const int64_t m0 = (int64_t)blockIdx.x * (32 * WMS) + (sub % WMS) * 32;
  const int64_t n0 = (int64_t)blockIdx.y * (32 * WNS) + (sub / WMS) * 32;
  const bool active = m0 < M && n0 < N;  // warp-uniform
  const int z = blockIdx.z;
  const int64_t kbeg = (int64_t)z * Kchunk;
  const int64_t kend = min(K, kbeg + Kchunk);
  // this warp group's part of [kbeg, kend): contiguous, a multiple of 4 rows
  const int64_t per = ((kend - kbeg + KG - 1) / KG + 3) & ~(int64_t)3;
  const int64_t gb = kbeg + (int64_t)kg * per;
  const int64_t ge = min(kend, gb + per);

  double acc[4][4][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  if (active && gb < ge) {
    // per-lane base pointers (clamped to a valid row / column) of the four m-tiles / n-tiles
    // op(A) may be given as TWO blocks (A2 != nullptr): TN -- columns [0, asplit) of A^T's row index m come from A,
    // [asplit, M) from A2 (the Gram / projection of [V | T] in one product); NN -- columns [0, asplit) of A from A,
    // [asplit, K) from A2 (the update [V | T] * Z in one product; asplit is a multiple of 4 there).
    const double* ap[4];
    const double* bp[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int64_t m = min(m0 + 8 * i + g, M - 1);
      if (TA) ap[i] = (A2 != nullptr && m >= asplit) ? A2 + (m - asplit) * lda2 + t : A + m * lda + t;
      else ap[i] = A + m + (int64_t)t * lda;
      const int64_t n = min(n0 + 8 * i + g, N - 1);
      bp[i] = B + n * ldb + t;
    }
    int64_t astep = TA ? 1 : lda;  // distance between consecutive k in op(A)
    const int nseg = (!TA && A2 != nullptr) ? 2 : 1;
    for (int seg = 0; seg < nseg; ++seg) {
    int64_t sb = gb, se = ge;  // this segment of the warp group's K range
    if (nseg == 2) {
      if (seg == 0) se = min(ge, asplit);
      else {
        sb = max(gb, asplit);
        astep = lda2;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int64_t m = min(m0 + 8 * i + g, M - 1);
          ap[i] = A2 + m + ((int64_t)t - asplit) * lda2;  // indexed with the global k below
        }
      }
      if (sb >= se) continue;
    }
    // which 8-wide tiles of this warp exist at all (warp-uniform): skipped MMAs for N = 16, edge tiles
    bool mi[4], nj[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      mi[i] = m0 + 8 * i < M;
      nj[i] = n0 + 8 * i < N;
    }
    int64_t kk = sb;
    if (VEC) {
      // TN, 16-byte aligned operands: the order of the k index inside an MMA is free as long as A and B agree, so
      // lane t takes 2 VH CONSECUTIVE rows of its column with VH 16-byte loads and feeds one of them to each of the
      // chunk's 2 VH MMAs -- half as many load instructions (and L1 wavefronts) per flop
      constexpr int VH = KG == 4 ? 1 : 2;  // 16-byte loads per tile and chunk (512 threads: 128 registers each)
      for (; kk + 8 * VH <= se; kk += 8 * VH) {
        double2 a[4][VH], b[4][VH];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int h = 0; h < VH; ++h) {
            // rows kk + 2 VH t + 2h, +1 (ap / bp carry + t already)
            a[i][h] = *reinterpret_cast<const double2*>(ap[i] + kk + (2 * VH - 1) * t + 2 * h);
            b[i][h] = *reinterpret_cast<const double2*>(bp[i] + kk + (2 * VH - 1) * t + 2 * h);
          }
#pragma unroll
        for (int u = 0; u < 2 * VH; ++u)
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            if (!mi[i]) continue;
            const double av = (u & 1) ? a[i][u >> 1].y : a[i][u >> 1].x;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              if (!nj[j]) continue;
              const double bv = (u & 1) ? b[j][u >> 1].y : b[j][u >> 1].x;
              asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                  : "+d"(acc[i][j][0]), "+d"(acc[i][j][1])
                  : "d"(av), "d"(bv));
            }
          }
      }
    }
    for (; kk + 4 * UN <= se; kk += 4 * UN) {
      double a[UN][4], b[UN][4];
#pragma unroll
      for (int u = 0; u < UN; ++u)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          a[u][i] = ap[i][(kk + 4 * u) * astep];
          b[u][i] = bp[i][kk + 4 * u];
        }
#pragma unroll
      for (int u = 0; u < UN; ++u)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (!mi[i]) continue;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (!nj[j]) continue;
            asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                : "+d"(acc[i][j][0]), "+d"(acc[i][j][1])
                : "d"(a[u][i]), "d"(b[u][j]));
          }
        }
    }
    for (; kk < se; kk += 4) {  // tail: fewer than 4 UN rows left, the last k-step may be partial
      const bool kin = kk + t < se;
      double a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        a[i] = kin ? ap[i][kk * astep] : 0.0;
        b[i] = kin ? bp[i][kk] : 0.0;
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (!mi[i]) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (!nj[j]) continue;
          asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
              : "+d"(acc[i][j][0]), "+d"(acc[i][j][1])
              : "d"(a[i]), "d"(b[j]));
        }
      }
    }
    }  // segments
  }

  // sum of the KG warp groups, fixed tree order: (g0 + g2) + (g1 + g3) for KG = 4, g0 + g1 for KG = 2
  if (KG > 1) {
#pragma unroll
    for (int half = KG / 2; half >= 1; half >>= 1) {
      if (kg >= half && kg < 2 * half) {
        double* slot = red_sm + ((size_t)((kg - half) * 4 + sub) * 32) * 32 + lane;
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            slot[((i * 4 + j) * 2) * 32] = acc[i][j][0];
            slot[((i * 4 + j) * 2 + 1) * 32] = acc[i][j][1];
          }
      }
      __syncthreads();
      if (kg < half) {
        const double* slot = red_sm + ((size_t)(kg * 4 + sub) * 32) * 32 + lane;
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            acc[i][j][0] += slot[((i * 4 + j) * 2) * 32];
            acc[i][j][1] += slot[((i * 4 + j) * 2 + 1) * 32];
          }
      }
      if (half > 1) __syncthreads();
    }
  }
  if (kg != 0 || !active) return;

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

// ---- r02: residuals + DPR corrections in ONE pass over the stored products (davidson.f90:163-170 via AV, (BV|V);
// :688-696; free :401-410, :484).  Was: GEMM AV*Y -> R, GEMM (BV|V)*Y -> C, residual_dpr_kernel (reads both, writes
// both) + stage 2 of the norms.  Here both products share the Y fragments of one k-loop (two accumulator sets per
// warp), and the epilogue forms r = AV y - theta (BV|V) y, the correction r / (theta dB_i - dA_i) (or, for GJD, the
// B-product (BV|V) y itself) and the column sums of r^2 of the warp's 32 rows (three shuffle stages over the 8 row
// lanes), written per (column, CTA row block) and summed in a fixed order by norm_partials_kernel.
template <int WNS>
__global__ void __launch_bounds__(128) resid_dmma_kernel(int64_t M, int64_t N, int64_t K,
                                                        const double* __restrict__ A1, const double* __restrict__ A2,
                                                        int64_t lda, const double* __restrict__ Y, int64_t ldy,
                                                        const double* __restrict__ theta,
                                                        const double* __restrict__ dA, const double* __restrict__ dB,
                                                        int write_correction, double* __restrict__ R, int64_t ldr,
                                                        double* __restrict__ C, int64_t ldc,
                                                        double* __restrict__ partial, int P) {
  constexpr int WMS = 4 / WNS;
  constexpr int UNR = 2;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int wm = warp % WMS, wn = warp / WMS;
  const int64_t m0 = (int64_t)blockIdx.x * (32 * WMS) + wm * 32;
  const int64_t n0 = (int64_t)blockIdx.y * (32 * WNS) + wn * 32;
  if (n0 >= N) return;  // warp-uniform (no barrier in this kernel)
  if (m0 >= M) {        // a warp below the last row still owns a slot of the norm partials
    if (g == 0) {
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int64_t n = n0 + 8 * (q >> 1) + 2 * t + (q & 1);
        if (n < N) partial[(size_t)n * P + (blockIdx.x * WMS + wm)] = 0.0;
      }
    }
    return;
  }
  double accA[4][4][2], accB[4][4][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) accA[i][j][0] = accA[i][j][1] = accB[i][j][0] = accB[i][j][1] = 0.0;
  int64_t aoff[4];
  const double* bp[4];
  bool mi[4], nj[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    aoff[i] = min(m0 + 8 * i + g, M - 1) + (int64_t)t * lda;
    bp[i] = Y + min(n0 + 8 * i + g, N - 1) * ldy + t;
    mi[i] = m0 + 8 * i < M;
    nj[i] = n0 + 8 * i < N;
  }
  int64_t kk = 0;
  for (; kk + 4 * UNR <= K; kk += 4 * UNR) {
    double a1[UNR][4], a2[UNR][4], b[UNR][4];
#pragma unroll
    for (int u = 0; u < UNR; ++u)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        a1[u][i] = A1[aoff[i] + (kk + 4 * u) * lda];
        a2[u][i] = A2[aoff[i] + (kk + 4 * u) * lda];
        b[u][i] = bp[i][kk + 4 * u];
      }
#pragma unroll
    for (int u = 0; u < UNR; ++u)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (!mi[i]) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (!nj[j]) continue;
          asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
              : "+d"(accA[i][j][0]), "+d"(accA[i][j][1])
              : "d"(a1[u][i]), "d"(b[u][j]));
          asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
              : "+d"(accB[i][j][0]), "+d"(accB[i][j][1])
              : "d"(a2[u][i]), "d"(b[u][j]));
        }
      }
  }
  for (; kk < K; kk += 4) {
    const bool kin = kk + t < K;
    double a1[4], a2[4], b[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      a1[i] = kin ? A1[aoff[i] + kk * lda] : 0.0;
      a2[i] = kin ? A2[aoff[i] + kk * lda] : 0.0;
      b[i] = kin ? bp[i][kk] : 0.0;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (!mi[i]) continue;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (!nj[j]) continue;
        asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
            : "+d"(accA[i][j][0]), "+d"(accA[i][j][1])
            : "d"(a1[i]), "d"(b[j]));
        asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
            : "+d"(accB[i][j][0]), "+d"(accB[i][j][1])
            : "d"(a2[i]), "d"(b[j]));
      }
    }
  }
  // ---- epilogue
  double dAi[4], dBi[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t m = min(m0 + 8 * i + g, M - 1);
    dAi[i] = dA[m];
    dBi[i] = dB ? dB[m] : 1.0;
  }
  const int pid = blockIdx.x * WMS + wm;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    if (!nj[j]) continue;  // warp-uniform
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const int64_t n = n0 + 8 * j + 2 * t + c;
      const bool nin = n < N;
      const double th = theta[nin ? n : N - 1];
      double ssq = 0.0;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int64_t m = m0 + 8 * i + g;
        if (m < M && nin) {
          const double r = accA[i][j][c] - th * accB[i][j][c];
          R[m + n * ldr] = r;
          ssq = fma(r, r, ssq);
          // davidson.f90:691,693 / :484, unguarded
          C[m + n * ldc] = write_correction ? r / (th * dBi[i] - dAi[i]) : accB[i][j][c];
        }
      }
      // sum over the 8 row lanes (same t)
      ssq += __shfl_xor_sync(0xffffffffu, ssq, 4);
      ssq += __shfl_xor_sync(0xffffffffu, ssq, 8);
      ssq += __shfl_xor_sync(0xffffffffu, ssq, 16);
      if (g == 0 && nin) partial[(size_t)n * P + pid] = ssq;
    }
  }
}

// out[j] = sum of partial[j * P + 0 .. P) in a fixed order: one warp per column, lanes strided, butterfly at the end
__global__ void __launch_bounds__(256) norm_partials_kernel(int N, int P, const double* __restrict__ partial,
                                                           double* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int j = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (j >= N) return;
  double s0 = 0.0, s1 = 0.0;
  int p = lane;
  for (; p + 32 < P; p += 64) {
    s0 += partial[(size_t)j * P + p];
    s1 += partial[(size_t)j * P + p + 32];
  }
  if (p < P) s0 += partial[(size_t)j * P + p];
  double s = s0 + s1;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) out[j] = s;
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

namespace {
template <bool TA, int KG, int WNS, bool VEC>
void launch_dmma(cudaStream_t s, dim3 grid, int64_t M, int64_t N, int64_t K, int64_t Kchunk, double alpha,
                 const double* A, int64_t lda, const double* B, int64_t ldb, double beta, double* C, int64_t ldc,
                 double* ws, int splits, int to_ws, const double* A2, int64_t lda2, int64_t asplit) {
  const size_t sm = KG > 1 ? (size_t)(KG / 2) * 4 * 32 * 32 * sizeof(double) : 0;
  if (sm > 48 * 1024) ensure_dyn_smem(gemm_dmma_kernel<TA, KG, WNS, VEC>, (int)sm);
  gemm_dmma_kernel<TA, KG, WNS, VEC><<<grid, 128 * KG, sm, s>>>(M, N, K, Kchunk, alpha, A, lda, B, ldb, beta, C, ldc,
                                                               ws, splits, to_ws, A2, lda2, asplit);
}
int env_or(const char* name, int dflt) {
  const char* e = std::getenv(name);
  return e ? std::atoi(e) : dflt;
}
}  // namespace

size_t residual_fused_partials(int64_t nl, int nc) {
  return (size_t)std::max(nc, 1) * (size_t)(ceil_div(std::max<int64_t>(nl, 1), 32) + 4);
}

void residual_fused(cudaStream_t s, int64_t nl, int nc, int k, const double* AV, const double* BV, int64_t ldv,
                    const double* Y, int64_t ldy, const double* theta, const double* dA, const double* dB,
                    bool write_correction, double* R, int64_t ldr, double* C, int64_t ldc, double* partial,
                    double* n2out) {
  if (nc <= 0) return;
  if (nl <= 0) {  // a rank without rows contributes zero norms
    CK(cudaMemsetAsync(n2out, 0, (size_t)nc * sizeof(double), s));
    return;
  }
  const int wns = nc <= 32 ? 1 : 2;
  const int bm = 32 * (4 / wns), bn = 32 * wns;
  const dim3 grid((unsigned)ceil_div(nl, bm), (unsigned)ceil_div(nc, bn));
  const int P = (int)grid.x * (4 / wns);
  if (wns == 1)
    resid_dmma_kernel<1><<<grid, 128, 0, s>>>(nl, nc, k, AV, BV, ldv, Y, ldy, theta, dA, dB, write_correction ? 1 : 0, R,
                                              ldr, C, ldc, partial, P);
  else
    resid_dmma_kernel<2><<<grid, 128, 0, s>>>(nl, nc, k, AV, BV, ldv, Y, ldy, theta, dA, dB, write_correction ? 1 : 0, R,
                                              ldr, C, ldc, partial, P);
  CK_LAUNCH();
  ++g_kernel_launches;
  norm_partials_kernel<<<(nc + 7) / 8, 256, 0, s>>>(nc, P, partial, n2out);
  CK_LAUNCH();
  ++g_kernel_launches;
}

void gemm(cudaStream_t s, bool transA, int64_t M, int64_t N, int64_t K, double alpha, const double* A, int64_t lda,
          const double* B, int64_t ldb, double beta, double* C, int64_t ldc, double* ws, size_t ws_doubles,
          int* partials_out, const GemmSplit* split) {
  if (M <= 0 || N <= 0) {
    if (partials_out) *partials_out = 0;
    return;
  }
  if (partials_out && (ws == nullptr || ws_doubles < (size_t)M * (size_t)N))
    DAV_THROW(DAV_ERR_STATE, "gemm: workspace too small for the partials of a %lld x %lld product", (long long)M,
              (long long)N);
  const int to_ws = partials_out ? 1 : 0;
  // read per call so one process can compare the implementations (tests, bench A/B)
  int impl = env_or("DAV_GEMM_IMPL", GEMM_IMPL_DEFAULT);
  const double* A2 = split ? split->A2 : nullptr;
  const int64_t lda2 = split ? split->lda2 : 0, asplit = split ? split->at : 0;
  if (A2) {
    impl = 1;  // the two-block operand exists on the tensor-pipe kernel only
    if (asplit <= 0 || asplit >= (transA ? M : K) || (!transA && asplit % 4 != 0))
      DAV_THROW(DAV_ERR_INVALID, "gemm: bad operand split %lld", (long long)asplit);
  }
  // tensor-pipe kernel: 128 x 32 CTA tiles for blocks of <= 32 columns, 64 x 64 otherwise
  const int wns = (impl == 1 && N <= 32) ? 1 : 2;
  const int bm = impl == 1 ? 32 * (4 / wns) : BM, bn = impl == 1 ? 32 * wns : BN;
  const int64_t gx = ceil_div(M, bm), gy = ceil_div(N, bn);
  // warp groups per CTA along K (tensor-pipe kernel): 4 whenever every group still gets a full chunk of UN k-steps
  int kg = env_or("DAV_GEMM_KG", 0);
  if (kg != 1 && kg != 2 && kg != 4) kg = 0;
  int splits = 1;
  if (K > 1024 && gx * gy < 592) {  // long reduction, few output tiles: split K over the grid
    // about one CTA (8-16 warps) per SM in total: more partials only lengthen the reduction (each is M x N doubles)
    const int64_t target = std::max(1, env_or("DAV_GEMM_SPLIT_TARGET", GEMM_SPLIT_TARGET_DEFAULT));
    // rows of K a CTA should at least own (measured, r02_gemm_bench: 12,500 rows -> ~130 CTAs of 96 rows beat both
    // 196 x 64 and 74 x 176; 100,000 rows -> 296 CTAs of ~340 rows beat 148 x 680 by up to 2x on wide outputs)
    const int64_t min_rows = impl == 1 ? 96 : 256;
    splits = (int)std::min<int64_t>(ceil_div(target, gx * gy), ceil_div(K, min_rows));
    (void)0;
    const size_t need = (size_t)M * (size_t)N;
    if (ws == nullptr || need == 0) splits = 1;
    else splits = (int)std::min<size_t>((size_t)splits, ws_doubles / need);
    if (splits < 1) splits = 1;
  }
  const int64_t Kchunk = round_up(ceil_div(std::max<int64_t>(K, 1), splits), BK);
  splits = (int)ceil_div(std::max<int64_t>(K, 1), Kchunk);
  if (gy > 65535 || splits > 65535) DAV_THROW(DAV_ERR_INVALID, "gemm grid too large");
  dim3 grid((unsigned)gx, (unsigned)gy, (unsigned)splits);
  if (impl == 1) {
    // warp groups along K: only the long split-K chunks (>= 256 rows per CTA: one rank holding ~100,000 rows) gain
    // from the shorter dependent chain (-8 % on wide outputs); short chunks and the NN shapes (K <= a few hundred,
    // thousands of CTAs) are faster with 4-warp CTAs and no shared-memory sum (measured, profiles/r02_gemm_bench*)
    if (kg == 0) kg = (transA && Kchunk >= 256) ? 4 : 1;
    const bool vec = transA && env_or("DAV_GEMM_VEC", 1) != 0 && lda % 2 == 0 && ldb % 2 == 0 &&
                     (reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(B) & 15) == 0 &&
                     (!A2 || (lda2 % 2 == 0 && (reinterpret_cast<uintptr_t>(A2) & 15) == 0));
#define DAV_GEMM_CASE(TA_, KG_, WNS_, VEC_)                                                                         \
  launch_dmma<TA_, KG_, WNS_, VEC_>(s, grid, M, N, K, Kchunk, alpha, A, lda, B, ldb, beta, C, ldc, ws, splits, to_ws, \
                                    A2, lda2, asplit)
#define DAV_GEMM_KG(TA_, WNS_, VEC_)                                   \
  do {                                                                 \
    if (kg == 4) DAV_GEMM_CASE(TA_, 4, WNS_, VEC_);                    \
    else if (kg == 2) DAV_GEMM_CASE(TA_, 2, WNS_, VEC_);               \
    else DAV_GEMM_CASE(TA_, 1, WNS_, VEC_);                            \
  } while (0)
    if (transA && vec) {
      if (wns == 1) DAV_GEMM_KG(true, 1, true); else DAV_GEMM_KG(true, 2, true);
    } else if (transA) {
      if (wns == 1) DAV_GEMM_KG(true, 1, false); else DAV_GEMM_KG(true, 2, false);
    } else {
      if (wns == 1) DAV_GEMM_KG(false, 1, false); else DAV_GEMM_KG(false, 2, false);
    }
#undef DAV_GEMM_KG
#undef DAV_GEMM_CASE
  } else if (transA) {
    gemm_kernel<true><<<grid, NT, 0, s>>>(M, N, K, Kchunk, alpha, A, lda, B, ldb, beta, C, ldc, ws, splits, to_ws);
  } else {
    gemm_kernel<false><<<grid, NT, 0, s>>>(M, N, K, Kchunk, alpha, A, lda, B, ldb, beta, C, ldc, ws, splits, to_ws);
  }
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
