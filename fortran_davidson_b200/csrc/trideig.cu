// Symmetric k x k eigenproblem for the Rayleigh-Ritz step (replaces DSYEV / the inner solves of DSYGV,
// lapack_wrapper.f90:14-91; call sites davidson.f90:153,155,394) -- the fast path for k >= ~48.
//
// The one-CTA Jacobi of smalldense.cu needs ~8 sweeps x (k-1) barrier-separated rounds: 3.0 ms at k = 128 and
// ~110 ms at k = 256, which is what bounds strong scaling once the block matvec is spread over 8 GPUs.  Here:
//   1. tridiag_kernel     one CTA, Householder tridiagonalisation S = Q T Q^T (k-2 steps, S in shared memory when it
//                         fits, otherwise L2-resident), reflectors kept for step 3
//   2. tri_eigvec_kernel  ONE WARP PER EIGENPAIR, k warps spread over the SMs, no communication between them:
//                         eigenvalue j by 32-way multisection on the Sturm count (each lane runs the count at its own
//                         shift), eigenvector of T by the twisted factorisation (forward and backward pivot
//                         recurrences in lanes 0 / 1 in lock-step), back-transformation by the k-2 reflectors with
//                         the vector held in registers (rows strided over the lanes)
//   3. guard              G = Y^T Y; one Newton-Schulz step Y <- Y (1.5 I - 0.5 G) squares the loss of orthogonality
//                         that independent eigenvector computations leave between close eigenvalues; Rayleigh
//                         quotients theta_j = y_j^T S y_j; the result is ACCEPTED only if max|G - I| <= 3e-8 (so that
//                         the corrected basis is orthonormal to ~1e-15) and max|S y - theta y| <= 64 k eps max|S|.
//   4. otherwise          (exactly degenerate / pathologically clustered spectra, NaN input) the Jacobi kernels run
//                         as before; when the fast path was accepted they return at once (device-side flag, no host
//                         synchronisation).
#include <algorithm>
#include <cstdlib>

#include "kernels.cuh"

namespace dav {
namespace {

constexpr double EPS = 2.220446049250313e-16;

// ---- 1. Householder tridiagonalisation (DSYTD2 on the full symmetric matrix) ----------------------------------
// Vh(:, j) = reflector j with absolute row indexing (rows <= j are 0, row j+1 is 1); tau[j] = 0 for H_j = I.
// Two block barriers per step: (1) p = tau S v as column dot products, one warp per column (S is symmetric, so
// column i serves as row i; contiguous shared-memory reads, warp-shuffle reduction), (2) the rank-2 update
// S -= v w^T + w v^T with w = p - (tau/2)(p.v) v formed on the fly; warp 0 owns the first trailing column and
// builds the NEXT reflector from it while the other warps finish their columns.

// warp 0 only: reflector for column jn of S (rows jn+1 .. k-1) -> vout (relative indexing), Vh, tau, d, e, ts
__device__ __forceinline__ void make_reflector(int k, int ld, int jn, const double* S, double* vout,
                                               double* __restrict__ Vh, double* __restrict__ tau,
                                               double* __restrict__ d, double* __restrict__ e, double* ts) {
  const int lane = threadIdx.x & 31;
  const int m = k - jn - 1;
  const double* col = S + (size_t)jn * ld + (jn + 1);
  double s2 = 0.0;
  for (int i = 1 + lane; i < m; i += 32) s2 = fma(col[i], col[i], s2);
  const double sigma = warp_sum(s2);
  const double alpha = col[0];
  double t = 0.0, beta = alpha, scale = 0.0;
  if (sigma > 0.0) {
    // beta = -sign(alpha) sqrt(alpha^2 + sigma); tau = (beta - alpha) / beta = 1 + |alpha| / |beta|;
    // scale = 1 / (alpha - beta) = sign(alpha) / (|alpha| + |beta|): one rsqrt and one reciprocal, no division
    const double h2 = fma(alpha, alpha, sigma);
    const double rs = rsqrt(h2);
    const double nrm = h2 * rs;
    beta = -copysign(nrm, alpha);
    t = fma(fabs(alpha), rs, 1.0);
    scale = copysign(__drcp_rn(fabs(alpha) + nrm), alpha);
  }
  for (int i = lane; i < m; i += 32) {
    const double vi = (i == 0) ? 1.0 : (t == 0.0 ? 0.0 : col[i] * scale);
    vout[i] = vi;
    Vh[(size_t)jn * k + (jn + 1 + i)] = vi;
  }
  if (lane == 0) {
    ts[0] = t;
    tau[jn] = t;
    d[jn] = S[jn + (size_t)jn * ld];
    e[jn] = beta;
  }
}

template <int RB>  // row blocks of 32 kept in registers during the rank-2 update
__global__ void __launch_bounds__(1024) tridiag_kernel(int k, const double* __restrict__ S_in, int64_t lds_in, double* gS,
                                                       int s_in_smem, double* __restrict__ Sfull,
                                                       double* __restrict__ Vh, double* __restrict__ tau,
                                                       double* __restrict__ d, double* __restrict__ e,
                                                       double* __restrict__ scal) {
  extern __shared__ __align__(16) double sm[];
  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nw = nt >> 5;
  const int ld = k;
  double* red = sm;             // 32
  double* tsb = red + 32;       // 2 (tau of the current / next reflector), padded to 4
  double* vb = tsb + 4;         // 2 x k
  double* p = vb + 2 * k;       // k
  double* part = p + k;         // column-chunk partials of S v: nch * m <= 32 * nw doubles
  double* S = s_in_smem ? part + (size_t)nw * 32 : gS;
  // (row group, column chunk) of every warp for each possible number of row groups: no divisions in the step loop
  __shared__ unsigned short map_tab[17 * 32];
  __shared__ int nch_tab[17];
  for (int q = tid; q < 17 * 32; q += nt) {
    const int rwq = max(1, q >> 5), wq = q & 31;
    map_tab[q] = (unsigned short)((wq % rwq) | ((wq / rwq) << 8));
    if (wq == 0) nch_tab[q >> 5] = max(1, nw / rwq);
  }

  // symmetrise from the upper triangle (DSYEV 'U'), keep a full copy for the guard
  double mx = 0.0;
  for (int idx = tid; idx < k * k; idx += nt) {
    const int i = idx % k, j = idx / k;
    const double x = (i <= j) ? S_in[i + (size_t)j * lds_in] : S_in[j + (size_t)i * lds_in];
    S[i + (size_t)j * ld] = x;
    Sfull[idx] = x;
    Vh[idx] = 0.0;
    mx = fmax(mx, fabs(x));  // NaN is dropped here and caught by the guard
  }
  for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  __syncthreads();
  if (lane == 0) red[warp] = mx;
  __syncthreads();
  if (tid == 0) {
    double m = 0.0;
    for (int i = 0; i < nw; ++i) m = fmax(m, red[i]);
    scal[0] = m;  // max |S|
  }
  if (warp == 0 && k > 2) make_reflector(k, ld, 0, S, vb, Vh, tau, d, e, tsb);
  __syncthreads();

#ifdef DAV_TRIDIAG_PROFILE
  long long cyc[4] = {0, 0, 0, 0};
  long long c0 = clock64();
#endif
  for (int j = 0; j + 2 < k; ++j) {
    const int m = k - j - 1;  // trailing size; rows/cols j+1 .. k-1
    const double* v = vb + (j & 1) * k;
    const double t = tsb[j & 1];
    double* St = S + (size_t)(j + 1) * ld + (j + 1);
    if (t != 0.0) {  // uniform
      // p = t S v without warp shuffles (SHFL issues one warp per clock per SM, which bounded the column-dot
      // form): lanes over rows, warps over (row group, column chunk), four independent FMA chains per thread;
      // the chunk partials go through shared memory.  p.v = t v^T S v is accumulated on the way.
      const int rw = (m + 31) >> 5;        // row groups of 32
      const int nch = nch_tab[rw];         // column chunks = max(1, nw / rw), columns strided over the chunks
      const int rg = map_tab[rw * 32 + warp] & 0xff, ch = map_tab[rw * 32 + warp] >> 8;
      double pvpart = 0.0;
      if (ch < nch) {
        const int i = rg * 32 + lane;
        if (i < m) {
          double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
          const size_t cs = (size_t)nch * ld;
          const double* sp = St + i + (size_t)ch * ld;
          int c = ch;
          for (; c + 3 * nch < m; c += 4 * nch, sp += 4 * cs) {
            a0 = fma(sp[0], v[c], a0);
            a1 = fma(sp[cs], v[c + nch], a1);
            a2 = fma(sp[2 * cs], v[c + 2 * nch], a2);
            a3 = fma(sp[3 * cs], v[c + 3 * nch], a3);
          }
          for (; c < m; c += nch, sp += cs) a0 = fma(sp[0], v[c], a0);
          const double acc = (a0 + a1) + (a2 + a3);
          part[ch * m + i] = acc;
          pvpart = acc * v[i];
        }
      }
      pvpart = warp_sum(pvpart);
      if (lane == 0) red[warp] = pvpart;
#ifdef DAV_TRIDIAG_PROFILE
      { const long long c1 = clock64(); cyc[0] += c1 - c0; c0 = c1; }
#endif
      __syncthreads();
      for (int i = tid; i < m; i += nt) {
        double b0 = 0.0, b1 = 0.0;
        int c = 0;
        for (; c + 1 < nch; c += 2) {
          b0 += part[c * m + i];
          b1 += part[(c + 1) * m + i];
        }
        if (c < nch) b0 += part[c * m + i];
        p[i] = (b0 + b1) * t;
      }
      double q0 = 0.0, q1 = 0.0, q2 = 0.0, q3 = 0.0;  // same order in every thread
      for (int q = 0; q + 3 < nw; q += 4) {
        q0 += red[q];
        q1 += red[q + 1];
        q2 += red[q + 2];
        q3 += red[q + 3];
      }
      const double pv = t * ((q0 + q1) + (q2 + q3));
      const double K = -0.5 * t * pv;
      __syncthreads();
#ifdef DAV_TRIDIAG_PROFILE
      { const long long c1 = clock64(); cyc[1] += c1 - c0; c0 = c1; }
#endif
      // S_trail -= v w^T + w v^T.  Warp 0 updates only trailing column 0 (the next reflector's column) and then
      // builds that reflector; warps 1.. share the other columns.
      for (int r0 = 0; r0 < m; r0 += 32 * RB) {
        double vi[RB], wi[RB];
#pragma unroll
        for (int r = 0; r < RB; ++r) {
          const int i = r0 + lane + 32 * r;
          vi[r] = i < m ? v[i] : 0.0;
          wi[r] = i < m ? fma(K, vi[r], p[i]) : 0.0;
        }
        const int cbeg = warp == 0 ? 0 : warp, cend = warp == 0 ? 1 : m, cstep = nw - 1;
        for (int c = cbeg; c < cend; c += cstep) {
          const double vc = v[c], wc = fma(K, vc, p[c]);
          double* sp = St + (size_t)c * ld + r0 + lane;
#pragma unroll
          for (int r = 0; r < RB; ++r) {
            const int i = r0 + lane + 32 * r;
            if (i < m) sp[32 * r] = sp[32 * r] - vi[r] * wc - wi[r] * vc;
          }
        }
      }
    }
    // warp 0 owns trailing column 0 (= column j+1 of S): next reflector while the others finish
    if (warp == 0 && j + 3 < k) {
      __syncwarp();
      make_reflector(k, ld, j + 1, S, vb + ((j + 1) & 1) * k, Vh, tau, d, e, tsb + ((j + 1) & 1));
    }
#ifdef DAV_TRIDIAG_PROFILE
    { const long long c1 = clock64(); cyc[2] += c1 - c0; c0 = c1; }
#endif
    __syncthreads();
#ifdef DAV_TRIDIAG_PROFILE
    { const long long c1 = clock64(); cyc[3] += c1 - c0; c0 = c1; }
#endif
  }
#ifdef DAV_TRIDIAG_PROFILE
  if (tid == 0) for (int q = 0; q < 4; ++q) scal[3 + q] = (double)cyc[q];
#endif
  if (tid == 0) {
    if (k >= 2) {
      d[k - 2] = S[(k - 2) + (size_t)(k - 2) * ld];
      e[k - 2] = S[(k - 1) + (size_t)(k - 2) * ld];
      tau[k - 2] = 0.0;
    }
    d[k - 1] = S[(k - 1) + (size_t)(k - 1) * ld];
    tau[k - 1] = 0.0;
  }
}

// ---- 1b. the same tridiagonalisation with the matrix in REGISTERS (k <= 128: every Rayleigh-Ritz problem of the
// headline solve).  The shared-memory kernel above spends ~35 issued instructions per matrix entry and step on
// addressing, bounds and LDS/STS (17.7 k warp-instructions per step at k = 128: 2.8 us per step, 353 us in total);
// here the full symmetric matrix lives in a 16 x 16 grid of threads, thread (ty, tx) holding the TS x TS tile of rows
// TS ty.. and columns TS tx.. (TS = 2 / 4 / 8 for k <= 32 / 64 / 128), and a step is
//   reflector scalars, v at my rows / columns       from the published row j (shared memory), redundantly per thread
//   p = tau S v                                      TS^2 FMAs + a reduce-scatter over the 16 lanes of a row of tiles
//   barrier 1
//   K = -tau/2 (p.v)                                 16-lane all-reduce, redundantly per half warp
//   S -= v w^T + w v^T,  w = p + K v                 2 TS^2 FMAs, registers only, tiles left of / above the front skip
//   owners of row j+1 publish it + its tail norm     barrier 2
// Dead rows / columns (<= j) are masked out of w, so they are never touched again; d, e, tau and the reflectors go to
// global memory from the published row.  Rounding differs from the kernel above only in summation order.
template <int TS>
__device__ __forceinline__ int reduce_scatter16(double (&v)[TS], int tx) {
  // sum over the 16 lanes of a half warp; afterwards v[0] of lane tx is the total of entry `return value`
  int idx = 0;
  if constexpr (TS == 8) {
    const bool up = tx & 8;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const double send = up ? v[q] : v[q + 4];
      const double keep = up ? v[q + 4] : v[q];
      v[q] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
    idx += up ? 4 : 0;
  }
  if constexpr (TS >= 4) {
    constexpr int M = TS == 8 ? 4 : 8;
    const bool up = tx & M;
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const double send = up ? v[q] : v[q + 2];
      const double keep = up ? v[q + 2] : v[q];
      v[q] = keep + __shfl_xor_sync(0xffffffffu, send, M);
    }
    idx += up ? 2 : 0;
  }
  {
    constexpr int M = TS == 8 ? 2 : (TS == 4 ? 4 : 8);
    const bool up = tx & M;
    const double send = up ? v[0] : v[1];
    const double keep = up ? v[1] : v[0];
    v[0] = keep + __shfl_xor_sync(0xffffffffu, send, M);
    idx += up ? 1 : 0;
#pragma unroll
    for (int m2 = M >> 1; m2 >= 1; m2 >>= 1) v[0] += __shfl_xor_sync(0xffffffffu, v[0], m2);
  }
  return idx;
}

__device__ __forceinline__ double allreduce16(double x) {
#pragma unroll
  for (int m = 8; m >= 1; m >>= 1) x += __shfl_xor_sync(0xffffffffu, x, m);
  return x;
}

template <int TSR, int TS>  // tile = TSR rows x TS columns; 16 TS / TSR x 16 threads
__global__ void __launch_bounds__(256 * TS / TSR) tridiag_reg_kernel(int k, const double* __restrict__ S_in, int64_t lds_in,
                                                          double* __restrict__ Sfull, double* __restrict__ Vh,
                                                          double* __restrict__ tau, double* __restrict__ d,
                                                          double* __restrict__ e, double* __restrict__ scal) {
  constexpr int KP = 16 * TS;
  __shared__ __align__(16) double rowb[KP];  // row j of the current matrix (owners only)
  __shared__ __align__(16) double vb[KP];    // reflector j: 0 for rows <= j, 1 at row j+1
  __shared__ __align__(16) double pb[KP];    // p = tau S v, exactly 0 on dead rows
  __shared__ double tjb;                     // tau of reflector j
  __shared__ double red[8 * TS / TSR];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4, warp = tid >> 5;
  // rows of a tile are contiguous (TS ty ..); its columns are the pairs 2 tx, 2 tx + 1 of every 32-column group, so
  // that the 16 lanes of a half warp read 256 contiguous bytes of v / p with each 16-byte load (conflict-free)
  const int r0 = ty * TSR, c0 = 2 * tx;
  auto col = [&](int q) { return c0 + 32 * (q >> 1) + (q & 1); };
  double t[TSR][TS];
  double mx = 0.0;
#pragma unroll
  for (int i = 0; i < TSR; ++i)
#pragma unroll
    for (int jj = 0; jj < TS; ++jj) {
      const int a = r0 + i, c = col(jj);
      double x = 0.0;
      if (a < k && c < k) {
        x = S_in[min(a, c) + (size_t)max(a, c) * lds_in];  // DSYEV 'U': only the upper triangle is read
        // the symmetrised copy for the guard, written at the MIRRORED position: the 16 lanes of a tile row then
        // store 256 contiguous bytes (entry (c, a) = entry (a, c))
        Sfull[c + (size_t)a * k] = x;
      }
      t[i][jj] = x;
      mx = fmax(mx, fabs(x));  // NaN is dropped here and caught by the guard
    }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((tid & 31) == 0) red[warp] = mx;
  for (int i = tid; i < KP; i += blockDim.x) pb[i] = 0.0;
  // reflector columns k-2, k-1 do not exist (the loop writes every other column of Vh completely)
  for (int i = tid; i < 2 * k; i += blockDim.x)
    if (k >= 2) Vh[(size_t)(k - 2) * k + i] = 0.0;

  // The warp that owns row jn builds reflector jn from it: v -> vb (and Vh), tau -> tjb (and tau[]), d, e.  The
  // step loop keeps its code small on purpose (one SM's instruction cache serves 8 warps in lock step): the row is
  // picked with a switch on jn % TS, and everyone else only ever loads finished vectors.
  auto publish = [&](int jn) {
    if ((ty >> 1) != ((jn / TSR) >> 1)) return;  // warp-uniform: the shuffles below need the whole warp
    const bool own = ty == jn / TSR;
    double x[TS];
    switch (jn % TSR) {
#define TRI_ROW(I_)                                   \
  case I_:                                            \
    if constexpr (I_ < TSR) {                         \
      _Pragma("unroll") for (int jj = 0; jj < TS; ++jj) x[jj] = t[I_][jj]; \
    }                                                 \
    break;
      TRI_ROW(0) TRI_ROW(1) TRI_ROW(2) TRI_ROW(3) TRI_ROW(4) TRI_ROW(5) TRI_ROW(6) TRI_ROW(7)
#undef TRI_ROW
      default: break;
    }
    double ssq = 0.0;
#pragma unroll
    for (int jj = 0; jj < TS; ++jj) {
      if (own) rowb[col(jj)] = x[jj];
      if (col(jj) > jn + 1) ssq = fma(x[jj], x[jj], ssq);
    }
    const double sigma = allreduce16(ssq);
    __syncwarp();
    const double alpha = rowb[min(jn + 1, KP - 1)];
    double tj = 0.0, beta = alpha, scale = 0.0;
    if (sigma > 0.0) {
      // beta = -sign(alpha) sqrt(alpha^2 + sigma); tau = (beta - alpha) / beta = 1 + |alpha| / |beta|;
      // scale = 1 / (alpha - beta) = sign(alpha) / (|alpha| + |beta|): one rsqrt and one reciprocal, no division
      const double h2 = fma(alpha, alpha, sigma);
      const double rs = rsqrt(h2);
      const double nrm = h2 * rs;
      beta = -copysign(nrm, alpha);
      tj = fma(fabs(alpha), rs, 1.0);
      scale = copysign(__drcp_rn(fabs(alpha) + nrm), alpha);
    }
    if (own) {
#pragma unroll
      for (int jj = 0; jj < TS; ++jj) {
        const int c = col(jj);
        const double v = c == jn + 1 ? 1.0 : ((c > jn + 1 && tj != 0.0) ? x[jj] * scale : 0.0);
        vb[c] = v;
        if (c < k && jn + 2 < k) Vh[(size_t)jn * k + c] = v;
      }
      if (tx == 0 && jn + 2 < k) {
        tjb = tj;
        tau[jn] = tj;
        d[jn] = rowb[jn];
        e[jn] = beta;
      }
    }
  };
  publish(0);
  __syncthreads();
  if (tid == 0) {
    double m = 0.0;
    for (int i = 0; i < 8 * TS / TSR; ++i) m = fmax(m, red[i]);
    scal[0] = m;  // max |S|
  }

#ifdef DAV_TRIDIAG_PROFILE
  long long cyc[4] = {0, 0, 0, 0};
  long long ck0 = clock64();
#define TRI_TICK(q) { const long long c1 = clock64(); if ((q) >= 0) cyc[(q) & 3] += c1 - ck0; ck0 = c1; }
#if DAV_TRIDIAG_PROFILE == 1
#define TRI_A(q) TRI_TICK(q)
#define TRI_B(q)
#else
#define TRI_A(q)
#define TRI_B(q) TRI_TICK(q)
#endif
#else
#define TRI_TICK(q)
#define TRI_A(q)
#define TRI_B(q)
#endif
  for (int j = 0; j + 2 < k; ++j) {
    const double tj = tjb;
    const int last_row = ((ty | 1) + 1) * TSR - 1;  // last row of this warp
    const bool warp_live = last_row > j;           // some row of this warp is still in the trailing matrix
    const bool tile_live = r0 + TSR - 1 > j;  // (the interleaved columns of a tile stay live until the last steps)
    double vc[TS], vr[TSR];
    if (tj != 0.0 && warp_live) {  // uniform per warp
#pragma unroll
      for (int q = 0; q < TS; q += 2) {
        const double2 a = *reinterpret_cast<const double2*>(vb + col(q));
        vc[q] = a.x; vc[q + 1] = a.y;
      }
#pragma unroll
      for (int q = 0; q < TSR; q += 2) {
        const double2 b = *reinterpret_cast<const double2*>(vb + r0 + q);
        vr[q] = b.x; vr[q + 1] = b.y;
      }
      // ---- p = tau S v: row sums of my tile, reduce-scatter over the 16 tiles of the row
      double pt[TSR];
#pragma unroll
      for (int i = 0; i < TSR; ++i) pt[i] = 0.0;
      if (tile_live) {
#pragma unroll
        for (int i = 0; i < TSR; ++i)
#pragma unroll
          for (int jj = 0; jj < TS; ++jj) pt[i] = fma(t[i][jj], vc[jj], pt[i]);
      }
      TRI_A(0)
      const int idx = reduce_scatter16<TSR>(pt, tx);
      constexpr int WMASK = TSR == 8 ? 1 : (TSR == 4 ? 3 : 7);
      if ((tx & WMASK) == 0) pb[r0 + idx] = (r0 + idx > j) ? pt[0] * tj : 0.0;
    } else if (tj != 0.0 && last_row == j) {
      if (tx < TSR) pb[r0 + tx] = 0.0;  // this warp's rows just died: p stays exactly 0 there from now on
    }
    TRI_A(1)
    TRI_B(3)
    __syncthreads();
    TRI_A(2)
    TRI_B(3)
    if (tj != 0.0 && warp_live) {
      double pc[TS], pr[TSR];
      double pvp = 0.0;
#pragma unroll
      for (int q = 0; q < TS; q += 2) {
        const double2 a = *reinterpret_cast<const double2*>(pb + col(q));
        pc[q] = a.x; pc[q + 1] = a.y;
      }
#pragma unroll
      for (int q = 0; q < TSR; q += 2) {
        const double2 b = *reinterpret_cast<const double2*>(pb + r0 + q);
        pr[q] = b.x; pr[q + 1] = b.y;
      }
#pragma unroll
      for (int q = 0; q < TS; ++q) pvp = fma(pc[q], vc[q], pvp);
      const double pv = allreduce16(pvp);
      const double K = -0.5 * tj * pv;
      TRI_B(0)
      if (tile_live) {
        // w = p + K v is 0 wherever v and p are (dead rows / columns): those entries are never touched again
#pragma unroll
        for (int q = 0; q < TS; ++q) pc[q] = fma(K, vc[q], pc[q]);
#pragma unroll
        for (int q = 0; q < TSR; ++q) pr[q] = fma(K, vr[q], pr[q]);
#pragma unroll
        for (int i = 0; i < TSR; ++i)
#pragma unroll
          for (int jj = 0; jj < TS; ++jj) t[i][jj] = fma(-pr[i], vc[jj], fma(-vr[i], pc[jj], t[i][jj]));
      }
    }
    TRI_B(1)
    publish(j + 1);
    TRI_B(2)
    __syncthreads();
    TRI_A(3)
    TRI_B(3)
  }
#ifdef DAV_TRIDIAG_PROFILE
  if (tid == (int)blockDim.x - 1) for (int q = 0; q < 4; ++q) scal[3 + q] = (double)cyc[q];
#endif
#undef TRI_TICK
#undef TRI_A
#undef TRI_B
  // the last 2 x 2 block and the trivial reflectors
#pragma unroll
  for (int i = 0; i < TSR; ++i)
#pragma unroll
    for (int jj = 0; jj < TS; ++jj) {
      const int a = r0 + i, c = col(jj);
      if (k >= 2 && a == k - 2 && c == k - 2) { d[k - 2] = t[i][jj]; tau[k - 2] = 0.0; }
      if (k >= 2 && a == k - 2 && c == k - 1) e[k - 2] = t[i][jj];
      if (a == k - 1 && c == k - 1) { d[k - 1] = t[i][jj]; tau[k - 1] = 0.0; }
    }
}

// ---- 2. one warp per eigenpair ----------------------------------------------------------------------------------
// warps per CTA.  r02: 4 instead of 8 -- the Sturm / twisted recurrences are chains of FP64 instructions, one warp
// issues one every ~4 cycles, and two such warps per scheduler already halve each other's speed (k = 128: 16 -> 32 CTAs)
constexpr int EW = 4;

template <int KT>  // rows per lane: k <= 32 * KT
__global__ void __launch_bounds__(EW * 32) tri_eigvec_kernel(int k, const double* __restrict__ d_in,
                                                            const double* __restrict__ e_in,
                                                            const double* __restrict__ Vh,
                                                            const double* __restrict__ tau, double* __restrict__ Y,
                                                            double* __restrict__ lam_out, int vh_smem,
                                                            double* prof) {
  extern __shared__ __align__(16) double sm[];
#ifdef DAV_TRIDIAG_PROFILE
  long long pt0 = clock64(), pt1 = 0, pt2 = 0, pt3 = 0;
#endif
  // T is scaled to unit norm (ds = d / tn, es = e / tn): counts and eigenvectors are scale invariant, and the
  // characteristic-polynomial recurrence below can then grow by at most 3x per step
  double* ds = sm;           // k
  double* es = ds + k;       // k (es[k-1] = 0)
  double* e2s = es + k;      // k
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double* qp = e2s + k + (size_t)warp * 5 * k;  // forward pivots
  double* qm = qp + k;                          // backward pivots
  double* z = qm + k;
  double* fa = z + k;                           // forward denominators, then the upward coefficients
  double* fb = fa + k;                          // backward denominators, then the downward coefficients
  // r02: the k x k reflector block staged in shared memory when it fits (k <= 128), copied asynchronously while the
  // eigenvalue is being located.  From global memory every one of the k-2 back-transformation steps waited for an
  // L2 round trip (one reflector of prefetch hides ~180 of ~700 cycles): 109 us at k = 128, of which ~60 us latency.
  double* vhs = e2s + k + (size_t)EW * 5 * k;
  if (vh_smem) {
    const unsigned sbase = (unsigned)__cvta_generic_to_shared(vhs);
    for (int idx = threadIdx.x; idx < k * k; idx += blockDim.x)
      asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sbase + 8u * (unsigned)idx), "l"(Vh + idx) : "memory");
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  // Gershgorin interval and norm (every warp computes the same values)
  double gl = 1.0e300, gu = -1.0e300;
  for (int i = lane; i < k; i += 32) {
    const double di = d_in[i];
    const double r = (i < k - 1 ? fabs(e_in[i]) : 0.0) + (i > 0 ? fabs(e_in[i - 1]) : 0.0);
    gl = fmin(gl, di - r);
    gu = fmax(gu, di + r);
  }
  for (int o = 16; o > 0; o >>= 1) {
    gl = fmin(gl, __shfl_xor_sync(0xffffffffu, gl, o));
    gu = fmax(gu, __shfl_xor_sync(0xffffffffu, gu, o));
  }
  const double tn = fmax(fabs(gl), fabs(gu));
  const double itn = (tn > 0.0 && tn < 1.0e300) ? 1.0 / tn : 1.0;
  for (int i = threadIdx.x; i < k; i += blockDim.x) {
    ds[i] = d_in[i] * itn;
    const double ei = (i < k - 1) ? e_in[i] * itn : 0.0;
    es[i] = ei;
    e2s[i] = ei * ei;
  }
  __syncthreads();
  const int jraw = blockIdx.x * EW + warp;  // eigenvalue index (ascending)
  const bool valid = jraw < k;              // idle warps of the last CTA compute a duplicate and store nothing
  const int j = valid ? jraw : k - 1;

  // ---- eigenvalue j by 32-way multisection.  count(x) = #{eigenvalues < x} = sign changes of the Sturm sequence
  // p_0 = 1, p_1 = d_0 - x, p_{i+1} = (d_i - x) p_i - e_{i-1}^2 p_{i-1}  (one dependent FMA per step; a zero takes
  // the sign opposite to its predecessor, like the pivot form's q = -pivmin).  Invariant: count(lo) <= j < count(hi).
  double lo = gl * itn - 4.0 * EPS * k - 1.0e-300, hi = gu * itn + 4.0 * EPS * k + 1.0e-300;
  for (int it = 0; it < 48; ++it) {
    const double width = hi - lo;
    if (!(width > fmax(4.0 * EPS * fmax(fabs(lo), fabs(hi)), 1.0e-3 * EPS))) break;
    const double x = lo + width * ((double)(lane + 1) * (1.0 / 33.0));
    double pm = 1.0, pc = ds[0] - x;
    bool neg = !(pc > 0.0);
    int cnt = neg;
    // r02: blocks of 8 steps with the NEXT block's coefficients already in registers -- the recurrence is one
    // dependent FMA per step, a shared-memory load inside the chain would triple its latency
    int i = 1;
    double dn[8], en[8];
    if (i + 8 <= k) {
#pragma unroll
      for (int q = 0; q < 8; ++q) { dn[q] = ds[i + q]; en[q] = e2s[i + q - 1]; }
    }
    for (; i + 8 <= k; i += 8) {
      double dc[8], ec[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) { dc[q] = dn[q]; ec[q] = en[q]; }
      if (i + 16 <= k) {
#pragma unroll
        for (int q = 0; q < 8; ++q) { dn[q] = ds[i + 8 + q]; en[q] = e2s[i + 8 + q - 1]; }
      }
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const double pn = fma(dc[q] - x, pc, -(ec[q] * pm));
        // sign tests on the integer pipe (the FP64 pipe is what this loop is bound by): a zero takes the sign
        // opposite to its predecessor (NaN input is caught by the guard, whatever it counts as here)
        const int hi = __double2hiint(pn);
        const bool zero = ((hi & 0x7fffffff) | __double2loint(pn)) == 0;
        const bool nneg = zero ? !neg : (hi < 0);
        cnt += nneg != neg;
        neg = nneg;
        pm = pc;
        pc = pn;
      }
      {  // keep the pair inside the exponent range
        const double mag = fmax(fabs(pc), fabs(pm));
        const double f = mag > 1.0e100 ? 1.0e-100 : (mag < 1.0e-100 ? 1.0e100 : 1.0);
        pc *= f;
        pm *= f;
      }
    }
    for (; i < k; ++i) {
      const double pn = fma(ds[i] - x, pc, -(e2s[i - 1] * pm));
      const bool nneg = (pn == 0.0) ? !neg : (pn < 0.0);
      cnt += nneg != neg;
      neg = nneg;
      pm = pc;
      pc = pn;
    }
    const unsigned ok = __ballot_sync(0xffffffffu, cnt >= j + 1);
    const int first = ok ? (__ffs(ok) - 1) : 32;
    const double xhi = __shfl_sync(0xffffffffu, x, min(first, 31));
    const double xlo = __shfl_sync(0xffffffffu, x, max(first - 1, 0));
    if (first < 32) hi = xhi;
    if (first > 0) lo = xlo;
  }
  const double lam = 0.5 * (lo + hi);  // scaled
#ifdef DAV_TRIDIAG_PROFILE
  pt1 = clock64();
#endif

  // ---- eigenvector of T: twisted factorisation.  Lane 0 runs the forward sequence (top down), lane 1 the backward
  // one (bottom up), one instruction stream for both.  The pivots q_i = p_i / p_{i-1} are NOT formed inside the
  // sequential loop (a reciprocal there costs ~5 dependent FP64 instructions per step): the loop carries the
  // division-free Sturm pair and stores (numerator, denominator); all lanes divide afterwards.
  const double tiny = EPS;
  if (lane < 2) {
    const int dir = lane == 0 ? 1 : -1;
    int i = lane == 0 ? 0 : k - 1;
    double* num = lane == 0 ? qp : qm;
    double* den = lane == 0 ? fa : fb;
    double pm = 1.0, pc = ds[i] - lam;
    num[i] = pc;
    den[i] = 1.0;
    for (int s = 1; s < k; ++s) {
      const double ee = lane == 0 ? e2s[i] : e2s[i - 1];
      i += dir;
      const double pn = fma(ds[i] - lam, pc, -(ee * pm));
      num[i] = pn;
      den[i] = pc;
      pm = pc;
      pc = pn;
      if ((s & 7) == 0) {
        const double mag = fmax(fabs(pc), fabs(pm));
        const double f = mag > 1.0e100 ? 1.0e-100 : (mag < 1.0e-100 ? 1.0e100 : 1.0);
        pc *= f;
        pm *= f;
      }
    }
  }
  __syncwarp();
  // pivots, gamma_i = q+_i + q-_i - (d_i - lam), r = argmin |gamma_i|
  double best = 1.0e308 * 10.0;
  int r = 0;
  for (int i = lane; i < k; i += 32) {
    double a = qp[i] / fa[i], b = qm[i] / fb[i];
    if (!(fabs(a) >= tiny)) a = (a < 0.0) ? -tiny : tiny;   // zero / NaN pivots -> +-eps |T|
    if (!(fabs(b) >= tiny)) b = (b < 0.0) ? -tiny : tiny;
    qp[i] = a;
    qm[i] = b;
    const double g = fabs(a + b - (ds[i] - lam));
    if (g < best) { best = g; r = i; }
  }
  for (int o = 16; o > 0; o >>= 1) {
    const double ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int orr = __shfl_xor_sync(0xffffffffu, r, o);
    if (ob < best || (ob == best && orr < r)) { best = ob; r = orr; }
  }
  __syncwarp();
  // recurrence coefficients: up z_i = cu_i z_{i+1}, cu_i = -e_i / q+_i; down z_{i+1} = cd_{i+1} z_i, cd_{i+1} = -e_i / q-_{i+1}
  for (int i = lane; i < k; i += 32) {
    fa[i] = (i < k - 1) ? -es[i] / qp[i] : 0.0;
    fb[i] = (i > 0) ? -es[i - 1] / qm[i] : 0.0;
    z[i] = 0.0;
  }
  __syncwarp();
  if (lane < 2) {
    if (lane == 0) z[r] = 1.0;
    double zc = 1.0;
    const int steps = lane == 0 ? r : k - 1 - r;
    const int nsteps = max(r, k - 1 - r);
    const double* cf = lane == 0 ? fa : fb;
    for (int s = 0; s < nsteps; ++s) {
      if (s < steps) {
        const int iq = lane == 0 ? r - 1 - s : r + 1 + s;
        zc *= cf[iq];
        z[iq] = zc;
      }
    }
  }
  __syncwarp();
  // normalise, move to registers (absolute row a = lane + 32 t)
  double zr[KT];
  double n2 = 0.0;
#pragma unroll
  for (int t = 0; t < KT; ++t) {
    const int a = lane + 32 * t;
    zr[t] = a < k ? z[a] : 0.0;
    n2 = fma(zr[t], zr[t], n2);
  }
  n2 = warp_sum(n2);
  const double inv = rsqrt(n2);
#pragma unroll
  for (int t = 0; t < KT; ++t) zr[t] *= inv;

#ifdef DAV_TRIDIAG_PROFILE
  pt2 = clock64();
#endif
  // ---- back-transformation y = H_0 H_1 ... H_{k-3} z, reflectors applied last to first
  if (vh_smem) {
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();
  }
#ifdef DAV_TRIDIAG_PROFILE
  pt3 = clock64();
#endif
  const double* VhP = vh_smem ? vhs : Vh;
  double vc[KT], vn[KT];
  int jj = k - 3;
#pragma unroll
  for (int t = 0; t < KT; ++t) {
    const int a = lane + 32 * t;
    vc[t] = (jj >= 0 && a < k) ? VhP[(size_t)jj * k + a] : 0.0;
  }
  for (; jj >= 0; --jj) {
#pragma unroll
    for (int t = 0; t < KT; ++t) {
      const int a = lane + 32 * t;
      vn[t] = (jj >= 1 && a < k) ? VhP[(size_t)(jj - 1) * k + a] : 0.0;  // prefetch the next reflector
    }
    const double tj = tau[jj];
    double dot = 0.0;
#pragma unroll
    for (int t = 0; t < KT; ++t) dot = fma(vc[t], zr[t], dot);
    dot = warp_sum(dot) * tj;
#pragma unroll
    for (int t = 0; t < KT; ++t) {
      zr[t] = fma(-dot, vc[t], zr[t]);
      vc[t] = vn[t];
    }
  }
#ifdef DAV_TRIDIAG_PROFILE
  if (prof && blockIdx.x == gridDim.x / 2 && threadIdx.x == 0) {
    const long long pt4 = clock64();
    prof[0] = (double)(pt1 - pt0); prof[1] = (double)(pt2 - pt1); prof[2] = (double)(pt3 - pt2); prof[3] = (double)(pt4 - pt3);
  }
#endif
  if (!valid) return;
#pragma unroll
  for (int t = 0; t < KT; ++t) {
    const int a = lane + 32 * t;
    if (a < k) Y[(size_t)j * k + a] = zr[t];
  }
  if (lane == 0) lam_out[j] = lam * tn;
}

// ---- small k x k products of the guard: C = op(A) * B, 32 x 32 tiles so that even k = 64 spreads over 4 SMs ----
template <bool TA>
__global__ void __launch_bounds__(256) small_gemm_kernel(int k, const double* __restrict__ A,
                                                         const double* __restrict__ B, double* __restrict__ C) {
  __shared__ double As[32][33];  // As[kk][m]
  __shared__ double Bs[32][33];  // Bs[kk][j]
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int m0 = blockIdx.x * 32, n0 = blockIdx.y * 32;
  double acc[4] = {0.0, 0.0, 0.0, 0.0};
  for (int k0 = 0; k0 < k; k0 += 32) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int q = ty + 8 * r;
      if (TA) {
        const int gk = k0 + tx, gm = m0 + q;
        As[tx][q] = (gk < k && gm < k) ? A[gk + (size_t)gm * k] : 0.0;
      } else {
        const int gm = m0 + tx, gk = k0 + q;
        As[q][tx] = (gk < k && gm < k) ? A[gm + (size_t)gk * k] : 0.0;
      }
      const int gk = k0 + tx, gj = n0 + q;
      Bs[tx][q] = (gk < k && gj < k) ? B[gk + (size_t)gj * k] : 0.0;
    }
    __syncthreads();
#pragma unroll 8
    for (int kk = 0; kk < 32; ++kk) {
      const double a = As[kk][tx];
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[c] = fma(a, Bs[kk][ty + 8 * c], acc[c]);
    }
    __syncthreads();
  }
  const int gm = m0 + tx;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const int gj = n0 + ty + 8 * c;
    if (gm < k && gj < k) C[gm + (size_t)gj * k] = acc[c];
  }
}

void small_gemm(cudaStream_t s, bool ta, int k, const double* A, const double* B, double* C) {
  const dim3 grid((k + 31) / 32, (k + 31) / 32);
  if (ta) small_gemm_kernel<true><<<grid, 256, 0, s>>>(k, A, B, C);
  else small_gemm_kernel<false><<<grid, 256, 0, s>>>(k, A, B, C);
  CK_LAUNCH();
  ++g_kernel_launches;
}

// ---- 3. guard ------------------------------------------------------------------------------------------------------
// G <- 1.5 I - 0.5 G; flag[1] = max |G - I| (before)
__global__ void __launch_bounds__(1024) ns_prepare_kernel(int k, double* G, double* flagv) {
  __shared__ double red[32];
  double mx = 0.0;
  for (int idx = threadIdx.x; idx < k * k; idx += blockDim.x) {
    const int i = idx % k, j = idx / k;
    const double g = G[idx];
    const double dev = fabs(g - (i == j ? 1.0 : 0.0));
    mx = (dev == dev) ? fmax(mx, dev) : 1.0e300;
    G[idx] = (i == j ? 1.5 : 0.0) - 0.5 * g;
  }
  for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
  __syncthreads();
  if (threadIdx.x == 0) {
    double m = 0.0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) m = fmax(m, red[i]);
    flagv[1] = m;
  }
}

// theta_j = y_j^T (S y_j); residual check; accept flag.  One warp per column, single CTA loop.
__global__ void __launch_bounds__(1024) guard_kernel(int k, const double* __restrict__ Y, const double* __restrict__ SY,
                                                     const double* __restrict__ scal, double* __restrict__ w,
                                                     double* flagv, int* accept) {
  __shared__ double red[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  double worst = 0.0;
  for (int j = warp; j < k; j += nw) {
    double th = 0.0, nn = 0.0;
    for (int i = lane; i < k; i += 32) {
      const double y = Y[(size_t)j * k + i];
      th = fma(y, SY[(size_t)j * k + i], th);
      nn = fma(y, y, nn);
    }
    th = warp_sum(th);
    nn = warp_sum(nn);
    th /= nn;
    double r = 0.0;
    for (int i = lane; i < k; i += 32) {
      const double x = fabs(SY[(size_t)j * k + i] - th * Y[(size_t)j * k + i]);
      r = (x == x) ? fmax(r, x) : 1.0e300;
    }
    for (int o = 16; o > 0; o >>= 1) r = fmax(r, __shfl_xor_sync(0xffffffffu, r, o));
    if (!(th == th)) r = 1.0e300;
    worst = fmax(worst, r);
    if (lane == 0) w[j] = th;
  }
  if (lane == 0) red[warp] = worst;
  __syncthreads();
  if (threadIdx.x == 0) {
    double m = 0.0;
    for (int i = 0; i < nw; ++i) m = fmax(m, red[i]);
    flagv[2] = m;
    const double smax = scal[0];
    const bool ok = (flagv[1] <= 3.0e-8) && (m <= 64.0 * k * EPS * smax) && (smax <= 1.0e150);
    *accept = ok ? 1 : 0;
  }
}

// ---- 3b. r02: the guard in two launches instead of five (small GEMM Y^T Y, ns_prepare, small GEMM Y G, small GEMM S Y,
// guard_kernel: ~50 us of launch-bound work per Rayleigh-Ritz step).  One CTA takes GJ eigenvector columns j:
//   g = Yraw^T yraw_j (one warp per column of Yraw)          defect_j = max |g - e_j|
//   y_j = Yraw (1.5 e_j - 0.5 g)                             (the Newton-Schulz step, column j of it)
//   s = S y_j,  theta_j = y_j.s / y_j.y_j,  resid_j = max |s - theta_j y_j|
// and guard_accept_kernel takes the maxima and sets the accept flag.  Same arithmetic per entry as the GEMM form.
constexpr int GJ = 4;
__global__ void __launch_bounds__(256) guard_cols_kernel(int k, const double* __restrict__ Yraw,
                                                         const double* __restrict__ Sfull, double* __restrict__ Y,
                                                         double* __restrict__ w, double* __restrict__ defect,
                                                         double* __restrict__ resid) {
  extern __shared__ __align__(16) double sm[];
  double* yr = sm;               // GJ x k: yraw_j
  double* cf = yr + GJ * k;      // GJ x k: g, then the coefficients 1.5 e_j - 0.5 g
  double* ys = cf + GJ * k;      // GJ x k: y_j
  __shared__ double red[3][GJ][8];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int j0 = blockIdx.x * GJ;
  const int nj = min(GJ, k - j0);
  for (int e = tid; e < GJ * k; e += 256) {
    const int q = e / k, r = e - q * k;
    yr[e] = q < nj ? Yraw[(size_t)(j0 + q) * k + r] : 0.0;
  }
  __syncthreads();
  // g(i, q) = Yraw(:, i) . yraw_q
  for (int i = warp; i < k; i += 8) {
    double a[GJ];
#pragma unroll
    for (int q = 0; q < GJ; ++q) a[q] = 0.0;
    for (int r = lane; r < k; r += 32) {
      const double v = Yraw[(size_t)i * k + r];
#pragma unroll
      for (int q = 0; q < GJ; ++q) a[q] = fma(v, yr[q * k + r], a[q]);
    }
#pragma unroll
    for (int q = 0; q < GJ; ++q) {
      a[q] = warp_sum(a[q]);
      if (lane == 0) cf[q * k + i] = a[q];
    }
  }
  __syncthreads();
  // defect and Newton-Schulz coefficients
  double dmax[GJ];
#pragma unroll
  for (int q = 0; q < GJ; ++q) dmax[q] = 0.0;
  for (int e = tid; e < GJ * k; e += 256) {
    const int q = e / k, i = e - q * k;
    const double gv = cf[e], id = (i == j0 + q) ? 1.0 : 0.0;
    const double dev = fabs(gv - id);
    const double dv = (dev == dev) ? dev : 1.0e300;
#pragma unroll
    for (int qq = 0; qq < GJ; ++qq)
      if (qq == q && q < nj) dmax[qq] = fmax(dmax[qq], dv);
    cf[e] = 1.5 * id - 0.5 * gv;
  }
#pragma unroll
  for (int q = 0; q < GJ; ++q) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) dmax[q] = fmax(dmax[q], __shfl_xor_sync(0xffffffffu, dmax[q], o));
    if (lane == 0) red[0][q][warp] = dmax[q];
  }
  __syncthreads();
  if (tid < nj) {
    double m = 0.0;
    for (int x = 0; x < 8; ++x) m = fmax(m, red[0][tid][x]);
    defect[j0 + tid] = m;
  }
  // y_q(r) = sum_i Yraw(r, i) c(i, q): row r by `parts` threads, each a strided share of the columns (k <= 128
  // would leave half of the CTA idle and every thread with a chain of k dependent-latency loads)
  const int parts = k >= 256 ? 1 : 256 / k;
  double* pp = ys + GJ * k;  // parts x GJ x k partial sums (parts > 1)
  for (int e = tid; e < parts * k || (parts == 1 && e < k); e += 256) {
    const int part = e / k, r = e - part * k;
    double a[GJ];
#pragma unroll
    for (int q = 0; q < GJ; ++q) a[q] = 0.0;
#pragma unroll 8
    for (int i = part; i < k; i += parts) {
      const double v = Yraw[(size_t)i * k + r];
#pragma unroll
      for (int q = 0; q < GJ; ++q) a[q] = fma(v, cf[q * k + i], a[q]);
    }
#pragma unroll
    for (int q = 0; q < GJ; ++q) pp[(part * GJ + q) * k + r] = a[q];
  }
  __syncthreads();
  for (int e = tid; e < GJ * k; e += 256) {
    const int q = e / k, r = e - q * k;
    double a = 0.0;
    for (int part = 0; part < parts; ++part) a += pp[(part * GJ + q) * k + r];
    ys[e] = a;
    if (q < nj) Y[(size_t)(j0 + q) * k + r] = a;
  }
  __syncthreads();
  // s = S y (same split), Rayleigh quotient and residual
  for (int e = tid; e < parts * k || (parts == 1 && e < k); e += 256) {
    const int part = e / k, r = e - part * k;
    double a[GJ];
#pragma unroll
    for (int q = 0; q < GJ; ++q) a[q] = 0.0;
#pragma unroll 8
    for (int c = part; c < k; c += parts) {
      const double v = Sfull[(size_t)c * k + r];  // S symmetric: column c read along r
#pragma unroll
      for (int q = 0; q < GJ; ++q) a[q] = fma(v, ys[q * k + c], a[q]);
    }
#pragma unroll
    for (int q = 0; q < GJ; ++q) pp[(part * GJ + q) * k + r] = a[q];
  }
  __syncthreads();
  double yy[GJ], sy[GJ];
#pragma unroll
  for (int q = 0; q < GJ; ++q) yy[q] = sy[q] = 0.0;
  for (int e = tid; e < GJ * k; e += 256) {
    const int q = e / k, r = e - q * k;
    double a = 0.0;
    for (int part = 0; part < parts; ++part) a += pp[(part * GJ + q) * k + r];
    yr[e] = a;  // yraw is no longer needed: keep s there
    const double yv = ys[e];
#pragma unroll
    for (int qq = 0; qq < GJ; ++qq)
      if (qq == q) {
        yy[qq] = fma(yv, yv, yy[qq]);
        sy[qq] = fma(yv, a, sy[qq]);
      }
  }
#pragma unroll
  for (int q = 0; q < GJ; ++q) {
    yy[q] = warp_sum(yy[q]);
    sy[q] = warp_sum(sy[q]);
    if (lane == 0) { red[1][q][warp] = yy[q]; red[2][q][warp] = sy[q]; }
  }
  __syncthreads();
  double th[GJ];
#pragma unroll
  for (int q = 0; q < GJ; ++q) {
    double a = 0.0, b = 0.0;
    for (int x = 0; x < 8; ++x) { a += red[1][q][x]; b += red[2][q][x]; }
    th[q] = b / a;
  }
  double rmax[GJ];
#pragma unroll
  for (int q = 0; q < GJ; ++q) rmax[q] = 0.0;
  for (int r = tid; r < k; r += 256) {
#pragma unroll
    for (int q = 0; q < GJ; ++q) {
      const double x = fabs(yr[q * k + r] - th[q] * ys[q * k + r]);
      rmax[q] = (x == x) ? fmax(rmax[q], x) : 1.0e300;
    }
  }
  __syncthreads();  // red[0] is reused
#pragma unroll
  for (int q = 0; q < GJ; ++q) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) rmax[q] = fmax(rmax[q], __shfl_xor_sync(0xffffffffu, rmax[q], o));
    if (lane == 0) red[0][q][warp] = rmax[q];
  }
  __syncthreads();
  if (tid < nj) {
    double m = 0.0;
    for (int x = 0; x < 8; ++x) m = fmax(m, red[0][tid][x]);
    // (a NaN quotient makes every |s - theta y| NaN, which the loop above turned into 1e300)
    resid[j0 + tid] = m;
    double tq = 0.0;
#pragma unroll
    for (int q = 0; q < GJ; ++q) if (q == tid) tq = th[q];
    w[j0 + tid] = tq;
  }
}

__global__ void __launch_bounds__(256) guard_accept_kernel(int k, const double* __restrict__ defect,
                                                           const double* __restrict__ resid,
                                                           const double* __restrict__ scal, double* flagv,
                                                           int* accept) {
  __shared__ double red[2][8];
  double d = 0.0, r = 0.0;
  for (int j = threadIdx.x; j < k; j += 256) {
    d = fmax(d, defect[j]);
    r = fmax(r, resid[j]);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    d = fmax(d, __shfl_xor_sync(0xffffffffu, d, o));
    r = fmax(r, __shfl_xor_sync(0xffffffffu, r, o));
  }
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = d; red[1][threadIdx.x >> 5] = r; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double dm = 0.0, rm = 0.0;
    for (int i = 0; i < 8; ++i) { dm = fmax(dm, red[0][i]); rm = fmax(rm, red[1][i]); }
    flagv[1] = dm;
    flagv[2] = rm;
    const double smax = scal[0];
    const bool ok = (dm <= 3.0e-8) && (rm <= 64.0 * k * EPS * smax) && (smax <= 1.0e150);
    *accept = ok ? 1 : 0;
  }
}

int env_int(const char* name, int dflt) {
  const char* e = std::getenv(name);
  return e ? std::atoi(e) : dflt;
}

}  // namespace

size_t sym_eigh_scratch_doubles(int k) {
  return jacobi_scratch_doubles(k) + 6 * (size_t)k * k + 4 * (size_t)k + 64;
}

// device location of the guard's outputs: double[8] {max|S|, max|G-I|, max residual, ...} then the int accept flag
double* sym_eigh_flags(double* scratch, int k) {
  return scratch + jacobi_scratch_doubles(k) + 6 * (size_t)k * k + 4 * (size_t)k;
}

bool sym_eigh_uses_tridiag(int k) {
  // r02: with the register-resident tridiagonalisation the fast path wins from k = 16 on (k = 32: 0.093 ms against
  // 0.284 ms for the one-CTA Jacobi; profiles/r02_eigh_bench*)
  static const int min_k = env_int("DAV_EIGH_TRIDIAG_MIN_K", 16);
  return k >= min_k && k <= 512;
}

void sym_eigh(cudaStream_t s, int k, double* S, double* Y, double* w, double* scratch, int* status, int64_t lds) {
  if (k <= 0) return;
  if (lds <= 0) lds = k;
  if (!sym_eigh_uses_tridiag(k)) {
    if (lds != k) DAV_THROW(DAV_ERR_INVALID, "sym_eigh: a strided input needs the tridiagonal path (k >= 16)");
    jacobi_eigh(s, k, S, Y, w, scratch, status, nullptr);
    return;
  }
  const int max_smem = device_max_smem_optin();
  ensure_dyn_smem(tridiag_kernel<2>, max_smem - 2048);  // per (kernel, device)
  ensure_dyn_smem(tridiag_kernel<5>, max_smem - 2048);
  ensure_dyn_smem(tridiag_kernel<8>, max_smem - 2048);
  ensure_dyn_smem(tri_eigvec_kernel<2>, max_smem);
  ensure_dyn_smem(tri_eigvec_kernel<4>, max_smem);
  ensure_dyn_smem(tri_eigvec_kernel<8>, max_smem);
  ensure_dyn_smem(tri_eigvec_kernel<16>, max_smem);
  const size_t kk = (size_t)k * k;
  double* base = scratch + jacobi_scratch_doubles(k);
  double* work = base;            // tridiagonalisation workspace when S does not fit in shared memory
  double* Sfull = work + kk;      // symmetrised copy of the input
  double* Vh = Sfull + kk;        // reflectors
  double* G = Vh + kk;            // Y^T Y -> 1.5 I - 0.5 G
  double* Yraw = G + kk;          // eigenvectors before the Newton-Schulz step
  double* SY = Yraw + kk;         // S * Y
  double* tau = SY + kk;
  double* d = tau + k;
  double* e = d + k;
  double* lam = e + k;
  double* flagv = lam + k;        // [0] max|S|, [1] max|G - I|, [2] max residual
  int* accept = reinterpret_cast<int*>(flagv + 8);

  static const int reg_env = env_int("DAV_TRIDIAG_REG", 1);
  if (reg_env != 0 && k <= 128) {
    // register-resident matrix (r02): 16 x 16 threads, TS x TS tile each
    // (16 warps for k > 32: one warp issues a DFMA only every ~4 cycles, so the FP64 pipe of the SM needs >= 4
    // warps per scheduler; DAV_TRIDIAG_WARPS=8 selects the 8-warp tiling)
    static const int w8 = env_int("DAV_TRIDIAG_WARPS", 16) == 8;
    if (k <= 32) tridiag_reg_kernel<2, 2><<<1, 256, 0, s>>>(k, S, lds, Sfull, Vh, tau, d, e, flagv);
    else if (k <= 64 && w8) tridiag_reg_kernel<4, 4><<<1, 256, 0, s>>>(k, S, lds, Sfull, Vh, tau, d, e, flagv);
    else if (k <= 64) tridiag_reg_kernel<2, 4><<<1, 512, 0, s>>>(k, S, lds, Sfull, Vh, tau, d, e, flagv);
    else if (w8) tridiag_reg_kernel<8, 8><<<1, 256, 0, s>>>(k, S, lds, Sfull, Vh, tau, d, e, flagv);
    else tridiag_reg_kernel<4, 8><<<1, 512, 0, s>>>(k, S, lds, Sfull, Vh, tau, d, e, flagv);
  } else {
    // the step loop is instruction-issue bound on per-warp bookkeeping, not on the m^2 FMAs: few warps for small k
    static const int thr_env = env_int("DAV_TRIDIAG_THREADS", 0);
    // (measured: k = 64 -> 256 threads, k = 128/160 -> 512, k >= 256 -> 1024; the row mapping needs >= k/32 warps)
    int threads = thr_env > 0 ? thr_env : (k <= 64 ? 256 : (k <= 160 ? 512 : 1024));
    threads = std::min(1024, std::max(threads, 32 * ((k + 31) / 32)));
    const size_t small = 32 + 4 + 3 * (size_t)k + (size_t)threads;
    const size_t need_in = (small + kk) * sizeof(double);
    const int s_in = need_in <= (size_t)max_smem - 2048 ? 1 : 0;  // 2 KB of static tables
    const size_t tsm = s_in ? need_in : small * sizeof(double);
    if (k <= 64) tridiag_kernel<2><<<1, threads, tsm, s>>>(k, S, lds, work, s_in, Sfull, Vh, tau, d, e, flagv);
    else if (k <= 160) tridiag_kernel<5><<<1, threads, tsm, s>>>(k, S, lds, work, s_in, Sfull, Vh, tau, d, e, flagv);
    else tridiag_kernel<8><<<1, threads, tsm, s>>>(k, S, lds, work, s_in, Sfull, Vh, tau, d, e, flagv);
  }
  CK_LAUNCH();
  ++g_kernel_launches;
  {
    const int grid = (k + EW - 1) / EW;
    size_t sm = (3 + 5 * (size_t)EW) * k * sizeof(double);
    static const int eprof_on = env_int("DAV_EIGVEC_PROFILE", 0);
    double* eprof = eprof_on ? flagv + 3 : nullptr;  // (overwrites the tridiagonalisation's profile slots)
    static const int vhs_env = env_int("DAV_EIGVEC_VH_SMEM", 1);
    const int vh_smem = (vhs_env != 0 && sm + kk * sizeof(double) <= (size_t)max_smem) ? 1 : 0;
    if (vh_smem) sm += kk * sizeof(double);
    if (k <= 64) tri_eigvec_kernel<2><<<grid, EW * 32, sm, s>>>(k, d, e, Vh, tau, Yraw, lam, vh_smem, eprof);
    else if (k <= 128) tri_eigvec_kernel<4><<<grid, EW * 32, sm, s>>>(k, d, e, Vh, tau, Yraw, lam, vh_smem, eprof);
    else if (k <= 256) tri_eigvec_kernel<8><<<grid, EW * 32, sm, s>>>(k, d, e, Vh, tau, Yraw, lam, vh_smem, eprof);
    else tri_eigvec_kernel<16><<<grid, EW * 32, sm, s>>>(k, d, e, Vh, tau, Yraw, lam, vh_smem, eprof);
    CK_LAUNCH();
    ++g_kernel_launches;
  }
  static const int fused_guard = env_int("DAV_GUARD_FUSED", 1);
  if (fused_guard != 0) {
    // G (k^2 doubles) is free in this form: its head holds the per-column defects and residuals
    const size_t gsm = (3 + (size_t)(k >= 256 ? 1 : 256 / k)) * GJ * k * sizeof(double);
    if (gsm > 40 * 1024) ensure_dyn_smem(guard_cols_kernel, 64 * 1024);
    guard_cols_kernel<<<(k + GJ - 1) / GJ, 256, gsm, s>>>(k, Yraw, Sfull, Y, w, G, G + k);
    CK_LAUNCH();
    ++g_kernel_launches;
    guard_accept_kernel<<<1, 256, 0, s>>>(k, G, G + k, flagv, flagv, accept);
    CK_LAUNCH();
    ++g_kernel_launches;
  } else {
    small_gemm(s, true, k, Yraw, Yraw, G);
    ns_prepare_kernel<<<1, 1024, 0, s>>>(k, G, flagv);
    CK_LAUNCH();
    ++g_kernel_launches;
    small_gemm(s, false, k, Yraw, G, Y);
    small_gemm(s, false, k, Sfull, Y, SY);
    guard_kernel<<<1, 1024, 0, s>>>(k, Y, SY, flagv, w, flagv, accept);
    CK_LAUNCH();
    ++g_kernel_launches;
  }
  // Jacobi runs only when the guard rejected the fast path (device-side decision)
  // (from the symmetrised dense copy: the caller's matrix may be strided)
  jacobi_eigh(s, k, Sfull, Y, w, scratch, status, accept);
}

}  // namespace dav
