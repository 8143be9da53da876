// Small dense (k x k) problems, each solved by ONE CTA: replaces lapack_wrapper.f90's DSYEV /
// DSYGV call sites of the Rayleigh-Ritz step (lapack_wrapper.f90:14-91; davidson.f90:153,155,394)
// and supplies the Gram-matrix transforms of the block orthonormalisation that replaces
// DGEQRF+DORGQR (lapack_wrapper.f90:176-236; davidson.f90:213,434).
//
// jacobi_eigh: two-sided Jacobi with the round-robin ("chess tournament") parallel ordering:
// each round rotates k/2 disjoint (p,q) pairs; a thread owns one 2x2 block S[{p1,q1},{p2,q2}] and
// applies J_P^T . J_Q to it, so a round needs two block barriers and no atomics.  S lives in
// shared memory whenever it fits (k <= ~165), the eigenvector accumulator too when both fit
// (k <= ~117); otherwise they stay in L2-resident global scratch.
#include <algorithm>

#include "kernels.cuh"

namespace dav {
namespace {

constexpr double EPS = 2.220446049250313e-16;
constexpr int JT = 1024;      // threads of the Jacobi CTA
constexpr int MAX_SWEEPS = 40;

__device__ __forceinline__ double block_reduce_sum(double v, double* red) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) red[wid] = v;
  __syncthreads();
  if (wid == 0) {
    double x = (lane < (blockDim.x + 31) / 32) ? red[lane] : 0.0;
    x = warp_sum(x);
    if (lane == 0) red[0] = x;
  }
  __syncthreads();
  const double r = red[0];
  __syncthreads();
  return r;
}

__device__ __forceinline__ double block_reduce_max(double v, double* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) red[wid] = v;
  __syncthreads();
  if (wid == 0) {
    double x = (lane < (blockDim.x + 31) / 32) ? red[lane] : -1.0e300;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x = fmax(x, __shfl_xor_sync(0xffffffffu, x, o));
    if (lane == 0) red[0] = x;
  }
  __syncthreads();
  const double r = red[0];
  __syncthreads();
  return r;
}

__global__ void __launch_bounds__(JT) jacobi_kernel(int k, const double* __restrict__ S_in, double* __restrict__ Y,
                                                    double* __restrict__ w, double* gS, double* gV, int s_in_smem,
                                                    int v_in_smem, int* status, const int* skip) {
  extern __shared__ __align__(16) double sm[];
  if (skip && *skip) return;  // the tridiagonal fast path (trideig.cu) was accepted
  const int kp = k + (k & 1), np = kp / 2;
  const int tid = threadIdx.x, nt = blockDim.x;
  // carve shared memory: small arrays first, then the big ones
  double* red = sm;                      // 32
  double* rc = red + 32;                 // np
  double* rs = rc + np;                  // np
  double* wv = rs + np;                  // kp
  int* rp = reinterpret_cast<int*>(wv + kp);  // np
  int* rq = rp + np;                          // np
  int* rk = rq + np;                          // kp
  size_t off = 32 + 2 * (size_t)np + kp + ((2 * (size_t)np + kp) * sizeof(int) + 7) / 8;
  off = (off + 1) & ~(size_t)1;
  double* S = gS;
  double* V = gV;
  if (s_in_smem) { S = sm + off; off += (size_t)kp * kp; }
  if (v_in_smem) { V = sm + off; }

  // 1. symmetrise from the upper triangle (DSYEV 'U'), V = I
  double part = 0.0;
  for (int e = tid; e < kp * kp; e += nt) {
    const int i = e % kp, j = e / kp;
    double v = 0.0;
    if (i < k && j < k) v = (i <= j) ? S_in[i + (size_t)j * k] : S_in[j + (size_t)i * k];
    S[e] = v;
    V[e] = (i == j) ? 1.0 : 0.0;
    part += v * v;
  }
  const double normF = sqrt(block_reduce_sum(part, red));
  if (!(normF <= 1.0e300)) {  // NaN or Inf in the input
    if (tid == 0) atomicOr(status, 1);
    return;
  }
  const double abs_thr = EPS * normF / (16.0 * kp);

  // 2. sweeps
  const int m = kp - 1;
  for (int sweep = 0; sweep < MAX_SWEEPS; ++sweep) {
    int rotated = 0;
    for (int r = 0; r < m; ++r) {
      if (tid < np) {
        int a, b;
        if (tid == 0) { a = kp - 1; b = r; }
        else { a = (r + tid) % m; b = (r - tid + m) % m; }
        const int p = min(a, b), q = max(a, b);
        double c = 1.0, s = 0.0;
        if (q < k) {
          const double apq = S[p + (size_t)q * kp], app = S[p + (size_t)p * kp], aqq = S[q + (size_t)q * kp];
          const double aa = fabs(apq);
          if (aa > abs_thr && aa > EPS * sqrt(fabs(app) * fabs(aqq))) {
            const double tau = (aqq - app) / (2.0 * apq);
            const double t = copysign(1.0, tau) / (fabs(tau) + sqrt(1.0 + tau * tau));
            c = 1.0 / sqrt(1.0 + t * t);
            s = t * c;
            rotated = 1;
          }
        }
        rc[tid] = c; rs[tid] = s; rp[tid] = p; rq[tid] = q;
      }
      __syncthreads();
      // S <- J^T S J, one 2x2 block per work item (upper block-triangle, mirrored)
      for (int e = tid; e < np * np; e += nt) {
        const int iP = e % np, iQ = e / np;
        if (iP > iQ) continue;
        const double cP = rc[iP], sP = rs[iP], cQ = rc[iQ], sQ = rs[iQ];
        if (sP == 0.0 && sQ == 0.0) continue;
        const int p1 = rp[iP], q1 = rq[iP], p2 = rp[iQ], q2 = rq[iQ];
        if (iP == iQ) {
          const double apq = S[p1 + (size_t)q1 * kp], app = S[p1 + (size_t)p1 * kp], aqq = S[q1 + (size_t)q1 * kp];
          const double t = sP / cP;
          S[p1 + (size_t)p1 * kp] = app - t * apq;
          S[q1 + (size_t)q1 * kp] = aqq + t * apq;
          S[p1 + (size_t)q1 * kp] = 0.0;
          S[q1 + (size_t)p1 * kp] = 0.0;
        } else {
          const double m00 = S[p1 + (size_t)p2 * kp], m01 = S[p1 + (size_t)q2 * kp];
          const double m10 = S[q1 + (size_t)p2 * kp], m11 = S[q1 + (size_t)q2 * kp];
          const double r00 = cP * m00 - sP * m10, r01 = cP * m01 - sP * m11;
          const double r10 = sP * m00 + cP * m10, r11 = sP * m01 + cP * m11;
          const double n00 = cQ * r00 - sQ * r01, n01 = sQ * r00 + cQ * r01;
          const double n10 = cQ * r10 - sQ * r11, n11 = sQ * r10 + cQ * r11;
          S[p1 + (size_t)p2 * kp] = n00; S[p2 + (size_t)p1 * kp] = n00;
          S[p1 + (size_t)q2 * kp] = n01; S[q2 + (size_t)p1 * kp] = n01;
          S[q1 + (size_t)p2 * kp] = n10; S[p2 + (size_t)q1 * kp] = n10;
          S[q1 + (size_t)q2 * kp] = n11; S[q2 + (size_t)q1 * kp] = n11;
        }
      }
      // V <- V J
      for (int e = tid; e < kp * np; e += nt) {
        const int row = e % kp, iP = e / kp;
        const double s = rs[iP];
        if (s == 0.0) continue;
        const double c = rc[iP];
        const int p = rp[iP], q = rq[iP];
        const double vp = V[row + (size_t)p * kp], vq = V[row + (size_t)q * kp];
        V[row + (size_t)p * kp] = c * vp - s * vq;
        V[row + (size_t)q * kp] = s * vp + c * vq;
      }
      __syncthreads();
    }
    const int any = __syncthreads_or(rotated);
    if (!any) break;
    if (sweep == MAX_SWEEPS - 1 && tid == 0) atomicOr(status, 1);
  }

  // 3. ascending order (stable), scatter eigenvectors
  for (int i = tid; i < k; i += nt) wv[i] = S[i + (size_t)i * kp];
  __syncthreads();
  for (int i = tid; i < k; i += nt) {
    const double wi = wv[i];
    int rank = 0;
    for (int j = 0; j < k; ++j) {
      const double wj = wv[j];
      rank += (wj < wi) || (wj == wi && j < i);
    }
    if (!(wi == wi)) { atomicOr(status, 1); rank = i; }
    rk[i] = rank;
    w[rank] = wi;
  }
  __syncthreads();
  for (int e = tid; e < k * k; e += nt) {
    const int row = e % k, col = e / k;
    Y[row + (size_t)rk[col] * k] = V[row + (size_t)col * kp];
  }
}


// ---------------------------------------------------------------------------------------------------
// Fast path (k <= ~165): the sweeps run on S alone (shared memory, one CTA, 2 barriers per round) and
// every rotation (c, s) is logged; a second kernel replays the log on the rows of V = I, one row
// per thread group, rows spread over several CTAs.  Same arithmetic as accumulating V inside the
// sweeps, but V no longer sits on the critical path of the (latency-bound) rounds.
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void rr_pair(int kp, int r, int i, int& p, int& q) {
  const int m = kp - 1;
  int a, b;
  if (i == 0) { a = kp - 1; b = r; }
  else { a = (r + i) % m; b = (r - i + m) % m; }
  p = min(a, b);
  q = max(a, b);
}

__global__ void __launch_bounds__(512) jacobi_sweeps_kernel(int k, const double* __restrict__ S_in, double* __restrict__ w,
                                                            int* __restrict__ rank_out, double2* __restrict__ rotlog,
                                                            int* __restrict__ nrounds_out, int max_rounds,
                                                            int* status, const int* skip) {
  if (skip && *skip) return;  // the tridiagonal fast path (trideig.cu) was accepted
  // Only the upper triangle of S is kept up to date: element (a, b) lives at [min(a,b) + max(a,b) * ld].  The
  // round is bound by shared-memory instruction issue, so the per-pair parameters are packed ((c,s) as one
  // double2, (p,q) as one int2, the work item (iP,iQ) as one 32-bit word) and nothing is mirrored.
  extern __shared__ __align__(16) double sm[];
  const int kp = k + (k & 1), np = kp / 2, ld = kp + 1;
  const int tid = threadIdx.x, nt = blockDim.x;
  double* red = sm;                                         // 32
  double2* rcs = reinterpret_cast<double2*>(red + 32);      // np  (c, s)
  double* wv = reinterpret_cast<double*>(rcs + np);         // kp
  int2* rpq = reinterpret_cast<int2*>(wv + kp);             // np  (p, q)
  const int nitems = np * (np + 1) / 2;
  unsigned* itab = reinterpret_cast<unsigned*>(rpq + np);   // nitems (iP | iQ << 16)
  size_t off = 32 + 2 * (size_t)np + kp + (size_t)np + ((size_t)nitems * 4 + 7) / 8;
  off = (off + 1) & ~(size_t)1;
  double* S = sm + off;

  for (int e = tid; e < np * np; e += nt) {
    const int iP = e % np, iQ = e / np;
    if (iP <= iQ) itab[iQ * (iQ + 1) / 2 + iP] = (unsigned)iP | ((unsigned)iQ << 16);
  }
  double mx = 0.0;
  for (int e = tid; e < kp * kp; e += nt) {
    const int i = e % kp, j = e / kp;
    double v = 0.0;
    if (i < k && j < k) v = (i <= j) ? S_in[i + (size_t)j * k] : S_in[j + (size_t)i * k];
    S[i + j * ld] = v;
    const double av = fabs(v);
    mx = (av == av) ? fmax(mx, av) : 1.0e308 * 10.0;
  }
  mx = block_reduce_max(mx, red);
  if (!(mx <= 1.0e300)) {
    if (tid == 0) { atomicOr(status, 1); *nrounds_out = 0; }
    for (int i = tid; i < k; i += nt) { rank_out[i] = i; w[i] = mx; }
    return;
  }
  // exact power-of-two scaling so that squares cannot overflow / underflow
  const int ex = (mx > 0.0) ? ilogb(mx) : 0;
  const double sc = scalbn(1.0, -ex), usc = scalbn(1.0, ex);
  double part = 0.0;
  for (int e = tid; e < kp * kp; e += nt) {
    const int i = e % kp, j = e / kp;
    const double v = S[i + j * ld] * sc;
    S[i + j * ld] = v;
    part += v * v;
  }
  const double normF = sqrt(block_reduce_sum(part, red));
  const double abs_thr = EPS * normF / (16.0 * kp);

  const int m = kp - 1;
  // this thread's pair of the current round, advanced incrementally (a, b -> a+1, b+1 mod m)
  int pa = (tid == 0) ? kp - 1 : tid % m, pb = (tid == 0) ? 0 : (m - (tid % m)) % m;
  int round = 0;
  for (int sweep = 0; sweep < MAX_SWEEPS; ++sweep) {
    int rotated = 0;
    for (int r = 0; r < m; ++r, ++round) {
      if (tid < np) {
        // pairs are kept UNSORTED: pa walks up and pb walks down with the pair slot, so a warp's 32 work items
        // touch 32 consecutive rows -- with ld = 1 (mod 16) element (r, c) sits in bank (r + c) mod 16 whichever
        // triangle it is read from, i.e. the accesses below are shared-memory bank-conflict free
        const int p = pa, q = pb;
        double c = 1.0, s = 0.0;
        if (p < k && q < k) {
          const double apq = S[min(p, q) + max(p, q) * ld], app = S[p + p * ld], aqq = S[q + q * ld];
          const double aa = fabs(apq);
          if (aa > abs_thr && aa * aa > (EPS * EPS) * fabs(app * aqq)) {
            // t = sign(tau) / (|tau| + sqrt(1 + tau^2)), tau = (aqq - app) / (2 apq), without forming tau;
            // rsqrt / reciprocal instead of sqrt / divide (the 2x2 block is transformed with the general formulas
            // below, so the similarity stays orthogonal to round-off whatever the last bits of t are)
            const double a = aqq - app, b = 2.0 * apq;
            const double h2 = a * a + b * b;
            const double rr = h2 * rsqrt(h2);
            const double t = b * __drcp_rn(a + copysign(rr, a));
            c = rsqrt(1.0 + t * t);
            s = t * c;
            rotated = 1;
          }
        }
        rcs[tid] = make_double2(c, s);
        rpq[tid] = make_int2(p, q);
        if (round < max_rounds) rotlog[(size_t)round * np + tid] = make_double2(c, s);
        // next round's pair
        if (tid == 0) { pb = (pb + 1 == m) ? 0 : pb + 1; }
        else { pa = (pa + 1 == m) ? 0 : pa + 1; pb = (pb + 1 == m) ? 0 : pb + 1; }
      }
      __syncthreads();
      for (int e = tid; e < nitems; e += nt) {
        const unsigned it = itab[e];
        const int iP = it & 0xffff, iQ = it >> 16;
        const double2 P = rcs[iP], Q = rcs[iQ];
        if (P.y == 0.0 && Q.y == 0.0) continue;
        const int2 pp = rpq[iP], qq = rpq[iQ];
        const int p1 = pp.x, q1 = pp.y, p2 = qq.x, q2 = qq.y;
        if (iP == iQ) {
          const int apq_i = min(p1, q1) + max(p1, q1) * ld;
          const double apq = S[apq_i], app = S[p1 + p1 * ld], aqq = S[q1 + q1 * ld];
          const double cc = P.x * P.x, ss = P.y * P.y, cs2 = 2.0 * P.x * P.y;
          S[p1 + p1 * ld] = cc * app - cs2 * apq + ss * aqq;
          S[q1 + q1 * ld] = ss * app + cs2 * apq + cc * aqq;
          S[apq_i] = P.x * P.y * (app - aqq) + (cc - ss) * apq;  // ~ eps |apq|: annihilated to round-off
        } else {
          // upper-triangle addresses of the four elements of the 2x2 block
          const int a00 = min(p1, p2) + max(p1, p2) * ld, a01 = min(p1, q2) + max(p1, q2) * ld;
          const int a10 = min(q1, p2) + max(q1, p2) * ld, a11 = min(q1, q2) + max(q1, q2) * ld;
          const double m00 = S[a00], m01 = S[a01], m10 = S[a10], m11 = S[a11];
          const double r00 = P.x * m00 - P.y * m10, r01 = P.x * m01 - P.y * m11;
          const double r10 = P.y * m00 + P.x * m10, r11 = P.y * m01 + P.x * m11;
          S[a00] = Q.x * r00 - Q.y * r01; S[a01] = Q.y * r00 + Q.x * r01;
          S[a10] = Q.x * r10 - Q.y * r11; S[a11] = Q.y * r10 + Q.x * r11;
        }
      }
      __syncthreads();
    }
    const int any = __syncthreads_or(rotated);
    if (!any) break;
    if (sweep == MAX_SWEEPS - 1 && tid == 0) atomicOr(status, 1);
  }
  if (round > max_rounds && tid == 0) atomicOr(status, 1);

  for (int i = tid; i < k; i += nt) wv[i] = S[i + i * ld];
  __syncthreads();
  for (int i = tid; i < k; i += nt) {
    const double wi = wv[i];
    int rank = 0;
    for (int j = 0; j < k; ++j) {
      const double wj = wv[j];
      rank += (wj < wi) || (wj == wi && j < i);
    }
    if (!(wi == wi)) { atomicOr(status, 1); rank = i; }
    rank_out[i] = rank;
    w[rank] = wi * usc;
  }
  if (tid == 0) *nrounds_out = min(round, max_rounds);
}

// Replays the rotation log on the rows of V = I.  A warp owns 2 rows x 16 lanes per row: the rotations of one
// round touch disjoint column pairs, so the 16 lanes of a row work independently and only a __syncwarp
// separates rounds.  The log is staged through shared memory in chunks of VCHUNK rounds (coalesced loads, one
// block barrier per chunk); pair indices advance incrementally (a, b -> a+1, b+1 mod m).  CTA = 8 warps = 16 rows.
constexpr int VROWS = 16, VPARTS = 16, VCHUNK = 16, VMAXI = 8;  // np <= VPARTS * VMAXI = 128 pairs (k <= 256)
__global__ void __launch_bounds__(VROWS * VPARTS) jacobi_vectors_kernel(int k, const double2* __restrict__ rotlog,
                                                                      const int* __restrict__ nrounds_in,
                                                                      const int* __restrict__ rank,
                                                                      double* __restrict__ Y, const int* skip) {
  extern __shared__ __align__(16) double sm[];
  if (skip && *skip) return;
  const int kp = k + (k & 1), np = kp / 2, ld = kp + 1, m = kp - 1;
  const int lrow = threadIdx.x / VPARTS, part = threadIdx.x % VPARTS;  // lanes 0-15: one row, 16-31: the next
  const int row = blockIdx.x * VROWS + lrow;
  double2* chunk = reinterpret_cast<double2*>(sm);                     // VCHUNK * np
  double* vr = sm + 2 * (size_t)VCHUNK * np + (size_t)lrow * ld;
  for (int j = part; j < kp; j += VPARTS) vr[j] = (j == row) ? 1.0 : 0.0;
  // pairs of round 0 for this lane's pair slots i = part, part + 16, ...
  int pa[VMAXI], pb[VMAXI];
#pragma unroll
  for (int t = 0; t < VMAXI; ++t) {
    const int i = part + t * VPARTS;
    if (i == 0) { pa[t] = kp - 1; pb[t] = 0; }
    else { pa[t] = i % m; pb[t] = (m - (i % m)) % m; }
  }
  const int nrounds = *nrounds_in;
  for (int r0 = 0; r0 < nrounds; r0 += VCHUNK) {
    const int nr = min(VCHUNK, nrounds - r0);
    __syncthreads();  // previous chunk fully consumed
    for (int e = threadIdx.x; e < nr * np; e += blockDim.x) chunk[e] = rotlog[(size_t)r0 * np + e];
    __syncthreads();
    for (int rr = 0; rr < nr; ++rr) {
      const double2* rl = chunk + rr * np;
      // gather the operands of all of this lane's rotations first (disjoint column pairs), then rotate
      double2 cs[VMAXI];
      double vp[VMAXI], vq[VMAXI];
#pragma unroll
      for (int t = 0; t < VMAXI; ++t) {
        const int i = part + t * VPARTS;
        if (i < np) {
          cs[t] = rl[i];
          vp[t] = vr[pa[t]];
          vq[t] = vr[pb[t]];
        }
      }
#pragma unroll
      for (int t = 0; t < VMAXI; ++t) {
        const int i = part + t * VPARTS;
        if (i < np) {
          if (cs[t].y != 0.0) {
            vr[pa[t]] = cs[t].x * vp[t] - cs[t].y * vq[t];
            vr[pb[t]] = cs[t].y * vp[t] + cs[t].x * vq[t];
          }
          if (i == 0) { pb[t] = (pb[t] + 1 == m) ? 0 : pb[t] + 1; }
          else { pa[t] = (pa[t] + 1 == m) ? 0 : pa[t] + 1; pb[t] = (pb[t] + 1 == m) ? 0 : pb[t] + 1; }
        }
      }
      __syncwarp();
    }
  }
  __syncwarp();
  if (row < k)
    for (int j = part; j < k; j += VPARTS) Y[row + (size_t)rank[j] * k] = vr[j];
}

__global__ void scale_cols_rsqrt_checked_kernel(int k, const double* __restrict__ U, const double* __restrict__ sv,
                                                double* __restrict__ T, int* status) {
  for (int e = threadIdx.x; e < k * k; e += blockDim.x) {
    const int j = e / k;
    const double s = sv[j];
    if (!(s > 0.0)) {
      if (e % k == 0) atomicOr(status, 2);
      T[e] = 0.0;
    } else {
      T[e] = U[e] / sqrt(s);
    }
  }
}

__global__ void gram_prescale_kernel(int k, double* G, double* D) {
  for (int i = threadIdx.x; i < k; i += blockDim.x) {
    const double d = G[i + (size_t)i * k];
    D[i] = (d > 0.0) ? 1.0 / sqrt(d) : 0.0;
  }
  __syncthreads();
  for (int e = threadIdx.x; e < k * k; e += blockDim.x) G[e] *= D[e % k] * D[e / k];
}

__global__ void svqb_make_T_kernel(int k, const double* __restrict__ U, const double* __restrict__ sv,
                                   const double* __restrict__ D, double* __restrict__ T, int* flags) {
  __shared__ double red[32];
  double mx = -1.0e300;
  for (int j = threadIdx.x; j < k; j += blockDim.x) mx = fmax(mx, sv[j]);
  const double smax = block_reduce_max(mx, red);
  const double thr = smax * (double)k * EPS * 16.0;
  for (int e = threadIdx.x; e < k * k; e += blockDim.x) {
    const int i = e % k, j = e / k;
    const double s = sv[j];
    const bool good = (smax > 0.0) && (s > thr);
    T[e] = good ? D[i] * U[e] / sqrt(s) : 0.0;
    if (i == 0) flags[j] = good ? 0 : 1;
  }
}

__global__ void symmetrize_from_upper_kernel(int k, double* S, int64_t ld) {
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < k * k; e += gridDim.x * blockDim.x) {
    const int i = e % k, j = e / k;
    if (i > j) S[i + j * ld] = S[j + i * ld];
  }
}

__global__ void max_abs_kernel(int rows, int cols, const double* __restrict__ G, int64_t ld, int minus_identity,
                               double* out) {
  __shared__ double red[32];
  double mx = 0.0;
  for (int e = threadIdx.x; e < rows * cols; e += blockDim.x) {
    const int i = e % rows, j = e / rows;
    const double v = fabs(G[i + j * ld] - ((minus_identity && i == j) ? 1.0 : 0.0));
    mx = (v == v) ? fmax(mx, v) : 1.0e300;
  }
  const double r = block_reduce_max(mx, red);
  if (threadIdx.x == 0) out[0] = r;
}

__global__ void __launch_bounds__(1024) cholesky_upper_kernel(int k, double* G, int64_t ld, int* status) {
  __shared__ double piv;
  __shared__ int bad;
  if (threadIdx.x == 0) bad = 0;
  __syncthreads();
  for (int j = 0; j < k; ++j) {
    if (threadIdx.x == 0) {
      const double d = G[j + (size_t)j * ld];
      if (!(d > 0.0)) { bad = 1; piv = 1.0; }
      else piv = sqrt(d);
      G[j + (size_t)j * ld] = piv;
    }
    __syncthreads();
    if (bad) break;
    const double r = piv;
    for (int i = j + 1 + threadIdx.x; i < k; i += blockDim.x) G[j + (size_t)i * ld] /= r;
    __syncthreads();
    const int rem = k - j - 1;
    for (int e = threadIdx.x; e < rem * rem; e += blockDim.x) {
      const int a = j + 1 + e % rem, b = j + 1 + e / rem;
      if (a <= b) G[a + (size_t)b * ld] -= G[j + (size_t)a * ld] * G[j + (size_t)b * ld];
    }
    __syncthreads();
  }
  if (bad && threadIdx.x == 0) atomicOr(status, 2);
  // zero the strict lower triangle so the result is a clean R
  for (int e = threadIdx.x; e < k * k; e += blockDim.x) {
    const int i = e % k, j = e / k;
    if (i > j) G[i + (size_t)j * ld] = 0.0;
  }
}


// Upper Cholesky factor G = R^T R and its inverse, one CTA, both triangles packed in shared memory
// ((i,j), i <= j, at i + j(j+1)/2).  T <- R^-1 (dense b x b, zero below the diagonal).  flag[0] = 1.0 when a
// pivot is not safely positive (the caller then falls back to the SVQB transform), else 0.0.  G is only read.
__global__ void __launch_bounds__(512) chol_inv_upper_kernel(int b, const double* __restrict__ G, double* __restrict__ T,
                                                             double* __restrict__ flag, double* gwork) {
  extern __shared__ __align__(16) double sm[];
  const int tri = b * (b + 1) / 2;
  // gwork != nullptr: the triangles do not fit in shared memory (b >= ~170) and live in global scratch
  double* R = gwork ? gwork : sm;
  double* X = R + tri;
  double* d0 = gwork ? sm : X + tri;
  __shared__ int bad;
  const int tid = threadIdx.x, nt = blockDim.x;
  if (tid == 0) bad = 0;
  for (int e = tid; e < b * b; e += nt) {
    const int i = e % b, j = e / b;
    if (i <= j) R[i + j * (j + 1) / 2] = G[i + (size_t)j * b];
    if (i == j) d0[i] = G[i + (size_t)j * b];
  }
  __syncthreads();
  for (int j = 0; j < b; ++j) {
    const double d = R[j + j * (j + 1) / 2];
    // pivot must stay well above the round-off of the eliminations (cond(G) < ~1e12)
    if (!(d > 1e-12 * fabs(d0[j])) || !(d0[j] > 0.0)) {
      if (tid == 0) bad = 1;
      break;  // uniform: every thread reads the same d
    }
    const double rinv = rsqrt(d);
    __syncthreads();  // everyone has read the pivot
    for (int i = j + tid; i < b; i += nt) {
      const int idx = j + i * (i + 1) / 2;
      R[idx] = (i == j) ? d * rinv : R[idx] * rinv;
    }
    __syncthreads();
    const int rem = b - j - 1;
    for (int e = tid; e < rem * rem; e += nt) {
      const int a = j + 1 + e % rem, c = j + 1 + e / rem;
      if (a <= c) R[a + c * (c + 1) / 2] -= R[j + a * (a + 1) / 2] * R[j + c * (c + 1) / 2];
    }
    __syncthreads();
  }
  __syncthreads();
  if (bad) {
    if (tid == 0) flag[0] = 1.0;
    return;
  }
  // X = R^-1, column j by thread j (back substitution)
  for (int j = tid; j < b; j += nt) {
    const int cj = j * (j + 1) / 2;
    X[j + cj] = 1.0 / R[j + cj];
    for (int i = j - 1; i >= 0; --i) {
      double acc = 0.0;
      for (int l = i + 1; l <= j; ++l) acc = fma(R[i + l * (l + 1) / 2], X[l + cj], acc);
      X[i + cj] = -acc / R[i + i * (i + 1) / 2];
    }
  }
  __syncthreads();
  for (int e = tid; e < b * b; e += nt) {
    const int i = e % b, j = e / b;
    T[e] = (i <= j) ? X[i + j * (j + 1) / 2] : 0.0;
  }
  if (tid == 0) flag[0] = 0.0;
}

// Register-tile variant for b <= 128 (the widths of every BASELINE config): T x T threads, T = ceil(b/4), thread
// (ty, tx) keeps the 4 x 4 tile (rows 4ty.., columns 4tx..) in registers.  Phase 1, right-looking Cholesky: per step
// ONE barrier -- the owners of row j+1 publish their freshly updated (unscaled) row and the diagonal thread its
// 1/sqrt(pivot) while everyone else is still applying step j (look-ahead), the update uses row_a * row_c / pivot.
// Phase 2, R^-1 by the column operations that turn R into I applied to an identity kept in the same registers
// (X(:,j) /= R(j,j); X(:,c) -= X(:,j) R(j,c)), same look-ahead.  ~2b barrier-separated steps of 16 FMAs per thread
// instead of b steps of index arithmetic plus a thread-per-column back substitution: 110 -> ~10 us at b = 64.
// body shared by chol_inv_tile_kernel and the fused pip_small_kernel: `active` threads (tid < T*T) hold the tile t of
// the symmetric matrix (both triangles filled, identity padding beyond b) and d0 = its original diagonal entries;
// on return t holds R^-1 (valid for row <= column) unless *bad.  Every thread of the CTA must call it (barriers).
__device__ __forceinline__ void chol_inv_tile_body(int T, bool active, int tx, int ty, double (&t)[4][4],
                                                   const double (&d0)[4], double* Rs, double* buf, double* rinv_s,
                                                   int* bad) {
  const int bp = 4 * T;
  const int r0 = 4 * ty, c0 = 4 * tx;
  // phase 1 look-ahead: row jn (unscaled) and 1/sqrt(pivot jn)
  auto publish_row = [&](int jn) {
    if (!active || ty != (jn >> 2)) return;
    const int i = jn & 3;
    double* rb = buf + (jn & 1) * bp;
#pragma unroll
    for (int ii = 0; ii < 4; ++ii)
      if (ii == i) {
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) rb[c0 + jj] = t[ii][jj];
        if (tx == ty) {
          const double d = t[ii][ii];
          // pivot must stay well above the round-off of the eliminations (cond(G) < ~1e12)
          if (!(d > 1e-12 * fabs(d0[ii])) || !(d0[ii] > 0.0)) {
            *bad = 1;
            rinv_s[jn] = 0.0;
          } else {
            rinv_s[jn] = rsqrt(d);
          }
        }
      }
  };
  publish_row(0);
  for (int j = 0; j < bp; ++j) {
    __syncthreads();
    if (*bad) break;  // uniform
    const double* rb = buf + (j & 1) * bp;
    const double rinv = rinv_s[j];
    if (active && ty == (j >> 2)) {
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) Rs[(size_t)j * bp + c0 + jj] = rb[c0 + jj] * rinv;
    }
    if (active && r0 + 3 > j && c0 + 3 > j) {
      const double dinv = rinv * rinv;
      double ra[4], rc[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        ra[i] = (r0 + i > j) ? rb[r0 + i] * dinv : 0.0;
        rc[i] = (c0 + i > j) ? rb[c0 + i] : 0.0;
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) t[i][jj] = fma(-ra[i], rc[jj], t[i][jj]);
    }
    if (j + 1 < bp) publish_row(j + 1);
  }
  __syncthreads();
  if (*bad) return;

  // phase 2: X = R^-1 in the same registers
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) t[i][jj] = (r0 + i == c0 + jj) ? 1.0 : 0.0;
  auto publish_col = [&](int jn) {  // column jn scaled by 1 / R(jn, jn)
    if (!active || tx != (jn >> 2)) return;
    const int jj = jn & 3;
    const double rinv = rinv_s[jn];
    double* xb = buf + (jn & 1) * bp;
#pragma unroll
    for (int q = 0; q < 4; ++q)
      if (q == jj) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          t[i][q] *= rinv;
          xb[r0 + i] = t[i][q];
        }
      }
  };
  publish_col(0);
  for (int j = 0; j < bp; ++j) {
    __syncthreads();
    if (active && c0 + 3 > j && r0 <= j) {  // X(a, j) = 0 for a > j
      const double* xb = buf + (j & 1) * bp;
      const double* rj = Rs + (size_t)j * bp;
      double xa[4], rc[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        xa[i] = xb[r0 + i];
        rc[i] = (c0 + i > j) ? rj[c0 + i] : 0.0;
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) t[i][jj] = fma(-xa[i], rc[jj], t[i][jj]);
    }
    if (j + 1 < bp) publish_col(j + 1);
  }
}

__global__ void __launch_bounds__(1024) chol_inv_tile_kernel(int b, const double* __restrict__ G,
                                                             double* __restrict__ Tout, double* __restrict__ flag) {
  extern __shared__ __align__(16) double sm[];
  const int T = (b + 3) >> 2, bp = 4 * T;
  double* Rs = sm;                       // bp x bp: Rs[j * bp + c] = R(j, c), c >= j
  double* buf = Rs + (size_t)bp * bp;    // 2 x bp: published row (phase 1) / column (phase 2), double buffered
  double* rinv_s = buf + 2 * bp;         // bp
  __shared__ int bad;
  const int tx = threadIdx.x % T, ty = threadIdx.x / T;
  const int r0 = 4 * ty, c0 = 4 * tx;
  double t[4][4];
  double d0[4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      const int a = r0 + i, c = c0 + jj;
      t[i][jj] = (a < b && c < b) ? G[min(a, c) + (size_t)max(a, c) * b] : (a == c ? 1.0 : 0.0);
    }
#pragma unroll
  for (int i = 0; i < 4; ++i) d0[i] = t[i][i];  // meaningful on the diagonal threads only
  if (threadIdx.x == 0) bad = 0;
  __syncthreads();
  chol_inv_tile_body(T, true, tx, ty, t, d0, Rs, buf, rinv_s, &bad);
  if (bad) {
    if (threadIdx.x == 0) flag[0] = 1.0;
    return;
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      const int a = r0 + i, c = c0 + jj;
      if (a < b && c < b) Tout[a + (size_t)c * b] = (a <= c) ? t[i][jj] : 0.0;
    }
  if (threadIdx.x == 0) flag[0] = 0.0;
}

__global__ void invert_upper_kernel(int k, const double* __restrict__ R, int64_t ld, double* __restrict__ X) {
  // column j of X solves R x = e_j (x_i = 0 for i > j), back substitution
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < k; j += gridDim.x * blockDim.x) {
    double* x = X + (size_t)j * k;
    for (int i = k - 1; i > j; --i) x[i] = 0.0;
    for (int i = j; i >= 0; --i) {
      double s = (i == j) ? 1.0 : 0.0;
      for (int l = i + 1; l <= j; ++l) s -= R[i + (size_t)l * ld] * x[l];
      x[i] = s / R[i + (size_t)i * ld];
    }
  }
}

// ---- fused block orthonormalisation (BCGS with the Pythagorean inner product) ---------------------------------
// Gall ((k+b) x b, ld k+b): rows 0..k = H = V^T C, rows k.. = C^T C.  P = H^T H.  Forms the Gram matrix of the
// projected block G' = C^T C - H^T H, scales it to unit diagonal (Gs = D G' D, D = diag(G')^-1/2), and measures how
// far the block was from being orthonormal to V and to itself:
//   metrics[0] = max_kj |H_kj| / sqrt((C^T C)_jj)      metrics[1] = max_ij |Gs_ij - delta_ij|
//   metrics[2] = 1.0 when a diagonal entry of G' is not safely positive (block numerically inside span(V))
__global__ void __launch_bounds__(1024) pip_prepare_kernel(int k, int b, const double* __restrict__ Gall,
                                                           const double* __restrict__ P, double* __restrict__ Gs,
                                                           double* __restrict__ D, double* __restrict__ metrics) {
  extern __shared__ __align__(16) double sm[];
  double* dsc = sm;       // b: D
  double* cn = sm + b;    // b: 1 / sqrt((C^T C)_jj)
  __shared__ double red[32];
  __shared__ int bad;
  const int ld = k + b, tid = threadIdx.x, nt = blockDim.x;
  if (tid == 0) bad = 0;
  __syncthreads();
  for (int j = tid; j < b; j += nt) {
    const double cc = Gall[(size_t)j * ld + k + j];
    const double d = cc - P[(size_t)j * b + j];
    // the projected column must keep a safe fraction of its norm, otherwise G' is round-off
    if (!(cc > 0.0) || !(d > 1e-10 * cc)) { bad = 1; dsc[j] = 0.0; cn[j] = 0.0; }
    else { dsc[j] = rsqrt(d); cn[j] = rsqrt(cc); }
  }
  __syncthreads();
  double m1 = 0.0, m0 = 0.0;
  for (int e = tid; e < b * b; e += nt) {
    const int i = e % b, j = e / b;
    const double g = (Gall[(size_t)j * ld + k + i] - P[e]) * dsc[i] * dsc[j];
    Gs[e] = g;
    const double dev = fabs(g - (i == j ? 1.0 : 0.0));
    m1 = (dev == dev) ? fmax(m1, dev) : 1.0e300;
  }
  for (int e = tid; e < k * b; e += nt) {
    const int i = e % k, j = e / k;
    const double h = fabs(Gall[(size_t)j * ld + i]) * cn[j];
    m0 = (h == h) ? fmax(m0, h) : 1.0e300;
  }
  for (int j = tid; j < b; j += nt) D[j] = dsc[j];
  m0 = block_reduce_max(m0, red);
  m1 = block_reduce_max(m1, red);
  if (tid == 0) {
    metrics[0] = m0;
    metrics[1] = m1;
    metrics[2] = bad ? 1.0 : 0.0;
  }
}

// M ((k+b) x b, ld k+b): rows k.. <- Tm = diag(D) * Tinv (upper triangular); Tm also stored densely (b x b) for the
// product H * Tm that fills rows 0..k.
__global__ void pip_finish_kernel(int k, int b, const double* __restrict__ Tinv, const double* __restrict__ D,
                                  double* __restrict__ Tm, double* __restrict__ M) {
  const int ld = k + b;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < b * b; e += gridDim.x * blockDim.x) {
    const int i = e % b, j = e / b;
    const double t = Tinv[e] * D[i];
    Tm[e] = t;
    M[(size_t)j * ld + k + i] = t;
  }
}

// ---- r02: everything between the tall-skinny products of one BCGS-PIP pass in ONE kernel (was: small GEMM H^T H,
// pip_prepare, chol_inv_upper, pip_finish, small GEMM H * Tm = 5 launches, 56 us at b = 32 and 103 us at b = 64).
// Gall ((kold + b) x b, ld kold + b): rows 0..kold = H = V^T C, rows kold.. = C^T C.  Z (same shape) <- [-H Tm; Tm]
// with Tm^T (C^T C - H^T H) Tm = I:
//   MODE 0 (first pass)   Tm = D R^-1, D G' D = R^T R: scaled Cholesky + inverse on the register tiles below
//   MODE 1 (second pass)  the block is already orthonormal to ~eps cond^2, G' = I + E with |E| << 1:
//                         Tm = (I + E)^-1/2 = I - E/2 + 3/8 E^2 + O(E^3) -- no factorisation, no sequential steps;
//                         metrics[2] is raised when |E| >= 1e-5 (the caller then takes the fallback path)
// metrics[0] = max |H_kj| / sqrt((C^T C)_jj), [1] = max |D G' D - I| (MODE 1: max |E|), [2] = 1.0 when a diagonal
// entry of G' is not safely positive, [3] = 1.0 when a Cholesky pivot is unsafe.
template <int MODE, int NTMAX>
__global__ void __launch_bounds__(NTMAX) pip_small_kernel(int kold, int b, const double* __restrict__ Gall,
                                                         double* __restrict__ Z, double* __restrict__ metrics) {
  extern __shared__ __align__(16) double sm[];
  const int T = (b + 3) >> 2, bp = 4 * T, ld = kold + b;
  double* Hs = sm;                       // kold x bp row-major: Hs[k * bp + j] = H(k, j), 0 for j >= b
  double* W1 = Hs + (size_t)kold * bp;   // bp x bp: rows of R (MODE 0) / E (MODE 1)
  double* W2 = W1 + (size_t)bp * bp;     // bp x bp row-major: Tm
  double* buf = W2 + (size_t)bp * bp;    // 2 bp
  double* rinv_s = buf + 2 * bp;         // bp
  double* dsc = rinv_s + bp;             // bp: D = diag(G')^-1/2 (MODE 1: 1)
  double* cn = dsc + bp;                 // bp: 1 / sqrt((C^T C)_jj)
  __shared__ double red[32];
  __shared__ int bad, badd;
  const int tid = threadIdx.x, nt = blockDim.x;
  const bool active = tid < T * T;
  const int tx = tid % T, ty = tid / T;
  const int r0 = 4 * ty, c0 = 4 * tx;
  if (tid == 0) { bad = 0; badd = 0; }
  for (int e = tid; e < kold * bp; e += nt) {
    const int k = e / bp, j = e - k * bp;
    Hs[e] = j < b ? Gall[(size_t)j * ld + k] : 0.0;
  }
  __syncthreads();

  // G' = C^T C - H^T H on the 4 x 4 register tiles of the T x T thread grid
  double t[4][4], d0[4];
  if (active) {
    double acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) acc[i][jj] = 0.0;
    for (int k = 0; k < kold; ++k) {
      const double2 a01 = *reinterpret_cast<const double2*>(Hs + (size_t)k * bp + r0);
      const double2 a23 = *reinterpret_cast<const double2*>(Hs + (size_t)k * bp + r0 + 2);
      const double2 b01 = *reinterpret_cast<const double2*>(Hs + (size_t)k * bp + c0);
      const double2 b23 = *reinterpret_cast<const double2*>(Hs + (size_t)k * bp + c0 + 2);
      const double a[4] = {a01.x, a01.y, a23.x, a23.y}, c[4] = {b01.x, b01.y, b23.x, b23.y};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) acc[i][jj] = fma(a[i], c[jj], acc[i][jj]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        const int r = r0 + i, c = c0 + jj;
        t[i][jj] = (r < b && c < b) ? Gall[(size_t)max(r, c) * ld + kold + min(r, c)] - acc[i][jj] : (r == c ? 1.0 : 0.0);
      }
    if (tx == ty) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r = r0 + i;
        if (r < b) {
          const double cc = Gall[(size_t)r * ld + kold + r], d = t[i][i];
          // the projected column must keep a safe fraction of its norm, otherwise G' is round-off
          if (!(cc > 0.0) || !(d > 1e-10 * cc)) { badd = 1; dsc[r] = 0.0; cn[r] = 0.0; }
          else { dsc[r] = MODE == 0 ? rsqrt(d) : 1.0; cn[r] = rsqrt(cc); }
        } else {
          dsc[r] = 1.0;
          cn[r] = 0.0;
        }
      }
    }
  }
  __syncthreads();
  double m0 = 0.0, m1 = 0.0;
  if (active) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        const int r = r0 + i, c = c0 + jj;
        if (r < b && c < b) t[i][jj] *= dsc[r] * dsc[c];
        const double dev = fabs(t[i][jj] - (r == c ? 1.0 : 0.0));
        m1 = (dev == dev) ? fmax(m1, dev) : 1.0e300;
        if (MODE == 1) W1[(size_t)r * bp + c] = t[i][jj] - (r == c ? 1.0 : 0.0);  // E
      }
#pragma unroll
    for (int i = 0; i < 4; ++i) d0[i] = t[i][i];
  }
  for (int e = tid; e < kold * bp; e += nt) {
    const int j = e % bp;
    const double h = fabs(Hs[e]) * cn[j];
    m0 = (h == h) ? fmax(m0, h) : 1.0e300;
  }
  m0 = block_reduce_max(m0, red);
  m1 = block_reduce_max(m1, red);  // (ends with a barrier: E is complete in W1)

  if (MODE == 0) {
    chol_inv_tile_body(T, active, tx, ty, t, d0, W1, buf, rinv_s, &bad);
    if (bad || badd) {
      if (tid == 0) { metrics[0] = m0; metrics[1] = m1; metrics[2] = badd ? 1.0 : 0.0; metrics[3] = bad ? 1.0 : 0.0; }
      return;  // uniform; Z is not used by the caller in this case
    }
    if (active) {
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          const int r = r0 + i, c = c0 + jj;
          t[i][jj] = (r <= c && r < b && c < b) ? t[i][jj] * dsc[r] : 0.0;
        }
    }
  } else {
    if (active) {
      double acc[4][4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) acc[i][jj] = 0.0;
      for (int l = 0; l < bp; ++l) {  // E^2 (E symmetric: column l of E = row l)
        const double2 a01 = *reinterpret_cast<const double2*>(W1 + (size_t)l * bp + r0);
        const double2 a23 = *reinterpret_cast<const double2*>(W1 + (size_t)l * bp + r0 + 2);
        const double2 b01 = *reinterpret_cast<const double2*>(W1 + (size_t)l * bp + c0);
        const double2 b23 = *reinterpret_cast<const double2*>(W1 + (size_t)l * bp + c0 + 2);
        const double a[4] = {a01.x, a01.y, a23.x, a23.y}, c[4] = {b01.x, b01.y, b23.x, b23.y};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) acc[i][jj] = fma(a[i], c[jj], acc[i][jj]);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          const int r = r0 + i, c = c0 + jj;
          const double E = t[i][jj] - (r == c ? 1.0 : 0.0);
          t[i][jj] = (r < b && c < b) ? (r == c ? 1.0 : 0.0) - 0.5 * E + 0.375 * acc[i][jj] : 0.0;
        }
    }
  }
  if (active) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        const int r = r0 + i, c = c0 + jj;
        W2[(size_t)r * bp + c] = t[i][jj];
        if (r < b && c < b) Z[(size_t)c * ld + kold + r] = t[i][jj];
      }
  }
  __syncthreads();
  // rows 0..kold of Z: -H * Tm, 4 x 4 register blocks over all threads
  const int ngrp = nt / T;  // row groups of 4 rows
  for (int i0 = 4 * (tid / T); i0 < kold; i0 += 4 * ngrp) {
    double acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) acc[i][jj] = 0.0;
    const int lmax = MODE == 0 ? min(bp, c0 + 4) : bp;  // Tm upper triangular in MODE 0
    for (int l = 0; l < lmax; ++l) {
      const double2 w01 = *reinterpret_cast<const double2*>(W2 + (size_t)l * bp + c0);
      const double2 w23 = *reinterpret_cast<const double2*>(W2 + (size_t)l * bp + c0 + 2);
      const double w[4] = {w01.x, w01.y, w23.x, w23.y};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const double h = (i0 + i < kold) ? Hs[(size_t)(i0 + i) * bp + l] : 0.0;
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) acc[i][jj] = fma(h, w[jj], acc[i][jj]);
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int jj = 0; jj < 4; ++jj)
        if (i0 + i < kold && c0 + jj < b) Z[(size_t)(c0 + jj) * ld + i0 + i] = -acc[i][jj];
  }
  if (tid == 0) {
    metrics[0] = m0;
    metrics[1] = m1;
    metrics[2] = (badd || (MODE == 1 && !(m1 < 1e-5))) ? 1.0 : 0.0;
    metrics[3] = 0.0;
  }
}

}  // namespace

void pip_prepare(cudaStream_t s, int k, int b, const double* Gall, const double* P, double* Gs, double* D,
                 double* metrics) {
  pip_prepare_kernel<<<1, 1024, 2 * (size_t)b * sizeof(double), s>>>(k, b, Gall, P, Gs, D, metrics);
  CK_LAUNCH();
  ++g_kernel_launches;
}

void pip_finish(cudaStream_t s, int k, int b, const double* Tinv, const double* D, double* Tm, double* M) {
  pip_finish_kernel<<<std::max(1, std::min(64, (b * b + 255) / 256)), 256, 0, s>>>(k, b, Tinv, D, Tm, M);
  CK_LAUNCH();
  ++g_kernel_launches;
}

bool pip_small(cudaStream_t s, int mode, int kold, int b, const double* Gall, double* Z, double* metrics) {
  static const bool off = [] { const char* e = std::getenv("DAV_PIP_FUSED"); return e && std::atoi(e) == 0; }();
  if (off || b < 1 || b > 128 || kold < 1) return false;
  const int T = (b + 3) / 4, bp = 4 * T;
  const size_t bytes = ((size_t)kold * bp + 2 * (size_t)bp * bp + 5 * (size_t)bp) * sizeof(double);
  const int max_smem = device_max_smem_optin();
  if (bytes > (size_t)max_smem - 1024) return false;
  const int nt = std::min(1024, std::max(256, (T * T + 31) / 32 * 32));
#define DAV_PIP_LAUNCH(MODE_, NTMAX_)                                                      \
  do {                                                                                     \
    ensure_dyn_smem(pip_small_kernel<MODE_, NTMAX_>, max_smem - 1024);                     \
    pip_small_kernel<MODE_, NTMAX_><<<1, nt, bytes, s>>>(kold, b, Gall, Z, metrics);       \
  } while (0)
  if (mode == 0) {
    if (nt <= 256) DAV_PIP_LAUNCH(0, 256); else DAV_PIP_LAUNCH(0, 1024);
  } else {
    if (nt <= 256) DAV_PIP_LAUNCH(1, 256); else DAV_PIP_LAUNCH(1, 1024);
  }
#undef DAV_PIP_LAUNCH
  CK_LAUNCH();
  ++g_kernel_launches;
  return true;
}

void jacobi_eigh(cudaStream_t s, int k, double* S, double* Y, double* w, double* scratch, int* status,
                 const int* skip) {
  if (k <= 0) return;
  const int kp = k + (k & 1), np = kp / 2;
  const int max_smem = device_max_smem_optin();
  ensure_dyn_smem(jacobi_kernel, max_smem);
  ensure_dyn_smem(jacobi_sweeps_kernel, max_smem);
  ensure_dyn_smem(jacobi_vectors_kernel, max_smem);
  // ---- fast path: S in shared memory, rotations logged, V replayed by a second kernel
  {
    const int nitems = np * (np + 1) / 2;
    size_t small = 32 + 2 * (size_t)np + kp + (size_t)np + ((size_t)nitems * 4 + 7) / 8;
    small = (small + 1) & ~(size_t)1;
    const size_t need = (small + (size_t)kp * (kp + 1)) * sizeof(double);
    // scratch layout: [rotation log: max_rounds * np double2][rank: k ints][nrounds: 1 int]
    const size_t scratch_doubles = jacobi_scratch_doubles(k);
    const size_t tail = ((size_t)k + 2) / 2 + 2;  // doubles reserved for the int arrays
    const int max_rounds = (int)std::min<size_t>((scratch_doubles - tail) / (2 * (size_t)np), (size_t)MAX_SWEEPS * (kp - 1));
    if (need <= (size_t)max_smem && max_rounds >= 8 * (kp - 1) && np <= VPARTS * VMAXI) {
      double2* rotlog = reinterpret_cast<double2*>(scratch);
      int* rank = reinterpret_cast<int*>(scratch + scratch_doubles - tail);
      int* nrounds = rank + k;
      const int threads = kp <= 32 ? 128 : (kp <= 64 ? 256 : 512);
      jacobi_sweeps_kernel<<<1, threads, need, s>>>(k, S, w, rank, rotlog, nrounds, max_rounds, status, skip);
      CK_LAUNCH();
      ++g_kernel_launches;
      const size_t vsm = (2 * (size_t)VCHUNK * np + (size_t)VROWS * (kp + 1)) * sizeof(double);
      jacobi_vectors_kernel<<<(k + VROWS - 1) / VROWS, VROWS * VPARTS, vsm, s>>>(k, rotlog, nrounds, rank, Y, skip);
      CK_LAUNCH();
      ++g_kernel_launches;
      return;
    }
  }
  // ---- general path (large k): S and/or V in L2-resident global scratch
  size_t small = 32 + 2 * (size_t)np + kp + ((2 * (size_t)np + kp) * sizeof(int) + 7) / 8;
  small = (small + 1) & ~(size_t)1;
  const size_t big = (size_t)kp * kp;
  const size_t cap = (size_t)max_smem / sizeof(double);
  int s_in = 0, v_in = 0;
  size_t doubles = small;
  if (small + 2 * big <= cap) { s_in = 1; v_in = 1; doubles += 2 * big; }
  else if (small + big <= cap) { s_in = 1; doubles += big; }
  jacobi_kernel<<<1, JT, doubles * sizeof(double), s>>>(k, S, Y, w, scratch, scratch + big, s_in, v_in, status, skip);
  CK_LAUNCH();
  ++g_kernel_launches;
}

void scale_cols_rsqrt_checked(cudaStream_t s, int k, const double* U, const double* sv, double* T, int* status) {
  scale_cols_rsqrt_checked_kernel<<<1, 1024, 0, s>>>(k, U, sv, T, status);
  CK_LAUNCH();
  ++g_kernel_launches;
}

void gram_prescale(cudaStream_t s, int k, double* G, double* D) {
  gram_prescale_kernel<<<1, 1024, 0, s>>>(k, G, D);
  CK_LAUNCH();
  ++g_kernel_launches;
}

void svqb_make_T(cudaStream_t s, int k, const double* U, const double* sv, const double* D, double* T, int* flags) {
  svqb_make_T_kernel<<<1, 1024, 0, s>>>(k, U, sv, D, T, flags);
  CK_LAUNCH();
  ++g_kernel_launches;
}

void symmetrize_from_upper(cudaStream_t s, int k, double* S, int64_t ld) {
  const int blocks = std::max(1, std::min(64, (k * k + 255) / 256));
  symmetrize_from_upper_kernel<<<blocks, 256, 0, s>>>(k, S, ld);
  CK_LAUNCH();
  ++g_kernel_launches;
}

void max_abs_dev(cudaStream_t s, int rows, int cols, const double* G, int64_t ld, bool minus_identity, double* out) {
  max_abs_kernel<<<1, 1024, 0, s>>>(rows, cols, G, ld, minus_identity ? 1 : 0, out);
  CK_LAUNCH();
  ++g_kernel_launches;
}

void cholesky_upper(cudaStream_t s, int k, double* G, int64_t ld, int* status) {
  cholesky_upper_kernel<<<1, 1024, 0, s>>>(k, G, ld, status);
  CK_LAUNCH();
  ++g_kernel_launches;
}

bool chol_inv_upper(cudaStream_t s, int b, const double* G, double* T_, double* flag, double* gwork,
                    size_t gwork_doubles) {
  double* T = T_;
  const int max_smem = device_max_smem_optin();
  static const bool tile_off = [] { const char* e = std::getenv("DAV_CHOL_TILE"); return e && std::atoi(e) == 0; }();
  if (b <= 128 && !tile_off) {
    const int T = (b + 3) / 4, bp = 4 * T;
    const size_t bytes = ((size_t)bp * bp + 3 * (size_t)bp) * sizeof(double);
    if (bytes <= (size_t)max_smem - 1024) {
      ensure_dyn_smem(chol_inv_tile_kernel, max_smem - 1024);
      chol_inv_tile_kernel<<<1, T * T, bytes, s>>>(b, G, T_, flag);
      CK_LAUNCH();
      ++g_kernel_launches;
      return true;
    }
  }
  ensure_dyn_smem(chol_inv_upper_kernel, max_smem - 1024);
  const size_t tri2 = (size_t)b * (b + 1);
  const size_t bytes = (tri2 + b) * sizeof(double);
  const int threads = b <= 32 ? 128 : (b <= 64 ? 256 : 512);
  if (bytes <= (size_t)max_smem - 1024) {
    chol_inv_upper_kernel<<<1, threads, bytes, s>>>(b, G, T, flag, nullptr);
  } else {
    // wide blocks (b >= ~170): both triangles live in L2-resident global scratch, only the diagonal in shared memory
    if (!gwork || gwork_doubles < tri2) return false;
    chol_inv_upper_kernel<<<1, 512, (size_t)b * sizeof(double), s>>>(b, G, T, flag, gwork);
  }
  CK_LAUNCH();
  ++g_kernel_launches;
  return true;
}

void invert_upper(cudaStream_t s, int k, const double* R, int64_t ld, double* Rinv) {
  invert_upper_kernel<<<(k + 63) / 64, 64, 0, s>>>(k, R, ld, Rinv);
  CK_LAUNCH();
  ++g_kernel_launches;
}

}  // namespace dav
