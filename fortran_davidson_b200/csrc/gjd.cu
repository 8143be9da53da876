// GJD correction on device (replaces compute_GJD_generalized_dense, davidson.f90:700-734).
//
// The reference forms, per Ritz pair, three dense n x n temporaries, two n^3 DGEMMs and a DSYSV
// factorisation (O(k n^3) per iteration) to solve  xs*ys*xs t = -r  with xs = I - u u^T,
// ys = A - theta B.  Here the same correction equation is solved matrix-free for ALL k Ritz pairs in
// lock step by diagonally preconditioned MINRES, so one inner iteration costs one block matvec
// A*P (and B*P):
//       (I - w u^T)(A - theta_j B)(I - u w^T) t_j = -r_j ,   w = B u_j   (w = u_j without second_matrix)
// For the standard problem this is exactly the reference's operator.  For the generalized problem the
// reference's literal operator (u u^T with a B-normalised u) has the exact solution -u/(1 - u^T u),
// i.e. no new direction -- it only converges through DSYSV round-off -- so the B-orthogonal (textbook)
// projector is used; outer iteration counts agree with the oracle to +-1 (tests/device_model.py
// restates this solver in numpy and tests/test_device_model.py checks it against the oracle).
// The preconditioner is projected as well,  K~^-1 r = K^-1 r - z (w^T K^-1 r)/(w^T z),  z = K^-1 w,
// K = |theta_j diag(B) - diag(A)|: the Krylov vectors then stay in the complement of u where the
// projected operator is non-singular (without it MINRES diverges once the residual is small).
#include <algorithm>
#include <cmath>

#include "solver.cuh"

namespace dav {
namespace {

constexpr int NCH = 64;  // row chunks of the two-stage column reductions (== vecops.cu NCHUNK)
// inner (MINRES) stopping rule, relative to ||r_j||: 1e-8 for the usual outer tolerances, tightened with the outer
// tolerance below that (an outer 1e-12 gets a 1e-12 inner solve and 8 more inner iterations per decade)
constexpr double GJD_RTOL_MAX = 1e-8;
constexpr int GJD_MAXIT_BASE = 40;
constexpr double GJD_DFLOOR = 1e-8;

enum { S_BETA, S_OLDB, S_BETA1, S_DBAR, S_EPSLN, S_OLDEPS, S_PHIBAR, S_CS, S_SN, S_ALFA, S_DELTA, S_GAMMA, S_PHI,
       S_D0, S_D1, S_WZ, S_WY, S_BETASQ, S_ROWS };

struct Vecs {
  int64_t nl, ld;
  int k, kcap;
  const double *theta, *dA, *dB;  // dB may be null
  const double *u, *w, *R;
  double *x, *r1, *r2, *y, *v, *w1, *w2, *wc, *z, *P, *AP, *BP;  // BP null when !gev
  double* st;       // S_ROWS x kcap scalars
  int* active;      // current / next flags: active[0..k), active[kcap..kcap+k)
  double* partial;  // k x NCH
};

__device__ __forceinline__ double dinv_at(const Vecs& a, int64_t i, double th) {
  const double d = fabs(a.dA[i] - th * (a.dB ? a.dB[i] : 1.0));
  return 1.0 / fmax(d, GJD_DFLOOR);
}

__device__ __forceinline__ void block_partial(double s, double* partial, int j, int c) {
  __shared__ double red[32];
  s = warp_sum(s);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) red[wid] = s;
  __syncthreads();
  if (wid == 0) {
    double x = lane < (blockDim.x >> 5) ? red[lane] : 0.0;
    x = warp_sum(x);
    if (lane == 0) partial[(size_t)j * NCH + c] = x;
  }
}

#define GJD_COLUMN_LOOP                                             \
  const int j = blockIdx.y, c = blockIdx.x;                         \
  const int64_t per = (a.nl + NCH - 1) / NCH;                       \
  const int64_t beg = (int64_t)c * per, end = min(a.nl, beg + per); \
  const int64_t off = (int64_t)j * a.ld;

// phase codes of the fused vector kernel
enum { PH_INIT_Z, PH_INIT_R, PH_INIT_Y, PH_A, PH_B, PH_C, PH_D, PH_E, PH_F, PH_G };

template <int PH>
__global__ void __launch_bounds__(256) gjd_vec_kernel(Vecs a, int itn) {
  GJD_COLUMN_LOOP
  if (PH >= PH_A && !a.active[j]) return;
  const double th = a.theta[j];
  double* st = a.st;
  const int K = a.kcap;
  double s = 0.0;
  if (PH == PH_INIT_Z) {  // z = K^-1 w ; partial <- w.z
    for (int64_t i = beg + threadIdx.x; i < end; i += blockDim.x) {
      const double w = a.w[off + i], z = w * dinv_at(a, i, th);
      a.z[off + i] = z;
      s = fma(w, z, s);
    }
  } else if (PH == PH_INIT_R) {  // r1 = r2 = -R ; y0 = K^-1 r1 ; x = w1 = w2 = 0 ; partial <- w.y0
    for (int64_t i = beg + threadIdx.x; i < end; i += blockDim.x) {
      const double r = -a.R[off + i], y0 = r * dinv_at(a, i, th);
      a.r1[off + i] = r; a.r2[off + i] = r; a.y[off + i] = y0;
      a.x[off + i] = 0.0; a.w1[off + i] = 0.0; a.w2[off + i] = 0.0; a.wc[off + i] = 0.0;
      s = fma(a.w[off + i], y0, s);
    }
  } else if (PH == PH_INIT_Y || PH == PH_F) {  // y = y0 - z (w.y0)/(w.z) ; partial <- r2.y
    const double f = st[S_WY * K + j] / st[S_WZ * K + j];
    for (int64_t i = beg + threadIdx.x; i < end; i += blockDim.x) {
      const double y = a.y[off + i] - a.z[off + i] * f;
      a.y[off + i] = y;
      s = fma(a.r2[off + i], y, s);
    }
  } else if (PH == PH_A) {  // v = y / beta ; partial <- w.v
    const double sc = 1.0 / st[S_BETA * K + j];
    for (int64_t i = beg + threadIdx.x; i < end; i += blockDim.x) {
      const double v = a.y[off + i] * sc;
      a.v[off + i] = v;
      s = fma(a.w[off + i], v, s);
    }
  } else if (PH == PH_B) {  // P = v - u (w.v)
    const double d0 = st[S_D0 * K + j];
    for (int64_t i = beg + threadIdx.x; i < end; i += blockDim.x) a.P[off + i] = a.v[off + i] - a.u[off + i] * d0;
    return;
  } else if (PH == PH_C) {  // q = A P - theta (B P | P) -> y ; partial <- u.q
    for (int64_t i = beg + threadIdx.x; i < end; i += blockDim.x) {
      const double q = a.AP[off + i] - th * (a.BP ? a.BP[off + i] : a.P[off + i]);
      a.y[off + i] = q;
      s = fma(a.u[off + i], q, s);
    }
  } else if (PH == PH_D) {  // y = q - w (u.q) [- (beta/oldb) r1] ; partial <- v.y
    const double d1 = st[S_D1 * K + j];
    const double f = (itn >= 2) ? st[S_BETA * K + j] / st[S_OLDB * K + j] : 0.0;
    for (int64_t i = beg + threadIdx.x; i < end; i += blockDim.x) {
      double y = a.y[off + i] - a.w[off + i] * d1;
      if (itn >= 2) y -= f * a.r1[off + i];
      a.y[off + i] = y;
      s = fma(a.v[off + i], y, s);
    }
  } else if (PH == PH_E) {  // y -= (alfa/beta) r2 ; r2new = y (stored over r1; host swaps) ; y0 = K^-1 r2new ; partial <- w.y0
    const double f = st[S_ALFA * K + j] / st[S_BETA * K + j];
    for (int64_t i = beg + threadIdx.x; i < end; i += blockDim.x) {
      const double r = a.y[off + i] - f * a.r2[off + i];
      a.r1[off + i] = r;
      const double y0 = r * dinv_at(a, i, th);
      a.y[off + i] = y0;
      s = fma(a.w[off + i], y0, s);
    }
  } else if (PH == PH_G) {  // wnew = (v - oldeps w1 - delta w2)/gamma -> wc (host rotates the three buffers) ; x += phi wnew
    const double oe = st[S_OLDEPS * K + j], de = st[S_DELTA * K + j], ga = 1.0 / st[S_GAMMA * K + j],
                 phi = st[S_PHI * K + j];
    for (int64_t i = beg + threadIdx.x; i < end; i += blockDim.x) {
      const double wn = (a.v[off + i] - oe * a.w1[off + i] - de * a.w2[off + i]) * ga;
      a.wc[off + i] = wn;
      a.x[off + i] += phi * wn;
    }
    return;
  }
  block_partial(s, a.partial, j, c);
}

// dst[row][j] = sum_c partial[j][c]
__global__ void gjd_reduce_kernel(int k, int kcap, const double* __restrict__ partial, double* st, int row) {
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < k; j += gridDim.x * blockDim.x) {
    double s = 0.0;
    for (int c = 0; c < NCH; ++c) s += partial[(size_t)j * NCH + c];
    st[(size_t)row * kcap + j] = s;
  }
}

__global__ void gjd_scalar_init(int k, int K, double* st, int* active) {
  for (int j = threadIdx.x; j < k; j += blockDim.x) {
    const double b1sq = st[S_BETASQ * K + j];
    const bool ok = b1sq > 0.0 && st[S_WZ * K + j] > 0.0;
    const double b1 = ok ? sqrt(b1sq) : 0.0;
    st[S_BETA1 * K + j] = b1; st[S_BETA * K + j] = b1; st[S_OLDB * K + j] = 0.0; st[S_DBAR * K + j] = 0.0;
    st[S_EPSLN * K + j] = 0.0; st[S_PHIBAR * K + j] = b1; st[S_CS * K + j] = -1.0; st[S_SN * K + j] = 0.0;
    active[j] = ok ? 1 : 0;
    active[K + j] = ok ? 1 : 0;
  }
}

// Lanczos / Givens recurrences of MINRES (Paige & Saunders) for every active column
__global__ void gjd_scalar_step(int k, int K, double* st, int* active, double rtol) {
  for (int j = threadIdx.x; j < k; j += blockDim.x) {
    if (!active[j]) continue;
    const double betasq = st[S_BETASQ * K + j];
    const bool bad = !(betasq >= 0.0);
    const double oldb = st[S_BETA * K + j];
    const double beta = bad ? 0.0 : sqrt(betasq);
    const double alfa = st[S_ALFA * K + j];
    const double cs = st[S_CS * K + j], sn = st[S_SN * K + j], dbar = st[S_DBAR * K + j];
    const double oldeps = st[S_EPSLN * K + j];
    const double delta = cs * dbar + sn * alfa;
    const double gbar = sn * dbar - cs * alfa;
    const double epsln = sn * beta;
    const double dbarn = -cs * beta;
    const double gamma = fmax(hypot(gbar, beta), 2.220446049250313e-16);
    const double csn = gbar / gamma, snn = beta / gamma;
    const double phibar = st[S_PHIBAR * K + j];
    const double phi = csn * phibar, phibarn = snn * phibar;
    st[S_OLDB * K + j] = oldb; st[S_BETA * K + j] = beta; st[S_OLDEPS * K + j] = oldeps; st[S_DELTA * K + j] = delta;
    st[S_EPSLN * K + j] = epsln; st[S_DBAR * K + j] = dbarn; st[S_GAMMA * K + j] = gamma; st[S_CS * K + j] = csn;
    st[S_SN * K + j] = snn; st[S_PHI * K + j] = bad ? 0.0 : phi; st[S_PHIBAR * K + j] = phibarn;
    const bool go = !bad && beta > 0.0 && phibarn > rtol * st[S_BETA1 * K + j];
    active[K + j] = go ? 1 : 0;
  }
}

__global__ void gjd_commit_active(int k, int K, int* active, int* nactive) {
  __shared__ int cnt;
  if (threadIdx.x == 0) cnt = 0;
  __syncthreads();
  for (int j = threadIdx.x; j < k; j += blockDim.x) {
    const int a = active[K + j];
    active[j] = a;
    if (a) atomicAdd(&cnt, 1);
  }
  __syncthreads();
  if (threadIdx.x == 0) *nactive = cnt;
}

}  // namespace
}  // namespace dav

using namespace dav;

// C <- GJD corrections for all k Ritz pairs.  On entry R holds the residuals and C holds
// w = (BV | V) * Y (residual_dpr leaves it untouched when no DPR correction is written).
void dav_solver::gjd_correction(int k, bool gev, double outer_tolerance) {
  const double GJD_RTOL = std::min(GJD_RTOL_MAX, std::max(outer_tolerance, 1e-14));
  const int GJD_MAXIT =
      GJD_MAXIT_BASE + (GJD_RTOL < GJD_RTOL_MAX ? (int)std::ceil(8.0 * std::log10(GJD_RTOL_MAX / GJD_RTOL)) : 0);
  const size_t nk = (size_t)ldv * (kcap / 2 > 0 ? std::max(kcap / 2, k) : k);
  const int nbuf = 13;
  gjd_buf.alloc(nk * nbuf);
  gjd_st.alloc((size_t)S_ROWS * kcap);
  gjd_active.alloc(2 * (size_t)kcap + 1);
  double* base = gjd_buf.p;
  auto take = [&]() { double* p = base; base += nk; return p; };
  Vecs a;
  a.nl = nl; a.ld = ldv; a.k = k; a.kcap = kcap;
  a.theta = theta.p; a.dA = mat[0].diag.p; a.dB = gev ? mat[1].diag.p : nullptr;
  double* Ug = take();
  a.x = take(); a.r1 = take(); a.r2 = take(); a.y = take(); a.v = take(); a.w1 = take(); a.w2 = take(); a.wc = take();
  a.z = take(); a.P = take(); a.AP = take();
  double* BPbuf = take();
  a.BP = gev ? BPbuf : nullptr;
  a.w = C.p;
  if (gev) {
    gemm(stream, false, nl, k, k, 1.0, V.p, ldv, Y.p, k, 0.0, Ug, ldv, nullptr, 0);  // u = V*Y
    a.u = Ug;
  } else {
    a.u = C.p;  // w == u
  }
  a.R = R.p;
  a.st = gjd_st.p;
  a.active = gjd_active.p;
  a.partial = partial.p;
  int* nactive = gjd_active.p + 2 * (size_t)kcap;
  const dim3 grid(NCH, k);
  auto reduce_to = [&](int row) {
    gjd_reduce_kernel<<<(k + 127) / 128, 128, 0, stream>>>(k, kcap, a.partial, a.st, row);
    CK_LAUNCH();
    ++g_kernel_launches;
    allreduce(a.st + (size_t)row * kcap, k);
  };
#define GJD_VEC(PH, itn)                                         \
  do {                                                           \
    gjd_vec_kernel<PH><<<grid, 256, 0, stream>>>(a, itn);        \
    CK_LAUNCH();                                                 \
    ++g_kernel_launches;                                         \
  } while (0)

  GJD_VEC(PH_INIT_Z, 0); reduce_to(S_WZ);
  GJD_VEC(PH_INIT_R, 0); reduce_to(S_WY);
  GJD_VEC(PH_INIT_Y, 0); reduce_to(S_BETASQ);
  gjd_scalar_init<<<1, 256, 0, stream>>>(k, kcap, a.st, a.active);
  CK_LAUNCH();
  ++g_kernel_launches;

  for (int itn = 1; itn <= GJD_MAXIT; ++itn) {
    GJD_VEC(PH_A, itn); reduce_to(S_D0);
    GJD_VEC(PH_B, itn);
    apply_both(a.P, k, a.AP, gev ? a.BP : nullptr);  // ONE gather of P feeds A*P and B*P
    GJD_VEC(PH_C, itn); reduce_to(S_D1);
    GJD_VEC(PH_D, itn); reduce_to(S_ALFA);
    GJD_VEC(PH_E, itn); reduce_to(S_WY);
    std::swap(a.r1, a.r2);  // r1 <- old r2, r2 <- new residual (written over the old r1)
    GJD_VEC(PH_F, itn); reduce_to(S_BETASQ);
    gjd_scalar_step<<<1, 256, 0, stream>>>(k, kcap, a.st, a.active, GJD_RTOL);
    CK_LAUNCH();
    ++g_kernel_launches;
    // w_new = (v - oldeps*w2_old - delta*w_old)/gamma, x += phi*w_new: reads a.w1 (= w2_old) and a.w2 (= w_old),
    // writes a.wc (the dead buffer); then rotate (w1, w2, wc) <- (w2, wc, w1)
    GJD_VEC(PH_G, itn);
    {
      double* dead = a.w1;
      a.w1 = a.w2; a.w2 = a.wc; a.wc = dead;
    }
    gjd_commit_active<<<1, 256, 0, stream>>>(k, kcap, a.active, nactive);
    CK_LAUNCH();
    ++g_kernel_launches;
    int h = 0;
    CK(cudaMemcpyAsync(&h, nactive, sizeof(int), cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    stats.gjd_inner_iterations += 1;
    if (h == 0) break;
  }
  copy_matrix(stream, nl, k, a.x, ldv, C.p, ldv);
#undef GJD_VEC
}
