// Matrix-free block matvec on the FP64 tensor pipe: W(rows row0 .. row0+nl, b) = Op * X(n x b) for the reference's
// on-the-fly operators (benchmark_free.f90:38-76, tests/test_utils.f90:37-116) applied as free_matmul does
// (davidson.f90:526-569).  Nothing n x n ever exists: every CTA GENERATES its 16-row x 32-column sub-tiles of the
// operator straight into shared memory, in the swizzled layout the DMMA fragment loads of matvec_dmma.cu use, and
// consumes them with mma.sync.m8n8k4.f64 against a pre-packed X tile.
//
// Entry generator.  a(i, l) = g(r) * 1e-4 (+ diagonal), r = e_lo / e_hi in (1/e, 1],
//   g(r) = cos(0.5 ln(atan r))  (sin for the test stx operator), e_t = dble(expf(t / n)) from the host-built table.
// libm's atan2 + log + sqrt + cos cost ~400 FP64-pipe operations per entry, 3x the 2*b flops the entry is used for.
// g is analytic on [1/e, 1], so the kernel evaluates a piecewise degree-5 polynomial (128 segments, coefficients
// fitted on the host in extended precision at Chebyshev nodes; measured max error 6e-16, i.e. the rounding of the
// evaluation itself; the builder verifies this against long-double libm on a dense sample and refuses the table
// otherwise): 1 multiplication by the tabulated reciprocal 1/e_hi, 3 LDS.128 and 5 FMA per entry.
// Bound: FP64 pipe, shared by DMMA and the generator (scripts/fp64_pipes.cu: DMMA and DFMA do not overlap on B200):
// (2 b + ~15) / (2 b) of the DMMA time.
#include <algorithm>
#include <cmath>
#include <type_traits>
#include <vector>

#include "kernels.cuh"

namespace dav {
namespace {

constexpr int FSEG = 128, FCOEF = 6;  // segments, coefficients per segment (degree 5)
constexpr int KS = 32;                // operator columns generated per step (two 16-column DMMA stages)

__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
      : "+d"(d0), "+d"(d1)
      : "d"(a), "d"(b));
}
__device__ __forceinline__ double2 lds128(uint32_t addr) {
  double2 v;
  asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

struct FreeParams {
  int op;
  int64_t n, row0, nl;
  int b;
  const double2* et;   // (e_t, 1 / e_t)
  const double* coef;  // FSEG x FCOEF
  double rlo, inv_h;
  const double* Xp;    // packed X: [k / 16][16 x BPAD] in fragment order (pack_x layout of matvec_dmma.cu)
  double* W;
  int64_t ldw;
};

// X packed index of element (k, j): ((k/8 * NTT + j/8) * 64 + (j%8)*8 + k%8)
__global__ void free_pack_x_kernel(int64_t K, int64_t Kpad, int b, int bpad, const double* __restrict__ X, int64_t ldx,
                                   double* __restrict__ Xp) {
  const int ntt = bpad / 8;
  const int64_t total = Kpad * bpad;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int kk = (int)(e & 7);
    const int g = (int)((e >> 3) & 7);
    const int64_t blk = e >> 6;
    const int jt = (int)(blk % ntt);
    const int64_t kq = blk / ntt;
    const int64_t k = kq * 8 + kk;
    const int j = jt * 8 + g;
    Xp[e] = (k < K && j < b) ? X[k + (int64_t)j * ldx] : 0.0;
  }
}

template <int NT, int WARPS_N>
__global__ void __launch_bounds__(256, 2) free_dmma_kernel(const FreeParams p) {
  constexpr int WARPS_M = 8 / WARPS_N;
  constexpr int BM = WARPS_M * 32;
  constexpr int SUBT = BM / 16;              // 16-row sub-tiles
  constexpr int NTT = NT * WARPS_N;
  constexpr int BPAD = NTT * 8;
  constexpr uint32_t A_STAGE = BM * 16 * 8;  // bytes of one 16-column stage
  constexpr uint32_t X_STAGE = 16 * BPAD * 8;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
  // layout: A[2 stages] | X[2 stages] | coef | column table (2 x KS double2)
  const uint32_t a_off = 0, x_off = 2 * A_STAGE, c_off = x_off + 2 * X_STAGE;
  double* coef_s = reinterpret_cast<double*>(gbase + c_off);
  double2* ct = reinterpret_cast<double2*>(gbase + c_off + FSEG * FCOEF * 8);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t i0 = (int64_t)blockIdx.x * BM;
  for (int e = tid; e < FSEG * FCOEF; e += 256) coef_s[e] = p.coef[e];
  if (tid < KS) {
    const int64_t gl = tid;
    ct[tid] = gl < p.n ? p.et[gl] : make_double2(1.0, 1.0);
  }

  // ---- generator mapping: a lane owns one row of a sub-tile and every second column of the step
  constexpr int SPW = SUBT >= 8 ? SUBT / 8 : 1;   // sub-tiles per warp
  constexpr int WPS = SUBT >= 8 ? 1 : 8 / SUBT;   // warps per sub-tile
  constexpr int NQ = (KS / 2) / WPS;              // columns per lane and sub-tile
  const int grow = lane & 15, gpar = lane >> 4;
  double2 erow[SPW];
  int64_t girow[SPW];
#pragma unroll
  for (int s = 0; s < SPW; ++s) {
    const int st = SUBT >= 8 ? warp * SPW + s : warp / WPS;
    const int64_t li = i0 + st * 16 + grow;
    girow[s] = p.row0 + li;
    erow[s] = (li < p.nl) ? p.et[girow[s]] : make_double2(1.0, 1.0);
    if (li >= p.nl) girow[s] = -1;  // never equals a column index: no diagonal term
  }
  const int kk_base = gpar + (SUBT >= 8 ? 0 : (warp % WPS) * (KS / WPS));
  const bool is_stx = p.op == DAV_OP_TEST_STX;

  // ---- consumer mapping (as matvec_dmma.cu)
  const int wr = warp / WARPS_N, wc = warp % WARPS_N;
  const int g = lane >> 2, t = lane & 3;
  double acc[2][2][NT][2];
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int e = 0; e < 2; ++e)
#pragma unroll
      for (int nn = 0; nn < NT; ++nn) acc[a][e][nn][0] = acc[a][e][nn][1] = 0.0;
  const uint32_t a_lane = a_off + (uint32_t)(wr * 2) * (16 * 128);
  const uint32_t x_lane = x_off + (uint32_t)((wc * NT) * 64 + g * 8 + 2 * t) * 8;

  const int64_t nsteps = (p.n + KS - 1) / KS;
  // entry generator: MODE 1 = every column index above every row index of the CTA (r = e_i / e_l),
  // 2 = every column below (r = e_l / e_i), 0 = mixed (diagonal tiles, last partial step)
  const double tt0 = -(p.rlo * p.inv_h) - 0.5;  // tt' = r * inv_h + tt0 = (r - rlo) / h - 1/2; segment = round(tt')
  const double MAGIC = 6755399441055744.0;      // 1.5 * 2^52: (tt' + MAGIC) holds round(tt') in its low mantissa bits
  int64_t l0 = 0;
  const double2* ctc = ct;
  // per-lane shared-memory offsets of the 4 distinct swizzle patterns (k16 & 7 = 2 (q & 3) + gpar)
  uint32_t sw_off[4];
#pragma unroll
  for (int q4 = 0; q4 < 4; ++q4)
    sw_off[q4] = (uint32_t)(((grow >> 1) ^ ((2 * q4 + gpar) & 7)) << 4) + (uint32_t)(grow & 1) * 8;
  auto gen_step = [&](auto mode_tag) {
    constexpr int MODE = decltype(mode_tag)::value;
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      const int kk = kk_base + 2 * q;              // kk_base in {0, 1, 16, 17}: stage and k16 follow q
      const double2 c = ctc[kk];
      const int64_t gl = l0 + kk;
#pragma unroll
      for (int s = 0; s < SPW; ++s) {
        const int st = SUBT >= 8 ? warp * SPW + s : warp / WPS;
        const int64_t gi = girow[s];
        // r = e_lo / e_hi with lo = the smaller INDEX (benchmark_free.f90:54-58)
        double r;
        if (MODE == 1) r = erow[s].x * c.y;
        else if (MODE == 2) r = c.x * erow[s].y;
        else r = (gi <= gl) ? erow[s].x * c.y : c.x * erow[s].y;
        const double tt = fma(r, p.inv_h, tt0);
        const double y = tt + MAGIC;
        int si = __double2loint(y);
        const double u = 2.0 * (tt - (y - MAGIC));
        si = min(FSEG - 1, max(0, si));
        const double2* cf = reinterpret_cast<const double2*>(coef_s + si * FCOEF);
        const double2 c01 = cf[0], c23 = cf[1], c45 = cf[2];
        double v = fma(c45.y, u, c45.x);
        v = fma(v, u, c23.y);
        v = fma(v, u, c23.x);
        v = fma(v, u, c01.y);
        v = fma(v, u, c01.x);  // the 1e-4 factor is folded into the coefficients
        if (MODE == 0) {
          if (gi == gl) v = is_stx ? 1.0 : v + (double)(float)(gi + 1);
          if (gl >= p.n) v = 0.0;
        }
        const int k16 = kk & 15;
        const uint32_t off = a_off + (uint32_t)(kk >> 4) * A_STAGE + (uint32_t)st * (16 * 128) +
                             (uint32_t)(k16 & ~1) * 128 + (uint32_t)gpar * 128 + sw_off[q & 3];
        *reinterpret_cast<double*>(gbase + off) = v;
      }
    }
  };
  __syncthreads();
  for (int64_t step = 0; step < nsteps; ++step) {
    l0 = step * KS;
    ctc = ct + (step & 1) * KS;
    // (a) generate the BM x 32 operator tile.  Steps entirely left / right of the CTA's rows (all but ~BM/32 of
    // them) need no index comparison, no diagonal term and no column bound: CTA-uniform three-way branch.
    const int64_t gr0 = p.row0 + i0, gr1 = gr0 + BM - 1;  // global rows of this CTA
    const int mode = (l0 + KS <= p.n && l0 > gr1) ? 1 : ((l0 + KS - 1 < gr0) ? 2 : 0);
    if (mode == 1) gen_step(std::integral_constant<int, 1>{});
    else if (mode == 2) gen_step(std::integral_constant<int, 2>{});
    else gen_step(std::integral_constant<int, 0>{});
    // (b) X tile of this step (already in fragment order) and the column table of the next step
    {
      const double2* src = reinterpret_cast<const double2*>(p.Xp + (size_t)step * (KS * BPAD));
      double2* dst = reinterpret_cast<double2*>(gbase + x_off);
#pragma unroll
      for (int e = tid; e < KS * BPAD / 2; e += 256) dst[e] = __ldg(src + e);
      if (tid < KS) {
        const int64_t gl = l0 + KS + tid;
        (ct + ((step + 1) & 1) * KS)[tid] = gl < p.n ? p.et[gl] : make_double2(1.0, 1.0);
      }
    }
    __syncthreads();
    // (c) consume: two 16-column stages
#pragma unroll
    for (int stage = 0; stage < 2; ++stage) {
      const uint32_t sa = base + (uint32_t)stage * A_STAGE, sx = base + (uint32_t)stage * X_STAGE;
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        double2 xf[NT];
#pragma unroll
        for (int nn = 0; nn < NT; ++nn) xf[nn] = lds128(sx + x_lane + (uint32_t)((q * NTT + nn) * 64) * 8);
#pragma unroll
        for (int rg = 0; rg < 2; ++rg) {
#pragma unroll
          for (int o = 0; o < 2; ++o) {
            const int kk = 8 * q + 2 * t + o;
            const double2 af =
                lds128(sa + a_lane + (uint32_t)rg * (16 * 128) + (uint32_t)kk * 128 + (uint32_t)((g ^ (kk & 7)) << 4));
#pragma unroll
            for (int nn = 0; nn < NT; ++nn) {
              const double xv = o ? xf[nn].y : xf[nn].x;
              dmma(acc[rg][0][nn][0], acc[rg][0][nn][1], af.x, xv);
              dmma(acc[rg][1][nn][0], acc[rg][1][nn][1], af.y, xv);
            }
          }
        }
      }
    }
    __syncthreads();
  }

  const int64_t row_base = i0 + wr * 32;
#pragma unroll
  for (int rg = 0; rg < 2; ++rg)
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int64_t row = row_base + rg * 16 + 2 * g + e;
      if (row < p.nl) {
#pragma unroll
        for (int nn = 0; nn < NT; ++nn) {
          const int j = (wc * NT + nn) * 8 + 2 * t;
          if (j < p.b) p.W[row + (int64_t)j * p.ldw] = acc[rg][e][nn][0];
          if (j + 1 < p.b) p.W[row + (int64_t)(j + 1) * p.ldw] = acc[rg][e][nn][1];
        }
      }
    }
}

template <int NT, int WARPS_N>
void launch_free(cudaStream_t s, FreeParams& p) {
  constexpr int BM = (8 / WARPS_N) * 32;
  constexpr int BPAD = NT * WARPS_N * 8;
  const size_t smem = 2 * (size_t)BM * 16 * 8 + 2 * (size_t)16 * BPAD * 8 + FSEG * FCOEF * 8 + 2 * KS * 16 + 1024;
  ensure_dyn_smem(free_dmma_kernel<NT, WARPS_N>, (int)smem);  // per (kernel, device)
  const unsigned grid = (unsigned)ceil_div(p.nl, BM);
  free_dmma_kernel<NT, WARPS_N><<<grid, 256, smem, s>>>(p);
  CK_LAUNCH();
  ++g_kernel_launches;
}

long double g_exact(long double r, bool use_sin) {
  const long double l = 0.5L * logl(atanl(r));
  return use_sin ? sinl(l) : cosl(l);
}

}  // namespace

struct FreeTables {
  int64_t n = 0;
  int op = -1;
  bool usable = false;
  double rlo = 0.0, inv_h = 0.0, max_err = 0.0;
  DevBuf<double2> et;
  DevBuf<double> coef;
  DevBuf<double> Xp;
};

FreeTables* free_tables_create(int op, int64_t n, const double* etab_host) {
  FreeTables* T = new FreeTables();
  T->n = n;
  T->op = op;
  const bool use_sin = op == DAV_OP_TEST_STX;
  std::vector<double2> et((size_t)n);
  for (int64_t t = 0; t < n; ++t) et[(size_t)t] = make_double2(etab_host[t], 1.0 / etab_host[t]);
  // fit interval: every ratio e_lo / e_hi of the table, with a little slack for the rounding of e_lo * (1 / e_hi)
  double emin = etab_host[0], emax = etab_host[0];
  for (int64_t t = 1; t < n; ++t) { emin = std::min(emin, etab_host[t]); emax = std::max(emax, etab_host[t]); }
  const long double rlo = (long double)(emin / emax) * (1.0L - 1e-6L), rhi = 1.0L + 1e-6L;
  const long double h = (rhi - rlo) / FSEG;
  std::vector<double> coef((size_t)FSEG * FCOEF);
  const long double PI = 3.14159265358979323846264338327950288L;
  for (int sgm = 0; sgm < FSEG; ++sgm) {
    const long double a = rlo + sgm * h;
    long double fv[FCOEF], th[FCOEF], c[FCOEF];
    for (int m = 0; m < FCOEF; ++m) {
      th[m] = PI * (m + 0.5L) / FCOEF;
      fv[m] = g_exact(a + (cosl(th[m]) + 1.0L) * h * 0.5L, use_sin);
    }
    for (int j = 0; j < FCOEF; ++j) {
      long double sum = 0.0L;
      for (int m = 0; m < FCOEF; ++m) sum += fv[m] * cosl(j * th[m]);
      c[j] = sum * 2.0L / FCOEF;
    }
    c[0] *= 0.5L;
    // Chebyshev -> monomial in u
    long double Tm[FCOEF][FCOEF] = {};
    Tm[0][0] = 1.0L;
    if (FCOEF > 1) Tm[1][1] = 1.0L;
    for (int j = 2; j < FCOEF; ++j)
      for (int q = 0; q < FCOEF; ++q) Tm[j][q] = (q > 0 ? 2.0L * Tm[j - 1][q - 1] : 0.0L) - Tm[j - 2][q];
    for (int q = 0; q < FCOEF; ++q) {
      long double mq = 0.0L;
      for (int j = 0; j < FCOEF; ++j) mq += c[j] * Tm[j][q];
      coef[(size_t)sgm * FCOEF + q] = (double)(mq * (long double)(double)1e-4f);  // x the reference's 1e-4 (single)
    }
  }
  T->rlo = (double)rlo;
  T->inv_h = (double)(1.0L / h);
  // verify against extended-precision libm on a dense sample (same arithmetic as the device evaluation)
  double worst = 0.0;
  const int samples = 20000;
  for (int sidx = 0; sidx <= samples; ++sidx) {
    const double r = (double)(rlo + (rhi - rlo) * ((long double)sidx / samples));
    const double tt = (r - T->rlo) * T->inv_h;
    int si = std::min(FSEG - 1, std::max(0, (int)tt));
    const double u = std::fma(2.0, tt - (double)si, -1.0);
    const double* cf = &coef[(size_t)si * FCOEF];
    double v = std::fma(cf[5], u, cf[4]);
    v = std::fma(v, u, cf[3]);
    v = std::fma(v, u, cf[2]);
    v = std::fma(v, u, cf[1]);
    v = std::fma(v, u, cf[0]);
    worst = std::max(worst, std::fabs(v / (double)1e-4f - (double)g_exact((long double)r, use_sin)));
  }
  T->max_err = worst;
  T->usable = worst < 4e-15 && emin > 0.0;
  if (T->usable) {
    T->et.alloc((size_t)n);
    CK(cudaMemcpy(T->et.p, et.data(), (size_t)n * sizeof(double2), cudaMemcpyHostToDevice));
    T->coef.alloc(coef.size());
    CK(cudaMemcpy(T->coef.p, coef.data(), coef.size() * 8, cudaMemcpyHostToDevice));
  }
  return T;
}

void free_tables_destroy(FreeTables* T) { delete T; }
bool free_tables_usable(const FreeTables* T) { return T && T->usable; }
double free_tables_max_err(const FreeTables* T) { return T ? T->max_err : -1.0; }

void free_matmul_dmma(cudaStream_t s, FreeTables* T, int64_t row0, int64_t nl, int b, const double* X, int64_t ldx,
                      double* W, int64_t ldw) {
  if (nl <= 0 || b <= 0) return;
  const int64_t n = T->n;
  const int64_t Kpad = round_up(n, KS);
  for (int j0 = 0; j0 < b; j0 += 128) {
    const int bc = std::min(128, b - j0);
    int bpad = (int)round_up(bc, 8);
    const int warps_n = bpad <= 32 ? 1 : (bpad <= 64 ? 2 : 4);
    const int nt = (bpad + 8 * warps_n - 1) / (8 * warps_n);
    bpad = nt * warps_n * 8;
    T->Xp.alloc((size_t)Kpad * bpad);
    {
      const int64_t total = Kpad * bpad;
      const int blocks = (int)std::min<int64_t>(ceil_div(total, 256), 1184);
      free_pack_x_kernel<<<blocks, 256, 0, s>>>(n, Kpad, bc, bpad, X + (int64_t)j0 * ldx, ldx, T->Xp.p);
      CK_LAUNCH();
      ++g_kernel_launches;
    }
    FreeParams p;
    p.op = T->op; p.n = n; p.row0 = row0; p.nl = nl; p.b = bc;
    p.et = T->et.p; p.coef = T->coef.p; p.rlo = T->rlo; p.inv_h = T->inv_h;
    p.Xp = T->Xp.p;
    p.W = W + (int64_t)j0 * ldw;
    p.ldw = ldw;
#define FCFG(NT_, WN_) launch_free<NT_, WN_>(s, p)
    if (warps_n == 1) {
      switch (nt) {
        case 1: FCFG(1, 1); break;
        case 2: FCFG(2, 1); break;
        case 3: FCFG(3, 1); break;
        default: FCFG(4, 1); break;
      }
    } else if (warps_n == 2) {
      if (nt == 3) FCFG(3, 2); else FCFG(4, 2);
    } else {
      if (nt == 3) FCFG(3, 4); else FCFG(4, 4);
    }
#undef FCFG
  }
}

}  // namespace dav
