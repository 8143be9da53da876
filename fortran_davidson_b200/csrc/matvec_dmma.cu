// The hot kernel: block matvec  W(M x b) = A(M x K, lda) * X(K x b)  in FP64, A streamed from HBM
// exactly once for all b columns.  Replaces the reference's DGEMM A*V / B*V (davidson.f90:131,134,
// 223,226 via lapack_wrapper.f90:279-328) and, through the stored products, its k DGEMVs per
// iteration (davidson.f90:163-170).
//
// sm_100a design
//   * persistent grid, one CTA per SM.  Blocks of up to 32 columns (HBM- or just FP64-bound): stream-K -- the
//     (row tile x k step) units are divided evenly and contiguously over the CTAs; partially covered tiles go to a
//     workspace and a tiny fixup kernel adds them in increasing-k order (bit-reproducible, no atomics).  Wider
//     blocks: full WAVES + a stream-K remainder -- while at least `grid` row tiles are left, CTA c takes tile
//     (wave*grid + c) and all CTAs sweep the k steps together, so one X tile serves the whole grid from L2 instead
//     of all of X (51 MB at n = 100,000, b = 64) staying live next to the A stream; same kernel time, DRAM reads
//     back to the algorithmic 80 GB (see schedule_for()).
//   * warp-specialised: warp 8 is the TMA producer, warps 0-7 consume (each 32 rows x up to 32 columns).  A tiles
//     arrive through a 2D tensor map (boxes of 16 rows x BK = 32 columns, 128-byte swizzle) with mbarrier
//     complete_tx; the X tile is pre-packed in fragment order so one 1D bulk copy per stage fetches it.  The slot
//     hand-back crosses memory proxies (ld.shared reads, async-proxy refill): block fence before the consumers'
//     arrive, fence.proxy.async before the producer's TMA issue.
//   * FP64 tensor-core math: mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4).  tcgen05 has no FP64 kind, so
//     TMEM / UTCMMA do not apply to this path.
//   * shared-memory fragment loads are 128-bit and bank-conflict free: the 128B swizzle puts rows
//     (2g, 2g+1) of column k at chunk g ^ (k & 7); a lane (g, t) reads column k0 + 2t + o, so the 8
//     lanes of a quarter warp hit 8 distinct chunks of 4 distinct lines.
//
// Algorithmic traffic per launch: 8*M*K (A) + 8*K*b (X) + 8*M*b (W) bytes; 2*M*K*b flops.
#include <cuda.h>

#include <algorithm>
#include <cstdlib>
#include <vector>

#include "kernels.cuh"

namespace dav {

namespace {

constexpr int BK_DEFAULT = 32;    // k columns of A per pipeline stage (DAV_MATVEC_BK=16 selects the shallower stage)
constexpr int CONSUMERS = 8;      // consumer warps
constexpr int THREADS = (CONSUMERS + 1) * 32;
constexpr int MAX_STAGES = 8;
constexpr int NUM_SMS_FALLBACK = 148;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE;\n"
      "bra WAIT_LOOP;\n"
      "DONE:\n"
      "}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}
// L2 policies: A is streamed exactly once (evict_first) while the packed X block is re-read by every row tile
// (evict_last) -- without the hints the A stream (~200 MB between two reads of an X line) pushes X out of the
// 126 MB L2 and X is re-fetched from HBM (ncu: +5.6 % DRAM reads at b = 16, +9 % at b = 32).
__device__ __forceinline__ uint64_t policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar,
                                            uint64_t pol) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3}], "
      "[%4], %5;" ::"r"(dst),
      "l"(map), "r"(c0), "r"(c1), "r"(bar), "l"(pol)
      : "memory");
}
__device__ __forceinline__ void bulk_load_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar,
                                             uint64_t pol) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
          dst),
      "l"(src), "r"(bytes), "r"(bar), "l"(pol)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_nohint(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
          dst),
      "l"(map), "r"(c0), "r"(c1), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void bulk_load_1d_nohint(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
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

// ---- the schedule: shared by the kernel, the fixup kernel and the host self-test (dav_debug_matvec_schedule) ----
#define DAV_HD __host__ __device__ __forceinline__
constexpr int WS_SLOTS = 4;  // workspace slots (partial tiles) per CTA
struct Sched {
  int tiles, ksteps;   // row tiles, k steps
  int grid;            // CTAs launched
  int waves;           // full waves: CTA c owns the tiles w*grid + c, w < waves, completely
  int tile_off;        // first tile of the remainder (= waves * grid)
  int rem_tiles;       // tiles - tile_off
  // remainder, variant A (split == 0), stream-K: the rem_tiles * ksteps units are cut into `grid` contiguous ranges
  long long total;     // rem_tiles * ksteps
  long long quota;     // units per CTA (0 when there is no remainder)
  // remainder, variant B (split >= 1), aligned split-K: every remainder tile is cut into `split` pieces of kchunk
  // k steps; piece v = j * rem_tiles + rt (j-major) goes to CTA v % grid as its (v / grid)-th remainder segment, so
  // the CTAs of one round sit on at most a few distinct k positions
  int split, kchunk;
};
// consecutive k steps [ks0, ks1) of one row tile; slot < 0: the whole tile (written to W directly), else the
// workspace slot of the CTA that receives the partial sums
struct Segment {
  int tile, ks0, ks1, slot;
};
struct SegCursor {
  int wave, round;
  long long u, u_begin, u_end;
};
DAV_HD SegCursor seg_begin(const Sched& sc, int cta) {
  SegCursor c;
  c.wave = 0;
  c.round = 0;
  const long long ub = (long long)cta * sc.quota;
  c.u_begin = ub < sc.total ? ub : sc.total;
  const long long ue = c.u_begin + sc.quota;
  c.u_end = ue < sc.total ? ue : sc.total;
  c.u = c.u_begin;
  return c;
}
DAV_HD bool seg_next(const Sched& sc, int cta, SegCursor& c, Segment& sg) {
  if (c.wave < sc.waves) {
    sg.tile = c.wave * sc.grid + cta;
    sg.ks0 = 0;
    sg.ks1 = sc.ksteps;
    sg.slot = -1;
    ++c.wave;
    return true;
  }
  if (sc.split > 0) {
    const long long v = (long long)c.round * sc.grid + cta;
    if (v >= (long long)sc.rem_tiles * sc.split) return false;
    const int j = (int)(v / sc.rem_tiles), rt = (int)(v - (long long)j * sc.rem_tiles);
    sg.tile = sc.tile_off + rt;
    sg.ks0 = j * sc.kchunk;
    sg.ks1 = sg.ks0 + sc.kchunk < sc.ksteps ? sg.ks0 + sc.kchunk : sc.ksteps;
    sg.slot = sc.split == 1 ? -1 : c.round;
    ++c.round;
    return true;
  }
  if (c.u >= c.u_end) return false;
  const int rt = (int)(c.u / sc.ksteps);
  sg.tile = sc.tile_off + rt;
  sg.ks0 = (int)(c.u - (long long)rt * sc.ksteps);
  const long long left = c.u_end - c.u;
  sg.ks1 = (long long)sg.ks0 + left < (long long)sc.ksteps ? (int)(sg.ks0 + left) : sc.ksteps;
  const bool complete = sg.ks0 == 0 && sg.ks1 == sc.ksteps;
  sg.slot = complete ? -1 : (c.u == c.u_begin ? 0 : 1);  // a CTA has at most a first and a last partial segment
  c.u += sg.ks1 - sg.ks0;
  return true;
}
// The fixup pass adds the fixup_count(rt) partial pieces of remainder tile rt in increasing-k order (0: the tile was
// written directly); piece i sits in workspace slot `slot` of CTA `cta`.
DAV_HD int fixup_count(const Sched& sc, int rt) {
  if (sc.split > 0) return sc.split == 1 ? 0 : sc.split;
  const long long u0 = (long long)rt * sc.ksteps, u1 = u0 + sc.ksteps;
  const int cA = (int)(u0 / sc.quota), cB = (int)((u1 - 1) / sc.quota);
  return cA == cB ? 0 : cB - cA + 1;
}
DAV_HD void fixup_piece(const Sched& sc, int rt, int i, int& cta, int& slot) {
  if (sc.split > 0) {
    const long long v = (long long)i * sc.rem_tiles + rt;
    cta = (int)(v % sc.grid);
    slot = (int)(v / sc.grid);
    return;
  }
  const long long u0 = (long long)rt * sc.ksteps;
  cta = (int)(u0 / sc.quota) + i;
  slot = ((long long)cta * sc.quota >= u0) ? 0 : 1;  // the CTA's first segment, or its last one
}
// schedule 0: everything stream-K.  1: full waves first (all CTAs on the same k step), the tiles that do not fill
// a wave as stream-K.  2: full waves, then the aligned split-K remainder when it keeps >= 90 % of the CTAs busy
// (else stream-K).  max_grid: SM count, already limited by the workspace (WS_SLOTS slots per CTA).
inline Sched make_sched(int tiles, int ksteps, int max_grid, int schedule) {
  Sched sc;
  sc.tiles = tiles;
  sc.ksteps = ksteps;
  const long long all_units = (long long)tiles * ksteps;
  int grid = (int)std::min<long long>(std::max(max_grid, 1), std::max<long long>(all_units, 1));
  sc.waves = (schedule != 0 && tiles >= grid) ? tiles / grid : 0;
  sc.tile_off = sc.waves * grid;
  sc.rem_tiles = tiles - sc.tile_off;
  sc.total = (long long)sc.rem_tiles * ksteps;
  sc.quota = sc.total > 0 ? (sc.total + grid - 1) / grid : 0;
  sc.split = 0;
  sc.kchunk = 0;
  if (schedule == 2 && sc.rem_tiles > 0) {
    const int R = sc.rem_tiles;
    int best_s = 0;
    double best_eff = 0.0;
    const long long smax = std::min<long long>(ksteps, (long long)WS_SLOTS * grid / R);
    for (int s = 1; s <= smax; ++s) {
      const int kc = (ksteps + s - 1) / s;
      if ((ksteps + kc - 1) / kc != s) continue;  // this s leaves an empty last piece
      const long long pieces = (long long)R * s, rounds = (pieces + grid - 1) / grid;
      // busy fraction of the rounds (every piece takes kc k steps, the useful work is R * ksteps)
      const double eff = (double)R * ksteps / ((double)rounds * grid * kc);
      if (eff > best_eff + 1e-12) { best_eff = eff; best_s = s; }
    }
    if (best_s > 0 && best_eff >= 0.9) {
      sc.split = best_s;
      sc.kchunk = (ksteps + best_s - 1) / best_s;
      sc.total = 0;  // no stream-K units
      sc.quota = 0;
    }
  }
  if (sc.waves == 0 && sc.split == 0 && sc.quota > 0) grid = (int)((sc.total + sc.quota - 1) / sc.quota);
  sc.grid = grid;
  return sc;
}

struct Params {
  int64_t M, K;        // rows of the local block, columns (= global n)
  int b;               // real column count (<= bpad)
  Sched sc;
  int stages;
  int l2_hints;        // bit 0: evict_first on the A stream, bit 1: evict_last on the packed X block
  const double* Xp;    // packed X: [kstep][BK x bpad] in fragment order
  double* W;
  int64_t ldw;
  double* ws;          // partial tiles: [cta][WS_SLOTS][BM x bpad]
};

// X packed index of element (k, j): ((k/8 * NTT + j/8) * 64 + (j%8)*8 + k%8)
__global__ void pack_x_kernel(int64_t K, int64_t Kpad, int b, int bpad, const double* __restrict__ X, int64_t ldx,
                              double* __restrict__ Xp) {
  const int ntt = bpad / 8;
  const int64_t total = Kpad * bpad;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    // e enumerates the packed layout so that stores are coalesced
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

template <int NT, int WARPS_N, int BK>
__global__ void __launch_bounds__(THREADS, 1)
    matvec_kernel(const __grid_constant__ CUtensorMap tmapA, const Params p) {
  constexpr int WARPS_M = CONSUMERS / WARPS_N;
  constexpr int BM = WARPS_M * 32;
  constexpr int NTT = NT * WARPS_N;          // 8-column tiles of the CTA tile
  constexpr int BPAD = NTT * 8;
  constexpr uint32_t A_BYTES = BM * BK * 8;  // per stage
  constexpr uint32_t X_BYTES = BK * BPAD * 8;
  constexpr uint32_t STAGE_BYTES = A_BYTES + X_BYTES;

  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[MAX_STAGES];
  __shared__ __align__(8) uint64_t empty_bar[MAX_STAGES];

  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int S = p.stages;

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&empty_bar[s]), CONSUMERS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const Sched& sc = p.sc;
  SegCursor cur = seg_begin(sc, (int)blockIdx.x);
  Segment sg;

  if (warp == CONSUMERS) {
    // ================= TMA producer (one elected lane) =================
    if (lane == 0) {
      int s = 0;        // ring position and phase bit, kept incrementally (no 64-bit division per stage)
      uint32_t ph = 0;
      const uint64_t pol_a = policy_evict_first(), pol_x = policy_evict_last();
      auto load_stage = [&](int tile, int ks) {
        mbar_wait(smem_u32(&empty_bar[s]), ph ^ 1u);
        // The consumers read this slot with ld.shared (generic proxy); the refill below writes it through the async
        // proxy.  Write-after-read across the two proxies needs a proxy fence between the acquire above and the TMA
        // issue -- without it the refill overtook late fragment loads once the producer stopped spending ~500
        // cycles per stage on 64-bit divisions (scripts/matvec_stress.py: 27 of 30 runs wrong at n=2048, b=128).
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        const uint32_t fb = smem_u32(&full_bar[s]);
        mbar_expect_tx(fb, STAGE_BYTES);
        const uint32_t a_dst = base + (uint32_t)s * STAGE_BYTES;
        if (p.l2_hints & 1) {
#pragma unroll
          for (int rg = 0; rg < BM / 16; ++rg)
            tma_load_2d(a_dst + rg * (BK * 128), &tmapA, tile * BM + rg * 16, ks * BK, fb, pol_a);
        } else {
#pragma unroll
          for (int rg = 0; rg < BM / 16; ++rg)
            tma_load_2d_nohint(a_dst + rg * (BK * 128), &tmapA, tile * BM + rg * 16, ks * BK, fb);
        }
        if (p.l2_hints & 2)
          bulk_load_1d(a_dst + A_BYTES, p.Xp + (size_t)ks * (BK * BPAD), X_BYTES, fb, pol_x);
        else
          bulk_load_1d_nohint(a_dst + A_BYTES, p.Xp + (size_t)ks * (BK * BPAD), X_BYTES, fb);
        if (++s == S) { s = 0; ph ^= 1u; }
      };
      while (seg_next(sc, (int)blockIdx.x, cur, sg))
        for (int ks = sg.ks0; ks < sg.ks1; ++ks) load_stage(sg.tile, ks);
    }
    return;
  }

  // ================= consumers =================
  const int wr = warp / WARPS_N, wc = warp % WARPS_N;
  const int g = lane >> 2, t = lane & 3;
  double acc[2][2][NT][2];  // [row group][even/odd row][col tile][c0,c1]

  // per-lane byte offsets inside a stage
  // A: sub-tile (wr*2 + rg) * (BK*128) + kk*128 + ((g ^ (kk&7)) << 4), kk = 8q + 2t + o
  // X: A_BYTES + ((q*NTT + wc*NT + nt)*64 + g*8 + 2t) * 8
  const uint32_t a_lane = (uint32_t)(wr * 2) * (BK * 128);
  const uint32_t x_lane = A_BYTES + (uint32_t)((wc * NT) * 64 + g * 8 + 2 * t) * 8;

  int s = 0;        // ring position and phase bit of the next stage to consume
  uint32_t ph = 0;
  while (seg_next(sc, (int)blockIdx.x, cur, sg)) {
    const int tile = sg.tile, ks0 = sg.ks0, ks1 = sg.ks1;
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int e = 0; e < 2; ++e)
#pragma unroll
        for (int n = 0; n < NT; ++n) acc[a][e][n][0] = acc[a][e][n][1] = 0.0;

    for (int ks = ks0; ks < ks1; ++ks) {
      mbar_wait(smem_u32(&full_bar[s]), ph);
      const uint32_t sb = base + (uint32_t)s * STAGE_BYTES;
#pragma unroll
      for (int q = 0; q < BK / 8; ++q) {
        double2 xf[NT];
#pragma unroll
        for (int n = 0; n < NT; ++n) xf[n] = lds128(sb + x_lane + (uint32_t)((q * NTT + n) * 64) * 8);
#pragma unroll
        for (int rg = 0; rg < 2; ++rg) {
#pragma unroll
          for (int o = 0; o < 2; ++o) {
            const int kk = 8 * q + 2 * t + o;
            const double2 af = lds128(sb + a_lane + (uint32_t)rg * (BK * 128) + (uint32_t)kk * 128 +
                                      (uint32_t)((g ^ (kk & 7)) << 4));
#pragma unroll
            for (int n = 0; n < NT; ++n) {
              const double xv = o ? xf[n].y : xf[n].x;
              dmma(acc[rg][0][n][0], acc[rg][0][n][1], af.x, xv);
              dmma(acc[rg][1][n][0], acc[rg][1][n][1], af.y, xv);
            }
          }
        }
      }
      // Release the slot only when every fragment load of this stage has returned: ptxas schedules the arrive
      // right behind the last LDS *issue* (it has no register dependency on the loaded data, the 30-odd DMMAs that
      // consume it come later), so the block fence -- which waits for the warp's outstanding shared-memory loads --
      // goes between them.  Measured cost: none (the other warp of the sub-partition owns the DMMA pipe meanwhile).
      __syncwarp();
      __threadfence_block();
      if (lane == 0) mbar_arrive(smem_u32(&empty_bar[s]));
      if (++s == S) { s = 0; ph ^= 1u; }
    }

    if (sg.slot < 0) {
      const int64_t row_base = (int64_t)tile * BM + wr * 32;
#pragma unroll
      for (int rg = 0; rg < 2; ++rg)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int64_t row = row_base + rg * 16 + 2 * g + e;
          if (row < p.M) {
#pragma unroll
            for (int n = 0; n < NT; ++n) {
              const int j = (wc * NT + n) * 8 + 2 * t;
              if (j < p.b) p.W[row + (int64_t)j * p.ldw] = acc[rg][e][n][0];
              if (j + 1 < p.b) p.W[row + (int64_t)(j + 1) * p.ldw] = acc[rg][e][n][1];
            }
          }
        }
    } else {
      // partial tile -> one of the CTA's workspace slots
      double* w = p.ws + ((size_t)blockIdx.x * WS_SLOTS + sg.slot) * (size_t)(BM * BPAD);
#pragma unroll
      for (int rg = 0; rg < 2; ++rg)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int r = wr * 32 + rg * 16 + 2 * g + e;
#pragma unroll
          for (int n = 0; n < NT; ++n) {
            const int j = (wc * NT + n) * 8 + 2 * t;
            w[r + j * BM] = acc[rg][e][n][0];
            w[r + (j + 1) * BM] = acc[rg][e][n][1];
          }
        }
    }
  }
}

// Adds the partial tiles of every row tile that was split over several CTAs, in CTA order.  grid = (remainder tiles,
// FIXUP_SPLIT): the pieces of a tile are looked up once per CTA (the lookup divides), every thread then adds its
// elements piece by piece (r02: was one CTA per tile with the lookup inside the element loop, 35-52 us per launch at
// 12,500 local rows).
constexpr int FIXUP_SPLIT = 4;
constexpr int FIXUP_MAX_PIECES = 64;
__global__ void __launch_bounds__(256) fixup_kernel(int BM, int BPAD, Params p) {
  __shared__ int piece_off[FIXUP_MAX_PIECES];
  const int rt = blockIdx.x;  // tile of the remainder
  const int tile = p.sc.tile_off + rt;
  const int count = fixup_count(p.sc, rt);
  if (count == 0) return;  // written directly
  const int elems = BM * BPAD;
  const int cached = min(count, FIXUP_MAX_PIECES);
  for (int i = threadIdx.x; i < cached; i += blockDim.x) {
    int cta, slot;
    fixup_piece(p.sc, rt, i, cta, slot);
    piece_off[i] = cta * WS_SLOTS + slot;
  }
  __syncthreads();
  for (int e = blockIdx.y * blockDim.x + threadIdx.x; e < elems; e += gridDim.y * blockDim.x) {
    const int r = e % BM, j = e / BM;
    const int64_t row = (int64_t)tile * BM + r;
    if (row >= p.M || j >= p.b) continue;
    double s = 0.0;
    for (int i = 0; i < cached; ++i) s += p.ws[(size_t)piece_off[i] * (size_t)elems + e];
    for (int i = cached; i < count; ++i) {
      int cta, slot;
      fixup_piece(p.sc, rt, i, cta, slot);
      s += p.ws[((size_t)cta * WS_SLOTS + slot) * (size_t)elems + e];
    }
    p.W[row + (int64_t)j * p.ldw] = s;
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)ptr;
    else
      (void)cudaGetLastError();
  }
  return fn;
}

template <int NT, int WARPS_N, int BK>
void launch_cfg(cudaStream_t s, const CUtensorMap& map, Params& p, int ksteps, int max_smem, int num_sms, double* ws,
                size_t ws_doubles, int schedule) {
  constexpr int BM = (CONSUMERS / WARPS_N) * 32;
  constexpr int BPAD = NT * WARPS_N * 8;
  constexpr int STAGE_BYTES = BM * BK * 8 + BK * BPAD * 8;
  const size_t slot = (size_t)BM * BPAD;
  const int max_grid = (int)std::min<size_t>((size_t)num_sms, std::max<size_t>(1, ws_doubles / (WS_SLOTS * slot)));
  p.sc = make_sched((int)ceil_div(p.M, BM), ksteps, max_grid, schedule);
  const int grid = p.sc.grid;
  const int rem_tiles = p.sc.rem_tiles;
  p.stages = std::min(MAX_STAGES, (max_smem - 1024 - 256) / STAGE_BYTES);
  if (p.stages < 2) DAV_THROW(DAV_ERR_CUDA, "not enough shared memory for the matvec pipeline");
  p.ws = ws;
  const size_t smem = (size_t)p.stages * STAGE_BYTES + 1024;
  ensure_dyn_smem(matvec_kernel<NT, WARPS_N, BK>, max_smem - 256);  // per (kernel, device)
  matvec_kernel<NT, WARPS_N, BK><<<grid, THREADS, smem, s>>>(map, p);
  CK_LAUNCH();
  ++g_kernel_launches;
  if (rem_tiles > 0) {
    fixup_kernel<<<dim3(rem_tiles, FIXUP_SPLIT), 256, 0, s>>>(BM, BPAD, p);
    CK_LAUNCH();
    ++g_kernel_launches;
  }
}

}  // namespace

// config: 8 consumer warps as (8/WN row warps) x (WN column warps); a warp owns 32 rows x NT*8 columns with
// NT <= 4 (the 9-warp CTA is granted at most 168 registers per thread)
//   bpad <= 32 : WN=1, BM=256   | bpad <= 64 : WN=2, BM=128   | bpad <= 128 : WN=4, BM=64
static void pick_cfg(int bc, int* warps_n, int* nt, int* bpad) {
  int bp = (int)round_up(bc, 8);
  *warps_n = bp <= 32 ? 1 : (bp <= 64 ? 2 : 4);
  *nt = (bp + 8 * *warps_n - 1) / (8 * *warps_n);
  if (*warps_n > 1 && *nt < 3) *nt = 3;  // instantiated tile shapes: NT 1..4 for WN=1, NT 3..4 for WN=2,4
  *bpad = *nt * *warps_n * 8;
}

static int bk_from_env() {
  const char* e = std::getenv("DAV_MATVEC_BK");
  const int v = e ? std::atoi(e) : BK_DEFAULT;
  return (v == 16 || v == 32) ? v : BK_DEFAULT;
}

// 0: pure stream-K, 1: full waves + stream-K remainder, 2: full waves + aligned split-K remainder.  DAV_MATVEC_SCHEDULE
// overrides (read per call so one process can compare them).  Default: waves for blocks wider than 32 columns, where
// the packed X block (n x b doubles, 51 MB at n = 100,000, b = 64) no longer stays in L2 next to the A stream when the
// 148 CTAs of stream-K sit at 148 different k positions -- ncu at n = 100,000: DRAM reads 105-110 GB -> 81.7 GB at b = 64,
// 142 GB -> 84.3 GB at b = 128 (algorithmic 80.1 / 80.2 GB), same kernel time; stream-K for the narrow, HBM-bound blocks
// (reads already 80.0 / 81.0 GB at b = 16 / 32).
static int schedule_for(int bpad) {
  const char* e = std::getenv("DAV_MATVEC_SCHEDULE");
  return e ? std::atoi(e) : (bpad > 32 ? 1 : 0);
}

// Host model of the schedule the kernels execute (same inline functions): enumerates the segments of every CTA and
// checks that each (row tile, k step) unit is covered exactly once, that a CTA uses each workspace slot at most
// once, and that the fixup kernel reads exactly the partial segments of each remainder tile.  No device needed.
// info[10] = {grid, waves, tile_off, quota, tiles, ksteps, partial segments, BM, split, kchunk}.  0 = consistent.
int matvec_schedule_selftest(int64_t M, int64_t K, int b, int num_sms, int schedule, long long* info) {
  if (M <= 0 || K <= 0 || b <= 0 || b > 128 || num_sms <= 0) return -1;
  int warps_n, nt, bpad;
  pick_cfg(b, &warps_n, &nt, &bpad);
  const int BM = (CONSUMERS / warps_n) * 32;
  const int tiles = (int)ceil_div(M, (int64_t)BM);
  const int BK = bk_from_env();
  const int ksteps = (int)(round_up(K, BK) / BK);
  if (schedule < 0) schedule = schedule_for(bpad);
  const Sched sc = make_sched(tiles, ksteps, num_sms, schedule);
  if (sc.grid < 1 || sc.grid > num_sms) return 1;
  if ((long long)tiles * ksteps > (1LL << 28)) return -2;  // the coverage map below is meant for test sizes
  std::vector<unsigned char> cover((size_t)tiles * ksteps, 0);
  // partial pieces per remainder tile: (cta, slot, ks0, ks1)
  struct Piece { int cta, slot, ks0, ks1; };
  std::vector<std::vector<Piece>> pieces((size_t)sc.rem_tiles);
  long long npartial = 0;
  for (int c = 0; c < sc.grid; ++c) {
    SegCursor cur = seg_begin(sc, c);
    Segment sg;
    int slot_used[WS_SLOTS] = {0, 0, 0, 0};
    while (seg_next(sc, c, cur, sg)) {
      if (sg.tile < 0 || sg.tile >= tiles || sg.ks0 < 0 || sg.ks1 > ksteps || sg.ks0 >= sg.ks1) return 2;
      for (int ks = sg.ks0; ks < sg.ks1; ++ks)
        if (cover[(size_t)sg.tile * ksteps + ks]++) return 3;  // covered twice
      const bool complete = sg.ks0 == 0 && sg.ks1 == ksteps;
      if (complete != (sg.slot < 0)) return 4;
      if (!complete) {
        if (sg.tile < sc.tile_off) return 5;  // a wave tile must be complete
        if (sg.slot >= WS_SLOTS || slot_used[sg.slot]++) return 6;
        pieces[(size_t)(sg.tile - sc.tile_off)].push_back(Piece{c, sg.slot, sg.ks0, sg.ks1});
        ++npartial;
      }
    }
  }
  for (unsigned char v : cover)
    if (v != 1) return 7;  // not covered
  for (int rt = 0; rt < sc.rem_tiles; ++rt) {
    std::vector<Piece>& pc = pieces[(size_t)rt];
    std::sort(pc.begin(), pc.end(), [](const Piece& x, const Piece& y) { return x.ks0 < y.ks0; });
    const int count = fixup_count(sc, rt);
    if (count == 0) {
      if (!pc.empty()) return 8;  // fixup would skip a tile that has partial pieces
      continue;
    }
    if ((int)pc.size() != count) return 9;
    int ks = 0;
    for (int i = 0; i < count; ++i) {  // the fixup pass must read exactly these pieces, in increasing-k order
      int cta, slot;
      fixup_piece(sc, rt, i, cta, slot);
      const Piece& q = pc[(size_t)i];
      if (q.cta != cta || q.slot != slot) return 10;
      if (q.ks0 != ks) return 11;
      ks = q.ks1;
    }
    if (ks != ksteps) return 12;
  }
  if (info) {
    info[0] = sc.grid; info[1] = sc.waves; info[2] = sc.tile_off; info[3] = sc.quota;
    info[4] = tiles; info[5] = ksteps; info[6] = npartial; info[7] = BM; info[8] = sc.split; info[9] = sc.kchunk;
  }
  return 0;
}

struct MatvecPlan {
  CUtensorMap map, map32;  // boxes of 16 rows x 16 / 32 columns
  const double* A;
  int64_t M, K, lda;
  int max_b;
  DevBuf<double> Xp, ws;
  int max_smem, num_sms;
};

bool matvec_dmma_supported() { return get_encode_fn() != nullptr; }

MatvecPlan* matvec_plan_create(const double* A, int64_t M, int64_t K, int64_t lda, int max_b) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) DAV_THROW(DAV_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
  if ((lda * 8) % 16 != 0 || ((uintptr_t)A & 15) != 0)
    DAV_THROW(DAV_ERR_INVALID, "matvec: matrix block must be 16-byte aligned with an even leading dimension");
  MatvecPlan* p = new MatvecPlan();
  p->A = A; p->M = M; p->K = K; p->lda = lda; p->max_b = max_b;
  int dev = 0;
  CK(cudaGetDevice(&dev));
  CK(cudaDeviceGetAttribute(&p->max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  CK(cudaDeviceGetAttribute(&p->num_sms, cudaDevAttrMultiProcessorCount, dev));
  if (p->num_sms <= 0) p->num_sms = NUM_SMS_FALLBACK;
  cuuint64_t gdim[2] = {(cuuint64_t)M, (cuuint64_t)K};
  cuuint64_t gstride[1] = {(cuuint64_t)lda * 8};
  cuuint32_t estr[2] = {1, 1};
  for (int v = 0; v < 2; ++v) {
    cuuint32_t box[2] = {16, (cuuint32_t)(v ? 32 : 16)};
    CUresult r = enc(v ? &p->map32 : &p->map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void*)A, gdim, gstride, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      delete p;
      DAV_THROW(DAV_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    }
  }
  const int bmax = (int)std::min<int64_t>(round_up(std::max(max_b, 8), 8), 128);
  p->Xp.alloc((size_t)round_up(K, 32) * bmax);
  p->ws.alloc((size_t)p->num_sms * WS_SLOTS * 256 * 32);  // BM x BPAD <= 8192 for every config
  return p;
}

void matvec_plan_destroy(MatvecPlan* p) { delete p; }

// one chunk of bc <= 128 columns whose packed X block (Kpad x bpad, fragment order) is at Xp
static void launch_chunk(cudaStream_t s, MatvecPlan* plan, int BK, int64_t Kpad, int bc, const double* Xp, double* W,
                         int64_t ldw) {
  int warps_n, nt, bpad;
  pick_cfg(bc, &warps_n, &nt, &bpad);
  Params p;
  p.M = plan->M; p.K = plan->K; p.b = bc;
  const int ksteps = (int)(Kpad / BK);
  p.Xp = Xp;
  p.W = W;
  p.ldw = ldw;
  {
    static const int hints = [] { const char* e = std::getenv("DAV_MATVEC_L2_HINTS"); return e ? std::atoi(e) : 2; }();
    p.l2_hints = hints;
  }
  const int schedule = schedule_for(bpad);
#define CFG(NT_, WN_)                                                                                          \
  do {                                                                                                         \
    if (BK == 32)                                                                                              \
      launch_cfg<NT_, WN_, 32>(s, plan->map32, p, ksteps, plan->max_smem, plan->num_sms, plan->ws.p, plan->ws.n, \
                               schedule);                                                                      \
    else                                                                                                       \
      launch_cfg<NT_, WN_, 16>(s, plan->map, p, ksteps, plan->max_smem, plan->num_sms, plan->ws.p, plan->ws.n,   \
                               schedule);                                                                      \
  } while (0)
  if (warps_n == 1) {
    switch (nt) {
      case 1: CFG(1, 1); break;
      case 2: CFG(2, 1); break;
      case 3: CFG(3, 1); break;
      default: CFG(4, 1); break;
    }
  } else if (warps_n == 2) {
    switch (nt) {
      case 3: CFG(3, 2); break;
      default: CFG(4, 2); break;
    }
  } else {
    switch (nt) {
      case 3: CFG(3, 4); break;
      default: CFG(4, 4); break;
    }
  }
#undef CFG
}

void matvec_dmma(cudaStream_t s, MatvecPlan* plan, int b, const double* X, int64_t ldx, double* W, int64_t ldw) {
  if (b <= 0 || plan->M <= 0) return;
  const int BK = bk_from_env();
  const int64_t Kpad = round_up(plan->K, BK);
  for (int j0 = 0; j0 < b; j0 += 128) {
    const int bc = std::min(128, b - j0);
    int warps_n, nt, bpad;
    pick_cfg(bc, &warps_n, &nt, &bpad);
    if ((size_t)Kpad * bpad > plan->Xp.n) plan->Xp.alloc((size_t)Kpad * bpad);
    {
      const int64_t total = Kpad * bpad;
      const int blocks = (int)std::min<int64_t>(ceil_div(total, 256), 1184);
      pack_x_kernel<<<blocks, 256, 0, s>>>(plan->K, Kpad, bc, bpad, X + (int64_t)j0 * ldx, ldx, plan->Xp.p);
      CK_LAUNCH();
      ++g_kernel_launches;
    }
    launch_chunk(s, plan, BK, Kpad, bc, plan->Xp.p, W + (int64_t)j0 * ldw, ldw);
  }
}

int64_t matvec_kpad(int64_t K) { return round_up(K, bk_from_env()); }

size_t matvec_packed_doubles(int64_t K, int b) {
  if (b <= 0) return 0;
  const int nchunks = (b + 127) / 128;
  int warps_n, nt, bpad;
  pick_cfg(b - (nchunks - 1) * 128, &warps_n, &nt, &bpad);
  return ((size_t)(nchunks - 1) * 128 + (size_t)bpad) * (size_t)matvec_kpad(K);
}

// X already packed (by Comm::gather_rows_packed: every rank stored its rows into this rank's copy): chunk c of 128
// columns starts at c * 128 * Kpad
void matvec_dmma_packed(cudaStream_t s, MatvecPlan* plan, int b, const double* Xpacked, double* W, int64_t ldw) {
  if (b <= 0 || plan->M <= 0) return;
  const int BK = bk_from_env();
  const int64_t Kpad = round_up(plan->K, BK);
  for (int j0 = 0; j0 < b; j0 += 128)
    launch_chunk(s, plan, BK, Kpad, std::min(128, b - j0), Xpacked + (size_t)j0 * (size_t)Kpad, W + (int64_t)j0 * ldw,
                 ldw);
}

}  // namespace dav
