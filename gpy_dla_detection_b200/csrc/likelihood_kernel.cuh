// likelihood_kernel.cuh : batched low-rank Gaussian log-likelihoods of absorber samples
// (SURVEY.md §8 a5, a7, a8).
//
// Reference: for every QMC sample the Python loop builds dla_mu = mu a, dla_M = M a,
// d = omega2 a^2 + v (dla_gp.py:388-394) and calls NullGP.log_mvnpdf_low_rank
// (null_gp.py:307-360): B = I + M'D^-1 M (dgemm), chol(B), quad and log-det.
//
// Here all samples of a spectrum share the interpolated model, so with
//     w_sp = a_sp^2 / d_sp ,  g_sp = a_sp r_sp / d_sp ,  r_sp = y_p - mu_p a_sp
// the per-sample Gram matrices and projections are ONE dense FP64 contraction over pixels
//     [B_s - I | c_s] = [W | G] x [P | M] ,   P[p,(i,j)] = m_pi m_pj  (i >= j, 210 pairs)
// on the FP64 tensor path (DMMA m8n8k4; on B200 it shares the 64 FMA lanes/SM/clk of the DFMA
// pipe - 37 TFLOP/s measured either way - so every scalar FP64 instruction in this kernel is
// paid for with DMMA slots).  Structure, per CTA of 32 samples x 8 warps (two CTAs per SM):
//   * the Gram basis [P | M] of the spectrum (n x 240, precomputed once by gram_basis_kernel,
//     L2-resident) streams through shared memory in 16-pixel panels by TMA bulk copies
//     (cp.async.bulk + mbarrier complete_tx): no DMUL and no column-index registers in the MMA loop;
//   * every warp is producer AND consumer (warps that only produce starve behind warps that only
//     issue DMMAs - measured): it turns its 2 x 16 share of the profile-cache rows (product of up to
//     8 absorber factors, dla_gp.py:370-386, prefetched one panel ahead in registers) into the W / G
//     operand tiles of the NEXT panel - MUFU-seeded Newton reciprocal, integer-renormalised running
//     product for sum log d (one log per lane per tile) - then runs its DMMAs on the current one;
//   * warp (rh, cq) owns 16 samples x 7 or 8 of the 30 column blocks, sized [8,7,8,7] / [7,8,7,8]
//     so that the two warps of every SM sub-partition issue 30 DMMAs per 4 pixels: no padding
//     block, four balanced FP64 pipes;
//   * no CTA-wide barrier in the main loop: W/G tiles go through a 4-deep ring (mbarrier with 8 warp
//     arrivals per stage), basis panels through a 2-deep ring whose refill is requested by the LAST
//     warp to leave the stage (shared-memory counter), so nobody ever waits for a stage to drain and
//     warps drift by up to a panel;
//   * the bordered 21 x 21 Cholesky (factor, z = L^-1 c, log-det) of every sample runs in shared
//     memory straight out of the accumulator fragments while the SM's other CTA keeps the
//     tensor pipe busy.
// HBM sees only the profile rows (read) and one double per sample (written).
#pragma once
#include <stdint.h>
#include <type_traits>

namespace dla {

constexpr int LK_K = 20;                         // rank of the learned covariance (Parameters.k)
constexpr int LK_PAIRS = LK_K * (LK_K + 1) / 2;  // 210 lower-triangle pairs
constexpr int LK_TS = 32;                        // samples per CTA tile
constexpr int LK_KC = 16;                        // pixels per panel
constexpr int LK_WSTRIDE = LK_KC + 4;            // row stride of the W/G tiles (20 == 4 mod 16: conflict-free A loads)
constexpr int LK_NBLK_PAIR = 27;                 // ceil(210 / 8) column blocks of the Gram part
constexpr int LK_NBLK = 30;                      // + 3 column blocks (24 >= 20) of the projection part
constexpr int LK_NCOLS = LK_NBLK * 8;            // 240
constexpr int LK_PSTRIDE = LK_NCOLS + 4;         // basis row stride (244 == 4 mod 16: conflict-free B loads)
constexpr int LK_PROJ_COL0 = LK_NBLK_PAIR * 8;   // 216: first projection column
constexpr int LK_WARPS = 8;
constexpr int LK_THREADS = LK_WARPS * 32;        // 256
constexpr int LK_PSTAGES = 2;                    // basis-panel ring (TMA)
constexpr int LK_WSTAGES = 4;                    // W/G operand ring
constexpr int LK_EP_STRIDE = LK_TS + 1;          // epilogue smem: [col][sample], stride 33
constexpr int LK_MAX_ROWS = 8;                   // max absorbers multiplied per sample (max_dlas <= 8)
constexpr int LK_MB = 2;                         // DMMA row blocks per warp (16 samples)
constexpr int LK_NB_MAX = 8;                     // DMMA column blocks per warp (7 or 8)
constexpr int LK_EPT = LK_TS * LK_KC / LK_THREADS;  // W/G elements per thread per panel (2)
constexpr double LK_LOG_2PI = 1.83787706640934534;  // null_gp.py:325
constexpr double LK_LN2 = 0.693147180559945309417232121458;

// pair index c -> (i, j), i >= j, row-major lower triangle; pads map to (0,0)
__device__ __constant__ uint8_t c_pair_i[LK_NBLK_PAIR * 8];
__device__ __constant__ uint8_t c_pair_j[LK_NBLK_PAIR * 8];

// One spectrum as the likelihood kernel sees it.
// Factor r of sample s is a row of a profile matrix:
//   r == 0 : base0[ row(0,s) * ld ]   (profile cache, or the running-product buffer of the previous level)
//   r >= 1 : cache[ row(r,s) * ld ]
//   row(0,s) = rows0 ? rows0[s] : row0 + s
//   row(r,s) = rows  ? rows[(r-1) * row_stride + s] : row0 + r * row_stride + s
// The absorption is the left-to-right product of the factors (dla_gp.py:370-386).  When
// prod_out is set the product is stored as row s of prod_out, so the next level of
// DLAGP.log_model_evidences reads two rows per sample instead of level+1.
struct LikelihoodSpectrum {
  const double* y;       // n   normalised flux of the modelled pixels
  const double* v;       // n   noise variance
  const double* mu;      // n   this_mu  (mean-flux suppressed)
  const double* omega2;  // n   this_omega2
  const double* M;       // n x 20 row-major this_M
  const double* P;       // Gram basis [ceil(n/16)*16][244]: 210 pair products, pad, 20 columns of M, pad; zero rows >= n
  const double* base0;   // profile rows of factor 0, stride ld
  const double* cache;   // profile rows of factors >= 1, stride ld
  const int32_t* rows0;  // factor-0 row per sample, or nullptr
  const int32_t* rows;   // factor >= 1 rows, or nullptr
  const int* alive;      // if non-null and *alive == 0 the spectrum left the level loop (NaN evidence): skip
  double* prod_out;      // optional: product rows out (may alias base0 rows of the same sample)
  double* out;           // num_samples raw log-likelihoods
  int n;                 // modelled pixels
  int ld;                // profile row stride
  int num_samples;       // samples in this launch
  int num_rows;          // factors per sample (1..LK_MAX_ROWS)
  int row_stride;        // stride between factor arrays in `rows`
  int row0;              // first profile row when the row arrays are null
};

// ---- Gram basis of one spectrum: P[p][c] = m_p,i(c) * m_p,j(c), then the columns of M ---------------
struct GramBasisTask {
  const double* M;  // n x 20
  double* P;        // [ceil(n/16)*16][LK_PSTRIDE]
  int n;
};
// grid = (ceil(max_rows / 8), num_spectra), block = 256 (8 pixel rows x 32 lanes)
__global__ void __launch_bounds__(256) gram_basis_kernel(const GramBasisTask* __restrict__ tasks) {
  const GramBasisTask t = tasks[blockIdx.y];
  const int rows = (t.n + LK_KC - 1) / LK_KC * LK_KC;
  const int p = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (p >= rows) return;
  const int lane = threadIdx.x & 31;
  const double* m = t.M + (size_t)p * LK_K;
  double* out = t.P + (size_t)p * LK_PSTRIDE;
  for (int c = lane; c < LK_PSTRIDE; c += 32) {
    double val = 0.0;
    if (p < t.n) {
      if (c < LK_PAIRS) val = __dmul_rn(m[c_pair_i[c]], m[c_pair_j[c]]);
      else if (c >= LK_PROJ_COL0 && c < LK_PROJ_COL0 + LK_K) val = m[c - LK_PROJ_COL0];
    }
    out[c] = val;
  }
}

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
      : "+d"(c0), "+d"(c1)
      : "d"(a), "d"(b));
}

// ---- mbarrier + TMA bulk copy -----------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// 1/d for the operand tiles: MUFU.RCP64H seed (SFU, not the FP64 pipe) + two Newton steps (4 DFMA);
// <= 1 ulp from the IEEE quotient on the normal range, IEEE division outside it.
__device__ __forceinline__ double fast_rcp(double d) {
  if (!(d > 1e-290 && d < 1e290)) return 1.0 / d;
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
  double e = fma(-d, r, 1.0);
  r = fma(r, e, r);
  e = fma(-d, r, 1.0);
  return fma(r, e, r);
}

// ---- shared memory plan (per CTA; two CTAs are resident per SM) -----------------------------------
//   basis ring : 2 x [16][244]                                              62 464 B
//   W/G ring   : 4 x { W [32][20] | G [32][20] }                            40 960 B
//   E          : epilogue matrix [240][33], overlays the rings              63 360 B
//   AUX        : per-sample sums [32][2], profile rows [8][32], mbarriers, drain counters
constexpr int LK_PANEL_DOUBLES = LK_KC * LK_PSTRIDE;               // 3904
constexpr int LK_WG_DOUBLES = LK_TS * LK_WSTRIDE;                  // 640
constexpr uint32_t LK_PANEL_BYTES = LK_PANEL_DOUBLES * sizeof(double);  // 31 232
constexpr int LK_RING_DOUBLES = LK_PSTAGES * LK_PANEL_DOUBLES + LK_WSTAGES * 2 * LK_WG_DOUBLES;  // 12 928
constexpr int LK_EP_DOUBLES = LK_NCOLS * LK_EP_STRIDE;             // 7920
constexpr int LK_MAIN_DOUBLES = LK_RING_DOUBLES > LK_EP_DOUBLES ? LK_RING_DOUBLES : LK_EP_DOUBLES;
constexpr size_t LK_AUX_BYTES = LK_TS * 2 * sizeof(double) + LK_MAX_ROWS * LK_TS * sizeof(int32_t) +
                                (LK_PSTAGES + LK_WSTAGES) * sizeof(uint64_t) + LK_PSTAGES * sizeof(int);
constexpr size_t LK_SMEM_BYTES = (size_t)LK_MAIN_DOUBLES * sizeof(double) + LK_AUX_BYTES;
static_assert(LK_PANEL_BYTES % 128 == 0, "TMA alignment");

// grid = (ceil(max num_samples / 32), num_spectra), block = 256, dynamic smem = LK_SMEM_BYTES
__global__ void __launch_bounds__(LK_THREADS, 2)
sample_likelihood_kernel(const LikelihoodSpectrum* __restrict__ specs) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const LikelihoodSpectrum sp = specs[blockIdx.y];
  const int tile_s0 = blockIdx.x * LK_TS;
  if (tile_s0 >= sp.num_samples) return;
  if (sp.alive && *sp.alive == 0) return;

  double* s_main = reinterpret_cast<double*>(smem_raw);
  double* s_sums = s_main + LK_MAIN_DOUBLES;                          // [32][2] : sum r^2/d, sum log d
  int32_t* s_rows = reinterpret_cast<int32_t*>(s_sums + LK_TS * 2);   // [num_rows][32]
  uint64_t* s_mbar = reinterpret_cast<uint64_t*>(s_rows + LK_MAX_ROWS * LK_TS);  // panel[2], wg[4]
  int* s_drain = reinterpret_cast<int*>(s_mbar + LK_PSTAGES + LK_WSTAGES);       // [2] warps done with a basis stage

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int n = sp.n;
  const int npanels = (n + LK_KC - 1) / LK_KC;
  const int num_rows = sp.num_rows;
  const uint32_t bar_base = (uint32_t)__cvta_generic_to_shared(s_mbar);
  auto panel_bar = [&](int stage) { return bar_base + 8u * stage; };             // basis panel landed (TMA bytes)
  auto wg_bar = [&](int stage) { return bar_base + 8u * (LK_PSTAGES + stage); };   // W/G written by the 8 warps
  double* s_wg = s_main + LK_PSTAGES * LK_PANEL_DOUBLES;

  // ---- profile rows of the tile's samples, barriers -------------------------------------------
  for (int e = tid; e < num_rows * LK_TS; e += LK_THREADS) {
    const int r = e / LK_TS, s = e % LK_TS;
    const int gs = tile_s0 + s;
    int row = 0;
    if (gs < sp.num_samples) {
      if (r == 0) row = sp.rows0 ? sp.rows0[gs] : sp.row0 + gs;
      else row = sp.rows ? sp.rows[(size_t)(r - 1) * sp.row_stride + gs] : sp.row0 + r * sp.row_stride + gs;
    }
    s_rows[r * LK_TS + s] = row;
  }
  // basis panel `panel` -> its ring stage (one thread; completion is the stage's mbarrier)
  auto issue_panel = [&](int panel) {
    const int stage = panel & (LK_PSTAGES - 1);
    mbar_expect_tx(panel_bar(stage), LK_PANEL_BYTES);
    tma_bulk_g2s((uint32_t)__cvta_generic_to_shared(s_main + stage * LK_PANEL_DOUBLES),
                 sp.P + (size_t)panel * LK_PANEL_DOUBLES, LK_PANEL_BYTES, panel_bar(stage));
  };
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < LK_PSTAGES; ++s) { mbar_init(panel_bar(s), 1); s_drain[s] = 0; }
#pragma unroll
    for (int s = 0; s < LK_WSTAGES; ++s) mbar_init(wg_bar(s), LK_WARPS);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    for (int c = 0; c < LK_PSTAGES && c < npanels; ++c) issue_panel(c);
  }
  __syncthreads();

  // ---- producer role: thread -> pixel lane pl of the panel, samples sgrp and sgrp + 16 ------------
  // (a half-warp stores 16 consecutive doubles of one W/G row: conflict-free; and reads 128
  //  contiguous bytes of a profile row)
  const int pl = tid & (LK_KC - 1);
  const int sgrp = tid >> 4;  // 0..15
  const bool two_rows = num_rows > 1;
  bool live[LK_EPT];
  const double* rowp0[LK_EPT];
  const double* rowp1[LK_EPT];
#pragma unroll
  for (int e = 0; e < LK_EPT; ++e) {
    const int s = sgrp + 16 * e;
    live[e] = tile_s0 + s < sp.num_samples;
    rowp0[e] = sp.base0 + (size_t)s_rows[s] * sp.ld;
    rowp1[e] = two_rows ? sp.cache + (size_t)s_rows[LK_TS + s] * sp.ld : rowp0[e];
  }
  double q_acc[LK_EPT], dprod[LK_EPT];
  int esum[LK_EPT];
#pragma unroll
  for (int e = 0; e < LK_EPT; ++e) { q_acc[e] = 0.0; dprod[e] = 1.0; esum[e] = 0; }

  // pull the binary exponent out of the running product of d (integer pipe): keeps it in range
  // for any panel count; non-finite / non-positive values are left alone so they still poison the log
  auto renorm = [&]() {
#pragma unroll
    for (int e = 0; e < LK_EPT; ++e) {
      const int hi = __double2hiint(dprod[e]);
      const int ex = (hi >> 20) & 0x7ff;
      if (hi > 0 && ex != 0 && ex != 0x7ff) {
        esum[e] += ex - 1023;
        dprod[e] = __hiloint2double(hi - ((ex - 1023) << 20), __double2loint(dprod[e]));
      }
    }
  };

  // register buffer of the next panel's inputs (loads issued one panel ahead).  The loads are
  // unconditional on clamped addresses - a predicated load drags a dependent select behind it and
  // the warp would sit on the DRAM latency instead of running its DMMAs; validity is applied at use.
  double f0[LK_EPT], f1[LK_EPT], yp, mup, omp, vp;
  auto load_panel = [&](int panel) {
    const int p = min(panel * LK_KC + pl, n - 1);
#pragma unroll
    for (int e = 0; e < LK_EPT; ++e) {
      f0[e] = __ldg(rowp0[e] + p);
      f1[e] = __ldg(rowp1[e] + p);
    }
    yp = __ldg(sp.y + p);
    mup = __ldg(sp.mu + p);
    omp = __ldg(sp.omega2 + p);
    vp = __ldg(sp.v + p);
  };

  // W/G tiles of `panel` into stage panel % 4.  No wait: a warp can only be here after the basis
  // panel of `panel - 1` landed, which was requested after ALL warps finished panel - 3 (the previous
  // user of this W/G stage).
  auto produce = [&](int panel) {
    double* Ws = s_wg + (panel & (LK_WSTAGES - 1)) * (2 * LK_WG_DOUBLES);
    double* Gs = Ws + LK_WG_DOUBLES;
    const int p = panel * LK_KC + pl;
    const bool pv = p < n;
    double w[LK_EPT], g[LK_EPT];
#pragma unroll
    for (int e = 0; e < LK_EPT; ++e) {
      const int s = sgrp + 16 * e;
      w[e] = 0.0;
      g[e] = 0.0;
      if (pv && live[e]) {
        // absorption = product of the factors' profiles, left to right (dla_gp.py:370-386)
        double a = f0[e];
        if (two_rows) a = a * f1[e];
        for (int r = 2; r < num_rows; ++r) a = a * sp.cache[(size_t)s_rows[r * LK_TS + s] * sp.ld + p];
        if (sp.prod_out) sp.prod_out[(size_t)(tile_s0 + s) * sp.ld + p] = a;
        const double a2 = a * a;
        const double d = fma(omp, a2, vp);   // dla_omega2 + v
        const double inv = fast_rcp(d);
        const double r = fma(-mup, a, yp);   // y - dla_mu
        const double t = r * inv;
        w[e] = a2 * inv;
        g[e] = a * t;
        q_acc[e] = fma(r, t, q_acc[e]);
        dprod[e] *= d;
      }
    }
    if (panel + 1 < npanels) load_panel(panel + 1);  // in flight while this warp runs its DMMAs
#pragma unroll
    for (int e = 0; e < LK_EPT; ++e) {
      Ws[(sgrp + 16 * e) * LK_WSTRIDE + pl] = w[e];
      Gs[(sgrp + 16 * e) * LK_WSTRIDE + pl] = g[e];
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(wg_bar(panel & (LK_WSTAGES - 1)));
  };

  // ---- MMA role: warp (rh, cq) owns samples 16 rh .. 16 rh + 15 and 7 or 8 column blocks -----------
  //   rh = 0 : blocks [0,8) [8,15) [15,23) [23,30)      rh = 1 : blocks [0,7) [7,15) [15,22) [22,30)
  // Blocks >= 27 are the projection part (A = G).  Local blocks 0..3 always take A = W, local block 4
  // and local blocks 5..7 read A through pointers (W or G), so all warps run one instruction stream
  // with no predicated DMMA; the eighth block sits behind a warp-uniform branch.
  const int grp = lane >> 2;      // DMMA groupID
  const int tig = lane & 3;       // DMMA threadID_in_group
  const int rh = warp >> 2, cq = warp & 3;
  const int first = rh == 0 ? (cq == 0 ? 0 : cq == 1 ? 8 : cq == 2 ? 15 : 23) : (cq == 0 ? 0 : cq == 1 ? 7 : cq == 2 ? 15 : 22);
  const bool has_eighth = ((cq + rh) & 1) == 0;
  const bool g4 = first + 4 >= LK_NBLK_PAIR, g5 = first + 5 >= LK_NBLK_PAIR;
  double acc[LK_MB][LK_NB_MAX][2];
#pragma unroll
  for (int m = 0; m < LK_MB; ++m)
#pragma unroll
    for (int nb = 0; nb < LK_NB_MAX; ++nb) acc[m][nb][0] = acc[m][nb][1] = 0.0;

  auto consume = [&](int panel) {
    const int pstage = panel & (LK_PSTAGES - 1), wstage = panel & (LK_WSTAGES - 1);
    const double* Ps = s_main + pstage * LK_PANEL_DOUBLES;
    const double* Ws = s_wg + wstage * (2 * LK_WG_DOUBLES);
    const double* Gs = Ws + LK_WG_DOUBLES;
    mbar_wait(wg_bar(wstage), (panel / LK_WSTAGES) & 1);     // W/G written by all warps
    mbar_wait(panel_bar(pstage), (panel / LK_PSTAGES) & 1);  // basis panel landed
    const int aoff = (rh * 16 + grp) * LK_WSTRIDE + tig;
    const double* arow = Ws + aoff;
    const double* a4row = (g4 ? Gs : Ws) + aoff;
    const double* a5row = (g5 ? Gs : Ws) + aoff;
    const double* brow = Ps + tig * LK_PSTRIDE + first * 8 + grp;
#pragma unroll
    for (int kb = 0; kb < LK_KC / 4; ++kb) {
      double a[LK_MB];
#pragma unroll
      for (int m = 0; m < LK_MB; ++m) a[m] = arow[m * 8 * LK_WSTRIDE + kb * 4];
#pragma unroll
      for (int nb = 0; nb < 4; ++nb) {
        const double b = brow[kb * 4 * LK_PSTRIDE + nb * 8];
#pragma unroll
        for (int m = 0; m < LK_MB; ++m) dmma884(acc[m][nb][0], acc[m][nb][1], a[m], b);
      }
#pragma unroll
      for (int m = 0; m < LK_MB; ++m) a[m] = a4row[m * 8 * LK_WSTRIDE + kb * 4];
      {
        const double b = brow[kb * 4 * LK_PSTRIDE + 4 * 8];
#pragma unroll
        for (int m = 0; m < LK_MB; ++m) dmma884(acc[m][4][0], acc[m][4][1], a[m], b);
      }
#pragma unroll
      for (int m = 0; m < LK_MB; ++m) a[m] = a5row[m * 8 * LK_WSTRIDE + kb * 4];
#pragma unroll
      for (int nb = 5; nb < 7; ++nb) {
        const double b = brow[kb * 4 * LK_PSTRIDE + nb * 8];
#pragma unroll
        for (int m = 0; m < LK_MB; ++m) dmma884(acc[m][nb][0], acc[m][nb][1], a[m], b);
      }
    }
    if (has_eighth) {
#pragma unroll
      for (int kb = 0; kb < LK_KC / 4; ++kb) {
        const double b = brow[kb * 4 * LK_PSTRIDE + 7 * 8];
#pragma unroll
        for (int m = 0; m < LK_MB; ++m)
          dmma884(acc[m][7][0], acc[m][7][1], a5row[m * 8 * LK_WSTRIDE + kb * 4], b);
      }
    }
    if (panel + LK_PSTAGES < npanels) {
      // the last warp to leave the basis stage requests the panel that reuses it: nobody waits
      __syncwarp();
      if (lane == 0 && atomicAdd(&s_drain[pstage], 1) == LK_WARPS - 1) {
        s_drain[pstage] = 0;
        issue_panel(panel + LK_PSTAGES);
      }
    }
  };

  // ---- main loop: every warp produces panel c+1, then runs its DMMAs on panel c --------------------
  load_panel(0);
  produce(0);
  for (int panel = 0; panel < npanels; ++panel) {
    if (panel + 1 < npanels) produce(panel + 1);
    consume(panel);
    if ((panel & 7) == 7) renorm();
  }

  // per-sample scalar sums: reduce over the 16 pixel lanes of the half-warp
  renorm();
#pragma unroll
  for (int e = 0; e < LK_EPT; ++e) {
    double q = q_acc[e];
    double l = fma((double)esum[e], LK_LN2, log(dprod[e]));
#pragma unroll
    for (int off = 8; off > 0; off >>= 1) {
      q += __shfl_xor_sync(0xffffffffu, q, off);
      l += __shfl_xor_sync(0xffffffffu, l, off);
    }
    if (pl == 0) {
      s_sums[(sgrp + 16 * e) * 2] = q;
      s_sums[(sgrp + 16 * e) * 2 + 1] = l;
    }
  }
  __syncthreads();  // every panel consumed by every warp; the stages can be overlaid

  // ---- accumulators -> E[col][sample] -----------------------------------------------------------
  double* E = s_main;
  {
    const int count = has_eighth ? 8 : 7;
#pragma unroll
    for (int m = 0; m < LK_MB; ++m) {
      const int s = rh * 16 + m * 8 + grp;
#pragma unroll
      for (int nb = 0; nb < LK_NB_MAX; ++nb) {
        if (nb < count) {
          const int col = (first + nb) * 8 + tig * 2;
          E[col * LK_EP_STRIDE + s] = acc[m][nb][0];
          E[(col + 1) * LK_EP_STRIDE + s] = acc[m][nb][1];
        }
      }
    }
  }
  __syncthreads();

  // ---- Cholesky of the bordered matrix [[B, c], [c', q]] : 4 threads per sample ------------
  // E col layout: pair (i,j) at i(i+1)/2 + j ; projection c_j at 216 + j.
  // Row 20 of the bordered factor is z = L^-1 c, so quad = q - z'z (null_gp.py:345-358).
  if (tid < LK_TS * 4) {
    const int s = tid >> 2;       // sample of this thread quad
    const int t = tid & 3;
    double* Es = E + s;
    auto at = [&](int i, int j) -> double& {  // element (i,j), i >= j, i <= 20
      const int col = (i < LK_K) ? (i * (i + 1) / 2 + j) : (LK_PROJ_COL0 + j);
      return Es[col * LK_EP_STRIDE];
    };
    double logdet_prod = 1.0, logdet = 0.0, zz = 0.0;
    for (int j = 0; j < LK_K; ++j) {
      // pivot (all four threads compute it redundantly)
      double piv = at(j, j) + 1.0;  // + I (null_gp.py:341)
      for (int k = 0; k < j; ++k) { const double l = at(j, k); piv = fma(-l, l, piv); }
      logdet_prod *= piv;
      if ((j % 5) == 4) { logdet += log(logdet_prod); logdet_prod = 1.0; }
      const double inv = 1.0 / sqrt(piv);
      // rows j+1 .. 20 of column j, interleaved over the quad
      for (int i = j + 1 + t; i <= LK_K; i += 4) {
        double x = at(i, j);
        for (int k = 0; k < j; ++k) x = fma(-at(i, k), at(j, k), x);
        x *= inv;
        at(i, j) = x;
        if (i == LK_K) zz = fma(x, x, zz);
      }
      __syncwarp();
    }
    // row 20 moved round the quad from column to column; sum the partial z'z
    zz += __shfl_xor_sync(0xffffffffu, zz, 1);
    zz += __shfl_xor_sync(0xffffffffu, zz, 2);
    if (t == 0 && tile_s0 + s < sp.num_samples) {
      const double quad = s_sums[s * 2] - zz;
      const double log_det = s_sums[s * 2 + 1] + logdet;  // sum log d + 2 sum log L_ii
      sp.out[tile_s0 + s] = -0.5 * (quad + log_det + (double)n * LK_LOG_2PI);
    }
  }
}

// -------------------------------------------------------------------------------------------
// Generic single evaluation, any k <= 64: NullGP.log_mvnpdf_low_rank (null_gp.py:307-360)
// called with explicit arrays.  One CTA; used by the staticmethod drop-in and the KATs.
// -------------------------------------------------------------------------------------------
constexpr int LG_MAXK = 64;
__global__ void __launch_bounds__(256)
log_mvnpdf_low_rank_kernel(const double* __restrict__ y, const double* __restrict__ mu, const double* __restrict__ M,
                           const double* __restrict__ d, int n, int k, double* out) {
  __shared__ double B[(LG_MAXK + 1) * (LG_MAXK + 1)];  // bordered (k+1) x (k+1), row-major, lower part used
  __shared__ double red[2][8];
  const int tid = threadIdx.x;
  const int kb = k + 1;
  // B[i][j] = sum_p M[p][i] M[p][j] / d[p]  (i >= j) ; B[k][j] = sum_p M[p][j] r[p] / d[p]
  for (int e = tid; e < kb * kb; e += blockDim.x) {
    const int i = e / kb, j = e % kb;
    double acc = 0.0;
    if (j <= i && j < k) {
      for (int p = 0; p < n; ++p) {
        const double left = (i < k) ? M[(size_t)p * k + i] : (y[p] - mu[p]);
        acc = fma(left / d[p], M[(size_t)p * k + j], acc);
      }
      if (i == j) acc += 1.0;
    }
    B[e] = acc;
  }
  // scalar sums
  double q = 0.0, ld = 0.0;
  for (int p = tid; p < n; p += blockDim.x) {
    const double r = y[p] - mu[p];
    q = fma(r / d[p], r, q);
    ld += log(d[p]);
  }
  for (int off = 16; off > 0; off >>= 1) {
    q += __shfl_xor_sync(0xffffffffu, q, off);
    ld += __shfl_xor_sync(0xffffffffu, ld, off);
  }
  if ((tid & 31) == 0) { red[0][tid >> 5] = q; red[1][tid >> 5] = ld; }
  __syncthreads();
  if (tid == 0) {
    q = 0.0; ld = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { q += red[0][w]; ld += red[1][w]; }
    double logdet = 0.0, zz = 0.0;
    for (int j = 0; j < k; ++j) {
      double piv = B[j * kb + j];
      for (int c = 0; c < j; ++c) piv = fma(-B[j * kb + c], B[j * kb + c], piv);
      const double ljj = sqrt(piv);
      logdet += log(ljj);
      for (int i = j + 1; i <= k; ++i) {
        double x = B[i * kb + j];
        for (int c = 0; c < j; ++c) x = fma(-B[i * kb + c], B[j * kb + c], x);
        x /= ljj;
        B[i * kb + j] = x;
        if (i == k) zz = fma(x, x, zz);
      }
    }
    out[0] = -0.5 * ((q - zz) + (ld + 2.0 * logdet) + (double)n * LK_LOG_2PI);
  }
}

}  // namespace dla
