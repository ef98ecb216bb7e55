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
// on the FP64 tensor path (DMMA m8n8k4).
//
// The design follows from one measured fact: on B200 the DMMA path and the scalar FP64 pipe are
// the same 64 FMA lanes/SM/clk (37 TFLOP/s either way; mixed stream 35).  A scalar FP64 instruction
// issued while DMMAs saturate that pipe waits for a DMMA slot (16 cycles each, several queued), so
// a dependent chain of a dozen scalar operations costs a warp more time than its 60 DMMAs - measured:
// the overlapped version of this kernel ran at 22 TFLOP/s, 30 with the producer arithmetic removed.
// So scalar and tensor work are SEPARATED IN TIME instead of overlapped:
//   * a CTA of 8 warps owns 32 samples and walks the pixels in 16-pixel panels, phase A | barrier |
//     phase B | barrier; two CTAs share an SM, so one CTA's scalar phase runs under the other's DMMAs;
//   * phase A (all warps of the CTA, none of them issuing DMMAs): every thread turns 2 profile-cache elements (product of
//     up to 8 absorber factors, dla_gp.py:370-386, loaded during the previous phase B) into the W / G
//     operand tiles - MUFU-seeded Newton reciprocal, integer-renormalised running product for
//     sum log d (one log per lane per tile); uncontended, the whole phase is a few hundred cycles;
//   * phase B (all warps, nothing but LDS + DMMA): warp (rq, cq) owns 16 samples x 7 or 8 of the 30
//     column blocks, sized [8,7,8,7] / [7,8,7,8] by row half so that the two warps a CTA has on every SM
//     sub-partition issue 30 DMMAs per 4 pixels: no padding block, four balanced pipes; the loads
//     of the next panel's profile rows are in flight underneath;
//   * the Gram basis [P | M] of the spectrum (n x 240, precomputed once by gram_basis_kernel,
//     L2-resident) streams through a 2-deep shared-memory ring by TMA bulk copies
//     (cp.async.bulk + mbarrier complete_tx), requested a full panel ahead;
//   * the bordered 21 x 21 Cholesky (factor, z = L^-1 c, log-det) of the 32 samples is a third,
//     purely scalar phase: 8 threads per sample, straight out of the accumulator fragments.
// HBM sees only the profile rows (read) and one double per sample (written).
#pragma once
#include <stdint.h>

namespace dla {

constexpr int LK_K = 20;                         // rank of the learned covariance (Parameters.k)
constexpr int LK_PAIRS = LK_K * (LK_K + 1) / 2;  // 210 lower-triangle pairs
constexpr int LK_TS = 32;                        // samples per CTA tile
constexpr int LK_KC = 16;                        // pixels per panel
constexpr int LK_WSTRIDE = LK_KC + 4;            // row stride of the W/G tiles (== 4 mod 16: conflict-free A loads)
constexpr int LK_NBLK_PAIR = 27;                 // ceil(210 / 8) column blocks of the Gram part
constexpr int LK_NBLK = 30;                      // + 3 column blocks (24 >= 20) of the projection part
constexpr int LK_NCOLS = LK_NBLK * 8;            // 240
constexpr int LK_PSTRIDE = LK_NCOLS + 4;         // basis row stride (244 == 4 mod 16: conflict-free B loads)
constexpr int LK_PROJ_COL0 = LK_NBLK_PAIR * 8;   // 216: first projection column
constexpr int LK_WARPS = 8;
constexpr int LK_THREADS = LK_WARPS * 32;        // 512
constexpr int LK_PSTAGES = 2;                    // basis-panel ring (TMA)
constexpr int LK_CTAS_PER_SM = 2;               // (64 x 32 tiles, 16 warps, 1 CTA/SM measured 3 % slower)
constexpr int LK_EP_STRIDE = LK_TS + 1;          // epilogue smem: [col][sample], stride 65
constexpr int LK_MAX_ROWS = 8;                   // max absorbers multiplied per sample (max_dlas <= 8)
constexpr int LK_MB = 2;                         // DMMA row blocks per warp (16 samples)
constexpr int LK_NB_MAX = 8;                     // DMMA column blocks per warp (7 or 8)
constexpr int LK_EPT = LK_TS * LK_KC / LK_THREADS;  // W/G elements per thread per panel
constexpr int LK_PSTEP = LK_THREADS / LK_KC;        // sample stride between the elements of a thread
static_assert(LK_KC <= 32 && LK_THREADS % LK_KC == 0 && LK_TS * 8 == LK_THREADS && LK_WARPS == LK_TS / 4, "tile shape");
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

// cp.async (LDGSTS): global -> shared without a register round trip
__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gmem_src) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async16_zfill(void* smem_dst, const void* gmem_src, int src_bytes) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(d), "l"(gmem_src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }

// 1/d for the operand tiles: MUFU.RCP64H seed (SFU, not the FP64 pipe) + two Newton steps (4 DFMA),
// <= 1 ulp from the IEEE quotient.  Branch-free, so the four elements of a thread interleave: outside
// [1e-290, 1e290] (0, inf, NaN, negative: never produced by a valid spectrum) the seed itself is
// returned, which has the IEEE special-value behaviour (1/inf = 0, 1/0 = inf, NaN stays NaN).
__device__ __forceinline__ double fast_rcp(double d) {
  double r0;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r0) : "d"(d));
  double e = fma(-d, r0, 1.0);
  double r = fma(r0, e, r0);
  e = fma(-d, r, 1.0);
  r = fma(r, e, r);
  return (d > 1e-290 && d < 1e290) ? r : r0;
}

// ---- shared memory plan (two CTAs per SM) -----------------------------------------------------------
//   basis ring : 2 x [16][244]                                              62 464 B
//   W | G      : [32][20] each                                              10 240 B
//   RAW        : profile-row panels of factor 0 and factor 1 [32][16] each, pixel scalars [4][16]
//                (filled by cp.async during phase B, read in phase A)           8 704 B
//   E          : epilogue matrix [240][33], overlays the above              63 360 B
//   AUX        : per-sample sums [32][2], profile rows [8][32], 2 mbarriers
constexpr int LK_PANEL_DOUBLES = LK_KC * LK_PSTRIDE;               // 3904
constexpr int LK_WG_DOUBLES = LK_TS * LK_WSTRIDE;                  // 640
constexpr uint32_t LK_PANEL_BYTES = LK_PANEL_DOUBLES * sizeof(double);  // 31 232
constexpr int LK_RAW_DOUBLES = 2 * LK_TS * LK_KC + 4 * LK_KC;       // 1088
constexpr int LK_RING_DOUBLES = LK_PSTAGES * LK_PANEL_DOUBLES + 2 * LK_WG_DOUBLES + LK_RAW_DOUBLES;  // 10 176
constexpr int LK_EP_DOUBLES = LK_NCOLS * LK_EP_STRIDE;             // 7920
constexpr int LK_MAIN_DOUBLES = LK_RING_DOUBLES > LK_EP_DOUBLES ? LK_RING_DOUBLES : LK_EP_DOUBLES;
constexpr size_t LK_AUX_BYTES = LK_TS * 2 * sizeof(double) + LK_MAX_ROWS * LK_TS * sizeof(int32_t) +
                                LK_PSTAGES * sizeof(uint64_t);
constexpr size_t LK_SMEM_BYTES = (size_t)LK_MAIN_DOUBLES * sizeof(double) + LK_AUX_BYTES;
static_assert(LK_PANEL_BYTES % 128 == 0, "TMA alignment");

// grid = (ceil(max num_samples / 32), num_spectra), block = 256, dynamic smem = LK_SMEM_BYTES
__global__ void __launch_bounds__(LK_THREADS, LK_CTAS_PER_SM)
sample_likelihood_kernel(const LikelihoodSpectrum* __restrict__ specs) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const LikelihoodSpectrum sp = specs[blockIdx.y];
  const int tile_s0 = blockIdx.x * LK_TS;
  if (tile_s0 >= sp.num_samples) return;
  if (sp.alive && *sp.alive == 0) return;

  double* s_main = reinterpret_cast<double*>(smem_raw);
  double* s_W = s_main + LK_PSTAGES * LK_PANEL_DOUBLES;
  double* s_G = s_W + LK_WG_DOUBLES;
  double* s_raw0 = s_G + LK_WG_DOUBLES;      // [64][32] factor-0 profile values of the staged panel
  double* s_raw1 = s_raw0 + LK_TS * LK_KC;   // [64][32] factor-1 profile values
  double* s_pix = s_raw1 + LK_TS * LK_KC;    // [4][32]  y, mu, omega2, v of the staged panel
  double* s_sums = s_main + LK_MAIN_DOUBLES;                          // [64][2] : sum r^2/d, sum log d
  int32_t* s_rows = reinterpret_cast<int32_t*>(s_sums + LK_TS * 2);   // [num_rows][64]
  uint64_t* s_mbar = reinterpret_cast<uint64_t*>(s_rows + LK_MAX_ROWS * LK_TS);  // basis panel landed [2]

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int n = sp.n;
  const int npanels = (n + LK_KC - 1) / LK_KC;
  const int num_rows = sp.num_rows;
  const uint32_t bar_base = (uint32_t)__cvta_generic_to_shared(s_mbar);

  // basis panel `panel` -> its ring stage (one thread; completion is the stage's mbarrier)
  auto issue_panel = [&](int panel) {
    const int stage = panel & (LK_PSTAGES - 1);
    mbar_expect_tx(bar_base + 8u * stage, LK_PANEL_BYTES);
    tma_bulk_g2s((uint32_t)__cvta_generic_to_shared(s_main + stage * LK_PANEL_DOUBLES),
                 sp.P + (size_t)panel * LK_PANEL_DOUBLES, LK_PANEL_BYTES, bar_base + 8u * stage);
  };

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
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < LK_PSTAGES; ++s) mbar_init(bar_base + 8u * s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    issue_panel(0);
  }
  __syncthreads();

  // ---- producer role: thread -> pixel pl of the panel and samples ps0 + e * LK_PSTEP ----------------
  // (consecutive threads store consecutive doubles of one W/G row: conflict-free; and read contiguous
  //  bytes of a profile row)
  const int pl = tid % LK_KC, ps0 = tid / LK_KC;
  const bool two_rows = num_rows > 1;
  unsigned live = 0;
#pragma unroll
  for (int e = 0; e < LK_EPT; ++e) live |= (tile_s0 + ps0 + LK_PSTEP * e < sp.num_samples ? 1u : 0u) << e;
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

  // stage the inputs of `panel` in shared memory with cp.async: issued at the start of phase B, so the
  // copies fly underneath the DMMAs and no register is held for them (a register prefetch of 4 elements
  // x 2 factors spilled, and the spill store waited for the load)
  // `nthreads` threads with rank `r` (the whole CTA in the prologue; in the main loop the warps that own
  // only 7 column blocks, i.e. the ones with 1/8 less DMMA work)
  auto stage_panel = [&](int panel, int r, int nthreads) {
    const int p0 = panel * LK_KC;
    for (int idx = r; idx < LK_TS * LK_KC / 2; idx += nthreads) {  // 16-byte chunks: KC / 2 per sample row
      const int s = idx / (LK_KC / 2), j = (idx % (LK_KC / 2)) * 2;
      const int p = p0 + j;
      const int ok = p < sp.ld ? 16 : 0;  // rows are padded to ld (multiple of 4): zero-fill beyond
      const size_t off = (size_t)min(p, sp.ld - 2);
      cp_async16_zfill(s_raw0 + s * LK_KC + j, sp.base0 + (size_t)s_rows[s] * sp.ld + off, ok);
      if (two_rows) cp_async16_zfill(s_raw1 + s * LK_KC + j, sp.cache + (size_t)s_rows[LK_TS + s] * sp.ld + off, ok);
    }
    if (r < 4 * LK_KC) {
      const int arr = r / LK_KC, j = r % LK_KC;
      const double* src = arr == 0 ? sp.y : arr == 1 ? sp.mu : arr == 2 ? sp.omega2 : sp.v;
      cp_async8(s_pix + arr * LK_KC + j, src + min(p0 + j, n - 1));
    }
    cp_async_commit();
  };

  // phase A: W/G tiles of `panel`.  Branch-free over the four elements of a thread so that their
  // dependent chains (load, products, reciprocal, ...) interleave.  Pixels beyond n are neutralised once per
  // thread (a = 0, y = mu = omega2 = 0, v = 1  =>  w = g = 0, d = 1); samples beyond num_samples (last
  // tile of a spectrum) compute on row 0 and are simply never written out - no per-element selects.
  auto produce = [&](int panel) {
    const int p = panel * LK_KC + pl;
    const bool pv = p < n;
    const double yp = pv ? s_pix[pl] : 0.0, mup = pv ? s_pix[LK_KC + pl] : 0.0;
    const double omp = pv ? s_pix[2 * LK_KC + pl] : 0.0, vp = pv ? s_pix[3 * LK_KC + pl] : 1.0;
    double a[LK_EPT];
#pragma unroll
    for (int e = 0; e < LK_EPT; ++e) a[e] = s_raw0[(ps0 + LK_PSTEP * e) * LK_KC + pl];
    // absorption = product of the factors' profiles, left to right (dla_gp.py:370-386)
    if (two_rows) {
#pragma unroll
      for (int e = 0; e < LK_EPT; ++e) a[e] = a[e] * s_raw1[(ps0 + LK_PSTEP * e) * LK_KC + pl];
    }
    if (num_rows > 2) {  // rare: dla_sample_log_likelihoods with more than two absorbers per sample
      const int pc = min(p, n - 1);
#pragma unroll
      for (int e = 0; e < LK_EPT; ++e)
        for (int r = 2; r < num_rows; ++r)
          a[e] = a[e] * sp.cache[(size_t)s_rows[r * LK_TS + ps0 + LK_PSTEP * e] * sp.ld + pc];
    }
    if (sp.prod_out && pv) {
#pragma unroll
      for (int e = 0; e < LK_EPT; ++e)
        if ((live >> e) & 1u) sp.prod_out[(size_t)(tile_s0 + ps0 + LK_PSTEP * e) * sp.ld + p] = a[e];
    }
    // the four elements advance in lock step (one loop per operation): four independent dependency
    // chains in flight instead of one after the other
    double a2[LK_EPT], d[LK_EPT], r0[LK_EPT], er[LK_EPT], inv[LK_EPT], res[LK_EPT], t[LK_EPT];
#pragma unroll
    for (int e = 0; e < LK_EPT; ++e) a[e] = pv ? a[e] : 0.0;  // pad columns of the profile rows hold no data
#pragma unroll
    for (int e = 0; e < LK_EPT; ++e) a2[e] = a[e] * a[e];
#pragma unroll
    for (int e = 0; e < LK_EPT; ++e) d[e] = fma(omp, a2[e], vp);          // dla_omega2 + v
#pragma unroll
    for (int e = 0; e < LK_EPT; ++e) asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r0[e]) : "d"(d[e]));
#pragma unroll
    for (int e = 0; e < LK_EPT; ++e) res[e] = fma(-mup, a[e], yp);        // y - dla_mu
#pragma unroll
    for (int e = 0; e < LK_EPT; ++e) er[e] = fma(-d[e], r0[e], 1.0);
#pragma unroll
    for (int e = 0; e < LK_EPT; ++e) inv[e] = fma(r0[e], er[e], r0[e]);
#pragma unroll
    for (int e = 0; e < LK_EPT; ++e) er[e] = fma(-d[e], inv[e], 1.0);
#pragma unroll
    for (int e = 0; e < LK_EPT; ++e) inv[e] = fma(inv[e], er[e], inv[e]);
#pragma unroll
    for (int e = 0; e < LK_EPT; ++e) {
      // see fast_rcp: outside roughly [1e-290, 1e290] (and for 0, inf, NaN, negative d) keep the seed.  The test
      // is on the exponent field with integer instructions - a DSETP would queue on the FP64 pipe like a DFMA
      const unsigned hi = (unsigned)__double2hiint(d[e]);
      inv[e] = (hi - 0x03d00000u < 0x7c300000u - 0x03d00000u) ? inv[e] : r0[e];
    }
#pragma unroll
    for (int e = 0; e < LK_EPT; ++e) t[e] = res[e] * inv[e];
#pragma unroll
    for (int e = 0; e < LK_EPT; ++e) {
      const int s = ps0 + LK_PSTEP * e;
      s_W[s * LK_WSTRIDE + pl] = a2[e] * inv[e];
      s_G[s * LK_WSTRIDE + pl] = a[e] * t[e];
    }
#pragma unroll
    for (int e = 0; e < LK_EPT; ++e) q_acc[e] = fma(res[e], t[e], q_acc[e]);
#pragma unroll
    for (int e = 0; e < LK_EPT; ++e) dprod[e] *= d[e];
  };

  // ---- MMA role: warp (rq, cq) owns samples 16 rq .. 16 rq + 15 and 7 or 8 column blocks -----------
  //   rq even : blocks [0,8) [8,15) [15,23) [23,30)      rq odd : blocks [0,7) [7,15) [15,22) [22,30)
  // Blocks >= 27 are the projection part (A = G).  Local blocks 0..3 always take A = W, local block 4
  // and local blocks 5..7 read A through pointers (W or G), so all warps run one instruction stream
  // with no predicated DMMA (a predicated-off DMMA still occupies its pipe slot); the eighth block
  // sits behind a warp-uniform branch.
  const int grp = lane >> 2;      // DMMA groupID
  const int tig = lane & 3;       // DMMA threadID_in_group
  const int rq = warp >> 2, cq = warp & 3;
  const bool odd = (rq & 1) != 0;
  const int first = !odd ? (cq == 0 ? 0 : cq == 1 ? 8 : cq == 2 ? 15 : 23) : (cq == 0 ? 0 : cq == 1 ? 7 : cq == 2 ? 15 : 22);
  const bool has_eighth = ((cq + (odd ? 1 : 0)) & 1) == 0;
  const bool g4 = first + 4 >= LK_NBLK_PAIR, g5 = first + 5 >= LK_NBLK_PAIR;
  const int stage_rank = (rq * 2 + (cq >> 1)) * 32 + lane;  // rank among the threads of the 7-block warps
  double acc[LK_MB][LK_NB_MAX][2];
#pragma unroll
  for (int m = 0; m < LK_MB; ++m)
#pragma unroll
    for (int nb = 0; nb < LK_NB_MAX; ++nb) acc[m][nb][0] = acc[m][nb][1] = 0.0;

  // phase B: DMMAs of `panel`
  auto consume = [&](int panel, bool stage_next) {
    const int pstage = panel & (LK_PSTAGES - 1);
    const double* Ps = s_main + pstage * LK_PANEL_DOUBLES;
    mbar_wait(bar_base + 8u * pstage, (panel / LK_PSTAGES) & 1);  // basis panel landed
    const int aoff = (rq * 16 + grp) * LK_WSTRIDE + tig;
    const double* arow = s_W + aoff;
    const double* a4row = (g4 ? s_G : s_W) + aoff;
    const double* a5row = (g5 ? s_G : s_W) + aoff;
    const double* brow = Ps + tig * LK_PSTRIDE + first * 8 + grp;
#pragma unroll
    for (int kb = 0; kb < LK_KC / 4; ++kb) {
      // the cp.async requests of the next panel go out after the first DMMAs, so that their address
      // arithmetic runs while the tensor pipe already has work queued; they are issued by the warps that
      // own 7 column blocks, which otherwise idle at the barrier while the 8-block warps finish
      if (kb == 1 && stage_next && !has_eighth) stage_panel(panel + 1, stage_rank, (LK_WARPS / 2) * 32);
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
  };

  // ---- main loop: phase A | barrier | phase B | barrier --------------------------------------------------
  stage_panel(0, tid, LK_THREADS);
  cp_async_wait_all();
  __syncthreads();
  for (int panel = 0; panel < npanels; ++panel) {
    if (tid == 0 && panel + 1 < npanels) issue_panel(panel + 1);  // its stage was drained by phase B of panel - 1
    produce(panel);
    __syncthreads();                                              // W/G complete, staged inputs consumed
    consume(panel, panel + 1 < npanels);                          // stages the next panel's inputs underneath the DMMAs
    if ((panel & 3) == 3) renorm();
    cp_async_wait_all();
    __syncthreads();                                              // W/G and the basis stage are free, inputs staged
  }

  // per-sample scalar sums: reduce over the pixel lanes of the panel
  renorm();
#pragma unroll
  for (int e = 0; e < LK_EPT; ++e) {
    double q = q_acc[e];
    double l = fma((double)esum[e], LK_LN2, log(dprod[e]));
#pragma unroll
    for (int off = (LK_KC < 32 ? LK_KC : 32) / 2; off > 0; off >>= 1) {
      q += __shfl_xor_sync(0xffffffffu, q, off);
      l += __shfl_xor_sync(0xffffffffu, l, off);
    }
    if (pl == 0) {
      s_sums[(ps0 + LK_PSTEP * e) * 2] = q;
      s_sums[(ps0 + LK_PSTEP * e) * 2 + 1] = l;
    }
  }

  // ---- accumulators -> E[col][sample] (overlays the ring; the last barrier of the loop freed it) --------
  double* E = s_main;
  {
    const int count = has_eighth ? 8 : 7;
#pragma unroll
    for (int m = 0; m < LK_MB; ++m) {
      const int s = rq * 16 + m * 8 + grp;
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

  // ---- Cholesky of the bordered matrix [[B, c], [c', q]] : 8 threads per sample, in registers --------------
  // E col layout: pair (i,j) at i(i+1)/2 + j ; projection c_j at 216 + j.  Row 20 of the bordered factor is
  // z = L^-1 c, so quad = q - z'z (null_gp.py:345-358).  Thread t of a sample's group owns rows t, t + 8 and
  // t + 16 of the lower triangle (row 20 = the projection row) in three register arrays with compile-time
  // indices; column j needs the finished entries of row j, which its owner broadcasts by shuffles, and every
  // dot product is two independent FMA chains.  (The first version walked the matrix in shared memory with
  // two loads per FMA and took 13 % of the kernel with the tensor pipe idle.)
  {
    const int s = tid >> 3;          // sample of this thread group (8 threads per sample)
    const int t = tid & 7;
    const int gbase = lane & ~7;
    const double* Es = E + s;
    double r0[8], r1[16], r2[20];
#pragma unroll
    for (int k = 0; k < 8; ++k) r0[k] = k <= t ? Es[(t * (t + 1) / 2 + k) * LK_EP_STRIDE] + (k == t ? 1.0 : 0.0) : 0.0;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const int i = t + 8;
      r1[k] = k <= i ? Es[(i * (i + 1) / 2 + k) * LK_EP_STRIDE] + (k == i ? 1.0 : 0.0) : 0.0;
    }
#pragma unroll
    for (int k = 0; k < 20; ++k) {
      const int i = t + 16;  // rows 16..19 of B, row 20 = c, rows 21..23 do not exist
      double val = 0.0;
      if (i < LK_K) { if (k <= i) val = Es[(i * (i + 1) / 2 + k) * LK_EP_STRIDE] + (k == i ? 1.0 : 0.0); }
      else if (i == LK_K) val = Es[(LK_PROJ_COL0 + k) * LK_EP_STRIDE];
      r2[k] = val;
    }
    double piv_prod = 1.0;
#pragma unroll
    for (int j = 0; j < LK_K; ++j) {
      const int owner = gbase + (j & 7);
      // finished entries of row j (k < j), from its owner
      double Lj[LK_K];
#pragma unroll
      for (int k = 0; k < j; ++k)
        Lj[k] = __shfl_sync(0xffffffffu, j < 8 ? r0[k < 8 ? k : 0] : j < 16 ? r1[k < 16 ? k : 0] : r2[k], owner);
      // x_q = A[i_q][j] - sum_k L[i_q][k] L[j][k] for the rows of every slot that reaches column j
      double x0 = 0.0, x1 = 0.0, x2 = r2[j];
      if (j < 8) {
        double e0 = r0[j], e1 = 0.0;
#pragma unroll
        for (int k = 0; k + 1 < j; k += 2) { e0 = fma(-r0[k], Lj[k], e0); e1 = fma(-r0[k + 1], Lj[k + 1], e1); }
        if (j & 1) e0 = fma(-r0[j - 1], Lj[j - 1], e0);
        x0 = e0 + e1;
      }
      if (j < 16) {
        double e0 = r1[j], e1 = 0.0;
#pragma unroll
        for (int k = 0; k + 1 < j; k += 2) { e0 = fma(-r1[k], Lj[k], e0); e1 = fma(-r1[k + 1], Lj[k + 1], e1); }
        if (j & 1) e0 = fma(-r1[j - 1], Lj[j - 1], e0);
        x1 = e0 + e1;
      }
      {
        double e0 = x2, e1 = 0.0;
#pragma unroll
        for (int k = 0; k + 1 < j; k += 2) { e0 = fma(-r2[k], Lj[k], e0); e1 = fma(-r2[k + 1], Lj[k + 1], e1); }
        if (j & 1) e0 = fma(-r2[j - 1], Lj[j - 1], e0);
        x2 = e0 + e1;
      }
      // the pivot is the x of row j itself
      const double piv = __shfl_sync(0xffffffffu, j < 8 ? x0 : j < 16 ? x1 : x2, owner);
      piv_prod *= piv;
      const double inv = rsqrt(piv);
      if (j < 8) r0[j] = x0 * inv;    // rows below j become column j of L (the pivot row's own entry is not used again)
      if (j < 16) r1[j] = x1 * inv;
      r2[j] = x2 * inv;
    }
    // z'z on the owner of row 20 (t = 4, third slot)
    if (t == 4 && tile_s0 + s < sp.num_samples) {
      double zz0 = 0.0, zz1 = 0.0;
#pragma unroll
      for (int k = 0; k < LK_K; k += 2) { zz0 = fma(r2[k], r2[k], zz0); zz1 = fma(r2[k + 1], r2[k + 1], zz1); }
      const double quad = s_sums[s * 2] - (zz0 + zz1);
      const double log_det = s_sums[s * 2 + 1] + log(piv_prod);  // sum log d + 2 sum log L_ii
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
