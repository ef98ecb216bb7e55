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
// The design follows from measurements on B200 (tools/fp64_peaks.cu, fp64_contention_probe.cu,
// fp64_interleave_probe.cu, dmma_feed_probe.cu):
//   * the DMMA path and the scalar FP64 pipe are the same 64 FMA lanes/SM/clk (37 TFLOP/s either way);
//   * a warp issuing scalar FP64 operations is STARVED while two or more other warps of its SM sub-partition
//     issue back-to-back DMMAs (a dependent DFMA then takes > 20 000 cycles), so producer warps, or a producer
//     phase of one CTA running under the DMMA phase of another, do not work (18.5 and 26.9 TFLOP/s);
//   * when every warp carries the same mix - two DMMAs, one scalar operation - the pipe stays 91 % busy;
//   * a DMMA loop fed from shared memory in this kernel's layout, with a CTA barrier per 16-pixel panel,
//     runs at 95-97 % of the peak, and every non-FP64 instruction a warp executes per panel (address
//     arithmetic, staging, selects) costs throughput: the warps of a sub-partition take turns on one pipe and
//     their own instruction streams are the think time of that queue.
// So every warp runs ONE instruction stream per panel: the DMMAs of panel p with the scalar arithmetic that
// produces the W / G operand tiles of panel p + 1 threaded through them, and as little else as possible:
//   * a CTA of 8 warps owns 32 samples and walks the pixels in 16-pixel panels (no CTA barrier in the loop: two split
//     mbarriers, full / empty, one arrival per warp, waited a quarter of a panel after the arrivals), two
//     panels per loop trip so that all buffer parities are compile-time and every shared-memory address is a
//     per-thread base plus an immediate; two CTAs share an SM;
//   * scalar work: every thread turns 2 profile-cache elements (product of up to 8 absorber factors,
//     dla_gp.py:370-386) into W / G entries - MUFU-seeded reciprocal with one cubic correction,
//     integer-renormalised running product for sum log d (one log per sample per tile), no validity selects
//     (the last panel is staged with neutral values beyond n);
//   * DMMA work: warp (rq, cq) owns 16 samples x 7 or 8 of the 30 column blocks, sized [8,7,8,7] / [7,8,7,8]
//     by row half so that the two warps a CTA has on every SM sub-partition issue 30 DMMAs per 4 pixels: no
//     padding block, four balanced pipes;
//   * the profile rows of panel p + 2 are staged by cp.async (row pointers from a shared-memory table: a
//     pointer load, an add and the LDGSTS per 16 bytes), issued by the warps that own only 7 blocks;
//   * the Gram basis [P | M] of the spectrum (n x 240, precomputed once by gram_basis_kernel,
//     L2-resident) streams through a 2-deep shared-memory ring by TMA bulk copies
//     (cp.async.bulk + mbarrier complete_tx), requested a full panel ahead;
//   * the bordered 21 x 21 Cholesky (factor, z = L^-1 c, log-det) of the 32 samples is a final,
//     purely scalar phase: 8 threads per sample, straight out of the accumulator fragments.
// HBM sees only the profile rows (read) and one double per sample (written).
#pragma once
#include <stdint.h>
#include <type_traits>

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
constexpr int LK_THREADS = LK_WARPS * 32;        // 256
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
  int num_samples;       // samples in this launch; a NEGATIVE value -i means "read the count from alive[i]": the launch
                         // was compacted on the device (compact_level_kernel; rows0 / rows / prod_out / out are then in
                         // slot order).  Encoded in the existing field because growing this struct by one pointer
                         // changed ptxas's register allocation of the main loop (8-byte spill per panel).
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

// Order-pinned variants for the interleaved main loop: volatile asm statements keep their program order,
// so the scalar operations stay where they were placed between the DMMAs.
__device__ __forceinline__ void dmma884_pinned(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}
__device__ __forceinline__ double fma_pinned(double a, double b, double c) {
  double r;
  asm volatile("fma.rn.f64 %0, %1, %2, %3;" : "=d"(r) : "d"(a), "d"(b), "d"(c));
  return r;
}
__device__ __forceinline__ double fnma_pinned(double a, double b, double c) {  // -a b + c
  double r;
  asm volatile("{\n.reg .f64 na;\nneg.f64 na, %1;\nfma.rn.f64 %0, na, %2, %3;\n}" : "=d"(r) : "d"(a), "d"(b), "d"(c));
  return r;
}
__device__ __forceinline__ double mul_pinned(double a, double b) {
  double r;
  asm volatile("mul.rn.f64 %0, %1, %2;" : "=d"(r) : "d"(a), "d"(b));
  return r;
}
__device__ __forceinline__ double rcp_seed(double d) {  // MUFU.RCP64H: SFU, not the FP64 pipe
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
  return r;
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
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
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
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }

// 1 / sqrt(x) for the Cholesky pivots: MUFU.RSQ64H seed y0 and one cubic correction y0 (1 + e/2 + 3 e^2/8),
// e = 1 - x y0^2 <= 2^-20 (5 dependent FP64 operations; the library routine issues twice as many and a branch).
// Pivots of I + M'D^-1 M are >= 1; zero, negative or NaN pivots (NaN input) come out as inf / NaN and poison the sample.
__device__ __forceinline__ double fast_rsqrt(double x) {
  double y0;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(x));
  const double e = fma(-(x * y0), y0, 1.0);
  const double p = fma(e, 0.375, 0.5) * e;
  return fma(y0, p, y0);
}

// factors 2 .. num_rows - 1 of a sample (dla_sample_log_likelihoods with more than two absorbers per sample):
// rare, kept out of line so that the main loop stays small
__device__ __noinline__ double times_extra_factors(double a, const double* cache, const int32_t* rows_of_sample,
                                                   int num_rows, int ld, int p) {
  for (int r = 2; r < num_rows; ++r) a = a * cache[(size_t)rows_of_sample[r * LK_TS] * ld + p];
  return a;
}

// 1/d as the scalar slots of sample_likelihood_kernel compute it (operations 4-6 of a chain; written out here as
// one function for reference and for the unit test of the formula): MUFU.RCP64H seed r0 (SFU, not the FP64 pipe,
// relative error e = 1 - d r0 <= 2^-20) and one cubic correction r0 (1 + e + e^2): 3 dependent DFMA, truncation
// error e^3 < 2^-60, <= 1 ulp from the IEEE quotient.  Outside [1e-290, 1e290] (0, inf, NaN, negative: never
// produced by a valid spectrum) the seed itself is returned, which has the IEEE special-value behaviour
// (1/inf = 0, 1/0 = inf, NaN stays NaN).
__device__ __forceinline__ double fast_rcp(double d) {
  double r0;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r0) : "d"(d));
  double e = fma(-d, r0, 1.0);
  e = fma(e, e, e);
  const double r = fma(r0, e, r0);
  return (d > 1e-290 && d < 1e290) ? r : r0;
}

// ---- shared memory plan (two CTAs per SM, 102 448 B each) ----------------------------------------------
//   basis ring : 2 x [16][244]                                                          62 464 B
//   W | G      : 2 buffers x { W [32][20], G [32][20] } (read by the DMMAs / written for the next panel) 20 480 B
//   RAW        : 2 buffers x { profile-row panels of factor 0 and factor 1 [32][16] each, y, mu, omega2, v [4][16] }
//                (read by the scalar slots / filled by cp.async for the panel after)           17 408 B
//   E          : epilogue matrix [240][33], overlays the above                          63 360 B
//   AUX        : per-sample sums [32][3], profile row indices [8][32], row pointers [2][32] + 4, 2 mbarriers
constexpr int LK_PANEL_DOUBLES = LK_KC * LK_PSTRIDE;               // 3904
constexpr int LK_WG_DOUBLES = LK_TS * LK_WSTRIDE;                  // 640
constexpr uint32_t LK_PANEL_BYTES = LK_PANEL_DOUBLES * sizeof(double);  // 31 232
constexpr int LK_RAW_DOUBLES = 2 * LK_TS * LK_KC + 4 * LK_KC;       // 1088
constexpr int LK_RING_DOUBLES = LK_PSTAGES * LK_PANEL_DOUBLES + 2 * (2 * LK_WG_DOUBLES + LK_RAW_DOUBLES);  // 12 544
constexpr int LK_EP_DOUBLES = LK_NCOLS * LK_EP_STRIDE;             // 7920
constexpr int LK_MAIN_DOUBLES = LK_RING_DOUBLES > LK_EP_DOUBLES ? LK_RING_DOUBLES : LK_EP_DOUBLES;
constexpr size_t LK_AUX_BYTES = LK_TS * 3 * sizeof(double) + LK_MAX_ROWS * LK_TS * sizeof(int32_t) +
                                (2 * LK_TS + 4) * sizeof(void*) + (LK_PSTAGES + 2) * sizeof(uint64_t);
constexpr size_t LK_SMEM_BYTES = (size_t)LK_MAIN_DOUBLES * sizeof(double) + LK_AUX_BYTES;
static_assert(LK_PANEL_BYTES % 128 == 0, "TMA alignment");

// grid = (ceil(max num_samples / 32), num_spectra), block = 256, dynamic smem = LK_SMEM_BYTES
__global__ void __launch_bounds__(LK_THREADS, LK_CTAS_PER_SM)
sample_likelihood_kernel(const LikelihoodSpectrum* __restrict__ specs) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  LikelihoodSpectrum sp = specs[blockIdx.y];
  const int tile_s0 = blockIdx.x * LK_TS;
  if (sp.alive && *sp.alive == 0) return;
  if (sp.num_samples < 0) sp.num_samples = sp.alive[-sp.num_samples];
  if (tile_s0 >= sp.num_samples) return;

  double* s_main = reinterpret_cast<double*>(smem_raw);
  double* s_WG = s_main + LK_PSTAGES * LK_PANEL_DOUBLES;  // 2 buffers x { W [32][20], G [32][20] }
  double* s_RAW = s_WG + 2 * 2 * LK_WG_DOUBLES;           // 2 buffers x { factor-0 rows [32][16], factor-1 rows [32][16],
                                                          //               y, mu, omega2, v [4][16] } of a staged panel
  double* s_sums = s_main + LK_MAIN_DOUBLES;                          // [32][3] : sum r^2/d; mantissa product and exponent sum of prod d
  int32_t* s_rows = reinterpret_cast<int32_t*>(s_sums + LK_TS * 3);   // [num_rows][32]
  const double** s_ptr0 = reinterpret_cast<const double**>(s_rows + LK_MAX_ROWS * LK_TS);  // [32] factor-0 row of each sample
  const double** s_ptr1 = s_ptr0 + LK_TS;                                                    // [32] factor-1 row
  const double** s_pixptr = s_ptr1 + LK_TS;                                                  // y, mu, omega2, v
  uint64_t* s_mbar = reinterpret_cast<uint64_t*>(s_pixptr + 4);  // basis panel landed [2]

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int n = sp.n;
  const int npanels = (n + LK_KC - 1) / LK_KC;
  const int num_rows = sp.num_rows;
  const uint32_t bar_base = (uint32_t)__cvta_generic_to_shared(s_mbar);
  // Split barriers between the warps of the CTA (one arrival per warp, mbarrier phase = panel): `full` completes when
  // every warp has stored its W/G entries of the next panel and its staged inputs have landed; `empty` when every
  // warp has issued its last DMMA of the panel.  A warp waits for `full` at the top of a panel and for the previous
  // panel's `empty` only before its first store (a quarter of the way in), so the warps may drift apart by that much
  // instead of meeting at a CTA barrier every panel (the barrier was 9 % of all warp stall samples).
  const uint32_t bar_full = bar_base + 8u * LK_PSTAGES, bar_empty = bar_full + 8u;
  auto warp_arrive = [&](uint32_t bar) {
    __syncwarp();
    if (lane == 0) mbar_arrive(bar);
  };

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
    if (r == 0) s_ptr0[s] = sp.base0 + (size_t)row * sp.ld;
    if (r == 1) s_ptr1[s] = sp.cache + (size_t)row * sp.ld;
  }
  if (tid < 4) s_pixptr[tid] = tid == 0 ? sp.y : tid == 1 ? sp.mu : tid == 2 ? sp.omega2 : sp.v;
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < LK_PSTAGES; ++s) mbar_init(bar_base + 8u * s, 1);
    mbar_init(bar_full, LK_WARPS);
    mbar_init(bar_empty, LK_WARPS);
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

  // stage the inputs of `panel` in RAW buffer PAR (= panel & 1) with cp.async: the copies fly underneath the
  // DMMAs and no register is held for them.  `nthreads` threads with rank `r` (the whole CTA in the
  // prologue; in the main loop the warps that own only 7 column blocks, i.e. 1/8 less DMMA work).
  // A thread always copies the same 16-byte column chunk jc of samples r / 8 + k nthreads / 8; the row base
  // pointers come from a shared-memory table, so a chunk costs a pointer load, an add and the LDGSTS.
  // Every panel but the last lies inside all rows (p0 + 16 <= n <= ld).  The last one is copied element by
  // element, and what lies beyond n is written as neutral values (a = 0, y = mu = omega2 = 0, v = 1 =>
  // w = g = 0, d = 1) so that the arithmetic needs no validity test.
  auto stage_panel = [&](const int panel, auto par_tag, const int r, const int nthreads) {
    constexpr int PAR = decltype(par_tag)::value;
    double* raw0 = s_RAW + PAR * LK_RAW_DOUBLES;
    double* raw1 = raw0 + LK_TS * LK_KC;
    double* pix = raw1 + LK_TS * LK_KC;
    const int p0 = panel * LK_KC;
    const int jc = (r % (LK_KC / 2)) * 2;
    if (panel + 1 < npanels) {
      const size_t boff = (size_t)(p0 + jc) * sizeof(double);
      for (int sidx = r / (LK_KC / 2); sidx < LK_TS; sidx += nthreads / (LK_KC / 2)) {
        cp_async16(raw0 + sidx * LK_KC + jc, reinterpret_cast<const char*>(s_ptr0[sidx]) + boff);
        if (two_rows) cp_async16(raw1 + sidx * LK_KC + jc, reinterpret_cast<const char*>(s_ptr1[sidx]) + boff);
      }
      if (r < 4 * LK_KC) cp_async8(pix + r, s_pixptr[r / LK_KC] + p0 + r % LK_KC);
    } else {
      for (int sidx = r / (LK_KC / 2); sidx < LK_TS; sidx += nthreads / (LK_KC / 2)) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int p = p0 + jc + h;
          if (p < n) {
            cp_async8(raw0 + sidx * LK_KC + jc + h, s_ptr0[sidx] + p);
            if (two_rows) cp_async8(raw1 + sidx * LK_KC + jc + h, s_ptr1[sidx] + p);
          } else {
            raw0[sidx * LK_KC + jc + h] = 0.0;
            raw1[sidx * LK_KC + jc + h] = 0.0;
          }
        }
      }
      if (r < 4 * LK_KC) {
        const int arr = r / LK_KC, p = p0 + r % LK_KC;
        if (p < n) cp_async8(pix + r, s_pixptr[arr] + p);
        else pix[r] = arr == 3 ? 1.0 : 0.0;
      }
    }
    cp_async_commit();
  };

  // ---- the scalar chain of one W/G element, cut into slots ------------------------------------------------
  // A thread owns two elements (pixel pl, samples ps0 and ps0 + 16) of every panel.  Their arithmetic
  //     a = prod of the factors' profiles (dla_gp.py:370-386);  d = omega2 a^2 + v;  r = y - mu a;
  //     w = a^2 / d;  g = a r / d;  q += r^2 / d;  prod d
  // is 12 dependent FP64 operations each (11 for a single factor).  Slot 12 e + i is operation i of element e (one chain
  // after the other measured 1.7 % faster than the two interleaved: fewer live registers); the main loop places one
  // slot after every pair of DMMAs.  Samples beyond num_samples compute on row 0 and are never written out.
  // PARN is the buffer parity of the panel being produced (compile time: every shared-memory address is a
  // per-thread base plus an immediate).
  double ca = 0.0, ca2 = 0.0, cd = 1.0, cr = 1.0, cer = 0.0, cres = 0.0, ct = 0.0;  // the chain in flight
  double c_om = 0.0, c_v = 1.0, c_mu = 0.0, c_y = 0.0;
  const double* raw_t = s_RAW + ps0 * LK_KC + pl;   // this thread's element 0 in RAW buffer 0
  double* wg_t = s_WG + ps0 * LK_WSTRIDE + pl;      // ... and in W of buffer 0
  auto scalar_slot = [&](const int slot, const int pn, auto parn_tag) {
    constexpr int PARN = decltype(parn_tag)::value;
    const int e = slot / 12, op = slot % 12;
    const double* raw0 = raw_t + PARN * LK_RAW_DOUBLES + e * LK_PSTEP * LK_KC;
    const double* pix = s_RAW + PARN * LK_RAW_DOUBLES + 2 * LK_TS * LK_KC + pl;
    double* Wn = wg_t + PARN * 2 * LK_WG_DOUBLES + e * LK_PSTEP * LK_WSTRIDE;
    switch (op) {
      case 0: {
        double a = raw0[0];
        if (two_rows) a = mul_pinned(a, raw0[LK_TS * LK_KC]);
        const int p = pn * LK_KC + pl;
        if (num_rows > 2) {  // rare: dla_sample_log_likelihoods with more than two absorbers per sample
          a = times_extra_factors(a, sp.cache, s_rows + ps0 + LK_PSTEP * e, num_rows, sp.ld, min(p, n - 1));
          if (p >= n) a = 0.0;
        }
        if (sp.prod_out && p < n && ((live >> e) & 1u)) sp.prod_out[(size_t)(tile_s0 + ps0 + LK_PSTEP * e) * sp.ld + p] = a;
        ca = a;
        break;
      }
      case 1: ca2 = mul_pinned(ca, ca); break;
      case 2:
        c_om = pix[2 * LK_KC];
        c_v = pix[3 * LK_KC];
        cd = fma_pinned(c_om, ca2, c_v);  // dla_omega2 + v
        break;
      case 3:
        c_y = pix[0];
        c_mu = pix[LK_KC];
        cres = fnma_pinned(c_mu, ca, c_y);  // y - dla_mu
        cr = rcp_seed(cd);
        break;
      // 1 / d = r0 / (1 - e) = r0 (1 + e + e^2 + ...), e = 1 - d r0 <= 2^-20: the cubic term is below 2^-60
      case 4: cer = fnma_pinned(cd, cr, 1.0); break;
      case 5: cer = fma_pinned(cer, cer, cer); break;
      case 6: {
        cr = fma_pinned(cr, cer, cr);
        // see fast_rcp: outside roughly [1e-290, 1e290] (and for 0, inf, NaN, negative d) keep the seed.  The test
        // is on the exponent field with integer instructions - a DSETP would queue on the FP64 pipe like a DFMA
        const unsigned hi = (unsigned)__double2hiint(cd);
        if (!(hi - 0x03d00000u < 0x7c300000u - 0x03d00000u)) cr = rcp_seed(cd);
        break;
      }
      case 7: dprod[e] = mul_pinned(dprod[e], cd); break;
      case 8: Wn[0] = mul_pinned(ca2, cr); break;
      case 9: ct = mul_pinned(cres, cr); break;
      case 10: Wn[LK_WG_DOUBLES] = mul_pinned(ca, ct); break;
      case 11: q_acc[e] = fma_pinned(cres, ct, q_acc[e]); break;
      default: break;
    }
  };
  constexpr int LK_SLOTS = 12 * LK_EPT;  // 24

  // ---- MMA role: warp (rq, cq) owns samples 16 rq .. 16 rq + 15 and 7 or 8 column blocks -----------
  //   rq even : blocks [0,8) [8,15) [15,23) [23,30)      rq odd : blocks [0,7) [7,15) [15,22) [22,30)
  // Blocks >= 27 are the projection part (A = G).  Local blocks 0..3 always take A = W, local block 4
  // and local blocks 5..7 read A through offsets (W or G), so all warps run one instruction stream
  // with no predicated DMMA (a predicated-off DMMA still occupies its pipe slot); the eighth block
  // sits behind a warp-uniform branch.
  const int grp = lane >> 2;      // DMMA groupID
  const int tig = lane & 3;       // DMMA threadID_in_group
  const int rq = warp >> 2, cq = warp & 3;
  const bool odd = (rq & 1) != 0;
  const int first = !odd ? (cq == 0 ? 0 : cq == 1 ? 8 : cq == 2 ? 15 : 23) : (cq == 0 ? 0 : cq == 1 ? 7 : cq == 2 ? 15 : 22);
  const bool has_eighth = ((cq + (odd ? 1 : 0)) & 1) == 0;
  const int stage_rank = (rq * 2 + (cq >> 1)) * 32 + lane;  // rank among the threads of the 7-block warps
  const double* arow = s_WG + (rq * 16 + grp) * LK_WSTRIDE + tig;                     // A fragment base, buffer 0
  const double* a4row = arow + (first + 4 >= LK_NBLK_PAIR ? LK_WG_DOUBLES : 0);        // W or G
  const double* a5row = arow + (first + 5 >= LK_NBLK_PAIR ? LK_WG_DOUBLES : 0);
  const double* brow = s_main + tig * LK_PSTRIDE + first * 8 + grp;                    // B fragment base, stage 0
  double acc[LK_MB][LK_NB_MAX][2];
#pragma unroll
  for (int m = 0; m < LK_MB; ++m)
#pragma unroll
    for (int nb = 0; nb < LK_NB_MAX; ++nb) acc[m][nb][0] = acc[m][nb][1] = 0.0;

  // One panel of parity PAR: its DMMAs with, when PRODUCE, the scalar slots of panel + 1 threaded through them
  // (28 steps of two DMMAs in the seven common blocks x four k-steps, 24 slots); STAGE: the 7-block warps
  // request the inputs of panel + 2.
  auto run_panel = [&](const int panel, auto par_tag, auto produce_tag, auto stage_tag) {
    constexpr int PAR = decltype(par_tag)::value;
    constexpr bool PRODUCE = decltype(produce_tag)::value, STAGE = decltype(stage_tag)::value;
    static_assert(LK_PSTAGES == 2, "basis stage == panel parity");
    mbar_wait(bar_full, PAR);                                     // W/G of this panel written, inputs of the next landed
    mbar_wait(bar_base + 8u * PAR, (panel >> 1) & 1);             // basis panel landed
    constexpr int WOFF = PAR * 2 * LK_WG_DOUBLES, POFF = PAR * LK_PANEL_DOUBLES;
#pragma unroll
    for (int j = 0; j < 7 * (LK_KC / 4); ++j) {
      const int kb = j / 7, nb = j % 7;
      // The inputs of panel + 2 go into the RAW buffer that held those of panel: every warp read them in slots 0-15
      // of panel - 1 and has passed slot 22 (`full`), so the buffer is free; the requests go out after the first
      // DMMAs, so that the pipe already has work queued, and have most of the panel to land.
      if (STAGE && j == 1 && !has_eighth) stage_panel(panel + 2, par_tag, stage_rank, (LK_WARPS / 2) * 32);
      if (j == 7) {
        // From here on this panel overwrites what panel - 1 read to its end: the other W/G buffer (first store in
        // slot 8) and the other basis stage (TMA).
        if (panel > 0) mbar_wait(bar_empty, 1 - PAR);
        if (tid == 0 && PRODUCE) issue_panel(panel + 1);
      }
      const double* ar = nb < 4 ? arow : nb == 4 ? a4row : a5row;
      const double b = brow[POFF + kb * 4 * LK_PSTRIDE + nb * 8];
#pragma unroll
      for (int m = 0; m < LK_MB; ++m)
        dmma884_pinned(acc[m][nb][0], acc[m][nb][1], ar[WOFF + m * 8 * LK_WSTRIDE + kb * 4], b);
      if (PRODUCE && j < LK_SLOTS) scalar_slot(j, panel + 1, std::integral_constant<int, 1 - PAR>{});
      if (PRODUCE && j == LK_SLOTS - 2) {  // the warp's last W/G store is behind it: publish (and its staged inputs)
        cp_async_wait_all();
        warp_arrive(bar_full);
      }
    }
    if (has_eighth) {
#pragma unroll
      for (int kb = 0; kb < LK_KC / 4; ++kb) {
        const double b = brow[POFF + kb * 4 * LK_PSTRIDE + 7 * 8];
#pragma unroll
        for (int m = 0; m < LK_MB; ++m)
          dmma884_pinned(acc[m][7][0], acc[m][7][1], a5row[WOFF + m * 8 * LK_WSTRIDE + kb * 4], b);
      }
    }
    warp_arrive(bar_empty);                                       // this warp is done reading panel's W/G, basis stage, inputs
  };
  using P0 = std::integral_constant<int, 0>;
  using P1 = std::integral_constant<int, 1>;

  // ---- prologue: inputs of panels 0 and 1, W/G of panel 0 ---------------------------------------------------
  stage_panel(0, P0{}, tid, LK_THREADS);
  if (npanels > 1) stage_panel(1, P1{}, tid, LK_THREADS);
  cp_async_wait_all();
  __syncthreads();
#pragma unroll
  for (int slot = 0; slot < LK_SLOTS; ++slot) scalar_slot(slot, 0, P0{});
  warp_arrive(bar_full);

  // ---- main loop: two panels per trip (buffer parities are compile-time) ------------------------------------------
  // panel p: DMMAs read W/G[p & 1] and basis stage p & 1; the scalar slots read RAW[(p + 1) & 1] and write
  // W/G[(p + 1) & 1]; cp.async fills RAW[p & 1] with the inputs of p + 2; TMA fills the other basis stage
  // with p + 1.  The `full` and `empty` mbarriers order all of it between the warps.
  {
    int panel = 0;
    for (; panel + 3 < npanels; panel += 2) {   // both panels have successors to produce and to stage
      run_panel(panel, P0{}, std::true_type{}, std::true_type{});
      run_panel(panel + 1, P1{}, std::true_type{}, std::true_type{});
      if (panel & 2) renorm();
    }
    renorm();
    const int left = npanels - panel;           // 1, 2 or 3
    if (left == 3) {
      run_panel(panel, P0{}, std::true_type{}, std::true_type{});
      run_panel(panel + 1, P1{}, std::true_type{}, std::false_type{});
      run_panel(panel + 2, P0{}, std::false_type{}, std::false_type{});
    } else if (left == 2) {
      run_panel(panel, P0{}, std::true_type{}, std::false_type{});
      run_panel(panel + 1, P1{}, std::false_type{}, std::false_type{});
    } else {
      run_panel(panel, P0{}, std::false_type{}, std::false_type{});
    }
  }

  // per-sample scalar sums: reduce over the pixel lanes of the panel
  renorm();
#pragma unroll
  for (int e = 0; e < LK_EPT; ++e) {
    // sum log d = ln2 * (sum of the exponents) + log(product of the renormalised mantissa products): the 16 lanes'
    // products (each in [1, 2)) are multiplied here and the one logarithm per sample is taken together with the
    // pivots' in the Cholesky phase (a log per lane was 80 FP64 instructions per thread in a phase that crawls
    // next to the other CTA's DMMAs).  A non-positive or NaN product poisons the sample like log() would.
    double q = q_acc[e];
    double dp = dprod[e] > 0.0 ? dprod[e] : __longlong_as_double(0x7ff8000000000000ll);
    int es = esum[e];
#pragma unroll
    for (int off = (LK_KC < 32 ? LK_KC : 32) / 2; off > 0; off >>= 1) {
      q += __shfl_xor_sync(0xffffffffu, q, off);
      dp *= __shfl_xor_sync(0xffffffffu, dp, off);
      es += __shfl_xor_sync(0xffffffffu, es, off);
    }
    if (pl == 0) {
      s_sums[(ps0 + LK_PSTEP * e) * 3] = q;
      s_sums[(ps0 + LK_PSTEP * e) * 3 + 1] = dp;
      s_sums[(ps0 + LK_PSTEP * e) * 3 + 2] = (double)es;
    }
  }

  // ---- accumulators -> E[col][sample] (overlays the ring once every warp has left the loop) -------------
  __syncthreads();
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
  // indices; the entries of a column that other threads need are broadcast by shuffles.  (The first version
  // walked the matrix in shared memory with two loads per FMA and took 13 % of the kernel with the tensor pipe idle.)
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
    // Right-looking elimination: column j is scaled by 1/sqrt(pivot) and the trailing entries are updated at once,
    // a_ik -= l_ij l_kj (j < k <= i).  The updates of a column are independent of each other, so the dependent
    // chain per column is pivot broadcast -> rsqrt -> scale -> broadcast -> one FMA (the left-looking form, which
    // collected a_ij - sum_k l_ik l_jk when it reached column j, had FMA chains of length j/2 in front of every
    // pivot; this phase is latency-bound, it runs next to the other CTA's DMMAs).  Entries above the diagonal of a
    // thread's rows are never read for a result; updating them costs nothing extra in SIMT.
    // product of the pivots with its binary exponent pulled out every four pivots (integer pipe): 20 pivots of an
    // unnormalised spectrum (v ~ 1e-30: pivots ~ 1e31) would overflow a plain product and turn a finite
    // log-likelihood into -inf; a non-positive or non-finite pivot is left alone and poisons the logarithm below
    double piv_prod = 1.0;
    int piv_exp = 0;
#pragma unroll
    for (int j = 0; j < LK_K; ++j) {
      const double piv = __shfl_sync(0xffffffffu, j < 8 ? r0[j < 8 ? j : 0] : j < 16 ? r1[j < 16 ? j : 0] : r2[j], gbase + (j & 7));
      piv_prod *= piv;
      if ((j & 3) == 3) {
        const int hi = __double2hiint(piv_prod);
        const int ex = (hi >> 20) & 0x7ff;
        if (hi > 0 && ex != 0 && ex != 0x7ff) {
          piv_exp += ex - 1023;
          piv_prod = __hiloint2double(hi - ((ex - 1023) << 20), __double2loint(piv_prod));
        }
      }
      const double inv = fast_rsqrt(piv);
      if (j < 8) r0[j < 8 ? j : 0] *= inv;  // rows below j: column j of L (row 20: z_j)
      if (j < 16) r1[j < 16 ? j : 0] *= inv;
      r2[j] *= inv;
#pragma unroll
      for (int k = j + 1; k < LK_K; ++k) {
        // l_kj from the owner of row k
        const double lkj = __shfl_sync(0xffffffffu, k < 8 ? r0[j < 8 ? j : 0] : k < 16 ? r1[j < 16 ? j : 0] : r2[j], gbase + (k & 7));
        if (k < 8) r0[k < 8 ? k : 0] = fma(-r0[j < 8 ? j : 0], lkj, r0[k < 8 ? k : 0]);
        if (k < 16) r1[k < 16 ? k : 0] = fma(-r1[j < 16 ? j : 0], lkj, r1[k < 16 ? k : 0]);
        r2[k] = fma(-r2[j], lkj, r2[k]);
      }
    }
    // z'z on the owner of row 20 (t = 4, third slot)
    if (t == 4 && tile_s0 + s < sp.num_samples) {
      double zz0 = 0.0, zz1 = 0.0;
#pragma unroll
      for (int k = 0; k < LK_K; k += 2) { zz0 = fma(r2[k], r2[k], zz0); zz1 = fma(r2[k + 1], r2[k + 1], zz1); }
      const double quad = s_sums[s * 3] - (zz0 + zz1);
      // sum log d + 2 sum log L_ii.  A zero, negative or NaN pivot has already poisoned the factor through fast_rsqrt
      // (inf / NaN), so quad is NaN and the sample with it: no +inf can come out of log(0) here
      const double log_det = fma(s_sums[s * 3 + 2] + (double)piv_exp, LK_LN2, log(s_sums[s * 3 + 1] * piv_prod));
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
