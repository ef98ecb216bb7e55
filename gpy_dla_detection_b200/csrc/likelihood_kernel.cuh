// likelihood_kernel.cuh : batched low-rank Gaussian log-likelihoods of absorber samples
// (SURVEY.md §8 a5, a7, a8).
//
// Reference: for every QMC sample the Python loop builds dla_mu = mu a, dla_M = M a,
// d = omega2 a^2 + v (dla_gp.py:388-394) and calls NullGP.log_mvnpdf_low_rank
// (null_gp.py:307-360): B = I + M'D^-1 M (dgemm), chol(B), quad and log-det.
//
// Here all samples of a tile share the interpolated model, so with
//     w_sp = a_sp^2 / d_sp ,  g_sp = a_sp r_sp / d_sp ,  r_sp = y_p - mu_p a_sp
// the per-sample Gram matrices and projections are ONE dense FP64 contraction
//     [B_s - I | c_s] = [W | G] x [P | M] ,   P[p,(i,j)] = m_pi m_pj  (i >= j, 210 pairs)
// over the pixel axis, run on the FP64 tensor path (DMMA m8n8k4; B200 measures 36.9 TFLOP/s,
// same as the DFMA pipe, at a fraction of the issue slots).  W/G tiles are produced on the
// fly from the profile cache (product of up to 4 absorber rows, dla_gp.py:370-386), P is
// formed in registers from the staged M tile, and the k x k Cholesky factor / solve /
// log-det of every sample of the tile runs in shared memory straight out of the accumulator
// fragments.  HBM sees only the profile rows (read) and one double per sample (written).
#pragma once
#include <stdint.h>
#include <type_traits>

namespace dla {

constexpr int LK_K = 20;                        // rank of the learned covariance (Parameters.k)
constexpr int LK_PAIRS = LK_K * (LK_K + 1) / 2;  // 210 lower-triangle pairs
constexpr int LK_TS = 32;                        // samples per CTA tile (two CTAs share an SM)
constexpr int LK_KC = 32;                        // pixels per chunk
constexpr int LK_WSTRIDE = LK_KC + 4;            // padded row stride of the W/G tiles (conflict-free DMMA A loads)
constexpr int LK_MSTRIDE = LK_K;                 // row stride of the M tile (20 == 4 mod 16: conflict-free B loads)
constexpr int LK_THREADS = 256;
constexpr int LK_NBLK_PAIR = 27;                 // ceil(210 / 8) column blocks of the Gram part
constexpr int LK_NBLK = 30;                      // + 3 column blocks (24 >= 20) of the projection part
constexpr int LK_EP_COLS = 216 + 24;             // columns staged for the epilogue
constexpr int LK_EP_STRIDE = LK_TS + 1;          // epilogue smem: [col][sample], stride 33
constexpr int LK_MAX_ROWS = 8;                   // max absorbers multiplied per sample (max_dlas <= 8)
constexpr double LK_LOG_2PI = 1.83787706640934534;  // null_gp.py:325

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

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// cp.async (LDGSTS): global -> shared without a register round trip
__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gmem_src) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async16_zfill(void* smem_dst, const void* gmem_src, int src_bytes) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;\n" ::"r"(d), "l"(gmem_src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }

// ---- shared memory plan (per CTA; two CTAs are resident per SM) -----------------------------------
//   WG   : 2 stages x { W[32][36], G[32][36] }                       36 864 B
//   MR   : 3-deep ring of M tiles [32][20]                            15 360 B
//   STG  : thread-private staging slots filled by cp.async one chunk ahead:
//          2 factor rows x 4 samples + y, mu, omega2, v  -> [12][256] 24 576 B
//   E    : epilogue matrix [240][33], overlays WG + MR + STG          63 360 B
//   AUX  : per-sample sums [32][2], profile rows [8][32]
constexpr int LK_STG_ROWS = 2;                                   // factors staged by cp.async (others load directly)
constexpr int LK_EPT = LK_TS * LK_KC / LK_THREADS;               // elements per thread per chunk (4)
constexpr int LK_WG_DOUBLES = 2 * 2 * LK_TS * LK_WSTRIDE;        // 4608
constexpr int LK_MR_DOUBLES = 3 * LK_KC * LK_MSTRIDE;            // 1920
constexpr int LK_STG_SLOTS = LK_STG_ROWS * LK_EPT + 4;           // 12
constexpr int LK_STG_DOUBLES = LK_STG_SLOTS * LK_THREADS;        // 3072
constexpr int LK_MAIN_DOUBLES_RAW = LK_WG_DOUBLES + LK_MR_DOUBLES + LK_STG_DOUBLES;
constexpr int LK_EP_DOUBLES = LK_EP_COLS * LK_EP_STRIDE;
constexpr int LK_MAIN_DOUBLES = LK_MAIN_DOUBLES_RAW > LK_EP_DOUBLES ? LK_MAIN_DOUBLES_RAW : LK_EP_DOUBLES;
constexpr size_t LK_AUX_BYTES = LK_TS * 2 * sizeof(double) + LK_MAX_ROWS * LK_TS * sizeof(int32_t);
constexpr size_t LK_SMEM_BYTES = (size_t)LK_MAIN_DOUBLES * sizeof(double) + LK_AUX_BYTES;
constexpr int LK_NWARPS = LK_THREADS / 32;                       // 8
constexpr int LK_MB = LK_TS / 8;                                 // DMMA row blocks per warp tile (4)
constexpr int LK_NB = 4;                                         // DMMA column blocks per warp (8 warps x 4 >= 30)

// grid = (ceil(max num_samples / 32), num_spectra), block = 256, dynamic smem = LK_SMEM_BYTES
__global__ void __launch_bounds__(LK_THREADS, 2)
sample_likelihood_kernel(const LikelihoodSpectrum* __restrict__ specs) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const LikelihoodSpectrum sp = specs[blockIdx.y];
  const int tile_s0 = blockIdx.x * LK_TS;
  if (tile_s0 >= sp.num_samples) return;
  if (sp.alive && *sp.alive == 0) return;

  double* s_main = reinterpret_cast<double*>(smem_raw);
  double* s_wg = s_main;
  double* s_mr = s_wg + LK_WG_DOUBLES;
  double* s_stg = s_mr + LK_MR_DOUBLES;
  double* s_sums = s_main + LK_MAIN_DOUBLES;                          // [32][2] : sum r^2/d, sum log d
  int32_t* s_rows = reinterpret_cast<int32_t*>(s_sums + LK_TS * 2);   // [num_rows][32]

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int n = sp.n;
  const int nchunks = (n + LK_KC - 1) / LK_KC;
  const int num_rows = sp.num_rows;

  // ---- profile rows of the tile's samples, zeroed sums -----------------------------------
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
  if (tid < LK_TS * 2) s_sums[tid] = 0.0;
  __syncthreads();

  // ---- MMA roles: warp w owns column blocks [4w, 4w+4) for all 32 samples -----------------
  //      blocks 0..26 = Gram pairs (A = W), blocks 27..29 = projection (A = G), 30,31 do not exist
  const int grp = lane >> 2;      // DMMA groupID
  const int tig = lane & 3;       // DMMA threadID_in_group
  int col_i[LK_NB], col_j[LK_NB];
#pragma unroll
  for (int nb = 0; nb < LK_NB; ++nb) {
    const int blk = warp * LK_NB + nb;
    if (blk < LK_NBLK_PAIR) {
      col_i[nb] = c_pair_i[blk * 8 + grp];
      col_j[nb] = c_pair_j[blk * 8 + grp];
    } else {
      const int j = (blk - LK_NBLK_PAIR) * 8 + grp;  // projection column (pads clamp to 0)
      col_i[nb] = -1;
      col_j[nb] = j < LK_K ? j : 0;
    }
  }
  double acc[LK_MB][LK_NB][2];
#pragma unroll
  for (int m = 0; m < LK_MB; ++m)
#pragma unroll
    for (int nb = 0; nb < LK_NB; ++nb) acc[m][nb][0] = acc[m][nb][1] = 0.0;

  // ---- producer roles: warp w owns samples w, w+8, w+16, w+24; lane = pixel within the chunk ----
  double q_acc[LK_EPT], dprod[LK_EPT];
#pragma unroll
  for (int e = 0; e < LK_EPT; ++e) { q_acc[e] = 0.0; dprod[e] = 1.0; }
  double* my_stg = s_stg + tid;  // slot k of this thread at my_stg[k * 256]

  // fold the running products of d into the per-sample log-determinant sums (only this warp touches its samples)
  auto fold_logdet = [&]() {
#pragma unroll
    for (int e = 0; e < LK_EPT; ++e) {
      double l = log(dprod[e]);
      dprod[e] = 1.0;
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) l += __shfl_xor_sync(0xffffffffu, l, off);
      if (lane == 0) s_sums[(warp + LK_NWARPS * e) * 2 + 1] += l;
    }
  };

  // issue the global->shared copies of one chunk: factor rows (first two), pixel scalars, M tile
  auto prefetch = [&](int chunk) {
    const int p = chunk * LK_KC + lane;
    if (p < n) {
#pragma unroll
      for (int e = 0; e < LK_EPT; ++e) {
        const int s = warp + LK_NWARPS * e;
        if (tile_s0 + s < sp.num_samples) {
          cp_async8(my_stg + e * LK_THREADS, sp.base0 + (size_t)s_rows[s] * sp.ld + p);
          if (num_rows > 1)
            cp_async8(my_stg + (LK_EPT + e) * LK_THREADS, sp.cache + (size_t)s_rows[LK_TS + s] * sp.ld + p);
        }
      }
      cp_async8(my_stg + (2 * LK_EPT + 0) * LK_THREADS, sp.y + p);
      cp_async8(my_stg + (2 * LK_EPT + 1) * LK_THREADS, sp.mu + p);
      cp_async8(my_stg + (2 * LK_EPT + 2) * LK_THREADS, sp.omega2 + p);
      cp_async8(my_stg + (2 * LK_EPT + 3) * LK_THREADS, sp.v + p);
    }
    // M tile: 640 doubles = 320 x 16 B, zero-filled beyond pixel n
    double* Mt = s_mr + (chunk % 3) * (LK_KC * LK_MSTRIDE);
    for (int e2 = tid; e2 < LK_KC * LK_K / 2; e2 += LK_THREADS) {
      const int pp = chunk * LK_KC + (2 * e2) / LK_K;
      // clamp the source so the address is always valid; src_bytes = 0 zero-fills
      const double* src = sp.M + (size_t)(pp < n ? (chunk * LK_KC * LK_K + 2 * e2) : 0);
      cp_async16_zfill(Mt + 2 * e2, src, pp < n ? 16 : 0);
    }
    cp_async_commit();
  };

  // turn the staged values of one chunk into the W / G operand tiles
  auto produce = [&](int chunk, int buf) {
    double* Ws = s_wg + buf * (2 * LK_TS * LK_WSTRIDE);
    double* Gs = Ws + LK_TS * LK_WSTRIDE;
    const int p = chunk * LK_KC + lane;
    const bool pv = p < n;
    double yp = 0, mup = 0, omp = 0, vp = 1;
    if (pv) {
      yp = my_stg[(2 * LK_EPT + 0) * LK_THREADS]; mup = my_stg[(2 * LK_EPT + 1) * LK_THREADS];
      omp = my_stg[(2 * LK_EPT + 2) * LK_THREADS]; vp = my_stg[(2 * LK_EPT + 3) * LK_THREADS];
    }
#pragma unroll
    for (int e = 0; e < LK_EPT; ++e) {
      const int s = warp + LK_NWARPS * e;
      double w = 0.0, g = 0.0;
      if (pv && tile_s0 + s < sp.num_samples) {
        // absorption = product of the factors' profiles, left to right (dla_gp.py:370-386)
        double a = my_stg[e * LK_THREADS];
        if (num_rows > 1) a = a * my_stg[(LK_EPT + e) * LK_THREADS];
        for (int r = 2; r < num_rows; ++r) a = a * sp.cache[(size_t)s_rows[r * LK_TS + s] * sp.ld + p];
        if (sp.prod_out) sp.prod_out[(size_t)(tile_s0 + s) * sp.ld + p] = a;
        const double a2 = a * a;
        const double d = fma(omp, a2, vp);   // dla_omega2 + v
        const double inv = 1.0 / d;
        const double r = fma(-mup, a, yp);   // y - dla_mu
        w = a2 * inv;
        g = a * r * inv;
        q_acc[e] = fma(r * r, inv, q_acc[e]);
        dprod[e] *= d;
      }
      Ws[s * LK_WSTRIDE + lane] = w;
      Gs[s * LK_WSTRIDE + lane] = g;
    }
    if ((chunk & 7) == 7) fold_logdet();  // keeps the product of d within range
  };

  // NPAIR Gram blocks followed by NPROJ projection blocks, both compile-time: the three warp
  // classes (0..5: 4+0, 6: 3+1, 7: 0+2) get separate loop bodies - a predicated-off DMMA still
  // costs its tensor-pipe slot.
  auto consume_kind = [&](int chunk, int buf, auto npair_c, auto nproj_c) {
    constexpr int NPAIR = decltype(npair_c)::value;
    constexpr int NPROJ = decltype(nproj_c)::value;
    const double* Ws = s_wg + buf * (2 * LK_TS * LK_WSTRIDE);
    const double* Gs = Ws + LK_TS * LK_WSTRIDE;
    const double* Ms = s_mr + (chunk % 3) * (LK_KC * LK_MSTRIDE);
    const int arow = grp * LK_WSTRIDE + tig;
#pragma unroll
    for (int kb = 0; kb < LK_KC / 4; ++kb) {
      const double* mrow = Ms + (kb * 4 + tig) * LK_MSTRIDE;
      if (NPAIR > 0) {
        double aw[LK_MB];
#pragma unroll
        for (int m = 0; m < LK_MB; ++m) aw[m] = Ws[arow + m * 8 * LK_WSTRIDE + kb * 4];
#pragma unroll
        for (int nb = 0; nb < NPAIR; ++nb) {
          const double b = mrow[col_i[nb]] * mrow[col_j[nb]];
#pragma unroll
          for (int m = 0; m < LK_MB; ++m) dmma884(acc[m][nb][0], acc[m][nb][1], aw[m], b);
        }
      }
      if (NPROJ > 0) {
        double ag[LK_MB];
#pragma unroll
        for (int m = 0; m < LK_MB; ++m) ag[m] = Gs[arow + m * 8 * LK_WSTRIDE + kb * 4];
#pragma unroll
        for (int nb = NPAIR; nb < NPAIR + NPROJ; ++nb) {
          const double b = mrow[col_j[nb]];
#pragma unroll
          for (int m = 0; m < LK_MB; ++m) dmma884(acc[m][nb][0], acc[m][nb][1], ag[m], b);
        }
      }
    }
  };
  auto consume = [&](int chunk, int buf) {
    using std::integral_constant;
    if (warp < 6) consume_kind(chunk, buf, integral_constant<int, 4>(), integral_constant<int, 0>());
    else if (warp == 6) consume_kind(chunk, buf, integral_constant<int, 3>(), integral_constant<int, 1>());
    else consume_kind(chunk, buf, integral_constant<int, 0>(), integral_constant<int, 2>());
  };

  // ---- main loop: chunk c+2 in flight (cp.async), chunk c+1 being produced, chunk c in the MMAs ----
  prefetch(0);
  cp_async_wait_all();
  produce(0, 0);          // reads only this thread's own staging slots
  if (nchunks > 1) prefetch(1);
  __syncthreads();        // W/G(0) and M(0) visible to all warps
  for (int chunk = 0; chunk < nchunks; ++chunk) {
    const int buf = chunk & 1;
    if (chunk + 1 < nchunks) {
      cp_async_wait_all();             // staged values + M tile of chunk+1 have landed (issued one MMA phase ago)
      produce(chunk + 1, buf ^ 1);
      if (chunk + 2 < nchunks) prefetch(chunk + 2);
    }
    consume(chunk, buf);
    __syncthreads();
  }

  // ---- per-sample scalar sums: reduce over the 32 pixel lanes -----------------------------
  fold_logdet();
#pragma unroll
  for (int e = 0; e < LK_EPT; ++e) {
    double q = q_acc[e];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) q += __shfl_xor_sync(0xffffffffu, q, off);
    if (lane == 0) s_sums[(warp + LK_NWARPS * e) * 2] = q;
  }

  // ---- accumulators -> E[col][sample] (overlays the stage buffers; all reads finished) -----
  double* E = s_main;
#pragma unroll
  for (int m = 0; m < LK_MB; ++m) {
    const int s = m * 8 + grp;
#pragma unroll
    for (int nb = 0; nb < LK_NB; ++nb) {
      const int blk = warp * LK_NB + nb;
      if (blk < LK_NBLK) {
        const int col = blk * 8 + tig * 2;
        E[col * LK_EP_STRIDE + s] = acc[m][nb][0];
        E[(col + 1) * LK_EP_STRIDE + s] = acc[m][nb][1];
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
      const int col = (i < LK_K) ? (i * (i + 1) / 2 + j) : (216 + j);
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
