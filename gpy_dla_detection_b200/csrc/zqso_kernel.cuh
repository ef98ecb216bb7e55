// zqso_kernel.cuh : quasar-redshift estimation, ZGP (BASELINE.json configs[4]).
//
// Reference: ZGP.inference_z_qso (zqso_gp.py:214-250) loops over 10 000 candidate redshifts and for
// each one runs set_data (:92-182: observed-frame window, nanmedian normalisation over rest
// 1176-1256 A, bluewards / redwards splits, mask + range filter), get_interp (:66-90: linear
// interpolation of mu and the 20 columns of M at x = lambda / (1 + z)) and log_model_evidence
// (:184-212: low-rank Gaussian inside the window + two i.i.d. Gaussians outside).
//
// Unlike the DLA path the basis changes with the sample (M is re-interpolated at every z), so there
// is no shared Gram basis: ONE WARP OWNS ONE (spectrum, z) SAMPLE.  The bordered matrix
//     [[B - I, c], [c', q]] = sum_p (1/v_p) mt_p mt_p' ,   mt_p = [m_p (20) ; r_p ; 0 0 0]  (24)
// is accumulated on the FP64 tensor path as the 6 lower 8x8 blocks of a 24 x 24 product: each lane
// interpolates exactly the three entries of mt it needs for its fragment (no redundancy across the
// warp: 96 values per 4 pixels = 24 per pixel), scales them by 1/v for the A operand and issues
// 6 DMMAs per 4 pixels.  Per-pixel scalars (x, grid index, normalised flux, 1/v, residual) are
// computed once by the lane that owns the pixel of a 32-pixel chunk and broadcast by shuffles.
// Masked and out-of-range pixels stay in the stream with weight 0, so nothing is compacted.
// The Cholesky of the bordered matrix, the i.i.d. sums and the median (bitonic sort of the <= 512
// pixels of the normalisation window) all run inside the warp.
#pragma once
#include <math.h>
#include <stdint.h>

namespace dla {

constexpr int ZQ_K = 20;           // rank of the learned covariance
constexpr int ZQ_STRIDE = 24;      // padded row stride of the M / slope tables (3 DMMA column blocks)
constexpr int ZQ_WARPS = 8;        // samples per CTA
constexpr int ZQ_TDIM = 24;        // bordered matrix is 21 x 21 inside 24 x 24
constexpr int ZQ_TSTRIDE = 25;
constexpr double ZQ_LOG_2PI = 1.83787706640934534;  // zqso_gp.py:263
constexpr double ZQ_LN2 = 0.693147180559945309417232121458;

struct ZqsoModelDev {
  const double* rest;       // n_rest grid
  const double* mu;         // n_rest
  const double* mu_slope;   // n_rest - 1 : (mu[i+1] - mu[i]) / (rest[i+1] - rest[i])
  const double* M;          // n_rest x 24 (columns >= 20 are zero)
  const double* M_slope;    // (n_rest - 1) x 24
  const double* MS;         // n_rest x 24 x 2 : (value, slope) pairs for the v2 kernel: M, then -mu in column 20
  int n_rest;
  int uniform;              // rest[i] == rest[0] + i * dl exactly
  double rest0, inv_dl;
  double bluewards_mu, redwards_mu, bluewards_var, redwards_var;  // var = sigma^2
};

struct ZqsoParamsDev {
  double min_lambda, max_lambda, norm_min_lambda, norm_max_lambda;
};

struct ZqsoSpectrum {
  const double* X;      // observed wavelengths, strictly increasing
  const double* Y;      // flux
  const double* V;      // noise variance
  const uint8_t* mask;  // 1 = bad pixel
  int n_raw;
};

// slope tables (scipy interp1d: slope = (y_hi - y_lo) / (x_hi - x_lo), formed once per model)
__global__ void zqso_slopes_kernel(const double* rest, const double* mu, const double* M /* n x k */, int n_rest, int k,
                                   double* mu_slope, double* Mp /* n x 24 */, double* Ms /* (n-1) x 24 */) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_rest) return;
  for (int j = 0; j < ZQ_STRIDE; ++j) Mp[(size_t)i * ZQ_STRIDE + j] = j < k ? M[(size_t)i * k + j] : 0.0;
  if (i + 1 < n_rest) {
    const double dx = __dsub_rn(rest[i + 1], rest[i]);
    mu_slope[i] = __ddiv_rn(__dsub_rn(mu[i + 1], mu[i]), dx);
    for (int j = 0; j < ZQ_STRIDE; ++j)
      Ms[(size_t)i * ZQ_STRIDE + j] = j < k ? __ddiv_rn(__dsub_rn(M[(size_t)(i + 1) * k + j], M[(size_t)i * k + j]), dx) : 0.0;
  }
}

// index lo of the interpolation interval for x: scipy interp1d = searchsorted(rest, x, 'left') clipped to
// [1, n-1], minus one
__device__ __forceinline__ int zqso_interval(const ZqsoModelDev& m, double x) {
  int hi;
  if (m.uniform) {
    hi = (int)ceil((x - m.rest0) * m.inv_dl);
    hi = max(1, min(hi, m.n_rest - 1));
    while (hi > 1 && m.rest[hi - 1] >= x) --hi;
    while (hi < m.n_rest - 1 && m.rest[hi] < x) ++hi;
  } else {
    int a = 0, b = m.n_rest;  // first index with rest[idx] >= x
    while (a < b) {
      const int c = (a + b) >> 1;
      if (m.rest[c] < x) a = c + 1; else b = c;
    }
    hi = max(1, min(a, m.n_rest - 1));
  }
  return hi - 1;
}

__device__ __forceinline__ double zq_rcp(double d) {  // see likelihood_kernel.cuh : fast_rcp
  double r0;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r0) : "d"(d));
  double e = fma(-d, r0, 1.0);
  double r = fma(r0, e, r0);
  e = fma(-d, r, 1.0);
  r = fma(r, e, r);
  return (d > 1e-290 && d < 1e290) ? r : 1.0 / d;
}

// running product with the binary exponent pulled out (sum of logs with one log at the end)
struct LogProd {
  double prod = 1.0;
  int esum = 0;
  __device__ __forceinline__ void mul(double d) { prod *= d; }
  __device__ __forceinline__ void renorm() {
    const int hi = __double2hiint(prod);
    const int ex = (hi >> 20) & 0x7ff;
    if (hi > 0 && ex != 0 && ex != 0x7ff) {
      esum += ex - 1023;
      prod = __hiloint2double(hi - ((ex - 1023) << 20), __double2loint(prod));
    }
  }
  __device__ __forceinline__ double value() { renorm(); return fma((double)esum, ZQ_LN2, log(prod)); }
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  return v;
}

__device__ __forceinline__ void zq_dmma(double& c0, double& c1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
      : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// first index in [a, b) with X[idx] > t (strict = true) or X[idx] >= t (strict = false)
__device__ __forceinline__ int zq_bound(const double* X, int a, int b, double t, bool strict) {
  while (a < b) {
    const int c = (a + b) >> 1;
    const bool go_right = strict ? (X[c] <= t) : (X[c] < t);
    if (go_right) a = c + 1; else b = c;
  }
  return a;
}
// same on x = X / opz (monotone in the index), with the reference's own division
__device__ __forceinline__ int zq_bound_rest(const double* X, double opz, int a, int b, double t, bool strict) {
  while (a < b) {
    const int c = (a + b) >> 1;
    const double x = __ddiv_rn(X[c], opz);
    const bool go_right = strict ? (x <= t) : (x < t);
    if (go_right) a = c + 1; else b = c;
  }
  return a;
}

// grid = (ceil(S / 8), num_spectra), block = 256, dynamic smem = 8 * per_warp doubles,
// per_warp = max(norm_cap, 24 * 25); norm_cap = power of two >= pixels in any normalisation window
__global__ void __launch_bounds__(ZQ_WARPS * 32, 3)
zqso_likelihood_kernel(const ZqsoSpectrum* __restrict__ spectra, const double* __restrict__ z_samples, int S,
                       ZqsoModelDev model, ZqsoParamsDev prm, int norm_cap, int per_warp,
                       double* __restrict__ out /* [num_spectra][S] */) {
  extern __shared__ double zq_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int s = blockIdx.x * ZQ_WARPS + warp;
  if (s >= S) return;
  const ZqsoSpectrum sp = spectra[blockIdx.y];
  double* buf = zq_smem + (size_t)warp * per_warp;
  const double* X = sp.X;
  const int n_raw = sp.n_raw;
  const double z = z_samples[s];
  const double opz = __dadd_rn(1.0, z);
  double* out_ll = out + (size_t)blockIdx.y * S + s;

  // ---- observed-frame window (zqso_gp.py:123-136): strict inequalities on both sides ------------
  const double max_pos = __dmul_rn(prm.max_lambda, opz), min_pos = __dmul_rn(prm.min_lambda, opz);
  const double max_obs = fmin(max_pos, X[n_raw - 1]), min_obs = fmax(min_pos, X[0]);
  const int lo = zq_bound(X, 0, n_raw, min_obs, true);          // first X > min_obs
  const int hi_end = zq_bound(X, 0, n_raw, max_obs, false);     // first X >= max_obs : window = [lo, hi_end)
  const int bw_end = zq_bound(X, 0, n_raw, min_obs, false);     // X < min_obs  <=> index < bw_end
  const int rw_begin = zq_bound(X, 0, n_raw, max_obs, true);    // X > max_obs  <=> index >= rw_begin

  // ---- flux normalisation: nanmedian of the window pixels with rest 1176..1256 (mask ignored, :142-148)
  double med;
  {
    const int nlo = zq_bound_rest(X, opz, lo, max(hi_end, lo), prm.norm_min_lambda, false);  // first x >= min
    const int nhi = zq_bound_rest(X, opz, lo, max(hi_end, lo), prm.norm_max_lambda, true);   // first x > max
    const double inf = __longlong_as_double(0x7ff0000000000000LL);
    int count = 0;
    for (int i = lane; i < norm_cap; i += 32) {
      const int p = nlo + i;
      double val = inf;
      if (p < nhi) {
        const double yv = sp.Y[p];
        if (!isnan(yv)) { val = yv; ++count; }
      }
      buf[i] = val;
    }
    count = (int)(warp_sum((double)count) + 0.5);
    __syncwarp();
    // bitonic sort, ascending (+inf pads and NaNs-as-inf go to the end; a genuine +inf flux sorts there too)
    for (int k = 2; k <= norm_cap; k <<= 1) {
      for (int j = k >> 1; j > 0; j >>= 1) {
        for (int i = lane; i < norm_cap; i += 32) {
          const int l = i ^ j;
          if (l > i) {
            const double a = buf[i], b = buf[l];
            const bool up = (i & k) == 0;
            if ((a > b) == up) { buf[i] = b; buf[l] = a; }
          }
        }
        __syncwarp();
      }
    }
    if (count == 0) med = __longlong_as_double(0x7ff8000000000000LL);
    else if (count & 1) med = buf[count >> 1];
    else med = (buf[(count >> 1) - 1] + buf[count >> 1]) * 0.5;
    __syncwarp();
  }
  const double invmed = 1.0 / med;
  const double invmed2 = 1.0 / (med * med);

  // ---- window pixels: bordered Gram matrix on the tensor path ----------------------------------------
  const int grp = lane >> 2, tig = lane & 3;
  double acc[6][2];  // blocks (0,0) (1,0) (1,1) (2,0) (2,1) (2,2)
#pragma unroll
  for (int b = 0; b < 6; ++b) acc[b][0] = acc[b][1] = 0.0;
  LogProd vprod;
  int n_sel = 0;
  for (int c0 = lo; c0 < hi_end; c0 += 32) {
    // per-pixel scalars, one pixel per lane
    const int p = c0 + lane;
    const int pc = min(p, n_raw - 1);
    const double xp = __ddiv_rn(X[pc], opz);  // emitted_wavelengths (:139)
    const bool sel = p < hi_end && !sp.mask[pc] && xp >= prm.min_lambda && xp <= prm.max_lambda;  // :169-170
    const int iv = zqso_interval(model, fmin(fmax(xp, model.rest[0]), model.rest[model.n_rest - 1]));
    const double xoff = xp - model.rest[iv];
    const double vn = sp.V[pc] * invmed2;
    const double mu_p = fma(model.mu_slope[iv], xoff, model.mu[iv]);
    double r = fma(sp.Y[pc], invmed, -mu_p);
    double dinv = zq_rcp(vn);
    if (!sel) { r = 0.0; dinv = 0.0; }
    vprod.mul(sel ? vn : 1.0);
    n_sel += sel ? 1 : 0;
    if (((c0 - lo) & 255) == 224) vprod.renorm();
#pragma unroll
    for (int kb = 0; kb < 8; ++kb) {
      const int src = kb * 4 + tig;
      const int iv_k = __shfl_sync(0xffffffffu, iv, src);
      const double xo_k = __shfl_sync(0xffffffffu, xoff, src);
      const double di_k = __shfl_sync(0xffffffffu, dinv, src);
      const double r_k = __shfl_sync(0xffffffffu, r, src);
      const double* Mrow = model.M + (size_t)iv_k * ZQ_STRIDE + grp;
      const double* Srow = model.M_slope + (size_t)iv_k * ZQ_STRIDE + grp;
      const double m0 = fma(__ldg(Srow), xo_k, __ldg(Mrow));
      const double m1 = fma(__ldg(Srow + 8), xo_k, __ldg(Mrow + 8));
      double m2 = fma(__ldg(Srow + 16), xo_k, __ldg(Mrow + 16));  // columns 16..19, zero pads beyond
      if (grp == 4) m2 = r_k;                                     // column 20 carries the residual
      const double a0 = m0 * di_k, a1 = m1 * di_k, a2 = m2 * di_k;
      zq_dmma(acc[0][0], acc[0][1], a0, m0);
      zq_dmma(acc[1][0], acc[1][1], a1, m0);
      zq_dmma(acc[2][0], acc[2][1], a1, m1);
      zq_dmma(acc[3][0], acc[3][1], a2, m0);
      zq_dmma(acc[4][0], acc[4][1], a2, m1);
      zq_dmma(acc[5][0], acc[5][1], a2, m2);
    }
  }
  const double sum_log_v = warp_sum(vprod.value());
  const int n_in = (int)(warp_sum((double)n_sel) + 0.5);

  // ---- fragments -> T (24 x 24, lower blocks), Cholesky of the bordered 21 x 21 matrix -------------------
  double* T = buf;
  {
    const int bi_of[6] = {0, 1, 1, 2, 2, 2}, bj_of[6] = {0, 0, 1, 0, 1, 2};
#pragma unroll
    for (int b = 0; b < 6; ++b) {
      const int row = bi_of[b] * 8 + grp, col = bj_of[b] * 8 + tig * 2;
      T[row * ZQ_TSTRIDE + col] = acc[b][0];
      T[row * ZQ_TSTRIDE + col + 1] = acc[b][1];
    }
  }
  __syncwarp();
  if (lane < ZQ_K) T[lane * ZQ_TSTRIDE + lane] += 1.0;  // + I (null_gp.py:341)
  __syncwarp();
  double piv_prod = 1.0;
  for (int j = 0; j < ZQ_K; ++j) {
    const double piv = T[j * ZQ_TSTRIDE + j];
    piv_prod *= piv;
    const double inv = 1.0 / sqrt(piv);
    __syncwarp();
    if (lane > j && lane <= ZQ_K) T[lane * ZQ_TSTRIDE + j] *= inv;  // column j of L, rows j+1..20
    __syncwarp();
    // trailing update of the lower triangle: rows i in (j, 20], columns k in (j, i]
    const int m = ZQ_K - j;                 // remaining rows
    const int pairs = m * (m + 1) / 2;
    for (int e = lane; e < pairs; e += 32) {
      // e -> (ii, kk), 0 <= kk <= ii < m
      int ii = (int)((sqrt(8.0 * e + 1.0) - 1.0) * 0.5);
      while ((ii + 1) * (ii + 2) / 2 <= e) ++ii;
      while (ii * (ii + 1) / 2 > e) --ii;
      const int kk = e - ii * (ii + 1) / 2;
      const int i = j + 1 + ii, k = j + 1 + kk;
      T[i * ZQ_TSTRIDE + k] = fma(-T[i * ZQ_TSTRIDE + j], T[k * ZQ_TSTRIDE + j], T[i * ZQ_TSTRIDE + k]);
    }
    __syncwarp();
  }
  const double quad = T[ZQ_K * ZQ_TSTRIDE + ZQ_K];  // q - z'z
  const double ll_window = -0.5 * (quad + (sum_log_v + log(piv_prod)) + (double)n_in * ZQ_LOG_2PI);

  // ---- i.i.d. Gaussians bluewards and redwards of the window (:160-166, :198-210, :252-278) -----------------
  double side_ll[2];
#pragma unroll
  for (int side = 0; side < 2; ++side) {
    const int begin = side == 0 ? 0 : rw_begin, end = side == 0 ? bw_end : n_raw;
    const double m_side = side == 0 ? model.bluewards_mu : model.redwards_mu;
    const double var_side = side == 0 ? model.bluewards_var : model.redwards_var;
    double qs = 0.0;
    LogProd dprod;
    int cnt = 0, it = 0;
    for (int p = begin + lane; p < end; p += 32, ++it) {
      if (!sp.mask[p]) {
        const double t = fma(sp.Y[p], invmed, -m_side);
        const double dd = fma(sp.V[p], invmed2, var_side);
        qs = fma(t * t, zq_rcp(dd), qs);
        dprod.mul(dd);
        ++cnt;
      }
      if ((it & 7) == 7) dprod.renorm();
    }
    const double q_tot = warp_sum(qs), ld_tot = warp_sum(dprod.value());
    const int c_tot = (int)(warp_sum((double)cnt) + 0.5);
    side_ll[side] = -0.5 * (q_tot + ld_tot + (double)c_tot * ZQ_LOG_2PI);
  }
  if (lane == 0) *out_ll = ll_window + side_ll[0] + side_ll[1];
}

// =====================================================================================================
// Round-2 path for uniform model grids (the published 910:0.25:3000 grid): zqso_median_batch_kernel +
// zqso_likelihood_kernel_v2.  What changed against the kernel above, and why (profiles/zqso_kernel_r01_ncu.txt:
// tensor pipe 33 %, long-scoreboard 34 %, in-warp sort 10 %):
//   * the nanmedian normalisation is a separate pass: a CTA takes 32 CONSECUTIVE redshift samples, whose
//     normalisation windows overlap almost entirely (the window slides by about half a pixel per sample), sorts
//     the union of the windows ONCE (bitonic, 256 threads) and every warp then picks its samples' medians by
//     rank counting over the sorted union (ballot / popc) - instead of one 512-element bitonic sort per sample
//     inside the warp that also has to feed the tensor pipe;
//   * rest-frame coordinates without the division and the search: x = X (1/(1+z)), interval = trunc((x - x0)/dl),
//     offset = x - (x0 + iv dl) - the interpolant is continuous, so a one-ulp difference in x moves the result by
//     one ulp; the two DISCONTINUOUS decisions (window edges, zqso_gp.py:132 and :169-170) are still taken with
//     the reference's own division, by binary search, once per sample;
//   * the Gram operand is m' = m sqrt(1/v): A and B fragments of a diagonal-or-not block are the same three
//     registers (the round-1 kernel carried m and m/v), and the per-pixel record (offset, sqrt weight, weighted
//     residual, interval) goes through a 1 KB shared-memory buffer per warp instead of 7 shuffles per 4 pixels;
//   * (value, slope) pairs are interleaved in one table: 3 LDG.128 per 4 pixels per lane instead of 6 LDG.64, issued
//     one step ahead of their use, and the 3 FMA + 3 MUL that turn them into the next step's operands are threaded
//     between the 6 DMMAs of the current step (measurement 3: a warp that carries its own
//     scalar work between its DMMAs keeps the shared FP64 pipe busy; a warp that alternates phases does not);
//   * the bordered 21 x 21 Cholesky runs in registers, one row per lane, rows broadcast by shuffles
//     (the round-1 version walked shared memory with a square root per index).
// =====================================================================================================
constexpr int ZQ2_ZPB = 32;              // redshift samples per CTA of the median pass
constexpr int ZQ2_REC = 4;               // doubles per pixel record: offset, sqrt weight, weighted residual, interval
constexpr int ZQ2_PER_WARP = ZQ_TDIM * ZQ_TSTRIDE;  // 600 doubles: the T matrix; the two 32 x 4 record buffers overlay it

// (value, slope) interleaved table for the v2 kernel: MS[i][24][2]; columns 0..19 = M, column 20 = -mu (so that the
// lane that owns it forms the weighted residual with one FMA), columns 21..23 zero; the last row's slopes are 0
__global__ void zqso_interleave_kernel(const double* mu, const double* mu_slope, const double* Mp, const double* Ms,
                                       int n_rest, double* MS) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_rest) return;
  const bool last = i + 1 >= n_rest;
  for (int j = 0; j < ZQ_STRIDE; ++j) {
    double v = Mp[(size_t)i * ZQ_STRIDE + j], sl = last ? 0.0 : Ms[(size_t)i * ZQ_STRIDE + j];
    if (j == ZQ_K) { v = -mu[i]; sl = last ? 0.0 : -mu_slope[i]; }
    MS[((size_t)i * ZQ_STRIDE + j) * 2] = v;
    MS[((size_t)i * ZQ_STRIDE + j) * 2 + 1] = sl;
  }
}

struct ZqWindow {
  int lo, hi_end, bw_end, rw_begin;  // observed-frame window [lo, hi_end), bluewards [0, bw_end), redwards [rw_begin, n)
  int wlo, whi;                      // modelled pixels: window and min_lambda <= X/(1+z) <= max_lambda (:169-170)
  int nlo, nhi;                      // normalisation window (:142-148)
};
__device__ __forceinline__ ZqWindow zq_window(const double* X, int n_raw, double opz, const ZqsoParamsDev& prm) {
  ZqWindow w;
  const double max_pos = __dmul_rn(prm.max_lambda, opz), min_pos = __dmul_rn(prm.min_lambda, opz);
  const double max_obs = fmin(max_pos, X[n_raw - 1]), min_obs = fmax(min_pos, X[0]);
  w.lo = zq_bound(X, 0, n_raw, min_obs, true);
  w.hi_end = zq_bound(X, 0, n_raw, max_obs, false);
  w.bw_end = zq_bound(X, 0, n_raw, min_obs, false);
  w.rw_begin = zq_bound(X, 0, n_raw, max_obs, true);
  const int top = max(w.hi_end, w.lo);
  w.wlo = zq_bound_rest(X, opz, w.lo, top, prm.min_lambda, false);  // first x >= min_lambda
  w.whi = zq_bound_rest(X, opz, w.lo, top, prm.max_lambda, true);   // first x >  max_lambda
  w.nlo = zq_bound_rest(X, opz, w.lo, top, prm.norm_min_lambda, false);
  w.nhi = zq_bound_rest(X, opz, w.lo, top, prm.norm_max_lambda, true);
  return w;
}

// CTA-wide bitonic sort of (key, tag) pairs in shared memory, ascending by key; n is a power of two
__device__ __forceinline__ void zq_block_bitonic(double* key, int* tag, int n) {
  for (int k = 2; k <= n; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int l = i ^ j;
        if (l > i) {
          const double a = key[i], b = key[l];
          const bool up = (i & k) == 0;
          if ((a > b) == up) {
            key[i] = b; key[l] = a;
            const int t = tag[i]; tag[i] = tag[l]; tag[l] = t;
          }
        }
      }
      __syncthreads();
    }
}

// the value of rank `k` (0-based) among the sorted entries whose tag lies in [a, b); one warp
__device__ __forceinline__ double zq_select_rank(const double* key, const int* tag, int n, int a, int b, int k, int lane) {
  int seen = 0;
  double val = 0.0;
  for (int base = 0; base < n; base += 32) {
    const int t = tag[base + lane];
    const unsigned in = __ballot_sync(0xffffffffu, t >= a && t < b);
    const int c = __popc(in);
    if (k < seen + c) {
      const int src = __fns(in, 0, k - seen + 1);  // lane of the (k - seen + 1)-th set bit
      val = key[base + src];
      break;
    }
    seen += c;
  }
  return val;
}

// grid = (ceil(S / 32), num_spectra), block = 256, dynamic smem = cap * 12 bytes; med_out [num_spectra][S]
__global__ void __launch_bounds__(256)
zqso_median_batch_kernel(const ZqsoSpectrum* __restrict__ spectra, const double* __restrict__ z_samples, int S,
                         ZqsoParamsDev prm, int cap, double* __restrict__ med_out) {
  extern __shared__ double zq_smem[];
  double* key = zq_smem;
  int* tag = reinterpret_cast<int*>(zq_smem + cap);
  __shared__ int s_nlo[ZQ2_ZPB], s_nhi[ZQ2_ZPB];
  __shared__ int s_union[2];
  const ZqsoSpectrum sp = spectra[blockIdx.y];
  const int s0 = blockIdx.x * ZQ2_ZPB;
  const int nz = min(ZQ2_ZPB, S - s0);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const double inf = __longlong_as_double(0x7ff0000000000000LL);
  const double qnan = __longlong_as_double(0x7ff8000000000000LL);
  if (threadIdx.x < ZQ2_ZPB) {
    int nlo = 0x7fffffff, nhi = 0;
    if (threadIdx.x < nz) {
      const ZqWindow w = zq_window(sp.X, sp.n_raw, __dadd_rn(1.0, z_samples[s0 + threadIdx.x]), prm);
      nlo = w.nlo;
      nhi = max(w.nhi, w.nlo);
      s_nlo[threadIdx.x] = nlo;
      s_nhi[threadIdx.x] = nhi;
    }
    // union of the non-empty windows
    int ulo = nhi > nlo ? nlo : 0x7fffffff, uhi = nhi > nlo ? nhi : 0;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      ulo = min(ulo, __shfl_xor_sync(0xffffffffu, ulo, off));
      uhi = max(uhi, __shfl_xor_sync(0xffffffffu, uhi, off));
    }
    if (threadIdx.x == 0) { s_union[0] = ulo; s_union[1] = uhi; }
  }
  __syncthreads();
  const int ulo = s_union[0], uhi = s_union[1];
  if (uhi <= ulo) {  // every window empty: nanmedian of nothing
    if (threadIdx.x < nz) med_out[(size_t)blockIdx.y * S + s0 + threadIdx.x] = qnan;
    return;
  }
  if (uhi - ulo <= cap) {
    // one sort for the CTA's 32 samples; NaN flux gets tag -1 (np.nanmedian ignores it), pads too
    for (int i = threadIdx.x; i < cap; i += blockDim.x) {
      const int p = ulo + i;
      double v = inf;
      int t = -1;
      if (p < uhi) {
        const double yv = sp.Y[p];
        if (!isnan(yv)) { v = yv; t = p; }
      }
      key[i] = v;
      tag[i] = t;
    }
    __syncthreads();
    zq_block_bitonic(key, tag, cap);
    for (int zi = warp; zi < nz; zi += 8) {
      const int a = s_nlo[zi], b = s_nhi[zi];
      int count = 0;
      for (int base = 0; base < cap; base += 32) {
        const int t = tag[base + lane];
        count += __popc(__ballot_sync(0xffffffffu, t >= a && t < b));
      }
      double med = qnan;
      if (count > 0) {
        const double hi = zq_select_rank(key, tag, cap, a, b, count >> 1, lane);
        med = (count & 1) ? hi : (zq_select_rank(key, tag, cap, a, b, (count >> 1) - 1, lane) + hi) * 0.5;
      }
      if (lane == 0) med_out[(size_t)blockIdx.y * S + s0 + zi] = med;
    }
  } else {
    // samples far apart (not a sorted sweep): one CTA-wide sort per sample
    for (int zi = 0; zi < nz; ++zi) {
      const int a = s_nlo[zi], b = s_nhi[zi];
      __syncthreads();
      for (int i = threadIdx.x; i < cap; i += blockDim.x) {
        const int p = a + i;
        double v = inf;
        int t = -1;
        if (p < b) {
          const double yv = sp.Y[p];
          if (!isnan(yv)) { v = yv; t = p; }
        }
        key[i] = v;
        tag[i] = t;
      }
      __syncthreads();
      zq_block_bitonic(key, tag, cap);
      if (warp == 0) {
        int count = 0;
        for (int base = 0; base < cap; base += 32) count += __popc(__ballot_sync(0xffffffffu, tag[base + lane] >= 0));
        double med = qnan;
        if (count > 0) med = (count & 1) ? key[count >> 1] : (key[(count >> 1) - 1] + key[count >> 1]) * 0.5;
        if (lane == 0) med_out[(size_t)blockIdx.y * S + s0 + zi] = med;
      }
    }
  }
}

__device__ __forceinline__ void zq_dmma_pinned(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ double zq_fma_pinned(double a, double b, double c) {
  double r;
  asm volatile("fma.rn.f64 %0, %1, %2, %3;" : "=d"(r) : "d"(a), "d"(b), "d"(c));
  return r;
}
__device__ __forceinline__ double zq_mul_pinned(double a, double b) {
  double r;
  asm volatile("mul.rn.f64 %0, %1, %2;" : "=d"(r) : "d"(a), "d"(b));
  return r;
}
// 1 / sqrt(x): MUFU.RSQ64H seed and one cubic correction (see likelihood_kernel.cuh : fast_rsqrt)
__device__ __forceinline__ double zq_rsqrt(double x) {
  double y0;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(x));
  const double e = fma(-(x * y0), y0, 1.0);
  const double p = fma(e, 0.375, 0.5) * e;
  const double y = fma(y0, p, y0);
  return (x > 1e-290 && x < 1e290) ? y : 1.0 / sqrt(x);
}

// grid = (ceil(S / ZQ2_WARPS), num_spectra), block = 32 ZQ2_WARPS, dynamic smem = ZQ2_WARPS * ZQ2_PER_WARP doubles.  Uniform model grid only.
#ifndef ZQ2_MIN_CTAS
#define ZQ2_MIN_CTAS 2
#endif
#ifndef ZQ2_WARPS
#define ZQ2_WARPS 8
#endif
#ifndef ZQ2_PREFETCH_PIX
#define ZQ2_PREFETCH_PIX 1
#endif
#ifndef ZQ2_SCALAR_FIRST
#define ZQ2_SCALAR_FIRST 0
#endif
#ifndef ZQ2_LOCKSTEP
#define ZQ2_LOCKSTEP 0   // experiment: CTA barrier every ZQ2_LOCKSTEP chunks so that the warps (adjacent redshifts) hit the same table rows
#endif
__global__ void __launch_bounds__(ZQ2_WARPS * 32, ZQ2_MIN_CTAS)
zqso_likelihood_kernel_v2(const ZqsoSpectrum* __restrict__ spectra, const double* __restrict__ z_samples, int S,
                          const double* __restrict__ med_all, ZqsoModelDev model, ZqsoParamsDev prm,
                          double* __restrict__ out /* [num_spectra][S] */) {
  extern __shared__ double zq_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int s = blockIdx.x * ZQ2_WARPS + warp;
  if (s >= S) return;
  const ZqsoSpectrum sp = spectra[blockIdx.y];
  double* buf = zq_smem + (size_t)warp * ZQ2_PER_WARP;
  const double* __restrict__ X = sp.X;
  const int n_raw = sp.n_raw;
  const double z = z_samples[s];
  const double opz = __dadd_rn(1.0, z);
  const ZqWindow w = zq_window(X, n_raw, opz, prm);
  const int wlo = max(w.lo, w.wlo), whi = min(w.hi_end, w.whi);
  const double med = med_all[(size_t)blockIdx.y * S + s];
  const double invmed = 1.0 / med;
  const double invmed2 = 1.0 / (med * med);
  const double ropz = 1.0 / opz;
  const double dl = 1.0 / model.inv_dl;
  const int iv_max = model.n_rest - 2;

  const int grp = lane >> 2, tig = lane & 3;
  double acc[6][2];  // blocks (0,0) (1,0) (1,1) (2,0) (2,1) (2,2)
#pragma unroll
  for (int b = 0; b < 6; ++b) acc[b][0] = acc[b][1] = 0.0;
  LogProd vprod;
  int n_sel = 0;

  // per-pixel record of chunk starting at c0 -> rec[lane] = {offset, sqrt weight, weighted flux, table row}.
  // The pixel's inputs are fetched one chunk ahead (fetch), so that the arithmetic never waits for a global load
  // (third version: the loads sat in front of the first DFMA of produce, 12 % of all stall samples).
  double pX = 0.0, pV = 1.0, pY = 0.0;
  bool pM = true;
  auto fetch = [&](int c0) {
    const int pc = min(c0 + lane, n_raw - 1);
    pX = X[pc];
    pV = sp.V[pc];
    pY = sp.Y[pc];
    pM = sp.mask[pc] != 0;
  };
  auto produce = [&](int c0, double* rec) {
#if !ZQ2_PREFETCH_PIX
    fetch(c0);
#endif
    const int p = c0 + lane;
    const bool sel = p < whi && !pM;
    const double x = pX * ropz;
    int iv = (int)((x - model.rest0) * model.inv_dl);
    iv = max(0, min(iv, iv_max));
    const double xo = fma(-(double)iv, dl, x - model.rest0);
    const double vn = pV * invmed2;
    double sw = zq_rsqrt(vn);
    double ys = (pY * invmed) * sw;
    if (!sel) { sw = 0.0; ys = 0.0; }
    vprod.mul(sel ? vn : 1.0);
    n_sel += sel ? 1 : 0;
    double4 v;
    v.x = xo; v.y = sw; v.z = ys; v.w = __longlong_as_double((long long)iv);
    reinterpret_cast<double4*>(rec)[lane] = v;
#if ZQ2_PREFETCH_PIX
    fetch(c0 + 32);
#endif
  };
  // LOAD LAYOUT: lane L = 8 pg + j reads, for pixel pg of the step, the (value, slope) pairs of columns j, j + 8, j + 16 -
  // the 8 lanes of a pixel cover 128 contiguous bytes per instruction, 4 L1 wavefronts per LDG.128 (round 2, first
  // version: the DMMA fragment layout itself, lane = 4 grp + tig, put 4 different table rows into every quarter-warp:
  // 16 wavefronts per LDG.128, l1tex__data_pipe_lsu_wavefronts at 95 % of peak and the tensor pipe at 38 %).
  // The operands are formed in this layout and moved to the fragment layout (pixel = tig, columns grp + 8 c) by three
  // 64-bit shuffles with the fixed lane transpose src = 8 tig + grp.
  const int pg = lane >> 3, lj = lane & 7;
  const int frag_src = tig * 8 + grp;
  const bool owns_residual = lj == 4;  // column 16 + 4 = 20 carries -mu: m' = (y - mu) sqrt(1/v)
  const double2* MS2 = reinterpret_cast<const double2*>(model.MS) + lj;
  auto load_pairs = [&](const double4& rc, double2& q0, double2& q1, double2& q2) {
    const double2* row = MS2 + (size_t)__double_as_longlong(rc.w) * ZQ_STRIDE;
    q0 = __ldg(row);
    q1 = __ldg(row + 8);
    q2 = __ldg(row + 16);
  };
  auto to_fragment = [&](double& v0, double& v1, double& v2) {
    v0 = __shfl_sync(0xffffffffu, v0, frag_src);
    v1 = __shfl_sync(0xffffffffu, v1, frag_src);
    v2 = __shfl_sync(0xffffffffu, v2, frag_src);
  };

  // Four-stage software pipeline over the 4-pixel steps g = 8 c + kb:
  //   step g :  SHFL  operands of step g + 1 (formed during step g - 1) -> fragment layout: six DMMAs of cover
  //             LDG   (value, slope) pairs of step g + 3 (its record was read from shared memory during step g - 1)
  //             LDS   record of step g + 4
  //             DMMA  x 6 of step g, with the 3 FMA + 3 MUL/FMA that form the operands of step g + 2 between them
  // so no consumer issues in the step that issued its producer (second version: LDS -> address -> LDG -> FMA inside one
  // step: long_scoreboard 33 %; third: operands shuffled at the end of the step that uses them next: short_scoreboard
  // on the first DMMAs).
  const int nchunks = whi > wlo ? (whi - wlo + 31) >> 5 : 0;
#if ZQ2_LOCKSTEP
  __shared__ int s_min_chunks;
  if (threadIdx.x == 0) s_min_chunks = 0x7fffffff;
  __syncthreads();
  if (lane == 0) atomicMin(&s_min_chunks, nchunks);
  __syncthreads();
  const int lock_chunks = s_min_chunks;  // every warp of the CTA runs at least this many chunks
#endif
  if (nchunks > 0) {
    double* rec0 = buf;
    double* rec1 = buf + 32 * ZQ2_REC;
#if ZQ2_PREFETCH_PIX
    fetch(wlo);
#endif
    produce(wlo, rec0);
    __syncwarp();
    const double4* r0 = reinterpret_cast<const double4*>(rec0);
    double2 qa0, qa1, qa2, qb0, qb1, qb2;
    auto form = [&](const double4& rc, const double2& q0, const double2& q1, const double2& q2, double& v0, double& v1,
                    double& v2) {
      v0 = fma(q0.y, rc.x, q0.x) * rc.y;
      v1 = fma(q1.y, rc.x, q1.x) * rc.y;
      v2 = fma(fma(q2.y, rc.x, q2.x), rc.y, owns_residual ? rc.z : 0.0);
    };
    // prologue: fragment operands of step 0, load-layout operands of step 1, pairs of step 2, records of steps 2 and 3
    double4 ra = r0[pg];
    load_pairs(ra, qa0, qa1, qa2);
    double m0, m1, m2, n0, n1, n2;
    form(ra, qa0, qa1, qa2, m0, m1, m2);
    to_fragment(m0, m1, m2);
    ra = r0[4 + pg];
    load_pairs(ra, qa0, qa1, qa2);
    form(ra, qa0, qa1, qa2, n0, n1, n2);
    ra = r0[8 + pg];                       // record of step 2 and its pairs
    load_pairs(ra, qa0, qa1, qa2);
    double4 rb = r0[12 + pg];              // record of step 3 (pairs loaded in step 0)
    for (int c = 0; c < nchunks; ++c) {
      double* cur = (c & 1) ? rec1 : rec0;
      double* nxt = (c & 1) ? rec0 : rec1;
      if ((c & 7) == 7) vprod.renorm();
#if ZQ2_LOCKSTEP
      if (c < lock_chunks && (c % ZQ2_LOCKSTEP) == 0) __syncthreads();
#endif
#pragma unroll
      for (int kb = 0; kb < 8; ++kb) {
        if (kb == 0) __syncwarp();                      // every lane has read its last record of chunk c - 1 from `nxt`
        if (kb == 1) produce(wlo + (c + 1) * 32, nxt);  // past the end: zero-weight records, never multiplied in
        if (kb == 3) __syncwarp();                      // records of chunk c + 1 visible before step kb = 4 reads them
        double f0 = n0, f1 = n1, f2 = n2;
        to_fragment(f0, f1, f2);                        // operands of step g + 1
        load_pairs(rb, qb0, qb1, qb2);                  // pairs of step g + 3
        const double4 rc = kb + 4 < 8 ? reinterpret_cast<const double4*>(cur)[(kb + 4) * 4 + pg]
                                      : reinterpret_cast<const double4*>(nxt)[(kb + 4 - 8) * 4 + pg];
        const double zadd = owns_residual ? ra.z : 0.0;
#if ZQ2_SCALAR_FIRST
        const double t0 = zq_fma_pinned(qa0.y, ra.x, qa0.x);
        const double t1 = zq_fma_pinned(qa1.y, ra.x, qa1.x);
        const double t2 = zq_fma_pinned(qa2.y, ra.x, qa2.x);
        n0 = zq_mul_pinned(t0, ra.y);
        n1 = zq_mul_pinned(t1, ra.y);
        n2 = zq_fma_pinned(t2, ra.y, zadd);
        zq_dmma_pinned(acc[0][0], acc[0][1], m0, m0);
        zq_dmma_pinned(acc[1][0], acc[1][1], m1, m0);
        zq_dmma_pinned(acc[2][0], acc[2][1], m1, m1);
        zq_dmma_pinned(acc[3][0], acc[3][1], m2, m0);
        zq_dmma_pinned(acc[4][0], acc[4][1], m2, m1);
        zq_dmma_pinned(acc[5][0], acc[5][1], m2, m2);
#else
        zq_dmma_pinned(acc[0][0], acc[0][1], m0, m0);
        const double t0 = zq_fma_pinned(qa0.y, ra.x, qa0.x);
        zq_dmma_pinned(acc[1][0], acc[1][1], m1, m0);
        const double t1 = zq_fma_pinned(qa1.y, ra.x, qa1.x);
        zq_dmma_pinned(acc[2][0], acc[2][1], m1, m1);
        const double t2 = zq_fma_pinned(qa2.y, ra.x, qa2.x);
        zq_dmma_pinned(acc[3][0], acc[3][1], m2, m0);
        n0 = zq_mul_pinned(t0, ra.y);
        zq_dmma_pinned(acc[4][0], acc[4][1], m2, m1);
        n1 = zq_mul_pinned(t1, ra.y);
        zq_dmma_pinned(acc[5][0], acc[5][1], m2, m2);
        n2 = zq_fma_pinned(t2, ra.y, zadd);
#endif
        m0 = f0; m1 = f1; m2 = f2;
        qa0 = qb0; qa1 = qb1; qa2 = qb2;
        ra = rb;
        rb = rc;
      }
    }
  }
  double lv = vprod.value();
  if (!(vprod.prod > 0.0)) lv = __longlong_as_double(0x7ff8000000000000LL);  // zero / negative / NaN variance: poison, never +inf
  const double sum_log_v = warp_sum(lv);
  const int n_in = (int)(warp_sum((double)n_sel) + 0.5);

  // ---- fragments -> T (24 x 24, lower blocks) -> one row per lane ------------------------------------------------
  __syncwarp();
  double* T = buf;
  {
    const int bi_of[6] = {0, 1, 1, 2, 2, 2}, bj_of[6] = {0, 0, 1, 0, 1, 2};
#pragma unroll
    for (int b = 0; b < 6; ++b) {
      const int row = bi_of[b] * 8 + grp, col = bj_of[b] * 8 + tig * 2;
      T[row * ZQ_TSTRIDE + col] = acc[b][0];
      T[row * ZQ_TSTRIDE + col + 1] = acc[b][1];
    }
  }
  __syncwarp();
  // Cholesky of the bordered matrix [[B, c], [c', q]] in registers: lane i owns row i (0..20; row 20 = [c', q]).
  // Right-looking: column j is scaled by 1/sqrt(pivot), then every row subtracts l_ij l_kj from its entries k > j;
  // after the 20 columns entry (20, 20) is q - z'z (null_gp.py:345-358).  Lanes > 20 compute on zeros.
  double row[ZQ_K + 1];
#pragma unroll
  for (int k = 0; k <= ZQ_K; ++k) {
    double v = (lane <= ZQ_K && k <= lane) ? T[lane * ZQ_TSTRIDE + k] : 0.0;
    if (k == lane && k < ZQ_K) v += 1.0;  // + I (null_gp.py:341); row 20 is the projection row
    row[k] = v;
  }
  LogProd pivots;
#pragma unroll
  for (int j = 0; j < ZQ_K; ++j) {
    const double piv = __shfl_sync(0xffffffffu, row[j], j);
    pivots.mul(piv);
    if ((j & 3) == 3) pivots.renorm();
    const double inv = zq_rsqrt(piv);
    const double l = row[j] * inv;
    row[j] = l;
#pragma unroll
    for (int k = j + 1; k <= ZQ_K; ++k) {
      const double lk = __shfl_sync(0xffffffffu, l, k);
      row[k] = fma(-l, lk, row[k]);
    }
  }
  const double quad = __shfl_sync(0xffffffffu, row[ZQ_K], ZQ_K);  // q - z'z
  double lp = pivots.value();
  if (!(pivots.prod > 0.0)) lp = __longlong_as_double(0x7ff8000000000000LL);
  const double ll_window = -0.5 * (quad + (sum_log_v + lp) + (double)n_in * ZQ_LOG_2PI);

  // ---- i.i.d. Gaussians bluewards and redwards of the window (:160-166, :198-210, :252-278) -----------------
  double side_ll[2];
#pragma unroll
  for (int side = 0; side < 2; ++side) {
    const int begin = side == 0 ? 0 : w.rw_begin, end = side == 0 ? w.bw_end : n_raw;
    const double m_side = side == 0 ? model.bluewards_mu : model.redwards_mu;
    const double var_side = side == 0 ? model.bluewards_var : model.redwards_var;
    double qs = 0.0;
    LogProd dprod;
    int cnt = 0, it = 0;
    for (int p = begin + lane; p < end; p += 32, ++it) {
      if (!sp.mask[p]) {
        const double t = fma(sp.Y[p], invmed, -m_side);
        const double dd = fma(sp.V[p], invmed2, var_side);
        qs = fma(t * t, zq_rcp(dd), qs);
        dprod.mul(dd);
        ++cnt;
      }
      if ((it & 7) == 7) dprod.renorm();
    }
    double ld = dprod.value();
    if (!(dprod.prod > 0.0)) ld = __longlong_as_double(0x7ff8000000000000LL);
    const double q_tot = warp_sum(qs), ld_tot = warp_sum(ld);
    const int c_tot = (int)(warp_sum((double)cnt) + 0.5);
    side_ll[side] = -0.5 * (q_tot + ld_tot + (double)c_tot * ZQ_LOG_2PI);
  }
  if (lane == 0) out[(size_t)blockIdx.y * S + s] = ll_window + side_ll[0] + side_ll[1];
}

// np.nanargmax over the samples of each spectrum (first maximum wins); -1 when every sample is NaN
__global__ void __launch_bounds__(256) zqso_argmax_kernel(const double* __restrict__ ll, int S, const double* z_samples,
                                                          double* z_map, int32_t* map_index) {
  __shared__ double s_val[256];
  __shared__ int s_idx[256];
  const double* row = ll + (size_t)blockIdx.x * S;
  double best = 0.0;
  int bi = -1;
  for (int i = threadIdx.x; i < S; i += blockDim.x) {
    const double v = row[i];
    if (!isnan(v) && (bi < 0 || v > best)) { best = v; bi = i; }
  }
  s_val[threadIdx.x] = best;
  s_idx[threadIdx.x] = bi;
  __syncthreads();
  for (int off = 128; off > 0; off >>= 1) {
    if (threadIdx.x < off) {
      const int oi = s_idx[threadIdx.x + off];
      const double ov = s_val[threadIdx.x + off];
      const int ci = s_idx[threadIdx.x];
      const double cv = s_val[threadIdx.x];
      if (oi >= 0 && (ci < 0 || ov > cv || (ov == cv && oi < ci))) { s_val[threadIdx.x] = ov; s_idx[threadIdx.x] = oi; }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const int i = s_idx[0];
    map_index[blockIdx.x] = i;
    z_map[blockIdx.x] = i >= 0 ? z_samples[i] : __longlong_as_double(0x7ff8000000000000LL);
  }
}

// ZGP.set_data + get_interp at ONE redshift, element-wise over the raw pixels (attribute path; exact
// two-rounding interpolation arithmetic of scipy interp1d).  cls: 0 none, 1 modelled, 2 bluewards, 3 redwards.
// The median is computed by the likelihood kernel's sorter (launched with one sample) and passed in.
__global__ void zqso_set_data_kernel(ZqsoSpectrum sp, double z, ZqsoModelDev model, ZqsoParamsDev prm, double med,
                                     double* x, double* yn, double* vn, double* this_mu, double* this_M /* n_raw x k */,
                                     uint8_t* cls, uint8_t* in_window) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= sp.n_raw) return;
  const double opz = __dadd_rn(1.0, z);
  const double max_pos = __dmul_rn(prm.max_lambda, opz), min_pos = __dmul_rn(prm.min_lambda, opz);
  const double max_obs = fmin(max_pos, sp.X[sp.n_raw - 1]), min_obs = fmax(min_pos, sp.X[0]);
  const double X = sp.X[p];
  const double xp = __ddiv_rn(X, opz);
  const bool inw = X > min_obs && X < max_obs;
  const bool masked = sp.mask[p] != 0;
  x[p] = xp;
  yn[p] = __ddiv_rn(sp.Y[p], med);
  vn[p] = __ddiv_rn(sp.V[p], __dmul_rn(med, med));
  uint8_t c = 0;
  if (inw && !masked && xp >= prm.min_lambda && xp <= prm.max_lambda) c = 1;
  else if (X < min_obs && !masked) c = 2;
  else if (X > max_obs && !masked) c = 3;
  cls[p] = c;
  in_window[p] = inw ? 1 : 0;
  if (c == 1) {
    const int iv = zqso_interval(model, xp);
    const double xoff = __dsub_rn(xp, model.rest[iv]);
    this_mu[p] = __dadd_rn(__dmul_rn(model.mu_slope[iv], xoff), model.mu[iv]);
    for (int j = 0; j < ZQ_K; ++j)
      this_M[(size_t)p * ZQ_K + j] =
          __dadd_rn(__dmul_rn(model.M_slope[(size_t)iv * ZQ_STRIDE + j], xoff), model.M[(size_t)iv * ZQ_STRIDE + j]);
  }
}

// nanmedian of the normalisation window at one redshift (same sorter as the likelihood kernel), one warp
__global__ void zqso_median_kernel(ZqsoSpectrum sp, double z, ZqsoParamsDev prm, int norm_cap, double* med_out) {
  extern __shared__ double zq_smem[];
  const int lane = threadIdx.x & 31;
  double* buf = zq_smem;
  const double opz = __dadd_rn(1.0, z);
  const double max_pos = __dmul_rn(prm.max_lambda, opz), min_pos = __dmul_rn(prm.min_lambda, opz);
  const double max_obs = fmin(max_pos, sp.X[sp.n_raw - 1]), min_obs = fmax(min_pos, sp.X[0]);
  const int lo = zq_bound(sp.X, 0, sp.n_raw, min_obs, true);
  const int hi_end = zq_bound(sp.X, 0, sp.n_raw, max_obs, false);
  const int nlo = zq_bound_rest(sp.X, opz, lo, max(hi_end, lo), prm.norm_min_lambda, false);
  const int nhi = zq_bound_rest(sp.X, opz, lo, max(hi_end, lo), prm.norm_max_lambda, true);
  const double inf = __longlong_as_double(0x7ff0000000000000LL);
  int count = 0;
  for (int i = lane; i < norm_cap; i += 32) {
    const int p = nlo + i;
    double val = inf;
    if (p < nhi) {
      const double yv = sp.Y[p];
      if (!isnan(yv)) { val = yv; ++count; }
    }
    buf[i] = val;
  }
  count = (int)(warp_sum((double)count) + 0.5);
  __syncwarp();
  for (int k = 2; k <= norm_cap; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = lane; i < norm_cap; i += 32) {
        const int l = i ^ j;
        if (l > i) {
          const double a = buf[i], b = buf[l];
          const bool up = (i & k) == 0;
          if ((a > b) == up) { buf[i] = b; buf[l] = a; }
        }
      }
      __syncwarp();
    }
  if (lane == 0) {
    double med;
    if (count == 0) med = __longlong_as_double(0x7ff8000000000000LL);
    else if (count & 1) med = buf[count >> 1];
    else med = (buf[(count >> 1) - 1] + buf[count >> 1]) * 0.5;
    *med_out = med;
  }
}

// log N(y; mu, diag(d)) (zqso_gp.py:252-278), one CTA
__global__ void __launch_bounds__(256) log_mvnpdf_iid_kernel(const double* y, const double* mu, const double* d, int n,
                                                             double* out) {
  __shared__ double red[2][8];
  double q = 0.0, ld = 0.0;
  for (int p = threadIdx.x; p < n; p += blockDim.x) {
    const double r = y[p] - mu[p];
    q = fma(r / d[p], r, q);
    ld += log(d[p]);
  }
  q = warp_sum(q);
  ld = warp_sum(ld);
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = q; red[1][threadIdx.x >> 5] = ld; }
  __syncthreads();
  if (threadIdx.x == 0) {
    q = 0.0; ld = 0.0;
    for (int w = 0; w < 8; ++w) { q += red[0][w]; ld += red[1][w]; }
    out[0] = -0.5 * (q + ld + (double)n * ZQ_LOG_2PI);
  }
}

}  // namespace dla
