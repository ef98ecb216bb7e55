// prep_kernel.cuh : per-spectrum preparation on the device.
//
// Reference: NullGP.set_data (null_gp.py:95-177), NullGP.get_interp (:179-242),
// effective_optical_depth (effective_optical_depth.py:10-80), Parameters.min_z_dla /
// max_z_dla (set_parameters.py:125-159).  One CTA per spectrum:
//   1. flux normalisation by the nanmedian over the rest-frame window, ignoring masked pixels
//   2. in-range mask `ind_unmasked`, modelled-pixel mask `ind` (bit-exact requirement) and
//      the compaction maps
//   3. linear interpolation of mu, the k columns of M and log omega (np.interp arithmetic:
//      slope * (x - x_lo) + y_lo, no FMA contraction)
//   4. mean-flux suppression with the Kim et al. parameters and the learned (tau_0, beta)
//      rescaling of omega^2 over num_forest_lines Lyman members
//   5. the padded wavelength grid for the instrument convolution and the z_DLA search range
#pragma once
#include <stdint.h>
#include <math.h>
#include "lyman_tables.h"

namespace dla {

__device__ __constant__ double c_tw_A[LYMAN_NUM_LINES] = LYMAN_TRANSITION_WAVELENGTHS_A;
__device__ __constant__ double c_osc[LYMAN_NUM_LINES] = LYMAN_OSCILLATOR_STRENGTHS;

struct ModelDev {
  const double* rest_wavelengths;  // n_rest
  const double* mu;                // n_rest
  const double* M;                 // n_rest x k row-major
  const double* log_omega;         // n_rest
  int n_rest, k;
  double log_c_0, log_tau_0, log_beta, prev_tau_0, prev_beta;
};

struct PrepParams {
  double min_lambda, max_lambda, norm_min_lambda, norm_max_lambda, pixel_spacing;
  double lya_wavelength, lyman_limit, max_z_cut, min_z_cut;
  int width, num_forest_lines, broadening, normalize;
};

// One spectrum's raw inputs and prepared outputs (all device pointers).
struct PrepTask {
  // inputs
  const double* X;        // n_raw rest wavelengths  (observed / (1 + z_qso)), or nullptr when Wobs is given
  const double* Wobs;     // n_raw observed wavelengths (catalogue path: X = Wobs / (1 + z_qso) formed here)
  const double* Y;        // n_raw flux
  const double* V;        // n_raw noise variance
  const uint8_t* mask;    // n_raw pixel mask (1 = bad)
  int n_raw;
  double z_qso;
  // outputs, capacities n_raw (+ 2*width for padded)
  uint8_t* ind_unmasked;  // n_raw
  uint8_t* ind;           // n_raw
  double* x;              // n   rest wavelengths of modelled pixels
  double* y;              // n   normalised flux
  double* v;              // n   normalised noise variance
  double* this_wl;        // n   observed wavelengths of modelled pixels
  double* mu;             // n   this_mu
  double* omega2;         // n   this_omega2
  double* M;              // n x k this_M
  int32_t* uidx;          // n   index of each modelled pixel within the in-range grid
  double* unmasked_wl;    // n_u
  double* wl_abs;         // n_u + 2*width (broadening) else n_u : grid the absorption is evaluated on
  double* padded_wl;      // n_u + 2*width (always, attribute parity with the reference)
  double* scratch;        // n_raw doubles (median selection)
  // scalars out: [0]=n_u [1]=n (as doubles), [2]=median, [3]=min_z_dla, [4]=max_z_dla (from this_wl),
  //              [5]=min_z_dla, [6]=max_z_dla from the full raw grid (run_bayes_select.py:194-195)
  double* scalars;
};

// np.interp for one point: binary search + slope formula (numpy/core/src/multiarray/compiled_base.c)
__device__ __forceinline__ int interp_locate(const double* xp, int n, double x) {
  // largest j with xp[j] <= x, clamped to [0, n-2]
  int lo = 0, hi = n - 1;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (xp[mid] <= x) lo = mid; else hi = mid;
  }
  return lo;
}
__device__ __forceinline__ double interp_eval(const double* xp, const double* fp, int stride, int n, int j, double x) {
  if (x == xp[n - 1]) return fp[(size_t)(n - 1) * stride];
  if (xp[j] == x) return fp[(size_t)j * stride];
  const double f0 = fp[(size_t)j * stride], f1 = fp[(size_t)(j + 1) * stride];
  const double slope = __ddiv_rn(__dsub_rn(f1, f0), __dsub_rn(xp[j + 1], xp[j]));
  double r = __dadd_rn(__dmul_rn(slope, __dsub_rn(x, xp[j])), f0);
  if (isnan(r)) {
    r = __dadd_rn(__dmul_rn(slope, __dsub_rn(x, xp[j + 1])), f1);
    if (isnan(r) && f0 == f1) r = f0;
  }
  return r;
}

// sum over the Lyman members of tau_0 (f_i lambda_i)/(f_1 lambda_1) (lambda/lambda_i)^beta [z_i <= z_qso]
// (effective_optical_depth.py:51-78)
__device__ __forceinline__ double total_optical_depth(double wl, double beta, double tau_0, double z_qso, int nlines) {
  double total = 0.0;
  for (int i = 0; i < nlines; ++i) {
    const double z_i = __ddiv_rn(__dsub_rn(wl, c_tw_A[i]), c_tw_A[i]);
    const double this_tau_0 =
        __ddiv_rn(__dmul_rn(__ddiv_rn(__dmul_rn(tau_0, c_osc[i]), c_osc[0]), c_tw_A[i]), c_tw_A[0]);
    double t = __dmul_rn(this_tau_0, pow(__dadd_rn(1.0, z_i), beta));
    t = (z_i <= z_qso) ? t : t * 0.0;
    total += t;
  }
  return total;
}

// NullGP.get_interp for one pixel (null_gp.py:179-242): interpolate mu, log omega and the k columns of M at rest
// wavelength x, apply the Kim et al. mean-flux suppression at observed wavelength wl to mu and M, and rescale
// omega^2 by the learned (tau_0, beta) Lyman-series factor.  Explicitly rounded operations: no FMA contraction.
__device__ __forceinline__ void interp_model_pixel(const ModelDev& model, int num_forest_lines, double x, double wl,
                                                   double z_qso, double beta2, double tau2, double c0, double* mu_out,
                                                   double* omega2_out, double* M_out) {
  const int j = interp_locate(model.rest_wavelengths, model.n_rest, x);
  double mu = interp_eval(model.rest_wavelengths, model.mu, 1, model.n_rest, j, x);
  const double log_omega = interp_eval(model.rest_wavelengths, model.log_omega, 1, model.n_rest, j, x);
  double omega2 = exp(__dmul_rn(2.0, log_omega));
  const double lya_abs = exp(-total_optical_depth(wl, model.prev_beta, model.prev_tau_0, z_qso, num_forest_lines));
  mu = __dmul_rn(mu, lya_abs);
  const double od2 = total_optical_depth(wl, beta2, tau2, z_qso, num_forest_lines);
  const double scaling = __dadd_rn(__dsub_rn(1.0, exp(-od2)), c0);
  omega2 = __dmul_rn(omega2, __dmul_rn(scaling, scaling));
  omega2 = __dmul_rn(omega2, __dmul_rn(lya_abs, lya_abs));
  *mu_out = mu;
  *omega2_out = omega2;
  for (int c = 0; c < model.k; ++c) {
    const double mv = interp_eval(model.rest_wavelengths, model.M + c, model.k, model.n_rest, j, x);
    M_out[c] = __dmul_rn(mv, lya_abs);
  }
}

// NullGP.get_interp on an arbitrary set of pixels (x rest, wl observed), one thread per pixel
__global__ void interp_model_kernel(ModelDev model, int num_forest_lines, const double* __restrict__ x,
                                    const double* __restrict__ wl, int n, double z_qso, double* __restrict__ mu,
                                    double* __restrict__ M, double* __restrict__ omega2) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n) return;
  const double beta2 = exp(model.log_beta), tau2 = exp(model.log_tau_0), c0 = exp(model.log_c_0);
  interp_model_pixel(model, num_forest_lines, x[q], wl[q], z_qso, beta2, tau2, c0, mu + q, omega2 + q, M + (size_t)q * model.k);
}

// block-wide exclusive scan of one int per thread (blockDim.x <= 1024)
__device__ __forceinline__ int block_exclusive_scan(int val, int* s_warp, int& block_total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int inc = val;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, inc, off);
    if (lane >= off) inc += t;
  }
  if (lane == 31) s_warp[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    int w = lane < (int)(blockDim.x >> 5) ? s_warp[lane] : 0;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, w, off);
      if (lane >= off) w += t;
    }
    s_warp[lane] = w;  // inclusive over warps
  }
  __syncthreads();
  const int warp_off = warp > 0 ? s_warp[warp - 1] : 0;
  block_total = s_warp[(blockDim.x >> 5) - 1];
  __syncthreads();
  return warp_off + inc - val;
}

// grid = num_spectra, block = 256
__global__ void __launch_bounds__(256)
prepare_spectrum_kernel(const PrepTask* __restrict__ tasks, ModelDev model, PrepParams P) {
  const PrepTask t = tasks[blockIdx.x];
  __shared__ int s_warp[32];
  __shared__ int s_count;
  __shared__ double s_median;
  __shared__ double s_mm[6];
  const int tid = threadIdx.x;
  const int n_raw = t.n_raw;
  const double zp1 = __dadd_rn(1.0, t.z_qso);
  // rest wavelength of raw pixel i: given, or Parameters.emitted_wavelengths(observed, z_qso)
  auto rest_at = [&](int i) -> double { return t.X ? t.X[i] : __ddiv_rn(t.Wobs[i], zp1); };

  // ---- 1. normalisation median (null_gp.py:125-136) -----------------------------------------
  double median = 1.0;
  if (P.normalize) {
    if (tid == 0) s_count = 0;
    __syncthreads();
    // gather window values in index order (order is irrelevant for a median; NaNs dropped = nanmedian)
    for (int base = 0; base < n_raw; base += blockDim.x) {
      const int i = base + tid;
      int flag = 0;
      double yv = 0.0;
      if (i < n_raw) {
        const double x = rest_at(i);
        yv = t.Y[i];
        flag = (x >= P.norm_min_lambda) && (x <= P.norm_max_lambda) && !t.mask[i] && !isnan(yv);
      }
      int tot;
      const int pos = block_exclusive_scan(flag, s_warp, tot);
      if (flag) t.scratch[s_count + pos] = yv;
      __syncthreads();
      if (tid == 0) s_count += tot;
      __syncthreads();
    }
    const int m = s_count;
    if (tid == 0) s_median = NAN;  // empty window: nanmedian -> NaN
    __syncthreads();
    // rank selection: element with rank m/2 (and m/2-1 when m is even)
    if (m > 0) {
      __shared__ double s_sel[2];
      const int r_hi = m / 2, r_lo = (m % 2) ? m / 2 : m / 2 - 1;
      for (int i = tid; i < m; i += blockDim.x) {
        const double vi = t.scratch[i];
        int rank = 0;
        for (int j = 0; j < m; ++j) {
          const double vj = t.scratch[j];
          rank += (vj < vi) || (vj == vi && j < i);
        }
        if (rank == r_hi) s_sel[1] = vi;
        if (rank == r_lo) s_sel[0] = vi;
      }
      __syncthreads();
      if (tid == 0) s_median = (m % 2) ? s_sel[1] : __ddiv_rn(__dadd_rn(s_sel[0], s_sel[1]), 2.0);
    }
    __syncthreads();
    median = s_median;
  }
  const double median2 = __dmul_rn(median, median);

  // ---- 2. masks + compaction (null_gp.py:139-152) -------------------------------------------
  int n_u = 0, n = 0;
  double mn_r = INFINITY, mx_r = -INFINITY;  // raw observed wavelengths of the in-range pixels
  __shared__ int s_nu, s_n;
  if (tid == 0) { s_nu = 0; s_n = 0; s_mm[0] = INFINITY; s_mm[1] = -INFINITY; s_mm[2] = INFINITY; s_mm[3] = -INFINITY; s_mm[4] = INFINITY; s_mm[5] = -INFINITY; }
  __syncthreads();
  for (int base = 0; base < n_raw; base += blockDim.x) {
    const int i = base + tid;
    int in_range = 0, keep = 0;
    double x = 0.0;
    if (i < n_raw) {
      x = rest_at(i);
      in_range = (x >= P.min_lambda) && (x <= P.max_lambda);
      keep = in_range && !t.mask[i];
      t.ind_unmasked[i] = (uint8_t)in_range;
      t.ind[i] = (uint8_t)keep;
    }
    int tot_u, tot_n;
    const int pos_u = block_exclusive_scan(in_range, s_warp, tot_u);
    const int pos_n = block_exclusive_scan(keep, s_warp, tot_n);
    const int off_u = s_nu, off_n = s_n;
    if (in_range) {
      const double obs = __dmul_rn(x, zp1);  // Parameters.observed_wavelengths
      t.unmasked_wl[off_u + pos_u] = obs;
      t.padded_wl[off_u + pos_u + P.width] = obs;
      const double raw = t.Wobs ? t.Wobs[i] : obs;
      mn_r = fmin(mn_r, raw);
      mx_r = fmax(mx_r, raw);
    }
    if (keep) {
      const int q = off_n + pos_n;
      const double obs = __dmul_rn(x, zp1);
      t.x[q] = x;
      t.this_wl[q] = obs;
      t.uidx[q] = off_u + pos_u;
      if (P.normalize) {
        t.y[q] = __ddiv_rn(t.Y[i], median);
        t.v[q] = __ddiv_rn(t.V[i], median2);
      } else {
        t.y[q] = t.Y[i];
        t.v[q] = t.V[i];
      }
    }
    __syncthreads();
    if (tid == 0) { s_nu += tot_u; s_n += tot_n; }
    __syncthreads();
  }
  n_u = s_nu;
  n = s_n;

  // ---- 3./4. interpolation + mean-flux suppression (null_gp.py:179-242) ---------------------
  const double beta2 = exp(model.log_beta), tau2 = exp(model.log_tau_0), c0 = exp(model.log_c_0);
  for (int q = tid; q < n; q += blockDim.x)
    interp_model_pixel(model, P.num_forest_lines, t.x[q], t.this_wl[q], t.z_qso, beta2, tau2, c0, t.mu + q, t.omega2 + q,
                       t.M + (size_t)q * model.k);

  // ---- 5. padded grid (null_gp.py:159-177) and z_DLA range (set_parameters.py:125-159) -------
  // min / max of the in-range observed wavelengths and of the modelled ones
  {
    double mn_u = INFINITY, mx_u = -INFINITY, mn_n = INFINITY, mx_n = -INFINITY;
    for (int i = tid; i < n_u; i += blockDim.x) { const double w = t.unmasked_wl[i]; mn_u = fmin(mn_u, w); mx_u = fmax(mx_u, w); }
    // Parameters.min_z_dla / max_z_dla re-filter the wavelengths they are given by emitted(w, z_qso) in
    // [min_lambda, max_lambda] (set_parameters.py:125-159); the samplers hand them this_wavelengths = x (1 + z_qso)
    // (dla_gp.py:122-124), and x (1 + z) / (1 + z) can fall one ulp outside the range at either end
    for (int i = tid; i < n; i += blockDim.x) {
      const double w = t.this_wl[i];
      const double rest = __ddiv_rn(w, zp1);
      if (rest >= P.min_lambda && rest <= P.max_lambda) { mn_n = fmin(mn_n, w); mx_n = fmax(mx_n, w); }
    }
    for (int off = 16; off > 0; off >>= 1) {
      mn_u = fmin(mn_u, __shfl_xor_sync(0xffffffffu, mn_u, off));
      mx_u = fmax(mx_u, __shfl_xor_sync(0xffffffffu, mx_u, off));
      mn_n = fmin(mn_n, __shfl_xor_sync(0xffffffffu, mn_n, off));
      mx_n = fmax(mx_n, __shfl_xor_sync(0xffffffffu, mx_n, off));
      mn_r = fmin(mn_r, __shfl_xor_sync(0xffffffffu, mn_r, off));
      mx_r = fmax(mx_r, __shfl_xor_sync(0xffffffffu, mx_r, off));
    }
    __shared__ double s_w[6][8];
    if ((tid & 31) == 0) {
      s_w[0][tid >> 5] = mn_u; s_w[1][tid >> 5] = mx_u; s_w[2][tid >> 5] = mn_n; s_w[3][tid >> 5] = mx_n;
      s_w[4][tid >> 5] = mn_r; s_w[5][tid >> 5] = mx_r;
    }
    __syncthreads();
    if (tid == 0) {
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w) {
        s_mm[0] = fmin(s_mm[0], s_w[0][w]); s_mm[1] = fmax(s_mm[1], s_w[1][w]);
        s_mm[2] = fmin(s_mm[2], s_w[2][w]); s_mm[3] = fmax(s_mm[3], s_w[3][w]);
        s_mm[4] = fmin(s_mm[4], s_w[4][w]); s_mm[5] = fmax(s_mm[5], s_w[5][w]);
      }
    }
    __syncthreads();
  }
  if (tid == 0) {
    t.scalars[0] = (double)n_u;
    t.scalars[1] = (double)n;
    t.scalars[2] = median;
    if (n_u > 0) {
      // np.logspace(a, b, 3) = 10 ** [a, a + (b-a)/2, b]
      const double lo = log10(s_mm[0]), hi = log10(s_mm[1]);
      const int w = P.width;
      for (int side = 0; side < 2; ++side) {
        const double a = side == 0 ? __dsub_rn(lo, __dmul_rn((double)w, P.pixel_spacing)) : __dadd_rn(hi, P.pixel_spacing);
        const double b = side == 0 ? __dsub_rn(lo, P.pixel_spacing) : __dadd_rn(hi, __dmul_rn((double)w, P.pixel_spacing));
        const double step = w > 1 ? __ddiv_rn(__dsub_rn(b, a), (double)(w - 1)) : 0.0;
        for (int i = 0; i < w; ++i) {
          double e = __dadd_rn(__dmul_rn((double)i, step), a);
          if (i == w - 1 && w > 1) e = b;
          t.padded_wl[(side == 0 ? 0 : n_u + w) + i] = pow(10.0, e);
        }
      }
    }
    // z range from the modelled pixels (what the samplers use, dla_gp.py:122-124)
    const double zc_lim = __dadd_rn(__dsub_rn(__ddiv_rn(__dmul_rn(P.lyman_limit, zp1), P.lya_wavelength), 1.0), P.min_z_cut);
    if (n > 0) {
      const double zmax_a = __dsub_rn(__dsub_rn(__ddiv_rn(s_mm[3], P.lya_wavelength), 1.0), P.max_z_cut);
      const double zmax_b = __dsub_rn(t.z_qso, P.max_z_cut);
      t.scalars[4] = fmin(zmax_a, zmax_b);
      const double zmin_a = __dsub_rn(__ddiv_rn(s_mm[2], P.lya_wavelength), 1.0);
      t.scalars[3] = fmax(zmin_a, zc_lim);
    } else {
      t.scalars[3] = NAN; t.scalars[4] = NAN;
    }
    // and from every in-range pixel of the raw grid (run_bayes_select.py:194-195)
    if (n_u > 0) {
      const double zmax_a = __dsub_rn(__dsub_rn(__ddiv_rn(s_mm[5], P.lya_wavelength), 1.0), P.max_z_cut);
      t.scalars[6] = fmin(zmax_a, __dsub_rn(t.z_qso, P.max_z_cut));
      t.scalars[5] = fmax(__dsub_rn(__ddiv_rn(s_mm[4], P.lya_wavelength), 1.0), zc_lim);
    } else {
      t.scalars[5] = NAN; t.scalars[6] = NAN;
    }
  }
  __syncthreads();
  // absorption grid: padded (broadening) or unmasked wavelengths (dla_gp.py:364-367)
  if (P.broadening) {
    for (int i = tid; i < n_u + 2 * P.width; i += blockDim.x) t.wl_abs[i] = t.padded_wl[i];
  } else {
    for (int i = tid; i < n_u; i += blockDim.x) t.wl_abs[i] = t.unmasked_wl[i];
  }
}

// a2 stand-alone: effective_optical_depth -> (n, num_forest_lines)
__global__ void effective_optical_depth_kernel(const double* wavelengths, int n, double beta, double tau_0, double z_qso,
                                               int num_forest_lines, double* out) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n * num_forest_lines) return;
  const int p = e / num_forest_lines, i = e % num_forest_lines;
  const double wl = wavelengths[p];
  const double z_i = __ddiv_rn(__dsub_rn(wl, c_tw_A[i]), c_tw_A[i]);
  const double this_tau_0 =
      __ddiv_rn(__dmul_rn(__ddiv_rn(__dmul_rn(tau_0, c_osc[i]), c_osc[0]), c_tw_A[i]), c_tw_A[0]);
  double tt = __dmul_rn(this_tau_0, pow(__dadd_rn(1.0, z_i), beta));
  out[e] = (z_i <= z_qso) ? tt : tt * 0.0;
}

}  // namespace dla
