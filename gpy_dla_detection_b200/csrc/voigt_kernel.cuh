// voigt_kernel.cuh : Lyman-series absorption profiles on the device.
//
// Reference: voigt.voigt_absorption (voigt.py:251-322) and its native twin voigt.c:253-304.
// One warp owns one (z_DLA, N_HI) sample of one spectrum and streams over the wavelength grid in
// 32-pixel chunks: raw profile exp(N_HI * sum_l -lc_l V_l(lambda)) per lane, a two-chunk ring in
// shared memory (512 B per warp - the full row is never staged, so occupancy is register-limited),
// then the 7-tap instrument convolution and the pixel-mask compaction (dla_gp.py:360,388) while
// writing the row of the profile cache.  The unbroadened profile never touches HBM.
#pragma once
#include <stdint.h>
#include "faddeeva.cuh"
#include "lyman_tables.h"

namespace dla {

__device__ __constant__ double c_tw_cm[LYMAN_NUM_LINES] = LYMAN_TRANSITION_WAVELENGTHS_CM;
__device__ __constant__ double c_damping_y[LYMAN_NUM_LINES] = LYMAN_DAMPING_Y;
__device__ __constant__ double c_coef[LYMAN_NUM_LINES] = LYMAN_NEG_LEADING_OVER_NORM;
__device__ __constant__ double c_instrument[2 * INSTRUMENT_WIDTH + 1] = INSTRUMENT_PROFILE;

// One absorption grid (one spectrum): what the profile kernel needs to know.
struct AbsorptionGrid {
  const double* wl;    // n_in wavelengths the raw profile is evaluated on (padded or unmasked grid)
  const int32_t* uidx; // n_out: for each output pixel, its index in the in-range grid
  const int32_t* qmap; // n_u: for each in-range pixel, its output index or -1 when masked (inverse of uidx)
  double* out;         // profile rows, row stride ld
  int n_in;            // n_u + 2*width (broadening) or n_u
  int n_out;           // modelled pixels n
  int ld;              // row stride (doubles)
  int num_samples;     // rows to produce for this spectrum
  const double* z;     // num_samples absorber redshifts
  const double* nhi;   // num_samples column densities
};

// qmap = inverse of uidx; one CTA per spectrum
__global__ void __launch_bounds__(256) build_qmap_kernel(const AbsorptionGrid* __restrict__ grids, int broadening) {
  const AbsorptionGrid g = grids[blockIdx.x];
  const int n_u = broadening ? g.n_in - 2 * INSTRUMENT_WIDTH : g.n_in;
  int32_t* qmap = const_cast<int32_t*>(g.qmap);
  for (int u = threadIdx.x; u < n_u; u += blockDim.x) qmap[u] = -1;
  __syncthreads();
  for (int q = threadIdx.x; q < g.n_out; q += blockDim.x) qmap[g.uidx[q]] = q;
}

// raw profile value at one wavelength (voigt.py:296-307); NL > 0: compile-time number of lines.
// Must be called by all 32 lanes (warp vote): when every lane is in the far wing (|x| >= 64) of
// every line - more than 9 chunks in 10 - the lines are evaluated in straight-line code; the
// arithmetic is the same sequence of operations as the general path, so a value does not depend on
// which path produced it.
template <int NL>
__device__ __forceinline__ double raw_profile_at(double lam, const double* mult, double nhi, int num_lines) {
  double total = 0.0;
  const int nl = NL > 0 ? NL : num_lines;
  if (NL > 0) {
    double x[NL > 0 ? NL : 1];
    bool far = true;
#pragma unroll
    for (int l = 0; l < NL; ++l) {
      // velocity = wavelengths * multipliers[l] - c   : two roundings, no FMA contraction
      const double vel = __dsub_rn(__dmul_rn(lam, mult[l]), LYMAN_C_CGS);
      // z = (v + i gamma) / (sqrt(2) sigma): numpy multiplies by the reciprocal of the real divisor
      x[l] = __dmul_rn(vel, LYMAN_INV_SQRT2_SIGMA);
      far = far && (fabs(x[l]) >= 64.0);
    }
    if (__all_sync(0xffffffffu, far)) {
#pragma unroll
      for (int l = 0; l < NL; ++l) {
        const double y = c_damping_y[l];
        const double ax = fabs(x[l]);
        const double h = dla_faddeeva_far(dla_wing_rcp(ax * ax), y, y * y);
        total += c_coef[l] * h;  // finite by construction: nansum has nothing to skip
      }
      return exp(nhi * total);
    }
#pragma unroll
    for (int l = 0; l < NL; ++l) {
      const double term = c_coef[l] * dla_faddeeva_re(x[l], c_damping_y[l]);
      if (!isnan(term)) total += term;  // np.nansum
    }
    return exp(nhi * total);
  }
  for (int l = 0; l < nl; ++l) {
    const double vel = __dsub_rn(__dmul_rn(lam, mult[l]), LYMAN_C_CGS);
    const double x = __dmul_rn(vel, LYMAN_INV_SQRT2_SIGMA);
    const double h = dla_faddeeva_re(x, c_damping_y[l]);
    // -leading_constants[l] * (Re w / (sqrt(2 pi) sigma)), folded into one constant (<= 1 ulp apart)
    const double term = c_coef[l] * h;
    if (!isnan(term)) total += term;  // np.nansum
  }
  return exp(nhi * total);
}

constexpr int VG_WARPS = 8;  // samples per CTA

// grid = (ceil(max_samples / 8), num_spectra), block = 256, static smem only
template <int NL>
__global__ void __launch_bounds__(VG_WARPS * 32)
voigt_profile_kernel(const AbsorptionGrid* __restrict__ grids, int num_lines, int broadening) {
  __shared__ double s_ring[VG_WARPS][2][32];
  __shared__ double s_mult[VG_WARPS][32];
  const AbsorptionGrid g = grids[blockIdx.y];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int sample = blockIdx.x * VG_WARPS + warp;
  if (sample >= g.num_samples) return;
  double* mult = s_mult[warp];

  const double zd = g.z[sample];
  const double nhi = g.nhi[sample];
  // multipliers = c / (transition_wavelengths * (1 + z_dla)) / 1e8   (voigt.py:296)
  if (lane < (NL > 0 ? NL : num_lines))
    mult[lane] = __ddiv_rn(__ddiv_rn(LYMAN_C_CGS, __dmul_rn(c_tw_cm[lane], __dadd_rn(1.0, zd))), 1e8);
  __syncwarp();

  double* out = g.out + (size_t)sample * g.ld;
  if (!broadening) {
    for (int p0 = 0; p0 < g.n_in; p0 += 32) {
      const int p = p0 + lane;
      const double a = raw_profile_at<NL>(g.wl[min(p, g.n_in - 1)], mult, nhi, num_lines);
      const int q = p < g.n_in ? g.qmap[p] : -1;
      if (q >= 0) out[q] = a;
    }
    return;
  }
  // np.convolve(raw, profile, 'valid')[u] = sum_k raw[u+k] * profile[6-k], u < n_u = n_in - 6
  const int n_u = g.n_in - 2 * INSTRUMENT_WIDTH;
  const int nchunks = (g.n_in + 31) >> 5;
  double (*ring)[32] = s_ring[warp];
  for (int j = 0; j <= nchunks; ++j) {
    if (j < nchunks) {
      const int p = (j << 5) + lane;
      ring[j & 1][lane] = raw_profile_at<NL>(g.wl[min(p, g.n_in - 1)], mult, nhi, num_lines);  // tail lanes: unused copies
    }
    __syncwarp();
    if (j >= 1) {
      const int u = ((j - 1) << 5) + lane;
      if (u < n_u) {
        const int q = g.qmap[u];
        if (q >= 0) {
          double acc = 0.0;
#pragma unroll
          for (int k = 0; k <= 2 * INSTRUMENT_WIDTH; ++k) {
            const int t = lane + k;  // element u + k lives in chunk j-1 (t < 32) or chunk j
            acc = fma(ring[(j - 1 + (t >> 5)) & 1][t & 31], c_instrument[2 * INSTRUMENT_WIDTH - k], acc);
          }
          out[q] = acc;
        }
      }
    }
    __syncwarp();  // chunk j-1 is overwritten by chunk j+1 in the next iteration
  }
  // the pad columns [n_out, ld) are never read as data (the likelihood kernel masks p >= n)
}

// plain element-wise Faddeeva evaluation for the accuracy tests
__global__ void faddeeva_kernel(const double* x, const double* y, double* out, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = dla_faddeeva_re(x[i], y[i]);
}

}  // namespace dla
