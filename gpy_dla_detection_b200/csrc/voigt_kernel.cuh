// voigt_kernel.cuh : Lyman-series absorption profiles on the device.
//
// Reference: voigt.voigt_absorption (voigt.py:251-322) and its native twin voigt.c:253-304.
// One warp owns one (z_DLA, N_HI) sample of one spectrum: it evaluates the raw profile
// exp(N_HI * sum_l -lc_l V_l(lambda)) on the padded wavelength grid into shared memory, then
// applies the 7-tap instrument convolution and the pixel-mask compaction
// (dla_gp.py:360,388) while writing the row of the profile cache, so the unbroadened
// profile never touches HBM.
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
  double* out;         // profile rows, row stride ld
  int n_in;            // n_u + 2*width (broadening) or n_u
  int n_out;           // modelled pixels n
  int ld;              // row stride (doubles)
  int num_samples;     // rows to produce for this spectrum
  const double* z;     // num_samples absorber redshifts
  const double* nhi;   // num_samples column densities
};

// raw profile value at one wavelength (voigt.py:296-307)
__device__ __forceinline__ double raw_profile_at(double lam, const double* mult, double nhi, int num_lines) {
  double total = 0.0;
  for (int l = 0; l < num_lines; ++l) {
    // velocity = wavelengths * multipliers[l] - c   : two roundings, no FMA contraction
    const double vel = __dsub_rn(__dmul_rn(lam, mult[l]), LYMAN_C_CGS);
    // z = (v + i gamma) / (sqrt(2) sigma): numpy multiplies by the reciprocal of the real divisor
    const double x = __dmul_rn(vel, LYMAN_INV_SQRT2_SIGMA);
    const double h = dla_faddeeva_re(x, c_damping_y[l]);
    // -leading_constants[l] * (Re w / (sqrt(2 pi) sigma)), folded into one constant (<= 1 ulp apart)
    const double term = c_coef[l] * h;
    if (!isnan(term)) total += term;  // np.nansum
  }
  return exp(nhi * total);
}

// grid = (ceil(max_samples / warps_per_cta), num_spectra), block = 32 * warps_per_cta,
// dynamic smem = warps_per_cta * smem_row doubles, smem_row >= n_in + 32
__global__ void __launch_bounds__(256)
voigt_profile_kernel(const AbsorptionGrid* __restrict__ grids, int num_lines, int broadening, int smem_row) {
  extern __shared__ double s_raw[];
  const AbsorptionGrid g = grids[blockIdx.y];
  const int warps_per_cta = blockDim.x >> 5;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int sample = blockIdx.x * warps_per_cta + warp;
  if (sample >= g.num_samples) return;
  double* raw = s_raw + (size_t)warp * smem_row;
  double* mult = raw + (smem_row - 32);  // last 32 doubles of the warp's row hold the multipliers

  const double zd = g.z[sample];
  const double nhi = g.nhi[sample];
  // multipliers = c / (transition_wavelengths * (1 + z_dla)) / 1e8   (voigt.py:296)
  if (lane < num_lines)
    mult[lane] = __ddiv_rn(__ddiv_rn(LYMAN_C_CGS, __dmul_rn(c_tw_cm[lane], __dadd_rn(1.0, zd))), 1e8);
  __syncwarp();

  for (int p = lane; p < g.n_in; p += 32) raw[p] = raw_profile_at(g.wl[p], mult, nhi, num_lines);
  __syncwarp();

  double* out = g.out + (size_t)sample * g.ld;
  if (broadening) {
    for (int q = lane; q < g.n_out; q += 32) {
      const int u = g.uidx[q];
      // np.convolve(raw, profile, 'valid')[u] = sum_k raw[u+k] * profile[6-k]
      double acc = 0.0;
#pragma unroll
      for (int k = 0; k <= 2 * INSTRUMENT_WIDTH; ++k) acc = fma(raw[u + k], c_instrument[2 * INSTRUMENT_WIDTH - k], acc);
      out[q] = acc;
    }
  } else {
    for (int q = lane; q < g.n_out; q += 32) out[q] = raw[g.uidx[q]];
  }
  // the pad columns [n_out, ld) are never read as data (the likelihood kernel masks p >= n)
}

// plain element-wise Faddeeva evaluation for the accuracy tests
__global__ void faddeeva_kernel(const double* x, const double* y, double* out, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = dla_faddeeva_re(x[i], y[i]);
}

}  // namespace dla
