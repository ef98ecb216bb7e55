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

// exp(x) for the profile, x = N_HI * sum of line terms (<= 0).  Same structure as any exp - argument
// reduction x = i ln2 + r by the magic-number rint, degree-13 Taylor polynomial on |r| <= ln2 / 2
// (truncation 4e-18 relative), scaling by 2^i in the exponent field - but with the coefficients in the
// constant bank, where a DFMA reads them as operands: the library routine materialises every
// coefficient with two uniform-register moves, a quarter of the instructions of the profile's inner
// loop.  <= 2 ulp from the correctly rounded value.  Outside [-708, 700] (deeply saturated cores,
// non-finite input) the library routine handles underflow, overflow and NaN.
__device__ __constant__ double c_exp_taylor[14] = {
    1.0 / 6227020800.0, 1.0 / 479001600.0, 1.0 / 39916800.0, 1.0 / 3628800.0, 1.0 / 362880.0, 1.0 / 40320.0,
    1.0 / 5040.0,       1.0 / 720.0,       1.0 / 120.0,      1.0 / 24.0,      1.0 / 6.0,      0.5,
    1.0,                1.0};
__device__ __constant__ double c_exp_consts[4] = {1.4426950408889634074, 6.93147180559945286227e-01,
                                                  2.31904681384629955842e-17, 6755399441055744.0};
__device__ __forceinline__ double profile_exp(double x) {
  if (!(x >= -708.0 && x <= 700.0)) return exp(x);
  double t = fma(x, c_exp_consts[0], c_exp_consts[3]);
  const int i = __double2loint(t);
  t -= c_exp_consts[3];
  double r = fma(-t, c_exp_consts[1], x);
  r = fma(-t, c_exp_consts[2], r);
  double p = c_exp_taylor[0];
#pragma unroll
  for (int k = 1; k < 14; ++k) p = fma(p, r, c_exp_taylor[k]);
  return __hiloint2double(__double2hiint(p) + (i << 20), __double2loint(p));
}

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
  int lls_break;       // add the Lyman-limit break optical depth to the exponent (voigt_lls.py:254-284)
  int pair_offset;     // 0, or S when samples i and i + S share their redshift (the reference's DLA and subDLA
                       // samples use the same offset_samples, set_lls_parameters.m:22 / subdla_samples.py:87):
                       // only the first S samples are launched, each warp evaluates the line sums once
                       // and writes rows i and i + S with the two column densities
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

// sum over the lines of -lc_l V_l(lambda) at one wavelength (voigt.py:296-307); the raw profile is
// exp(N_HI * sum).  NL > 0: compile-time number of lines.
// Must be called by all 32 lanes (warp vote): when every lane is in the far wing (|x| >= 64) of
// every line - more than 9 chunks in 10 - the lines are evaluated in straight-line code; the
// arithmetic is the same sequence of operations as the general path, so a value does not depend on
// which path produced it.
template <int NL>
__device__ __forceinline__ double line_sum_at(double lam, const double* mult, const float* multf, int num_lines) {
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
      return total;
    }
#pragma unroll
    for (int l = 0; l < NL; ++l) {
      const double term = c_coef[l] * dla_faddeeva_re(x[l], c_damping_y[l]);
      if (!isnan(term)) total += term;  // np.nansum
    }
    return total;
  }
  // Any number of lines (31 in BASELINE configs[3]).  Round 2: the same warp-uniform far-wing shortcut as the 3-line
  // instantiation.  The test runs in FP32 on the FP32 pipe with a margin (|x| >= 65 there implies |x| >= 64 in
  // FP64: the rounding error of x in FP32 is below 2e-5 of 65), so it costs the FP64 pipe nothing; when every lane is
  // beyond every line - all but the one or two chunks that hold a line core, the high-order lines lie bluewards of
  // the spectrum for most absorbers - the lines are summed in straight-line code, four at a time, by the very
  // function the general path calls for |x| >= 64, so a value does not depend on the path that produced it.
  bool far = true;
  {
    const float lamf = (float)lam;
    for (int l = 0; l < nl; ++l) {
      const float xf = (lamf * multf[l] - (float)LYMAN_C_CGS) * (float)LYMAN_INV_SQRT2_SIGMA;
      far = far && (fabsf(xf) >= 65.0f);
    }
  }
  if (__all_sync(0xffffffffu, far)) {
#pragma unroll 4
    for (int l = 0; l < nl; ++l) {
      const double vel = __dsub_rn(__dmul_rn(lam, mult[l]), LYMAN_C_CGS);
      const double x = __dmul_rn(vel, LYMAN_INV_SQRT2_SIGMA);
      const double y = c_damping_y[l];
      const double ax = fabs(x);
      total += c_coef[l] * dla_faddeeva_far(dla_wing_rcp(ax * ax), y, y * y);  // finite by construction
    }
    return total;
  }
  for (int l = 0; l < nl; ++l) {
    const double vel = __dsub_rn(__dmul_rn(lam, mult[l]), LYMAN_C_CGS);
    const double x = __dmul_rn(vel, LYMAN_INV_SQRT2_SIGMA);
    const double h = dla_faddeeva_re(x, c_damping_y[l]);
    // -leading_constants[l] * (Re w / (sqrt(2 pi) sigma)), folded into one constant (<= 1 ulp apart)
    const double term = c_coef[l] * h;
    if (!isnan(term)) total += term;  // np.nansum
  }
  return total;
}

// Lyman-limit break of an absorber at redshift z (voigt_lls.py:254-284):
//   tau = nhi / 10^17.2 * (lambda_rest / 911.7641)^3  for lambda_rest <= 911.7641 A, else 0;  returned per unit
//   (nhi / 10^17.2), i.e. the cube alone, so that paired samples share it
constexpr double LLS_LIMIT_A = 911.7641;
__device__ __forceinline__ double lls_break_cube(double lam, double one_plus_z) {
  const double rest = __ddiv_rn(lam, one_plus_z);
  return rest > LLS_LIMIT_A ? 0.0 : pow(__ddiv_rn(rest, LLS_LIMIT_A), 3.0);
}

constexpr int VG_WARPS = 8;  // samples per CTA

// grid = (ceil(max_samples / 8), num_spectra), block = 256, static smem only
template <int NL>
__global__ void __launch_bounds__(VG_WARPS * 32)
voigt_profile_kernel(const AbsorptionGrid* __restrict__ grids, int num_lines, int broadening) {
  // ring of raw chunks: even chunks at [0,32) and again at [64,96), odd chunks at [32,64): the 7-tap
  // window of an output pixel is contiguous whatever the parity, so the taps are immediate offsets
  __shared__ double s_ring[VG_WARPS][2][96];
  __shared__ double s_mult[VG_WARPS][32];
  __shared__ float s_multf[VG_WARPS][32];  // FP32 copies for the far-wing test of the any-line-count instantiation
  const AbsorptionGrid g = grids[blockIdx.y];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int sample = blockIdx.x * VG_WARPS + warp;
  const bool paired = g.pair_offset > 0;
  if (sample >= (paired ? g.pair_offset : g.num_samples) || g.num_samples == 0) return;
  double* mult = s_mult[warp];
  float* multf = s_multf[warp];

  const double zd = g.z[sample];
  const double nhi = g.nhi[sample];
  const double nhi2 = paired ? g.nhi[sample + g.pair_offset] : 0.0;
  const bool lls = g.lls_break != 0;
  const double opz = __dadd_rn(1.0, zd);
  const double lls_scale = lls ? __ddiv_rn(nhi, pow(10.0, 17.2)) : 0.0;    // np.float64(nhi) / 10**17.2
  const double lls_scale2 = lls ? __ddiv_rn(nhi2, pow(10.0, 17.2)) : 0.0;
  // multipliers = c / (transition_wavelengths * (1 + z_dla)) / 1e8   (voigt.py:296)
  if (lane < (NL > 0 ? NL : num_lines)) {
    mult[lane] = __ddiv_rn(__ddiv_rn(LYMAN_C_CGS, __dmul_rn(c_tw_cm[lane], __dadd_rn(1.0, zd))), 1e8);
    multf[lane] = (float)mult[lane];
  }
  __syncwarp();

  double* out = g.out + (size_t)sample * g.ld;
  double* out2 = g.out + (size_t)(sample + g.pair_offset) * g.ld;
  if (!broadening) {
    for (int p0 = 0; p0 < g.n_in; p0 += 32) {
      const int p = p0 + lane;
      const double lam = g.wl[min(p, g.n_in - 1)];
      const double total = line_sum_at<NL>(lam, mult, multf, num_lines);
      const double cube = lls ? lls_break_cube(lam, opz) : 0.0;
      const int q = p < g.n_in ? g.qmap[p] : -1;
      if (q >= 0) {
        out[q] = profile_exp(__dsub_rn(__dmul_rn(nhi, total), __dmul_rn(lls_scale, cube)));
        if (paired) out2[q] = profile_exp(__dsub_rn(__dmul_rn(nhi2, total), __dmul_rn(lls_scale2, cube)));
      }
    }
    return;
  }
  // np.convolve(raw, profile, 'valid')[u] = sum_k raw[u+k] * profile[6-k], u < n_u = n_in - 6
  const int n_u = g.n_in - 2 * INSTRUMENT_WIDTH;
  const int nchunks = (g.n_in + 31) >> 5;
  double* ring = s_ring[warp][0];
  double* ring2 = s_ring[warp][1];
  // the two global loads of an iteration are issued an iteration (wavelength) or a chunk of arithmetic (output
  // index) ahead of their use (they were the kernel's only long-scoreboard stalls; with 8 resident warps per
  // sub-partition the other warps covered most of them already: + 0.1 % of the step)
  double lam_next = g.wl[min(lane, g.n_in - 1)];
  for (int j = 0; j <= nchunks; ++j) {
    const int u = ((j - 1) << 5) + lane;     // output pixel of the convolution step of this iteration
    const int q = (j >= 1 && u < n_u) ? g.qmap[u] : -1;
    if (j < nchunks) {
      const double lam = lam_next;                                      // tail lanes: unused copies
      lam_next = g.wl[min(((j + 1) << 5) + lane, g.n_in - 1)];
      const double total = line_sum_at<NL>(lam, mult, multf, num_lines);
      const double cube = lls ? lls_break_cube(lam, opz) : 0.0;        // warp-uniform branch
      const double rawv = profile_exp(__dsub_rn(__dmul_rn(nhi, total), __dmul_rn(lls_scale, cube)));
      ring[((j & 1) << 5) + lane] = rawv;
      if ((j & 1) == 0) ring[64 + lane] = rawv;
      if (paired) {
        const double rawv2 = profile_exp(__dsub_rn(__dmul_rn(nhi2, total), __dmul_rn(lls_scale2, cube)));
        ring2[((j & 1) << 5) + lane] = rawv2;
        if ((j & 1) == 0) ring2[64 + lane] = rawv2;
      }
    }
    __syncwarp();
    if (q >= 0) {
      const int w0 = (((j - 1) & 1) << 5) + lane;  // elements u .. u + 6
      double acc = 0.0;
#pragma unroll
      for (int k = 0; k <= 2 * INSTRUMENT_WIDTH; ++k) acc = fma(ring[w0 + k], c_instrument[2 * INSTRUMENT_WIDTH - k], acc);
      out[q] = acc;
      if (paired) {
        double acc2 = 0.0;
#pragma unroll
        for (int k = 0; k <= 2 * INSTRUMENT_WIDTH; ++k)
          acc2 = fma(ring2[w0 + k], c_instrument[2 * INSTRUMENT_WIDTH - k], acc2);
        out2[q] = acc2;
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
