// faddeeva.cuh : Re w(x + i y) for the Lyman-series Voigt profile, FP64, sm_100a.
//
// Replaces the reference's call to scipy.special.wofz inside voigt.Voigt (voigt.py:241-248;
// libcerf voigt() in voigt.c:288).  The damping parameter y = gamma_l / (sqrt(2) sigma) is a
// per-line constant between 7.2e-8 (Ly-31) and 4.72e-4 (Ly-alpha); |x| reaches ~2e4.
// For y this small the Taylor expansion about the real axis converges in three terms:
//
//   Re w = e^{-x^2} [1 - y^2 (2x^2-1) + y^4 (16x^4 - 48x^2 + 12)/24] + y K1(x) + y^3 K3(x)
//
// (truncation < 1e-15 relative for y <= 5e-4, 2.3e-14 at y = 1e-3; tools/gen_faddeeva_tables.py).
// K1, K3 are piecewise polynomials: 32 core intervals on |x| < 8, two wing intervals in
// t = 1/x^2 on [8, 64) and one short polynomial for |x| >= 64, which is where > 97 % of the
// pixels of a spectrum fall.  Max relative error against 60-digit mpmath over the domain is
// checked in tests/test_faddeeva_host.py (host build of this header) and on the GPU.
//
// The header compiles for host and device: the host build is used by the CPU test-suite only.
#pragma once
#include <math.h>
#include "faddeeva_tables.h"

#if defined(__CUDACC__)
#define DLA_HD __host__ __device__ __forceinline__
#define DLA_TABLE_QUAL __device__ __constant__
#else
#define DLA_HD inline
#define DLA_TABLE_QUAL static const
#endif

// Tables live in global memory on the device (divergent indexing near line cores would
// serialise in the constant cache); a second constant copy holds the tiny far-wing data.
#if defined(__CUDACC__)
__device__ const double g_fadd_k1[FADD_CORE_N * (FADD_K1_DEG + 1)] = FADD_K1_TABLE;
__device__ const double g_fadd_k3[FADD_CORE_N * (FADD_K3_DEG + 1)] = FADD_K3_TABLE;
__device__ const double g_fadd_w1[2 * (FADD_W1_DEG + 1)] = FADD_W1_TABLE;
__device__ const double g_fadd_w3[2 * (FADD_W3_DEG + 1)] = FADD_W3_TABLE;
__device__ const double g_fadd_wc[2] = FADD_W_CENTER;
__device__ const double g_fadd_ws[2] = FADD_W_SCALE;
#endif
static const double h_fadd_k1[FADD_CORE_N * (FADD_K1_DEG + 1)] = FADD_K1_TABLE;
static const double h_fadd_k3[FADD_CORE_N * (FADD_K3_DEG + 1)] = FADD_K3_TABLE;
static const double h_fadd_w1[2 * (FADD_W1_DEG + 1)] = FADD_W1_TABLE;
static const double h_fadd_w3[2 * (FADD_W3_DEG + 1)] = FADD_W3_TABLE;
static const double h_fadd_wc[2] = FADD_W_CENTER;
static const double h_fadd_ws[2] = FADD_W_SCALE;

#if defined(__CUDA_ARCH__)
#define FADD_TAB(name) g_fadd_##name
#else
#define FADD_TAB(name) h_fadd_##name
#endif

// 1/x^2 for the wing expansions.  Device: MUFU.RCP64H seed (SFU) + two Newton steps (4 DFMA), <= 1 ulp
// from the IEEE quotient, against ~9 FP64-pipe slots for the IEEE division; x2 is in [64, 1e9] here.
DLA_HD double dla_wing_rcp(double x2) {
#if defined(__CUDA_ARCH__)
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x2));
  double e = fma(-x2, r, 1.0);
  r = fma(r, e, r);
  e = fma(-x2, r, 1.0);
  return fma(r, e, r);
#else
  return 1.0 / x2;
#endif
}

// Far-wing coefficients: on the device they sit in the constant bank, where a DFMA reads them as a
// direct operand (as literals each one costs two uniform-register moves per use).
#if defined(__CUDACC__)
__device__ __constant__ double c_fadd_f1[FADD_F1_DEG + 1] = FADD_F1_COEF;
__device__ __constant__ double c_fadd_f3[FADD_F3_DEG + 1] = FADD_F3_COEF;
#endif

// Far wing, |x| >= 64, given t = 1/x^2 : returns Re w.
DLA_HD double dla_faddeeva_far(double t, double y, double y2) {
#if defined(__CUDA_ARCH__)
  const double* f1 = c_fadd_f1;
  const double* f3 = c_fadd_f3;
#else
  const double f1[FADD_F1_DEG + 1] = FADD_F1_COEF;
  const double f3[FADD_F3_DEG + 1] = FADD_F3_COEF;
#endif
  double a1 = f1[0];
#pragma unroll
  for (int i = 1; i <= FADD_F1_DEG; ++i) a1 = fma(a1, t, f1[i]);
  double a3 = fma(f3[0], t, f3[1]);
  return (y * FADD_INV_SQRT_PI) * t * fma(y2 * t, a3, a1);
}

// General entry.  x any sign; 0 <= y <= 1e-3.
DLA_HD double dla_faddeeva_re(double x, double y) {
  const double ax = fabs(x);
  const double y2 = y * y;
  const double x2 = ax * ax;
  if (ax >= 64.0) {
    return dla_faddeeva_far(dla_wing_rcp(x2), y, y2);
  }
  if (ax >= 8.0) {
    const int j = ax >= 16.0 ? 1 : 0;
    const double t = dla_wing_rcp(x2);
    const double u = (t - FADD_TAB(wc)[j]) * FADD_TAB(ws)[j];
    const double* c1 = FADD_TAB(w1) + j * (FADD_W1_DEG + 1);
    const double* c3 = FADD_TAB(w3) + j * (FADD_W3_DEG + 1);
    double a1 = c1[0];
#pragma unroll
    for (int i = 1; i <= FADD_W1_DEG; ++i) a1 = fma(a1, u, c1[i]);
    double a3 = c3[0];
#pragma unroll
    for (int i = 1; i <= FADD_W3_DEG; ++i) a3 = fma(a3, u, c3[i]);
    return (y * FADD_INV_SQRT_PI) * t * fma(y2 * t, a3, a1);
  }
  // core
  int i = (int)(ax * 4.0);
  i = i > FADD_CORE_N - 1 ? FADD_CORE_N - 1 : i;
  const double u = fma(ax, 8.0, -(double)(2 * i + 1));
  const double* c1 = FADD_TAB(k1) + i * (FADD_K1_DEG + 1);
  const double* c3 = FADD_TAB(k3) + i * (FADD_K3_DEG + 1);
  double k1 = c1[0];
#pragma unroll
  for (int q = 1; q <= FADD_K1_DEG; ++q) k1 = fma(k1, u, c1[q]);
  double k3 = c3[0];
#pragma unroll
  for (int q = 1; q <= FADD_K3_DEG; ++q) k3 = fma(k3, u, c3[q]);
  // e^{-x^2} with the rounding error of x^2 folded back in
  const double lo = fma(ax, ax, -x2);
  double g = exp(-x2);
  g = fma(-g, lo, g);
  const double even = fma(y2 * y2, fma(x2, fma(x2, 16.0, -48.0), 12.0) * (1.0 / 24.0), fma(-y2, fma(2.0, x2, -1.0), 1.0));
  return fma(g, even, y * fma(y2, k3, k1));
}
