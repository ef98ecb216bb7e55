// Host build of csrc/faddeeva.cuh for the CPU test-suite (tests/test_faddeeva_host.py).
// TEST INFRASTRUCTURE: validates the polynomial tables and the region logic without a GPU.
#include "faddeeva.cuh"
extern "C" void fadd_re_host(const double* x, const double* y, double* out, long n) {
  for (long i = 0; i < n; ++i) out[i] = dla_faddeeva_re(x[i], y[i]);
}
