// Single-stream interleave probe for B200 (sm_100a): every warp issues NDM DMMAs (m8n8k4.f64) followed by one
// scalar DFMA of one of ILP dependent chains, repeatedly.  If the shared FP64 pipe stays busy, the DMMA rate
// drops only by the DFMAs' own pipe time (2 cycles each against 16 per DMMA); if a dependent DFMA stalls its
// warp behind the other warps' queued DMMAs, it drops much further.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_interleave_probe tools/fp64_interleave_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

template <int NDM, int ILP, bool SCALAR>
__global__ void __launch_bounds__(256, 2) k_mix(double* out, int iters, double a, double b) {
  double c0[16], c1[16], x[ILP];
#pragma unroll
  for (int i = 0; i < 16; ++i) { c0[i] = i; c1[i] = -i; }
#pragma unroll
  for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x * 1e-3 + i;
  double fa = a + threadIdx.x * 1e-9, fb = b;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int g = 0; g < 16 / NDM; ++g) {
#pragma unroll
      for (int i = 0; i < NDM; ++i)
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                     : "+d"(c0[g * NDM + i]), "+d"(c1[g * NDM + i]) : "d"(fa), "d"(fb));
      if (SCALAR) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(x[g % ILP]) : "d"(a), "d"(b));
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += c0[i] + c1[i];
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += x[i];
  if (s == 123.456) out[0] = s;
}

template <int NDM, int ILP, bool SCALAR>
static void run(int sms, int ctas, double* out) {
  const int iters = 20000;
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  k_mix<NDM, ILP, SCALAR><<<sms * ctas, 256>>>(out, iters, 0.999, 1e-3);
  CK(cudaEventRecord(e0));
  k_mix<NDM, ILP, SCALAR><<<sms * ctas, 256>>>(out, iters, 0.999, 1e-3);
  CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
  double flops = (double)sms * ctas * 8 * iters * 16.0 * 512.0;
  printf("{\"dmma_per_dfma\": %d, \"ilp\": %d, \"scalar\": %d, \"warps_per_smsp\": %d, \"dmma_tflops\": %.2f}\n", NDM, ILP, (int)SCALAR,
         2 * ctas, flops / (ms * 1e-3) / 1e12);
}

int main() {
  CK(cudaSetDevice(0));
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  const int sms = p.multiProcessorCount;
  double* out; CK(cudaMalloc(&out, 8));
  for (int c = 1; c <= 2; ++c) {
    run<2, 1, false>(sms, c, out);
    run<2, 1, true>(sms, c, out);
    run<2, 2, true>(sms, c, out);
    run<2, 4, true>(sms, c, out);
    run<4, 1, true>(sms, c, out);
    run<4, 2, true>(sms, c, out);
    run<1, 2, true>(sms, c, out);
    run<1, 8, true>(sms, c, out);
  }
  return 0;
}
