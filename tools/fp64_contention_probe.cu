// FP64 pipe contention probe for B200 (sm_100a): how long does a dependent chain of scalar DFMAs take on an SM
// sub-partition whose FP64 pipe is being saturated by k other warps issuing back-to-back DMMAs (m8n8k4.f64)?
// This is the "phase A under the other units' phase B" situation of sample_likelihood_kernel.
// One CTA per SM, 16 warps: warp w sits on sub-partition w % 4.  Warps 0..3 run the DFMA chain (ILP independent
// chains per thread) and time it with clock64; warps with (w / 4) in 1..k run DMMAs for the whole duration.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_contention_probe tools/fp64_contention_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

template <int ILP>
__global__ void __launch_bounds__(512, 1) k_probe(long long* cycles, double* sink, int k, int chain, double a, double b) {
  const int warp = threadIdx.x >> 5;
  const int role = warp >> 2;  // 0: chain, 1..3: DMMA load
  __shared__ volatile int done;
  if (threadIdx.x == 0) done = 0;
  __syncthreads();
  if (role == 0) {
    double x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x * 1e-3 + i;
    // let the DMMA warps fill the pipe first
    long long t0 = clock64();
    while (clock64() - t0 < 20000) {}
    t0 = clock64();
    for (int it = 0; it < chain; ++it) {
#pragma unroll
      for (int i = 0; i < ILP; ++i) x[i] = fma(x[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += x[i];
    long long t1 = clock64();
    if (s == 123.456) sink[0] = s;
    if ((threadIdx.x & 31) == 0) cycles[blockIdx.x * 4 + warp] = t1 - t0;
    __syncwarp();
    if (threadIdx.x == 0) done = 1;
  } else if (role <= k) {
    double c0[8], c1[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { c0[i] = i; c1[i] = -i; }
    double fa = a + threadIdx.x * 1e-9, fb = b;
    while (!done) {
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int i = 0; i < 8; ++i)
          asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                       : "+d"(c0[i]), "+d"(c1[i]) : "d"(fa), "d"(fb));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c0[i] + c1[i];
    if (s == 123.456) sink[1] = s;
  }
}

template <int ILP>
static void run(int sms, long long* d_cycles, double* sink) {
  const int chain = 2000;
  for (int k = 0; k <= 3; ++k) {
    k_probe<ILP><<<sms, 512>>>(d_cycles, sink, k, chain, 0.999, 1e-3);
    CK(cudaDeviceSynchronize());
    long long h[4];
    CK(cudaMemcpy(h, d_cycles, sizeof(h), cudaMemcpyDeviceToHost));
    double avg = (h[0] + h[1] + h[2] + h[3]) / 4.0 / chain;
    printf("{\"ilp\": %d, \"dmma_warps_per_smsp\": %d, \"cycles_per_dependent_step\": %.1f, \"cycles_per_dfma\": %.1f}\n", ILP, k, avg,
           avg / ILP);
  }
}

int main() {
  CK(cudaSetDevice(0));
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  long long* d_cycles; CK(cudaMalloc(&d_cycles, p.multiProcessorCount * 4 * sizeof(long long)));
  double* sink; CK(cudaMalloc(&sink, 16));
  run<1>(p.multiProcessorCount, d_cycles, sink);
  run<2>(p.multiProcessorCount, d_cycles, sink);
  run<4>(p.multiProcessorCount, d_cycles, sink);
  return 0;
}
