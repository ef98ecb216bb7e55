// FP64 peak micro-benchmark for B200 (sm_100a): the roofline denominators for the
// DLA likelihood kernels.  MEASURED_PEAKS.json only carries HBM and bf16 numbers,
// so the DFMA (vector pipe) and DMMA (mma.sync f64 tensor path) peaks are measured
// here with CUDA events.  Prints one JSON line.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peaks tools/fp64_peaks.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

// ---- DFMA: 16 independent accumulators per thread, ITERS x 16 FMAs -------------
template <int ILP>
__global__ void __launch_bounds__(256) k_dfma(double* out, double a, double b, int iters) {
  double acc[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) acc[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = fma(acc[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += acc[i];
  if (s == 123.456) out[0] = s;
}

// ---- DMMA m8n8k4: NACC independent 8x8 accumulators per warp ---------------------
template <int NACC>
__global__ void __launch_bounds__(256) k_dmma884(double* out, double a, double b, int iters) {
  double c0[NACC], c1[NACC];
#pragma unroll
  for (int i = 0; i < NACC; ++i) { c0[i] = i; c1[i] = -i; }
  double fa = a + threadIdx.x * 1e-9, fb = b;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) {
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                   : "+d"(c0[i]), "+d"(c1[i]) : "d"(fa), "d"(fb));
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += c0[i] + c1[i];
  if (s == 123.456) out[0] = s;
}

// ---- mixed: are the DFMA pipe and the DMMA tensor sub-pipe independent units? ------------------
template <int NMMA, int NFMA>
__global__ void __launch_bounds__(256) k_mixed(double* out, double a, double b, int iters) {
  double c0[NMMA], c1[NMMA], acc[NFMA];
#pragma unroll
  for (int i = 0; i < NMMA; ++i) { c0[i] = i; c1[i] = -i; }
#pragma unroll
  for (int i = 0; i < NFMA; ++i) acc[i] = threadIdx.x * 1e-3 + i;
  double fa = a + threadIdx.x * 1e-9, fb = b;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NMMA; ++i) {
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                   : "+d"(c0[i]), "+d"(c1[i]) : "d"(fa), "d"(fb));
#pragma unroll
      for (int j = 0; j < NFMA / NMMA; ++j) acc[i * (NFMA / NMMA) + j] = fma(acc[i * (NFMA / NMMA) + j], a, b);
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NMMA; ++i) s += c0[i] + c1[i];
#pragma unroll
  for (int i = 0; i < NFMA; ++i) s += acc[i];
  if (s == 123.456) out[0] = s;
}

#ifdef TRY_M16
// ---- DMMA m16n8k8 (sm_90+ shapes) ------------------------------------------------
template <int NACC>
__global__ void __launch_bounds__(256) k_dmma1688(double* out, double a, double b, int iters) {
  double c[NACC][4];
#pragma unroll
  for (int i = 0; i < NACC; ++i) { c[i][0] = i; c[i][1] = -i; c[i][2] = 1; c[i][3] = 2; }
  double a0 = a + threadIdx.x * 1e-9, a1 = a * 2, a2 = a * 3, a3 = a * 4, b0 = b, b1 = b * 0.5;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) {
      asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                   : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                   : "d"(a0), "d"(a1), "d"(a2), "d"(a3), "d"(b0), "d"(b1));
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  if (s == 123.456) out[0] = s;
}
#endif

// ---- throughput of the special functions the producer stage needs --------------
template <int OP>
__global__ void __launch_bounds__(256) k_special(double* out, double a, int iters) {
  double x[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = 1.0 + (threadIdx.x + i) * 1e-3;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (OP == 0) x[i] = 1.0 / (x[i] + a);          // IEEE division
      if (OP == 1) x[i] = exp(-x[i] * a) + 1.0;      // exp
      if (OP == 2) x[i] = log(x[i] + a) + 2.0;       // log
      if (OP == 3) x[i] = sqrt(x[i] + a);            // sqrt
      if (OP == 4) x[i] = __drcp_rn(x[i] + a);       // reciprocal intrinsic
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += x[i];
  if (s == 123.456) out[0] = s;
}

template <typename F>
static double time_ms(F launch, int reps) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int i = 0; i < 3; ++i) launch();
  CK(cudaDeviceSynchronize());
  double best = 1e30, tot = 0;
  for (int r = 0; r < reps; ++r) {
    CK(cudaEventRecord(e0));
    launch();
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
    tot += ms;
  }
  CK(cudaGetLastError());
  return best;
}

int main(int argc, char** argv) {
  int dev = 0; CK(cudaSetDevice(dev));
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, dev));
  int sms = p.multiProcessorCount;
  double* out; CK(cudaMalloc(&out, 8));
  const int iters = 4096;
  const int blocks = sms * 8, threads = 256;
  double nthreads = (double)blocks * threads;
  double nwarps = nthreads / 32;

  double ms_fma8  = time_ms([&] { k_dfma<8><<<blocks, threads>>>(out, 1.0000001, 1e-9, iters); }, 10);
  double ms_fma16 = time_ms([&] { k_dfma<16><<<blocks, threads>>>(out, 1.0000001, 1e-9, iters); }, 10);
  double tf_fma8  = nthreads * iters * 8  * 2 / (ms_fma8  * 1e-3) / 1e12;
  double tf_fma16 = nthreads * iters * 16 * 2 / (ms_fma16 * 1e-3) / 1e12;

  double ms_mma8  = time_ms([&] { k_dmma884<8><<<blocks, threads>>>(out, 1.0000001, 1e-9, iters); }, 10);
  double ms_mma16 = time_ms([&] { k_dmma884<16><<<blocks, threads>>>(out, 1.0000001, 1e-9, iters); }, 10);
  double tf_mma8  = nwarps * iters * 8  * (8.0 * 8 * 4 * 2) / (ms_mma8  * 1e-3) / 1e12;
  double tf_mma16 = nwarps * iters * 16 * (8.0 * 8 * 4 * 2) / (ms_mma16 * 1e-3) / 1e12;
  // occupancy sweep for DMMA: 1 block of 128/256 threads per SM
  double ms_mma_lo = time_ms([&] { k_dmma884<16><<<sms, 128>>>(out, 1.0000001, 1e-9, iters); }, 10);
  double tf_mma_lo = (double)sms * 4 * iters * 16 * 512.0 / (ms_mma_lo * 1e-3) / 1e12;
  double ms_mma_md = time_ms([&] { k_dmma884<16><<<sms, 256>>>(out, 1.0000001, 1e-9, iters); }, 10);
  double tf_mma_md = (double)sms * 8 * iters * 16 * 512.0 / (ms_mma_md * 1e-3) / 1e12;
#ifdef TRY_M16
  double ms_m16 = time_ms([&] { k_dmma1688<8><<<blocks, threads>>>(out, 1.0000001, 1e-9, iters); }, 10);
  double tf_m16 = nwarps * iters * 8 * (16.0 * 8 * 8 * 2) / (ms_m16 * 1e-3) / 1e12;
#else
  double tf_m16 = -1;
#endif
  // mixed DMMA + DFMA: 8 DMMA (4096 flop/warp) + 16 DFMA (1024 flop/warp) and 8 + 32
  double ms_mix1 = time_ms([&] { k_mixed<8, 16><<<blocks, threads>>>(out, 1.0000001, 1e-9, iters); }, 10);
  double tf_mix1 = nwarps * iters * (8 * 512.0 + 16 * 64.0) / (ms_mix1 * 1e-3) / 1e12;
  double ms_mix2 = time_ms([&] { k_mixed<8, 32><<<blocks, threads>>>(out, 1.0000001, 1e-9, iters); }, 10);
  double tf_mix2 = nwarps * iters * (8 * 512.0 + 32 * 64.0) / (ms_mix2 * 1e-3) / 1e12;
  printf("{\"mixed_8dmma_16dfma_tflops\": %.2f, \"mixed_8dmma_32dfma_tflops\": %.2f}\n", tf_mix1, tf_mix2);
  // sustained DFMA (about 3 s) to see the power-capped figure
  double t_sus = 0; int n_sus = 0;
  {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0));
    while (true) {
      for (int i = 0; i < 20; ++i) k_dfma<16><<<blocks, threads>>>(out, 1.0000001, 1e-9, iters);
      n_sus += 20;
      CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
      float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
      t_sus = ms;
      if (ms > 3000) break;
    }
  }
  double tf_fma_sus = nthreads * iters * 16 * 2 * n_sus / (t_sus * 1e-3) / 1e12;
  double t_sus2 = 0; int n_sus2 = 0;
  {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0));
    while (true) {
      for (int i = 0; i < 20; ++i) k_dmma884<16><<<blocks, threads>>>(out, 1.0000001, 1e-9, iters);
      n_sus2 += 20;
      CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
      float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
      t_sus2 = ms;
      if (ms > 3000) break;
    }
  }
  double tf_mma_sus = nwarps * iters * 16 * 512.0 * n_sus2 / (t_sus2 * 1e-3) / 1e12;

  const int it2 = 512;
  double gops[5];
  {
    double ms;
    ms = time_ms([&] { k_special<0><<<blocks, threads>>>(out, 1e-9, it2); }, 5); gops[0] = nthreads * it2 * 8 / (ms * 1e-3) / 1e9;
    ms = time_ms([&] { k_special<1><<<blocks, threads>>>(out, 1e-9, it2); }, 5); gops[1] = nthreads * it2 * 8 / (ms * 1e-3) / 1e9;
    ms = time_ms([&] { k_special<2><<<blocks, threads>>>(out, 1e-9, it2); }, 5); gops[2] = nthreads * it2 * 8 / (ms * 1e-3) / 1e9;
    ms = time_ms([&] { k_special<3><<<blocks, threads>>>(out, 1e-9, it2); }, 5); gops[3] = nthreads * it2 * 8 / (ms * 1e-3) / 1e9;
    ms = time_ms([&] { k_special<4><<<blocks, threads>>>(out, 1e-9, it2); }, 5); gops[4] = nthreads * it2 * 8 / (ms * 1e-3) / 1e9;
  }

  printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz\": %d, "
         "\"dfma_tflops_ilp8\": %.2f, \"dfma_tflops_ilp16\": %.2f, \"dfma_tflops_sustained\": %.2f, "
         "\"dmma884_tflops_acc8\": %.2f, \"dmma884_tflops_acc16\": %.2f, \"dmma884_tflops_sustained\": %.2f, "
         "\"dmma884_tflops_4warps_per_sm\": %.2f, \"dmma884_tflops_8warps_per_sm\": %.2f, "
         "\"dmma1688_tflops\": %.2f, "
         "\"gops_div\": %.1f, \"gops_exp\": %.1f, \"gops_log\": %.1f, \"gops_sqrt\": %.1f, \"gops_rcp\": %.1f}\n",
         p.name, sms, p.clockRate, tf_fma8, tf_fma16, tf_fma_sus, tf_mma8, tf_mma16, tf_mma_sus,
         tf_mma_lo, tf_mma_md, tf_m16, gops[0], gops[1], gops[2], gops[3], gops[4]);
  return 0;
}
