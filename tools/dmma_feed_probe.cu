// DMMA feed probe for B200 (sm_100a): how close can an mma.sync.m8n8k4.f64 loop get to the FP64 peak as a
// function of (a) warps per SM sub-partition, (b) where the operands come from (registers / shared memory in
// the layout of sample_likelihood_kernel's phase B), (c) a CTA barrier every 60 DMMAs per warp.
// Prints one JSON line per variant.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_feed_probe tools/dmma_feed_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

constexpr int KC = 16, WSTRIDE = 20, PSTRIDE = 244, TS = 32;

// MODE 0: operands in registers.  MODE 1: operands from shared memory (phase-B addressing).
// BARRIER: __syncthreads() after every panel (60 DMMAs per warp).   NBLK: column blocks per warp (7 or 8 -> 7.5 avg; here 8 = 64 DMMAs)
template <int MODE, bool BARRIER, int KCT>
__global__ void __launch_bounds__(256, 2) k_feed(double* out, int panels, double seed) {
  extern __shared__ double smem[];
  double* s_W = smem;                       // [TS][KCT + 4]
  double* s_P = smem + TS * (KCT + 4);        // [KCT][PSTRIDE]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < TS * (KCT + 4) + KCT * PSTRIDE; i += blockDim.x) smem[i] = seed + i * 1e-9;
  __syncthreads();
  const int grp = lane >> 2, tig = lane & 3;
  const int rq = warp >> 2, cq = warp & 3;
  double acc[2][8][2];
#pragma unroll
  for (int m = 0; m < 2; ++m)
#pragma unroll
    for (int nb = 0; nb < 8; ++nb) acc[m][nb][0] = acc[m][nb][1] = 0.0;
  const double* arow = s_W + (rq * 16 + grp) * (KCT + 4) + tig;
  const double* brow = s_P + tig * PSTRIDE + cq * 56 + grp;
  double ra = seed + lane, rb = seed * 0.5;
  for (int p = 0; p < panels; ++p) {
#pragma unroll
    for (int kb = 0; kb < KCT / 4; ++kb) {
      double a[2];
      if (MODE == 1) {
        a[0] = arow[kb * 4];
        a[1] = arow[8 * (KCT + 4) + kb * 4];
      } else { a[0] = ra; a[1] = rb; }
#pragma unroll
      for (int nb = 0; nb < 8; ++nb) {
        if (nb == 7 && (warp & 1)) continue;  // 7.5 blocks per warp on average, like the kernel
        const double b = MODE == 1 ? brow[kb * 4 * PSTRIDE + nb * 8] : rb;
        dmma884(acc[0][nb][0], acc[0][nb][1], a[0], b);
        dmma884(acc[1][nb][0], acc[1][nb][1], a[1], b);
      }
    }
    if (BARRIER) __syncthreads();
  }
  double s = 0;
#pragma unroll
  for (int m = 0; m < 2; ++m)
#pragma unroll
    for (int nb = 0; nb < 8; ++nb) s += acc[m][nb][0] + acc[m][nb][1];
  if (s == 123.456) out[0] = s;
}

template <typename F>
static double time_ms(F launch, int reps) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int i = 0; i < 3; ++i) launch();
  CK(cudaDeviceSynchronize());
  double best = 1e30;
  for (int r = 0; r < reps; ++r) {
    CK(cudaEventRecord(e0));
    launch();
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  CK(cudaGetLastError());
  return best;
}

template <int MODE, bool BARRIER, int KCT>
static void run(const char* name, int sms, int ctas_per_sm, double* out) {
  const int panels = 4096 * 16 / KCT;
  const size_t smem = (size_t)(TS * (KCT + 4) + KCT * PSTRIDE) * sizeof(double);
  CK(cudaFuncSetAttribute(k_feed<MODE, BARRIER, KCT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int blocks = sms * ctas_per_sm;
  double ms = time_ms([&] { k_feed<MODE, BARRIER, KCT><<<blocks, 256, smem>>>(out, panels, 1.0); }, 5);
  // DMMAs per warp per panel: KCT/4 * (2 * 7.5) ; 512 flops each
  double flops = (double)blocks * 8 * panels * (KCT / 4) * 15.0 * 512.0;
  printf("{\"variant\": \"%s\", \"kc\": %d, \"ctas_per_sm\": %d, \"warps_per_smsp\": %d, \"tflops\": %.2f}\n", name, KCT,
         ctas_per_sm, 2 * ctas_per_sm, flops / (ms * 1e-3) / 1e12);
}

int main() {
  CK(cudaSetDevice(0));
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  const int sms = p.multiProcessorCount;
  double* out; CK(cudaMalloc(&out, 8));
  for (int c = 1; c <= 2; ++c) {
    run<0, false, 16>("reg", sms, c, out);
    run<0, true, 16>("reg+barrier", sms, c, out);
    run<1, false, 16>("smem", sms, c, out);
    run<1, true, 16>("smem+barrier", sms, c, out);
    run<1, true, 32>("smem+barrier", sms, c, out);
  }
  return 0;
}
