// dla_b200.cu : C-ABI of libdla_b200.so (see include/dla_b200.h).
// Single translation unit: the kernels live in the *.cuh headers next to this file.
#include "../../include/dla_b200.h"

#include <math.h>
#include <string.h>

#include <algorithm>
#include <limits>
#include <memory>
#include <string>
#include <vector>

#include <nvtx3/nvToolsExt.h>

#include "common.cuh"
#include "evidence_kernel.cuh"
#include "likelihood_kernel.cuh"
#include "prep_kernel.cuh"
#include "voigt_kernel.cuh"
#include "zqso_kernel.cuh"

namespace dla {

// ------------------------------------------------------------------------------------------
// runtime
// ------------------------------------------------------------------------------------------
static thread_local std::string t_error;
Runtime& runtime() {
  static Runtime rt;
  return rt;
}
BufferPool& buffer_pool() {
  static BufferPool pool;
  return pool;
}
void set_error(const std::string& msg) { t_error = msg; }
int fail(const std::string& msg) {
  t_error = msg;
  return 1;
}

static int init_device(int device) {
  Runtime& rt = runtime();
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0)
    return fail(std::string("no usable CUDA device (") + cudaGetErrorString(e) +
                "); libdla_b200 has no CPU fallback");
  if (device < 0 || device >= count) return fail("dla_init: device index out of range");
  if (rt.ready && rt.device == device) return 0;
  DLA_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  DLA_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10) return fail(std::string("device '") + prop.name + "' is not sm_100-class; this library is built for sm_100a only");
  if (rt.stream) {
    buffer_pool().flush();  // pooled allocations belong to the device being left
    cudaStreamDestroy(rt.stream);
    if (rt.copy_stream) cudaStreamDestroy(rt.copy_stream);
    cudaEventDestroy(rt.ev_begin);
    cudaEventDestroy(rt.ev_end);
  }
  DLA_CUDA(cudaStreamCreateWithFlags(&rt.stream, cudaStreamNonBlocking));
  DLA_CUDA(cudaStreamCreateWithFlags(&rt.copy_stream, cudaStreamNonBlocking));
  DLA_CUDA(cudaEventCreate(&rt.ev_begin));
  DLA_CUDA(cudaEventCreate(&rt.ev_end));
  rt.device = device;
  rt.sm_count = prop.multiProcessorCount;
  rt.smem_optin = prop.sharedMemPerBlockOptin;

  // pair tables of the likelihood kernel
  uint8_t pi[LK_NBLK_PAIR * 8], pj[LK_NBLK_PAIR * 8];
  memset(pi, 0, sizeof(pi));
  memset(pj, 0, sizeof(pj));
  int c = 0;
  for (int i = 0; i < LK_K; ++i)
    for (int j = 0; j <= i; ++j, ++c) {
      pi[c] = (uint8_t)i;
      pj[c] = (uint8_t)j;
    }
  DLA_CUDA(cudaMemcpyToSymbol(c_pair_i, pi, sizeof(pi)));
  DLA_CUDA(cudaMemcpyToSymbol(c_pair_j, pj, sizeof(pj)));
  DLA_CUDA(cudaFuncSetAttribute(sample_likelihood_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LK_SMEM_BYTES));
  rt.ready = true;
  return 0;
}

int ensure_ready() {
  Runtime& rt = runtime();
  if (rt.ready) {
    cudaError_t e = cudaSetDevice(rt.device);
    if (e != cudaSuccess) return fail(std::string("cudaSetDevice failed: ") + cudaGetErrorString(e));
    return 0;
  }
  return init_device(0);
}

cudaError_t KernelTimer::begin() {
  active = true;
  return cudaEventRecord(runtime().ev_begin, runtime().stream);
}
cudaError_t KernelTimer::end() {
  Runtime& rt = runtime();
  cudaError_t e = cudaEventRecord(rt.ev_end, rt.stream);
  if (e != cudaSuccess) return e;
  e = cudaEventSynchronize(rt.ev_end);
  if (e != cudaSuccess) return e;
  float ms = 0.f;
  e = cudaEventElapsedTime(&ms, rt.ev_begin, rt.ev_end);
  rt.last_kernel_ms = ms;
  active = false;
  return e;
}

// small fill kernels
__global__ void fill_double_kernel(double* p, size_t n, double value) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = value;
}
__global__ void fill_rows_kernel(int32_t* rows, int S, int nrows) {
  // row 0 = identity (the sample's own profile); rows 1.. = resampled indices, zero-initialised
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < (size_t)S * nrows) rows[i] = i < (size_t)S ? (int32_t)i : 0;
}
__global__ void product_rows_kernel(const double* cache, int ld, int n, int nrows, double* out) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  double a = cache[p];
  for (int r = 1; r < nrows; ++r) a = a * cache[(size_t)r * ld + p];
  out[p] = a;
}
__global__ void keep_to_uidx_kernel(const uint8_t* keep, int n_u, int32_t* uidx) {
  // single thread: tiny; builds the compaction map from a keep-mask
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    int q = 0;
    for (int i = 0; i < n_u; ++i)
      if (keep[i]) uidx[q++] = i;
  }
}

static inline size_t round_up(size_t a, size_t b) { return (a + b - 1) / b * b; }

}  // namespace dla

using namespace dla;

// ------------------------------------------------------------------------------------------
// handles
// ------------------------------------------------------------------------------------------
struct dla_model {
  DevBuf<double> rest, mu, M, log_omega;
  ModelDev dev;
  int device = -1;  // every handle remembers its device; calls check it against the one dla_init selected
};

struct dla_spectrum {
  int device = -1;
  int n_raw = 0, n_u = 0, n = 0, k = 0, width = 0, broadening = 1, n_abs = 0, ld = 0;
  int lls_break = 0;  // profiles include the Lyman-limit break (voigt_lls.py)
  double z_qso = 0.0;
  double scalars[8] = {0};
  bool from_raw = false;
  // raw-side buffers
  DevBuf<double> X, Y, V, scratch;
  DevBuf<uint8_t> mask, ind_unmasked, ind;
  // prepared arrays
  DevBuf<double> x, y, v, this_wl, mu, omega2, M, unmasked_wl, wl_abs, padded_wl, d_scalars;
  DevBuf<int32_t> uidx, qmap;
  // work buffers (grown on demand)
  DevBuf<double> cache, prod, z_dev, nhi_dev, uniforms, raw_ll, sample_ll, log_ev, cdf;
  DevBuf<int32_t> rows, sel, pos_a, pos_b, rows0_c, rows1_c;
  DevBuf<double> raw_slots;
  DevBuf<int> alive;  // [0] alive, [1] status, [4 + level] samples evaluated at that level
  DevBuf<CompactTask> compact_desc;
  DevBuf<ScatterTask> scatter_desc;
  DevBuf<LikelihoodSpectrum> lk_desc;
  DevBuf<EvidenceLevel> ev_desc;
  DevBuf<AbsorptionGrid> grid_desc;
  // Gram basis [P | M] of the likelihood kernel (built on first use)
  DevBuf<double> basis;
  DevBuf<GramBasisTask> basis_desc;
  bool basis_ready = false;
};

// ------------------------------------------------------------------------------------------
// launch helpers (single spectrum = batch of one)
// ------------------------------------------------------------------------------------------
// profiles of `max_samples` absorbers for each of `num_spectra` grids; num_lines = 3 (the published
// configuration) runs the fully unrolled instantiation
static int launch_voigt_grids(const AbsorptionGrid* d_grids, int max_samples, int num_spectra, int num_lines, int broadening) {
  Runtime& rt = runtime();
  dim3 grid((max_samples + VG_WARPS - 1) / VG_WARPS, num_spectra);
  if (num_lines == 3)
    voigt_profile_kernel<3><<<grid, VG_WARPS * 32, 0, rt.stream>>>(d_grids, num_lines, broadening);
  else
    voigt_profile_kernel<0><<<grid, VG_WARPS * 32, 0, rt.stream>>>(d_grids, num_lines, broadening);
  DLA_LAUNCHED();
  return 0;
}

static int launch_voigt(dla_spectrum* sp, const double* d_z, const double* d_nhi, int num_samples, int num_lines,
                        double* d_out, int ld) {
  Runtime& rt = runtime();
  DLA_REQUIRE(num_lines >= 1 && num_lines <= LYMAN_NUM_LINES, "num_lines must be in [1, 31]");
  AbsorptionGrid g;
  g.wl = sp->wl_abs.p;
  g.uidx = sp->uidx.p;
  g.out = d_out;
  g.n_in = sp->n_abs;
  g.n_out = sp->n;
  g.ld = ld;
  g.num_samples = num_samples;
  g.z = d_z;
  g.nhi = d_nhi;
  g.pair_offset = 0;
  g.lls_break = sp->lls_break;
  const int n_u = sp->broadening ? sp->n_abs - 2 * INSTRUMENT_WIDTH : sp->n_abs;
  DLA_CUDA(sp->qmap.ensure(std::max(n_u, 1)));
  g.qmap = sp->qmap.p;
  DLA_CUDA(sp->grid_desc.ensure(1));
  DLA_CUDA(cudaMemcpyAsync(sp->grid_desc.p, &g, sizeof(g), cudaMemcpyHostToDevice, rt.stream));
  build_qmap_kernel<<<1, 256, 0, rt.stream>>>(sp->grid_desc.p, sp->broadening);
  DLA_LAUNCHED();
  return launch_voigt_grids(sp->grid_desc.p, num_samples, 1, num_lines, sp->broadening);
}

static int ensure_cache(dla_spectrum* sp, size_t rows) {
  DLA_CUDA(sp->cache.ensure(rows * (size_t)sp->ld));
  return 0;
}

// Gram basis of the spectrum's interpolated M (pair products + M columns), streamed by the
// likelihood kernel in 16-pixel panels
static int ensure_basis(dla_spectrum* sp) {
  if (sp->basis_ready) return 0;
  Runtime& rt = runtime();
  const size_t rows = round_up((size_t)std::max(sp->n, 1), LK_KC);
  DLA_CUDA(sp->basis.ensure(rows * LK_PSTRIDE));
  DLA_CUDA(sp->basis_desc.ensure(1));
  GramBasisTask t;
  t.M = sp->M.p;
  t.P = sp->basis.p;
  t.n = sp->n;
  DLA_CUDA(cudaMemcpyAsync(sp->basis_desc.p, &t, sizeof(t), cudaMemcpyHostToDevice, rt.stream));
  gram_basis_kernel<<<dim3((unsigned)((rows + 7) / 8), 1), 256, 0, rt.stream>>>(sp->basis_desc.p);
  DLA_LAUNCHED();
  sp->basis_ready = true;
  return 0;
}

// ------------------------------------------------------------------------------------------
// library / device
// ------------------------------------------------------------------------------------------
extern "C" int dla_init(int device) { return init_device(device); }

extern "C" int dla_device_count(void) {
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess) return 0;
  return count;
}

extern "C" const char* dla_last_error(void) { return t_error.c_str(); }
extern "C" const char* dla_version(void) { return "dla_b200 0.1 (sm_100a)"; }
extern "C" double dla_last_kernel_ms(void) { return runtime().last_kernel_ms; }
extern "C" long long dla_kernel_launch_count(void) { return runtime().launches; }

// ------------------------------------------------------------------------------------------
// FP64 peak probes (roofline denominators)
// ------------------------------------------------------------------------------------------
namespace dla {
__global__ void __launch_bounds__(256) peak_dfma_kernel(double* out, double a, double b, int iters) {
  double acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = fma(acc[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += acc[i];
  if (s == 123.456) out[0] = s;
}
__global__ void __launch_bounds__(256) peak_dmma_kernel(double* out, double a, double b, int iters) {
  double c0[8], c1[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { c0[i] = i; c1[i] = -i; }
  const double fa = a + threadIdx.x * 1e-9, fb = b;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) dmma884(c0[i], c1[i], fa, fb);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += c0[i] + c1[i];
  if (s == 123.456) out[0] = s;
}
}  // namespace dla

extern "C" int dla_measure_fp64_peaks(double* dfma_tflops, double* dmma_tflops) {
  DLA_CHECK_READY();
  Runtime& rt = runtime();
  DevBuf<double> sink;
  DLA_CUDA(sink.alloc(1));
  const int iters = 4096, blocks = rt.sm_count * 8, threads = 256;
  cudaEvent_t e0, e1;
  DLA_CUDA(cudaEventCreate(&e0));
  DLA_CUDA(cudaEventCreate(&e1));
  double best[2] = {1e30, 1e30};
  for (int which = 0; which < 2; ++which) {
    for (int rep = 0; rep < 8; ++rep) {
      DLA_CUDA(cudaEventRecord(e0, rt.stream));
      if (which == 0) peak_dfma_kernel<<<blocks, threads, 0, rt.stream>>>(sink.p, 1.0000001, 1e-9, iters);
      else peak_dmma_kernel<<<blocks, threads, 0, rt.stream>>>(sink.p, 1.0000001, 1e-9, iters);
      DLA_CUDA(cudaGetLastError());
      DLA_CUDA(cudaEventRecord(e1, rt.stream));
      DLA_CUDA(cudaEventSynchronize(e1));
      float ms = 0.f;
      DLA_CUDA(cudaEventElapsedTime(&ms, e0, e1));
      if (rep >= 2 && ms < best[which]) best[which] = ms;
    }
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  const double nthreads = (double)blocks * threads;
  if (dfma_tflops) *dfma_tflops = nthreads * iters * 16 * 2 / (best[0] * 1e-3) / 1e12;
  if (dmma_tflops) *dmma_tflops = (nthreads / 32) * iters * 8 * 512.0 / (best[1] * 1e-3) / 1e12;
  return 0;
}

// ------------------------------------------------------------------------------------------
// a1: voigt
// ------------------------------------------------------------------------------------------
static int voigt_batch_impl(const double* wavelengths, int n_in, const double* nhis, const double* z_dlas, int S,
                            int num_lines, int broadening, int lls_break, double* out);

extern "C" int dla_voigt_absorption_batch(const double* wavelengths, int n_in, const double* nhis,
                                          const double* z_dlas, int S, int num_lines, int broadening, double* out) {
  return voigt_batch_impl(wavelengths, n_in, nhis, z_dlas, S, num_lines, broadening, 0, out);
}

extern "C" int dla_voigt_lls_absorption_batch(const double* wavelengths, int n_in, const double* nhis,
                                              const double* z_llss, int S, int num_lines, int broadening, double* out) {
  return voigt_batch_impl(wavelengths, n_in, nhis, z_llss, S, num_lines, broadening, 1, out);
}

extern "C" int dla_spectrum_set_lls_break(dla_spectrum* spec, int on) {
  DLA_REQUIRE(spec, "null spectrum");
  spec->lls_break = on ? 1 : 0;
  return 0;
}

static int voigt_batch_impl(const double* wavelengths, int n_in, const double* nhis, const double* z_dlas, int S,
                            int num_lines, int broadening, int lls_break, double* out) {
  DLA_CHECK_READY();
  Runtime& rt = runtime();
  DLA_REQUIRE(wavelengths && nhis && z_dlas && out, "null pointer argument");
  const int n_out = broadening ? n_in - 2 * INSTRUMENT_WIDTH : n_in;
  DLA_REQUIRE(n_out >= 1 && S >= 1, "voigt_absorption: need at least 7 wavelengths with broadening and S >= 1");
  dla_spectrum sp;
  sp.n_abs = n_in;
  sp.n = n_out;
  sp.broadening = broadening ? 1 : 0;
  sp.lls_break = lls_break;
  sp.ld = n_out;
  DLA_CUDA(sp.wl_abs.alloc(n_in));
  DLA_CUDA(sp.wl_abs.upload(wavelengths, n_in, rt.stream));
  std::vector<int32_t> iota(n_out);
  for (int i = 0; i < n_out; ++i) iota[i] = i;
  DLA_CUDA(sp.uidx.alloc(n_out));
  DLA_CUDA(sp.uidx.upload(iota.data(), n_out, rt.stream));
  DLA_CUDA(sp.z_dev.alloc(S));
  DLA_CUDA(sp.nhi_dev.alloc(S));
  DLA_CUDA(sp.z_dev.upload(z_dlas, S, rt.stream));
  DLA_CUDA(sp.nhi_dev.upload(nhis, S, rt.stream));
  DLA_CUDA(sp.cache.alloc((size_t)S * n_out));
  KernelTimer timer;
  DLA_CUDA(timer.begin());
  int rc = launch_voigt(&sp, sp.z_dev.p, sp.nhi_dev.p, S, num_lines, sp.cache.p, n_out);
  if (rc) return rc;
  DLA_CUDA(timer.end());
  DLA_CUDA(sp.cache.download(out, (size_t)S * n_out, rt.stream));
  DLA_CUDA(cudaStreamSynchronize(rt.stream));
  return 0;
}

extern "C" int dla_voigt_absorption(const double* wavelengths, int n_in, double nhi, double z_dla, int num_lines,
                                    int broadening, double* out) {
  return dla_voigt_absorption_batch(wavelengths, n_in, &nhi, &z_dla, 1, num_lines, broadening, out);
}

extern "C" int dla_faddeeva_re(const double* x, const double* y, int n, double* out) {
  DLA_CHECK_READY();
  Runtime& rt = runtime();
  DLA_REQUIRE(x && y && out && n >= 1, "bad argument");
  DevBuf<double> dx, dy, dout;
  DLA_CUDA(dx.alloc(n));
  DLA_CUDA(dy.alloc(n));
  DLA_CUDA(dout.alloc(n));
  DLA_CUDA(dx.upload(x, n, rt.stream));
  DLA_CUDA(dy.upload(y, n, rt.stream));
  faddeeva_kernel<<<(n + 255) / 256, 256, 0, rt.stream>>>(dx.p, dy.p, dout.p, n);
  DLA_LAUNCHED();
  DLA_CUDA(dout.download(out, n, rt.stream));
  DLA_CUDA(cudaStreamSynchronize(rt.stream));
  return 0;
}

// ------------------------------------------------------------------------------------------
// a2: effective optical depth
// ------------------------------------------------------------------------------------------
extern "C" int dla_effective_optical_depth(const double* wavelengths, int n, double beta, double tau_0, double z_qso,
                                           int num_forest_lines, double* out) {
  DLA_CHECK_READY();
  Runtime& rt = runtime();
  DLA_REQUIRE(wavelengths && out && n >= 1, "bad argument");
  DLA_REQUIRE(num_forest_lines >= 1 && num_forest_lines <= LYMAN_NUM_LINES, "num_forest_lines must be in [1, 31]");
  DevBuf<double> dw, dout;
  const size_t total = (size_t)n * num_forest_lines;
  DLA_CUDA(dw.alloc(n));
  DLA_CUDA(dout.alloc(total));
  DLA_CUDA(dw.upload(wavelengths, n, rt.stream));
  effective_optical_depth_kernel<<<(unsigned)((total + 255) / 256), 256, 0, rt.stream>>>(dw.p, n, beta, tau_0, z_qso,
                                                                                       num_forest_lines, dout.p);
  DLA_LAUNCHED();
  DLA_CUDA(dout.download(out, total, rt.stream));
  DLA_CUDA(cudaStreamSynchronize(rt.stream));
  return 0;
}

// ------------------------------------------------------------------------------------------
// a5: generic low-rank mvn
// ------------------------------------------------------------------------------------------
extern "C" int dla_log_mvnpdf_low_rank(const double* y, const double* mu, const double* M, const double* d, int n, int k,
                                       double* out) {
  DLA_CHECK_READY();
  Runtime& rt = runtime();
  DLA_REQUIRE(y && mu && M && d && out, "null pointer argument");
  DLA_REQUIRE(n >= 1 && k >= 1 && k <= LG_MAXK, "log_mvnpdf_low_rank: need n >= 1 and 1 <= k <= 64");
  DevBuf<double> dy, dmu, dM, dd, dout;
  DLA_CUDA(dy.alloc(n));
  DLA_CUDA(dmu.alloc(n));
  DLA_CUDA(dM.alloc((size_t)n * k));
  DLA_CUDA(dd.alloc(n));
  DLA_CUDA(dout.alloc(1));
  DLA_CUDA(dy.upload(y, n, rt.stream));
  DLA_CUDA(dmu.upload(mu, n, rt.stream));
  DLA_CUDA(dM.upload(M, (size_t)n * k, rt.stream));
  DLA_CUDA(dd.upload(d, n, rt.stream));
  log_mvnpdf_low_rank_kernel<<<1, 256, 0, rt.stream>>>(dy.p, dmu.p, dM.p, dd.p, n, k, dout.p);
  DLA_LAUNCHED();
  DLA_CUDA(dout.download(out, 1, rt.stream));
  DLA_CUDA(cudaStreamSynchronize(rt.stream));
  return 0;
}

// ------------------------------------------------------------------------------------------
// a4: model
// ------------------------------------------------------------------------------------------
extern "C" int dla_model_create(const double* rest_wavelengths, const double* mu, const double* M,
                                const double* log_omega, int n_rest, int k, double log_c_0, double log_tau_0,
                                double log_beta, double prev_tau_0, double prev_beta, dla_model** out) {
  DLA_CHECK_READY();
  Runtime& rt = runtime();
  DLA_REQUIRE(rest_wavelengths && mu && M && log_omega && out, "null pointer argument");
  DLA_REQUIRE(n_rest >= 2 && k >= 1, "model needs at least two grid points");
  std::unique_ptr<dla_model> m(new dla_model());
  DLA_CUDA(m->rest.alloc(n_rest));
  DLA_CUDA(m->mu.alloc(n_rest));
  DLA_CUDA(m->M.alloc((size_t)n_rest * k));
  DLA_CUDA(m->log_omega.alloc(n_rest));
  DLA_CUDA(m->rest.upload(rest_wavelengths, n_rest, rt.stream));
  DLA_CUDA(m->mu.upload(mu, n_rest, rt.stream));
  DLA_CUDA(m->M.upload(M, (size_t)n_rest * k, rt.stream));
  DLA_CUDA(m->log_omega.upload(log_omega, n_rest, rt.stream));
  DLA_CUDA(cudaStreamSynchronize(rt.stream));
  m->dev.rest_wavelengths = m->rest.p;
  m->dev.mu = m->mu.p;
  m->dev.M = m->M.p;
  m->dev.log_omega = m->log_omega.p;
  m->dev.n_rest = n_rest;
  m->dev.k = k;
  m->dev.log_c_0 = log_c_0;
  m->dev.log_tau_0 = log_tau_0;
  m->dev.log_beta = log_beta;
  m->dev.prev_tau_0 = prev_tau_0;
  m->dev.prev_beta = prev_beta;
  m->device = rt.device;
  *out = m.release();
  return 0;
}

extern "C" int dla_model_destroy(dla_model* model) {
  delete model;
  return 0;
}

// NullGP.get_interp (null_gp.py:179-242) on caller-supplied pixels: x rest wavelengths (inside the model grid),
// wavelengths observed, both (n); outputs this_mu (n), this_M (n, k) row-major, this_omega2 (n)
extern "C" int dla_model_interp(const dla_model* model, int num_forest_lines, const double* x, const double* wavelengths,
                                int n, double z_qso, double* this_mu, double* this_M, double* this_omega2) {
  DLA_CHECK_READY();
  Runtime& rt = runtime();
  DLA_REQUIRE(model && x && wavelengths && this_mu && this_M && this_omega2, "null pointer argument");
  DLA_REQUIRE(model->device == rt.device, "the model lives on another device than the one selected by dla_init");
  DLA_REQUIRE(n >= 0, "negative length");
  DLA_REQUIRE(num_forest_lines >= 1 && num_forest_lines <= LYMAN_NUM_LINES, "num_forest_lines must be in [1, 31]");
  if (n == 0) return 0;
  const int k = model->dev.k;
  DevBuf<double> dx, dw, dmu, dM, dom;
  DLA_CUDA(dx.alloc(n));
  DLA_CUDA(dw.alloc(n));
  DLA_CUDA(dmu.alloc(n));
  DLA_CUDA(dM.alloc((size_t)n * k));
  DLA_CUDA(dom.alloc(n));
  DLA_CUDA(dx.upload(x, n, rt.stream));
  DLA_CUDA(dw.upload(wavelengths, n, rt.stream));
  interp_model_kernel<<<(n + 127) / 128, 128, 0, rt.stream>>>(model->dev, num_forest_lines, dx.p, dw.p, n, z_qso, dmu.p,
                                                               dM.p, dom.p);
  DLA_LAUNCHED();
  DLA_CUDA(dmu.download(this_mu, n, rt.stream));
  DLA_CUDA(dM.download(this_M, (size_t)n * k, rt.stream));
  DLA_CUDA(dom.download(this_omega2, n, rt.stream));
  DLA_CUDA(cudaStreamSynchronize(rt.stream));
  return 0;
}

static PrepParams to_prep_params(const dla_params* p, int normalize) {
  PrepParams P;
  P.min_lambda = p->min_lambda;
  P.max_lambda = p->max_lambda;
  P.norm_min_lambda = p->normalization_min_lambda;
  P.norm_max_lambda = p->normalization_max_lambda;
  P.pixel_spacing = p->pixel_spacing;
  P.lya_wavelength = p->lya_wavelength;
  P.lyman_limit = p->lyman_limit;
  P.max_z_cut = p->max_z_cut;
  P.min_z_cut = p->min_z_cut;
  P.width = p->width;
  P.num_forest_lines = p->num_forest_lines;
  P.broadening = p->broadening;
  P.normalize = normalize;
  return P;
}

// ------------------------------------------------------------------------------------------
// a3/a4: spectrum preparation
// ------------------------------------------------------------------------------------------
extern "C" int dla_spectrum_create(const dla_model* model, const dla_params* params, const double* X, const double* Y,
                                   const double* V, const uint8_t* pixel_mask, int n_raw, double z_qso, int normalize,
                                   dla_spectrum** out) {
  DLA_CHECK_READY();
  Runtime& rt = runtime();
  DLA_REQUIRE(model && params && X && Y && V && pixel_mask && out, "null pointer argument");
  DLA_REQUIRE(n_raw >= 1, "empty spectrum");
  DLA_REQUIRE(model->device == rt.device, "the model lives on another device than the one selected by dla_init");
  DLA_REQUIRE(params->width == INSTRUMENT_WIDTH, "instrument profile width must be 3");
  DLA_REQUIRE(params->num_forest_lines >= 1 && params->num_forest_lines <= LYMAN_NUM_LINES,
              "num_forest_lines must be in [1, 31]");
  std::unique_ptr<dla_spectrum> sp(new dla_spectrum());
  const int w = params->width;
  const int k = model->dev.k;
  sp->n_raw = n_raw;
  sp->k = k;
  sp->width = w;
  sp->broadening = params->broadening ? 1 : 0;
  sp->z_qso = z_qso;
  sp->from_raw = true;
  sp->device = rt.device;
  DLA_CUDA(sp->X.alloc(n_raw));
  DLA_CUDA(sp->Y.alloc(n_raw));
  DLA_CUDA(sp->V.alloc(n_raw));
  DLA_CUDA(sp->mask.alloc(n_raw));
  DLA_CUDA(sp->scratch.alloc(n_raw));
  DLA_CUDA(sp->ind_unmasked.alloc(n_raw));
  DLA_CUDA(sp->ind.alloc(n_raw));
  DLA_CUDA(sp->x.alloc(n_raw));
  DLA_CUDA(sp->y.alloc(n_raw));
  DLA_CUDA(sp->v.alloc(n_raw));
  DLA_CUDA(sp->this_wl.alloc(n_raw));
  DLA_CUDA(sp->mu.alloc(n_raw));
  DLA_CUDA(sp->omega2.alloc(n_raw));
  DLA_CUDA(sp->M.alloc((size_t)n_raw * k));
  DLA_CUDA(sp->uidx.alloc(n_raw));
  DLA_CUDA(sp->unmasked_wl.alloc(n_raw));
  DLA_CUDA(sp->wl_abs.alloc(n_raw + 2 * w));
  DLA_CUDA(sp->padded_wl.alloc(n_raw + 2 * w));
  DLA_CUDA(sp->d_scalars.alloc(8));
  DLA_CUDA(sp->X.upload(X, n_raw, rt.stream));
  DLA_CUDA(sp->Y.upload(Y, n_raw, rt.stream));
  DLA_CUDA(sp->V.upload(V, n_raw, rt.stream));
  DLA_CUDA(sp->mask.upload(pixel_mask, n_raw, rt.stream));

  PrepTask t;
  t.X = sp->X.p;
  t.Wobs = nullptr;
  t.Y = sp->Y.p;
  t.V = sp->V.p;
  t.mask = sp->mask.p;
  t.n_raw = n_raw;
  t.z_qso = z_qso;
  t.ind_unmasked = sp->ind_unmasked.p;
  t.ind = sp->ind.p;
  t.x = sp->x.p;
  t.y = sp->y.p;
  t.v = sp->v.p;
  t.this_wl = sp->this_wl.p;
  t.mu = sp->mu.p;
  t.omega2 = sp->omega2.p;
  t.M = sp->M.p;
  t.uidx = sp->uidx.p;
  t.unmasked_wl = sp->unmasked_wl.p;
  t.wl_abs = sp->wl_abs.p;
  t.padded_wl = sp->padded_wl.p;
  t.scratch = sp->scratch.p;
  t.scalars = sp->d_scalars.p;
  DevBuf<PrepTask> d_task;
  DLA_CUDA(d_task.alloc(1));
  DLA_CUDA(cudaMemcpyAsync(d_task.p, &t, sizeof(t), cudaMemcpyHostToDevice, rt.stream));
  KernelTimer timer;
  DLA_CUDA(timer.begin());
  prepare_spectrum_kernel<<<1, 256, 0, rt.stream>>>(d_task.p, model->dev, to_prep_params(params, normalize));
  DLA_LAUNCHED();
  DLA_CUDA(timer.end());
  DLA_CUDA(sp->d_scalars.download(sp->scalars, 8, rt.stream));
  DLA_CUDA(cudaStreamSynchronize(rt.stream));
  sp->n_u = (int)sp->scalars[0];
  sp->n = (int)sp->scalars[1];
  sp->n_abs = sp->broadening ? sp->n_u + 2 * w : sp->n_u;
  sp->ld = (int)round_up(std::max(sp->n, 1), 4);
  *out = sp.release();
  return 0;
}

extern "C" int dla_spectrum_create_prepared(const double* y, const double* v, const double* mu, const double* M,
                                            const double* omega2, int n, int k, const double* wl_abs, int n_abs,
                                            const uint8_t* keep, int n_u, int broadening, dla_spectrum** out) {
  DLA_CHECK_READY();
  Runtime& rt = runtime();
  DLA_REQUIRE(y && v && mu && M && omega2 && wl_abs && keep && out, "null pointer argument");
  DLA_REQUIRE(n >= 1 && k >= 1, "empty spectrum");
  DLA_REQUIRE(n_abs == (broadening ? n_u + 2 * INSTRUMENT_WIDTH : n_u), "absorption grid length does not match n_u");
  int kept = 0;
  for (int i = 0; i < n_u; ++i) kept += keep[i] ? 1 : 0;
  DLA_REQUIRE(kept == n, "keep mask does not select n pixels");  // dla_gp.py:390 assert
  std::unique_ptr<dla_spectrum> sp(new dla_spectrum());
  sp->n = n;
  sp->n_u = n_u;
  sp->k = k;
  sp->width = INSTRUMENT_WIDTH;
  sp->broadening = broadening ? 1 : 0;
  sp->n_abs = n_abs;
  sp->ld = (int)round_up(n, 4);
  sp->device = rt.device;
  DLA_CUDA(sp->y.alloc(n));
  DLA_CUDA(sp->v.alloc(n));
  DLA_CUDA(sp->mu.alloc(n));
  DLA_CUDA(sp->omega2.alloc(n));
  DLA_CUDA(sp->M.alloc((size_t)n * k));
  DLA_CUDA(sp->wl_abs.alloc(n_abs));
  DLA_CUDA(sp->uidx.alloc(n));
  DLA_CUDA(sp->y.upload(y, n, rt.stream));
  DLA_CUDA(sp->v.upload(v, n, rt.stream));
  DLA_CUDA(sp->mu.upload(mu, n, rt.stream));
  DLA_CUDA(sp->omega2.upload(omega2, n, rt.stream));
  DLA_CUDA(sp->M.upload(M, (size_t)n * k, rt.stream));
  DLA_CUDA(sp->wl_abs.upload(wl_abs, n_abs, rt.stream));
  std::vector<int32_t> uidx;
  uidx.reserve(n);
  for (int i = 0; i < n_u; ++i)
    if (keep[i]) uidx.push_back(i);
  DLA_CUDA(sp->uidx.upload(uidx.data(), n, rt.stream));
  DLA_CUDA(cudaStreamSynchronize(rt.stream));
  *out = sp.release();
  return 0;
}

extern "C" int dla_spectrum_destroy(dla_spectrum* spec) {
  delete spec;
  return 0;
}

extern "C" int dla_spectrum_sizes(const dla_spectrum* spec, int* n_raw, int* n_u, int* n) {
  DLA_REQUIRE(spec, "null spectrum");
  if (n_raw) *n_raw = spec->n_raw;
  if (n_u) *n_u = spec->n_u;
  if (n) *n = spec->n;
  return 0;
}

extern "C" int dla_spectrum_get(const dla_spectrum* sp, double* x, double* y, double* v, double* this_wavelengths,
                                double* this_mu, double* this_M, double* this_omega2, double* unmasked_wavelengths,
                                double* padded_wavelengths, uint8_t* ind_unmasked, uint8_t* ind,
                                double* normalization_median) {
  DLA_CHECK_READY();
  Runtime& rt = runtime();
  DLA_REQUIRE(sp, "null spectrum");
  DLA_REQUIRE(sp->device == rt.device, "the spectrum lives on another device than the one selected by dla_init");
  const size_t n = sp->n, nu = sp->n_u;
  if (x && sp->x.p) DLA_CUDA(sp->x.download(x, n, rt.stream));
  if (y) DLA_CUDA(sp->y.download(y, n, rt.stream));
  if (v) DLA_CUDA(sp->v.download(v, n, rt.stream));
  if (this_wavelengths && sp->this_wl.p) DLA_CUDA(sp->this_wl.download(this_wavelengths, n, rt.stream));
  if (this_mu) DLA_CUDA(sp->mu.download(this_mu, n, rt.stream));
  if (this_M) DLA_CUDA(sp->M.download(this_M, n * sp->k, rt.stream));
  if (this_omega2) DLA_CUDA(sp->omega2.download(this_omega2, n, rt.stream));
  if (unmasked_wavelengths && sp->unmasked_wl.p) DLA_CUDA(sp->unmasked_wl.download(unmasked_wavelengths, nu, rt.stream));
  if (padded_wavelengths && sp->padded_wl.p)
    DLA_CUDA(sp->padded_wl.download(padded_wavelengths, nu + 2 * sp->width, rt.stream));
  if (ind_unmasked && sp->ind_unmasked.p) DLA_CUDA(sp->ind_unmasked.download(ind_unmasked, sp->n_raw, rt.stream));
  if (ind && sp->ind.p) DLA_CUDA(sp->ind.download(ind, sp->n_raw, rt.stream));
  DLA_CUDA(cudaStreamSynchronize(rt.stream));
  if (normalization_median) *normalization_median = sp->scalars[2];
  return 0;
}

// ------------------------------------------------------------------------------------------
// likelihood launches on one spectrum
// ------------------------------------------------------------------------------------------
static LikelihoodSpectrum base_desc(dla_spectrum* sp) {
  LikelihoodSpectrum d;
  d.y = sp->y.p;
  d.v = sp->v.p;
  d.mu = sp->mu.p;
  d.omega2 = sp->omega2.p;
  d.M = sp->M.p;
  d.P = sp->basis.p;
  d.base0 = sp->cache.p;
  d.cache = sp->cache.p;
  d.rows0 = nullptr;
  d.rows = nullptr;
  d.prod_out = nullptr;
  d.out = nullptr;
  d.n = sp->n;
  d.ld = sp->ld;
  d.num_samples = 0;
  d.num_rows = 1;
  d.row_stride = 0;
  d.row0 = 0;
  d.alive = nullptr;
  return d;
}

static int launch_likelihood(const LikelihoodSpectrum* d_desc, int max_samples, int num_spectra) {
  Runtime& rt = runtime();
  dim3 grid((max_samples + LK_TS - 1) / LK_TS, num_spectra);
  sample_likelihood_kernel<<<grid, LK_THREADS, LK_SMEM_BYTES, rt.stream>>>(d_desc);
  DLA_LAUNCHED();
  return 0;
}

extern "C" int dla_null_log_model_evidence(dla_spectrum* sp, double* out) {
  DLA_CHECK_READY();
  Runtime& rt = runtime();
  DLA_REQUIRE(sp && out, "null pointer argument");
  DLA_REQUIRE(sp->device == rt.device, "the spectrum lives on another device than the one selected by dla_init");
  DLA_REQUIRE(sp->k == LK_K, "the batched likelihood path is built for k = 20");
  DLA_REQUIRE(sp->n >= 1, "spectrum has no modelled pixels");
  if (int rcb = ensure_basis(sp)) return rcb;
  DevBuf<double> ones, res;
  DLA_CUDA(ones.alloc(sp->ld));
  DLA_CUDA(res.alloc(1));
  KernelTimer timer;
  DLA_CUDA(timer.begin());
  fill_double_kernel<<<(sp->ld + 255) / 256, 256, 0, rt.stream>>>(ones.p, sp->ld, 1.0);
  DLA_LAUNCHED();
  LikelihoodSpectrum d = base_desc(sp);
  d.base0 = ones.p;
  d.cache = ones.p;
  d.out = res.p;
  d.num_samples = 1;
  DLA_CUDA(sp->lk_desc.ensure(16));
  DLA_CUDA(cudaMemcpyAsync(sp->lk_desc.p, &d, sizeof(d), cudaMemcpyHostToDevice, rt.stream));
  int rc = launch_likelihood(sp->lk_desc.p, 1, 1);
  if (rc) return rc;
  DLA_CUDA(timer.end());
  DLA_CUDA(res.download(out, 1, rt.stream));
  DLA_CUDA(cudaStreamSynchronize(rt.stream));
  return 0;
}

extern "C" int dla_sample_log_likelihoods(dla_spectrum* sp, const double* z_dlas, const double* nhis, int S, int k_dlas,
                                          int num_lines, double* out) {
  DLA_CHECK_READY();
  Runtime& rt = runtime();
  DLA_REQUIRE(sp && z_dlas && nhis && out, "null pointer argument");
  DLA_REQUIRE(sp->device == rt.device, "the spectrum lives on another device than the one selected by dla_init");
  DLA_REQUIRE(S >= 1 && k_dlas >= 1 && k_dlas <= LK_MAX_ROWS, "need S >= 1 and 1 <= k_dlas <= 8");
  DLA_REQUIRE(sp->k == LK_K, "the batched likelihood path is built for k = 20");
  DLA_REQUIRE(sp->n >= 1, "spectrum has no modelled pixels");
  const size_t rows = (size_t)S * k_dlas;
  // factor r of sample s lives in profile row r*S + s
  std::vector<double> zt(rows), nt(rows);
  for (int s = 0; s < S; ++s)
    for (int r = 0; r < k_dlas; ++r) {
      zt[(size_t)r * S + s] = z_dlas[(size_t)s * k_dlas + r];
      nt[(size_t)r * S + s] = nhis[(size_t)s * k_dlas + r];
    }
  DLA_CUDA(sp->z_dev.ensure(rows));
  DLA_CUDA(sp->nhi_dev.ensure(rows));
  DLA_CUDA(sp->z_dev.upload(zt.data(), rows, rt.stream));
  DLA_CUDA(sp->nhi_dev.upload(nt.data(), rows, rt.stream));
  int rc = ensure_cache(sp, rows);
  if (rc) return rc;
  if ((rc = ensure_basis(sp))) return rc;
  DLA_CUDA(sp->raw_ll.ensure(S));
  KernelTimer timer;
  DLA_CUDA(timer.begin());
  rc = launch_voigt(sp, sp->z_dev.p, sp->nhi_dev.p, (int)rows, num_lines, sp->cache.p, sp->ld);
  if (rc) return rc;
  LikelihoodSpectrum d = base_desc(sp);
  d.out = sp->raw_ll.p;
  d.num_samples = S;
  d.num_rows = k_dlas;
  d.row_stride = S;
  DLA_CUDA(sp->lk_desc.ensure(16));
  DLA_CUDA(cudaMemcpyAsync(sp->lk_desc.p, &d, sizeof(d), cudaMemcpyHostToDevice, rt.stream));
  rc = launch_likelihood(sp->lk_desc.p, S, 1);
  if (rc) return rc;
  DLA_CUDA(timer.end());
  DLA_CUDA(sp->raw_ll.download(out, S, rt.stream));
  DLA_CUDA(cudaStreamSynchronize(rt.stream));
  return 0;
}

extern "C" int dla_absorption_k_dlas(dla_spectrum* sp, const double* z_dlas, const double* nhis, int k_dlas,
                                     int num_lines, double* out) {
  DLA_CHECK_READY();
  Runtime& rt = runtime();
  DLA_REQUIRE(sp && z_dlas && nhis && out, "null pointer argument");
  DLA_REQUIRE(sp->device == rt.device, "the spectrum lives on another device than the one selected by dla_init");
  DLA_REQUIRE(k_dlas >= 1, "need at least one absorber");
  DLA_REQUIRE(sp->n >= 1, "spectrum has no modelled pixels");
  DLA_CUDA(sp->z_dev.ensure(k_dlas));
  DLA_CUDA(sp->nhi_dev.ensure(k_dlas));
  DLA_CUDA(sp->z_dev.upload(z_dlas, k_dlas, rt.stream));
  DLA_CUDA(sp->nhi_dev.upload(nhis, k_dlas, rt.stream));
  int rc = ensure_cache(sp, (size_t)k_dlas + 1);
  if (rc) return rc;
  rc = launch_voigt(sp, sp->z_dev.p, sp->nhi_dev.p, k_dlas, num_lines, sp->cache.p, sp->ld);
  if (rc) return rc;
  double* d_out = sp->cache.p + (size_t)k_dlas * sp->ld;
  product_rows_kernel<<<(sp->n + 255) / 256, 256, 0, rt.stream>>>(sp->cache.p, sp->ld, sp->n, k_dlas, d_out);
  DLA_LAUNCHED();
  DLA_CUDA(cudaMemcpyAsync(out, d_out, sizeof(double) * sp->n, cudaMemcpyDeviceToHost, rt.stream));
  DLA_CUDA(cudaStreamSynchronize(rt.stream));
  return 0;
}

// ------------------------------------------------------------------------------------------
// a9: evidence levels on one spectrum
// ------------------------------------------------------------------------------------------
extern "C" int dla_log_model_evidences(dla_spectrum* sp, const double* z_samples, const double* nhi_samples, int S,
                                       int max_dlas, const double* uniforms, double min_z_separation, int num_lines,
                                       double* sample_log_likelihoods, int32_t* base_sample_inds, double* log_evidences,
                                       int* uniform_rows_used) {
  DLA_CHECK_READY();
  Runtime& rt = runtime();
  DLA_REQUIRE(sp && z_samples && nhi_samples && log_evidences, "null pointer argument");
  DLA_REQUIRE(sp->device == rt.device, "the spectrum lives on another device than the one selected by dla_init");
  DLA_REQUIRE(S >= 1 && max_dlas >= 1 && max_dlas <= LK_MAX_ROWS, "need S >= 1 and 1 <= max_dlas <= 8");
  DLA_REQUIRE(max_dlas == 1 || uniforms, "uniforms are required when max_dlas > 1");
  DLA_REQUIRE(sp->k == LK_K, "the batched likelihood path is built for k = 20");
  DLA_REQUIRE(sp->n >= 1, "spectrum has no modelled pixels");

  DLA_CUDA(sp->z_dev.ensure(S));
  DLA_CUDA(sp->nhi_dev.ensure(S));
  DLA_CUDA(sp->z_dev.upload(z_samples, S, rt.stream));
  DLA_CUDA(sp->nhi_dev.upload(nhi_samples, S, rt.stream));
  if (max_dlas > 1) {
    DLA_CUDA(sp->uniforms.ensure((size_t)(max_dlas - 1) * S));
    DLA_CUDA(sp->uniforms.upload(uniforms, (size_t)(max_dlas - 1) * S, rt.stream));
  }
  int rc = ensure_cache(sp, S);
  if (rc) return rc;
  if ((rc = ensure_basis(sp))) return rc;
  const size_t prod_half = (size_t)S * sp->ld;
  if (max_dlas >= 3) DLA_CUDA(sp->prod.ensure(prod_half * (max_dlas >= 4 ? 2 : 1)));  // ping-pong, see catalogue.inc.cuh
  DLA_CUDA(sp->raw_ll.ensure(S));
  DLA_CUDA(sp->sample_ll.ensure((size_t)S * max_dlas));
  DLA_CUDA(sp->log_ev.ensure(max_dlas));
  DLA_CUDA(sp->cdf.ensure(S));
  DLA_CUDA(sp->rows.ensure((size_t)S * max_dlas));
  DLA_CUDA(sp->alive.ensure(4 + LK_MAX_ROWS));
  DLA_CUDA(sp->lk_desc.ensure(16));
  DLA_CUDA(sp->ev_desc.ensure(16));
  DLA_CUDA(sp->sel.ensure(S));
  DLA_CUDA(sp->pos_a.ensure(S));
  DLA_CUDA(sp->pos_b.ensure(S));
  DLA_CUDA(sp->rows0_c.ensure(S));
  DLA_CUDA(sp->rows1_c.ensure(S));
  DLA_CUDA(sp->raw_slots.ensure(S));
  DLA_CUDA(sp->compact_desc.ensure(16));
  DLA_CUDA(sp->scatter_desc.ensure(1));

  // descriptors of all levels (they do not depend on results)
  std::vector<LikelihoodSpectrum> lk(max_dlas);
  std::vector<EvidenceLevel> ev(max_dlas);
  std::vector<CompactTask> compact(max_dlas);
  for (int level = 0; level < max_dlas; ++level) {
    LikelihoodSpectrum d = base_desc(sp);
    d.out = sp->raw_ll.p;
    d.num_samples = S;
    // level L >= 1 multiplies the running product of level L-1 by the profile of the newly drawn absorber
    // base_sample_inds[L-1][s], for the samples that pass the separation test only: compacted launch in slot order,
    // products ping-pong between two buffers (same scheme as the catalogue engine, catalogue.inc.cuh)
    d.alive = sp->alive.p;
    if (level == 0) {
      d.num_rows = 1;
      d.base0 = sp->cache.p;
    } else {
      double* prod_write = sp->prod.p + ((max_dlas >= 4 && (level & 1) == 0) ? prod_half : 0);
      const double* prod_read = sp->prod.p + ((max_dlas >= 4 && ((level - 1) & 1) == 0) ? prod_half : 0);
      d.num_rows = 2;
      d.base0 = level == 1 ? sp->cache.p : prod_read;
      d.rows0 = sp->rows0_c.p;
      d.rows = sp->rows1_c.p;
      d.row_stride = S;
      d.prod_out = (level + 1 < max_dlas) ? prod_write : nullptr;
      d.out = sp->raw_slots.p;
      d.num_samples = -(4 + level);  // the count is alive[4 + level], written by compact_level_kernel
    }
    CompactTask ct;
    ct.z_samples = sp->z_dev.p;
    ct.base_inds = sp->rows.p + S;
    ct.alive = sp->alive.p;
    ct.pos_prev = level <= 1 ? nullptr : ((level & 1) ? sp->pos_b.p : sp->pos_a.p);
    ct.pos = (level & 1) ? sp->pos_a.p : sp->pos_b.p;
    ct.sel = sp->sel.p;
    ct.rows0 = sp->rows0_c.p;
    ct.rows1 = sp->rows1_c.p;
    ct.raw_ll = sp->raw_ll.p;
    ct.num_sel = sp->alive.p + 4 + level;
    ct.S = S;
    ct.level = level;
    ct.min_z_separation = min_z_separation;
    compact[level] = ct;
    lk[level] = d;
    EvidenceLevel e;
    e.raw_ll = sp->raw_ll.p;
    e.sample_ll = sp->sample_ll.p + level;
    e.ll_stride = max_dlas;
    e.z_samples = sp->z_dev.p;
    e.base_inds = sp->rows.p + S;
    e.base_out = (level + 1 < max_dlas) ? sp->rows.p + (size_t)(level + 1) * S : nullptr;
    e.uniforms = (level + 1 < max_dlas) ? sp->uniforms.p + (size_t)level * S : nullptr;
    e.log_evidence = sp->log_ev.p + level;
    e.cdf_scratch = sp->cdf.p;
    e.alive = sp->alive.p;
    e.status = sp->alive.p + 1;
    e.S = S;
    e.level = level;
    e.min_z_separation = min_z_separation;
    ev[level] = e;
  }
  DLA_CUDA(cudaMemcpyAsync(sp->lk_desc.p, lk.data(), sizeof(LikelihoodSpectrum) * max_dlas, cudaMemcpyHostToDevice, rt.stream));
  DLA_CUDA(cudaMemcpyAsync(sp->ev_desc.p, ev.data(), sizeof(EvidenceLevel) * max_dlas, cudaMemcpyHostToDevice, rt.stream));
  DLA_CUDA(cudaMemcpyAsync(sp->compact_desc.p, compact.data(), sizeof(CompactTask) * max_dlas, cudaMemcpyHostToDevice, rt.stream));
  ScatterTask sc;
  sc.raw_slots = sp->raw_slots.p;
  sc.sel = sp->sel.p;
  sc.num_sel = sp->alive.p + 4;
  sc.raw_ll = sp->raw_ll.p;
  DLA_CUDA(cudaMemcpyAsync(sp->scatter_desc.p, &sc, sizeof(sc), cudaMemcpyHostToDevice, rt.stream));
  int alive_init[4 + LK_MAX_ROWS] = {1, 0};
  DLA_CUDA(cudaMemcpyAsync(sp->alive.p, alive_init, sizeof(alive_init), cudaMemcpyHostToDevice, rt.stream));

  KernelTimer timer;
  DLA_CUDA(timer.begin());
  const double nan = std::numeric_limits<double>::quiet_NaN();
  fill_double_kernel<<<(unsigned)(((size_t)S * max_dlas + 255) / 256), 256, 0, rt.stream>>>(sp->sample_ll.p, (size_t)S * max_dlas, nan);
  DLA_LAUNCHED();
  fill_double_kernel<<<1, 256, 0, rt.stream>>>(sp->log_ev.p, max_dlas, nan);
  DLA_LAUNCHED();
  fill_rows_kernel<<<(unsigned)(((size_t)S * max_dlas + 255) / 256), 256, 0, rt.stream>>>(sp->rows.p, S, max_dlas);
  DLA_LAUNCHED();
  rc = launch_voigt(sp, sp->z_dev.p, sp->nhi_dev.p, S, num_lines, sp->cache.p, sp->ld);
  if (rc) return rc;
  for (int level = 0; level < max_dlas; ++level) {
    if (level > 0) {
      compact_level_kernel<<<1, 1024, 0, rt.stream>>>(sp->compact_desc.p + level);
      DLA_LAUNCHED();
    }
    rc = launch_likelihood(sp->lk_desc.p + level, S, 1);
    if (rc) return rc;
    if (level > 0) {
      scatter_ll_kernel<<<dim3((S + 255) / 256, 1), 256, 0, rt.stream>>>(sp->scatter_desc.p, level);
      DLA_LAUNCHED();
    }
    evidence_level_kernel<<<1, 1024, 0, rt.stream>>>(sp->ev_desc.p + level);
    DLA_LAUNCHED();
  }
  DLA_CUDA(timer.end());
  DLA_CUDA(sp->log_ev.download(log_evidences, max_dlas, rt.stream));
  if (sample_log_likelihoods) DLA_CUDA(sp->sample_ll.download(sample_log_likelihoods, (size_t)S * max_dlas, rt.stream));
  if (base_sample_inds && max_dlas > 1)
    DLA_CUDA(cudaMemcpyAsync(base_sample_inds, sp->rows.p + S, sizeof(int32_t) * (size_t)(max_dlas - 1) * S,
                             cudaMemcpyDeviceToHost, rt.stream));
  DLA_CUDA(cudaStreamSynchronize(rt.stream));
  if (uniform_rows_used) {
    int used = 0;
    for (int level = 0; level + 1 < max_dlas; ++level) {
      if (isnan(log_evidences[level])) break;
      ++used;
    }
    *uniform_rows_used = used;
  }
  return 0;
}

extern "C" int dla_resample_indices(const double* W, const double* uniforms, int S, int32_t* out) {
  DLA_CHECK_READY();
  Runtime& rt = runtime();
  DLA_REQUIRE(W && uniforms && out && S >= 1, "bad argument");
  // scratch kept between calls (the class API calls this once per level per spectrum; four cudaMalloc / cudaFree pairs
  // per call each synchronised the device); re-created when dla_init moves the library to another device
  static DevBuf<double> dW, dU, scratch;
  static DevBuf<int32_t> dout;
  static int scratch_device = -1;
  if (scratch_device != rt.device) {
    dW.release(); dU.release(); scratch.release(); dout.release();
    scratch_device = rt.device;
  }
  DLA_CUDA(dW.ensure(S));
  DLA_CUDA(dU.ensure(S));
  DLA_CUDA(scratch.ensure(S));
  DLA_CUDA(dout.ensure(S));
  DLA_CUDA(dW.upload(W, S, rt.stream));
  DLA_CUDA(dU.upload(uniforms, S, rt.stream));
  resample_kernel<<<1, 1024, 0, rt.stream>>>(dW.p, dU.p, S, scratch.p, dout.p);
  DLA_LAUNCHED();
  DLA_CUDA(dout.download(out, S, rt.stream));
  DLA_CUDA(cudaStreamSynchronize(rt.stream));
  return 0;
}

#include "catalogue.inc.cuh"
#include "zqso.inc.cuh"
