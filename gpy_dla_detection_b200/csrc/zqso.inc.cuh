// zqso.inc.cuh : C-ABI of the quasar-redshift estimation path - included by dla_b200.cu.
// Reference: ZGP (gpy_dla_detection/zqso_gp.py).

struct dla_zqso_model {
  DevBuf<double> rest, mu, mu_slope, M, M_slope, MS;
  ZqsoModelDev dev;
  int device = -1;
  // workspace of dla_zqso_inference, kept between calls (a cudaMalloc / cudaFree pair per call synchronises the device)
  struct Workspace {
    DevBuf<double> X, Y, V, z, ll, zmap, med;
    DevBuf<uint8_t> mask;
    DevBuf<int32_t> mapi;
    DevBuf<ZqsoSpectrum> desc;
  } ws;
};

// timing of the last dla_zqso_inference call: kernels only (inputs resident) and the whole call
static double g_zqso_kernel_ms = 0.0, g_zqso_total_ms = 0.0;
static int g_zqso_force_generic = 0;  // tests: run the generic (non-uniform grid) kernel on a uniform grid too

extern "C" int dla_zqso_model_create(const double* rest_wavelengths, const double* mu, const double* M, int n_rest, int k,
                                     double bluewards_mu, double redwards_mu, double bluewards_sigma,
                                     double redwards_sigma, dla_zqso_model** out) {
  DLA_CHECK_READY();
  Runtime& rt = runtime();
  DLA_REQUIRE(rest_wavelengths && mu && M && out, "null pointer argument");
  DLA_REQUIRE(n_rest >= 2, "the model grid needs at least two points");
  DLA_REQUIRE(k == ZQ_K, "the zQSO path is built for k = 20");
  for (int i = 1; i < n_rest; ++i) DLA_REQUIRE(rest_wavelengths[i] > rest_wavelengths[i - 1], "rest_wavelengths must increase");
  std::unique_ptr<dla_zqso_model> m(new dla_zqso_model());
  DevBuf<double> M_in;
  DLA_CUDA(m->rest.alloc(n_rest));
  DLA_CUDA(m->mu.alloc(n_rest));
  DLA_CUDA(m->mu_slope.alloc(n_rest));
  DLA_CUDA(m->M.alloc((size_t)n_rest * ZQ_STRIDE));
  DLA_CUDA(m->M_slope.alloc((size_t)n_rest * ZQ_STRIDE));
  DLA_CUDA(M_in.alloc((size_t)n_rest * k));
  DLA_CUDA(m->rest.upload(rest_wavelengths, n_rest, rt.stream));
  DLA_CUDA(m->mu.upload(mu, n_rest, rt.stream));
  DLA_CUDA(M_in.upload(M, (size_t)n_rest * k, rt.stream));
  zqso_slopes_kernel<<<(n_rest + 127) / 128, 128, 0, rt.stream>>>(m->rest.p, m->mu.p, M_in.p, n_rest, k, m->mu_slope.p,
                                                                   m->M.p, m->M_slope.p);
  DLA_LAUNCHED();
  DLA_CUDA(m->MS.alloc((size_t)n_rest * ZQ_STRIDE * 2));
  zqso_interleave_kernel<<<(n_rest + 127) / 128, 128, 0, rt.stream>>>(m->mu.p, m->mu_slope.p, m->M.p, m->M_slope.p, n_rest,
                                                                       m->MS.p);
  DLA_LAUNCHED();
  DLA_CUDA(cudaStreamSynchronize(rt.stream));
  // uniform grid (the published models: 910:0.25:3000): constant-time interval lookup
  const double dl = rest_wavelengths[1] - rest_wavelengths[0];
  bool uniform = true;
  for (int i = 0; i < n_rest && uniform; ++i) uniform = rest_wavelengths[i] == rest_wavelengths[0] + i * dl;
  ZqsoModelDev& d = m->dev;
  d.rest = m->rest.p;
  d.mu = m->mu.p;
  d.mu_slope = m->mu_slope.p;
  d.M = m->M.p;
  d.M_slope = m->M_slope.p;
  d.MS = m->MS.p;
  d.n_rest = n_rest;
  d.uniform = uniform ? 1 : 0;
  d.rest0 = rest_wavelengths[0];
  d.inv_dl = 1.0 / dl;
  d.bluewards_mu = bluewards_mu;
  d.redwards_mu = redwards_mu;
  d.bluewards_var = bluewards_sigma * bluewards_sigma;  // sigma ** 2 (zqso_gp.py:202,208)
  d.redwards_var = redwards_sigma * redwards_sigma;
  m->device = rt.device;
  *out = m.release();
  return 0;
}

extern "C" int dla_zqso_model_destroy(dla_zqso_model* model) {
  delete model;
  return 0;
}

// power of two >= the largest number of pixels whose rest wavelength can fall in [nmin, nmax] at any z
static int zqso_norm_cap(const double* X, int n, double ratio) {
  int best = 1, j = 0;
  for (int i = 0; i < n; ++i) {
    while (j < n && X[j] <= X[i] * ratio * (1.0 + 1e-12)) ++j;
    best = std::max(best, j - i);
  }
  int cap = 32;
  while (cap < best + 2) cap <<= 1;
  return cap;
}

static ZqsoParamsDev to_zqso_params(const dla_zqso_params* p) {
  ZqsoParamsDev d;
  d.min_lambda = p->min_lambda;
  d.max_lambda = p->max_lambda;
  d.norm_min_lambda = p->normalization_min_lambda;
  d.norm_max_lambda = p->normalization_max_lambda;
  return d;
}

extern "C" int dla_zqso_inference(const dla_zqso_model* model, const dla_zqso_params* params, int num_spectra,
                                  const int64_t* pixel_offsets, const double* wavelengths, const double* flux,
                                  const double* noise_variance, const uint8_t* pixel_mask, const double* z_samples,
                                  int S, double* sample_log_likelihoods, double* z_map, int32_t* map_index) {
  DLA_CHECK_READY();
  Runtime& rt = runtime();
  DLA_REQUIRE(model && params && pixel_offsets && wavelengths && flux && noise_variance && pixel_mask && z_samples,
              "null pointer argument");
  DLA_REQUIRE(num_spectra >= 1 && S >= 1, "need at least one spectrum and one redshift sample");
  DLA_REQUIRE(model->device == rt.device, "the model lives on another device than the one selected by dla_init");
  DLA_REQUIRE(params->normalization_min_lambda > 0 && params->normalization_max_lambda >= params->normalization_min_lambda,
              "bad normalisation window");
  const int64_t total = pixel_offsets[num_spectra];
  DLA_REQUIRE(pixel_offsets[0] == 0 && total >= 1, "pixel_offsets must start at 0");
  int norm_cap = 32;
  const double ratio = params->normalization_max_lambda / params->normalization_min_lambda;
  for (int q = 0; q < num_spectra; ++q) {
    const int64_t a = pixel_offsets[q], b = pixel_offsets[q + 1];
    DLA_REQUIRE(b - a >= 2 && b - a < (1 << 30), "every spectrum needs at least two pixels");
    for (int64_t i = a + 1; i < b; ++i)
      DLA_REQUIRE(wavelengths[i] > wavelengths[i - 1], "observed wavelengths must be strictly increasing");
    norm_cap = std::max(norm_cap, zqso_norm_cap(wavelengths + a, (int)(b - a), ratio));
  }
  DLA_REQUIRE(norm_cap <= 2048, "normalisation window holds more than 2048 pixels");
  const int per_warp = std::max(norm_cap, ZQ_TDIM * ZQ_TSTRIDE);
  const size_t smem = (size_t)ZQ_WARPS * per_warp * sizeof(double);
  DLA_CUDA(cudaFuncSetAttribute(zqso_likelihood_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  // uniform model grid (the published one): median pass + the v2 kernel; any other grid: the generic kernel
  const bool v2 = model->dev.uniform != 0 && g_zqso_force_generic == 0;
  const size_t smem_v2 = (size_t)ZQ2_WARPS * ZQ2_PER_WARP * sizeof(double);
  const size_t smem_med = (size_t)norm_cap * (sizeof(double) + sizeof(int));
  DevBuf<double>& dmed = const_cast<dla_zqso_model*>(model)->ws.med;

  dla_zqso_model::Workspace& ws = const_cast<dla_zqso_model*>(model)->ws;
  DevBuf<double>&dX = ws.X, &dY = ws.Y, &dV = ws.V, &dz = ws.z, &dll = ws.ll, &dzmap = ws.zmap;
  DevBuf<uint8_t>& dmask = ws.mask;
  DevBuf<int32_t>& dmapi = ws.mapi;
  DevBuf<ZqsoSpectrum>& ddesc = ws.desc;
  DLA_CUDA(dX.ensure(total));
  DLA_CUDA(dY.ensure(total));
  DLA_CUDA(dV.ensure(total));
  DLA_CUDA(dmask.ensure(total));
  DLA_CUDA(dz.ensure(S));
  DLA_CUDA(dX.upload(wavelengths, total, rt.stream));
  DLA_CUDA(dY.upload(flux, total, rt.stream));
  DLA_CUDA(dV.upload(noise_variance, total, rt.stream));
  DLA_CUDA(dmask.upload(pixel_mask, total, rt.stream));
  DLA_CUDA(dz.upload(z_samples, S, rt.stream));
  const int chunk = std::max(1, std::min(num_spectra, (int)std::min<size_t>(16384, ((size_t)1 << 28) / (size_t)S)));  // <= 2 GiB of ll
  DLA_CUDA(dll.ensure((size_t)chunk * S));
  DLA_CUDA(dzmap.ensure(chunk));
  DLA_CUDA(dmapi.ensure(chunk));
  DLA_CUDA(ddesc.ensure(chunk));
  std::vector<ZqsoSpectrum> h_desc(chunk);
  cudaEvent_t e_k0, e_k1;
  DLA_CUDA(cudaEventCreate(&e_k0));
  DLA_CUDA(cudaEventCreate(&e_k1));
  g_zqso_kernel_ms = 0.0;
  KernelTimer timer;
  DLA_CUDA(timer.begin());
  for (int q0 = 0; q0 < num_spectra; q0 += chunk) {
    const int nb = std::min(chunk, num_spectra - q0);
    for (int b = 0; b < nb; ++b) {
      const int64_t off = pixel_offsets[q0 + b];
      h_desc[b].X = dX.p + off;
      h_desc[b].Y = dY.p + off;
      h_desc[b].V = dV.p + off;
      h_desc[b].mask = dmask.p + off;
      h_desc[b].n_raw = (int)(pixel_offsets[q0 + b + 1] - off);
    }
    DLA_CUDA(cudaMemcpyAsync(ddesc.p, h_desc.data(), sizeof(ZqsoSpectrum) * nb, cudaMemcpyHostToDevice, rt.stream));
    dim3 grid((S + ZQ_WARPS - 1) / ZQ_WARPS, nb);
    DLA_CUDA(cudaEventRecord(e_k0, rt.stream));
    if (v2) {
      DLA_CUDA(dmed.ensure((size_t)chunk * S));
      dim3 mgrid((S + ZQ2_ZPB - 1) / ZQ2_ZPB, nb);
      zqso_median_batch_kernel<<<mgrid, 256, smem_med, rt.stream>>>(ddesc.p, dz.p, S, to_zqso_params(params), norm_cap, dmed.p);
      DLA_LAUNCHED();
      dim3 grid2((S + ZQ2_WARPS - 1) / ZQ2_WARPS, nb);
      zqso_likelihood_kernel_v2<<<grid2, ZQ2_WARPS * 32, smem_v2, rt.stream>>>(ddesc.p, dz.p, S, dmed.p, model->dev,
                                                                            to_zqso_params(params), dll.p);
      DLA_LAUNCHED();
    } else {
      zqso_likelihood_kernel<<<grid, ZQ_WARPS * 32, smem, rt.stream>>>(ddesc.p, dz.p, S, model->dev, to_zqso_params(params),
                                                                       norm_cap, per_warp, dll.p);
      DLA_LAUNCHED();
    }
    zqso_argmax_kernel<<<nb, 256, 0, rt.stream>>>(dll.p, S, dz.p, dzmap.p, dmapi.p);
    DLA_LAUNCHED();
    DLA_CUDA(cudaEventRecord(e_k1, rt.stream));
    if (sample_log_likelihoods) DLA_CUDA(dll.download(sample_log_likelihoods + (size_t)q0 * S, (size_t)nb * S, rt.stream));
    if (z_map) DLA_CUDA(dzmap.download(z_map + q0, nb, rt.stream));
    if (map_index) DLA_CUDA(dmapi.download(map_index + q0, nb, rt.stream));
    DLA_CUDA(cudaStreamSynchronize(rt.stream));  // h_desc is reused by the next chunk
    float ms = 0.f;
    DLA_CUDA(cudaEventElapsedTime(&ms, e_k0, e_k1));
    g_zqso_kernel_ms += ms;
  }
  DLA_CUDA(timer.end());
  g_zqso_total_ms = rt.last_kernel_ms;
  cudaEventDestroy(e_k0);
  cudaEventDestroy(e_k1);
  return 0;
}

extern "C" int dla_zqso_force_generic_kernel(int on) {
  g_zqso_force_generic = on ? 1 : 0;
  return 0;
}

extern "C" int dla_zqso_last_timing(double* kernel_ms, double* total_ms) {
  if (kernel_ms) *kernel_ms = g_zqso_kernel_ms;
  if (total_ms) *total_ms = g_zqso_total_ms;
  return 0;
}

extern "C" int dla_zqso_set_data(const dla_zqso_model* model, const dla_zqso_params* params, const double* X,
                                 const double* Y, const double* noise_variance, const uint8_t* pixel_mask, int n_raw,
                                 double z_qso, double* x, double* y_normalized, double* v_normalized, double* this_mu,
                                 double* this_M, uint8_t* cls, uint8_t* in_window, double* this_median) {
  DLA_CHECK_READY();
  Runtime& rt = runtime();
  DLA_REQUIRE(model && params && X && Y && noise_variance && pixel_mask && x && y_normalized && v_normalized && this_mu &&
                  this_M && cls && in_window && this_median,
              "null pointer argument");
  DLA_REQUIRE(n_raw >= 2, "a spectrum needs at least two pixels");
  for (int i = 1; i < n_raw; ++i) DLA_REQUIRE(X[i] > X[i - 1], "observed wavelengths must be strictly increasing");
  const int norm_cap = zqso_norm_cap(X, n_raw, params->normalization_max_lambda / params->normalization_min_lambda);
  DLA_REQUIRE(norm_cap <= 2048, "normalisation window holds more than 2048 pixels");
  DevBuf<double> dX, dY, dV, dx, dyn, dvn, dmu, dM, dmed;
  DevBuf<uint8_t> dmask, dcls, dinw;
  DLA_CUDA(dX.alloc(n_raw));
  DLA_CUDA(dY.alloc(n_raw));
  DLA_CUDA(dV.alloc(n_raw));
  DLA_CUDA(dmask.alloc(n_raw));
  DLA_CUDA(dx.alloc(n_raw));
  DLA_CUDA(dyn.alloc(n_raw));
  DLA_CUDA(dvn.alloc(n_raw));
  DLA_CUDA(dmu.alloc(n_raw));
  DLA_CUDA(dM.alloc((size_t)n_raw * ZQ_K));
  DLA_CUDA(dcls.alloc(n_raw));
  DLA_CUDA(dinw.alloc(n_raw));
  DLA_CUDA(dmed.alloc(1));
  DLA_CUDA(dX.upload(X, n_raw, rt.stream));
  DLA_CUDA(dY.upload(Y, n_raw, rt.stream));
  DLA_CUDA(dV.upload(noise_variance, n_raw, rt.stream));
  DLA_CUDA(dmask.upload(pixel_mask, n_raw, rt.stream));
  DLA_CUDA(cudaMemsetAsync(dmu.p, 0, sizeof(double) * n_raw, rt.stream));
  DLA_CUDA(cudaMemsetAsync(dM.p, 0, sizeof(double) * (size_t)n_raw * ZQ_K, rt.stream));
  ZqsoSpectrum sp;
  sp.X = dX.p;
  sp.Y = dY.p;
  sp.V = dV.p;
  sp.mask = dmask.p;
  sp.n_raw = n_raw;
  const ZqsoParamsDev prm = to_zqso_params(params);
  KernelTimer timer;
  DLA_CUDA(timer.begin());
  zqso_median_kernel<<<1, 32, (size_t)norm_cap * sizeof(double), rt.stream>>>(sp, z_qso, prm, norm_cap, dmed.p);
  DLA_LAUNCHED();
  DLA_CUDA(dmed.download(this_median, 1, rt.stream));
  DLA_CUDA(cudaStreamSynchronize(rt.stream));
  zqso_set_data_kernel<<<(n_raw + 127) / 128, 128, 0, rt.stream>>>(sp, z_qso, model->dev, prm, *this_median, dx.p, dyn.p,
                                                                    dvn.p, dmu.p, dM.p, dcls.p, dinw.p);
  DLA_LAUNCHED();
  DLA_CUDA(timer.end());
  DLA_CUDA(dx.download(x, n_raw, rt.stream));
  DLA_CUDA(dyn.download(y_normalized, n_raw, rt.stream));
  DLA_CUDA(dvn.download(v_normalized, n_raw, rt.stream));
  DLA_CUDA(dmu.download(this_mu, n_raw, rt.stream));
  DLA_CUDA(dM.download(this_M, (size_t)n_raw * ZQ_K, rt.stream));
  DLA_CUDA(dcls.download(cls, n_raw, rt.stream));
  DLA_CUDA(dinw.download(in_window, n_raw, rt.stream));
  DLA_CUDA(cudaStreamSynchronize(rt.stream));
  return 0;
}

extern "C" int dla_log_mvnpdf_iid(const double* y, const double* mu, const double* d, int n, double* out) {
  DLA_CHECK_READY();
  Runtime& rt = runtime();
  DLA_REQUIRE(y && mu && d && out && n >= 0, "bad argument");
  DevBuf<double> dy, dmu, dd, dout;
  DLA_CUDA(dy.alloc(n));
  DLA_CUDA(dmu.alloc(n));
  DLA_CUDA(dd.alloc(n));
  DLA_CUDA(dout.alloc(1));
  if (n > 0) {
    DLA_CUDA(dy.upload(y, n, rt.stream));
    DLA_CUDA(dmu.upload(mu, n, rt.stream));
    DLA_CUDA(dd.upload(d, n, rt.stream));
  }
  KernelTimer timer;
  DLA_CUDA(timer.begin());
  log_mvnpdf_iid_kernel<<<1, 256, 0, rt.stream>>>(dy.p, dmu.p, dd.p, n, dout.p);
  DLA_LAUNCHED();
  DLA_CUDA(timer.end());
  DLA_CUDA(dout.download(out, 1, rt.stream));
  DLA_CUDA(cudaStreamSynchronize(rt.stream));
  return 0;
}
