// catalogue.inc.cuh : batched catalogue engine - included by dla_b200.cu.
//
// Reference: the per-spectrum loop of run_bayes_select.process_qso (run_bayes_select.py:141-230):
// seed, read, three set_data calls, BayesModelSelect.model_selection, MAP, copy into the
// (num_quasars, ...) result arrays.  Spectra are independent, so a batch of B spectra is
// resident on the device at once and every stage is ONE launch over the batch:
//   prepare (B CTAs) -> z samples -> profiles of the 2S unique absorbers of every spectrum ->
//   level 0 likelihoods of DLA + subDLA + null samples -> per-level {evidence, resample,
//   likelihood} -> MAP -> model posteriors.
// The level dependency (level k+1 needs level k's normalised weights) is a kernel boundary.

namespace dla {

__global__ void z_samples_kernel(const double* __restrict__ scalars, int scalars_stride,
                                 const double* __restrict__ dla_offsets, const double* __restrict__ sub_offsets, int S,
                                 double* __restrict__ z_out /* [B][2S] */) {
  // z_i = min_z + (max_z - min_z) * offset_i  (dla_samples.py:94-104, subdla_samples.py:115-125)
  const int b = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 2 * S) return;
  const double lo = scalars[(size_t)b * scalars_stride + 3], hi = scalars[(size_t)b * scalars_stride + 4];
  const double off = i < S ? dla_offsets[i] : sub_offsets[i - S];
  z_out[(size_t)b * 2 * S + i] = __dadd_rn(lo, __dmul_rn(__dsub_rn(hi, lo), off));
}

struct GatherTask {
  const double* raw_ll0;   // [2S+1] level-0 raw log-likelihoods (DLA, subDLA, null)
  const double* log_ev_dla;  // [max_dlas]
  const double* log_ev_sub;  // [1]
  double* log_lik;         // [2+max_dlas] = [null, sub, dla...]
  int S, max_dlas;
};
__global__ void gather_evidences_kernel(const GatherTask* __restrict__ tasks, int num) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= num) return;
  const GatherTask t = tasks[q];
  t.log_lik[0] = t.raw_ll0[2 * t.S];
  t.log_lik[1] = t.log_ev_sub[0];
  for (int i = 0; i < t.max_dlas; ++i) t.log_lik[2 + i] = t.log_ev_dla[i];
}

__global__ void transpose_inds_kernel(const int32_t* __restrict__ rows /* [B][(max)][S], row 0 = identity */, int S,
                                      int max_dlas, int32_t* __restrict__ out /* [B][S][max-1] */) {
  // run_bayes_select.py:214 stores base_sample_inds transposed: (num_dla_samples, max_dlas - 1)
  const int b = blockIdx.y;
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= S) return;
  rows += (size_t)b * S * max_dlas;
  out += (size_t)b * S * (max_dlas - 1);
  for (int r = 1; r < max_dlas; ++r) out[(size_t)s * (max_dlas - 1) + (r - 1)] = rows[(size_t)r * S + s];
}

// identity row + zeroed resample rows of every spectrum of the batch; ones row (null model) of its profile cache
__global__ void init_rows_kernel(int32_t* __restrict__ rows, int S, int nrows, const AbsorptionGrid* __restrict__ grids) {
  const int b = blockIdx.y;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < (size_t)S * nrows) rows[(size_t)b * S * nrows + i] = i < (size_t)S ? (int32_t)i : 0;
  const AbsorptionGrid g = grids[b];
  if (i < (size_t)g.ld) g.out[(size_t)2 * S * g.ld + i] = 1.0;
}

}  // namespace dla

struct dla_catalogue {
  const dla_model* model = nullptr;
  dla_params params;
  int S = 0, max_dlas = 0, B = 0, keep = 0;
  bool paired_offsets = false;  // DLA and subDLA samples share their redshift offsets: one line-sum evaluation for both
  // constants
  DevBuf<double> dla_offsets, dla_log_nhi, sub_offsets, nhi_all /* [dla_nhi ; sub_nhi] */, uniforms;
  // staged inputs
  int Q = 0;
  std::vector<int64_t> pix_off;
  std::vector<double> z_qsos;
  DevBuf<double> wl, flux, var, log_priors_in;
  DevBuf<uint8_t> mask;
  int max_n_raw = 0;
  // per-batch workspace
  DevBuf<uint8_t> ind_unmasked, ind;
  DevBuf<double> x, y, v, this_wl, mu, omega2, M, unmasked_wl, wl_abs, padded_wl, scratch, scalars;
  DevBuf<int32_t> uidx, qmap;
  DevBuf<double> z_samples, cache, prod, raw_ll0, raw_ll, sample_ll_dla, sample_ll_sub, log_ev_dla, log_ev_sub, cdf;
  DevBuf<double> log_lik, log_priors, log_post, model_post, p_dla, p_no_dla, map_z, map_lognhi;
  DevBuf<double> basis;  // Gram basis panels of the batch's spectra
  DevBuf<GramBasisTask> basis_desc;
  DevBuf<int32_t> rows, inds_t, map_ind;
  DevBuf<int> alive;  // [B][4] : DLA level-loop alive flag, status, usable (constant), pad
  DevBuf<PrepTask> prep_desc;
  DevBuf<AbsorptionGrid> grid_desc;
  DevBuf<LikelihoodSpectrum> lk_desc;
  DevBuf<EvidenceLevel> ev_desc;
  DevBuf<MapTask> map_desc;
  DevBuf<GatherTask> gather_desc;
  std::vector<cudaEvent_t> events;
  // timing of the last run
  double total_ms = 0, gram_ms = 0, voigt_ms = 0, gram_flops = 0;
  long long launches = 0;
  ~dla_catalogue() {
    for (cudaEvent_t e : events) cudaEventDestroy(e);
  }
};

extern "C" int dla_catalogue_create(const dla_model* model, const dla_params* params, const dla_catalogue_config* config,
                                    const double* dla_offset_samples, const double* dla_log_nhi_samples,
                                    const double* dla_nhi_samples, const double* sub_offset_samples,
                                    const double* sub_nhi_samples, const double* uniforms, dla_catalogue** out) {
  DLA_CHECK_READY();
  Runtime& rt = runtime();
  DLA_REQUIRE(model && params && config && out, "null pointer argument");
  DLA_REQUIRE(dla_offset_samples && dla_log_nhi_samples && dla_nhi_samples && sub_offset_samples && sub_nhi_samples,
              "null sample array");
  const int S = config->num_dla_samples, max_dlas = config->max_dlas;
  DLA_REQUIRE(S >= 1 && max_dlas >= 1 && max_dlas <= LK_MAX_ROWS, "need S >= 1 and 1 <= max_dlas <= 8");
  DLA_REQUIRE(max_dlas == 1 || uniforms, "uniforms are required when max_dlas > 1");
  DLA_REQUIRE(model->dev.k == LK_K, "the batched likelihood path is built for k = 20");
  DLA_REQUIRE(params->width == INSTRUMENT_WIDTH, "instrument profile width must be 3");
  DLA_REQUIRE(params->num_lines >= 1 && params->num_lines <= LYMAN_NUM_LINES, "num_lines must be in [1, 31]");
  std::unique_ptr<dla_catalogue> cat(new dla_catalogue());
  cat->model = model;
  cat->params = *params;
  cat->S = S;
  cat->max_dlas = max_dlas;
  cat->B = config->batch_spectra > 0 ? config->batch_spectra : 64;
  cat->keep = config->keep_sample_likelihoods;
  cat->paired_offsets = memcmp(dla_offset_samples, sub_offset_samples, sizeof(double) * (size_t)S) == 0;
  DLA_CUDA(cat->dla_offsets.alloc(S));
  DLA_CUDA(cat->dla_log_nhi.alloc(S));
  DLA_CUDA(cat->sub_offsets.alloc(S));
  DLA_CUDA(cat->nhi_all.alloc(2 * (size_t)S));
  DLA_CUDA(cat->dla_offsets.upload(dla_offset_samples, S, rt.stream));
  DLA_CUDA(cat->dla_log_nhi.upload(dla_log_nhi_samples, S, rt.stream));
  DLA_CUDA(cat->sub_offsets.upload(sub_offset_samples, S, rt.stream));
  DLA_CUDA(cudaMemcpyAsync(cat->nhi_all.p, dla_nhi_samples, sizeof(double) * S, cudaMemcpyHostToDevice, rt.stream));
  DLA_CUDA(cudaMemcpyAsync(cat->nhi_all.p + S, sub_nhi_samples, sizeof(double) * S, cudaMemcpyHostToDevice, rt.stream));
  if (max_dlas > 1) {
    DLA_CUDA(cat->uniforms.alloc((size_t)(max_dlas - 1) * S));
    DLA_CUDA(cat->uniforms.upload(uniforms, (size_t)(max_dlas - 1) * S, rt.stream));
  }
  DLA_CUDA(cudaStreamSynchronize(rt.stream));
  *out = cat.release();
  return 0;
}

extern "C" int dla_catalogue_destroy(dla_catalogue* cat) {
  delete cat;
  return 0;
}

extern "C" int dla_catalogue_stage(dla_catalogue* cat, int num_spectra, const int64_t* pixel_offsets,
                                   const double* wavelengths, const double* flux, const double* noise_variance,
                                   const uint8_t* pixel_mask, const double* z_qsos, const double* log_priors_in) {
  DLA_CHECK_READY();
  Runtime& rt = runtime();
  DLA_REQUIRE(cat && pixel_offsets && wavelengths && flux && noise_variance && pixel_mask && z_qsos && log_priors_in,
              "null pointer argument");
  DLA_REQUIRE(num_spectra >= 1, "empty catalogue");
  const int64_t total = pixel_offsets[num_spectra];
  DLA_REQUIRE(pixel_offsets[0] == 0 && total >= 1, "pixel_offsets must start at 0");
  cat->Q = num_spectra;
  cat->pix_off.assign(pixel_offsets, pixel_offsets + num_spectra + 1);
  cat->z_qsos.assign(z_qsos, z_qsos + num_spectra);
  int max_n_raw = 0;
  for (int q = 0; q < num_spectra; ++q) {
    const int64_t nr = pixel_offsets[q + 1] - pixel_offsets[q];
    DLA_REQUIRE(nr >= 1 && nr < (1 << 30), "bad pixel_offsets");
    max_n_raw = std::max(max_n_raw, (int)nr);
  }
  cat->max_n_raw = max_n_raw;
  const int m = 2 + cat->max_dlas;
  DLA_CUDA(cat->wl.ensure(total));
  DLA_CUDA(cat->flux.ensure(total));
  DLA_CUDA(cat->var.ensure(total));
  DLA_CUDA(cat->mask.ensure(total));
  DLA_CUDA(cat->log_priors_in.ensure((size_t)num_spectra * m));
  DLA_CUDA(cat->wl.upload(wavelengths, total, rt.stream));
  DLA_CUDA(cat->flux.upload(flux, total, rt.stream));
  DLA_CUDA(cat->var.upload(noise_variance, total, rt.stream));
  DLA_CUDA(cat->mask.upload(pixel_mask, total, rt.stream));
  DLA_CUDA(cat->log_priors_in.upload(log_priors_in, (size_t)num_spectra * m, rt.stream));
  DLA_CUDA(cudaStreamSynchronize(rt.stream));
  return 0;
}

static cudaEvent_t cat_event(dla_catalogue* cat, size_t i) {
  while (cat->events.size() <= i) {
    cudaEvent_t e;
    cudaEventCreate(&e);
    cat->events.push_back(e);
  }
  return cat->events[i];
}

static int cat_ensure_workspace(dla_catalogue* cat) {
  const size_t B = cat->B, cap = cat->max_n_raw, S = cat->S, md = cat->max_dlas, m = 2 + md;
  const size_t w = cat->params.width;
  DLA_CUDA(cat->ind_unmasked.ensure(B * cap));
  DLA_CUDA(cat->ind.ensure(B * cap));
  DLA_CUDA(cat->x.ensure(B * cap));
  DLA_CUDA(cat->y.ensure(B * cap));
  DLA_CUDA(cat->v.ensure(B * cap));
  DLA_CUDA(cat->this_wl.ensure(B * cap));
  DLA_CUDA(cat->mu.ensure(B * cap));
  DLA_CUDA(cat->omega2.ensure(B * cap));
  DLA_CUDA(cat->M.ensure(B * cap * LK_K));
  DLA_CUDA(cat->uidx.ensure(B * cap));
  DLA_CUDA(cat->qmap.ensure(B * cap));
  DLA_CUDA(cat->unmasked_wl.ensure(B * cap));
  DLA_CUDA(cat->wl_abs.ensure(B * (cap + 2 * w)));
  DLA_CUDA(cat->padded_wl.ensure(B * (cap + 2 * w)));
  DLA_CUDA(cat->scratch.ensure(B * cap));
  DLA_CUDA(cat->scalars.ensure(B * 8));
  DLA_CUDA(cat->z_samples.ensure(B * 2 * S));
  DLA_CUDA(cat->raw_ll0.ensure(B * (2 * S + 1)));
  DLA_CUDA(cat->raw_ll.ensure(B * S));
  DLA_CUDA(cat->sample_ll_dla.ensure(B * S * md));
  DLA_CUDA(cat->sample_ll_sub.ensure(B * S));
  DLA_CUDA(cat->log_ev_dla.ensure(B * md));
  DLA_CUDA(cat->log_ev_sub.ensure(B));
  DLA_CUDA(cat->cdf.ensure(B * S));
  DLA_CUDA(cat->rows.ensure(B * S * md));
  DLA_CUDA(cat->inds_t.ensure(B * S * std::max<size_t>(md - 1, 1)));
  DLA_CUDA(cat->map_ind.ensure(B * md));
  DLA_CUDA(cat->alive.ensure(B * 4));
  DLA_CUDA(cat->log_lik.ensure(B * m));
  DLA_CUDA(cat->log_priors.ensure(B * m));
  DLA_CUDA(cat->log_post.ensure(B * m));
  DLA_CUDA(cat->model_post.ensure(B * m));
  DLA_CUDA(cat->p_dla.ensure(B));
  DLA_CUDA(cat->p_no_dla.ensure(B));
  DLA_CUDA(cat->map_z.ensure(B * md * md));
  DLA_CUDA(cat->map_lognhi.ensure(B * md * md));
  DLA_CUDA(cat->prep_desc.ensure(B));
  DLA_CUDA(cat->grid_desc.ensure(B));
  DLA_CUDA(cat->lk_desc.ensure(B * md));
  DLA_CUDA(cat->ev_desc.ensure(B * (md + 1)));
  DLA_CUDA(cat->map_desc.ensure(B));
  DLA_CUDA(cat->gather_desc.ensure(B));
  return 0;
}

extern "C" int dla_catalogue_run_staged(dla_catalogue* cat, dla_catalogue_outputs* o) {
  DLA_CHECK_READY();
  Runtime& rt = runtime();
  DLA_REQUIRE(cat && o, "null pointer argument");
  DLA_REQUIRE(cat->Q >= 1, "nothing staged");
  const int S = cat->S, md = cat->max_dlas, m = 2 + md, w = cat->params.width;
  const size_t cap = cat->max_n_raw;
  const double nan = std::numeric_limits<double>::quiet_NaN();
  int rc = cat_ensure_workspace(cat);
  if (rc) return rc;
  cat->total_ms = cat->gram_ms = cat->voigt_ms = cat->gram_flops = 0;
  const long long launches_before = rt.launches;
  PrepParams P = to_prep_params(&cat->params, 1);

  cudaEvent_t e_call0 = cat_event(cat, 0), e_call1 = cat_event(cat, 1);
  DLA_CUDA(cudaEventRecord(e_call0, rt.stream));

  std::vector<PrepTask> h_prep;
  std::vector<AbsorptionGrid> h_grid;
  std::vector<LikelihoodSpectrum> h_lk;
  std::vector<EvidenceLevel> h_ev;
  std::vector<MapTask> h_map;
  std::vector<GatherTask> h_gather;
  std::vector<double> h_scalars;
  std::vector<int> h_alive;

  for (int q0 = 0; q0 < cat->Q; q0 += cat->B) {
    const int nb = std::min(cat->B, cat->Q - q0);
    size_t ev_i = 2;
    // ---- 1. prepare --------------------------------------------------------------------------
    h_prep.resize(nb);
    for (int b = 0; b < nb; ++b) {
      const int64_t off = cat->pix_off[q0 + b];
      PrepTask t;
      t.X = nullptr;
      t.Wobs = cat->wl.p + off;
      t.Y = cat->flux.p + off;
      t.V = cat->var.p + off;
      t.mask = cat->mask.p + off;
      t.n_raw = (int)(cat->pix_off[q0 + b + 1] - off);
      t.z_qso = cat->z_qsos[q0 + b];
      t.ind_unmasked = cat->ind_unmasked.p + b * cap;
      t.ind = cat->ind.p + b * cap;
      t.x = cat->x.p + b * cap;
      t.y = cat->y.p + b * cap;
      t.v = cat->v.p + b * cap;
      t.this_wl = cat->this_wl.p + b * cap;
      t.mu = cat->mu.p + b * cap;
      t.omega2 = cat->omega2.p + b * cap;
      t.M = cat->M.p + b * cap * LK_K;
      t.uidx = cat->uidx.p + b * cap;
      t.unmasked_wl = cat->unmasked_wl.p + b * cap;
      t.wl_abs = cat->wl_abs.p + b * (cap + 2 * w);
      t.padded_wl = cat->padded_wl.p + b * (cap + 2 * w);
      t.scratch = cat->scratch.p + b * cap;
      t.scalars = cat->scalars.p + (size_t)b * 8;
      h_prep[b] = t;
    }
    DLA_CUDA(cudaMemcpyAsync(cat->prep_desc.p, h_prep.data(), sizeof(PrepTask) * nb, cudaMemcpyHostToDevice, rt.stream));
    cudaEvent_t e_begin = cat_event(cat, ev_i++);
    DLA_CUDA(cudaEventRecord(e_begin, rt.stream));
    prepare_spectrum_kernel<<<nb, 256, 0, rt.stream>>>(cat->prep_desc.p, cat->model->dev, P);
    DLA_LAUNCHED();
    h_scalars.resize((size_t)nb * 8);
    DLA_CUDA(cat->scalars.download(h_scalars.data(), (size_t)nb * 8, rt.stream));
    DLA_CUDA(cudaStreamSynchronize(rt.stream));

    // ---- 2. sizes, cache layout, descriptors ---------------------------------------------------
    std::vector<int> n_b(nb), nu_b(nb), ld_b(nb);
    std::vector<size_t> cache_off(nb), prod_off(nb), basis_off(nb);
    size_t cache_total = 0, prod_total = 0, basis_total = 0;
    int max_n_abs = 1, max_basis_rows = LK_KC;
    for (int b = 0; b < nb; ++b) {
      nu_b[b] = (int)h_scalars[(size_t)b * 8 + 0];
      n_b[b] = (int)h_scalars[(size_t)b * 8 + 1];
      ld_b[b] = (int)round_up(std::max(n_b[b], 1), 4);
      cache_off[b] = cache_total;
      cache_total += (size_t)(2 * S + 1) * ld_b[b];
      prod_off[b] = prod_total;
      if (md >= 3) prod_total += (size_t)S * ld_b[b];
      max_n_abs = std::max(max_n_abs, cat->params.broadening ? nu_b[b] + 2 * w : nu_b[b]);
      const size_t brows = round_up((size_t)std::max(n_b[b], 1), LK_KC);
      basis_off[b] = basis_total;
      basis_total += brows * LK_PSTRIDE;
      max_basis_rows = std::max(max_basis_rows, (int)brows);
    }
    DLA_CUDA(cat->cache.ensure(cache_total));
    DLA_CUDA(cat->prod.ensure(prod_total));
    DLA_CUDA(cat->basis.ensure(basis_total));
    DLA_CUDA(cat->basis_desc.ensure(nb));
    std::vector<GramBasisTask> h_basis(nb);
    for (int b = 0; b < nb; ++b) {
      h_basis[b].M = h_prep[b].M;
      h_basis[b].P = cat->basis.p + basis_off[b];
      h_basis[b].n = n_b[b];
    }
    DLA_CUDA(cudaMemcpyAsync(cat->basis_desc.p, h_basis.data(), sizeof(GramBasisTask) * nb, cudaMemcpyHostToDevice, rt.stream));

    h_grid.resize(nb);
    h_lk.assign((size_t)nb * md, LikelihoodSpectrum());
    h_ev.assign((size_t)nb * (md + 1), EvidenceLevel());
    h_map.resize(nb);
    h_gather.resize(nb);
    h_alive.assign((size_t)nb * 4, 0);
    for (int b = 0; b < nb; ++b) {
      const bool usable = n_b[b] >= 1 && isfinite(h_scalars[(size_t)b * 8 + 3]) && isfinite(h_scalars[(size_t)b * 8 + 4]);
      h_alive[(size_t)b * 4 + 0] = usable ? 1 : 0;
      h_alive[(size_t)b * 4 + 1] = usable ? 0 : 1;  // status 1: nothing to model
      h_alive[(size_t)b * 4 + 2] = usable ? 1 : 0;
      double* cache_b = cat->cache.p + cache_off[b];
      AbsorptionGrid g;
      g.wl = h_prep[b].wl_abs;
      g.uidx = h_prep[b].uidx;
      g.qmap = cat->qmap.p + (size_t)b * cap;
      g.out = cache_b;
      g.n_in = cat->params.broadening ? nu_b[b] + 2 * w : nu_b[b];
      g.n_out = n_b[b];
      g.ld = ld_b[b];
      g.num_samples = usable ? 2 * S : 0;
      g.z = cat->z_samples.p + (size_t)b * 2 * S;
      g.nhi = cat->nhi_all.p;
      g.pair_offset = cat->paired_offsets ? S : 0;
      g.lls_break = 0;
      h_grid[b] = g;
      for (int level = 0; level < md; ++level) {
        LikelihoodSpectrum d;
        d.y = h_prep[b].y;
        d.v = h_prep[b].v;
        d.mu = h_prep[b].mu;
        d.omega2 = h_prep[b].omega2;
        d.M = h_prep[b].M;
        d.P = cat->basis.p + basis_off[b];
        d.cache = cache_b;
        d.rows0 = nullptr;
        d.alive = cat->alive.p + (size_t)b * 4;
        d.n = n_b[b];
        d.ld = ld_b[b];
        d.row0 = 0;
        if (level == 0) {  // DLA + subDLA + null rows in one go
          d.base0 = cache_b;
          d.rows = nullptr;
          d.prod_out = nullptr;
          d.out = cat->raw_ll0.p + (size_t)b * (2 * S + 1);
          d.num_samples = 2 * S + 1;
          d.num_rows = 1;
          d.row_stride = 0;
        } else {
          // running product of the previous level (row s) x profile of the newly drawn absorber
          double* prod_b = cat->prod.p + prod_off[b];
          d.base0 = level == 1 ? cache_b : prod_b;
          d.rows = cat->rows.p + (size_t)b * S * md + (size_t)level * S;
          d.prod_out = (level + 1 < md) ? prod_b : nullptr;
          d.out = cat->raw_ll.p + (size_t)b * S;
          d.num_samples = S;
          d.num_rows = 2;
          d.row_stride = S;
        }
        h_lk[(size_t)level * nb + b] = d;
        EvidenceLevel e;
        e.raw_ll = level == 0 ? cat->raw_ll0.p + (size_t)b * (2 * S + 1) : cat->raw_ll.p + (size_t)b * S;
        e.sample_ll = cat->sample_ll_dla.p + (size_t)b * S * md + level;
        e.ll_stride = md;
        e.z_samples = cat->z_samples.p + (size_t)b * 2 * S;
        e.base_inds = cat->rows.p + (size_t)b * S * md + S;
        e.base_out = (level + 1 < md) ? cat->rows.p + (size_t)b * S * md + (size_t)(level + 1) * S : nullptr;
        e.uniforms = (level + 1 < md) ? cat->uniforms.p + (size_t)level * S : nullptr;
        e.log_evidence = cat->log_ev_dla.p + (size_t)b * md + level;
        e.cdf_scratch = cat->cdf.p + (size_t)b * S;
        e.alive = cat->alive.p + (size_t)b * 4;
        e.status = cat->alive.p + (size_t)b * 4 + 1;
        e.S = S;
        e.level = level;
        e.min_z_separation = cat->params.min_z_separation;
        h_ev[(size_t)level * nb + b] = e;
      }
      {  // subDLA model: one level, no resampling (subdla_gp.py, max_dlas = 1)
        EvidenceLevel e;
        e.raw_ll = cat->raw_ll0.p + (size_t)b * (2 * S + 1) + S;
        e.sample_ll = cat->sample_ll_sub.p + (size_t)b * S;
        e.ll_stride = 1;
        e.z_samples = cat->z_samples.p + (size_t)b * 2 * S + S;
        e.base_inds = nullptr;
        e.base_out = nullptr;
        e.uniforms = nullptr;
        e.log_evidence = cat->log_ev_sub.p + b;
        e.cdf_scratch = nullptr;
        e.alive = cat->alive.p + (size_t)b * 4 + 2;  // not affected by the DLA model's early exit
        e.status = nullptr;
        e.S = S;
        e.level = 0;
        e.min_z_separation = cat->params.min_z_separation;
        h_ev[(size_t)md * nb + b] = e;
      }
      MapTask mt;
      mt.sample_ll = cat->sample_ll_dla.p + (size_t)b * S * md;
      mt.base_inds = cat->rows.p + (size_t)b * S * md + S;
      mt.z_samples = cat->z_samples.p + (size_t)b * 2 * S;
      mt.log_nhi = cat->dla_log_nhi.p;
      mt.map_z = cat->map_z.p + (size_t)b * md * md;
      mt.map_log_nhi = cat->map_lognhi.p + (size_t)b * md * md;
      mt.map_ind = cat->map_ind.p + (size_t)b * md;
      mt.S = S;
      mt.max_dlas = md;
      h_map[b] = mt;
      GatherTask gt;
      gt.raw_ll0 = cat->raw_ll0.p + (size_t)b * (2 * S + 1);
      gt.log_ev_dla = cat->log_ev_dla.p + (size_t)b * md;
      gt.log_ev_sub = cat->log_ev_sub.p + b;
      gt.log_lik = cat->log_lik.p + (size_t)b * m;
      gt.S = S;
      gt.max_dlas = md;
      h_gather[b] = gt;
    }
    DLA_CUDA(cudaMemcpyAsync(cat->grid_desc.p, h_grid.data(), sizeof(AbsorptionGrid) * nb, cudaMemcpyHostToDevice, rt.stream));
    DLA_CUDA(cudaMemcpyAsync(cat->lk_desc.p, h_lk.data(), sizeof(LikelihoodSpectrum) * nb * md, cudaMemcpyHostToDevice, rt.stream));
    DLA_CUDA(cudaMemcpyAsync(cat->ev_desc.p, h_ev.data(), sizeof(EvidenceLevel) * nb * (md + 1), cudaMemcpyHostToDevice, rt.stream));
    DLA_CUDA(cudaMemcpyAsync(cat->map_desc.p, h_map.data(), sizeof(MapTask) * nb, cudaMemcpyHostToDevice, rt.stream));
    DLA_CUDA(cudaMemcpyAsync(cat->gather_desc.p, h_gather.data(), sizeof(GatherTask) * nb, cudaMemcpyHostToDevice, rt.stream));
    DLA_CUDA(cudaMemcpyAsync(cat->alive.p, h_alive.data(), sizeof(int) * nb * 4, cudaMemcpyHostToDevice, rt.stream));

    // ---- 3. launches ---------------------------------------------------------------------------
    cudaEvent_t e_begin2 = cat_event(cat, ev_i++);
    DLA_CUDA(cudaEventRecord(e_begin2, rt.stream));
    auto fill = [&](double* p, size_t count, double value) -> int {
      fill_double_kernel<<<(unsigned)((count + 255) / 256), 256, 0, rt.stream>>>(p, count, value);
      DLA_LAUNCHED();
      return 0;
    };
    if ((rc = fill(cat->sample_ll_dla.p, (size_t)nb * S * md, nan))) return rc;
    if ((rc = fill(cat->sample_ll_sub.p, (size_t)nb * S, nan))) return rc;
    if ((rc = fill(cat->log_ev_dla.p, (size_t)nb * md, nan))) return rc;
    if ((rc = fill(cat->log_ev_sub.p, (size_t)nb, nan))) return rc;
    if ((rc = fill(cat->raw_ll0.p, (size_t)nb * (2 * S + 1), nan))) return rc;
    {
      int max_ld = 1;
      for (int b = 0; b < nb; ++b) max_ld = std::max(max_ld, ld_b[b]);
      const size_t work = std::max((size_t)S * md, (size_t)max_ld);
      dim3 grid((unsigned)((work + 255) / 256), nb);
      init_rows_kernel<<<grid, 256, 0, rt.stream>>>(cat->rows.p, S, md, cat->grid_desc.p);
      DLA_LAUNCHED();
    }
    gram_basis_kernel<<<dim3((unsigned)((max_basis_rows + 7) / 8), nb), 256, 0, rt.stream>>>(cat->basis_desc.p);
    DLA_LAUNCHED();
    {
      dim3 grid((2 * S + 255) / 256, nb);
      z_samples_kernel<<<grid, 256, 0, rt.stream>>>(cat->scalars.p, 8, cat->dla_offsets.p, cat->sub_offsets.p, S, cat->z_samples.p);
      DLA_LAUNCHED();
    }
    // profiles
    cudaEvent_t e_v0 = cat_event(cat, ev_i++), e_v1 = cat_event(cat, ev_i++);
    {
      build_qmap_kernel<<<nb, 256, 0, rt.stream>>>(cat->grid_desc.p, cat->params.broadening);
      DLA_LAUNCHED();
      DLA_CUDA(cudaEventRecord(e_v0, rt.stream));
      if ((rc = launch_voigt_grids(cat->grid_desc.p, cat->paired_offsets ? S : 2 * S, nb, cat->params.num_lines,
                                   cat->params.broadening)))
        return rc;
      DLA_CUDA(cudaEventRecord(e_v1, rt.stream));
    }
    // levels
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> lk_events;
    for (int level = 0; level < md; ++level) {
      cudaEvent_t e0 = cat_event(cat, ev_i++), e1 = cat_event(cat, ev_i++);
      DLA_CUDA(cudaEventRecord(e0, rt.stream));
      const int ns = level == 0 ? 2 * S + 1 : S;
      dim3 grid((ns + LK_TS - 1) / LK_TS, nb);
      sample_likelihood_kernel<<<grid, LK_THREADS, LK_SMEM_BYTES, rt.stream>>>(cat->lk_desc.p + (size_t)level * nb);
      DLA_LAUNCHED();
      DLA_CUDA(cudaEventRecord(e1, rt.stream));
      lk_events.push_back({e0, e1});
      evidence_level_kernel<<<nb, 1024, 0, rt.stream>>>(cat->ev_desc.p + (size_t)level * nb);
      DLA_LAUNCHED();
      if (level == 0) {
        evidence_level_kernel<<<nb, 1024, 0, rt.stream>>>(cat->ev_desc.p + (size_t)md * nb);
        DLA_LAUNCHED();
      }
      for (int b = 0; b < nb; ++b)
        if (h_alive[(size_t)b * 4]) cat->gram_flops += (double)ns * (472.0 * n_b[b] + 3.1e3);
    }
    {
      dim3 grid(md, nb);
      map_kernel<<<grid, 256, 0, rt.stream>>>(cat->map_desc.p);
      DLA_LAUNCHED();
      gather_evidences_kernel<<<(nb + 127) / 128, 128, 0, rt.stream>>>(cat->gather_desc.p, nb);
      DLA_LAUNCHED();
      model_selection_kernel<<<(nb + 127) / 128, 128, 0, rt.stream>>>(
          cat->log_priors_in.p + (size_t)q0 * m, cat->log_lik.p, nb, md, cat->log_priors.p, cat->log_post.p,
          cat->model_post.p, cat->p_dla.p, cat->p_no_dla.p);
      DLA_LAUNCHED();
      if (o->base_sample_inds && md > 1) {
        dim3 tgrid((S + 255) / 256, nb);
        transpose_inds_kernel<<<tgrid, 256, 0, rt.stream>>>(cat->rows.p, S, md, cat->inds_t.p);
        DLA_LAUNCHED();
      }
    }
    cudaEvent_t e_end = cat_event(cat, ev_i++);
    DLA_CUDA(cudaEventRecord(e_end, rt.stream));

    // ---- 4. results back -------------------------------------------------------------------------
    auto d2h = [&](void* dst, const void* src, size_t bytes) -> int {
      DLA_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, rt.stream));
      return 0;
    };
    if (o->log_priors && (rc = d2h(o->log_priors + (size_t)q0 * m, cat->log_priors.p, sizeof(double) * nb * m))) return rc;
    if (o->log_likelihoods && (rc = d2h(o->log_likelihoods + (size_t)q0 * m, cat->log_lik.p, sizeof(double) * nb * m))) return rc;
    if (o->log_posteriors && (rc = d2h(o->log_posteriors + (size_t)q0 * m, cat->log_post.p, sizeof(double) * nb * m))) return rc;
    if (o->model_posteriors && (rc = d2h(o->model_posteriors + (size_t)q0 * m, cat->model_post.p, sizeof(double) * nb * m))) return rc;
    if (o->p_dlas && (rc = d2h(o->p_dlas + q0, cat->p_dla.p, sizeof(double) * nb))) return rc;
    if (o->p_no_dlas && (rc = d2h(o->p_no_dlas + q0, cat->p_no_dla.p, sizeof(double) * nb))) return rc;
    if (o->MAP_z_dlas && (rc = d2h(o->MAP_z_dlas + (size_t)q0 * md * md, cat->map_z.p, sizeof(double) * nb * md * md))) return rc;
    if (o->MAP_log_nhis && (rc = d2h(o->MAP_log_nhis + (size_t)q0 * md * md, cat->map_lognhi.p, sizeof(double) * nb * md * md))) return rc;
    if (o->sample_log_likelihoods_dla &&
        (rc = d2h(o->sample_log_likelihoods_dla + (size_t)q0 * S * md, cat->sample_ll_dla.p, sizeof(double) * nb * S * md)))
      return rc;
    if (o->sample_log_likelihoods_lls &&
        (rc = d2h(o->sample_log_likelihoods_lls + (size_t)q0 * S, cat->sample_ll_sub.p, sizeof(double) * nb * S)))
      return rc;
    if (o->base_sample_inds && md > 1 &&
        (rc = d2h(o->base_sample_inds + (size_t)q0 * S * (md - 1), cat->inds_t.p, sizeof(int32_t) * nb * S * (md - 1))))
      return rc;
    h_alive.resize((size_t)nb * 4);
    if ((rc = d2h(h_alive.data(), cat->alive.p, sizeof(int) * nb * 4))) return rc;
    DLA_CUDA(cudaStreamSynchronize(rt.stream));
    for (int b = 0; b < nb; ++b) {
      if (o->min_z_dlas) o->min_z_dlas[q0 + b] = h_scalars[(size_t)b * 8 + 5];
      if (o->max_z_dlas) o->max_z_dlas[q0 + b] = h_scalars[(size_t)b * 8 + 6];
      if (o->num_pixels) o->num_pixels[q0 + b] = n_b[b];
      if (o->status) o->status[q0 + b] = h_alive[(size_t)b * 4 + 1];
    }
    // ---- timing ------------------------------------------------------------------------------------
    float ms = 0.f;
    (void)e_begin; (void)e_begin2; (void)e_end;
    DLA_CUDA(cudaEventElapsedTime(&ms, e_v0, e_v1));
    cat->voigt_ms += ms;
    for (auto& pr : lk_events) {
      DLA_CUDA(cudaEventElapsedTime(&ms, pr.first, pr.second));
      cat->gram_ms += ms;
    }
  }
  // whole call on the library stream: kernels, descriptor uploads, result copies and host gaps
  DLA_CUDA(cudaEventRecord(e_call1, rt.stream));
  DLA_CUDA(cudaEventSynchronize(e_call1));
  {
    float ms = 0.f;
    DLA_CUDA(cudaEventElapsedTime(&ms, e_call0, e_call1));
    cat->total_ms = ms;
  }
  cat->launches = rt.launches - launches_before;
  rt.last_kernel_ms = cat->total_ms;
  return 0;
}

extern "C" int dla_catalogue_process(dla_catalogue* cat, int num_spectra, const int64_t* pixel_offsets,
                                     const double* wavelengths, const double* flux, const double* noise_variance,
                                     const uint8_t* pixel_mask, const double* z_qsos, const double* log_priors_in,
                                     dla_catalogue_outputs* outputs) {
  int rc = dla_catalogue_stage(cat, num_spectra, pixel_offsets, wavelengths, flux, noise_variance, pixel_mask, z_qsos,
                               log_priors_in);
  if (rc) return rc;
  return dla_catalogue_run_staged(cat, outputs);
}

extern "C" int dla_catalogue_last_timing(const dla_catalogue* cat, double* total_ms, double* gram_ms, double* voigt_ms,
                                         long long* launches, double* gram_flops) {
  DLA_REQUIRE(cat, "null catalogue");
  if (total_ms) *total_ms = cat->total_ms;
  if (gram_ms) *gram_ms = cat->gram_ms;
  if (voigt_ms) *voigt_ms = cat->voigt_ms;
  if (launches) *launches = cat->launches;
  if (gram_flops) *gram_flops = cat->gram_flops;
  return 0;
}
