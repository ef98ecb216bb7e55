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

// ---------------------------------------------------------------------------------------------------
// The engine is a two-deep software pipeline over batches of B spectra (tasks 5/6):
//
//   copy stream :  H2D(i+2) ...............                      (host source only; raw-input slot i & 1)
//   main stream :  prep(i+1) | compute(i) | D2H(i) | prep(i+2) | compute(i+1) | D2H(i+1) | ...
//   host        :  wait prep(i) -> size the batch, write descriptors -> enqueue compute(i) ->
//                  consume results(i-1) -> enqueue H2D(i+2), prep(i+2)
//
// prepare_spectrum_kernel of batch i+1 is queued AHEAD of batch i's kernels, so the host has the pixel counts it
// needs to size batch i+1 (profile cache, Gram basis, descriptors) long before batch i finishes: the device
// never waits for the host between batches (round 1 synchronised twice per batch).  Everything the host hands
// to or takes from an in-flight batch lives in page-locked staging owned by one of two slots (parity of the
// batch index): descriptors, the prep scalars, the result block, the optional per-sample arrays.  The large
// per-batch device workspace (profile cache, running product, basis, ...) is shared by all batches - their
// kernels are serialised by the main stream.  Device memory holds at most two batches of raw spectra, so a
// catalogue of any length streams through (the staged variant keeps the raw catalogue resident instead and
// only measures the device path).
// ---------------------------------------------------------------------------------------------------
// per-spectrum control block on the device: [0] DLA level-loop alive flag, [1] status, [2] usable (constant), [3] pad,
// [4 + level] samples evaluated at that level (written by compact_level_kernel, read by the likelihood launch)
constexpr int CAT_ALIVE = 4 + LK_MAX_ROWS;

struct CatSlot {
  // raw spectra of the batch (host source)
  DevBuf<double> wl, flux, var;
  DevBuf<uint8_t> mask;
  // prepared arrays (written by prep(i), read by compute(i))
  DevBuf<uint8_t> ind_unmasked, ind;
  DevBuf<double> x, y, v, this_wl, mu, omega2, M, unmasked_wl, wl_abs, padded_wl, scratch, scalars;
  DevBuf<int32_t> uidx;
  DevBuf<PrepTask> prep_desc;
  // page-locked host staging
  PinnedBuf<PrepTask> h_prep;
  PinnedBuf<double> h_scalars;          // [B][8]
  PinnedBuf<unsigned char> h_desc;      // all descriptors of compute(i), one block
  PinnedBuf<double> h_res;              // result block of the batch
  PinnedBuf<int> h_alive;               // [B][CAT_ALIVE]
  std::vector<unsigned char> usable_b;  // spectrum had pixels to model (level-0 evaluations ran)
  PinnedBuf<double> h_sample_dla, h_sample_sub;  // optional per-sample arrays
  PinnedBuf<int32_t> h_inds;
  // bookkeeping: (q0, nb) of the batch whose prep is queued; (c_q0, c_nb) of the batch whose compute is queued - the
  // slot's next prep is enqueued before the previous compute's results are consumed
  int q0 = 0, nb = 0, c_q0 = 0, c_nb = 0;
  std::vector<int> n_b;
  std::vector<double> zmin_b, zmax_b;
  cudaEvent_t ev_upload = nullptr, ev_prep = nullptr, ev_done = nullptr, ev_v0 = nullptr, ev_v1 = nullptr;
  cudaEvent_t ev_lk[2 * LK_MAX_ROWS] = {nullptr};
  bool prep_recorded = false;
  ~CatSlot() {
    for (cudaEvent_t e : {ev_upload, ev_prep, ev_done, ev_v0, ev_v1})
      if (e) cudaEventDestroy(e);
    for (cudaEvent_t e : ev_lk)
      if (e) cudaEventDestroy(e);
  }
};

struct dla_catalogue {
  const dla_model* model = nullptr;
  dla_params params;
  int device = -1;
  int S = 0, max_dlas = 0, B = 0, keep = 0;
  bool paired_offsets = false;  // DLA and subDLA samples share their redshift offsets: one line-sum evaluation for both
  // constants
  DevBuf<double> dla_offsets, dla_log_nhi, sub_offsets, nhi_all /* [dla_nhi ; sub_nhi] */, uniforms;
  // staged inputs (dla_catalogue_stage)
  int Q = 0;
  std::vector<int64_t> pix_off;
  std::vector<double> z_qsos;
  DevBuf<double> wl, flux, var, log_priors_in;
  DevBuf<uint8_t> mask;
  int max_n_raw = 0;
  DevBuf<double> log_priors_call;  // priors of a dla_catalogue_process call
  // the two pipeline slots
  CatSlot slot[2];
  // per-batch compute workspace, shared by all batches (stream-ordered)
  DevBuf<int32_t> qmap;
  DevBuf<double> z_samples, cache, prod, raw_ll0, raw_ll, sample_ll_dla, sample_ll_sub, log_ev_dla, log_ev_sub, cdf;
  DevBuf<double> res;    // [log_priors | log_lik | log_post | model_post] (B x m each), p_dla, p_no_dla (B), map_z, map_lognhi (B x md x md)
  DevBuf<double> basis;  // Gram basis panels of the batch's spectra
  DevBuf<int32_t> rows, inds_t, map_ind;
  // compacted launches of levels >= 1 (compact_level_kernel): per spectrum S entries each
  DevBuf<int32_t> sel, pos_a, pos_b, rows0_c, rows1_c;
  DevBuf<double> raw_slots;  // [B][S] raw log-likelihoods of the current level in slot order
  DevBuf<int> alive;  // [B][CAT_ALIVE] control blocks
  DevBuf<unsigned char> desc;  // device copy of the descriptor block
  cudaEvent_t ev_call0 = nullptr, ev_call1 = nullptr;
  // timing of the last run
  double total_ms = 0, gram_ms = 0, voigt_ms = 0, gram_flops = 0;
  long long launches = 0;
  long long lk_evaluated = 0, lk_masked = 0;  // likelihood evaluations run / left out by the separation mask
  ~dla_catalogue() {
    if (ev_call0) cudaEventDestroy(ev_call0);
    if (ev_call1) cudaEventDestroy(ev_call1);
  }
};

// where the raw spectra of a run come from
struct CatSource {
  bool on_device;            // staged: pointers are device arrays of the whole catalogue
  const int64_t* pix_off;    // Q + 1
  const double* wl;
  const double* flux;
  const double* var;
  const uint8_t* mask;
  const double* z_qsos;
  int Q;
  int max_n_raw;
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
  DLA_REQUIRE(model->device == rt.device, "the model lives on another device than the one selected by dla_init");
  std::unique_ptr<dla_catalogue> cat(new dla_catalogue());
  cat->model = model;
  cat->params = *params;
  cat->device = rt.device;
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
  DLA_CUDA(cudaEventCreate(&cat->ev_call0));
  DLA_CUDA(cudaEventCreate(&cat->ev_call1));
  for (CatSlot& sl : cat->slot) {
    DLA_CUDA(cudaEventCreateWithFlags(&sl.ev_upload, cudaEventDisableTiming));
    DLA_CUDA(cudaEventCreateWithFlags(&sl.ev_prep, cudaEventDisableTiming));
    DLA_CUDA(cudaEventCreateWithFlags(&sl.ev_done, cudaEventDisableTiming));
    DLA_CUDA(cudaEventCreate(&sl.ev_v0));
    DLA_CUDA(cudaEventCreate(&sl.ev_v1));
    for (int i = 0; i < 2 * max_dlas; ++i) DLA_CUDA(cudaEventCreate(&sl.ev_lk[i]));
  }
  DLA_CUDA(cudaStreamSynchronize(rt.stream));
  *out = cat.release();
  return 0;
}

extern "C" int dla_catalogue_destroy(dla_catalogue* cat) {
  if (cat) {
    cudaStreamSynchronize(runtime().stream);
    cudaStreamSynchronize(runtime().copy_stream);
  }
  delete cat;
  return 0;
}

static int cat_check_offsets(int num_spectra, const int64_t* pixel_offsets, int* max_n_raw) {
  DLA_REQUIRE(num_spectra >= 1, "empty catalogue");
  const int64_t total = pixel_offsets[num_spectra];
  DLA_REQUIRE(pixel_offsets[0] == 0 && total >= 1, "pixel_offsets must start at 0");
  int mx = 0;
  for (int q = 0; q < num_spectra; ++q) {
    const int64_t nr = pixel_offsets[q + 1] - pixel_offsets[q];
    DLA_REQUIRE(nr >= 1 && nr < (1 << 30), "bad pixel_offsets");
    mx = std::max(mx, (int)nr);
  }
  *max_n_raw = mx;
  return 0;
}

extern "C" int dla_catalogue_stage(dla_catalogue* cat, int num_spectra, const int64_t* pixel_offsets,
                                   const double* wavelengths, const double* flux, const double* noise_variance,
                                   const uint8_t* pixel_mask, const double* z_qsos, const double* log_priors_in) {
  DLA_CHECK_READY();
  Runtime& rt = runtime();
  DLA_REQUIRE(cat && pixel_offsets && wavelengths && flux && noise_variance && pixel_mask && z_qsos && log_priors_in,
              "null pointer argument");
  DLA_REQUIRE(cat->device == rt.device, "the catalogue lives on another device than the one selected by dla_init");
  int max_n_raw = 0;
  if (int rc = cat_check_offsets(num_spectra, pixel_offsets, &max_n_raw)) return rc;
  const int64_t total = pixel_offsets[num_spectra];
  cat->Q = num_spectra;
  cat->pix_off.assign(pixel_offsets, pixel_offsets + num_spectra + 1);
  cat->z_qsos.assign(z_qsos, z_qsos + num_spectra);
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

// layout of the device result block for a batch capacity of B spectra (offsets in doubles)
struct CatResLayout {
  size_t log_priors, log_lik, log_post, model_post, p_dla, p_no_dla, map_z, map_lognhi, total;
};
static CatResLayout cat_res_layout(size_t B, size_t m, size_t md) {
  CatResLayout L;
  size_t o = 0;
  L.log_priors = o; o += B * m;
  L.log_lik = o; o += B * m;
  L.log_post = o; o += B * m;
  L.model_post = o; o += B * m;
  L.p_dla = o; o += B;
  L.p_no_dla = o; o += B;
  L.map_z = o; o += B * md * md;
  L.map_lognhi = o; o += B * md * md;
  L.total = o;
  return L;
}

// descriptor block of one batch: byte offsets of the typed arrays inside it
struct CatDescLayout {
  size_t basis, grid, lk, ev, map, gather, compact, scatter, total;
};
static CatDescLayout cat_desc_layout(size_t B, size_t md) {
  auto up = [](size_t x) { return (x + 255) / 256 * 256; };
  CatDescLayout L;
  size_t o = 0;
  L.basis = o; o = up(o + sizeof(GramBasisTask) * B);
  L.grid = o; o = up(o + sizeof(AbsorptionGrid) * B);
  L.lk = o; o = up(o + sizeof(LikelihoodSpectrum) * B * md);
  L.ev = o; o = up(o + sizeof(EvidenceLevel) * B * (md + 1));
  L.map = o; o = up(o + sizeof(MapTask) * B);
  L.gather = o; o = up(o + sizeof(GatherTask) * B);
  L.compact = o; o = up(o + sizeof(CompactTask) * B * md);
  L.scatter = o; o = up(o + sizeof(ScatterTask) * B);
  L.total = o;
  return L;
}

static int cat_ensure_workspace(dla_catalogue* cat, size_t cap, bool host_source, const dla_catalogue_outputs* o) {
  const size_t B = cat->B, S = cat->S, md = cat->max_dlas, m = 2 + md;
  const size_t w = cat->params.width;
  for (CatSlot& sl : cat->slot) {
    if (host_source) {
      DLA_CUDA(sl.wl.ensure(B * cap));
      DLA_CUDA(sl.flux.ensure(B * cap));
      DLA_CUDA(sl.var.ensure(B * cap));
      DLA_CUDA(sl.mask.ensure(B * cap));
    }
    DLA_CUDA(sl.ind_unmasked.ensure(B * cap));
    DLA_CUDA(sl.ind.ensure(B * cap));
    DLA_CUDA(sl.x.ensure(B * cap));
    DLA_CUDA(sl.y.ensure(B * cap));
    DLA_CUDA(sl.v.ensure(B * cap));
    DLA_CUDA(sl.this_wl.ensure(B * cap));
    DLA_CUDA(sl.mu.ensure(B * cap));
    DLA_CUDA(sl.omega2.ensure(B * cap));
    DLA_CUDA(sl.M.ensure(B * cap * LK_K));
    DLA_CUDA(sl.uidx.ensure(B * cap));
    DLA_CUDA(sl.unmasked_wl.ensure(B * cap));
    DLA_CUDA(sl.wl_abs.ensure(B * (cap + 2 * w)));
    DLA_CUDA(sl.padded_wl.ensure(B * (cap + 2 * w)));
    DLA_CUDA(sl.scratch.ensure(B * cap));
    DLA_CUDA(sl.scalars.ensure(B * 8));
    DLA_CUDA(sl.prep_desc.ensure(B));
    DLA_CUDA(sl.h_prep.ensure(B));
    DLA_CUDA(sl.h_scalars.ensure(B * 8));
    DLA_CUDA(sl.h_desc.ensure(cat_desc_layout(B, md).total));
    DLA_CUDA(sl.h_res.ensure(cat_res_layout(B, m, md).total));
    DLA_CUDA(sl.h_alive.ensure(B * CAT_ALIVE));
    if (o->sample_log_likelihoods_dla) DLA_CUDA(sl.h_sample_dla.ensure(B * S * md));
    if (o->sample_log_likelihoods_lls) DLA_CUDA(sl.h_sample_sub.ensure(B * S));
    if (o->base_sample_inds && md > 1) DLA_CUDA(sl.h_inds.ensure(B * S * (md - 1)));
  }
  DLA_CUDA(cat->qmap.ensure(B * cap));
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
  DLA_CUDA(cat->alive.ensure(B * CAT_ALIVE));
  DLA_CUDA(cat->sel.ensure(B * S));
  DLA_CUDA(cat->pos_a.ensure(B * S));
  DLA_CUDA(cat->pos_b.ensure(B * S));
  DLA_CUDA(cat->rows0_c.ensure(B * S));
  DLA_CUDA(cat->rows1_c.ensure(B * S));
  DLA_CUDA(cat->raw_slots.ensure(B * S));
  DLA_CUDA(cat->res.ensure(cat_res_layout(B, m, md).total));
  DLA_CUDA(cat->desc.ensure(cat_desc_layout(B, md).total));
  return 0;
}

// ---- pipeline stage: raw spectra of batch `bi` -> device slot (host source), on the copy stream ---------------
static int cat_enqueue_upload(dla_catalogue* cat, const CatSource& src, int bi) {
  Runtime& rt = runtime();
  CatSlot& sl = cat->slot[bi & 1];
  const int q0 = bi * cat->B, nb = std::min(cat->B, src.Q - q0);
  const int64_t a = src.pix_off[q0], total = src.pix_off[q0 + nb] - a;
  // the slot's raw buffers were last read by prep(bi - 2)
  if (sl.prep_recorded) DLA_CUDA(cudaStreamWaitEvent(rt.copy_stream, sl.ev_prep, 0));
  DLA_CUDA(cudaMemcpyAsync(sl.wl.p, src.wl + a, sizeof(double) * total, cudaMemcpyHostToDevice, rt.copy_stream));
  DLA_CUDA(cudaMemcpyAsync(sl.flux.p, src.flux + a, sizeof(double) * total, cudaMemcpyHostToDevice, rt.copy_stream));
  DLA_CUDA(cudaMemcpyAsync(sl.var.p, src.var + a, sizeof(double) * total, cudaMemcpyHostToDevice, rt.copy_stream));
  DLA_CUDA(cudaMemcpyAsync(sl.mask.p, src.mask + a, total, cudaMemcpyHostToDevice, rt.copy_stream));
  DLA_CUDA(cudaEventRecord(sl.ev_upload, rt.copy_stream));
  return 0;
}

// ---- pipeline stage: prepare_spectrum_kernel of batch `bi` + read-back of its scalars --------------------------
static int cat_enqueue_prep(dla_catalogue* cat, const CatSource& src, int bi, const PrepParams& P, size_t cap) {
  Runtime& rt = runtime();
  CatSlot& sl = cat->slot[bi & 1];
  const int w = cat->params.width;
  const int q0 = bi * cat->B, nb = std::min(cat->B, src.Q - q0);
  sl.q0 = q0;
  sl.nb = nb;
  const int64_t base = src.pix_off[q0];
  for (int b = 0; b < nb; ++b) {
    const int64_t off = src.pix_off[q0 + b];
    PrepTask t;
    t.X = nullptr;
    if (src.on_device) {
      t.Wobs = src.wl + off;
      t.Y = src.flux + off;
      t.V = src.var + off;
      t.mask = src.mask + off;
    } else {
      t.Wobs = sl.wl.p + (off - base);
      t.Y = sl.flux.p + (off - base);
      t.V = sl.var.p + (off - base);
      t.mask = sl.mask.p + (off - base);
    }
    t.n_raw = (int)(src.pix_off[q0 + b + 1] - off);
    t.z_qso = src.z_qsos[q0 + b];
    t.ind_unmasked = sl.ind_unmasked.p + b * cap;
    t.ind = sl.ind.p + b * cap;
    t.x = sl.x.p + b * cap;
    t.y = sl.y.p + b * cap;
    t.v = sl.v.p + b * cap;
    t.this_wl = sl.this_wl.p + b * cap;
    t.mu = sl.mu.p + b * cap;
    t.omega2 = sl.omega2.p + b * cap;
    t.M = sl.M.p + b * cap * LK_K;
    t.uidx = sl.uidx.p + b * cap;
    t.unmasked_wl = sl.unmasked_wl.p + b * cap;
    t.wl_abs = sl.wl_abs.p + b * (cap + 2 * w);
    t.padded_wl = sl.padded_wl.p + b * (cap + 2 * w);
    t.scratch = sl.scratch.p + b * cap;
    t.scalars = sl.scalars.p + (size_t)b * 8;
    sl.h_prep.p[b] = t;
  }
  if (!src.on_device) DLA_CUDA(cudaStreamWaitEvent(rt.stream, sl.ev_upload, 0));
  DLA_CUDA(cudaMemcpyAsync(sl.prep_desc.p, sl.h_prep.p, sizeof(PrepTask) * nb, cudaMemcpyHostToDevice, rt.stream));
  prepare_spectrum_kernel<<<nb, 256, 0, rt.stream>>>(sl.prep_desc.p, cat->model->dev, P);
  DLA_LAUNCHED();
  DLA_CUDA(cudaMemcpyAsync(sl.h_scalars.p, sl.scalars.p, sizeof(double) * nb * 8, cudaMemcpyDeviceToHost, rt.stream));
  DLA_CUDA(cudaEventRecord(sl.ev_prep, rt.stream));
  sl.prep_recorded = true;
  return 0;
}

// ---- pipeline stage: size batch `bi`, write its descriptors, enqueue every kernel and the result read-back ------
static int cat_enqueue_compute(dla_catalogue* cat, int bi, size_t cap, const dla_catalogue_outputs* o,
                               const double* d_log_priors_in, int* open_ranges) {
  Runtime& rt = runtime();
  CatSlot& sl = cat->slot[bi & 1];
  const int S = cat->S, md = cat->max_dlas, m = 2 + md, w = cat->params.width;
  const int nb = sl.nb, q0 = sl.q0;
  sl.c_nb = nb;
  sl.c_q0 = q0;
  const size_t B = cat->B;
  const double nan = std::numeric_limits<double>::quiet_NaN();
  int rc = 0;
  DLA_CUDA(cudaEventSynchronize(sl.ev_prep));  // long done: prep(bi) was queued ahead of compute(bi - 1)
  const double* h_scalars = sl.h_scalars.p;

  // ---- sizes, cache layout --------------------------------------------------------------------
  sl.n_b.resize(nb);
  sl.zmin_b.resize(nb);
  sl.zmax_b.resize(nb);
  std::vector<int> nu_b(nb), ld_b(nb);
  std::vector<size_t> cache_off(nb), prod_off(nb), basis_off(nb);
  size_t cache_total = 0, prod_total = 0, basis_total = 0;
  int max_basis_rows = LK_KC, max_ld = 1;
  for (int b = 0; b < nb; ++b) {
    nu_b[b] = (int)h_scalars[(size_t)b * 8 + 0];
    sl.n_b[b] = (int)h_scalars[(size_t)b * 8 + 1];
    sl.zmin_b[b] = h_scalars[(size_t)b * 8 + 5];
    sl.zmax_b[b] = h_scalars[(size_t)b * 8 + 6];
    ld_b[b] = (int)round_up(std::max(sl.n_b[b], 1), 4);
    max_ld = std::max(max_ld, ld_b[b]);
    cache_off[b] = cache_total;
    cache_total += (size_t)(2 * S + 1) * ld_b[b];
    prod_off[b] = prod_total;
    if (md >= 3) prod_total += (size_t)S * ld_b[b] * (md >= 4 ? 2 : 1);  // ping-pong: level L reads what L - 1 wrote
    const size_t brows = round_up((size_t)std::max(sl.n_b[b], 1), LK_KC);
    basis_off[b] = basis_total;
    basis_total += brows * LK_PSTRIDE;
    max_basis_rows = std::max(max_basis_rows, (int)brows);
  }
  // grown with head-room: a re-allocation frees memory the previous batch may still be using, which CUDA
  // resolves by synchronising the device - correct, but a pipeline bubble
  if (cache_total > cat->cache.n) DLA_CUDA(cat->cache.ensure(cache_total + cache_total / 8));
  if (prod_total > cat->prod.n) DLA_CUDA(cat->prod.ensure(prod_total + prod_total / 8));
  if (basis_total > cat->basis.n) DLA_CUDA(cat->basis.ensure(basis_total + basis_total / 8));

  // ---- descriptors, written straight into the slot's page-locked block ---------------------------------
  const CatDescLayout DL = cat_desc_layout(B, md);
  const CatResLayout RL = cat_res_layout(B, m, md);
  unsigned char* hd = sl.h_desc.p;
  GramBasisTask* h_basis = reinterpret_cast<GramBasisTask*>(hd + DL.basis);
  AbsorptionGrid* h_grid = reinterpret_cast<AbsorptionGrid*>(hd + DL.grid);
  LikelihoodSpectrum* h_lk = reinterpret_cast<LikelihoodSpectrum*>(hd + DL.lk);
  EvidenceLevel* h_ev = reinterpret_cast<EvidenceLevel*>(hd + DL.ev);
  MapTask* h_map = reinterpret_cast<MapTask*>(hd + DL.map);
  GatherTask* h_gather = reinterpret_cast<GatherTask*>(hd + DL.gather);
  CompactTask* h_compact = reinterpret_cast<CompactTask*>(hd + DL.compact);
  const CompactTask* d_compact = reinterpret_cast<const CompactTask*>(cat->desc.p + DL.compact);
  ScatterTask* h_scatter = reinterpret_cast<ScatterTask*>(hd + DL.scatter);
  const ScatterTask* d_scatter = reinterpret_cast<const ScatterTask*>(cat->desc.p + DL.scatter);
  sl.usable_b.assign(nb, 0);
  const GramBasisTask* d_basis = reinterpret_cast<const GramBasisTask*>(cat->desc.p + DL.basis);
  const AbsorptionGrid* d_grid = reinterpret_cast<const AbsorptionGrid*>(cat->desc.p + DL.grid);
  const LikelihoodSpectrum* d_lk = reinterpret_cast<const LikelihoodSpectrum*>(cat->desc.p + DL.lk);
  const EvidenceLevel* d_ev = reinterpret_cast<const EvidenceLevel*>(cat->desc.p + DL.ev);
  const MapTask* d_map = reinterpret_cast<const MapTask*>(cat->desc.p + DL.map);
  const GatherTask* d_gather = reinterpret_cast<const GatherTask*>(cat->desc.p + DL.gather);
  int* h_alive = sl.h_alive.p;
  double* res = cat->res.p;
  for (int b = 0; b < nb; ++b) {
    const PrepTask& pt = sl.h_prep.p[b];
    const int n = sl.n_b[b];
    const bool usable = n >= 1 && isfinite(h_scalars[(size_t)b * 8 + 3]) && isfinite(h_scalars[(size_t)b * 8 + 4]);
    h_alive[(size_t)b * CAT_ALIVE + 0] = usable ? 1 : 0;
    h_alive[(size_t)b * CAT_ALIVE + 1] = usable ? 0 : 1;  // status 1: nothing to model
    h_alive[(size_t)b * CAT_ALIVE + 2] = usable ? 1 : 0;
    h_alive[(size_t)b * CAT_ALIVE + 3] = 0;
    for (int level = 0; level < LK_MAX_ROWS; ++level) h_alive[(size_t)b * CAT_ALIVE + 4 + level] = 0;  // samples evaluated per level
    sl.usable_b[b] = usable ? 1 : 0;
    double* cache_b = cat->cache.p + cache_off[b];
    h_basis[b].M = pt.M;
    h_basis[b].P = cat->basis.p + basis_off[b];
    h_basis[b].n = n;
    AbsorptionGrid g;
    g.wl = pt.wl_abs;
    g.uidx = pt.uidx;
    g.qmap = cat->qmap.p + (size_t)b * cap;
    g.out = cache_b;
    g.n_in = cat->params.broadening ? nu_b[b] + 2 * w : nu_b[b];
    g.n_out = n;
    g.ld = ld_b[b];
    g.num_samples = usable ? 2 * S : 0;
    g.z = cat->z_samples.p + (size_t)b * 2 * S;
    g.nhi = cat->nhi_all.p;
    g.pair_offset = cat->paired_offsets ? S : 0;
    g.lls_break = 0;
    h_grid[b] = g;
    for (int level = 0; level < md; ++level) {
      LikelihoodSpectrum d;
      d.y = pt.y;
      d.v = pt.v;
      d.mu = pt.mu;
      d.omega2 = pt.omega2;
      d.M = pt.M;
      d.P = cat->basis.p + basis_off[b];
      d.cache = cache_b;
      d.rows0 = nullptr;
      d.alive = cat->alive.p + (size_t)b * CAT_ALIVE;
      d.n = n;
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
        // Running product of the previous level x profile of the newly drawn absorber, over the samples that pass the
        // separation test only (compact_level_kernel, launched just before): slot i evaluates sample sel[i]; its
        // factor-0 row is rows0_c[i] (level 1: the sample's own profile in the cache; later: the slot it had in the
        // previous level's launch, whose product rows are in slot order), its factor-1 row rows1_c[i] is the profile of
        // the absorber drawn for this level.  Products ping-pong between two buffers (level L writes buffer L & 1 and
        // reads buffer (L - 1) & 1), so no launch reads rows that the same launch writes.
        double* prod_b = cat->prod.p + prod_off[b];
        const size_t prod_half = (size_t)S * ld_b[b];
        double* prod_write = prod_b + ((md >= 4 && (level & 1) == 0) ? prod_half : 0);
        const double* prod_read = prod_b + ((md >= 4 && ((level - 1) & 1) == 0) ? prod_half : 0);
        d.base0 = level == 1 ? cache_b : prod_read;
        d.rows0 = cat->rows0_c.p + (size_t)b * S;
        d.rows = cat->rows1_c.p + (size_t)b * S;
        d.prod_out = (level + 1 < md) ? prod_write : nullptr;
        d.out = cat->raw_slots.p + (size_t)b * S;
        d.num_samples = -(4 + level);  // the count is alive[4 + level], written by compact_level_kernel
        d.num_rows = 2;
        d.row_stride = S;
        CompactTask ct;
        ct.z_samples = cat->z_samples.p + (size_t)b * 2 * S;
        ct.base_inds = cat->rows.p + (size_t)b * S * md + S;
        ct.alive = cat->alive.p + (size_t)b * CAT_ALIVE;
        ct.pos_prev = level == 1 ? nullptr : ((level & 1) ? cat->pos_b.p : cat->pos_a.p) + (size_t)b * S;
        ct.pos = ((level & 1) ? cat->pos_a.p : cat->pos_b.p) + (size_t)b * S;
        ct.sel = cat->sel.p + (size_t)b * S;
        ct.rows0 = cat->rows0_c.p + (size_t)b * S;
        ct.rows1 = cat->rows1_c.p + (size_t)b * S;
        ct.raw_ll = cat->raw_ll.p + (size_t)b * S;
        ct.num_sel = cat->alive.p + (size_t)b * CAT_ALIVE + 4 + level;
        ct.S = S;
        ct.level = level;
        ct.min_z_separation = cat->params.min_z_separation;
        h_compact[(size_t)level * nb + b] = ct;
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
      e.alive = cat->alive.p + (size_t)b * CAT_ALIVE;
      e.status = cat->alive.p + (size_t)b * CAT_ALIVE + 1;
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
      e.alive = cat->alive.p + (size_t)b * CAT_ALIVE + 2;  // not affected by the DLA model's early exit
      e.status = nullptr;
      e.S = S;
      e.level = 0;
      e.min_z_separation = cat->params.min_z_separation;
      h_ev[(size_t)md * nb + b] = e;
    }
    ScatterTask sc;
    sc.raw_slots = cat->raw_slots.p + (size_t)b * S;
    sc.sel = cat->sel.p + (size_t)b * S;
    sc.num_sel = cat->alive.p + (size_t)b * CAT_ALIVE + 4;  // scatter_ll_kernel reads entry `level`
    sc.raw_ll = cat->raw_ll.p + (size_t)b * S;
    h_scatter[b] = sc;
    MapTask mt;
    mt.sample_ll = cat->sample_ll_dla.p + (size_t)b * S * md;
    mt.base_inds = cat->rows.p + (size_t)b * S * md + S;
    mt.z_samples = cat->z_samples.p + (size_t)b * 2 * S;
    mt.log_nhi = cat->dla_log_nhi.p;
    mt.map_z = res + RL.map_z + (size_t)b * md * md;
    mt.map_log_nhi = res + RL.map_lognhi + (size_t)b * md * md;
    mt.map_ind = cat->map_ind.p + (size_t)b * md;
    mt.S = S;
    mt.max_dlas = md;
    h_map[b] = mt;
    GatherTask gt;
    gt.raw_ll0 = cat->raw_ll0.p + (size_t)b * (2 * S + 1);
    gt.log_ev_dla = cat->log_ev_dla.p + (size_t)b * md;
    gt.log_ev_sub = cat->log_ev_sub.p + b;
    gt.log_lik = res + RL.log_lik + (size_t)b * m;
    gt.S = S;
    gt.max_dlas = md;
    h_gather[b] = gt;
  }
  DLA_CUDA(cudaMemcpyAsync(cat->desc.p, hd, DL.total, cudaMemcpyHostToDevice, rt.stream));
  DLA_CUDA(cudaMemcpyAsync(cat->alive.p, h_alive, sizeof(int) * nb * CAT_ALIVE, cudaMemcpyHostToDevice, rt.stream));

  // ---- launches ------------------------------------------------------------------------------
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
    const size_t work = std::max((size_t)S * md, (size_t)max_ld);
    dim3 grid((unsigned)((work + 255) / 256), nb);
    init_rows_kernel<<<grid, 256, 0, rt.stream>>>(cat->rows.p, S, md, d_grid);
    DLA_LAUNCHED();
  }
  gram_basis_kernel<<<dim3((unsigned)((max_basis_rows + 7) / 8), nb), 256, 0, rt.stream>>>(d_basis);
  DLA_LAUNCHED();
  {
    dim3 grid((2 * S + 255) / 256, nb);
    z_samples_kernel<<<grid, 256, 0, rt.stream>>>(sl.scalars.p, 8, cat->dla_offsets.p, cat->sub_offsets.p, S, cat->z_samples.p);
    DLA_LAUNCHED();
  }
  build_qmap_kernel<<<nb, 256, 0, rt.stream>>>(d_grid, cat->params.broadening);
  DLA_LAUNCHED();
  DLA_CUDA(cudaEventRecord(sl.ev_v0, rt.stream));
  if ((rc = launch_voigt_grids(d_grid, cat->paired_offsets ? S : 2 * S, nb, cat->params.num_lines, cat->params.broadening)))
    return rc;
  DLA_CUDA(cudaEventRecord(sl.ev_v1, rt.stream));
  for (int level = 0; level < md; ++level) {
    char lvl[32];
    snprintf(lvl, sizeof(lvl), "level %d", level);
    nvtxRangePushA(lvl);  // popped below; the error paths inside return with it open and cat_run pops what is left
    ++*open_ranges;
    if (level > 0) {
      compact_level_kernel<<<nb, 1024, 0, rt.stream>>>(d_compact + (size_t)level * nb);
      DLA_LAUNCHED();
    }
    DLA_CUDA(cudaEventRecord(sl.ev_lk[2 * level], rt.stream));
    const int ns = level == 0 ? 2 * S + 1 : S;
    dim3 grid((ns + LK_TS - 1) / LK_TS, nb);
    sample_likelihood_kernel<<<grid, LK_THREADS, LK_SMEM_BYTES, rt.stream>>>(d_lk + (size_t)level * nb);
    DLA_LAUNCHED();
    DLA_CUDA(cudaEventRecord(sl.ev_lk[2 * level + 1], rt.stream));
    if (level > 0) {
      scatter_ll_kernel<<<dim3((S + 255) / 256, nb), 256, 0, rt.stream>>>(d_scatter, level);
      DLA_LAUNCHED();
    }
    evidence_level_kernel<<<nb, 1024, 0, rt.stream>>>(d_ev + (size_t)level * nb);
    DLA_LAUNCHED();
    if (level == 0) {
      evidence_level_kernel<<<nb, 1024, 0, rt.stream>>>(d_ev + (size_t)md * nb);
      DLA_LAUNCHED();
    }
    nvtxRangePop();
    --*open_ranges;
  }
  {
    dim3 grid(md, nb);
    map_kernel<<<grid, 256, 0, rt.stream>>>(d_map);
    DLA_LAUNCHED();
    gather_evidences_kernel<<<(nb + 127) / 128, 128, 0, rt.stream>>>(d_gather, nb);
    DLA_LAUNCHED();
    model_selection_kernel<<<(nb + 127) / 128, 128, 0, rt.stream>>>(
        d_log_priors_in + (size_t)q0 * m, res + RL.log_lik, nb, md, res + RL.log_priors, res + RL.log_post,
        res + RL.model_post, res + RL.p_dla, res + RL.p_no_dla);
    DLA_LAUNCHED();
    if (o->base_sample_inds && md > 1) {
      dim3 tgrid((S + 255) / 256, nb);
      transpose_inds_kernel<<<tgrid, 256, 0, rt.stream>>>(cat->rows.p, S, md, cat->inds_t.p);
      DLA_LAUNCHED();
    }
  }
  // ---- results -> the slot's page-locked staging (the host picks them up one batch later) ----------------
  DLA_CUDA(cudaMemcpyAsync(sl.h_res.p, res, sizeof(double) * RL.total, cudaMemcpyDeviceToHost, rt.stream));
  DLA_CUDA(cudaMemcpyAsync(sl.h_alive.p, cat->alive.p, sizeof(int) * nb * CAT_ALIVE, cudaMemcpyDeviceToHost, rt.stream));
  if (o->sample_log_likelihoods_dla)
    DLA_CUDA(cudaMemcpyAsync(sl.h_sample_dla.p, cat->sample_ll_dla.p, sizeof(double) * nb * S * md, cudaMemcpyDeviceToHost, rt.stream));
  if (o->sample_log_likelihoods_lls)
    DLA_CUDA(cudaMemcpyAsync(sl.h_sample_sub.p, cat->sample_ll_sub.p, sizeof(double) * nb * S, cudaMemcpyDeviceToHost, rt.stream));
  if (o->base_sample_inds && md > 1)
    DLA_CUDA(cudaMemcpyAsync(sl.h_inds.p, cat->inds_t.p, sizeof(int32_t) * nb * S * (md - 1), cudaMemcpyDeviceToHost, rt.stream));
  DLA_CUDA(cudaEventRecord(sl.ev_done, rt.stream));
  return 0;
}

// ---- pipeline stage: results of batch `bi` from the slot's staging into the caller's arrays ----------------------
static int cat_consume(dla_catalogue* cat, int bi, dla_catalogue_outputs* o) {
  CatSlot& sl = cat->slot[bi & 1];
  const size_t S = cat->S, md = cat->max_dlas, m = 2 + md, B = cat->B;
  const size_t nb = sl.c_nb, q0 = sl.c_q0;
  DLA_CUDA(cudaEventSynchronize(sl.ev_done));
  const CatResLayout RL = cat_res_layout(B, m, md);
  const double* r = sl.h_res.p;
  auto put = [&](double* dst, size_t dst_off, size_t src_off, size_t count) {
    if (dst) memcpy(dst + dst_off, r + src_off, sizeof(double) * count);
  };
  put(o->log_priors, q0 * m, RL.log_priors, nb * m);
  put(o->log_likelihoods, q0 * m, RL.log_lik, nb * m);
  put(o->log_posteriors, q0 * m, RL.log_post, nb * m);
  put(o->model_posteriors, q0 * m, RL.model_post, nb * m);
  put(o->p_dlas, q0, RL.p_dla, nb);
  put(o->p_no_dlas, q0, RL.p_no_dla, nb);
  put(o->MAP_z_dlas, q0 * md * md, RL.map_z, nb * md * md);
  put(o->MAP_log_nhis, q0 * md * md, RL.map_lognhi, nb * md * md);
  if (o->sample_log_likelihoods_dla)
    memcpy(o->sample_log_likelihoods_dla + q0 * S * md, sl.h_sample_dla.p, sizeof(double) * nb * S * md);
  if (o->sample_log_likelihoods_lls)
    memcpy(o->sample_log_likelihoods_lls + q0 * S, sl.h_sample_sub.p, sizeof(double) * nb * S);
  if (o->base_sample_inds && md > 1)
    memcpy(o->base_sample_inds + q0 * S * (md - 1), sl.h_inds.p, sizeof(int32_t) * nb * S * (md - 1));
  for (size_t b = 0; b < nb; ++b) {
    if (o->min_z_dlas) o->min_z_dlas[q0 + b] = sl.zmin_b[b];
    if (o->max_z_dlas) o->max_z_dlas[q0 + b] = sl.zmax_b[b];
    if (o->num_pixels) o->num_pixels[q0 + b] = sl.n_b[b];
    if (o->status) o->status[q0 + b] = sl.h_alive.p[b * CAT_ALIVE + 1];
  }
  // algorithmic work of the likelihood launches: the evaluations that were run (masked samples are not evaluated)
  for (size_t b = 0; b < nb; ++b) {
    if (!sl.usable_b[b]) continue;
    const double per_eval = 472.0 * sl.n_b[b] + 3.1e3;
    long long evals = 2 * (long long)S + 1;
    for (size_t level = 1; level < md; ++level) {
      const int kept = sl.h_alive.p[b * CAT_ALIVE + 4 + level];
      evals += kept;
      // a spectrum that left the level loop (NaN evidence) has kept == 0 and nothing masked
      if (kept > 0 || sl.h_alive.p[b * CAT_ALIVE] != 0) cat->lk_masked += (long long)S - kept;
    }
    cat->lk_evaluated += evals;
    cat->gram_flops += (double)evals * per_eval;
  }
  float ms = 0.f;
  DLA_CUDA(cudaEventElapsedTime(&ms, sl.ev_v0, sl.ev_v1));
  cat->voigt_ms += ms;
  for (size_t level = 0; level < md; ++level) {
    DLA_CUDA(cudaEventElapsedTime(&ms, sl.ev_lk[2 * level], sl.ev_lk[2 * level + 1]));
    cat->gram_ms += ms;
  }
  return 0;
}

static int cat_run_pipeline(dla_catalogue* cat, const CatSource& src, const double* d_log_priors_in, dla_catalogue_outputs* o,
                            int* open_ranges) {
  Runtime& rt = runtime();
  const size_t cap = src.max_n_raw;
  int rc = cat_ensure_workspace(cat, cap, !src.on_device, o);
  if (rc) return rc;
  cat->total_ms = cat->gram_ms = cat->voigt_ms = cat->gram_flops = 0;
  cat->lk_evaluated = cat->lk_masked = 0;
  const long long launches_before = rt.launches;
  const PrepParams P = to_prep_params(&cat->params, 1);
  const int nbatches = (src.Q + cat->B - 1) / cat->B;
  for (CatSlot& sl : cat->slot) sl.prep_recorded = false;
  DLA_CUDA(cudaEventRecord(cat->ev_call0, rt.stream));
  // the copy stream must not run ahead of work queued before this call (log_priors upload, previous run)
  DLA_CUDA(cudaStreamWaitEvent(rt.copy_stream, cat->ev_call0, 0));
  nvtxRangePushA("dla_catalogue_run");
  ++*open_ranges;
  for (int bi = 0; bi < std::min(2, nbatches); ++bi) {
    if (!src.on_device && (rc = cat_enqueue_upload(cat, src, bi))) return rc;
    if ((rc = cat_enqueue_prep(cat, src, bi, P, cap))) return rc;
  }
  for (int bi = 0; bi < nbatches; ++bi) {
    char label[64];
    snprintf(label, sizeof(label), "batch %d (spectra %d..%d)", bi, bi * cat->B, std::min(src.Q, (bi + 1) * cat->B) - 1);
    nvtxRangePushA(label);
    ++*open_ranges;
    if ((rc = cat_enqueue_compute(cat, bi, cap, o, d_log_priors_in, open_ranges))) return rc;
    if (bi >= 1 && (rc = cat_consume(cat, bi - 1, o))) return rc;
    if (bi + 2 < nbatches) {
      if (!src.on_device && (rc = cat_enqueue_upload(cat, src, bi + 2))) return rc;
      if ((rc = cat_enqueue_prep(cat, src, bi + 2, P, cap))) return rc;
    }
    nvtxRangePop();
    --*open_ranges;
  }
  if ((rc = cat_consume(cat, nbatches - 1, o))) return rc;
  nvtxRangePop();
  --*open_ranges;
  // whole call on the library stream: kernels, descriptor uploads, result copies and host gaps
  DLA_CUDA(cudaEventRecord(cat->ev_call1, rt.stream));
  DLA_CUDA(cudaEventSynchronize(cat->ev_call1));
  {
    float ms = 0.f;
    DLA_CUDA(cudaEventElapsedTime(&ms, cat->ev_call0, cat->ev_call1));
    cat->total_ms = ms;
  }
  cat->launches = rt.launches - launches_before;
  rt.last_kernel_ms = cat->total_ms;
  return 0;
}

static int cat_run(dla_catalogue* cat, const CatSource& src, const double* d_log_priors_in, dla_catalogue_outputs* o) {
  int open_ranges = 0;
  const int rc = cat_run_pipeline(cat, src, d_log_priors_in, o, &open_ranges);
  if (rc) {
    // an error left batches in flight: nothing may still read the caller's host buffers or write the staging slots
    // when this call returns (the error text of the first failure is kept)
    Runtime& rt = runtime();
    cudaStreamSynchronize(rt.copy_stream);
    cudaStreamSynchronize(rt.stream);
    cudaGetLastError();
    while (open_ranges-- > 0) nvtxRangePop();
  }
  return rc;
}

extern "C" int dla_catalogue_run_staged(dla_catalogue* cat, dla_catalogue_outputs* o) {
  DLA_CHECK_READY();
  DLA_REQUIRE(cat && o, "null pointer argument");
  DLA_REQUIRE(cat->Q >= 1, "nothing staged");
  DLA_REQUIRE(cat->device == runtime().device, "the catalogue lives on another device than the one selected by dla_init");
  CatSource src;
  src.on_device = true;
  src.pix_off = cat->pix_off.data();
  src.wl = cat->wl.p;
  src.flux = cat->flux.p;
  src.var = cat->var.p;
  src.mask = cat->mask.p;
  src.z_qsos = cat->z_qsos.data();
  src.Q = cat->Q;
  src.max_n_raw = cat->max_n_raw;
  return cat_run(cat, src, cat->log_priors_in.p, o);
}

extern "C" int dla_catalogue_process(dla_catalogue* cat, int num_spectra, const int64_t* pixel_offsets,
                                     const double* wavelengths, const double* flux, const double* noise_variance,
                                     const uint8_t* pixel_mask, const double* z_qsos, const double* log_priors_in,
                                     dla_catalogue_outputs* outputs) {
  DLA_CHECK_READY();
  Runtime& rt = runtime();
  DLA_REQUIRE(cat && pixel_offsets && wavelengths && flux && noise_variance && pixel_mask && z_qsos && log_priors_in && outputs,
              "null pointer argument");
  DLA_REQUIRE(cat->device == rt.device, "the catalogue lives on another device than the one selected by dla_init");
  CatSource src;
  if (int rc = cat_check_offsets(num_spectra, pixel_offsets, &src.max_n_raw)) return rc;
  src.on_device = false;
  src.pix_off = pixel_offsets;
  src.wl = wavelengths;
  src.flux = flux;
  src.var = noise_variance;
  src.mask = pixel_mask;
  src.z_qsos = z_qsos;
  src.Q = num_spectra;
  const int m = 2 + cat->max_dlas;
  // the host-side priors of this call; a staged catalogue keeps its own copy
  DLA_CUDA(cat->log_priors_call.ensure((size_t)num_spectra * m));
  DLA_CUDA(cat->log_priors_call.upload(log_priors_in, (size_t)num_spectra * m, rt.stream));
  return cat_run(cat, src, cat->log_priors_call.p, outputs);
}

extern "C" int dla_catalogue_last_counts(const dla_catalogue* cat, long long* evaluated, long long* masked) {
  DLA_REQUIRE(cat, "null catalogue");
  if (evaluated) *evaluated = cat->lk_evaluated;
  if (masked) *masked = cat->lk_masked;
  return 0;
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
