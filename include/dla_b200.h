/*
 * dla_b200.h : C-ABI of libdla_b200.so - the B200 (sm_100a) implementation of the
 * per-spectrum Bayesian model-selection hot path of gpy_dla_detection.
 *
 * The reference is pure Python (NumPy/SciPy); its "FFI" for this path is the set of Python
 * methods listed beside each entry point below (paths relative to the reference tree).
 * A maintainer binds these functions with ctypes; the package
 * gpy_dla_detection_b200/ is exactly that binding.
 *
 * Conventions
 *   - every array is a caller-owned, C-contiguous host buffer (float64 / int32 / uint8);
 *     the library borrows the pointer for the duration of the call and writes results into
 *     caller-preallocated outputs.  Device memory is owned by opaque handles.
 *   - return value: 0 = ok, non-zero = error (dla_last_error() gives the text).  NaNs are
 *     in-band results exactly where the reference produces them (dla_gp.py:200-206).
 *   - no CPU fallback: every call fails if no CUDA device is usable.
 *   - threading: the library keeps ONE runtime per process (device, stream, copy stream): one host thread per
 *     process, one process per GPU (torchrun).  Calls are not re-entrant; calls on one handle are stream-ordered.
 *     Every handle records the device it was created on and every call checks it against the device selected by
 *     dla_init, so a handle used after dla_init(another device) fails instead of touching foreign memory.
 */
#ifndef DLA_B200_H
#define DLA_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct dla_model dla_model;       /* learned GP model resident on the device   */
typedef struct dla_spectrum dla_spectrum; /* one prepared spectrum + its profile cache */
typedef struct dla_catalogue dla_catalogue; /* batched catalogue engine                */

/* ---- library / device ------------------------------------------------------------- */
int dla_init(int device);                 /* select the CUDA device for this thread/process */
int dla_device_count(void);
const char* dla_last_error(void);
const char* dla_version(void);
/* milliseconds the GPU spent in the kernels of the last API call (CUDA events on the
 * library's own stream), and how many kernels that call launched */
double dla_last_kernel_ms(void);
long long dla_kernel_launch_count(void);  /* cumulative since dla_init */

/* FP64 peaks of this device, measured now (about 0.3 s): DFMA vector pipe and DMMA m8n8k4
 * tensor path, in TFLOP/s.  bench.py uses them as the roofline denominators because
 * MEASURED_PEAKS.json carries no FP64 figure. */
int dla_measure_fp64_peaks(double* dfma_tflops, double* dmma_tflops);

/* ---- a1: voigt.voigt_absorption (voigt.py:251-322; voigt.c:253-304) ---------------- */
/* out has n_in - 6 entries when broadening != 0, else n_in */
int dla_voigt_absorption(const double* wavelengths, int n_in, double nhi, double z_dla,
                         int num_lines, int broadening, double* out);
/* S profiles on one wavelength grid; out is (S, n_out) row-major */
int dla_voigt_absorption_batch(const double* wavelengths, int n_in, const double* nhis,
                               const double* z_dlas, int S, int num_lines, int broadening,
                               double* out);
/* voigt_lls.voigt_absorption (voigt_lls.py:287-363): the same profile with the Lyman-limit break optical
 * depth tau = nhi / 10^17.2 (lambda_rest / 911.7641)^3 (lambda_rest <= 911.7641 A) added to the exponent */
int dla_voigt_lls_absorption_batch(const double* wavelengths, int n_in, const double* nhis,
                                   const double* z_llss, int S, int num_lines, int broadening,
                                   double* out);
/* Re w(x + i y) of the profile kernel, exposed for accuracy tests (0 <= y <= 1e-3) */
int dla_faddeeva_re(const double* x, const double* y, int n, double* out);

/* ---- a2: effective_optical_depth.effective_optical_depth (:10-80) ------------------ */
/* out is (n, num_forest_lines) row-major */
int dla_effective_optical_depth(const double* wavelengths, int n, double beta, double tau_0,
                                double z_qso, int num_forest_lines, double* out);

/* ---- a5: NullGP.log_mvnpdf_low_rank (null_gp.py:307-360) --------------------------- */
/* M is (n, k) row-major, any k <= 64 */
int dla_log_mvnpdf_low_rank(const double* y, const double* mu, const double* M,
                            const double* d, int n, int k, double* out);

/* ---- a4: learned model (NullGP.__init__, null_gp.py:36-71) ------------------------- */
int dla_model_create(const double* rest_wavelengths, const double* mu, const double* M,
                     const double* log_omega, int n_rest, int k, double log_c_0,
                     double log_tau_0, double log_beta, double prev_tau_0, double prev_beta,
                     dla_model** out);
int dla_model_destroy(dla_model* model);

/* NullGP.get_interp (null_gp.py:179-242) for caller-supplied pixels: x = rest wavelengths (must lie inside the
 * model grid, as scipy interp1d requires), wavelengths = observed, both (n).  Outputs this_mu (n), this_M (n, k)
 * row-major, this_omega2 (n): interpolated model with the Kim et al. mean-flux suppression applied. */
int dla_model_interp(const dla_model* model, int num_forest_lines, const double* x,
                     const double* wavelengths, int n, double z_qso, double* this_mu, double* this_M,
                     double* this_omega2);

/* pipeline parameters read by the path (set_parameters.py:21-102) */
typedef struct dla_params {
  double min_lambda, max_lambda;                             /* modelling range, rest A  */
  double normalization_min_lambda, normalization_max_lambda; /* flux normalisation window */
  double pixel_spacing;                                      /* dex                      */
  int width;                                                 /* instrument half-width    */
  int num_forest_lines;                                      /* mean-flux suppression    */
  int num_lines;                                             /* absorber Lyman members   */
  int broadening;                                            /* instrumental broadening  */
  double lya_wavelength, lyman_limit;                        /* A                        */
  double max_z_cut, min_z_cut;                               /* already as redshift      */
  double min_z_separation;                                   /* already as redshift      */
} dla_params;

/* ---- a3/a4: NullGP.set_data + get_interp on the device (null_gp.py:95-242) --------- */
/* X = rest wavelengths (observed / (1+z_qso)), Y flux, V noise variance, mask uint8 */
int dla_spectrum_create(const dla_model* model, const dla_params* params, const double* X,
                        const double* Y, const double* V, const uint8_t* pixel_mask,
                        int n_raw, double z_qso, int normalize, dla_spectrum** out);
/* user-supplied interpolated model (for subclasses that override get_interp / set_data):
 * y, v, mu, omega2 (n), M (n,k) of the modelled pixels; wl_abs = padded (n_u+6) or unmasked
 * (n_u) wavelengths; keep (n_u) uint8 = ~pixel_mask[ind_unmasked] */
int dla_spectrum_create_prepared(const double* y, const double* v, const double* mu,
                                 const double* M, const double* omega2, int n, int k,
                                 const double* wl_abs, int n_abs, const uint8_t* keep, int n_u,
                                 int broadening, dla_spectrum** out);
int dla_spectrum_destroy(dla_spectrum* spec);
/* absorbers of this spectrum use the Lyman-limit-system profile of voigt_lls.py in every later call
 * (the extension pattern of examples/gp_find_lls.py:159-224: a DLAGP subclass overriding this_dla_gp) */
int dla_spectrum_set_lls_break(dla_spectrum* spec, int on);
/* sizes: n_raw, n_u (in range), n (in range & unmasked) */
int dla_spectrum_sizes(const dla_spectrum* spec, int* n_raw, int* n_u, int* n);
/* copy-back of the attributes NullGP holds after set_data (any pointer may be NULL):
 * x,y,v,this_wavelengths,this_mu,this_omega2 (n); this_M (n,k); unmasked_wavelengths (n_u);
 * padded_wavelengths (n_u+2*width); ind_unmasked, ind (n_raw, uint8); normalization_median */
int dla_spectrum_get(const dla_spectrum* spec, double* x, double* y, double* v,
                     double* this_wavelengths, double* this_mu, double* this_M,
                     double* this_omega2, double* unmasked_wavelengths,
                     double* padded_wavelengths, uint8_t* ind_unmasked, uint8_t* ind,
                     double* normalization_median);

/* ---- a6: NullGP.log_model_evidence (null_gp.py:294-305) ---------------------------- */
int dla_null_log_model_evidence(dla_spectrum* spec, double* out);

/* ---- a7/a8: DLAGP.sample_log_likelihood_k_dlas for S parameter sets ----------------- */
/* z_dlas, nhis are (S, k_dlas) row-major; out (S); product order = column order */
int dla_sample_log_likelihoods(dla_spectrum* spec, const double* z_dlas, const double* nhis,
                               int S, int k_dlas, int num_lines, double* out);
/* a7: DLAGP.this_dla_gp absorption for one parameter set: out (n) masked absorption */
int dla_absorption_k_dlas(dla_spectrum* spec, const double* z_dlas, const double* nhis,
                          int k_dlas, int num_lines, double* out);

/* ---- a9: DLAGP/SubDLAGP.log_model_evidences (dla_gp.py:92-225, subdla_gp.py:90-222) - */
/* z_samples, nhi_samples (S); uniforms ((max_dlas-1), S) = the MT19937 draws the reference
 * takes through np.random.choice; outputs: sample_log_likelihoods (S, max_dlas) NaN-filled,
 * base_sample_inds ((max_dlas-1), S) int32, log_evidences (max_dlas),
 * uniform_rows_used = number of resampling steps actually performed */
int dla_log_model_evidences(dla_spectrum* spec, const double* z_samples,
                            const double* nhi_samples, int S, int max_dlas,
                            const double* uniforms, double min_z_separation, int num_lines,
                            double* sample_log_likelihoods, int32_t* base_sample_inds,
                            double* log_evidences, int* uniform_rows_used);
/* resampling step alone (np.random.choice with p = W / W.sum(), dla_gp.py:209-218) */
int dla_resample_indices(const double* W, const double* uniforms, int S, int32_t* out);

/* ---- a15: run_bayes_select.process_qso, batched (run_bayes_select.py:141-230) ------- */
typedef struct dla_catalogue_config {
  int num_dla_samples;   /* S */
  int max_dlas;
  int batch_spectra;     /* spectra resident per batch (0 = choose from free memory) */
  int keep_sample_likelihoods; /* write the (Q,S,max_dlas) / (Q,S) sample arrays */
} dla_catalogue_config;

/* The catalogue BORROWS `model`: it must stay alive until dla_catalogue_destroy (the Python binding keeps both in one
 * object).  Sample arrays and uniforms are copied to the device by this call. */
int dla_catalogue_create(const dla_model* model, const dla_params* params,
                         const dla_catalogue_config* config,
                         const double* dla_offset_samples, const double* dla_log_nhi_samples,
                         const double* dla_nhi_samples, const double* sub_offset_samples,
                         const double* sub_nhi_samples, const double* uniforms,
                         dla_catalogue** out);
int dla_catalogue_destroy(dla_catalogue* cat);
/* spectra are ragged: pixel_offsets (Q+1) index into wavelengths/flux/noise_variance/mask.
 * log_priors (Q, 2+max_dlas) = [unused, subDLA, DLA 1..max] from the host prior catalogue
 * (entry 0 is filled with log(1 - sum others), bayesian_model_selection.py:79-80).
 * Outputs (any may be NULL): per run_bayes_select.py:107-139. */
typedef struct dla_catalogue_outputs {
  double* min_z_dlas;                 /* Q */
  double* max_z_dlas;                 /* Q */
  double* log_priors;                 /* Q x (2+max) : completed priors */
  double* log_likelihoods;            /* Q x (2+max) : [null, sub, dla 1..max] */
  double* log_posteriors;             /* Q x (2+max) */
  double* model_posteriors;           /* Q x (2+max) */
  double* p_dlas;                     /* Q */
  double* p_no_dlas;                  /* Q */
  double* MAP_z_dlas;                 /* Q x max x max */
  double* MAP_log_nhis;               /* Q x max x max */
  double* sample_log_likelihoods_dla; /* Q x S x max   (optional) */
  double* sample_log_likelihoods_lls; /* Q x S         (optional) */
  int32_t* base_sample_inds;          /* Q x S x (max-1) (optional, transposed as :214) */
  int32_t* num_pixels;                /* Q : modelled pixels n per spectrum */
  int32_t* status;                    /* Q : 0 ok, 1 = no pixels in range, 2 = NaN early exit */
} dla_catalogue_outputs;
int dla_catalogue_process(dla_catalogue* cat, int num_spectra, const int64_t* pixel_offsets,
                          const double* wavelengths, const double* flux,
                          const double* noise_variance, const uint8_t* pixel_mask,
                          const double* z_qsos, const double* log_priors_in,
                          dla_catalogue_outputs* outputs);
/* same, inputs already staged on the device by dla_catalogue_stage (kernel-only timing) */
int dla_catalogue_stage(dla_catalogue* cat, int num_spectra, const int64_t* pixel_offsets,
                        const double* wavelengths, const double* flux,
                        const double* noise_variance, const uint8_t* pixel_mask,
                        const double* z_qsos, const double* log_priors_in);
int dla_catalogue_run_staged(dla_catalogue* cat, dla_catalogue_outputs* outputs);
/* timing of the last process/run call: total GPU ms and the share of the likelihood kernel */
int dla_catalogue_last_timing(const dla_catalogue* cat, double* total_ms, double* gram_ms,
                              double* voigt_ms, long long* launches, double* gram_flops);

/* likelihood evaluations of the last process/run call: how many were run, and how many samples of levels >= 1 were
 * left out because the separation test (dla_gp.py:164-177) turns their result into NaN whatever its value */
int dla_catalogue_last_counts(const dla_catalogue* cat, long long* evaluated, long long* masked);

/* ---- a14: quasar-redshift estimation, ZGP (zqso_gp.py) ---------------------------------- */
typedef struct dla_zqso_model dla_zqso_model; /* learned zQSO model resident on the device */
/* ZGP.__init__ (zqso_gp.py:36-64): rest grid (n_rest), mu (n_rest), M (n_rest, k = 20) row-major, and the
 * means / standard deviations of the i.i.d. models outside the modelling window */
int dla_zqso_model_create(const double* rest_wavelengths, const double* mu, const double* M, int n_rest, int k,
                          double bluewards_mu, double redwards_mu, double bluewards_sigma,
                          double redwards_sigma, dla_zqso_model** out);
int dla_zqso_model_destroy(dla_zqso_model* model);
/* the attributes of ZParameters the path reads (zqso_set_parameters.py:19-54) */
typedef struct dla_zqso_params {
  double min_lambda, max_lambda;                             /* modelling window, rest A */
  double normalization_min_lambda, normalization_max_lambda; /* flux normalisation window, rest A */
} dla_zqso_params;
/* ZGP.inference_z_qso (zqso_gp.py:214-250) for num_spectra ragged spectra (observed wavelengths,
 * strictly increasing per spectrum) and S candidate redshifts shared by all spectra.
 * sample_log_likelihoods (num_spectra, S) may be NULL; z_map (num_spectra) = z of np.nanargmax, NaN and
 * map_index -1 when every sample is NaN (the reference raises ValueError there). */
int dla_zqso_inference(const dla_zqso_model* model, const dla_zqso_params* params, int num_spectra,
                       const int64_t* pixel_offsets, const double* wavelengths, const double* flux,
                       const double* noise_variance, const uint8_t* pixel_mask, const double* z_samples,
                       int S, double* sample_log_likelihoods, double* z_map, int32_t* map_index);
/* timing of the last dla_zqso_inference call (CUDA events on the library stream): the kernels alone, i.e. with the
 * spectra already resident in HBM, and the whole call including the host <-> device copies */
int dla_zqso_last_timing(double* kernel_ms, double* total_ms);
/* Testing hook: dla_zqso_inference picks its kernel by the model grid (uniform spacing: median pass +
 * zqso_likelihood_kernel_v2; otherwise the generic kernel).  on != 0 forces the generic kernel for every model so
 * that both can be checked on the same inputs. */
int dla_zqso_force_generic_kernel(int on);
/* ZGP.set_data + get_interp at one redshift (zqso_gp.py:66-182), element-wise over the n_raw pixels:
 * x = X / (1 + z), normalised flux / variance, interpolated mu (n_raw) and M (n_raw, k) where cls == 1,
 * cls (0 none, 1 modelled, 2 bluewards, 3 redwards), in_window (the first `ind` of :132), this_median.
 * The caller selects rows by cls (boolean indexing) to form the reference's attributes. */
int dla_zqso_set_data(const dla_zqso_model* model, const dla_zqso_params* params, const double* X,
                      const double* Y, const double* noise_variance, const uint8_t* pixel_mask, int n_raw,
                      double z_qso, double* x, double* y_normalized, double* v_normalized, double* this_mu,
                      double* this_M, uint8_t* cls, uint8_t* in_window, double* this_median);
/* ZGP.log_mvnpdf_iid (zqso_gp.py:252-278) */
int dla_log_mvnpdf_iid(const double* y, const double* mu, const double* d, int n, double* out);

#ifdef __cplusplus
}
#endif
#endif /* DLA_B200_H */
