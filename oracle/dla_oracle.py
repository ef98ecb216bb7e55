"""
dla_oracle.py : CPU restatement (NumPy/SciPy float64) of the reference's per-spectrum
Bayesian model-selection path.  TEST INFRASTRUCTURE ONLY.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline legs may import this
module, and only as the checker / reported baseline; nothing under `gpy_dla_detection_b200/`
imports it.  The product path calls hand-written CUDA through the C-ABI and raises when the
library is missing.

Parity status: PINNED.  `tests/golden/make_golden.py` runs the live reference
(`/root/reference`, imported unmodified with `h5py`/`emcee` stubbed) on seeded synthetic
inputs and stores its outputs under `tests/golden/`; `tests/test_oracle.py` checks this
restatement against those vectors and against the reference's own data-free known-answer
tests (tests/test_model.py:52-72, tests/test_voigt.py:8-57).

Third-party arithmetic: the Faddeeva function is `scipy.special.wofz` (scipy 1.18.1 here;
S. G. Johnson's Faddeeva package), exactly the call the reference makes (voigt.py:15,248).

Every function cites the reference lines it restates (paths relative to /root/reference).
The restatement is vectorised over QMC samples where the reference loops in Python; the
per-sample arithmetic (operation order inside one sample) follows the reference.
"""
from typing import Dict, Optional, Tuple

import numpy as np
from scipy.special import wofz, logsumexp

# ---- constants: gpy_dla_detection/voigt.py:18-224 (values checked bit-identical in tests) ----
from gpy_dla_detection_b200 import _tables as T  # data only (no compute code)

LOG_2PI = 1.83787706640934534  # null_gp.py:325
LYA_WAVELENGTH = 1215.6701  # set_parameters.py:16
LYMAN_LIMIT = 911.7633  # set_parameters.py:18
SPEED_OF_LIGHT_MS = 299792458.0  # set_parameters.py:19


def kms_to_z(kms: float) -> float:
    """set_parameters.py:104-109"""
    return (kms * 1000) / SPEED_OF_LIGHT_MS


# ------------------------------------------------------------------------------------------
# a1  Voigt absorption profile
# ------------------------------------------------------------------------------------------
def voigt_profile(x: np.ndarray, sigma: float, gamma: float) -> np.ndarray:
    """voigt.py:241-248 : Re w((x + i gamma)/(sqrt(2) sigma)) / (sqrt(2 pi) sigma)"""
    z = (x + 1j * gamma) / (np.sqrt(2) * sigma)
    return np.real(wofz(z)) / (np.sqrt(2 * np.pi) * sigma)


def voigt_absorption(
    wavelengths: np.ndarray, nhi: float, z_dla: float, num_lines: int = 3, broadening: bool = True
) -> np.ndarray:
    """
    voigt.py:251-322.  `wavelengths` observed Angstrom; returns len-6 points when
    `broadening` (7-tap 'valid' convolution with the instrument profile), else len points.
    """
    c = T.SPEED_OF_LIGHT_CGS
    multipliers = c / (T.TRANSITION_WAVELENGTHS[:num_lines] * (1 + z_dla)) / 1e8  # voigt.py:296
    total = np.empty((num_lines, wavelengths.shape[0]))
    for l in range(num_lines):  # voigt.py:301-305
        velocity = wavelengths * multipliers[l] - c
        total[l, :] = -T.LEADING_CONSTANTS[l] * voigt_profile(velocity, T.SIGMA, T.GAMMAS[l])
    raw_profile = np.exp(np.float64(nhi) * np.nansum(total, axis=0))  # voigt.py:307
    if broadening:
        return np.convolve(raw_profile, T.INSTRUMENT_PROFILE, "valid")  # voigt.py:318
    return raw_profile


def voigt_absorption_batch(
    wavelengths: np.ndarray,
    nhis: np.ndarray,
    z_dlas: np.ndarray,
    num_lines: int = 3,
    broadening: bool = True,
    chunk: int = 512,
) -> np.ndarray:
    """
    voigt.py:251-322 for many (nhi, z_dla) pairs at once -> (S, n_out).  Same elementwise
    arithmetic as `voigt_absorption`; the sum over lines is accumulated in line order with
    NaN terms counted as zero (np.nansum semantics, voigt.py:307).
    """
    nhis = np.asarray(nhis, dtype=np.float64)
    z_dlas = np.asarray(z_dlas, dtype=np.float64)
    S = nhis.shape[0]
    n_in = wavelengths.shape[0]
    n_out = n_in - 2 * T.WIDTH if broadening else n_in
    out = np.empty((S, n_out))
    c = T.SPEED_OF_LIGHT_CGS
    inv_norm = np.sqrt(2 * np.pi) * T.SIGMA
    denom = np.sqrt(2) * T.SIGMA
    for s0 in range(0, S, chunk):
        s1 = min(S, s0 + chunk)
        total = np.zeros((s1 - s0, n_in))
        for l in range(num_lines):
            mult = c / (T.TRANSITION_WAVELENGTHS[l] * (1 + z_dlas[s0:s1])) / 1e8
            velocity = wavelengths[None, :] * mult[:, None] - c
            zz = (velocity + 1j * T.GAMMAS[l]) / denom
            term = -T.LEADING_CONSTANTS[l] * (np.real(wofz(zz)) / inv_norm)
            total += np.where(np.isnan(term), 0.0, term)
        raw = np.exp(nhis[s0:s1, None] * total)
        if broadening:
            # np.convolve(raw, profile, 'valid')[i] = sum_k raw[i+k] * profile[6-k] (symmetric)
            acc = np.zeros((s1 - s0, n_out))
            for k in range(2 * T.WIDTH + 1):
                acc += raw[:, k : k + n_out] * T.INSTRUMENT_PROFILE[2 * T.WIDTH - k]
            out[s0:s1] = acc
        else:
            out[s0:s1] = raw
    return out


# ------------------------------------------------------------------------------------------
# a2  effective optical depth of the Lyman-series forest
# ------------------------------------------------------------------------------------------
def effective_optical_depth(
    wavelengths: np.ndarray, beta: float, tau_0: float, z_qso: float, num_forest_lines: int
) -> np.ndarray:
    """
    effective_optical_depth.py:10-80 -> (n, num_forest_lines).  The indicator
    z_absorber <= z_qso multiplies every member including Ly-alpha (the
    `skip_lya_indicator` argument of the reference is never read).
    """
    tw = T.TRANSITION_WAVELENGTHS * 1e8
    f = T.OSCILLATOR_STRENGTHS
    out = np.empty((wavelengths.shape[0], num_forest_lines))
    for i in range(num_forest_lines):
        z_i = (wavelengths - tw[i]) / tw[i]
        this_tau_0 = tau_0 * f[i] / f[0] * tw[i] / tw[0]
        out[:, i] = this_tau_0 * (1 + z_i) ** beta
        out[:, i] = out[:, i] * (z_i <= z_qso)
    return out


# ------------------------------------------------------------------------------------------
# a3/a4  spectrum preparation (NullGP.set_data / get_interp)
# ------------------------------------------------------------------------------------------
def prepare_spectrum(
    model: Dict[str, np.ndarray],
    rest_wavelengths_obs: np.ndarray,
    flux: np.ndarray,
    noise_variance: np.ndarray,
    pixel_mask: np.ndarray,
    z_qso: float,
    min_lambda: float = 911.75,
    max_lambda: float = 1215.75,
    normalization_min_lambda: float = 1310.0,
    normalization_max_lambda: float = 1325.0,
    num_forest_lines: int = 31,
    width: int = 3,
    pixel_spacing: float = 1e-4,
    prev_tau_0: float = 0.0023,
    prev_beta: float = 3.65,
    normalize: bool = True,
) -> Dict[str, np.ndarray]:
    """
    null_gp.py:95-242.  `rest_wavelengths_obs` is X = observed wavelengths / (1 + z_qso).
    Returns the attributes the reference object holds after `set_data(..., build_model=True)`.
    """
    x = rest_wavelengths_obs
    y = flux
    v = noise_variance
    out = {}
    if normalize:  # null_gp.py:125-136
        ind = (x >= normalization_min_lambda) & (x <= normalization_max_lambda) & (~pixel_mask)
        this_median = np.nanmedian(y[ind])
        y = y / this_median
        v = v / this_median**2
        out["normalization_median"] = this_median

    ind_unmasked = (x >= min_lambda) & (x <= max_lambda)  # null_gp.py:139-140
    observed = x * (1 + z_qso)
    unmasked_wavelengths = observed[ind_unmasked]  # null_gp.py:144
    ind = ind_unmasked & (~pixel_mask)  # null_gp.py:146
    this_wavelengths = observed[ind]
    xs, ys, vs = x[ind], y[ind], v[ind]

    # get_interp, null_gp.py:179-242 (interp1d linear on float64 1-D data == np.interp)
    rw = model["rest_wavelengths"]
    this_mu = np.interp(xs, rw, model["mu"])
    this_M = np.stack([np.interp(xs, rw, col) for col in model["M"].T], axis=1)
    this_omega2 = np.exp(2 * np.interp(xs, rw, model["log_omega"]))

    tod = effective_optical_depth(this_wavelengths, prev_beta, prev_tau_0, z_qso, num_forest_lines)
    lya_absorption = np.exp(-np.sum(tod, axis=1))
    this_mu = this_mu * lya_absorption
    this_M = this_M * lya_absorption[:, None]

    lod = effective_optical_depth(
        this_wavelengths, np.exp(model["log_beta"]), np.exp(model["log_tau_0"]), z_qso, num_forest_lines
    )
    scaling = 1 - np.exp(-np.sum(lod, axis=1)) + np.exp(model["log_c_0"])
    this_omega2 = this_omega2 * scaling**2
    this_omega2 = this_omega2 * lya_absorption**2

    # padded grid, null_gp.py:159-177
    lo = np.log10(unmasked_wavelengths.min())
    hi = np.log10(unmasked_wavelengths.max())
    padded = np.concatenate(
        [
            np.logspace(lo - width * pixel_spacing, lo - pixel_spacing, width),
            unmasked_wavelengths,
            np.logspace(hi + pixel_spacing, hi + width * pixel_spacing, width),
        ]
    )
    out.update(
        x=xs,
        y=ys,
        v=vs,
        z_qso=z_qso,
        pixel_mask=pixel_mask,
        ind_unmasked=ind_unmasked,
        ind=ind,
        unmasked_wavelengths=unmasked_wavelengths,
        this_wavelengths=this_wavelengths,
        padded_wavelengths=padded,
        this_mu=this_mu,
        this_M=this_M,
        this_omega2=this_omega2,
        mask_ind=~pixel_mask[ind_unmasked],  # dla_gp.py:360
    )
    return out


# ------------------------------------------------------------------------------------------
# a5  low-rank Gaussian log-density
# ------------------------------------------------------------------------------------------
def log_mvnpdf_low_rank(y: np.ndarray, mu: np.ndarray, M: np.ndarray, d: np.ndarray) -> float:
    """
    null_gp.py:307-360 : log N(y; mu, M M' + diag d) through the Woodbury identity and a
    k x k Cholesky factor.  The reference forms C = B^-1 M' D^-1 (k x n) with two dtrtri
    calls; here the same quantity is reached with triangular solves on the k-vector
    M' D^-1 r, which is algebraically identical.
    """
    n, k = M.shape
    r = y - mu
    d_inv = 1 / d
    D_inv_M = d_inv[:, None] * M
    B = M.T @ D_inv_M
    B[np.diag_indices(k)] += 1
    L = np.linalg.cholesky(B)
    c = D_inv_M.T @ r
    zvec = np.linalg.solve(L, c)  # L z = c
    quad = np.dot(r, d_inv * r) - np.dot(zvec, zvec)
    log_det = np.sum(np.log(d)) + 2 * np.sum(np.log(np.diag(L)))
    return -0.5 * (quad + log_det + n * LOG_2PI)


def log_mvnpdf_low_rank_batch(y: np.ndarray, mu: np.ndarray, M: np.ndarray, d: np.ndarray) -> np.ndarray:
    """
    null_gp.py:307-360 for a batch, literal form: mu (S, n), M (S, n, k), d (S, n) -> (S,).
    Slow (einsum); kept as the literal restatement that `batch_log_likelihoods` is tested against.
    """
    S, n = mu.shape
    k = M.shape[-1]
    r = y[None, :] - mu
    d_inv = 1 / d
    D_inv_M = d_inv[:, :, None] * M
    B = np.einsum("sni,snj->sij", M, D_inv_M)
    B[:, np.arange(k), np.arange(k)] += 1
    L = np.linalg.cholesky(B)
    c = np.einsum("sni,sn->si", D_inv_M, r)
    zvec = np.linalg.solve(L, c[:, :, None])[:, :, 0]
    quad = np.sum(r * d_inv * r, axis=1) - np.sum(zvec * zvec, axis=1)
    log_det = np.sum(np.log(d), axis=1) + 2 * np.sum(np.log(np.diagonal(L, axis1=1, axis2=2)), axis=1)
    return -0.5 * (quad + log_det + n * LOG_2PI)


# ------------------------------------------------------------------------------------------
# a7/a8  absorber model applied to the GP and the per-sample likelihood
# ------------------------------------------------------------------------------------------
def absorption_k_dlas(
    prep: Dict[str, np.ndarray], z_dlas: np.ndarray, nhis: np.ndarray, num_lines: int = 3, broadening: bool = True
) -> np.ndarray:
    """dla_gp.py:358-388 : product of the k profiles in the given order, then mask."""
    wl = prep["padded_wavelengths"] if broadening else prep["unmasked_wavelengths"]
    absorption = voigt_absorption(wl, nhis[0], z_dlas[0], num_lines, broadening)
    for j in range(1, len(z_dlas)):
        absorption = absorption * voigt_absorption(wl, nhis[j], z_dlas[j], num_lines, broadening)
    return absorption[prep["mask_ind"]]


def sample_log_likelihood_k_dlas(
    prep: Dict[str, np.ndarray], z_dlas: np.ndarray, nhis: np.ndarray, num_lines: int = 3, broadening: bool = True
) -> float:
    """dla_gp.py:311-329 + :331-396"""
    a = absorption_k_dlas(prep, z_dlas, nhis, num_lines, broadening)
    dla_mu = prep["this_mu"] * a
    dla_M = prep["this_M"] * a[:, None]
    dla_omega2 = prep["this_omega2"] * a**2
    return log_mvnpdf_low_rank(prep["y"], dla_mu, dla_M, dla_omega2 + prep["v"])


def null_log_model_evidence(prep: Dict[str, np.ndarray]) -> float:
    """null_gp.py:294-305"""
    return log_mvnpdf_low_rank(prep["y"], prep["this_mu"], prep["this_M"], prep["this_omega2"] + prep["v"])


def batch_log_likelihoods(prep: Dict[str, np.ndarray], absorption: np.ndarray, chunk: int = 256) -> np.ndarray:
    """
    dla_gp.py:392-394 + null_gp.py:307-360 for a (S, n) block of masked absorption rows a_s.
    With dla_mu = mu a, dla_M = M a, d = omega2 a^2 + v the reference's quantities are
      B_s = I + M' diag(a^2 / d) M ,   M_dla' D^-1 r = M' (a r / d) ,   r = y - mu a,
    so the batch is one dense contraction per chunk (BLAS) instead of S small ones.
    """
    y, mu0, M0, om, v = prep["y"], prep["this_mu"], prep["this_M"], prep["this_omega2"], prep["v"]
    n, k = M0.shape
    S = absorption.shape[0]
    out = np.empty(S)
    diag = np.arange(k)
    for s0 in range(0, S, chunk):
        a = absorption[s0 : s0 + chunk]
        d = om[None, :] * a**2 + v[None, :]
        r = y[None, :] - mu0[None, :] * a
        d_inv = 1 / d
        w = a * a * d_inv
        g = a * r * d_inv
        B = np.matmul(M0.T[None, :, :] * w[:, None, :], M0)
        B[:, diag, diag] += 1
        L = np.linalg.cholesky(B)
        c = g @ M0
        zvec = np.linalg.solve(L, c[:, :, None])[:, :, 0]
        quad = np.sum(r * r * d_inv, axis=1) - np.sum(zvec * zvec, axis=1)
        log_det = np.sum(np.log(d), axis=1) + 2 * np.sum(np.log(np.diagonal(L, axis1=1, axis2=2)), axis=1)
        out[s0 : s0 + chunk] = -0.5 * (quad + log_det + n * LOG_2PI)
    return out


# ------------------------------------------------------------------------------------------
# a10  QMC z samples
# ------------------------------------------------------------------------------------------
def z_dla_range(this_wavelengths: np.ndarray, z_qso: float, min_lambda=911.75, max_lambda=1215.75,
                max_z_cut_kms=3000.0, min_z_cut_kms=3000.0) -> Tuple[float, float]:
    """set_parameters.py:125-159 -> (min_z_dla, max_z_dla)"""
    rest = this_wavelengths / (1 + z_qso)
    ind = (rest >= min_lambda) & (rest <= max_lambda)
    max_z = np.min([(np.max(this_wavelengths[ind]) / LYA_WAVELENGTH - 1) - kms_to_z(max_z_cut_kms),
                    z_qso - kms_to_z(max_z_cut_kms)])
    min_z = np.max([np.min(this_wavelengths[ind]) / LYA_WAVELENGTH - 1,
                    LYMAN_LIMIT * (1 + z_qso) / LYA_WAVELENGTH - 1 + kms_to_z(min_z_cut_kms)])
    return min_z, max_z


def sample_z_dlas(offset_samples: np.ndarray, this_wavelengths: np.ndarray, z_qso: float) -> np.ndarray:
    """dla_samples.py:94-104 / subdla_samples.py:115-125"""
    min_z, max_z = z_dla_range(this_wavelengths, z_qso)
    return min_z + (max_z - min_z) * offset_samples


# ------------------------------------------------------------------------------------------
# a9  evidence levels with conditional resampling
# ------------------------------------------------------------------------------------------
def resample_indices(W: np.ndarray, uniforms: np.ndarray) -> np.ndarray:
    """
    dla_gp.py:213-218 : np.random.choice(arange(S, int32), S, True, p) for p = W / W.sum()
    == searchsorted(cumsum(p) / cumsum(p)[-1], U, 'right') with U the next S draws of the
    legacy MT19937 stream (numpy/random/mtrand.pyx `choice`).
    """
    p = W / W.sum()
    cdf = p.cumsum()
    cdf /= cdf[-1]
    return cdf.searchsorted(uniforms, side="right").astype(np.int32)


def log_model_evidences(
    prep: Dict[str, np.ndarray],
    offset_samples: np.ndarray,
    nhi_samples: np.ndarray,
    max_dlas: int,
    uniforms: Optional[np.ndarray],
    min_z_separation_kms: float = 3000.0,
    num_lines: int = 3,
    broadening: bool = True,
) -> Dict[str, np.ndarray]:
    """
    dla_gp.py:92-225 (and its twin subdla_gp.py:90-222).  `uniforms` is the
    (max_dlas-1, S) block of MT19937 draws the reference would take from the global NumPy
    stream; rows are consumed only for levels that are actually resampled.
    Returns log evidences, sample_log_likelihoods (S, max_dlas), base_sample_inds
    (max_dlas-1, S) int32 and the number of uniform rows consumed.
    """
    S = offset_samples.shape[0]
    min_z_separation = kms_to_z(min_z_separation_kms)
    log_likelihoods_dla = np.full((max_dlas,), np.nan)
    base_sample_inds = np.zeros((max_dlas - 1, S), dtype=np.int32)
    sample_log_likelihoods = np.full((S, max_dlas), np.nan)
    z_samples = sample_z_dlas(offset_samples, prep["this_wavelengths"], prep["z_qso"])

    wl = prep["padded_wavelengths"] if broadening else prep["unmasked_wavelengths"]
    # unique single-absorber profiles; every multi-DLA absorption is a product of these rows
    profiles = voigt_absorption_batch(wl, nhi_samples, z_samples, num_lines, broadening)
    rows_used = 0
    for num_dlas in range(max_dlas):
        absorption = profiles.copy()
        for j in range(num_dlas):  # left-to-right product [i, b0, b1, ...], dla_gp.py:135-150
            absorption = absorption * profiles[base_sample_inds[j]]
        absorption = absorption[:, prep["mask_ind"]]
        ll = batch_log_likelihoods(prep, absorption) - np.log(S)  # dla_gp.py:155-159
        sample_log_likelihoods[:, num_dlas] = ll

        if num_dlas > 0:  # dla_gp.py:164-177
            all_z = np.concatenate([z_samples[None, :], z_samples[base_sample_inds[:num_dlas]]], axis=0)
            too_close = np.any(np.diff(np.sort(all_z, axis=0), axis=0) < min_z_separation, axis=0)
            sample_log_likelihoods[too_close, num_dlas] = np.nan

        col = sample_log_likelihoods[:, num_dlas]
        if np.all(np.isnan(col)):
            max_ll = np.nan
        else:
            max_ll = np.nanmax(col)
        probs = np.exp(col - max_ll)
        with np.errstate(invalid="ignore", divide="ignore"):
            log_likelihoods_dla[num_dlas] = (
                max_ll + np.log(np.nanmean(probs) if not np.all(np.isnan(probs)) else np.nan)
                - np.log(S) * num_dlas
            )  # dla_gp.py:180-190
        if (num_dlas + 1) == max_dlas:
            break
        if np.isnan(log_likelihoods_dla[num_dlas]):  # dla_gp.py:200-206
            break
        W = probs
        W[np.isnan(W)] = 0.0
        base_sample_inds[num_dlas, :] = resample_indices(W, uniforms[rows_used])
        rows_used += 1
    return dict(
        log_likelihoods=log_likelihoods_dla,
        sample_log_likelihoods=sample_log_likelihoods,
        base_sample_inds=base_sample_inds,
        sample_z_dlas=z_samples,
        uniform_rows_used=rows_used,
    )


# ------------------------------------------------------------------------------------------
# a11  model priors
# ------------------------------------------------------------------------------------------
def dla_log_priors(num_dlas_below: float, num_quasars_below: float, max_dlas: int, scale: float = 1.0) -> np.ndarray:
    """dla_gp.py:398-426 ; subdla_gp.py:311-346 with scale = Z_lls / Z_dla"""
    p = scale * (num_dlas_below / num_quasars_below) ** np.arange(1, max_dlas + 1)
    for i in range(max_dlas - 1):
        p[i] = p[i] - p[i + 1]
    return np.log(p)


# ------------------------------------------------------------------------------------------
# a12  Bayesian model selection
# ------------------------------------------------------------------------------------------
def model_selection(log_priors_sub: np.ndarray, log_priors_dla: np.ndarray, log_ev_null: float,
                    log_ev_sub: np.ndarray, log_ev_dla: np.ndarray) -> Dict[str, np.ndarray]:
    """bayesian_model_selection.py:48-149 for model_list = [null, subDLA, DLA]."""
    log_priors = np.concatenate([[np.nan], log_priors_sub, log_priors_dla])
    log_priors[0] = np.log(1 - np.exp(logsumexp(log_priors[1:])))  # :79-80
    log_likelihoods = np.concatenate([[log_ev_null], log_ev_sub, log_ev_dla])
    log_posteriors = log_likelihoods + log_priors
    model_posteriors = np.exp(log_posteriors - logsumexp(log_posteriors))  # :126-129
    max_dlas = len(log_ev_dla)
    p_dla = np.sum(model_posteriors[-max_dlas:])  # :141-145
    return dict(log_priors=log_priors, log_likelihoods=log_likelihoods, log_posteriors=log_posteriors,
                model_posteriors=model_posteriors, p_dla=p_dla, p_no_dla=1 - p_dla)


# ------------------------------------------------------------------------------------------
# a13  maximum a posteriori absorber parameters
# ------------------------------------------------------------------------------------------
def maximum_a_posteriori(sample_log_likelihoods: np.ndarray, base_sample_inds: np.ndarray,
                         z_samples: np.ndarray, log_nhi_samples: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """dla_gp.py:428-472 (raises ValueError on an all-NaN column like np.nanargmax)."""
    maxinds = np.nanargmax(sample_log_likelihoods, axis=0)
    max_dlas = sample_log_likelihoods.shape[1]
    MAP_z = np.full((max_dlas, max_dlas), np.nan)
    MAP_lognhi = np.full((max_dlas, max_dlas), np.nan)
    for num_dlas, maxind in enumerate(maxinds):
        inds = np.concatenate([[maxind], base_sample_inds[:num_dlas, maxind]]).astype(int)
        MAP_z[num_dlas, : num_dlas + 1] = z_samples[inds]
        MAP_lognhi[num_dlas, : num_dlas + 1] = log_nhi_samples[inds]
    return MAP_z, MAP_lognhi


# ------------------------------------------------------------------------------------------
# a15  one spectrum, end to end (run_bayes_select.py:141-230)
# ------------------------------------------------------------------------------------------
def process_spectrum(
    model: Dict[str, np.ndarray],
    dla_samples: Dict[str, np.ndarray],
    subdla_samples: Dict[str, np.ndarray],
    prior_counts: Tuple[float, float],
    wavelengths: np.ndarray,
    flux: np.ndarray,
    noise_variance: np.ndarray,
    pixel_mask: np.ndarray,
    z_qso: float,
    max_dlas: int = 4,
    num_lines: int = 3,
    broadening: bool = True,
    uniforms: Optional[np.ndarray] = None,
) -> Dict[str, np.ndarray]:
    """
    run_bayes_select.py:141-230 for one spectrum with BayesModelSelect([0, 1, max_dlas], 2).
    `uniforms` defaults to the first (max_dlas-1) x S draws of RandomState(0)
    (run_bayes_select.py:144; the subDLA model at max_dlas=1 draws nothing).
    """
    S = dla_samples["offset_samples"].shape[0]
    if uniforms is None:
        uniforms = np.random.RandomState(0).random_sample((max(max_dlas - 1, 1), S))
    rest = wavelengths / (1 + z_qso)
    prep = prepare_spectrum(model, rest, flux, noise_variance, pixel_mask, z_qso)
    ev_null = null_log_model_evidence(prep)
    sub = log_model_evidences(prep, subdla_samples["offset_samples"], subdla_samples["nhi_samples"], 1, None,
                              num_lines=num_lines, broadening=broadening)
    dla = log_model_evidences(prep, dla_samples["offset_samples"], dla_samples["nhi_samples"], max_dlas,
                              uniforms, num_lines=num_lines, broadening=broadening)
    m, nq = prior_counts
    lp_sub = dla_log_priors(m, nq, 1, subdla_samples["Z_lls"] / subdla_samples["Z_dla"])
    lp_dla = dla_log_priors(m, nq, max_dlas)
    sel = model_selection(lp_sub, lp_dla, ev_null, sub["log_likelihoods"], dla["log_likelihoods"])
    out = dict(sel)
    out.update(
        prep=prep,
        sample_log_likelihoods_dla=dla["sample_log_likelihoods"],
        base_sample_inds=dla["base_sample_inds"],
        sample_log_likelihoods_lls=sub["sample_log_likelihoods"][:, 0],
        sample_z_dlas=dla["sample_z_dlas"],
        min_z_dla=z_dla_range(wavelengths, z_qso)[0],
        max_z_dla=z_dla_range(wavelengths, z_qso)[1],
    )
    try:
        out["MAP_z_dlas"], out["MAP_log_nhis"] = maximum_a_posteriori(
            dla["sample_log_likelihoods"], dla["base_sample_inds"], dla["sample_z_dlas"], dla_samples["log_nhi_samples"]
        )
    except ValueError:
        out["MAP_z_dlas"] = out["MAP_log_nhis"] = None
    return out
