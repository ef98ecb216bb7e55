"""
zqso_oracle.py : CPU restatement (NumPy float64) of the reference's quasar-redshift estimation
path, ZGP (gpy_dla_detection/zqso_gp.py).  TEST INFRASTRUCTURE ONLY - same rules as
oracle/dla_oracle.py: imported by tests/ (and bench.py's CPU legs) as the checker, never by the
product package.

Parity status: PINNED - tests/golden/zqso_*.npz are written by tests/golden/make_golden.py from
the live reference's ZGP.inference_z_qso / set_data / log_model_evidence on seeded synthetic
inputs, and tests/test_zqso_oracle.py checks this restatement against them.

The per-sample arithmetic follows the reference line by line; only the Python object plumbing is
dropped.  np.interp replaces scipy.interpolate.interp1d(kind='linear') (same
slope * (x - x_lo) + y_lo arithmetic; checked against the golden this_mu / this_M).
"""
from typing import Dict

import numpy as np

from oracle.dla_oracle import log_mvnpdf_low_rank

LOG_2PI = 1.83787706640934534  # zqso_gp.py:263


def log_mvnpdf_iid(y: np.ndarray, mu: np.ndarray, d: np.ndarray) -> float:
    """zqso_gp.py:252-278 : log N(y; mu, diag(d))"""
    n = d.shape[0]
    r = y[:, None] - mu[:, None]
    d_inv = 1 / d[:, None]
    log_det = np.sum(np.log(d))
    return -0.5 * (np.matmul(r.T, d_inv * r).sum() + log_det + n * LOG_2PI)


def interp_model(model: Dict[str, np.ndarray], x: np.ndarray):
    """zqso_gp.py:66-90 (interp1d linear, bounds_error) -> this_mu (n,), this_M (n, k)"""
    rw = model["rest_wavelengths"]
    if x.size and (x.min() < rw[0] or x.max() > rw[-1]):
        raise ValueError("A value in x_new is outside the interpolation range.")
    from scipy import interpolate

    this_mu = interpolate.interp1d(rw, model["mu"])(x)
    this_M = np.stack([interpolate.interp1d(rw, col)(x) for col in model["M"].T], axis=1) if x.size else np.empty(
        (0, model["M"].shape[1]))
    return this_mu, this_M


def set_data(model: Dict[str, np.ndarray], X: np.ndarray, Y: np.ndarray, noise_variance: np.ndarray,
             pixel_mask: np.ndarray, z_qso: float, min_lambda: float = 910.0, max_lambda: float = 3000.0,
             normalization_min_lambda: float = 1176.0, normalization_max_lambda: float = 1256.0) -> Dict[str, np.ndarray]:
    """zqso_gp.py:92-182 with normalize=True, build_model=True.  X are OBSERVED wavelengths."""
    max_pos_lambda = max_lambda * (1 + z_qso)
    min_pos_lambda = min_lambda * (1 + z_qso)
    max_observed_lambda = np.min((max_pos_lambda, np.max(X)))
    min_observed_lambda = np.max((min_pos_lambda, np.min(X)))
    ind = (X > min_observed_lambda) * (X < max_observed_lambda)  # strict on both sides (:132)
    y = Y[ind]
    this_wavelengths = X[ind]
    v = noise_variance[ind]
    mask_in = pixel_mask[ind]
    x = X[ind] / (1 + z_qso)
    ind_n = (x >= normalization_min_lambda) & (x <= normalization_max_lambda)  # ignores the pixel mask (:143-146)
    with np.errstate(all="ignore"):
        this_median = np.nanmedian(y[ind_n]) if np.any(ind_n) else np.nan
    y = y / this_median
    v = v / this_median**2
    this_normalized_flux = Y / this_median
    this_normalized_v = noise_variance / this_median**2
    ind_bw = (X < min_observed_lambda) & (~pixel_mask)
    ind_rw = (X > max_observed_lambda) & (~pixel_mask)
    ind2 = (x >= min_lambda) & (x <= max_lambda) & (~mask_in)
    this_wavelengths, x, y, v = this_wavelengths[ind2], x[ind2], y[ind2], v[ind2]
    with np.errstate(all="ignore"):
        v[np.isinf(v)] = np.nanmean(v) if v.size else np.nan  # :177
    this_mu, this_M = interp_model(model, x)
    return dict(x=x, y=y, v=v, this_wavelengths=this_wavelengths, ind=ind2, this_mu=this_mu, this_M=this_M,
                y_bw=this_normalized_flux[ind_bw], v_bw=this_normalized_v[ind_bw],
                y_rw=this_normalized_flux[ind_rw], v_rw=this_normalized_v[ind_rw], this_median=this_median, z_qso=z_qso)


def log_model_evidence(model: Dict[str, np.ndarray], d: Dict[str, np.ndarray]) -> float:
    """zqso_gp.py:184-212"""
    with np.errstate(all="ignore"):
        try:
            ll = log_mvnpdf_low_rank(d["y"], d["this_mu"], d["this_M"], d["v"])
        except np.linalg.LinAlgError:
            ll = np.nan
        n_bw, n_rw = d["y_bw"].shape[0], d["y_rw"].shape[0]
        bw = log_mvnpdf_iid(d["y_bw"], model["bluewards_mu"] * np.ones((n_bw,)),
                            model["bluewards_sigma"] ** 2 * np.ones((n_bw,)) + d["v_bw"])
        rw = log_mvnpdf_iid(d["y_rw"], model["redwards_mu"] * np.ones((n_rw,)),
                            model["redwards_sigma"] ** 2 * np.ones((n_rw,)) + d["v_rw"])
    return ll + bw + rw


def inference_z_qso(model: Dict[str, np.ndarray], wavelengths: np.ndarray, flux: np.ndarray,
                    noise_variance: np.ndarray, pixel_mask: np.ndarray, sample_z_qsos: np.ndarray,
                    **window) -> Dict[str, np.ndarray]:
    """zqso_gp.py:214-250 : sample log-likelihoods over the z_QSO samples and the MAP redshift."""
    ll = np.full((sample_z_qsos.shape[0],), np.nan)
    for i, z in enumerate(sample_z_qsos):
        ll[i] = log_model_evidence(model, set_data(model, wavelengths, flux, noise_variance, pixel_mask, float(z), **window))
    return dict(sample_log_likelihoods=ll, z_map=sample_z_qsos[np.nanargmax(ll)])
