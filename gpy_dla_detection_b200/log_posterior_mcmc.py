"""
log_posterior_mcmc.py : log posterior of absorber parameters for MCMC refinement, on the device.

Drop-in for the reference module of the same name (log_posterior_mcmc.py:16-250): `log_prior`,
`log_posterior`, `sample_log_likelihood_k_dlas`, `this_dla_gp`, `log_mvnpdf_low_rank` keep their
signatures (free functions taking the prepared arrays of a DLAGP, as emcee calls them).  The arrays are
uploaded once per distinct set (a small cache keyed on the arrays' contents) into a prepared-spectrum handle
of the C-ABI (`dla_spectrum_create_prepared`); every call is then one Voigt + one likelihood launch.
`log_posteriors` evaluates a whole ensemble of walkers in one call (emcee `vectorize=True`), which is
how `DLAGP.run_mcmc` drives it.
"""
import ctypes
from typing import Sequence, Tuple

import numpy as np

from . import _lib
from .null_gp import NullGP, _Handle
from .voigt import width as _instrument_width


def log_prior(z_dla: float, log_nhi: float, min_z_dla: float, max_z_dla: float, min_log_nhi: float,
              max_log_nhi: float, pdf) -> float:
    """Uniform prior on z_DLA, data-driven prior on log N_HI (log_posterior_mcmc.py:16-43)."""
    if (z_dla < max_z_dla) and (z_dla > min_z_dla) and (log_nhi > min_log_nhi) and (log_nhi < max_log_nhi):
        return np.log(pdf(log_nhi))
    return -np.inf


class _Prepared:
    """Device handle of one prepared spectrum (y, v, this_mu, this_M, this_omega2, padded grid, mask)."""

    def __init__(self, y, v, padded_wavelengths, this_mu, this_M, this_omega2, pixel_mask, ind_unmasked):
        y, v, mu, om = _lib.f64(y), _lib.f64(v), _lib.f64(this_mu), _lib.f64(this_omega2)
        M = _lib.f64(this_M)
        wl = _lib.f64(padded_wavelengths)
        keep = _lib.u8(~np.asarray(pixel_mask).astype(bool)[np.asarray(ind_unmasked).astype(bool)])  # dla_gp.py:360
        n, k = M.shape
        n_u = keep.shape[0]
        assert wl.shape[0] == n_u + 2 * _instrument_width and int(keep.sum()) == n == y.shape[0]
        ptr = ctypes.c_void_p()
        _lib.check(
            _lib.load_library().dla_spectrum_create_prepared(
                _lib.dptr(y), _lib.dptr(v), _lib.dptr(mu), _lib.dptr(M), _lib.dptr(om), n, k, _lib.dptr(wl),
                wl.shape[0], _lib.bptr(keep), n_u, 1, ctypes.byref(ptr),
            )
        )
        self.handle = _Handle(ptr, "dla_spectrum_destroy")
        self.n = n
        self.this_mu, self.this_M, self.this_omega2 = mu, M, om


_CACHE = {}        # (data fingerprint, model fingerprint) -> _Prepared, at most 8 entries, oldest evicted first


def _fingerprint(*arrays) -> tuple:
    """
    Content key of a set of arrays: shape, dtype and a 128-bit hash of the bytes (about 0.1 ms for a 1 000 x 20
    model).  Keying on id() would silently reuse a stale device copy after an in-place update of y, v or the model
    arrays (re-normalising a spectrum between emcee runs, say).
    """
    import hashlib

    out = []
    for a in arrays:
        a = np.ascontiguousarray(a)
        out.append((a.shape, a.dtype.str, hashlib.blake2b(a.view(np.uint8).reshape(-1).data, digest_size=16).digest()))
    return tuple(out)


def _prepared(y, v, padded_wavelengths, this_mu, this_M, this_omega2, pixel_mask, ind_unmasked) -> _Prepared:
    key = (_fingerprint(y, v), _fingerprint(padded_wavelengths, this_mu, this_M, this_omega2, pixel_mask, ind_unmasked))
    hit = _CACHE.get(key)
    if hit is None:
        if len(_CACHE) >= 8:
            _CACHE.pop(next(iter(_CACHE)))
        hit = _CACHE[key] = _Prepared(y, v, padded_wavelengths, this_mu, this_M, this_omega2, pixel_mask, ind_unmasked)
    return hit


def _prepared_for_model(padded_wavelengths, this_mu, this_M, this_omega2, pixel_mask, ind_unmasked) -> _Prepared:
    """Any cached handle with these model arrays (y and v do not enter the absorption); a new one otherwise."""
    model_key = _fingerprint(padded_wavelengths, this_mu, this_M, this_omega2, pixel_mask, ind_unmasked)
    for (_, mk), prep in _CACHE.items():
        if mk == model_key:
            return prep
    dummy = np.ones(np.asarray(this_mu).shape[0])
    return _prepared(dummy, dummy, padded_wavelengths, this_mu, this_M, this_omega2, pixel_mask, ind_unmasked)


def sample_log_likelihoods(z_dlas: np.ndarray, nhis: np.ndarray, y, v, padded_wavelengths, this_mu, this_M,
                           this_omega2, pixel_mask, ind_unmasked, num_lines: int) -> np.ndarray:
    """Vectorised `sample_log_likelihood_k_dlas`: z_dlas, nhis of shape (W, k_dlas) -> (W,)."""
    prep = _prepared(y, v, padded_wavelengths, this_mu, this_M, this_omega2, pixel_mask, ind_unmasked)
    zz, nn = _lib.f64(z_dlas), _lib.f64(nhis)
    assert zz.shape == nn.shape and zz.ndim == 2
    out = np.empty((zz.shape[0],))
    _lib.check(
        _lib.load_library().dla_sample_log_likelihoods(
            prep.handle.ptr, _lib.dptr(zz), _lib.dptr(nn), zz.shape[0], zz.shape[1], int(num_lines), _lib.dptr(out)
        )
    )
    return out


def sample_log_likelihood_k_dlas(z_dlas: np.ndarray, nhis: np.ndarray, y, v, padded_wavelengths, this_mu, this_M,
                                 this_omega2, pixel_mask, ind_unmasked, num_lines: int) -> float:
    """log p(y | k absorbers at (z_dlas, nhis)) (log_posterior_mcmc.py:100-136)."""
    assert len(z_dlas) == len(nhis)
    return float(sample_log_likelihoods(np.asarray(z_dlas)[None, :], np.asarray(nhis)[None, :], y, v,
                                        padded_wavelengths, this_mu, this_M, this_omega2, pixel_mask, ind_unmasked,
                                        num_lines)[0])


def this_dla_gp(z_dlas: np.ndarray, nhis: np.ndarray, padded_wavelengths, this_mu, this_M, this_omega2, pixel_mask,
                ind_unmasked, num_lines: int) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """(dla_mu, dla_M, dla_omega2) with k absorbers applied (log_posterior_mcmc.py:139-198)."""
    assert len(z_dlas) == len(nhis)
    # y and v are not part of this signature: the handle cached for the same model arrays is reused
    prep = _prepared_for_model(padded_wavelengths, this_mu, this_M, this_omega2, pixel_mask, ind_unmasked)
    zz, nn = _lib.f64(z_dlas), _lib.f64(nhis)
    absorption = np.empty((prep.n,))
    _lib.check(
        _lib.load_library().dla_absorption_k_dlas(prep.handle.ptr, _lib.dptr(zz), _lib.dptr(nn), zz.shape[0],
                                                  int(num_lines), _lib.dptr(absorption))
    )
    return prep.this_mu * absorption, prep.this_M * absorption[:, None], prep.this_omega2 * absorption**2


def log_mvnpdf_low_rank(y, mu, M, d, scipy_lapack: bool = True) -> float:
    """log N(y; mu, MM' + diag(d)) (log_posterior_mcmc.py:200-250)."""
    return NullGP.log_mvnpdf_low_rank(y, mu, M, d, scipy_lapack)


def log_posterior(theta: Sequence[float], this_wavelengths, y, v, z_qso, min_z_dla, max_z_dla, min_log_nhi,
                  max_log_nhi, pdf, padded_wavelengths, this_mu, this_M, this_omega2, pixel_mask, ind_unmasked,
                  num_lines) -> float:
    """log p(theta | y) up to a constant for one (z_dla, log_nhi) pair (log_posterior_mcmc.py:46-97)."""
    z_dla, log_nhi = theta
    lp = log_prior(z_dla, log_nhi, min_z_dla, max_z_dla, min_log_nhi, max_log_nhi, pdf)
    if not np.isfinite(lp):
        return -np.inf
    return lp + sample_log_likelihood_k_dlas(np.array([z_dla]), 10 ** np.array([log_nhi]), y, v, padded_wavelengths,
                                             this_mu, this_M, this_omega2, pixel_mask, ind_unmasked, num_lines)


def log_posteriors(thetas: np.ndarray, this_wavelengths, y, v, z_qso, min_z_dla, max_z_dla, min_log_nhi,
                   max_log_nhi, pdf, padded_wavelengths, this_mu, this_M, this_omega2, pixel_mask, ind_unmasked,
                   num_lines) -> np.ndarray:
    """`log_posterior` for an ensemble: thetas (W, 2) -> (W,), one device call for the walkers inside the prior."""
    thetas = np.asarray(thetas, dtype=np.float64)
    lp = np.array([log_prior(t[0], t[1], min_z_dla, max_z_dla, min_log_nhi, max_log_nhi, pdf) for t in thetas])
    out = np.full(thetas.shape[0], -np.inf)
    ok = np.isfinite(lp)
    if np.any(ok):
        ll = sample_log_likelihoods(thetas[ok, 0:1], 10 ** thetas[ok, 1:2], y, v, padded_wavelengths, this_mu, this_M,
                                    this_omega2, pixel_mask, ind_unmasked, num_lines)
        out[ok] = lp[ok] + ll
    return out
