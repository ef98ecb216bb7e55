"""
voigt_lls.py : Voigt absorption profile of a Lyman-limit system, computed on the B200.

Drop-in for the reference module of the same name (voigt_lls.py:1-363): the Lyman-series profile of
`voigt.voigt_absorption` with the optical depth of the Lyman-limit break folded into the exponent,
    raw = exp(nhi * sum_l(-lc_l V_l) - tau_LLS_break),   tau = nhi / 10^17.2 (lambda_rest / 911.7641 A)^3
for rest wavelengths bluewards of the limit.  Same module-level tables as `voigt`.
"""
import numpy as np

from . import _lib
from .voigt import (  # noqa: F401  (the reference re-declares the same literals, voigt_lls.py:18-224)
    Gammas, c, gammas, instrument_profile, leading_constants, oscillator_strengths, sigma, transition_wavelengths, width,
)

lambda_Lyman_limit: float = 911.7641  # A (voigt_lls.py:226)


def tau_LLS_break(wavelengths: np.ndarray, nhi: float, z_lls: float) -> np.ndarray:
    """Optical depth of the Lyman-limit break (voigt_lls.py:254-284); a 3-operation host formula, kept for callers."""
    rest_wavelengths = np.asarray(wavelengths, dtype=np.float64) / (1 + z_lls)
    tau = np.float64(nhi) / 10**17.2 * (rest_wavelengths / lambda_Lyman_limit) ** 3
    tau[rest_wavelengths > lambda_Lyman_limit] = 0
    return tau


def voigt_absorption(wavelengths: np.ndarray, nhi: float, z_lls: float, num_lines: int = 3,
                     broadening: bool = True) -> np.ndarray:
    """Absorption profile exp(-tau) of one Lyman-limit system (voigt_lls.py:287-363)."""
    return voigt_absorption_batch(wavelengths, np.array([nhi], dtype=np.float64), np.array([z_lls], dtype=np.float64),
                                  num_lines, broadening)[0]


def voigt_absorption_batch(wavelengths: np.ndarray, nhis: np.ndarray, z_llss: np.ndarray, num_lines: int = 3,
                           broadening: bool = True) -> np.ndarray:
    """S profiles on one wavelength grid -> (S, n_out)."""
    wl, nh, zz = _lib.f64(wavelengths), _lib.f64(nhis), _lib.f64(z_llss)
    assert nh.shape == zz.shape and nh.ndim == 1
    n_in = wl.shape[0]
    n_out = n_in - 2 * width if broadening else n_in
    out = np.empty((nh.shape[0], max(n_out, 0)))
    _lib.check(
        _lib.load_library().dla_voigt_lls_absorption_batch(
            _lib.dptr(wl), n_in, _lib.dptr(nh), _lib.dptr(zz), nh.shape[0], int(num_lines), 1 if broadening else 0,
            _lib.dptr(out),
        )
    )
    return out
