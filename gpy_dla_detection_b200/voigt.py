"""
voigt.py : Lyman-series Voigt absorption profile, computed on the B200.

Drop-in for the reference module of the same name (voigt.py:1-322): the Lyman-series
tables keep their names and `voigt_absorption` keeps its signature and output length;
the arithmetic runs in the fused CUDA profile kernel (csrc/voigt_kernel.cuh) through
`dla_voigt_absorption` of the C-ABI.
"""
import numpy as np

from . import _lib
from . import _tables

# same module-level names as the reference (voigt.py:18-224)
c: float = _tables.SPEED_OF_LIGHT_CGS
transition_wavelengths: np.ndarray = _tables.TRANSITION_WAVELENGTHS
oscillator_strengths: np.ndarray = _tables.OSCILLATOR_STRENGTHS
Gammas: np.ndarray = _tables.TRANSITION_RATES
sigma: float = _tables.SIGMA
leading_constants: np.ndarray = _tables.LEADING_CONSTANTS
gammas: np.ndarray = _tables.GAMMAS
width: int = _tables.WIDTH
instrument_profile: np.ndarray = _tables.INSTRUMENT_PROFILE


def voigt_absorption(
    wavelengths: np.ndarray,
    nhi: float,
    z_dla: float,
    num_lines: int = 3,
    broadening: bool = True,
) -> np.ndarray:
    """
    Absorption profile exp(-tau) of one absorber (voigt.py:251-322).

    :param wavelengths: observed wavelengths (Angstrom)
    :param nhi: column density (cm^-2); note the argument order (nhi, z_dla)
    :param z_dla: absorber redshift
    :param num_lines: members of the Lyman series to include (1..31)
    :param broadening: apply the 7-tap SDSS instrument profile ('valid' convolution,
        output is 6 points shorter than the input)
    """
    wl = _lib.f64(wavelengths)
    n_in = wl.shape[0]
    n_out = n_in - 2 * width if broadening else n_in
    out = np.empty((max(n_out, 0),))
    _lib.check(
        _lib.load_library().dla_voigt_absorption(
            _lib.dptr(wl), n_in, float(nhi), float(z_dla), int(num_lines), 1 if broadening else 0, _lib.dptr(out)
        )
    )
    return out


def voigt_absorption_batch(
    wavelengths: np.ndarray,
    nhis: np.ndarray,
    z_dlas: np.ndarray,
    num_lines: int = 3,
    broadening: bool = True,
) -> np.ndarray:
    """S profiles on one wavelength grid -> (S, n_out); same arithmetic as `voigt_absorption`."""
    wl = _lib.f64(wavelengths)
    nh = _lib.f64(nhis)
    zz = _lib.f64(z_dlas)
    assert nh.shape == zz.shape and nh.ndim == 1
    n_in = wl.shape[0]
    n_out = n_in - 2 * width if broadening else n_in
    out = np.empty((nh.shape[0], max(n_out, 0)))
    _lib.check(
        _lib.load_library().dla_voigt_absorption_batch(
            _lib.dptr(wl), n_in, _lib.dptr(nh), _lib.dptr(zz), nh.shape[0], int(num_lines),
            1 if broadening else 0, _lib.dptr(out),
        )
    )
    return out


def faddeeva_re(x: np.ndarray, y: np.ndarray) -> np.ndarray:
    """Re w(x + iy) as the profile kernel evaluates it (0 <= y <= 1e-3); for accuracy tests."""
    xx = _lib.f64(x).ravel()
    yy = _lib.f64(np.broadcast_to(y, np.shape(x))).ravel()
    out = np.empty_like(xx)
    _lib.check(_lib.load_library().dla_faddeeva_re(_lib.dptr(xx), _lib.dptr(yy), xx.shape[0], _lib.dptr(out)))
    return out.reshape(np.shape(x))
