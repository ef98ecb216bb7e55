"""
lls_gp.py : DLAGP whose absorbers carry the Lyman-limit break (voigt_lls.py).

The reference shows the pattern in examples/gp_find_lls.py:159-224 - a DLAGP subclass that overrides
`this_dla_gp` to call `voigt_lls.voigt_absorption`; everything else (QMC loop, evidences, resampling, MAP)
is inherited.  Here the override is a flag on the device-resident spectrum: every profile the handle
evaluates afterwards includes the break.
"""
import numpy as np

from . import _lib
from .dla_gp import DLAGP


class LLSGP(DLAGP):
    def set_data(self, X: np.ndarray, Y: np.ndarray, noise_variance: np.ndarray, pixel_mask: np.ndarray, z_qso: float,
                 normalize: bool = True, build_model: bool = True) -> None:
        super().set_data(X, Y, noise_variance, pixel_mask, z_qso, normalize, build_model)
        _lib.check(_lib.load_library().dla_spectrum_set_lls_break(self._spectrum.ptr, 1))
