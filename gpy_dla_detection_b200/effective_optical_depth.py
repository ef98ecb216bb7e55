"""
effective_optical_depth.py : effective optical depth of the Lyman-series forest.

Drop-in for effective_optical_depth.effective_optical_depth (effective_optical_depth.py:10-80),
evaluated on the device (csrc/prep_kernel.cuh).  As in the reference, `skip_lya_indicator`
is accepted and not used: the indicator z_absorber <= z_qso multiplies every member.
"""
import numpy as np

from . import _lib


def effective_optical_depth(
    wavelengths: np.ndarray,
    beta: float,
    tau_0: float,
    z_qso: float,
    num_forest_lines: int,
    skip_lya_indicator: bool = True,
) -> np.ndarray:
    """-> (n_points, num_forest_lines) optical depths tau_0 (f_i l_i)/(f_1 l_1) (1 + z_i)^beta."""
    wl = _lib.f64(wavelengths)
    out = np.empty((wl.shape[0], int(num_forest_lines)))
    _lib.check(
        _lib.load_library().dla_effective_optical_depth(
            _lib.dptr(wl), wl.shape[0], float(beta), float(tau_0), float(z_qso), int(num_forest_lines), _lib.dptr(out)
        )
    )
    return out
