"""
subdla_samples.py : QMC samples of the sub-DLA alternative model.

Mirrors SubDLASamples / SubDLASamplesMAT (subdla_samples.py:12-125): the same offsets,
log N_HI ~ U(extrapolate_min_log_nhi, 20) and the partition functions `Z_lls`, `Z_dla`
(also exposed as `_Z_lls`, `_Z_dla`, which SubDLAGP.log_priors reads).
"""
import numpy as np

from .set_parameters import Parameters
from .dla_samples import DLASamples


class SubDLASamples(DLASamples):
    def __init__(self, params: Parameters, prior, extrapolate_min_log_nhi: float):
        self.extrapolate_min_log_nhi = extrapolate_min_log_nhi
        super().__init__(params, prior)


class SubDLASamplesArrays(SubDLASamples):
    def __init__(
        self, params: Parameters, prior, offset_samples, log_nhi_samples, nhi_samples, Z_lls: float, Z_dla: float,
        extrapolate_min_log_nhi: float = 19.5,
    ):
        super().__init__(params, prior, extrapolate_min_log_nhi)
        self._offset_samples = np.ascontiguousarray(offset_samples, dtype=np.float64)
        self._log_nhi_samples = np.ascontiguousarray(log_nhi_samples, dtype=np.float64)
        self._nhi_samples = np.ascontiguousarray(nhi_samples, dtype=np.float64)
        self._Z_lls = float(Z_lls)
        self._Z_dla = float(Z_dla)

    Z_dla = property(lambda self: self._Z_dla)
    Z_lls = property(lambda self: self._Z_lls)
    offset_samples = property(lambda self: self._offset_samples)
    log_nhi_samples = property(lambda self: self._log_nhi_samples)
    nhi_samples = property(lambda self: self._nhi_samples)

    def sample_z_lls(self, wavelengths: np.ndarray, z_qso: float) -> np.ndarray:
        """subdla_samples.py:115-125"""
        lo = self.params.min_z_dla(wavelengths, z_qso)
        return lo + (self.params.max_z_dla(wavelengths, z_qso) - lo) * self._offset_samples


class SubDLASamplesMAT(SubDLASamplesArrays):
    """Samples from the published subdla_samples.mat (subdla_samples.py:66-125); needs h5py."""

    def __init__(self, params: Parameters, prior, sub_dla_samples_file: str = "subdla_samples.mat"):
        import h5py

        with h5py.File(sub_dla_samples_file, "r") as f:
            assert params.alpha == f["alpha"][0, 0]
            assert params.num_dla_samples == f["num_dla_samples"][0, 0]
            super().__init__(
                params, prior, f["offset_samples"][:, 0], f["lls_log_nhi_samples"][:, 0], f["lls_nhi_samples"][:, 0],
                f["Z_lls"][0, 0], f["Z_dla"][0, 0], f["extrapolate_min_log_nhi"][0, 0],
            )
