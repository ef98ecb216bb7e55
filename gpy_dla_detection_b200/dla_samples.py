"""
dla_samples.py : QMC samples of the DLA parameters theta = (z_DLA, log N_HI).

Mirrors the reference's DLASamples / DLASamplesMAT (dla_samples.py:13-131).  The .mat loader
needs h5py and the published sample file; `DLASamplesArrays` holds the same attributes from
in-memory arrays (the pattern examples/gp_find_lls.py:227-351 uses).
"""
import numpy as np

from .set_parameters import Parameters


class DLASamples:
    """Base class: holds the prior hyper-parameters (dla_samples.py:13-50)."""

    def __init__(self, params: Parameters, prior):
        self.params = params
        self.prior = prior
        self.num_dla_samples = params.num_dla_samples
        self.uniform_min_log_nhi = params.uniform_min_log_nhi
        self.uniform_max_log_nhi = params.uniform_max_log_nhi
        self.fit_min_log_nhi = params.fit_min_log_nhi
        self.fit_max_log_nhi = params.fit_max_log_nhi
        self.alpha = params.alpha


class DLASamplesArrays(DLASamples):
    """offset / log N_HI / N_HI sample arrays given directly."""

    def __init__(self, params: Parameters, prior, offset_samples, log_nhi_samples, nhi_samples=None):
        super().__init__(params, prior)
        self._offset_samples = np.ascontiguousarray(offset_samples, dtype=np.float64)
        self._log_nhi_samples = np.ascontiguousarray(log_nhi_samples, dtype=np.float64)
        self._nhi_samples = (
            10.0**self._log_nhi_samples if nhi_samples is None else np.ascontiguousarray(nhi_samples, dtype=np.float64)
        )

    @property
    def offset_samples(self) -> np.ndarray:
        return self._offset_samples

    @property
    def log_nhi_samples(self) -> np.ndarray:
        return self._log_nhi_samples

    @property
    def nhi_samples(self) -> np.ndarray:
        return self._nhi_samples

    def sample_z_dlas(self, wavelengths: np.ndarray, z_qso: float) -> np.ndarray:
        """z_i = z_min + (z_max - z_min) * offset_i (dla_samples.py:94-104)."""
        lo = self.params.min_z_dla(wavelengths, z_qso)
        return lo + (self.params.max_z_dla(wavelengths, z_qso) - lo) * self._offset_samples

    def pdf(self, log_nhi):
        """log N_HI mixture prior of Garnett et al. 2017 (dla_samples.py:106-131), unit-normalised."""
        from scipy.integrate import quad

        unnorm = lambda x: np.exp(-1.2695 * x**2 + 50.863 * x - 509.33)  # noqa: E731
        Z = quad(unnorm, self.fit_min_log_nhi, 25.0)[0]
        width = self.uniform_max_log_nhi - self.uniform_min_log_nhi
        uniform = ((log_nhi >= self.uniform_min_log_nhi) & (log_nhi <= self.uniform_max_log_nhi)) / width
        return self.alpha * unnorm(log_nhi) / Z + (1 - self.alpha) * uniform


class DLASamplesMAT(DLASamplesArrays):
    """Samples from the published dla_samples_a03.mat (dla_samples.py:53-92); needs h5py."""

    def __init__(self, params: Parameters, prior, dla_samples_file: str = "dla_samples_a03.mat"):
        import h5py  # not in the offline image; imported lazily

        with h5py.File(dla_samples_file, "r") as f:
            assert params.alpha == f["alpha"][0, 0]
            assert params.uniform_min_log_nhi == f["uniform_min_log_nhi"][0, 0]
            super().__init__(
                params, prior, f["offset_samples"][:, 0], f["log_nhi_samples"][:, 0], f["nhi_samples"][:, 0]
            )
            self.uniform_min_log_nhi = f["uniform_min_log_nhi"][0, 0]
            self.uniform_max_log_nhi = f["uniform_max_log_nhi"][0, 0]
