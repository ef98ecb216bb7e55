"""
zqso_samples.py : parameter samples for the zQSO estimation (reference zqso_samples.py:10-29).
"""
import numpy as np

from .zqso_set_parameters import ZParameters


class ZSamples:
    """Linearly spaced z_QSO samples, as the reference's Python code uses instead of QMC samples."""

    def __init__(self, params: ZParameters):
        self.params = params
        self.num_zqso_samples = params.num_zqso_samples

    def sample_z_qsos(self, z_qso_min: float = 2.14, z_qso_max: float = 6.16) -> np.ndarray:
        return np.linspace(z_qso_min, z_qso_max, self.num_zqso_samples)
