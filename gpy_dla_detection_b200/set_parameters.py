"""
set_parameters.py : pipeline parameters of the DLA finder.

Drop-in for the reference's `Parameters` (set_parameters.py:14-165): same constructor
keywords, attribute names, class-level physical constants and helper methods
(`kms_to_z`, `emitted_wavelengths`, `observed_wavelengths`, `min_z_dla`, `max_z_dla`).
The z-range arithmetic (set_parameters.py:125-159) is reproduced operation for operation
because the QMC z_DLA samples are derived from it.
"""
import numpy as np

_DEFAULTS = dict(
    # file loading
    loading_min_lambda=910.0,
    loading_max_lambda=1217.0,
    # preprocessing
    z_qso_cut=2.15,
    min_num_pixels=200,
    # flux normalisation window (rest frame, Angstrom)
    normalization_min_lambda=1310.0,
    normalization_max_lambda=1325.0,
    # null model
    min_lambda=911.75,
    max_lambda=1215.75,
    dlambda=0.25,
    k=20,
    max_noise_variance=3.0**2,
    # optimiser start values (training only; kept for API parity)
    initial_c_0=0.1,
    initial_tau_0=0.0023,
    initial_beta=3.65,
    minFunc_options=None,
    # QMC samples
    num_dla_samples=10000,
    alpha=0.97,
    uniform_min_log_nhi=20.0,
    uniform_max_log_nhi=23.0,
    fit_min_log_nhi=20.0,
    fit_max_log_nhi=22.0,
    # model prior
    prior_z_qso_increase=30000.0,
    # instrumental broadening
    width=3,
    pixel_spacing=1e-4,
    # absorber model
    num_lines=3,
    max_z_cut=3000.0,
    min_z_cut=3000.0,
    num_forest_lines=31,
)

# attributes given in km/s and stored as redshift differences
_KMS_ATTRS = ("prior_z_qso_increase", "max_z_cut", "min_z_cut")


class Parameters:
    # physical constants (set_parameters.py:15-19)
    lya_wavelength: float = 1215.6701
    lyb_wavelength: float = 1025.7223
    lyman_limit: float = 911.7633
    speed_of_light: float = 299792458.0

    def __init__(self, **kwargs):
        unknown = set(kwargs) - set(_DEFAULTS)
        if unknown:
            raise TypeError("unknown Parameters keyword(s): {}".format(sorted(unknown)))
        for name, default in _DEFAULTS.items():
            value = kwargs.get(name, default)
            if name == "minFunc_options" and value is None:
                value = {"MaxIter": 2000, "MaxFunEvals": 4000}
            if name in _KMS_ATTRS:
                value = self.kms_to_z(value)
            setattr(self, name, value)

    @classmethod
    def kms_to_z(cls, kms: float) -> float:
        """relative velocity in km/s -> redshift difference (set_parameters.py:104-109)"""
        return (kms * 1000) / cls.speed_of_light

    @staticmethod
    def emitted_wavelengths(observed_wavelengths: np.ndarray, z: float) -> np.ndarray:
        return observed_wavelengths / (1 + z)

    @staticmethod
    def observed_wavelengths(emitted_wavelengths: np.ndarray, z: float) -> np.ndarray:
        return emitted_wavelengths * (1 + z)

    def _in_model_range(self, wavelengths: np.ndarray, z_qso: float) -> np.ndarray:
        rest = self.emitted_wavelengths(wavelengths, z_qso)
        return wavelengths[(rest >= self.min_lambda) & (rest <= self.max_lambda)]

    def max_z_dla(self, wavelengths: np.ndarray, z_qso: float) -> float:
        """largest z_DLA searched (set_parameters.py:125-140)"""
        inside = self._in_model_range(wavelengths, z_qso)
        return np.min(
            [(np.max(inside) / self.lya_wavelength - 1) - self.max_z_cut, z_qso - self.max_z_cut]
        )

    def min_z_dla(self, wavelengths: np.ndarray, z_qso: float) -> float:
        """smallest z_DLA searched (set_parameters.py:142-159)"""
        inside = self._in_model_range(wavelengths, z_qso)
        return np.max(
            [
                np.min(inside) / self.lya_wavelength - 1,
                self.observed_wavelengths(self.lyman_limit, z_qso) / self.lya_wavelength
                - 1
                + self.min_z_cut,
            ]
        )

    def __repr__(self):
        return str(self.__dict__)
