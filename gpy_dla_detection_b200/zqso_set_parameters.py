"""
zqso_set_parameters.py : parameters of the quasar-redshift (zQSO) estimation model.

Drop-in for the reference's ZParameters (zqso_set_parameters.py:19-54).  Like the reference it
does NOT run Parameters.__init__: only the attributes below exist on an instance (plus the
physical constants and helper methods inherited from the class).
"""
from .set_parameters import Parameters


class ZParameters(Parameters):
    def __init__(
        self,
        normalization_min_lambda: float = 1216.0 - 40.0,  # rest-frame window used for flux normalisation (A)
        normalization_max_lambda: float = 1216.0 + 40.0,
        min_lambda: float = 910.0,   # rest wavelengths modelled by the GP (A)
        max_lambda: float = 3000.0,
        dlambda: float = 0.25,
        k: int = 20,
        max_noise_variance: float = 4.0 ** 2,
        num_zqso_samples: int = 10000,
        minFunc_options: dict = None,
    ):
        self.normalization_min_lambda = normalization_min_lambda
        self.normalization_max_lambda = normalization_max_lambda
        self.min_lambda = min_lambda
        self.max_lambda = max_lambda
        self.dlambda = dlambda
        self.k = k
        self.max_noise_variance = max_noise_variance
        self.num_zqso_samples = num_zqso_samples
        self.minFunc_options = minFunc_options or {"MaxIter": 4000, "MaxFunEvals": 8000}
