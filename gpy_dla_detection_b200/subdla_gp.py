"""
subdla_gp.py : GP model with a sub-DLA (log N_HI in [19.5, 20)) as the alternative model.

Drop-in for the reference's SubDLAGP (subdla_gp.py:23-346).  Differences from DLAGP as in
the reference: z samples come from `dla_samples.sample_z_lls`, the default `max_dlas` is 1,
`base_sample_inds` is not stored, and the priors carry the factor Z_lls / Z_dla.
"""
import numpy as np

from .set_parameters import Parameters
from ._absorber_gp import AbsorberGP


class SubDLAGP(AbsorberGP):
    def __init__(
        self,
        params: Parameters,
        prior,
        dla_samples,
        rest_wavelengths: np.ndarray,
        mu: np.ndarray,
        M: np.ndarray,
        log_omega: np.ndarray,
        log_c_0: float,
        log_tau_0: float,
        log_beta: float,
        prev_tau_0: float = 0.0023,
        prev_beta: float = 3.65,
        min_z_separation: float = 3000.0,
        broadening: bool = True,
    ):
        super().__init__(
            params, prior, rest_wavelengths, mu, M, log_omega, log_c_0, log_tau_0, log_beta, prev_tau_0, prev_beta
        )
        self._init_absorber(dla_samples, min_z_separation, broadening)

    def _sample_z(self) -> np.ndarray:
        return self.dla_samples.sample_z_lls(self.this_wavelengths, self.z_qso)  # subdla_gp.py:120-122

    def log_model_evidences(self, max_dlas: int = 1) -> np.ndarray:
        """[log p(D | 1 subDLA), ...] (subdla_gp.py:90-222)."""
        log_ev, sample_ll, _ = self._log_model_evidences(max_dlas)
        self.sample_log_likelihoods = sample_ll
        return log_ev

    def log_priors(self, z_qso: float, max_dlas: int) -> np.ndarray:
        """DLA priors scaled by the ratio of the log N_HI partition functions (subdla_gp.py:311-346)."""
        this_num_dlas, this_num_quasars = self.prior.less_ind(z_qso)
        p_dlas = (
            self.dla_samples._Z_lls
            / self.dla_samples._Z_dla
            * (this_num_dlas / this_num_quasars) ** np.arange(1, max_dlas + 1)
        )
        for i in range(max_dlas - 1):
            p_dlas[i] = p_dlas[i] - p_dlas[i + 1]
        return np.log(p_dlas)
