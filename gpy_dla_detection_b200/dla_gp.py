"""
dla_gp.py : GP model with up to `max_dlas` intervening DLAs.

Drop-in for the reference's DLAGP (dla_gp.py:25-472): same constructor,
`log_model_evidences`, `sample_log_likelihood_k_dlas`, `this_dla_gp`, `log_priors`,
`maximum_a_posteriori`, and the attributes `sample_log_likelihoods` (S, max_dlas) and
`base_sample_inds` (max_dlas-1, S) int32.  `run_mcmc` (emcee) is outside the hot path.
"""
import numpy as np

from .set_parameters import Parameters
from ._absorber_gp import AbsorberGP


class DLAGP(AbsorberGP):
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
        return self.dla_samples.sample_z_dlas(self.this_wavelengths, self.z_qso)  # dla_gp.py:122-124

    def log_model_evidences(self, max_dlas: int) -> np.ndarray:
        """[log p(D | 1 DLA), ..., log p(D | max_dlas DLAs)] (dla_gp.py:92-225)."""
        log_ev, sample_ll, base_inds = self._log_model_evidences(max_dlas)
        self.sample_log_likelihoods = sample_ll
        self.base_sample_inds = base_inds
        return log_ev

    def log_priors(self, z_qso: float, max_dlas: int) -> np.ndarray:
        """P(k DLAs | z_QSO) = (M/N)^k - (M/N)^(k+1), last one (M/N)^max (dla_gp.py:398-426)."""
        this_num_dlas, this_num_quasars = self.prior.less_ind(z_qso)
        p_dlas = (this_num_dlas / this_num_quasars) ** np.arange(1, max_dlas + 1)
        for i in range(max_dlas - 1):
            p_dlas[i] = p_dlas[i] - p_dlas[i + 1]
        return np.log(p_dlas)

    def maximum_a_posteriori(self):
        """
        MAP (z_dla, log N_HI) of every DLA(k) model -> two (max_dlas, max_dlas) NaN-padded arrays
        (dla_gp.py:428-472).  Like np.nanargmax, raises ValueError on an all-NaN level.
        """
        maxinds = np.nanargmax(self.sample_log_likelihoods, axis=0)
        max_dlas = self.sample_log_likelihoods.shape[1]
        MAP_z_dla = np.full((max_dlas, max_dlas), np.nan)
        MAP_log_nhi = np.full((max_dlas, max_dlas), np.nan)
        sample_z_dlas = self._sample_z()
        log_nhi = self.dla_samples.log_nhi_samples
        for num_dlas, maxind in enumerate(maxinds):
            chain = np.concatenate([[maxind], self.base_sample_inds[:num_dlas, maxind]]).astype(int)
            MAP_z_dla[num_dlas, : num_dlas + 1] = sample_z_dlas[chain]
            MAP_log_nhi[num_dlas, : num_dlas + 1] = log_nhi[chain]
        return MAP_z_dla, MAP_log_nhi

    def run_mcmc(self, *args, **kwargs):
        raise NotImplementedError("MCMC refinement (emcee) is outside the B200 hot path; see SURVEY.md §8f")
