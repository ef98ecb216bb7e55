"""
dla_gp.py : GP model with up to `max_dlas` intervening DLAs.

Drop-in for the reference's DLAGP (dla_gp.py:25-472): same constructor,
`log_model_evidences`, `sample_log_likelihood_k_dlas`, `this_dla_gp`, `log_priors`,
`maximum_a_posteriori`, and the attributes `sample_log_likelihoods` (S, max_dlas) and
`base_sample_inds` (max_dlas-1, S) int32.  `run_mcmc` (emcee) is outside the hot path.
"""
from typing import Optional

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

    def mcmc_log_posterior_args(self):
        """The argument tuple of log_posterior_mcmc.log_posterior for this spectrum (dla_gp.py:265-287)."""
        from scipy import stats
        from scipy.integrate import quad

        ds = self.dla_samples
        min_z_dla = self.params.min_z_dla(self.this_wavelengths, self.z_qso)
        max_z_dla = self.params.max_z_dla(self.this_wavelengths, self.z_qso)
        u = stats.uniform(loc=ds.uniform_min_log_nhi, scale=ds.uniform_max_log_nhi - ds.uniform_min_log_nhi)

        def unnormalized_pdf(nhi):
            return np.exp(-1.2695 * nhi**2 + 50.863 * nhi - 509.33)

        Z = quad(unnormalized_pdf, ds.fit_min_log_nhi, 25.0)[0]

        def normalized_pdf(nhi):
            return ds.alpha * (unnormalized_pdf(nhi) / Z) + (1 - ds.alpha) * u.pdf(nhi)

        return (self.this_wavelengths, self.y, self.v, self.z_qso, min_z_dla, max_z_dla, ds.uniform_min_log_nhi,
                ds.uniform_max_log_nhi, normalized_pdf, self.padded_wavelengths, self.this_mu, self.this_M,
                self.this_omega2, self.pixel_mask, self.ind_unmasked, self.params.num_lines)

    def run_mcmc(self, nwalkers: int, kth_dla: int = 1, nsamples: int = 5000, pos: Optional[np.ndarray] = None,
                 skip_initial_state_check: bool = True):
        """
        emcee sampling of the 1-DLA posterior (dla_gp.py:227-309).  The walkers of a step are evaluated in one
        device call (`vectorize=True` with log_posterior_mcmc.log_posteriors).  Needs emcee, which is not part
        of this image.
        """
        import emcee  # noqa: F401  (ImportError here when the package is absent, as for the reference)

        from .log_posterior_mcmc import log_posteriors

        sampler = emcee.EnsembleSampler(nwalkers, kth_dla * 2, log_posteriors, args=self.mcmc_log_posterior_args(),
                                        vectorize=True)
        if pos is None:
            sample_z_dlas = self._sample_z()
            pos = np.concatenate([np.random.choice(sample_z_dlas, size=nwalkers)[:, None],
                                  np.random.choice(self.dla_samples.log_nhi_samples, size=nwalkers)[:, None]], axis=1)
            assert pos.shape[0] == nwalkers
        sampler.run_mcmc(pos, nsamples, progress=True, skip_initial_state_check=skip_initial_state_check)
        return sampler
