"""
bayesian_model_selection.py : DLA classification by Bayesian model selection.

Drop-in for BayesModelSelect (bayesian_model_selection.py:21-149): priors, evidences and
posteriors of [null, subDLA, DLA(1..k)].  The evidences come from the device path of the
model objects; the remaining arithmetic is over 2 + max_dlas numbers.
"""
from itertools import chain
from typing import List

import numpy as np
from scipy.special import logsumexp

from .null_gp import NullGP
from .dla_gp import DLAGP


class BayesModelSelect:
    def __init__(self, all_max_dlas: List[int] = [0, 1, 4], dla_model_ind: int = 2):
        self.all_max_dlas = all_max_dlas
        self.dla_model_ind = dla_model_ind

    def model_selection(self, model_list: List[NullGP], z_qso: float) -> np.ndarray:
        """log posteriors of every model in `model_list` (bayesian_model_selection.py:48-109)."""
        assert isinstance(model_list[-1], DLAGP)
        assert isinstance(model_list[-1], NullGP)
        assert len(model_list) > self.dla_model_ind

        # priors first: the null prior is one minus all the others (:66-80)
        log_priors = [
            [np.nan] if num_dlas == 0 else model.log_priors(z_qso, num_dlas)
            for model, num_dlas in zip(model_list, self.all_max_dlas)
        ]
        log_priors = np.array(list(chain(*log_priors)))
        log_priors[0] = np.log(1 - np.exp(logsumexp(log_priors[1:])))

        # evidences in list order: null -> subDLA -> DLA (the RNG consumption order, :84-98)
        log_likelihoods = [
            [model.log_model_evidence()] if num_dlas == 0 else model.log_model_evidences(num_dlas)
            for model, num_dlas in zip(model_list, self.all_max_dlas)
        ]
        log_likelihoods = np.array(list(chain(*log_likelihoods)))
        log_posteriors = log_likelihoods + log_priors

        self.log_priors = log_priors
        self.log_likelihoods = log_likelihoods
        self.log_posteriors = log_posteriors
        return log_posteriors

    @property
    def dla_model_posterior_ind(self):
        ind = np.zeros((self.log_posteriors.shape[0],), dtype=np.bool_)
        ind[-self.all_max_dlas[self.dla_model_ind] :] = True
        self._dla_model_posterior_ind = ind
        return ind

    @property
    def model_posteriors(self):
        return np.exp(self.log_posteriors - logsumexp(self.log_posteriors))

    @property
    def model_evidences(self):
        return np.exp(self.log_likelihoods - logsumexp(self.log_likelihoods))

    @property
    def model_priors(self):
        return np.exp(self.log_priors - logsumexp(self.log_priors))

    @property
    def p_dla(self):
        self._p_dla = np.sum(self.model_posteriors[self.dla_model_posterior_ind])
        return self._p_dla

    @property
    def p_no_dla(self):
        return 1 - self.p_dla
