"""
_absorber_gp.py : shared machinery of the DLA and subDLA GP models.

The reference keeps two near-identical copies of the QMC loop (dla_gp.py:92-225 and
subdla_gp.py:90-222).  Here both models run the same device path: unique single-absorber
profiles -> per-level batched likelihoods -> evidence / separation mask / resampling, all
inside `dla_log_model_evidences` of the C-ABI.  Only the sampler method, the stored
attributes and the priors differ between the two public classes.
"""
import ctypes
from typing import Tuple

import numpy as np

from . import _lib
from .null_gp import NullGP


class AbsorberGP(NullGP):
    """NullGP + k intervening absorbers parameterised by (z, N_HI); not used directly."""

    def _init_absorber(self, dla_samples, min_z_separation: float, broadening: bool) -> None:
        self.min_z_separation = self.params.kms_to_z(min_z_separation)
        self.dla_samples = dla_samples
        self.broadening = broadening

    # subclasses say where the z samples come from
    def _sample_z(self) -> np.ndarray:
        raise NotImplementedError

    def _log_model_evidences(self, max_dlas: int) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
        S = int(self.params.num_dla_samples)
        sample_z = _lib.f64(self._sample_z())
        nhi = _lib.f64(self.dla_samples.nhi_samples)
        assert sample_z.shape == (S,) and nhi.shape == (S,)  # dla_gp.py:112-119,132

        # The reference draws S uniforms from the global MT19937 stream per resampled level
        # (np.random.choice, dla_gp.py:213).  Draw the same numbers up front and rewind the
        # stream if the level loop exits early, so the global RNG ends in the same state.
        uniforms = None
        rng_state = None
        if max_dlas > 1:
            rng_state = np.random.get_state()
            uniforms = np.random.random_sample((max_dlas - 1, S))

        sample_ll = np.empty((S, max_dlas))
        base_inds = np.zeros((max(max_dlas - 1, 0), S), dtype=np.int32)
        log_ev = np.empty((max_dlas,))
        used = ctypes.c_int(0)
        _lib.check(
            _lib.load_library().dla_log_model_evidences(
                self._spectrum.ptr, _lib.dptr(sample_z), _lib.dptr(nhi), S, int(max_dlas),
                _lib.dptr(uniforms) if uniforms is not None else None, float(self.min_z_separation),
                int(self.params.num_lines), _lib.dptr(sample_ll),
                _lib.iptr(base_inds) if max_dlas > 1 else None, _lib.dptr(log_ev), ctypes.byref(used),
            )
        )
        if max_dlas > 1 and used.value < max_dlas - 1:
            print(
                "Finish the loop earlier because NaN value in log p(D | z_QSO, {} DLAs)".format(used.value)
            )
            np.random.set_state(rng_state)
            if used.value > 0:
                np.random.random_sample(used.value * S)
        return log_ev, sample_ll, base_inds

    def sample_log_likelihood_k_dlas(self, z_dlas: np.ndarray, nhis: np.ndarray) -> float:
        """
        log p(y | k absorbers at (z_dlas, nhis)) (dla_gp.py:311-329, subdla_gp.py:224-242).
        """
        assert len(z_dlas) == len(nhis)
        return float(self.sample_log_likelihoods_batch(np.asarray(z_dlas)[None, :], np.asarray(nhis)[None, :])[0])

    def sample_log_likelihoods_batch(self, z_dlas: np.ndarray, nhis: np.ndarray) -> np.ndarray:
        """Vectorised `sample_log_likelihood_k_dlas`: z_dlas, nhis of shape (S, k_dlas) -> (S,)."""
        zz, nn = _lib.f64(z_dlas), _lib.f64(nhis)
        assert zz.shape == nn.shape and zz.ndim == 2
        out = np.empty((zz.shape[0],))
        _lib.check(
            _lib.load_library().dla_sample_log_likelihoods(
                self._spectrum.ptr, _lib.dptr(zz), _lib.dptr(nn), zz.shape[0], zz.shape[1],
                int(self.params.num_lines), _lib.dptr(out),
            )
        )
        return out

    def this_dla_gp(self, z_dlas: np.ndarray, nhis: np.ndarray) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
        """
        (dla_mu, dla_M, dla_omega2) with k absorbers applied (dla_gp.py:331-396,
        subdla_gp.py:244-309).  The absorption product (profiles, instrument convolution,
        pixel mask) is evaluated on the device; the three scalings of the already-fetched
        model attributes are formed on the host for the caller (plotting / MCMC read them).
        """
        assert len(z_dlas) == len(nhis)
        zz, nn = _lib.f64(z_dlas), _lib.f64(nhis)
        absorption = np.empty((self.this_mu.shape[0],))
        _lib.check(
            _lib.load_library().dla_absorption_k_dlas(
                self._spectrum.ptr, _lib.dptr(zz), _lib.dptr(nn), zz.shape[0], int(self.params.num_lines),
                _lib.dptr(absorption),
            )
        )
        assert len(absorption) == len(self.this_mu)
        return self.this_mu * absorption, self.this_M * absorption[:, None], self.this_omega2 * absorption**2
