"""
zqso_gp.py : GP model for quasar-redshift estimation, device-resident.

Drop-in for the reference's ZGP (zqso_gp.py:14-281): same constructor, `set_data`, `get_interp`,
`log_model_evidence`, `inference_z_qso`, `log_mvnpdf_iid`, and the attributes the reference leaves
behind (`x, y, v, this_wavelengths, pixel_mask, ind, this_mu, this_M, y_bw, v_bw, y_rw, v_rw, z_qso,
sample_log_likelihoods, z_map`).  The 10 000-sample redshift sweep of `inference_z_qso` is ONE call of
`dla_zqso_inference` (one warp per candidate redshift, csrc/zqso_kernel.cuh); `inference_z_qsos`
does the same for a list of spectra.
"""
import ctypes
from typing import Dict, Sequence, Tuple

import numpy as np

from . import _lib
from .null_gp import NullGP, _Handle
from .zqso_samples import ZSamples
from .zqso_set_parameters import ZParameters


class ZGP(NullGP):
    """
    p(y | lambda, sigma^2, M, blue_sigma^2, red_sigma^2) with i.i.d. Gaussians outside the modelling window
    (zqso_gp.py:14-64, arXiv:2006.07343).
    """

    def __init__(
        self,
        params: ZParameters,
        z_qso_samples: ZSamples,
        rest_wavelengths: np.ndarray,
        mu: np.ndarray,
        M: np.ndarray,
        bluewards_mu: float,
        redwards_mu: float,
        bluewards_sigma: float,
        redwards_sigma: float,
    ):
        # like the reference, NullGP.__init__ is not run: a ZGP has no omega / mean-flux parameters
        self.params = params
        self.z_qso_samples = z_qso_samples
        self.rest_wavelengths = _lib.f64(rest_wavelengths)
        self.mu = _lib.f64(mu)
        self.M = _lib.f64(M)
        self.bluewards_mu = float(bluewards_mu)
        self.redwards_mu = float(redwards_mu)
        self.bluewards_sigma = float(bluewards_sigma)
        self.redwards_sigma = float(redwards_sigma)
        assert self.M.shape == (self.rest_wavelengths.shape[0], self.params.k)
        ptr = ctypes.c_void_p()
        _lib.check(
            _lib.load_library().dla_zqso_model_create(
                _lib.dptr(self.rest_wavelengths), _lib.dptr(self.mu), _lib.dptr(self.M), self.rest_wavelengths.shape[0],
                self.M.shape[1], self.bluewards_mu, self.redwards_mu, self.bluewards_sigma, self.redwards_sigma,
                ctypes.byref(ptr),
            )
        )
        self._zmodel = _Handle(ptr, "dla_zqso_model_destroy")

    def _zparams(self) -> "_lib.ZqsoParamsStruct":
        p = self.params
        return _lib.ZqsoParamsStruct(p.min_lambda, p.max_lambda, p.normalization_min_lambda, p.normalization_max_lambda)

    # -- one redshift: the reference's attribute-setting interface ------------------------------------
    def set_data(
        self,
        X: np.ndarray,
        Y: np.ndarray,
        noise_variance: np.ndarray,
        pixel_mask: np.ndarray,
        z_qso: float,
        normalize: bool = True,
        build_model: bool = True,
    ) -> None:
        """
        Select, normalise and (build_model) interpolate at one candidate redshift (zqso_gp.py:92-182).
        X are OBSERVED wavelengths.  The arithmetic (rest-frame conversion, nanmedian, normalisation,
        interpolation) runs on the device (`dla_zqso_set_data`); the boolean selections are NumPy views.
        """
        if not normalize:
            raise NameError("name 'this_median' is not defined")  # the reference fails the same way (:153-157)
        X, Y, V = _lib.f64(X), _lib.f64(Y), _lib.f64(noise_variance)
        mask = np.asarray(pixel_mask).astype(bool)
        n_raw = X.shape[0]
        x = np.empty(n_raw)
        yn = np.empty(n_raw)
        vn = np.empty(n_raw)
        mu = np.empty(n_raw)
        M = np.empty((n_raw, self.params.k))
        cls = np.empty(n_raw, dtype=np.uint8)
        inw = np.empty(n_raw, dtype=np.uint8)
        med = ctypes.c_double()
        zp = self._zparams()
        mask_u8 = _lib.u8(mask)
        _lib.check(
            _lib.load_library().dla_zqso_set_data(
                self._zmodel.ptr, ctypes.byref(zp), _lib.dptr(X), _lib.dptr(Y), _lib.dptr(V), _lib.bptr(mask_u8), n_raw,
                float(z_qso), _lib.dptr(x), _lib.dptr(yn), _lib.dptr(vn), _lib.dptr(mu), _lib.dptr(M), _lib.bptr(cls),
                _lib.bptr(inw), ctypes.byref(med),
            )
        )
        self.z_qso = z_qso
        in_window = inw.astype(bool)
        sel = cls == 1
        self.pixel_mask = mask[in_window]
        self.ind = sel[in_window]                      # the second `ind` of the reference (:169-181)
        self.this_wavelengths = X[sel]
        self.x = x[sel]
        self.y = yn[sel]
        self.v = vn[sel]
        if np.any(np.isinf(self.v)):
            self.v[np.isinf(self.v)] = np.nanmean(self.v)  # the reference's kludge (:177), a no-op for +inf
        self.y_bw, self.v_bw = yn[cls == 2], vn[cls == 2]
        self.y_rw, self.v_rw = yn[cls == 3], vn[cls == 3]
        self.this_median = med.value
        if build_model:
            self.this_mu = mu[sel]
            self.this_M = np.ascontiguousarray(M[sel])
            assert self.this_M.shape[1] == self.params.k

    def get_interp(self, x: np.ndarray, y: np.ndarray, wavelengths: np.ndarray, z_qso: float) -> None:
        """Interpolate mu and M at rest wavelengths x (zqso_gp.py:66-90); done inside `set_data` on the device."""
        raise NotImplementedError("ZGP.get_interp is folded into set_data(build_model=True) on the device")

    def log_model_evidence(self) -> float:
        """Low-rank Gaussian inside the window + two i.i.d. Gaussians outside (zqso_gp.py:184-212)."""
        log_likelihood = self.log_mvnpdf_low_rank(self.y, self.this_mu, self.this_M, self.v)
        n_bw, n_rw = self.y_bw.shape[0], self.y_rw.shape[0]
        bw = self.log_mvnpdf_iid(self.y_bw, self.bluewards_mu * np.ones((n_bw,)),
                                 self.bluewards_sigma ** 2 * np.ones((n_bw,)) + self.v_bw)
        rw = self.log_mvnpdf_iid(self.y_rw, self.redwards_mu * np.ones((n_rw,)),
                                 self.redwards_sigma ** 2 * np.ones((n_rw,)) + self.v_rw)
        return log_likelihood + bw + rw

    @staticmethod
    def log_mvnpdf_iid(y: np.ndarray, mu: np.ndarray, d: np.ndarray) -> float:
        """log N(y; mu, diag(d)) (zqso_gp.py:252-278)."""
        y, mu, d = _lib.f64(y), _lib.f64(mu), _lib.f64(d)
        assert y.shape == mu.shape == d.shape
        out = np.empty(1)
        _lib.check(_lib.load_library().dla_log_mvnpdf_iid(_lib.dptr(y), _lib.dptr(mu), _lib.dptr(d), y.shape[0],
                                                          _lib.dptr(out)))
        return float(out[0])

    # -- the redshift sweep -------------------------------------------------------------------------------
    def inference_z_qso(
        self,
        wavelengths: np.ndarray,
        flux: np.ndarray,
        noise_variance: np.ndarray,
        pixel_mask: np.ndarray,
        z_qso_min: float = 2.14,
        z_qso_max: float = 6.16,
    ):
        """Sample log-likelihoods over the prior volume and the MAP redshift (zqso_gp.py:214-250)."""
        sample_z_qsos = self.z_qso_samples.sample_z_qsos(z_qso_min=z_qso_min, z_qso_max=z_qso_max)
        out = self.inference_z_qsos([(wavelengths, flux, noise_variance, pixel_mask)], sample_z_qsos)
        self.sample_log_likelihoods = out["sample_log_likelihoods"][0]
        if out["map_index"][0] < 0:
            raise ValueError("All-NaN slice encountered")  # np.nanargmax (:248)
        self.z_map = float(out["z_map"][0])
        print("[Info] Z MAP = {:.3g}".format(self.z_map))

    @staticmethod
    def pack(spectra: Sequence[Tuple[np.ndarray, np.ndarray, np.ndarray, np.ndarray]]):
        """List of (wavelengths, flux, noise_variance, pixel_mask) -> (offsets, wl, fl, nv, pm) ragged contiguous arrays."""
        lengths = np.array([len(sp[0]) for sp in spectra], dtype=np.int64)
        offsets = np.concatenate([[0], np.cumsum(lengths)]).astype(np.int64)
        wl = np.concatenate([_lib.f64(sp[0]) for sp in spectra])
        fl = np.concatenate([_lib.f64(sp[1]) for sp in spectra])
        nv = np.concatenate([_lib.f64(sp[2]) for sp in spectra])
        pm = np.concatenate([_lib.u8(sp[3]) for sp in spectra])
        return offsets, wl, fl, nv, pm

    def inference_z_qsos(self, spectra: Sequence[Tuple[np.ndarray, np.ndarray, np.ndarray, np.ndarray]],
                         sample_z_qsos: np.ndarray, keep_samples: bool = True) -> Dict[str, np.ndarray]:
        """
        The sweep for a list of spectra (wavelengths, flux, noise_variance, pixel_mask) in one call:
        `z_map` (Q,), `map_index` (Q,) and, when keep_samples, `sample_log_likelihoods` (Q, S).
        """
        return self.inference_packed(self.pack(spectra), sample_z_qsos, keep_samples)

    def inference_packed(self, packed, sample_z_qsos: np.ndarray, keep_samples: bool = True) -> Dict[str, np.ndarray]:
        """The sweep on ragged arrays as `pack` (or a preload.PreloadedSpectra chunk) lays them out."""
        offsets, wl, fl, nv, pm = packed
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        zs = _lib.f64(sample_z_qsos)
        Q, S = offsets.shape[0] - 1, zs.shape[0]
        ll = np.empty((Q, S)) if keep_samples else None
        z_map = np.empty(Q)
        map_index = np.empty(Q, dtype=np.int32)
        zp = self._zparams()
        _lib.check(
            _lib.load_library().dla_zqso_inference(
                self._zmodel.ptr, ctypes.byref(zp), Q, offsets.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)),
                _lib.dptr(wl), _lib.dptr(fl), _lib.dptr(nv), _lib.bptr(pm), _lib.dptr(zs), S,
                _lib.dptr(ll) if ll is not None else None, _lib.dptr(z_map), _lib.iptr(map_index),
            )
        )
        out = dict(z_map=z_map, map_index=map_index)
        if keep_samples:
            out["sample_log_likelihoods"] = ll
        return out

    @staticmethod
    def last_timing() -> Dict[str, float]:
        """GPU milliseconds of the last sweep: the kernels alone (spectra resident in HBM) and the whole call."""
        k, t = ctypes.c_double(), ctypes.c_double()
        _lib.check(_lib.load_library().dla_zqso_last_timing(ctypes.byref(k), ctypes.byref(t)))
        return dict(kernel_ms=k.value, total_ms=t.value)

    @property
    def this_noise(self):
        """noise kernel: instrumental noise (zqso_gp.py:280-285)"""
        return self.v


class ZGPMAT(ZGP):
    """Load the learned model from the reference's .mat file (zqso_gp.py:283-319); needs h5py."""

    def __init__(self, params: ZParameters, z_qso_samples: ZSamples,
                 learned_file: str = "learned_zqso_only_model_outdata_full_dr9q_minus_concordance_norm_1176-1256.mat"):
        import h5py  # not part of this image; only needed for the published .mat model

        with h5py.File(learned_file, "r") as learned:
            rest_wavelengths = learned["rest_wavelengths"][:, 0]
            mu = learned["mu"][:, 0]
            M = learned["M"][()].T
            bluewards_mu = learned["bluewards_mu"][0, 0]
            redwards_mu = learned["redwards_mu"][0, 0]
            bluewards_sigma = learned["bluewards_sigma"][0, 0]
            redwards_sigma = learned["redwards_sigma"][0, 0]
        super().__init__(params, z_qso_samples, rest_wavelengths, mu, M, bluewards_mu, redwards_mu, bluewards_sigma,
                         redwards_sigma)
