"""
null_gp.py : the null (no absorber) GP model, device-resident.

Drop-in for the reference's NullGP (null_gp.py:15-387): same constructor, `set_data`,
`get_interp`, `log_model_evidence`, `log_mvnpdf_low_rank`, `log_prior` and the same
attributes after `set_data` (`x, y, v, ind_unmasked, ind, unmasked_wavelengths,
this_wavelengths, padded_wavelengths, this_mu, this_M, this_omega2, normalization_median`).
The learned model and the prepared spectrum live on the GPU behind opaque handles of the
C-ABI; NumPy copies of the attributes are fetched once after preparation.
"""
import ctypes

import numpy as np

from . import _lib
from .set_parameters import Parameters


class _Handle:
    """Owns one C-ABI handle and frees it with the matching destroy function."""

    def __init__(self, ptr: ctypes.c_void_p, destroy_name: str):
        self.ptr = ptr
        self._destroy_name = destroy_name

    def __del__(self):
        try:
            if self.ptr:
                getattr(_lib.load_library(), self._destroy_name)(self.ptr)
                self.ptr = None
        except Exception:
            pass


class NullGP:
    """
    p(y | lambda, sigma^2, M, omega, c_0, tau_0, beta, tau_kim, beta_kim)   (null_gp.py:15-71)
    """

    def __init__(
        self,
        params: Parameters,
        prior,
        rest_wavelengths: np.ndarray,
        mu: np.ndarray,
        M: np.ndarray,
        log_omega: np.ndarray,
        log_c_0: float,
        log_tau_0: float,
        log_beta: float,
        prev_tau_0: float = 0.0023,
        prev_beta: float = 3.65,
    ):
        self.params = params
        self.prior = prior

        self.rest_wavelengths = _lib.f64(rest_wavelengths)
        self.mu = _lib.f64(mu)
        self.M = _lib.f64(M)
        self.log_omega = _lib.f64(log_omega)
        self.log_c_0 = float(log_c_0)
        self.log_tau_0 = float(log_tau_0)
        self.log_beta = float(log_beta)
        self.prev_tau_0 = float(prev_tau_0)
        self.prev_beta = float(prev_beta)

        assert self.M.shape == (self.rest_wavelengths.shape[0], self.params.k)

        self._model_handle()
        self._spectrum = None
        self.broadening = True

    def _model_handle(self) -> "_Handle":
        """
        The learned model on the device, uploaded on first use.  (Lazy so that the methods of this class can be
        installed on objects built by the REFERENCE's constructors - interop.patch_reference.)
        """
        h = self.__dict__.get("_model")
        if h is None:
            rest, mu = _lib.f64(self.rest_wavelengths), _lib.f64(self.mu)
            M, lo = _lib.f64(self.M), _lib.f64(self.log_omega)
            ptr = ctypes.c_void_p()
            _lib.check(
                _lib.load_library().dla_model_create(
                    _lib.dptr(rest), _lib.dptr(mu), _lib.dptr(M), _lib.dptr(lo), rest.shape[0], M.shape[1],
                    float(self.log_c_0), float(self.log_tau_0), float(self.log_beta), float(self.prev_tau_0),
                    float(self.prev_beta), ctypes.byref(ptr),
                )
            )
            h = self._model = _Handle(ptr, "dla_model_destroy")
        return h

    # -- device-side preparation ---------------------------------------------------------------
    def _params_struct(self) -> "_lib.DLAParamsStruct":
        return _lib.params_struct(
            self.params, getattr(self, "broadening", True), getattr(self, "min_z_separation", 0.0)
        )

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
        Load one spectrum (null_gp.py:95-177): X rest wavelengths, Y flux, noise variance and
        pixel mask of all raw pixels.  Normalisation, range/mask filtering, interpolation of
        the learned model and the mean-flux suppression run in one device kernel.
        """
        if self.params.min_lambda < self.rest_wavelengths[0] or self.params.max_lambda > self.rest_wavelengths[-1]:
            # scipy.interpolate.interp1d raises for out-of-range x (null_gp.py:67-71)
            raise ValueError("modelling range exceeds the learned model's rest wavelength grid")
        X = _lib.f64(X)
        Y = _lib.f64(Y)
        V = _lib.f64(noise_variance)
        mask = _lib.u8(pixel_mask)
        assert X.shape == Y.shape == V.shape == mask.shape

        self.pixel_mask = np.asarray(pixel_mask).astype(np.bool_)
        self.z_qso = z_qso

        ptr = ctypes.c_void_p()
        ps = self._params_struct()
        _lib.check(
            _lib.load_library().dla_spectrum_create(
                self._model_handle().ptr, ctypes.byref(ps), _lib.dptr(X), _lib.dptr(Y), _lib.dptr(V), _lib.bptr(mask),
                X.shape[0], float(z_qso), 1 if normalize else 0, ctypes.byref(ptr),
            )
        )
        self._spectrum = _Handle(ptr, "dla_spectrum_destroy")
        self._fetch_attributes(X.shape[0], normalize, build_model)

    def _fetch_attributes(self, n_raw: int, normalize: bool, build_model: bool) -> None:
        lib = _lib.load_library()
        n_u, n = ctypes.c_int(), ctypes.c_int()
        _lib.check(lib.dla_spectrum_sizes(self._spectrum.ptr, None, ctypes.byref(n_u), ctypes.byref(n)))
        n_u, n = n_u.value, n.value
        k, w = self.params.k, self.params.width
        x, y, v, twl = np.empty(n), np.empty(n), np.empty(n), np.empty(n)
        mu, M, om = np.empty(n), np.empty((n, k)), np.empty(n)
        uwl, pwl = np.empty(n_u), np.empty(n_u + 2 * w)
        iu, ind = np.empty(n_raw, dtype=np.uint8), np.empty(n_raw, dtype=np.uint8)
        med = ctypes.c_double()
        _lib.check(
            lib.dla_spectrum_get(
                self._spectrum.ptr, _lib.dptr(x), _lib.dptr(y), _lib.dptr(v), _lib.dptr(twl), _lib.dptr(mu),
                _lib.dptr(M), _lib.dptr(om), _lib.dptr(uwl), _lib.dptr(pwl), _lib.bptr(iu), _lib.bptr(ind),
                ctypes.byref(med),
            )
        )
        self.x, self.y, self.v = x, y, v
        self.this_wavelengths = twl
        self.unmasked_wavelengths = uwl
        self.padded_wavelengths = pwl
        self.ind_unmasked = iu.astype(np.bool_)
        self.ind = ind.astype(np.bool_)
        if normalize:
            self.normalization_median = med.value
        if build_model:
            self.this_mu, self.this_M, self.this_omega2 = mu, M, om

    def get_interp(self, x: np.ndarray, y: np.ndarray, wavelengths: np.ndarray, z_qso: float) -> None:
        """
        Build and interpolate the GP model onto the given pixels (null_gp.py:179-242): x rest wavelengths,
        wavelengths observed (y is unused, as in the reference).  Sets `this_mu`, `this_M`, `this_omega2`.
        `set_data(build_model=True)` already does this for the loaded spectrum inside its own kernel; called
        with any other grid (or after `build_model=False`) the interpolation runs through `dla_model_interp`,
        and - when the grid has the loaded spectrum's length - the device copy the likelihoods use is rebuilt
        from the new arrays, so the evidences follow the attributes exactly as they do in the reference.
        """
        x, wl = _lib.f64(x), _lib.f64(wavelengths)
        assert x.shape == wl.shape and x.ndim == 1
        if x.shape[0] and (np.min(x) < self.rest_wavelengths[0] or np.max(x) > self.rest_wavelengths[-1]):
            # scipy.interpolate.interp1d raises for out-of-range x (null_gp.py:67-71)
            raise ValueError("A value in x_new is outside the interpolation range.")
        n, k = x.shape[0], self.params.k
        mu, M, om = np.empty(n), np.empty((n, k)), np.empty(n)
        _lib.check(
            _lib.load_library().dla_model_interp(
                self._model_handle().ptr, int(self.params.num_forest_lines), _lib.dptr(x), _lib.dptr(wl), n, float(z_qso),
                _lib.dptr(mu), _lib.dptr(M), _lib.dptr(om),
            )
        )
        same = (getattr(self, "this_mu", None) is not None and self.this_mu.shape == mu.shape
                and np.array_equal(self.this_mu, mu) and np.array_equal(self.this_omega2, om)
                and np.array_equal(self.this_M, M))
        self.this_mu, self.this_M, self.this_omega2 = mu, M, om
        if not same and getattr(self, "_spectrum", None) is not None and hasattr(self, "y") and n == self.y.shape[0]:
            self._rebuild_prepared()

    def _rebuild_prepared(self) -> None:
        """Device spectrum from the object's own attributes (y, v, this_mu, this_M, this_omega2, grids, mask)."""
        broadening = bool(getattr(self, "broadening", True))
        wl = _lib.f64(self.padded_wavelengths if broadening else self.unmasked_wavelengths)
        keep = _lib.u8(~self.pixel_mask[self.ind_unmasked])
        y, v = _lib.f64(self.y), _lib.f64(self.v)
        mu, M, om = _lib.f64(self.this_mu), _lib.f64(self.this_M), _lib.f64(self.this_omega2)
        ptr = ctypes.c_void_p()
        _lib.check(
            _lib.load_library().dla_spectrum_create_prepared(
                _lib.dptr(y), _lib.dptr(v), _lib.dptr(mu), _lib.dptr(M), _lib.dptr(om), y.shape[0], M.shape[1],
                _lib.dptr(wl), wl.shape[0], _lib.bptr(keep), keep.shape[0], 1 if broadening else 0, ctypes.byref(ptr),
            )
        )
        self._spectrum = _Handle(ptr, "dla_spectrum_destroy")

    # -- properties of the reference ---------------------------------------------------------------
    mean = property(lambda self: self.mu)
    K = property(lambda self: np.matmul(self.M, self.M.T))
    this_mean = property(lambda self: self.this_mu)
    this_noise = property(lambda self: self.this_omega2 + self.v)
    this_K = property(lambda self: np.matmul(self.this_M, self.this_M.T))
    X = property(lambda self: self.x)
    Y = property(lambda self: self.y)
    V = property(lambda self: self.v)

    # -- likelihoods ---------------------------------------------------------------------------------
    def log_model_evidence(self) -> float:
        """log p(y | null model) (null_gp.py:294-305), on the device-resident spectrum."""
        out = ctypes.c_double()
        _lib.check(_lib.load_library().dla_null_log_model_evidence(self._spectrum.ptr, ctypes.byref(out)))
        return out.value

    @staticmethod
    def log_mvnpdf_low_rank(
        y: np.ndarray, mu: np.ndarray, M: np.ndarray, d: np.ndarray, scipy_lapack: bool = True
    ) -> float:
        """
        log N(y; mu, M M' + diag(d)) via the Woodbury identity (null_gp.py:307-360);
        `scipy_lapack` is accepted for signature parity and has no effect.
        """
        y, mu, M, d = _lib.f64(y), _lib.f64(mu), _lib.f64(M), _lib.f64(d)
        n, k = M.shape
        assert y.shape == (n,) and mu.shape == (n,) and d.shape == (n,)
        out = ctypes.c_double()
        _lib.check(
            _lib.load_library().dla_log_mvnpdf_low_rank(
                _lib.dptr(y), _lib.dptr(mu), _lib.dptr(M), _lib.dptr(d), n, k, ctypes.byref(out)
            )
        )
        return out.value

    def log_prior(self, z_qso: float, without_subDLAs: bool = True) -> float:
        """P(no DLA | z_QSO) = 1 - M / N without the subDLA share (null_gp.py:362-387)."""
        this_num_dlas, this_num_quasars = self.prior.less_ind(z_qso)
        return np.log(1 - (this_num_dlas / this_num_quasars))
