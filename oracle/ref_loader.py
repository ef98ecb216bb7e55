"""
ref_loader.py : import the UNMODIFIED reference package - from /root/reference in the build
container, else from oracle/_ref/ (the git-ignored copy staged by oracle/make_ref.sh, which
travels to the GPU box with the snapshot).  TEST / BENCH INFRASTRUCTURE ONLY: used by
tests/golden/make_golden.py to generate the committed golden vectors, by the live-reference
tests, and by bench.py's CPU arm (`--impl reference`, `cpu_baseline`, kind = "reference").
Nothing under gpy_dla_detection_b200/ imports it.

`h5py` and `emcee` are absent from the image and are only used by the reference for .mat
I/O and MCMC, so empty stand-in modules are registered before the import.
"""
import os
import sys
import types

REFERENCE_ROOT = "/root/reference"
STAGED_ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def reference_root():
    """The directory holding the reference's `gpy_dla_detection` package, or None."""
    for root in (REFERENCE_ROOT, STAGED_ROOT):
        if os.path.isfile(os.path.join(root, "gpy_dla_detection", "dla_gp.py")):
            return root
    return None


def reference_available() -> bool:
    return reference_root() is not None


def load_reference():
    """Returns the reference's `gpy_dla_detection` package (modules imported lazily by caller)."""
    root = reference_root()
    if root is None:
        raise RuntimeError("reference package found neither at %s nor at %s (run oracle/make_ref.sh in the build "
                           "container)" % (REFERENCE_ROOT, STAGED_ROOT))
    if "h5py" not in sys.modules:
        sys.modules["h5py"] = types.ModuleType("h5py")
    if "emcee" not in sys.modules:
        emcee = types.ModuleType("emcee")
        emcee.EnsembleSampler = type("EnsembleSampler", (), {})
        sys.modules["emcee"] = emcee
    if root not in sys.path:
        sys.path.insert(0, root)
    import gpy_dla_detection  # noqa: F401
    from gpy_dla_detection import (  # noqa: F401
        voigt, set_parameters, effective_optical_depth, null_gp, dla_gp, subdla_gp,
        dla_samples, subdla_samples, bayesian_model_selection,
    )
    return gpy_dla_detection


class RefDLASamples:
    """In-memory sample object with the attributes the reference reads (dla_samples.py:67-104)."""

    def __init__(self, params, arrays, sub=False):
        self.params = params
        self._offset_samples = arrays["offset_samples"]
        self._log_nhi_samples = arrays["log_nhi_samples"]
        self._nhi_samples = arrays["nhi_samples"]
        if sub:
            self._Z_lls = arrays["Z_lls"]
            self._Z_dla = arrays["Z_dla"]

    offset_samples = property(lambda self: self._offset_samples)
    log_nhi_samples = property(lambda self: self._log_nhi_samples)
    nhi_samples = property(lambda self: self._nhi_samples)

    def _sample(self, wavelengths, z_qso):
        lo = self.params.min_z_dla(wavelengths, z_qso)
        hi = self.params.max_z_dla(wavelengths, z_qso)
        return lo + (hi - lo) * self._offset_samples

    sample_z_dlas = _sample
    sample_z_lls = _sample


def run_reference_spectrum(model, dla, sub, prior, spectrum, z_qso, S, max_dlas=4, num_lines=3, broadening=True):
    """
    One spectrum through the reference's own classes exactly as run_bayes_select.py:141-230 drives them:
    seed, three set_data calls, BayesModelSelect.model_selection, maximum_a_posteriori.  `model`, `dla`, `sub`
    are the synthetic arrays of gpy_dla_detection_b200.synthetic; `prior` needs `less_ind`.
    """
    import numpy as np

    load_reference()
    from gpy_dla_detection.set_parameters import Parameters as RParameters
    from gpy_dla_detection.null_gp import NullGP as RNullGP
    from gpy_dla_detection.dla_gp import DLAGP as RDLAGP
    from gpy_dla_detection.subdla_gp import SubDLAGP as RSubDLAGP
    from gpy_dla_detection.bayesian_model_selection import BayesModelSelect as RBayes

    rp = RParameters(num_dla_samples=S, num_lines=num_lines)
    wl, fl, nv, pm = spectrum
    rest = rp.emitted_wavelengths(wl, z_qso)
    margs = (model["rest_wavelengths"], model["mu"], model["M"], model["log_omega"], model["log_c_0"],
             model["log_tau_0"], model["log_beta"])
    gp = RNullGP(rp, prior, *margs)
    dgp = RDLAGP(rp, prior, RefDLASamples(rp, dla), *margs, broadening=broadening)
    sgp = RSubDLAGP(rp, prior, RefDLASamples(rp, sub, True), *margs, broadening=broadening)
    for m in (gp, dgp, sgp):
        m.set_data(rest, fl, nv, pm, z_qso, build_model=True)
    np.random.seed(0)  # run_bayes_select.py:144
    bayes = RBayes([0, 1, max_dlas], 2)
    log_post = bayes.model_selection([gp, sgp, dgp], z_qso)
    try:
        map_z, map_n = dgp.maximum_a_posteriori()
    except ValueError:
        map_z = map_n = np.full((max_dlas, max_dlas), np.nan)
    return dict(log_priors=bayes.log_priors, log_likelihoods=bayes.log_likelihoods, log_posteriors=log_post,
                model_posteriors=bayes.model_posteriors, p_dla=bayes.p_dla, p_no_dla=bayes.p_no_dla,
                sample_log_likelihoods_dla=dgp.sample_log_likelihoods, base_sample_inds=dgp.base_sample_inds,
                sample_log_likelihoods_lls=sgp.sample_log_likelihoods[:, 0], MAP_z_dlas=map_z, MAP_log_nhis=map_n)


def run_reference_zqso(model, spectrum, num_zqso_samples=10000):
    """ZGP.inference_z_qso of the reference on one spectrum (zqso_gp.py:214-250); returns (sample ll, z_map)."""
    load_reference()
    from gpy_dla_detection.zqso_gp import ZGP as RZGP
    from gpy_dla_detection.zqso_set_parameters import ZParameters as RZParameters
    from gpy_dla_detection.zqso_samples import ZSamples as RZSamples

    rp = RZParameters(num_zqso_samples=num_zqso_samples)
    gp = RZGP(rp, RZSamples(rp), model["rest_wavelengths"], model["mu"], model["M"], model["bluewards_mu"],
              model["redwards_mu"], model["bluewards_sigma"], model["redwards_sigma"])
    wl, fl, nv, pm = spectrum
    gp.inference_z_qso(wl, fl, nv, pm)
    return gp.sample_log_likelihoods, gp.z_map
