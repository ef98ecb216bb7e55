"""Shared builders for the test-suite: seeded synthetic inputs and the drop-in model objects."""
import glob
import os

import numpy as np

from gpy_dla_detection_b200 import synthetic
from gpy_dla_detection_b200.set_parameters import Parameters

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def small_spectrum_fixtures():
    return sorted(f for f in glob.glob(os.path.join(GOLDEN, "spec_S*.npz")) if "S10000" not in f)


def fixture_id(path):
    return os.path.basename(path)[:-4]


def model_args(model):
    return (model["rest_wavelengths"], model["mu"], model["M"], model["log_omega"], model["log_c_0"],
            model["log_tau_0"], model["log_beta"])


class Setup:
    """Parameters + learned model + prior + sample arrays for S samples."""

    def __init__(self, S, num_lines=3):
        self.params = Parameters(num_dla_samples=S, num_lines=num_lines)
        self.model = synthetic.make_learned_model(0)
        self.prior = synthetic.SyntheticPrior(self.params)
        self.dla = synthetic.make_dla_sample_arrays(self.params)
        self.sub = synthetic.make_subdla_sample_arrays(self.params)

    def sample_objects(self):
        from gpy_dla_detection_b200.dla_samples import DLASamplesArrays
        from gpy_dla_detection_b200.subdla_samples import SubDLASamplesArrays

        d = DLASamplesArrays(self.params, self.prior, self.dla["offset_samples"], self.dla["log_nhi_samples"],
                             self.dla["nhi_samples"])
        s = SubDLASamplesArrays(self.params, self.prior, self.sub["offset_samples"], self.sub["log_nhi_samples"],
                                self.sub["nhi_samples"], self.sub["Z_lls"], self.sub["Z_dla"])
        return d, s

    def gp_objects(self, broadening=True):
        """(NullGP, SubDLAGP, DLAGP) built exactly as run_bayes_select.py:151-183 builds them."""
        from gpy_dla_detection_b200.null_gp import NullGP
        from gpy_dla_detection_b200.dla_gp import DLAGP
        from gpy_dla_detection_b200.subdla_gp import SubDLAGP

        d, s = self.sample_objects()
        margs = model_args(self.model)
        gp = NullGP(self.params, self.prior, *margs)
        sub = SubDLAGP(self.params, self.prior, s, *margs, broadening=broadening)
        dla = DLAGP(self.params, self.prior, d, *margs, broadening=broadening)
        return gp, sub, dla

    def catalogue(self, max_dlas=4, broadening=True, batch_spectra=8):
        from gpy_dla_detection_b200.run_bayes_select import CatalogueProcessor

        d, s = self.sample_objects()
        return CatalogueProcessor(self.params, self.prior, self.model, d, s, max_dlas, broadening,
                                  batch_spectra=batch_spectra)


def rel_err(a, b, floor=1e-300):
    """max |a-b| / max(|b|, floor) over the non-NaN entries (NaN patterns are compared separately)."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    ok = ~np.isnan(b)
    if not np.any(ok):
        return 0.0
    return float(np.max(np.abs(a[ok] - b[ok]) / np.maximum(np.abs(b[ok]), floor)))


def ll_err(a, b):
    """
    Error measure of the per-sample log-likelihood tolerance (north_star: 1e-9 relative):
    |a-b| / max(|b|, 1).  A log-likelihood is -(quad + logdet + n log 2 pi)/2, a sum of O(1e3)
    terms that can cancel to |ll| << 1 for a few of the 40 000 samples; there a pure relative
    measure is ill-conditioned for ANY float64 implementation - the NumPy restatement and the
    live reference (same LAPACK, different GEMM blocking) already differ by 1.2e-8 relative
    (9e-12 absolute) at ll = -7.7e-4 of the S = 10 000 golden spectrum - so entries with
    |ll| < 1 are held to 1e-9 absolute instead.
    """
    return rel_err(a, b, floor=1.0)
