"""
ref_loader.py : import the UNMODIFIED reference package from /root/reference in this
container.  TEST INFRASTRUCTURE ONLY - used by tests/golden/make_golden.py to generate the
committed golden vectors and by the (container-only) live-reference tests.  /root/reference
does not exist on the GPU box; nothing that runs there may call this.

`h5py` and `emcee` are absent from the image and are only used by the reference for .mat
I/O and MCMC, so empty stand-in modules are registered before the import.
"""
import os
import sys
import types

REFERENCE_ROOT = "/root/reference"


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "gpy_dla_detection"))


def load_reference():
    """Returns the reference's `gpy_dla_detection` package (modules imported lazily by caller)."""
    if not reference_available():
        raise RuntimeError("reference tree not present at " + REFERENCE_ROOT)
    if "h5py" not in sys.modules:
        sys.modules["h5py"] = types.ModuleType("h5py")
    if "emcee" not in sys.modules:
        emcee = types.ModuleType("emcee")
        emcee.EnsembleSampler = type("EnsembleSampler", (), {})
        sys.modules["emcee"] = emcee
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
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
