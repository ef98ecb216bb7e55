"""
interop.py : run the REFERENCE's own classes on the device.

The reference's BayesModelSelect.model_selection type-checks its models
(bayesian_model_selection.py:58-61: `isinstance(model_list[-1], DLAGP)` against ITS DLAGP), so the
classes of this package - same names and attributes, but not subclasses of the reference's - cannot
be handed to it.  A maintainer who wants to keep every reference object (constructors, .mat loaders,
`BayesModelSelect`, plotting, their own subclasses) and only move the arithmetic to the GPU calls

    import gpy_dla_detection                      # the reference package
    from gpy_dla_detection_b200 import interop
    undo = interop.patch_reference()              # installs the device methods on the reference's classes
    ...                                           # NullGPMAT / DLAGPMAT / SubDLAGPMAT / BayesModelSelect as before
    undo()                                        # optional: restore the NumPy methods

Installed (the hot path of SURVEY.md §8a; everything else of the reference is untouched):
  NullGP   : set_data, get_interp, log_model_evidence, log_mvnpdf_low_rank
  DLAGP    : log_model_evidences, sample_log_likelihood_k_dlas, this_dla_gp      (+ SubDLAGP)
  voigt.voigt_absorption, effective_optical_depth.effective_optical_depth (module functions, also
  re-bound in the modules that imported them by name)
Objects built by the reference's constructors carry the attributes these methods read (`params`, `prior`,
`rest_wavelengths`, `mu`, `M`, `log_omega`, `log_c_0`, ..., `dla_samples`, `min_z_separation`, `broadening`);
the device handles are created lazily on first use.
"""
import inspect
from typing import Callable

_NULL_METHODS = ("set_data", "get_interp", "log_model_evidence", "log_mvnpdf_low_rank", "_fetch_attributes",
                 "_params_struct", "_model_handle", "_rebuild_prepared")
_ABSORBER_METHODS = ("log_model_evidences", "sample_log_likelihood_k_dlas", "sample_log_likelihoods_batch",
                     "this_dla_gp", "_log_model_evidences", "_sample_z")
_MISSING = object()


def patch_reference(pkg=None) -> Callable[[], None]:
    """Install the device methods on the reference's classes; returns a function that restores the originals."""
    if pkg is None:
        import gpy_dla_detection as pkg  # noqa: F811  (the reference must be importable)
    import importlib

    from . import dla_gp, effective_optical_depth, null_gp, subdla_gp, voigt

    name = pkg.__name__
    r_null = importlib.import_module(name + ".null_gp")
    r_dla = importlib.import_module(name + ".dla_gp")
    r_sub = importlib.import_module(name + ".subdla_gp")
    r_voigt = importlib.import_module(name + ".voigt")
    r_eod = importlib.import_module(name + ".effective_optical_depth")

    saved = []

    def install(target, attr, value):
        saved.append((target, attr, target.__dict__.get(attr, _MISSING)))
        setattr(target, attr, value)

    for ref_cls, our_cls, names in ((r_null.NullGP, null_gp.NullGP, _NULL_METHODS),
                                    (r_dla.DLAGP, dla_gp.DLAGP, _ABSORBER_METHODS),
                                    (r_sub.SubDLAGP, subdla_gp.SubDLAGP, _ABSORBER_METHODS)):
        for meth in names:
            install(ref_cls, meth, inspect.getattr_static(our_cls, meth))  # keeps staticmethod objects intact
    install(r_voigt, "voigt_absorption", voigt.voigt_absorption)
    install(r_eod, "effective_optical_depth", effective_optical_depth.effective_optical_depth)
    for mod in (r_dla, r_sub):
        if "voigt_absorption" in mod.__dict__:
            install(mod, "voigt_absorption", voigt.voigt_absorption)
    if "effective_optical_depth" in r_null.__dict__:
        install(r_null, "effective_optical_depth", effective_optical_depth.effective_optical_depth)

    def undo() -> None:
        while saved:
            target, attr, old = saved.pop()
            if old is _MISSING:
                delattr(target, attr)
            else:
                setattr(target, attr, old)

    return undo
