"""
_lib.py : ctypes binding of libdla_b200.so (include/dla_b200.h).

The library is the only compute path: if it is missing, or no B200 is usable, every call
raises - there is no NumPy fallback anywhere in this package.
"""
import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_int, c_int32, c_int64, c_longlong, c_uint8, c_void_p

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# DLA_B200_LIB: developer override used for A/B runs of kernel variants (tools/); the product path is the in-tree build
LIB_PATH = os.environ.get("DLA_B200_LIB") or os.path.join(_HERE, "libdla_b200.so")

_dp = POINTER(c_double)
_ip = POINTER(c_int32)
_bp = POINTER(c_uint8)


class DLAParamsStruct(ctypes.Structure):
    _fields_ = [
        ("min_lambda", c_double),
        ("max_lambda", c_double),
        ("normalization_min_lambda", c_double),
        ("normalization_max_lambda", c_double),
        ("pixel_spacing", c_double),
        ("width", c_int),
        ("num_forest_lines", c_int),
        ("num_lines", c_int),
        ("broadening", c_int),
        ("lya_wavelength", c_double),
        ("lyman_limit", c_double),
        ("max_z_cut", c_double),
        ("min_z_cut", c_double),
        ("min_z_separation", c_double),
    ]


class ZqsoParamsStruct(ctypes.Structure):
    _fields_ = [
        ("min_lambda", c_double),
        ("max_lambda", c_double),
        ("normalization_min_lambda", c_double),
        ("normalization_max_lambda", c_double),
    ]


class CatalogueConfigStruct(ctypes.Structure):
    _fields_ = [
        ("num_dla_samples", c_int),
        ("max_dlas", c_int),
        ("batch_spectra", c_int),
        ("keep_sample_likelihoods", c_int),
    ]


class CatalogueOutputsStruct(ctypes.Structure):
    _fields_ = [
        ("min_z_dlas", _dp),
        ("max_z_dlas", _dp),
        ("log_priors", _dp),
        ("log_likelihoods", _dp),
        ("log_posteriors", _dp),
        ("model_posteriors", _dp),
        ("p_dlas", _dp),
        ("p_no_dlas", _dp),
        ("MAP_z_dlas", _dp),
        ("MAP_log_nhis", _dp),
        ("sample_log_likelihoods_dla", _dp),
        ("sample_log_likelihoods_lls", _dp),
        ("base_sample_inds", _ip),
        ("num_pixels", _ip),
        ("status", _ip),
    ]


# every symbol include/dla_b200.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "dla_init": (c_int, [c_int]),
    "dla_device_count": (c_int, []),
    "dla_last_error": (c_char_p, []),
    "dla_version": (c_char_p, []),
    "dla_last_kernel_ms": (c_double, []),
    "dla_kernel_launch_count": (c_longlong, []),
    "dla_measure_fp64_peaks": (c_int, [_dp, _dp]),
    "dla_voigt_absorption": (c_int, [_dp, c_int, c_double, c_double, c_int, c_int, _dp]),
    "dla_voigt_absorption_batch": (c_int, [_dp, c_int, _dp, _dp, c_int, c_int, c_int, _dp]),
    "dla_voigt_lls_absorption_batch": (c_int, [_dp, c_int, _dp, _dp, c_int, c_int, c_int, _dp]),
    "dla_spectrum_set_lls_break": (c_int, [c_void_p, c_int]),
    "dla_faddeeva_re": (c_int, [_dp, _dp, c_int, _dp]),
    "dla_effective_optical_depth": (c_int, [_dp, c_int, c_double, c_double, c_double, c_int, _dp]),
    "dla_log_mvnpdf_low_rank": (c_int, [_dp, _dp, _dp, _dp, c_int, c_int, _dp]),
    "dla_model_create": (
        c_int,
        [_dp, _dp, _dp, _dp, c_int, c_int, c_double, c_double, c_double, c_double, c_double, POINTER(c_void_p)],
    ),
    "dla_model_destroy": (c_int, [c_void_p]),
    "dla_model_interp": (c_int, [c_void_p, c_int, _dp, _dp, c_int, c_double, _dp, _dp, _dp]),
    "dla_spectrum_create": (
        c_int,
        [c_void_p, POINTER(DLAParamsStruct), _dp, _dp, _dp, _bp, c_int, c_double, c_int, POINTER(c_void_p)],
    ),
    "dla_spectrum_create_prepared": (
        c_int,
        [_dp, _dp, _dp, _dp, _dp, c_int, c_int, _dp, c_int, _bp, c_int, c_int, POINTER(c_void_p)],
    ),
    "dla_spectrum_destroy": (c_int, [c_void_p]),
    "dla_spectrum_sizes": (c_int, [c_void_p, POINTER(c_int), POINTER(c_int), POINTER(c_int)]),
    "dla_spectrum_get": (c_int, [c_void_p, _dp, _dp, _dp, _dp, _dp, _dp, _dp, _dp, _dp, _bp, _bp, _dp]),
    "dla_null_log_model_evidence": (c_int, [c_void_p, _dp]),
    "dla_sample_log_likelihoods": (c_int, [c_void_p, _dp, _dp, c_int, c_int, c_int, _dp]),
    "dla_absorption_k_dlas": (c_int, [c_void_p, _dp, _dp, c_int, c_int, _dp]),
    "dla_log_model_evidences": (
        c_int,
        [c_void_p, _dp, _dp, c_int, c_int, _dp, c_double, c_int, _dp, _ip, _dp, POINTER(c_int)],
    ),
    "dla_resample_indices": (c_int, [_dp, _dp, c_int, _ip]),
    "dla_catalogue_create": (
        c_int,
        [c_void_p, POINTER(DLAParamsStruct), POINTER(CatalogueConfigStruct), _dp, _dp, _dp, _dp, _dp, _dp, POINTER(c_void_p)],
    ),
    "dla_catalogue_destroy": (c_int, [c_void_p]),
    "dla_catalogue_process": (
        c_int,
        [c_void_p, c_int, POINTER(c_int64), _dp, _dp, _dp, _bp, _dp, _dp, POINTER(CatalogueOutputsStruct)],
    ),
    "dla_catalogue_stage": (c_int, [c_void_p, c_int, POINTER(c_int64), _dp, _dp, _dp, _bp, _dp, _dp]),
    "dla_catalogue_run_staged": (c_int, [c_void_p, POINTER(CatalogueOutputsStruct)]),
    "dla_catalogue_last_timing": (
        c_int,
        [c_void_p, _dp, _dp, _dp, POINTER(c_longlong), _dp],
    ),
    "dla_catalogue_last_counts": (c_int, [c_void_p, POINTER(c_longlong), POINTER(c_longlong)]),
    "dla_zqso_model_create": (
        c_int,
        [_dp, _dp, _dp, c_int, c_int, c_double, c_double, c_double, c_double, POINTER(c_void_p)],
    ),
    "dla_zqso_model_destroy": (c_int, [c_void_p]),
    "dla_zqso_inference": (
        c_int,
        [c_void_p, POINTER(ZqsoParamsStruct), c_int, POINTER(c_int64), _dp, _dp, _dp, _bp, _dp, c_int, _dp, _dp, _ip],
    ),
    "dla_zqso_last_timing": (c_int, [_dp, _dp]),
    "dla_zqso_force_generic_kernel": (c_int, [c_int]),
    "dla_zqso_set_data": (
        c_int,
        [c_void_p, POINTER(ZqsoParamsStruct), _dp, _dp, _dp, _bp, c_int, c_double, _dp, _dp, _dp, _dp, _dp, _bp, _bp, _dp],
    ),
    "dla_log_mvnpdf_iid": (c_int, [_dp, _dp, _dp, c_int, _dp]),
}

_lib = None


class DLALibraryError(RuntimeError):
    pass


def load_library():
    """dlopen libdla_b200.so and attach the prototypes; raises when the build is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise DLALibraryError(
            "{} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). There is no CPU fallback.".format(LIB_PATH)
        )
    lib = ctypes.CDLL(LIB_PATH)
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def last_error() -> str:
    return load_library().dla_last_error().decode("utf-8", "replace")


def check(rc: int) -> None:
    if rc != 0:
        raise DLALibraryError(last_error())


def init(device: int = None) -> None:
    """Select the device (default: LOCAL_RANK, else 0)."""
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", "0"))
    check(load_library().dla_init(device))


# ---- array helpers ---------------------------------------------------------------------------
def f64(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float64)


def u8(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a).astype(np.uint8))


def dptr(a: np.ndarray):
    return a.ctypes.data_as(_dp) if a is not None else None


def iptr(a: np.ndarray):
    return a.ctypes.data_as(_ip) if a is not None else None


def bptr(a: np.ndarray):
    return a.ctypes.data_as(_bp) if a is not None else None


def params_struct(params, broadening: bool = True, min_z_separation: float = 0.0) -> DLAParamsStruct:
    """Pack the attributes of a `Parameters` object that the device path reads."""
    return DLAParamsStruct(
        params.min_lambda,
        params.max_lambda,
        params.normalization_min_lambda,
        params.normalization_max_lambda,
        params.pixel_spacing,
        int(params.width),
        int(params.num_forest_lines),
        int(params.num_lines),
        1 if broadening else 0,
        params.lya_wavelength,
        params.lyman_limit,
        params.max_z_cut,
        params.min_z_cut,
        float(min_z_separation),
    )
