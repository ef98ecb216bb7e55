"""
CPU tests of the C-ABI boundary: libdla_b200.so builds for sm_100a, loads, exports every
symbol include/dla_b200.h declares (and the ctypes binding covers exactly that set), and
fails loudly - no CPU fallback - when no CUDA device is usable.
"""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "dla_b200.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(dla_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_expected_entry_points():
    syms = declared_symbols()
    for name in ("dla_voigt_absorption", "dla_effective_optical_depth", "dla_log_mvnpdf_low_rank",
                 "dla_spectrum_create", "dla_null_log_model_evidence", "dla_sample_log_likelihoods",
                 "dla_log_model_evidences", "dla_resample_indices", "dla_catalogue_process"):
        assert name in syms


def test_library_exports_every_declared_symbol(built):
    from gpy_dla_detection_b200 import _lib

    lib = _lib.load_library()
    syms = declared_symbols()
    assert sorted(_lib.SIGNATURES) == syms  # the ctypes binding is the header, nothing more, nothing less
    for name in syms:
        assert getattr(lib, name) is not None
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r"\bT (dla_[a-z0-9_]+)", out))
    assert set(syms) <= exported


def test_library_is_built_for_sm_100a(built):
    from gpy_dla_detection_b200 import _lib

    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    assert "sm_90" not in out and "sm_80" not in out  # one architecture, no multi-backend dispatch


def test_struct_layouts_match_the_header(built):
    """sizes the C side assumes for the by-pointer structs (ints pad to 8-byte alignment)"""
    from gpy_dla_detection_b200 import _lib

    assert ctypes.sizeof(_lib.DLAParamsStruct) == 5 * 8 + 4 * 4 + 5 * 8
    assert ctypes.sizeof(_lib.CatalogueConfigStruct) == 16
    assert ctypes.sizeof(_lib.CatalogueOutputsStruct) == 15 * 8


def test_version_string(built):
    from gpy_dla_detection_b200 import _lib

    v = _lib.load_library().dla_version().decode()
    assert "sm_100a" in v


@pytest.mark.skipif(os.path.exists("/dev/nvidia0"), reason="needs a box without a GPU")
def test_no_cpu_fallback_without_a_device(built):
    """every compute entry refuses to run when there is no CUDA device"""
    from gpy_dla_detection_b200 import _lib, voigt

    lib = _lib.load_library()
    assert lib.dla_init(0) != 0
    assert len(_lib.last_error()) > 0
    with pytest.raises(_lib.DLALibraryError):
        _lib.init(0)
    with pytest.raises(_lib.DLALibraryError):
        voigt.voigt_absorption(np.linspace(4000.0, 5000.0, 100), 1e21, 2.5)
    out = np.empty(3)
    y = np.zeros(3)
    assert lib.dla_log_mvnpdf_low_rank(_lib.dptr(y), _lib.dptr(y), _lib.dptr(y), _lib.dptr(y), 3, 1, _lib.dptr(out)) != 0


def test_missing_library_raises(monkeypatch, built):
    from gpy_dla_detection_b200 import _lib

    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libdla_b200.so")
    with pytest.raises(_lib.DLALibraryError, match="no CPU fallback"):
        _lib.load_library()


def test_product_package_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under the package may import it"""
    pkg = os.path.join(ROOT, "gpy_dla_detection_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert "dla_oracle" not in text, f
