"""
CPU test of the Faddeeva approximation used by the Voigt kernel: csrc/faddeeva.cuh compiles for the host
too (tests/host_shim/faddeeva_host.cpp, built by __graft_entry__.build()), so its accuracy against
scipy.special.wofz - the reference's own call, voigt.py:15,248 - is pinned without a GPU.  The device
build differs only in using the MUFU-seeded reciprocal (<= 1 ulp) instead of the IEEE division.
"""
import ctypes
import os

import numpy as np
from scipy.special import wofz

from tests import helpers as H

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def host_faddeeva(x, y):
    lib = ctypes.CDLL(os.path.join(ROOT, "tests", "_build", "libfadd_host.so"))
    dp = ctypes.POINTER(ctypes.c_double)
    lib.fadd_re_host.argtypes = [dp, dp, dp, ctypes.c_long]
    lib.fadd_re_host.restype = None
    x = np.ascontiguousarray(x, dtype=np.float64)
    yy = np.full_like(x, float(y))
    out = np.empty_like(x)
    lib.fadd_re_host(x.ctypes.data_as(dp), yy.ctypes.data_as(dp), out.ctypes.data_as(dp), x.shape[0])
    return out


def test_host_faddeeva_against_scipy(built):
    x = np.concatenate([np.linspace(0, 80, 80001), np.geomspace(80, 3e4, 20000)])
    for y in (4.72e-4, 1.2e-4, 3e-5, 7.2e-8):  # gamma_l / (sqrt(2) sigma) of Ly-alpha ... Ly-31
        ref = np.real(wofz(x + 1j * y))
        got = host_faddeeva(x, y)
        assert np.max(np.abs(got - ref) / ref) < 1e-12, y
        assert np.array_equal(host_faddeeva(-x, y), got)  # even in x


def test_host_faddeeva_golden(built):
    g = H.golden("voigt_golden.npz")
    for y, ref in zip(g["fadd_y"], g["fadd_re"]):
        assert np.max(np.abs(host_faddeeva(g["fadd_x"], y) - ref) / ref) < 1e-12
