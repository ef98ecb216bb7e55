"""
GPU tests of the interop layer (VERDICT r1 task 9): the REFERENCE's own classes and its own
BayesModelSelect.model_selection (with its isinstance checks, bayesian_model_selection.py:58-61) running on the
device after interop.patch_reference(), and NullGP.get_interp as a callable on an arbitrary grid.
The reference package is imported from /root/reference (build container) or oracle/_ref (GPU box, staged by
oracle/make_ref.sh); the tests skip when neither is present.
"""
import numpy as np
import pytest

from gpy_dla_detection_b200 import synthetic
from tests import helpers as H

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ref():
    from oracle import ref_loader

    if not ref_loader.reference_available():
        pytest.skip("reference package neither at /root/reference nor staged in oracle/_ref")
    return ref_loader.load_reference(), ref_loader


def test_reference_classes_run_on_the_device_after_patch_reference(gpu, ref):
    pkg, ref_loader = ref
    from gpy_dla_detection.bayesian_model_selection import BayesModelSelect as RBayes
    from gpy_dla_detection.dla_gp import DLAGP as RDLAGP
    from gpy_dla_detection.null_gp import NullGP as RNullGP
    from gpy_dla_detection.set_parameters import Parameters as RParameters
    from gpy_dla_detection.subdla_gp import SubDLAGP as RSubDLAGP

    from gpy_dla_detection_b200 import interop

    g = H.golden("spec_S256_z2p3.npz")
    S, z_qso, md = int(g["S"]), float(g["z_qso"]), int(g["max_dlas"])
    st = H.Setup(S, int(g["num_lines"]))
    rp = RParameters(num_dla_samples=S, num_lines=int(g["num_lines"]))
    spec = (g["wavelengths"], g["flux"], g["noise_variance"], g["pixel_mask"])
    margs = H.model_args(st.model)
    numpy_set_data = RNullGP.__dict__["set_data"]
    lib = gpu.load_library()

    undo = interop.patch_reference(pkg)
    try:
        assert RNullGP.__dict__["set_data"] is not numpy_set_data
        # objects built by the REFERENCE's constructors (interp1d tables and all)
        gp = RNullGP(rp, st.prior, *margs)
        dla = RDLAGP(rp, st.prior, ref_loader.RefDLASamples(rp, st.dla), *margs, broadening=True)
        sub = RSubDLAGP(rp, st.prior, ref_loader.RefDLASamples(rp, st.sub, True), *margs, broadening=True)
        launches0 = lib.dla_kernel_launch_count()
        rest = rp.emitted_wavelengths(spec[0], z_qso)
        for m in (gp, sub, dla):
            m.set_data(rest, spec[1], spec[2], spec[3], z_qso, build_model=True)
        np.random.seed(0)
        bayes = RBayes([0, 1, md], 2)                      # the reference's own class, isinstance checks included
        log_post = bayes.model_selection([gp, sub, dla], z_qso)
        assert lib.dla_kernel_launch_count() - launches0 > 10  # the arithmetic ran in our kernels
        map_z, map_n = dla.maximum_a_posteriori()           # the reference's own NumPy method, on device results
        assert np.array_equal(gp.ind, g["ind"]) and np.array_equal(gp.ind_unmasked, g["ind_unmasked"])
        assert np.array_equal(dla.base_sample_inds, g["base_sample_inds"])
        assert H.ll_err(dla.sample_log_likelihoods, g["sample_log_likelihoods_dla"]) < 1e-9
        assert H.ll_err(sub.sample_log_likelihoods[:, 0], g["sample_log_likelihoods_lls"]) < 1e-9
        assert np.max(np.abs(log_post - g["log_posteriors"])) < 1e-6
        assert abs(bayes.p_dla - float(g["p_dla"])) < 1e-6
        assert np.array_equal(map_z, g["MAP_z_dlas"], equal_nan=True) and np.array_equal(map_n, g["MAP_log_nhis"], equal_nan=True)
        # single-sample entry points and this_dla_gp through the patched methods
        zs = dla.dla_samples.sample_z_dlas(dla.this_wavelengths, z_qso)
        i = int(g["pick"][2])
        got = dla.sample_log_likelihood_k_dlas(np.array([zs[i]]), np.array([st.dla["nhi_samples"][i]]))
        assert abs(got - float(g["single_ll"][2])) < 1e-9 * abs(float(g["single_ll"][2]))
    finally:
        undo()
    assert RNullGP.__dict__["set_data"] is numpy_set_data and "_model_handle" not in RNullGP.__dict__


def test_get_interp_on_an_arbitrary_grid(gpu, ref):
    """NullGP.get_interp(x, y, wavelengths, z_qso) vs the reference's (null_gp.py:179-242) on pixels of our choosing."""
    from gpy_dla_detection.null_gp import NullGP as RNullGP
    from gpy_dla_detection.set_parameters import Parameters as RParameters

    from gpy_dla_detection_b200.null_gp import NullGP

    st = H.Setup(64)
    z_qso = 3.05
    rng = np.random.default_rng(2)
    x = np.sort(rng.uniform(911.75, 1215.75, 333))
    x[0], x[-1], x[7] = 911.75, 1215.75, st.model["rest_wavelengths"][400]  # grid ends and an exact grid point
    wl = x * (1 + z_qso)
    rgp = RNullGP(RParameters(), st.prior, *H.model_args(st.model))
    rgp.get_interp(x, None, wl, z_qso)
    gp = NullGP(st.params, st.prior, *H.model_args(st.model))
    gp.get_interp(x, None, wl, z_qso)
    assert np.max(np.abs(gp.this_mu - rgp.this_mu)) < 1e-13
    assert np.max(np.abs(gp.this_M - rgp.this_M)) < 1e-13
    assert np.max(np.abs(gp.this_omega2 / rgp.this_omega2 - 1)) < 1e-12
    with pytest.raises(ValueError):
        gp.get_interp(np.array([900.0, 1000.0]), None, np.array([3600.0, 4000.0]), z_qso)


def test_get_interp_after_set_data_rebuilds_the_device_model(gpu):
    """set_data(build_model=False) + get_interp(own grid) == set_data(build_model=True); a modified model is honoured."""
    from gpy_dla_detection_b200.null_gp import NullGP

    st = H.Setup(64)
    z_qso = 2.7
    wl, fl, nv, pm = synthetic.make_spectrum(st.model, z_qso, seed=77)
    a = NullGP(st.params, st.prior, *H.model_args(st.model))
    a.set_data(wl / (1 + z_qso), fl, nv, pm, z_qso, build_model=True)
    b = NullGP(st.params, st.prior, *H.model_args(st.model))
    b.set_data(wl / (1 + z_qso), fl, nv, pm, z_qso, build_model=False)
    assert not hasattr(b, "this_mu")
    b.get_interp(b.x, b.y, b.this_wavelengths, z_qso)
    assert np.array_equal(a.this_mu, b.this_mu) and np.array_equal(a.this_M, b.this_M)
    assert np.array_equal(a.this_omega2, b.this_omega2)
    ev_a, ev_b = a.log_model_evidence(), b.log_model_evidence()
    assert abs(ev_a - ev_b) < 1e-9 * abs(ev_a)
    # interpolating at shifted rest wavelengths changes the model the likelihood sees, as in the reference
    b.get_interp(np.clip(b.x + 0.1, 911.75, 1215.75), b.y, b.this_wavelengths, z_qso)
    ev_c = b.log_model_evidence()
    expect = NullGP.log_mvnpdf_low_rank(b.y, b.this_mu, b.this_M, b.this_omega2 + b.v)
    assert abs(ev_c - expect) < 1e-9 * abs(expect) and abs(ev_c - ev_a) > 1e-6
