"""
GPU parity tests of the Lyman-limit-system variant: voigt_lls.voigt_absorption and a DLAGP
whose absorbers carry the Lyman-limit break, against golden values from the live reference
(voigt_lls.py:254-363; subclass pattern of examples/gp_find_lls.py:159-224).
"""
import numpy as np
import pytest

from gpy_dla_detection_b200 import synthetic
from tests import helpers as H

pytestmark = pytest.mark.gpu


def test_voigt_lls_golden(gpu):
    from gpy_dla_detection_b200 import voigt, voigt_lls

    g = H.golden("lls_golden.npz")
    wl = g["wavelengths"]
    for i, (zl, ln, nl, br) in enumerate(g["cases"]):
        ref = g["profile_%d" % i]
        got = voigt_lls.voigt_absorption(wl, 10.0**ln, zl, num_lines=int(nl), broadening=bool(br))
        assert got.shape == ref.shape and np.max(np.abs(got - ref)) < 1e-13, i
        assert np.array_equal(voigt_lls.tau_LLS_break(wl, 10.0**ln, zl), g["tau_%d" % i])
        # redwards of the limit the break vanishes: identical to the plain Voigt profile there
        plain = voigt.voigt_absorption(wl, 10.0**ln, zl, num_lines=int(nl), broadening=bool(br))
        red = (wl[: got.shape[0]] / (1 + zl)) > voigt_lls.lambda_Lyman_limit + 2.0
        assert np.array_equal(got[red], plain[red])
        if np.any(g["tau_%d" % i] > 0):  # the grid reaches bluewards of this absorber's Lyman limit
            assert not np.array_equal(got, plain)


def test_lls_gp_evidences_golden(gpu):
    from gpy_dla_detection_b200.dla_samples import DLASamplesArrays
    from gpy_dla_detection_b200.lls_gp import LLSGP

    g = H.golden("lls_golden.npz")
    st = H.Setup(int(g["S"]), 4)
    z_qso = float(g["z_qso"])
    wl, fl, nv, pm = synthetic.make_spectrum(st.model, z_qso, seed=int(g["seed"]))
    d = DLASamplesArrays(st.params, st.prior, st.dla["offset_samples"], st.dla["log_nhi_samples"], st.dla["nhi_samples"])
    gp = LLSGP(st.params, st.prior, d, *H.model_args(st.model))
    gp.set_data(wl / (1 + z_qso), fl, nv, pm, z_qso, build_model=True)
    np.random.seed(0)
    ev = gp.log_model_evidences(3)
    assert np.max(np.abs(ev - g["log_evidences"])) < 1e-6
    assert H.ll_err(gp.sample_log_likelihoods, g["sample_log_likelihoods"]) < 1e-9
    assert np.array_equal(gp.base_sample_inds, g["base_sample_inds"])


def test_lls_extended_model_grid_golden(gpu):
    """
    The model range of examples/gp_find_lls.py:102,162-170: rest grid 850.75-1420.75 A (2 281 points), normalisation
    window 1425-1475 A, num_lines = 4, absorbers from log N = 17; n = 2 194 modelled pixels (the DLA range has <= 1 250).
    """
    from gpy_dla_detection_b200.dla_samples import DLASamplesArrays
    from gpy_dla_detection_b200.lls_gp import LLSGP
    from gpy_dla_detection_b200.set_parameters import Parameters

    g = H.golden("lls_golden.npz")
    S, z_qso = int(g["ext_S"]), float(g["ext_z_qso"])
    p = Parameters(num_dla_samples=S, num_lines=4, min_lambda=850.75, max_lambda=1420.75,
                   normalization_min_lambda=1425.0, normalization_max_lambda=1475.0)
    model = synthetic.make_learned_model(1, rest_min=850.75, rest_max=1420.75)
    assert model["rest_wavelengths"].shape == (2281,) and model["rest_wavelengths"][-1] == 1420.75
    prior = synthetic.SyntheticPrior(p)
    lls = synthetic.make_lls_sample_arrays(S)
    wl, fl, nv, pm = synthetic.make_spectrum(model, z_qso, seed=int(g["ext_seed"]), params=p)
    d = DLASamplesArrays(p, prior, lls["offset_samples"], lls["log_nhi_samples"], lls["nhi_samples"])
    gp = LLSGP(p, prior, d, *H.model_args(model))
    gp.set_data(wl / (1 + z_qso), fl, nv, pm, z_qso, build_model=True)
    assert np.array_equal(gp.ind, g["ext_ind"]) and np.array_equal(gp.x, g["ext_x"])
    assert gp.normalization_median == float(g["ext_normalization_median"])
    assert np.max(np.abs(gp.this_mu - g["ext_this_mu"])) < 1e-13
    assert np.max(np.abs(gp.this_omega2 / g["ext_this_omega2"] - 1)) < 1e-12
    np.random.seed(0)
    ev = gp.log_model_evidences(2)
    assert np.max(np.abs(ev - g["ext_log_evidences"])) < 1e-6
    assert H.ll_err(gp.sample_log_likelihoods, g["ext_sample_log_likelihoods"]) < 1e-9
    assert np.array_equal(gp.base_sample_inds, g["ext_base_sample_inds"])
