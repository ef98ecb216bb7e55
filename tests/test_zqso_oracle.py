"""
CPU tests of the zQSO-estimation oracle (oracle/zqso_oracle.py) against the golden vectors written
from the live reference's ZGP (gpy_dla_detection/zqso_gp.py) by tests/golden/make_golden.py.
"""
import numpy as np
import pytest

from gpy_dla_detection_b200 import synthetic
from gpy_dla_detection_b200.zqso_samples import ZSamples
from gpy_dla_detection_b200.zqso_set_parameters import ZParameters
from oracle import ref_loader
from oracle import zqso_oracle as ZO
from tests import helpers as H


def test_zparameters_and_samples_match_reference_defaults():
    """zqso_set_parameters.py:19-54, zqso_samples.py:26-29"""
    p = ZParameters()
    assert (p.min_lambda, p.max_lambda, p.dlambda, p.k) == (910.0, 3000.0, 0.25, 20)
    assert (p.normalization_min_lambda, p.normalization_max_lambda) == (1176.0, 1256.0)
    assert p.num_zqso_samples == 10000 and p.max_noise_variance == 16.0
    assert not hasattr(p, "num_dla_samples")  # the reference skips Parameters.__init__
    z = ZSamples(ZParameters(num_zqso_samples=7)).sample_z_qsos()
    assert np.array_equal(z, np.linspace(2.14, 6.16, 7))
    assert p.observed_wavelengths(910.0, 3.0) == 910.0 * 4.0


def test_log_mvnpdf_iid_golden():
    g = H.golden("zqso_golden.npz")
    assert ZO.log_mvnpdf_iid(g["iid_y"], g["iid_mu"], g["iid_d"]) == float(g["iid_value"])


@pytest.mark.parametrize("case", [0, 1, 2])
def test_zqso_oracle_golden(case):
    g = H.golden("zqso_golden.npz")
    z_true, seed = g["cases"][case]
    model = synthetic.make_zqso_model(0)
    wl, fl, nv, pm = synthetic.make_zqso_spectrum(model, float(z_true), seed=int(seed))
    zs = np.linspace(2.14, 6.16, 96)
    out = ZO.inference_z_qso(model, wl, fl, nv, pm, zs)
    ref = g["ll_%d" % case]
    assert np.array_equal(np.isnan(out["sample_log_likelihoods"]), np.isnan(ref))
    assert H.ll_err(out["sample_log_likelihoods"], ref) < 1e-12
    assert out["z_map"] == float(g["z_map_%d" % case])
    zs2 = np.linspace(z_true - 0.05, z_true + 0.05, 96)
    out2 = ZO.inference_z_qso(model, wl, fl, nv, pm, zs2)
    assert H.ll_err(out2["sample_log_likelihoods"], g["ll_narrow_%d" % case]) < 1e-12
    assert out2["z_map"] == float(g["z_map_narrow_%d" % case])
    d = ZO.set_data(model, wl, fl, nv, pm, float(z_true) + 0.013)
    for k in ("x", "y", "v", "this_wavelengths", "this_mu", "this_M", "y_bw", "v_bw", "y_rw", "v_rw", "ind"):
        assert np.array_equal(d[k], g["%s_%d" % (k, case)], equal_nan=True), k
    assert abs(ZO.log_model_evidence(model, d) - float(g["evidence_%d" % case])) < 1e-12 * abs(float(g["evidence_%d" % case]))


@pytest.mark.skipif(not ref_loader.reference_available(), reason="/root/reference not present")
def test_zqso_oracle_against_live_reference():
    ref_loader.load_reference()
    from gpy_dla_detection.zqso_gp import ZGP as RZGP
    from gpy_dla_detection.zqso_samples import ZSamples as RZSamples
    from gpy_dla_detection.zqso_set_parameters import ZParameters as RZParameters

    model = synthetic.make_zqso_model(0)
    wl, fl, nv, pm = synthetic.make_zqso_spectrum(model, 3.9, seed=77)
    rp = RZParameters(num_zqso_samples=24)
    gp = RZGP(rp, RZSamples(rp), model["rest_wavelengths"], model["mu"], model["M"], model["bluewards_mu"],
              model["redwards_mu"], model["bluewards_sigma"], model["redwards_sigma"])
    gp.inference_z_qso(wl, fl, nv, pm)
    out = ZO.inference_z_qso(model, wl, fl, nv, pm, np.linspace(2.14, 6.16, 24))
    assert H.ll_err(out["sample_log_likelihoods"], gp.sample_log_likelihoods) < 1e-12
    assert out["z_map"] == gp.z_map
