"""
GPU parity tests of the quasar-redshift estimation path (ZGP, SURVEY.md §8 a14, BASELINE.json configs[4])
through the C-ABI, against the golden vectors written from the live reference and against the oracle.
Tolerances: sample log-likelihoods 1e-9 relative (|d| / max(|ll|, 1)), identical MAP redshift.
"""
import numpy as np
import pytest

from gpy_dla_detection_b200 import synthetic
from tests import helpers as H

pytestmark = pytest.mark.gpu
LL_RTOL = 1e-9


def build_zgp(num_samples):
    from gpy_dla_detection_b200.zqso_gp import ZGP
    from gpy_dla_detection_b200.zqso_samples import ZSamples
    from gpy_dla_detection_b200.zqso_set_parameters import ZParameters

    model = synthetic.make_zqso_model(0)
    p = ZParameters(num_zqso_samples=num_samples)
    gp = ZGP(p, ZSamples(p), model["rest_wavelengths"], model["mu"], model["M"], model["bluewards_mu"],
             model["redwards_mu"], model["bluewards_sigma"], model["redwards_sigma"])
    return model, gp


@pytest.mark.parametrize("case", [0, 1, 2])
def test_inference_z_qso_golden(gpu, case):
    g = H.golden("zqso_golden.npz")
    z_true, seed = g["cases"][case]
    model, gp = build_zgp(96)
    wl, fl, nv, pm = synthetic.make_zqso_spectrum(model, float(z_true), seed=int(seed))
    gp.inference_z_qso(wl, fl, nv, pm)
    ref = g["ll_%d" % case]
    assert np.array_equal(np.isnan(gp.sample_log_likelihoods), np.isnan(ref))
    assert H.ll_err(gp.sample_log_likelihoods, ref) < LL_RTOL
    assert gp.z_map == float(g["z_map_%d" % case])
    gp.inference_z_qso(wl, fl, nv, pm, z_qso_min=z_true - 0.05, z_qso_max=z_true + 0.05)
    assert H.ll_err(gp.sample_log_likelihoods, g["ll_narrow_%d" % case]) < LL_RTOL
    assert gp.z_map == float(g["z_map_narrow_%d" % case])


@pytest.mark.parametrize("case", [0, 1, 2])
def test_set_data_and_evidence_golden(gpu, case):
    g = H.golden("zqso_golden.npz")
    z_true, seed = g["cases"][case]
    model, gp = build_zgp(8)
    wl, fl, nv, pm = synthetic.make_zqso_spectrum(model, float(z_true), seed=int(seed))
    gp.set_data(wl, fl, nv, pm, z_qso=float(z_true) + 0.013, normalize=True, build_model=True)
    assert np.array_equal(gp.ind, g["ind_%d" % case])  # pixel selection: bit-exact
    for k in ("x", "y", "v", "this_wavelengths", "this_mu", "this_M", "y_bw", "v_bw", "y_rw", "v_rw"):
        assert np.array_equal(getattr(gp, k), g["%s_%d" % (k, case)], equal_nan=True), k  # same IEEE operations
    ref = float(g["evidence_%d" % case])
    assert abs(gp.log_model_evidence() - ref) < LL_RTOL * abs(ref)
    # the sweep kernel at that single redshift agrees with the attribute path
    out = gp.inference_z_qsos([(wl, fl, nv, pm)], np.array([float(z_true) + 0.013]))
    assert abs(out["sample_log_likelihoods"][0, 0] - ref) < LL_RTOL * abs(ref)


def test_log_mvnpdf_iid_golden(gpu):
    from gpy_dla_detection_b200.zqso_gp import ZGP

    g = H.golden("zqso_golden.npz")
    got = ZGP.log_mvnpdf_iid(g["iid_y"], g["iid_mu"], g["iid_d"])
    assert abs(got - float(g["iid_value"])) < 1e-12 * abs(float(g["iid_value"]))
    assert ZGP.log_mvnpdf_iid(np.empty(0), np.empty(0), np.empty(0)) == 0.0


def test_full_size_sweep_golden(gpu):
    """10 000 candidate redshifts on one spectrum against the live reference's output (28.8 s there)"""
    g = H.golden("zqso_full_S10000.npz")
    model, gp = build_zgp(10000)
    wl, fl, nv, pm = synthetic.make_zqso_spectrum(model, float(g["z_true"]), seed=int(g["seed"]))
    gp.inference_z_qso(wl, fl, nv, pm)
    ref = g["sample_log_likelihoods"]
    assert np.array_equal(np.isnan(gp.sample_log_likelihoods), np.isnan(ref))
    assert H.ll_err(gp.sample_log_likelihoods, ref) < LL_RTOL
    assert gp.z_map == float(g["z_map"])
    assert abs(gp.z_map - float(g["z_true"])) < 0.05  # the reference's accuracy criterion is 0.5 (test_zestimation.py:70)


def test_batched_sweep_against_oracle(gpu):
    """ragged batch (different lengths, heavy masking, NaN flux) == one-by-one oracle"""
    from oracle import zqso_oracle as ZO

    model, gp = build_zgp(40)
    zs = np.linspace(2.14, 6.16, 40)
    spectra = [synthetic.make_zqso_spectrum(model, z, seed=50 + i) for i, z in enumerate((2.2, 3.1, 4.4, 5.6))]
    wl, fl, nv, pm = spectra[1]
    rng = np.random.default_rng(3)
    pm = pm | (rng.random(pm.shape[0]) < 0.25)
    fl = np.where(pm & (rng.random(fl.shape[0]) < 0.5), np.nan, fl)  # NaN flux at half of the masked pixels:
    spectra[1] = (wl, fl, np.where(pm, np.inf, nv), pm)            # nanmedian has to skip them (zqso_gp.py:146)
    spectra[2] = tuple(a[500:4300] for a in spectra[2])
    out = gp.inference_z_qsos(spectra, zs)
    for q, spec in enumerate(spectra):
        ref = ZO.inference_z_qso(model, *spec, zs)
        assert np.array_equal(np.isnan(out["sample_log_likelihoods"][q]), np.isnan(ref["sample_log_likelihoods"])), q
        assert H.ll_err(out["sample_log_likelihoods"][q], ref["sample_log_likelihoods"]) < LL_RTOL, q
        assert out["z_map"][q] == ref["z_map"], q
    # without the sample array
    out2 = gp.inference_z_qsos(spectra, zs, keep_samples=False)
    assert np.array_equal(out2["z_map"], out["z_map"]) and "sample_log_likelihoods" not in out2


def test_all_nan_sweep(gpu):
    """an unmasked NaN flux inside every window makes every sample NaN: nanargmax raises, as in the reference"""
    model, gp = build_zgp(16)
    wl, fl, nv, pm = synthetic.make_zqso_spectrum(model, 3.0, seed=2)
    fl = fl.copy()
    fl[np.flatnonzero(~pm)[::50]] = np.nan
    out = gp.inference_z_qsos([(wl, fl, nv, pm)], np.linspace(2.14, 6.16, 16))
    assert np.all(np.isnan(out["sample_log_likelihoods"])) and out["map_index"][0] == -1 and np.isnan(out["z_map"][0])
    with pytest.raises(ValueError):
        gp.inference_z_qso(wl, fl, nv, pm)


def test_unsorted_wavelengths_are_rejected(gpu):
    model, gp = build_zgp(4)
    wl, fl, nv, pm = synthetic.make_zqso_spectrum(model, 3.0, seed=1)
    with pytest.raises(gpu.DLALibraryError, match="increasing"):
        gp.inference_z_qsos([(wl[::-1].copy(), fl, nv, pm)], np.array([3.0]))


def test_generic_kernel_and_uniform_grid_kernel_agree(gpu):
    """
    dla_zqso_inference picks zqso_likelihood_kernel_v2 (+ the median pass) for a uniform model grid and the generic
    kernel otherwise: both on the same inputs - a dense sweep (neighbouring samples share the sorted normalisation
    window) and a coarse one (one sort per sample in the median pass).
    """
    lib = gpu.load_library()
    model, gp = build_zgp(512)
    spectra = [synthetic.make_zqso_spectrum(model, z, seed=80 + i) for i, z in enumerate((2.3, 3.9, 5.2))]
    spectra[1] = tuple(a[250:4400] for a in spectra[1])
    for zs in (np.linspace(2.14, 6.16, 10000)[3000:3512], np.linspace(2.14, 6.16, 37)):
        fast = gp.inference_z_qsos(spectra, zs)
        try:
            gpu.check(lib.dla_zqso_force_generic_kernel(1))
            slow = gp.inference_z_qsos(spectra, zs)
        finally:
            gpu.check(lib.dla_zqso_force_generic_kernel(0))
        assert np.array_equal(np.isnan(fast["sample_log_likelihoods"]), np.isnan(slow["sample_log_likelihoods"]))
        assert H.ll_err(fast["sample_log_likelihoods"], slow["sample_log_likelihoods"]) < 1e-10
        assert np.array_equal(fast["map_index"], slow["map_index"])


def test_flux_scale_invariance(gpu):
    """
    ZGP normalises every spectrum by its own median, so flux in units of 1e-30 or 1e+30 must give the same sample
    likelihoods (products of variances / pivots must not overflow or underflow into +-inf).
    """
    model, gp = build_zgp(64)
    wl, fl, nv, pm = synthetic.make_zqso_spectrum(model, 3.3, seed=9)
    zs = np.linspace(2.14, 6.16, 64)
    base = gp.inference_z_qsos([(wl, fl, nv, pm)], zs)["sample_log_likelihoods"][0]
    assert np.all(np.isfinite(base))
    for c in (1e-30, 1e30):
        got = gp.inference_z_qsos([(wl, fl * c, nv * c * c, pm)], zs)["sample_log_likelihoods"][0]
        assert np.all(np.isfinite(got))
        assert H.ll_err(got, base) < 1e-9, c
