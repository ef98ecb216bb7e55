"""
CPU tests of the oracle (oracle/dla_oracle.py): it must reproduce
  * the reference's own data-free known-answer tests (tests/test_model.py:52-72,
    tests/test_voigt.py:8-57, tests/test_set_parameters.py of the reference),
  * the golden vectors written by tests/golden/make_golden.py from the LIVE reference,
  * and, when /root/reference is present (build container), the live reference itself.
"""
import glob
import os

import numpy as np
import pytest
from scipy.stats import multivariate_normal

from oracle import dla_oracle as O
from oracle import ref_loader
from gpy_dla_detection_b200 import _tables as T
from gpy_dla_detection_b200 import synthetic
from gpy_dla_detection_b200.set_parameters import Parameters

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def spectrum_fixtures():
    return sorted(f for f in glob.glob(os.path.join(GOLDEN, "spec_S*.npz")) if "S10000" not in f)


# ---- the reference's own known-answer tests -------------------------------------------------
def test_log_mvnpdf_known_answers():
    """reference tests/test_model.py:52-72, tolerance 1e-4 there; 1e-10 here"""
    mu = np.array([1.0, 2.0])
    M = np.array([[2.0, 3.0, 1.0], [1.0, 2.0, 4.0]])
    rv = multivariate_normal(mu, M @ M.T + np.eye(2) * 2)
    for y in ([1.0, 2.0], [2.0, 3.0], [100.0, 100.0]):
        y = np.array(y)
        lp = O.log_mvnpdf_low_rank(y, mu, M, np.ones(2) * 2)
        assert abs(lp - rv.logpdf(y)) < 1e-10 * max(1.0, abs(lp))
    assert abs(O.log_mvnpdf_low_rank(np.array([1.0, 2.0]), mu, M, np.ones(2) * 2) - (-4.5437000923)) < 1e-9


def test_instrumental_broadening_structure():
    """reference tests/test_voigt.py:8-57: explicit 7-tap loop == the convolution of the profile"""
    for z_qso, npix, z_dla, lognhi, lines in ((3.15, 1000, 3.1, 20.3, 3), (5.0, 50, 4.5, 21.0, 5)):
        wl = np.linspace(911, 1216, npix) * (1 + z_qso)
        raw = O.voigt_absorption(wl, 10**lognhi, z_dla, num_lines=lines, broadening=False)
        prof = np.zeros(wl.shape[0] - 2 * T.WIDTH)
        for i in range(prof.shape[0]):
            for k, j in enumerate(range(i, i + 2 * T.WIDTH + 1)):
                prof[i] += raw[j] * T.INSTRUMENT_PROFILE[k]
        assert np.all(np.abs(prof - O.voigt_absorption(wl, 10**lognhi, z_dla, lines, True)) < 1e-12)
        assert np.all(np.abs(prof - O.voigt_absorption_batch(wl, [10**lognhi], [z_dla], lines, True)[0]) < 1e-12)


def test_parameters_known_answers():
    """reference tests/test_set_parameters.py"""
    p = Parameters()
    assert abs(p.kms_to_z(3000) - 0.01) < 1e-4
    wl = np.linspace(911, 1216, 100)
    assert np.all(np.abs(p.emitted_wavelengths(p.observed_wavelengths(wl, 3.0), 3.0) - wl) < 1e-10)
    assert O.kms_to_z(3000) == p.kms_to_z(3000)


# ---- golden vectors from the live reference ---------------------------------------------------
def test_voigt_golden():
    g = load("voigt_golden.npz")
    wl = g["wavelengths"]
    for i, (zd, ln, nl, br) in enumerate(g["cases"]):
        a = O.voigt_absorption(wl, 10.0**ln, zd, int(nl), bool(br))
        assert np.array_equal(a, g["profile_%d" % i])  # same NumPy/SciPy calls: bit-identical
        b = O.voigt_absorption_batch(wl, [10.0**ln], [zd], int(nl), bool(br))[0]
        assert np.max(np.abs(b - g["profile_%d" % i])) < 5e-16
    assert np.array_equal(O.effective_optical_depth(wl, 3.65, 0.0023, 3.2, 31), g["eod_kim"])
    assert np.array_equal(O.effective_optical_depth(wl, 3.1, 0.0019, 2.9, 5), g["eod_learned"])


@pytest.mark.parametrize("path", spectrum_fixtures(), ids=lambda p: os.path.basename(p)[:-4])
def test_spectrum_golden(path):
    g = np.load(path)
    S, md, nl, br = int(g["S"]), int(g["max_dlas"]), int(g["num_lines"]), bool(g["broadening"])
    p = Parameters(num_dla_samples=S, num_lines=nl)
    model = synthetic.make_learned_model(0)
    dla = synthetic.make_dla_sample_arrays(p)
    sub = synthetic.make_subdla_sample_arrays(p)
    out = O.process_spectrum(model, dla, sub, tuple(g["prior_counts"]), g["wavelengths"], g["flux"],
                             g["noise_variance"], g["pixel_mask"], float(g["z_qso"]), md, nl, br)
    prep = out["prep"]
    # pixel masks and the prepared model: bit-exact
    assert np.array_equal(prep["ind"], g["ind"]) and np.array_equal(prep["ind_unmasked"], g["ind_unmasked"])
    for k in ("x", "y", "v", "this_wavelengths", "unmasked_wavelengths", "padded_wavelengths", "this_mu", "this_M",
              "this_omega2"):
        assert np.array_equal(prep[k], g[k]), k
    assert prep["normalization_median"] == g["normalization_median"]
    # per-sample log-likelihoods: 1e-9 relative (north_star); QMC indices: bit-exact
    ref_ll = g["sample_log_likelihoods_dla"]
    assert np.array_equal(np.isnan(out["sample_log_likelihoods_dla"]), np.isnan(ref_ll))
    assert np.nanmax(np.abs(out["sample_log_likelihoods_dla"] - ref_ll) / np.abs(ref_ll)) < 1e-9
    assert np.nanmax(np.abs(out["sample_log_likelihoods_lls"] - g["sample_log_likelihoods_lls"])
                     / np.abs(g["sample_log_likelihoods_lls"])) < 1e-9
    assert np.array_equal(out["base_sample_inds"], g["base_sample_inds"])
    # evidences, priors, posteriors: 1e-6 absolute
    for k in ("log_priors", "log_likelihoods", "log_posteriors", "model_posteriors"):
        assert np.max(np.abs(out[k] - g[k])) < 1e-6, k
    assert abs(out["p_dla"] - g["p_dla"]) < 1e-6
    assert np.argmax(out["model_posteriors"]) == np.argmax(g["model_posteriors"])
    assert np.array_equal(out["MAP_z_dlas"], g["MAP_z_dlas"], equal_nan=True)
    assert np.array_equal(out["MAP_log_nhis"], g["MAP_log_nhis"], equal_nan=True)
    assert out["min_z_dla"] == g["min_z_dla"] and out["max_z_dla"] == g["max_z_dla"]
    # the reference's single-sample entry points
    zs, pick = g["sample_z_dlas"], g["pick"]
    for i, ref in zip(pick, g["single_ll"]):
        got = O.sample_log_likelihood_k_dlas(prep, np.array([zs[i]]), np.array([dla["nhi_samples"][i]]), nl, br)
        assert abs(got - ref) < 1e-9 * abs(ref)
    for i, ref in zip(pick, g["pair_ll"]):
        j = (i * 7 + 3) % S
        got = O.sample_log_likelihood_k_dlas(prep, np.array([zs[i], zs[j]]),
                                             np.array([dla["nhi_samples"][i], dla["nhi_samples"][j]]), nl, br)
        assert abs(got - ref) < 1e-9 * abs(ref)


def test_full_size_spectrum_golden_oracle():
    """S = 10 000, max_dlas = 4 (config 1/2 size): the restatement against the live reference's output (~30 s)"""
    from tests import helpers as H

    g = load("spec_S10000_z2p9_full.npz")
    st = H.Setup(10000, 3)
    out = O.process_spectrum(st.model, st.dla, st.sub, tuple(g["prior_counts"]), g["wavelengths"], g["flux"],
                             g["noise_variance"], g["pixel_mask"], float(g["z_qso"]), 4, 3, True)
    ref_ll = g["sample_log_likelihoods_dla"]
    assert np.array_equal(out["prep"]["ind"], g["ind"]) and np.array_equal(out["prep"]["ind_unmasked"], g["ind_unmasked"])
    assert np.array_equal(np.isnan(out["sample_log_likelihoods_dla"]), np.isnan(ref_ll))
    assert H.ll_err(out["sample_log_likelihoods_dla"], ref_ll) < 1e-9
    assert H.ll_err(out["sample_log_likelihoods_lls"], g["sample_log_likelihoods_lls"]) < 1e-9
    assert np.array_equal(out["base_sample_inds"], g["base_sample_inds"])  # 30 000 resampled indices, bit-exact
    for k in ("log_priors", "log_likelihoods", "log_posteriors", "model_posteriors"):
        assert np.max(np.abs(out[k] - g[k])) < 1e-6, k
    assert abs(out["p_dla"] - float(g["p_dla"])) < 1e-6
    assert np.array_equal(out["MAP_z_dlas"], g["MAP_z_dlas"], equal_nan=True)
    assert np.array_equal(out["MAP_log_nhis"], g["MAP_log_nhis"], equal_nan=True)


def test_batch_likelihood_matches_literal_form():
    """the BLAS-friendly batch used by the oracle == the literal restatement of null_gp.py:307-360"""
    g = np.load(spectrum_fixtures()[0])
    model = synthetic.make_learned_model(0)
    prep = O.prepare_spectrum(model, g["wavelengths"] / (1 + float(g["z_qso"])), g["flux"], g["noise_variance"],
                              g["pixel_mask"], float(g["z_qso"]))
    rng = np.random.default_rng(0)
    a = np.clip(1 - 0.6 * rng.random((9, prep["y"].shape[0])) ** 6, 0, 1)
    fast = O.batch_log_likelihoods(prep, a)
    mu = prep["this_mu"][None] * a
    M = prep["this_M"][None] * a[:, :, None]
    d = prep["this_omega2"][None] * a**2 + prep["v"][None]
    literal = O.log_mvnpdf_low_rank_batch(prep["y"], mu, M, d)
    single = np.array([O.log_mvnpdf_low_rank(prep["y"], mu[i], M[i], d[i]) for i in range(a.shape[0])])
    assert np.max(np.abs(fast - literal) / np.abs(literal)) < 1e-11
    assert np.max(np.abs(single - literal) / np.abs(literal)) < 1e-11


def test_resampling_is_numpy_choice():
    """np.random.choice(S, S, p) == searchsorted(cumsum(p)/cumsum(p)[-1], U, 'right') on the same MT19937 draws"""
    rng = np.random.default_rng(3)
    for S in (17, 300, 10000):
        W = np.exp(-rng.exponential(6.0, S))
        W[rng.random(S) < 0.2] = 0.0
        np.random.seed(0)
        ref = np.random.choice(np.arange(S).astype(np.int32), size=S, replace=True, p=W / W.sum())
        U = np.random.RandomState(0).random_sample(S)
        assert np.array_equal(O.resample_indices(W.copy(), U), ref)


def numpy_pairwise_sum(a):
    """Python transliteration of NumPy's pairwise summation (what csrc/evidence_kernel.cuh restates)."""
    n = len(a)
    if n < 8:
        res = 0.0
        for v in a:
            res += v
        return res
    if n <= 128:
        r = [a[i] for i in range(8)]
        i = 8
        while i < n - (n % 8):
            for k in range(8):
                r[k] += a[i + k]
            i += 8
        res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]))
        while i < n:
            res += a[i]
            i += 1
        return res
    n2 = n // 2
    n2 -= n2 % 8
    return numpy_pairwise_sum(a[:n2]) + numpy_pairwise_sum(a[n2:])


def test_pairwise_sum_algorithm_matches_numpy():
    rng = np.random.default_rng(1)
    for n in (5, 8, 100, 128, 129, 1000, 4097, 10000, 30000):
        a = np.exp(-rng.exponential(5.0, n))
        assert numpy_pairwise_sum(a.tolist()) == float(a.sum()), n


def test_tables_match_golden_checksum():
    """the Lyman-series data must stay bit-identical to the reference's literals (voigt.py:18-224)"""
    g = load("voigt_golden.npz")
    assert np.array_equal(g["tables"], np.stack([T.TRANSITION_WAVELENGTHS, T.OSCILLATOR_STRENGTHS, T.TRANSITION_RATES,
                                                 T.LEADING_CONSTANTS, T.GAMMAS]))
    assert np.array_equal(g["instrument_profile"], T.INSTRUMENT_PROFILE)
    assert float(g["sigma"]) == T.SIGMA and float(g["c"]) == T.SPEED_OF_LIGHT_CGS


# ---- live reference (build container only) ------------------------------------------------------
@pytest.mark.skipif(not ref_loader.reference_available(), reason="/root/reference not present")
def test_oracle_against_live_reference():
    ref_loader.load_reference()
    from gpy_dla_detection import voigt as rv
    from gpy_dla_detection.null_gp import NullGP as RNullGP
    from gpy_dla_detection.effective_optical_depth import effective_optical_depth as r_eod

    assert np.array_equal(rv.leading_constants, T.LEADING_CONSTANTS) and np.array_equal(rv.gammas, T.GAMMAS)
    assert np.array_equal(rv.transition_wavelengths, T.TRANSITION_WAVELENGTHS)
    wl = 10 ** (3.56 + 1e-4 * np.arange(900))
    for zd, ln, nl, br in ((2.1, 20.5, 3, True), (2.3, 21.9, 7, False)):
        assert np.array_equal(rv.voigt_absorption(wl, 10**ln, zd, nl, br), O.voigt_absorption(wl, 10**ln, zd, nl, br))
    assert np.array_equal(r_eod(wl, 3.65, 0.0023, 2.4, 31), O.effective_optical_depth(wl, 3.65, 0.0023, 2.4, 31))
    rng = np.random.default_rng(0)
    n, k = 300, 20
    y, mu, d = rng.standard_normal(n), rng.standard_normal(n) * 0.1, 0.05 + rng.random(n)
    M = rng.standard_normal((n, k)) * 0.3
    a, b = RNullGP.log_mvnpdf_low_rank(y, mu, M, d), O.log_mvnpdf_low_rank(y, mu, M, d)
    assert abs(a - b) < 1e-10 * abs(a)
