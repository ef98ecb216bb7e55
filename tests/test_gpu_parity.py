"""
GPU parity tests (run on the B200 box with `pytest -m gpu`): the CUDA path, called through
the C-ABI of libdla_b200.so, against
  * the golden vectors the LIVE reference wrote (tests/golden/, made by make_golden.py),
  * the CPU oracle (oracle/dla_oracle.py) on the same seeded inputs,
  * size-independent properties at the full published sizes (S = 10 000, max_dlas = 4).

Tolerances are the ones BASELINE.json's north_star states:
  per-sample log-likelihoods 1e-9 relative; log evidences, posteriors, p(DLA) 1e-6 absolute;
  identical MAP model / DLA count; bit-exact QMC sample indices and pixel masks.
"""
import ctypes

import numpy as np
import pytest

from tests import helpers as H

pytestmark = pytest.mark.gpu

LL_RTOL = 1e-9   # north_star: per-sample log-likelihoods within 1e-9 relative (see helpers.ll_err for |ll| < 1)
EV_ATOL = 1e-6   # north_star: log evidences and p(DLA)/p(subDLA) within 1e-6 absolute


@pytest.fixture(scope="module")
def O():
    from oracle import dla_oracle

    return dla_oracle


# ---- a1: Faddeeva / Voigt profile -------------------------------------------------------------
def test_faddeeva_against_scipy_wofz_golden(gpu):
    from gpy_dla_detection_b200 import voigt

    g = H.golden("voigt_golden.npz")
    x = g["fadd_x"]
    for y, ref in zip(g["fadd_y"], g["fadd_re"]):
        got = voigt.faddeeva_re(x, y)
        assert np.max(np.abs(got - ref) / ref) < 1e-12


def test_faddeeva_dense_grid_against_scipy(gpu):
    from scipy.special import wofz
    from gpy_dla_detection_b200 import voigt

    x = np.concatenate([np.linspace(0, 80, 160001), np.geomspace(80, 3e4, 20000)])
    for y in (4.7e-4, 1.2e-4, 3e-5, 7e-8):  # gamma_l / (sqrt(2) sigma) spans [7.2e-8, 4.7e-4] for the 31 lines
        ref = np.real(wofz(x + 1j * y))
        got = voigt.faddeeva_re(x, y)
        ok = ref > 1e-300
        assert np.max(np.abs(got[ok] - ref[ok]) / ref[ok]) < 1e-12, y


def test_voigt_absorption_golden(gpu):
    from gpy_dla_detection_b200 import voigt

    g = H.golden("voigt_golden.npz")
    wl = g["wavelengths"]
    for i, (zd, ln, nl, br) in enumerate(g["cases"]):
        ref = g["profile_%d" % i]
        got = voigt.voigt_absorption(wl, 10.0**ln, zd, num_lines=int(nl), broadening=bool(br))
        assert got.shape == ref.shape  # len-6 with broadening, len without (voigt.py:293,319-322)
        assert np.max(np.abs(got - ref)) < 1e-13, (i, np.max(np.abs(got - ref)))
        batch = voigt.voigt_absorption_batch(wl, np.array([10.0**ln, 10.0**ln]), np.array([zd, zd]), int(nl), bool(br))
        assert np.array_equal(batch[0], got) and np.array_equal(batch[1], got)


def test_voigt_reference_structural_test(gpu):
    """reference tests/test_voigt.py:8-57 run against the drop-in module"""
    from gpy_dla_detection_b200 import voigt

    for z_qso, npix, z_dla, lognhi, lines in ((3.15, 1000, 3.1, 20.3, 3), (5.0, 50, 4.5, 21.0, 5)):
        wl = np.linspace(911, 1216, npix) * (1 + z_qso)
        raw = voigt.voigt_absorption(wl, 10**lognhi, z_dla, num_lines=lines, broadening=False)
        prof = np.zeros(wl.shape[0] - 2 * voigt.width)
        for i in range(prof.shape[0]):
            for k, j in enumerate(range(i, i + 2 * voigt.width + 1)):
                prof[i] += raw[j] * voigt.instrument_profile[k]
        got = voigt.voigt_absorption(wl, 10**lognhi, z_dla, num_lines=lines, broadening=True)
        assert np.all(np.abs(prof - got) < 1e-4)      # the reference's tolerance
        assert np.all(np.abs(prof - got) < 1e-13)     # ours


def test_voigt_edge_cases(gpu, O):
    from gpy_dla_detection_b200 import voigt

    wl = 10 ** (3.6 + 1e-4 * np.arange(7))  # shortest grid the broadened profile accepts -> 1 output
    got = voigt.voigt_absorption(wl, 1e21, 2.3, 3, True)
    assert got.shape == (1,) and abs(got[0] - O.voigt_absorption(wl, 1e21, 2.3, 3, True)[0]) < 1e-13
    # nhi = 0: no absorption at all
    wl = 10 ** (3.6 + 1e-4 * np.arange(500))
    assert np.all(voigt.voigt_absorption(wl, 0.0, 2.4, 3, False) == 1.0)
    # saturated core underflows to exactly 0 like the reference's exp()
    core = voigt.voigt_absorption(wl, 10**22.5, 1215.6701 ** -1 * wl[250] - 1, 3, False)
    ref = O.voigt_absorption(wl, 10**22.5, 1215.6701 ** -1 * wl[250] - 1, 3, False)
    assert core.min() == 0.0 == ref.min()
    assert np.max(np.abs(core - ref)) < 1e-13
    # all 31 members
    a = voigt.voigt_absorption(wl, 10**20.7, 3.2, 31, True)
    assert np.max(np.abs(a - O.voigt_absorption(wl, 10**20.7, 3.2, 31, True))) < 1e-13


# ---- a2: effective optical depth --------------------------------------------------------------
def test_effective_optical_depth_golden(gpu):
    from gpy_dla_detection_b200.effective_optical_depth import effective_optical_depth

    g = H.golden("voigt_golden.npz")
    wl = g["wavelengths"]
    for args, key in (((3.65, 0.0023, 3.2, 31), "eod_kim"), ((3.1, 0.0019, 2.9, 5), "eod_learned")):
        got = effective_optical_depth(wl, *args)
        ref = g[key]
        assert got.shape == ref.shape
        assert np.array_equal(got == 0.0, ref == 0.0)  # the z <= z_qso indicator, on every line incl. Lya
        assert H.rel_err(got, ref) < 1e-14


# ---- a5: low-rank Gaussian log-pdf --------------------------------------------------------------
def test_log_mvnpdf_known_answers(gpu):
    """reference tests/test_model.py:52-72"""
    from scipy.stats import multivariate_normal
    from gpy_dla_detection_b200.null_gp import NullGP

    mu = np.array([1.0, 2.0])
    M = np.array([[2.0, 3.0, 1.0], [1.0, 2.0, 4.0]])
    d = np.ones(2) * 2
    rv = multivariate_normal(mu, M @ M.T + np.eye(2) * 2)
    for y in ([1.0, 2.0], [2.0, 3.0], [100.0, 100.0]):
        y = np.array(y)
        got = NullGP.log_mvnpdf_low_rank(y, mu, M, d)
        assert abs(got - rv.logpdf(y)) < 1e-4                      # the reference's tolerance
        assert abs(got - rv.logpdf(y)) < 1e-10 * max(1.0, abs(got))  # ours
    assert abs(NullGP.log_mvnpdf_low_rank(np.array([1.0, 2.0]), mu, M, d) - (-4.5437000923)) < 1e-9


def test_log_mvnpdf_random_shapes(gpu, O):
    from gpy_dla_detection_b200.null_gp import NullGP

    rng = np.random.default_rng(0)
    for n, k in ((1, 1), (7, 3), (300, 20), (1250, 20), (513, 64)):
        y, mu, d = rng.standard_normal(n), 0.1 * rng.standard_normal(n), 0.05 + rng.random(n)
        M = 0.3 * rng.standard_normal((n, k))
        got, ref = NullGP.log_mvnpdf_low_rank(y, mu, M, d), O.log_mvnpdf_low_rank(y, mu, M, d)
        assert abs(got - ref) < 1e-11 * abs(ref), (n, k)


# ---- a3/a4: set_data + get_interp ---------------------------------------------------------------
@pytest.mark.parametrize("path", H.small_spectrum_fixtures(), ids=H.fixture_id)
def test_set_data_golden(gpu, path):
    g = np.load(path)
    st = H.Setup(int(g["S"]), int(g["num_lines"]))
    gp, _, _ = st.gp_objects()
    z_qso = float(g["z_qso"])
    gp.set_data(g["wavelengths"] / (1 + z_qso), g["flux"], g["noise_variance"], g["pixel_mask"], z_qso,
                build_model=True)
    # pixel masks: bit-exact
    assert gp.ind.dtype == np.bool_ and np.array_equal(gp.ind, g["ind"])
    assert np.array_equal(gp.ind_unmasked, g["ind_unmasked"])
    # selections / normalisation: same IEEE operations as the reference -> identical
    for k in ("x", "y", "v", "this_wavelengths", "unmasked_wavelengths"):
        assert np.array_equal(getattr(gp, k), g[k]), k
    assert gp.normalization_median == float(g["normalization_median"])
    # pow / exp / log based quantities: a few ulp
    assert H.rel_err(gp.padded_wavelengths, g["padded_wavelengths"]) < 1e-14
    assert H.rel_err(gp.this_mu, g["this_mu"]) < 1e-13
    assert np.max(np.abs(gp.this_M - g["this_M"])) < 1e-13 * np.max(np.abs(g["this_M"]))
    assert H.rel_err(gp.this_omega2, g["this_omega2"]) < 1e-13


# ---- a6-a13: full model selection through the reference's class API ------------------------------
def _run_class_api(st, g, max_dlas, broadening):
    from gpy_dla_detection_b200.bayesian_model_selection import BayesModelSelect

    z_qso = float(g["z_qso"])
    gp, sub, dla = st.gp_objects(broadening)
    rest = g["wavelengths"] / (1 + z_qso)
    for m in (gp, sub, dla):
        m.set_data(rest, g["flux"], g["noise_variance"], g["pixel_mask"], z_qso, build_model=True)
    np.random.seed(0)  # run_bayes_select.py:144
    bayes = BayesModelSelect([0, 1, max_dlas], 2)
    log_post = bayes.model_selection([gp, sub, dla], z_qso)
    return gp, sub, dla, bayes, log_post


def _check_against_golden(g, dla_ll, base_inds, sub_ll, log_priors, log_lik, log_post, model_post, p_dla,
                          map_z, map_n):
    ref_ll = g["sample_log_likelihoods_dla"]
    assert np.array_equal(np.isnan(dla_ll), np.isnan(ref_ll))
    assert H.ll_err(dla_ll, ref_ll) < LL_RTOL
    assert H.ll_err(sub_ll, g["sample_log_likelihoods_lls"]) < LL_RTOL
    assert base_inds.dtype == np.int32 and np.array_equal(base_inds, g["base_sample_inds"])  # bit-exact
    assert np.max(np.abs(log_priors - g["log_priors"])) < EV_ATOL
    assert np.max(np.abs(log_lik - g["log_likelihoods"])) < EV_ATOL
    assert np.max(np.abs(log_post - g["log_posteriors"])) < EV_ATOL
    assert np.max(np.abs(model_post - g["model_posteriors"])) < EV_ATOL
    assert abs(p_dla - float(g["p_dla"])) < EV_ATOL
    assert int(np.argmax(model_post)) == int(np.argmax(g["model_posteriors"]))  # MAP model index / DLA count
    assert np.array_equal(map_z, g["MAP_z_dlas"], equal_nan=True)
    assert np.array_equal(map_n, g["MAP_log_nhis"], equal_nan=True)


@pytest.mark.parametrize("path", H.small_spectrum_fixtures(), ids=H.fixture_id)
def test_model_selection_golden_class_api(gpu, path):
    g = np.load(path)
    S, md, nl, br = int(g["S"]), int(g["max_dlas"]), int(g["num_lines"]), bool(g["broadening"])
    st = H.Setup(S, nl)
    gp, sub, dla, bayes, log_post = _run_class_api(st, g, md, br)
    map_z, map_n = dla.maximum_a_posteriori()
    _check_against_golden(g, dla.sample_log_likelihoods, dla.base_sample_inds, sub.sample_log_likelihoods[:, 0],
                          bayes.log_priors, bayes.log_likelihoods, log_post, bayes.model_posteriors, bayes.p_dla,
                          map_z, map_n)
    assert abs(bayes.p_no_dla - float(g["p_no_dla"])) < EV_ATOL
    assert st.params.min_z_dla(g["wavelengths"], float(g["z_qso"])) == float(g["min_z_dla"])
    assert st.params.max_z_dla(g["wavelengths"], float(g["z_qso"])) == float(g["max_z_dla"])
    # the reference's single-sample entry points (dla_gp.py:311-396)
    zs, pick = g["sample_z_dlas"], g["pick"]
    nhi = st.dla["nhi_samples"]
    for i, ref in zip(pick, g["single_ll"]):
        got = dla.sample_log_likelihood_k_dlas(np.array([zs[i]]), np.array([nhi[i]]))
        assert abs(got - ref) < LL_RTOL * abs(ref)
    for i, ref in zip(pick, g["pair_ll"]):
        j = (i * 7 + 3) % S
        got = dla.sample_log_likelihood_k_dlas(np.array([zs[i], zs[j]]), np.array([nhi[i], nhi[j]]))
        assert abs(got - ref) < LL_RTOL * abs(ref)
    dmu, dM, dom = dla.this_dla_gp(np.array([zs[pick[2]], zs[pick[3]]]), np.array([nhi[pick[2]], nhi[pick[3]]]))
    assert np.max(np.abs(dmu - g["this_dla_mu"])) < 1e-13
    assert np.max(np.abs(dM - g["this_dla_M"])) < 1e-13
    assert np.max(np.abs(dom - g["this_dla_omega2"])) < 1e-13


@pytest.mark.parametrize("path", H.small_spectrum_fixtures(), ids=H.fixture_id)
def test_model_selection_golden_catalogue_api(gpu, path):
    """the batched engine (dla_catalogue_process) on the same fixture"""
    g = np.load(path)
    S, md, nl, br = int(g["S"]), int(g["max_dlas"]), int(g["num_lines"]), bool(g["broadening"])
    st = H.Setup(S, nl)
    proc = st.catalogue(md, br, batch_spectra=2)
    spec = (g["wavelengths"], g["flux"], g["noise_variance"], g["pixel_mask"])
    out = proc.process(*proc.pack([spec]), np.array([float(g["z_qso"])]), keep_samples=True)
    assert out["status"][0] == 0 and out["num_pixels"][0] == g["x"].shape[0]
    _check_against_golden(g, out["sample_log_likelihoods_dla"][0], out["base_sample_inds"][0].T,
                          out["sample_log_likelihoods_lls"][0], out["log_priors"][0], out["log_likelihoods"][0],
                          out["log_posteriors"][0], out["model_posteriors"][0], out["p_dlas"][0],
                          out["MAP_z_dlas"][0], out["MAP_log_nhis"][0])
    assert out["min_z_dlas"][0] == float(g["min_z_dla"]) and out["max_z_dlas"][0] == float(g["max_z_dla"])


def test_full_size_spectrum_golden(gpu):
    """config 1/2 size: S = 10 000 DLA + 10 000 subDLA samples, max_dlas = 4, against the live reference's output"""
    g = H.golden("spec_S10000_z2p9_full.npz")
    st = H.Setup(10000, 3)
    gp, sub, dla, bayes, log_post = _run_class_api(st, g, 4, True)
    assert np.array_equal(gp.ind, g["ind"]) and np.array_equal(gp.ind_unmasked, g["ind_unmasked"])
    map_z, map_n = dla.maximum_a_posteriori()
    _check_against_golden(g, dla.sample_log_likelihoods, dla.base_sample_inds, sub.sample_log_likelihoods[:, 0],
                          bayes.log_priors, bayes.log_likelihoods, log_post, bayes.model_posteriors, bayes.p_dla,
                          map_z, map_n)
    # and through the batched engine
    proc = st.catalogue(4, True, batch_spectra=2)
    spec = (g["wavelengths"], g["flux"], g["noise_variance"], g["pixel_mask"])
    out = proc.process(*proc.pack([spec]), np.array([float(g["z_qso"])]), keep_samples=True)
    _check_against_golden(g, out["sample_log_likelihoods_dla"][0], out["base_sample_inds"][0].T,
                          out["sample_log_likelihoods_lls"][0], out["log_priors"][0], out["log_likelihoods"][0],
                          out["log_posteriors"][0], out["model_posteriors"][0], out["p_dlas"][0],
                          out["MAP_z_dlas"][0], out["MAP_log_nhis"][0])


# ---- catalogue engine against the oracle: ragged batch, edge cases --------------------------------
def _ragged_batch(st):
    from gpy_dla_detection_b200 import synthetic

    z_qsos = [2.2, 2.55, 3.3, 4.9, 2.9, 3.7, 2.35]
    spectra = [synthetic.make_spectrum(st.model, z, seed=100 + i) for i, z in enumerate(z_qsos)]
    # heavy masking (30 % of the pixels)
    wl, fl, nv, pm = spectra[4]
    rng = np.random.default_rng(9)
    pm = pm | (rng.random(pm.shape[0]) < 0.3)
    spectra[4] = (wl, fl, np.where(pm, np.nan, nv), pm)
    # a truncated spectrum (different raw length)
    spectra[5] = tuple(a[300:4200] for a in spectra[5])
    return np.array(z_qsos), spectra


def test_catalogue_ragged_batch_against_oracle(gpu, O):
    S, md = 192, 4
    st = H.Setup(S)
    z_qsos, spectra = _ragged_batch(st)
    proc = st.catalogue(md, True, batch_spectra=3)  # 7 spectra -> batches of 3, 3, 1
    out = proc.process(*proc.pack(spectra), z_qsos, keep_samples=True)
    assert np.all(out["status"] == 0)
    for q, (z, spec) in enumerate(zip(z_qsos, spectra)):
        ref = O.process_spectrum(st.model, st.dla, st.sub, st.prior.less_ind(z), *spec, float(z), md)
        assert out["num_pixels"][q] == ref["prep"]["y"].shape[0]
        ll, rl = out["sample_log_likelihoods_dla"][q], ref["sample_log_likelihoods_dla"]
        assert np.array_equal(np.isnan(ll), np.isnan(rl)), q
        assert H.ll_err(ll, rl) < LL_RTOL, q
        assert H.ll_err(out["sample_log_likelihoods_lls"][q], ref["sample_log_likelihoods_lls"]) < LL_RTOL
        assert np.array_equal(out["base_sample_inds"][q].T, ref["base_sample_inds"]), q
        for k in ("log_priors", "log_likelihoods", "log_posteriors", "model_posteriors"):
            assert np.max(np.abs(out[k][q] - ref[k])) < EV_ATOL, (q, k)
        assert abs(out["p_dlas"][q] - ref["p_dla"]) < EV_ATOL and abs(out["p_no_dlas"][q] - ref["p_no_dla"]) < EV_ATOL
        assert np.argmax(out["model_posteriors"][q]) == np.argmax(ref["model_posteriors"])
        assert np.array_equal(out["MAP_z_dlas"][q], ref["MAP_z_dlas"], equal_nan=True)
        assert np.array_equal(out["MAP_log_nhis"][q], ref["MAP_log_nhis"], equal_nan=True)
        assert out["min_z_dlas"][q] == ref["min_z_dla"] and out["max_z_dlas"][q] == ref["max_z_dla"]


def test_catalogue_batching_is_invisible(gpu):
    """same spectra, different batch sizes and orders -> bit-identical results"""
    S, md = 128, 3
    st = H.Setup(S)
    z_qsos, spectra = _ragged_batch(st)
    a = st.catalogue(md, True, batch_spectra=7)
    oa = a.process(*a.pack(spectra), z_qsos, keep_samples=True)
    b = st.catalogue(md, True, batch_spectra=2)
    perm = np.array([3, 0, 6, 2, 5, 1, 4])
    ob = b.process(*b.pack([spectra[i] for i in perm]), z_qsos[perm], keep_samples=True)
    for k in ("sample_log_likelihoods_dla", "sample_log_likelihoods_lls", "base_sample_inds", "log_posteriors",
              "model_posteriors", "MAP_z_dlas", "p_dlas"):
        assert np.array_equal(oa[k][perm], ob[k], equal_nan=True), k
    # staged (device-resident) run == host-buffer run
    b.stage(*b.pack(spectra), z_qsos)
    oc = b.run_staged(keep_samples=True)
    for k in ("sample_log_likelihoods_dla", "base_sample_inds", "log_posteriors"):
        assert np.array_equal(oa[k], oc[k], equal_nan=True), k


def test_catalogue_unusable_spectra(gpu):
    """no pixel in the modelling range -> status 1 and NaN results, neighbours unaffected"""
    from gpy_dla_detection_b200 import synthetic

    S, md = 64, 2
    st = H.Setup(S)
    good = synthetic.make_spectrum(st.model, 2.8, seed=1)
    wl, fl, nv, pm = synthetic.make_spectrum(st.model, 2.8, seed=2)
    all_masked = (wl, fl, np.full_like(nv, np.nan), np.ones_like(pm))
    proc = st.catalogue(md, True, batch_spectra=4)
    z = np.array([2.8, 1.2, 2.8, 2.8])  # z_qso = 1.2: Lyman-alpha forest entirely bluewards of the spectrograph
    out = proc.process(*proc.pack([good, good, all_masked, good]), z, keep_samples=True)
    assert list(out["status"]) == [0, 1, 1, 0]
    assert out["num_pixels"][1] == 0 and out["num_pixels"][2] == 0
    assert np.all(np.isnan(out["log_likelihoods"][1])) and np.all(np.isnan(out["p_dlas"][[1, 2]]))
    assert np.array_equal(out["sample_log_likelihoods_dla"][0], out["sample_log_likelihoods_dla"][3], equal_nan=True)
    assert np.all(np.isfinite(out["log_posteriors"][0]))


# ---- a9 pieces ---------------------------------------------------------------------------------------
def test_resample_indices_is_numpy_choice(gpu):
    """dla_resample_indices == np.random.choice(S, S, p=W/W.sum()) on the same MT19937 stream (dla_gp.py:209-218)"""
    from gpy_dla_detection_b200 import _lib

    rng = np.random.default_rng(5)
    for n in (1, 7, 100, 129, 1000, 4097, 10000, 30000):
        W = np.exp(-rng.exponential(8.0, n))
        W[rng.random(n) < 0.1] = 0.0
        W[0] = max(W[0], 1e-3)
        np.random.seed(n)
        ref = np.random.choice(np.arange(n).astype(np.int32), size=n, replace=True, p=W / W.sum())
        U = np.random.RandomState(n).random_sample(n)
        out = np.empty(n, dtype=np.int32)
        _lib.check(_lib.load_library().dla_resample_indices(_lib.dptr(W), _lib.dptr(U), n, _lib.iptr(out)))
        assert np.array_equal(out, ref), n


def test_global_rng_stream_consumed_like_the_reference(gpu):
    """after log_model_evidences(4) the global MT19937 has advanced by 3 x S draws (np.random.choice x 3)"""
    g = np.load(H.small_spectrum_fixtures()[1])
    S = int(g["S"])
    st = H.Setup(S)
    _, _, dla = st.gp_objects()
    z_qso = float(g["z_qso"])
    dla.set_data(g["wavelengths"] / (1 + z_qso), g["flux"], g["noise_variance"], g["pixel_mask"], z_qso)
    np.random.seed(0)
    dla.log_model_evidences(4)
    after = np.random.random_sample()
    rs = np.random.RandomState(0)
    rs.random_sample(3 * S)
    assert after == rs.random_sample()


def test_nan_early_exit(gpu, O):
    """a level whose evidence is NaN ends the loop: later columns stay NaN, later index rows stay 0 (dla_gp.py:200-206)"""
    S = 96
    st = H.Setup(S)
    _, _, dla = st.gp_objects()
    g = np.load(H.small_spectrum_fixtures()[0])
    z_qso = float(g["z_qso"])
    dla.set_data(g["wavelengths"] / (1 + z_qso), g["flux"], g["noise_variance"], g["pixel_mask"], z_qso)
    # a separation so large that every pair of absorbers collides -> level 2 is all-NaN
    dla.min_z_separation = 10.0
    np.random.seed(0)
    ev = dla.log_model_evidences(4)
    assert np.isfinite(ev[0]) and np.all(np.isnan(ev[1:]))
    assert np.all(np.isnan(dla.sample_log_likelihoods[:, 1:]))
    assert np.any(dla.base_sample_inds[0] != 0) and np.all(dla.base_sample_inds[1:] == 0)
    prep = O.prepare_spectrum(st.model, g["wavelengths"] / (1 + z_qso), g["flux"], g["noise_variance"],
                              g["pixel_mask"], z_qso)
    U = np.random.RandomState(0).random_sample((3, S))
    ref = O.log_model_evidences(prep, st.dla["offset_samples"], st.dla["nhi_samples"], 4, U,
                                min_z_separation_kms=10.0 * 299792.458)
    assert np.array_equal(dla.base_sample_inds, ref["base_sample_inds"])
    assert np.array_equal(np.isnan(ev), np.isnan(ref["log_likelihoods"]))
    with pytest.raises(ValueError):
        dla.maximum_a_posteriori()  # nanargmax on an all-NaN column, as the reference (dla_gp.py:436)


# ---- size-independent properties at the full size ----------------------------------------------------
def test_full_size_properties(gpu):
    from gpy_dla_detection_b200 import synthetic

    S = 10000
    st = H.Setup(S)
    _, _, dla = st.gp_objects()
    gp, _, _ = st.gp_objects()
    z_qso = 3.1
    wl, fl, nv, pm = synthetic.make_spectrum(st.model, z_qso, seed=77)
    rest = wl / (1 + z_qso)
    dla.set_data(rest, fl, nv, pm, z_qso)
    gp.set_data(rest, fl, nv, pm, z_qso)
    zs = dla.dla_samples.sample_z_dlas(dla.this_wavelengths, z_qso)
    nhi = st.dla["nhi_samples"]
    # (1) an absorber with zero column density is the null model
    ll0 = dla.sample_log_likelihoods_batch(zs[:64, None], np.zeros((64, 1)))
    assert np.max(np.abs(ll0 - gp.log_model_evidence())) < 1e-9 * abs(gp.log_model_evidence())
    # (2) the likelihood of k absorbers does not depend on their order beyond rounding
    idx = np.random.default_rng(0).integers(0, S, size=(S, 4))
    a = dla.sample_log_likelihoods_batch(zs[idx], nhi[idx])
    b = dla.sample_log_likelihoods_batch(zs[idx[:, ::-1]], nhi[idx[:, ::-1]])
    assert H.rel_err(a, b) < 1e-11
    # (3) determinism: two runs of the whole level loop are bit-identical
    np.random.seed(0)
    ev1 = dla.log_model_evidences(4)
    ll1, bi1 = dla.sample_log_likelihoods.copy(), dla.base_sample_inds.copy()
    np.random.seed(0)
    ev2 = dla.log_model_evidences(4)
    assert np.array_equal(ev1, ev2) and np.array_equal(ll1, dla.sample_log_likelihoods, equal_nan=True)
    assert np.array_equal(bi1, dla.base_sample_inds)
    # (4) column 0 of the level loop == the batched single-absorber entry point, minus log S (dla_gp.py:157-159)
    single = dla.sample_log_likelihoods_batch(zs[:, None], nhi[:, None]) - np.log(S)
    assert H.rel_err(ll1[:, 0], single) < 1e-12
    # (5) resampled indices are a valid index set drawn only from samples with non-zero weight
    assert bi1.min() >= 0 and bi1.max() < S
    assert np.all(np.isfinite(ll1[bi1[0], 0]))
    # (6) evidences are log-mean-exp of the columns (dla_gp.py:180-190)
    for k in range(4):
        col = ll1[:, k]
        mx = np.nanmax(col)
        assert abs(ev1[k] - (mx + np.log(np.nanmean(np.exp(col - mx))) - k * np.log(S))) < 1e-9


# ---- no CPU fallback -----------------------------------------------------------------------------------
def test_kernels_actually_launch(gpu):
    from gpy_dla_detection_b200 import voigt

    lib = gpu.load_library()
    before = lib.dla_kernel_launch_count()
    voigt.voigt_absorption(10 ** (3.6 + 1e-4 * np.arange(100)), 1e21, 2.3, 3, True)
    assert lib.dla_kernel_launch_count() > before
    assert lib.dla_last_kernel_ms() > 0.0
    dfma, dmma = ctypes.c_double(), ctypes.c_double()
    gpu.check(lib.dla_measure_fp64_peaks(ctypes.byref(dfma), ctypes.byref(dmma)))
    assert 5.0 < dfma.value < 80.0 and 5.0 < dmma.value < 80.0  # TFLOP/s, B200 FP64


# ---- BASELINE.json configs[3]: full Lyman series (31 lines), 30 000 samples, broadening -----------------
def test_config4_full_lyman_series_30k_samples(gpu, O):
    """
    num_lines = 31, S = 30 000, max_dlas = 4 on one spectrum.  The full oracle would need 2e9 wofz calls,
    so parity is checked (a) on a random subset of the 120 000 sample likelihoods, re-evaluated one by one
    by the oracle's restatement of sample_log_likelihood_k_dlas with the absorber chains the GPU drew, and
    (b) through the size-independent identities of the level loop.
    """
    from gpy_dla_detection_b200 import synthetic

    S, md, nl = 30000, 4, 31
    st = H.Setup(S, nl)
    z_qso = 3.4
    spec = synthetic.make_spectrum(st.model, z_qso, seed=404)
    proc = st.catalogue(md, True, batch_spectra=1)
    out = proc.process(*proc.pack([spec]), np.array([z_qso]), keep_samples=True)
    assert out["status"][0] == 0
    ll = out["sample_log_likelihoods_dla"][0]          # (S, 4)
    inds = out["base_sample_inds"][0].T                # (3, S)
    assert inds.min() >= 0 and inds.max() < S
    prep = O.prepare_spectrum(st.model, spec[0] / (1 + z_qso), spec[1], spec[2], spec[3], z_qso)
    zs = O.sample_z_dlas(st.dla["offset_samples"], prep["this_wavelengths"], z_qso)
    nhi = st.dla["nhi_samples"]
    rng = np.random.default_rng(1)
    for level in range(md):
        finite = np.flatnonzero(np.isfinite(ll[:, level]))
        for i in rng.choice(finite, size=6, replace=False):
            chain = np.concatenate([[i], inds[:level, i]]).astype(int)
            ref = O.sample_log_likelihood_k_dlas(prep, zs[chain], nhi[chain], nl, True) - np.log(S)
            assert abs(ll[i, level] - ref) < LL_RTOL * max(abs(ref), 1.0), (level, i)
    # subDLA column against the oracle on a subset
    zs_sub = O.sample_z_dlas(st.sub["offset_samples"], prep["this_wavelengths"], z_qso)
    for i in rng.choice(S, size=6, replace=False):
        ref = O.sample_log_likelihood_k_dlas(prep, zs_sub[[i]], st.sub["nhi_samples"][[i]], nl, True) - np.log(S)
        assert abs(out["sample_log_likelihoods_lls"][0][i] - ref) < LL_RTOL * max(abs(ref), 1.0)
    # level identities (dla_gp.py:164-190): separation mask, log-mean-exp evidences
    sep = st.params.kms_to_z(3000.0)
    for level in range(1, md):
        allz = np.concatenate([zs[None, :], zs[inds[:level]]], axis=0)
        too_close = np.any(np.diff(np.sort(allz, axis=0), axis=0) < sep, axis=0)
        assert np.array_equal(np.isnan(ll[:, level]), too_close)
    for level in range(md):
        col = ll[:, level]
        mx = np.nanmax(col)
        ev = mx + np.log(np.nanmean(np.exp(col - mx))) - level * np.log(S)
        assert abs(out["log_likelihoods"][0][2 + level] - ev) < EV_ATOL
    # resampling reproduces NumPy's choice on the GPU's own weights (bit-exact)
    U = np.random.RandomState(0).random_sample((md - 1, S))
    for level in range(md - 1):
        col = ll[:, level]
        W = np.exp(col - np.nanmax(col))
        W[np.isnan(W)] = 0.0
        assert np.array_equal(inds[level], O.resample_indices(W, U[level]))


def test_catalogue_unpaired_subdla_offsets(gpu, O):
    """
    The reference's subDLA samples reuse the DLA redshift offsets (subdla_samples.py:87) and the profile kernel
    then evaluates the line sums once for both; with different offsets it must fall back to separate profiles.
    """
    from gpy_dla_detection_b200 import synthetic
    from gpy_dla_detection_b200.run_bayes_select import CatalogueProcessor
    from gpy_dla_detection_b200.subdla_samples import SubDLASamplesArrays

    S, md = 160, 2
    st = H.Setup(S)
    sub = dict(st.sub)
    sub["offset_samples"] = np.ascontiguousarray(st.sub["offset_samples"][::-1])
    d, _ = st.sample_objects()
    s = SubDLASamplesArrays(st.params, st.prior, sub["offset_samples"], sub["log_nhi_samples"], sub["nhi_samples"],
                            sub["Z_lls"], sub["Z_dla"])
    proc = CatalogueProcessor(st.params, st.prior, st.model, d, s, md, True, batch_spectra=2)
    z_qsos = np.array([2.6, 3.5])
    spectra = [synthetic.make_spectrum(st.model, z, seed=200 + i) for i, z in enumerate(z_qsos)]
    out = proc.process(*proc.pack(spectra), z_qsos, keep_samples=True)
    for q, (z, spec) in enumerate(zip(z_qsos, spectra)):
        ref = O.process_spectrum(st.model, st.dla, sub, st.prior.less_ind(z), *spec, float(z), md)
        assert H.ll_err(out["sample_log_likelihoods_lls"][q], ref["sample_log_likelihoods_lls"]) < LL_RTOL
        assert H.ll_err(out["sample_log_likelihoods_dla"][q], ref["sample_log_likelihoods_dla"]) < LL_RTOL
        assert np.array_equal(out["base_sample_inds"][q].T, ref["base_sample_inds"])
        assert np.max(np.abs(out["log_posteriors"][q] - ref["log_posteriors"])) < EV_ATOL


def test_likelihood_kernel_panel_tails_and_many_absorbers(gpu, O):
    """sample_likelihood_kernel at pixel counts around every panel boundary (1 .. 5 panels of 16 pixels, partial
    last panels, masked pixels), sample counts that leave a partial last tile, and 1 .. 5 absorbers per sample
    (more than two factors take the out-of-line path), against the oracle's sample_log_likelihood_k_dlas
    (dla_gp.py:311-396)."""
    from gpy_dla_detection_b200 import log_posterior_mcmc as M

    rng = np.random.default_rng(5)
    k, W = 20, 70
    worst = 0.0
    for n_u in [1, 2, 3, 7, 15, 16, 17, 31, 32, 33, 47, 48, 49, 63, 64, 65, 80, 81]:
        padded = 10 ** (3.6 + 1e-4 * np.arange(n_u + 6))           # observed wavelengths, BOSS pixel scale
        keep = rng.random(n_u) > 0.1
        keep[rng.integers(n_u)] = True
        n = int(keep.sum())
        prep = {
            "padded_wavelengths": padded, "mask_ind": keep,
            "this_mu": 1.0 + 0.1 * rng.standard_normal(n), "this_M": 0.2 * rng.standard_normal((n, k)),
            "this_omega2": rng.uniform(0.01, 0.1, n), "y": 1.0 + 0.3 * rng.standard_normal(n),
            "v": rng.uniform(0.01, 0.2, n),
        }
        z_mid = padded[3 + n_u // 2] / 1215.6701 - 1.0
        for kd in (1, 2, 3, 5):
            zz = z_mid + 0.01 * rng.standard_normal((W, kd))
            nn = 10 ** rng.uniform(19.5, 21.5, (W, kd))
            got = M.sample_log_likelihoods(zz, nn, prep["y"], prep["v"], padded, prep["this_mu"], prep["this_M"],
                                           prep["this_omega2"], ~keep, np.ones(n_u, bool), 3)
            ref = np.array([O.sample_log_likelihood_k_dlas(prep, zz[i], nn[i], 3) for i in range(W)])
            assert np.all(np.isfinite(got)), (n_u, kd)
            worst = max(worst, H.ll_err(got, ref))
    assert worst < 1e-9, worst  # |d ll| / max(|ll|, 1), the tolerance of the full-size tests


def test_catalogue_repeats_are_bit_identical(gpu):
    """The likelihood kernel's warps run unsynchronised between split barriers: a lost phase or a race would show
    as run-to-run differences.  Full-size samples (S = 10 000, max_dlas = 4), 24 spectra, 4 repeats; the long
    version is tools/soak.py."""
    from gpy_dla_detection_b200 import synthetic

    st = H.Setup(10000)
    z_qsos = synthetic.sample_z_qsos(24, seed=99)
    spectra = [synthetic.make_spectrum(st.model, float(z), seed=500 + i) for i, z in enumerate(z_qsos)]
    cat = st.catalogue(4, True, batch_spectra=16)
    packed = cat.pack(spectra)
    first = None
    for _ in range(4):
        out = cat.process(*packed, z_qsos, keep_samples=True)
        arrays = {k: v for k, v in out.items() if isinstance(v, np.ndarray)}
        if first is None:
            first = {k: v.copy() for k, v in arrays.items()}
            assert np.isfinite(first["p_dlas"]).all()
            continue
        for k, v in first.items():
            assert np.array_equal(v, arrays[k], equal_nan=True), k


def test_unnormalised_flux_known_answers(gpu, O):
    """
    ADVICE r1: the log-determinant is a log of exponent-renormalised products, which must stay finite where the
    reference (a sum of logs) does: noise variance ~ 1e-30 (20 pivots of ~1e31: their plain product overflows),
    ~ 1e+30, and a consistent rescaling of flux and model by 1e+-15.
    """
    import ctypes

    from gpy_dla_detection_b200.log_posterior_mcmc import _Prepared

    rng = np.random.default_rng(5)
    n, k = 700, 20
    M = 0.1 * rng.standard_normal((n, k))
    mu = 1.0 + 0.1 * rng.standard_normal(n)
    y = mu + M @ rng.standard_normal(k) + 0.05 * rng.standard_normal(n)
    om2 = 0.01 * (1.0 + rng.random(n))
    wl = 10 ** (3.6 + 1e-4 * np.arange(n + 6))
    mask, ind_unmasked = np.zeros(n, dtype=bool), np.ones(n, dtype=bool)
    lib = gpu.load_library()

    def null_evidence(y_, v_, mu_, M_, om2_):
        prep = _Prepared(y_, v_, wl, mu_, M_, om2_, mask, ind_unmasked)
        out = ctypes.c_double()
        gpu.check(lib.dla_null_log_model_evidence(prep.handle.ptr, ctypes.byref(out)))
        return out.value

    cases = [
        ("v=1e-30", y, np.full(n, 1e-30), mu, M, np.zeros(n)),
        ("v=1e+30", y, np.full(n, 1e30), mu, M, om2),
        ("scale 1e-15", y * 1e-15, np.full(n, 0.04) * 1e-30, mu * 1e-15, M * 1e-15, om2 * 1e-30),
        ("scale 1e+15", y * 1e15, np.full(n, 0.04) * 1e30, mu * 1e15, M * 1e15, om2 * 1e30),
    ]
    for name, y_, v_, mu_, M_, om2_ in cases:
        ref = O.log_mvnpdf_low_rank(y_, mu_, M_, om2_ + v_)
        got = null_evidence(y_, v_, mu_, M_, om2_)
        assert np.isfinite(ref) and np.isfinite(got), name
        assert abs(got - ref) < 1e-9 * max(abs(ref), 1.0), (name, got, ref)
    # a zero variance makes d = 0: the reference's log(0) / division by zero gives nan or -inf, never +inf
    bad = null_evidence(y, np.zeros(n), mu, M, np.zeros(n))
    assert not (bad == np.inf)
