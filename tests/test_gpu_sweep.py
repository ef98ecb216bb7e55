"""
GPU parity at the published sizes on the workload bench.py measures (VERDICT r1 task 1):
  * >= 128 spectra of the configs[1] bench workload (S = 10 000, max_dlas = 4, z_QSO 2.15 ... 5): CUDA catalogue engine
    vs the CPU oracle on the box's host cores, and vs the live-reference goldens of 20 of them
    (tests/golden/bench_sweep_S10000.npz, written by make_golden.py --only-sweep from /root/reference);
  * configs[3] (31 lines, S = 30 000) on a short spectrum: WHOLE columns of per-sample likelihoods vs the oracle.
Tolerances: north_star's (ll 1e-9 relative, evidences / posteriors 1e-6 absolute, indices / masks / MAP identical).
"""
import json
import os

import numpy as np
import pytest

from tests import helpers as H

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_bench_workload_sweep_against_oracle_and_live_reference_goldens(gpu):
    import sys

    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import parity_sweep

    num = int(os.environ.get("DLA_SWEEP_SPECTRA", "128"))
    report, failures = parity_sweep.run_sweep(num, verbose=False)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "parity_sweep.json"), "w") as f:
        json.dump(report, f, indent=1)
    summary = {k: v for k, v in report.items() if k not in ("per_spectrum", "per_golden")}
    print(json.dumps(summary))
    assert not failures, "\n".join(failures[:20])
    assert report["num_spectra"] == num
    assert report["all_base_sample_inds_identical"] and report["all_nan_patterns_identical"]
    assert report["all_map_models_identical"] and report["all_map_parameters_identical"]
    assert report["worst_vs_oracle"]["ll"] < 1e-9 and report["worst_vs_oracle"]["ev"] < 1e-6
    if num >= 128:
        assert report["live_reference_goldens"]["count"] == 20
        assert report["live_reference_goldens"]["all_indices_masks_maps_identical"]
        assert report["z_qso_max"] > 3.5  # the sweep reaches the high-redshift tail of the draw


def test_config3_whole_columns_short_spectrum(gpu):
    """
    BASELINE configs[3] (num_lines = 31, S = 30 000 + 30 000, max_dlas = 4, broadening) on a z_QSO = 2.15 sightline
    (n ~ 300 pixels): every one of the 150 001 likelihoods, the resampled indices, evidences, posteriors and MAP
    against the oracle (about 600 M Faddeeva evaluations on the host).
    """
    from gpy_dla_detection_b200 import synthetic
    from oracle import dla_oracle as O

    S, md, nl, z_qso = 30000, 4, 31, 2.15
    st = H.Setup(S, nl)
    spec = synthetic.make_spectrum(st.model, z_qso, seed=505)
    proc = st.catalogue(md, True, batch_spectra=1)
    out = proc.process(*proc.pack([spec]), np.array([z_qso]), keep_samples=True)
    assert out["status"][0] == 0
    ref = O.process_spectrum(st.model, st.dla, st.sub, st.prior.less_ind(z_qso), *spec, z_qso, md, nl, True)
    assert out["num_pixels"][0] == ref["prep"]["y"].shape[0]
    ll, rl = out["sample_log_likelihoods_dla"][0], ref["sample_log_likelihoods_dla"]
    assert ll.shape == (S, md)
    assert np.array_equal(np.isnan(ll), np.isnan(rl))
    assert H.ll_err(ll, rl) < 1e-9
    assert H.ll_err(out["sample_log_likelihoods_lls"][0], ref["sample_log_likelihoods_lls"]) < 1e-9
    assert np.array_equal(out["base_sample_inds"][0].T, ref["base_sample_inds"])
    for k in ("log_priors", "log_likelihoods", "log_posteriors", "model_posteriors"):
        assert np.max(np.abs(out[k][0] - ref[k])) < 1e-6, k
    assert abs(out["p_dlas"][0] - ref["p_dla"]) < 1e-6
    assert np.argmax(out["model_posteriors"][0]) == np.argmax(ref["model_posteriors"])
    assert np.array_equal(out["MAP_z_dlas"][0], ref["MAP_z_dlas"], equal_nan=True)
    assert np.array_equal(out["MAP_log_nhis"][0], ref["MAP_log_nhis"], equal_nan=True)


def _zqso_oracle_one(job):
    from oracle import zqso_oracle as ZO

    model, spec, zs = job
    out = ZO.inference_z_qso(model, *spec, zs)
    return out["sample_log_likelihoods"], float(out["z_map"])


def test_zqso_bench_workload_against_live_reference_goldens_and_oracle(gpu):
    """
    configs[4] at the published size (10 000 z_QSO samples per spectrum) on spectra of bench.py's own workload: the
    round-2 ZGP kernels vs (a) the live reference's output for 8 of them (tests/golden/zqso_bench_sweep_S10000.npz) and
    (b) the oracle on every 10th redshift sample of 16 of them, on the box's host cores.
    """
    import multiprocessing as mp

    from gpy_dla_detection_b200 import synthetic
    from gpy_dla_detection_b200.zqso_gp import ZGP
    from gpy_dla_detection_b200.zqso_samples import ZSamples
    from gpy_dla_detection_b200.zqso_set_parameters import ZParameters

    Q = 22
    model, z_true, spectra = synthetic.make_zqso_workload(Q, 0)
    p = ZParameters(num_zqso_samples=10000)
    gp = ZGP(p, ZSamples(p), model["rest_wavelengths"], model["mu"], model["M"], model["bluewards_mu"], model["redwards_mu"],
             model["bluewards_sigma"], model["redwards_sigma"])
    zs = ZSamples(p).sample_z_qsos()
    out = gp.inference_z_qsos(spectra, zs)
    ll = out["sample_log_likelihoods"]
    # (a) live reference
    g = H.golden("zqso_bench_sweep_S10000.npz")
    stride = int(g["stride"])
    for j, q in enumerate(g["indices"]):
        q = int(q)
        assert float(g["z_true"][j]) == float(z_true[q])
        assert H.ll_err(ll[q][::stride], g["ll_strided"][j]) < 1e-9, q
        assert int(out["map_index"][q]) == int(g["argmax"][j]) and out["z_map"][q] == float(g["z_map"][j]), q
        assert abs(np.nanmax(ll[q]) - float(g["ll_max"][j])) < 1e-9 * abs(float(g["ll_max"][j]))
    # (b) oracle on a coarse subset of the same sweep (every 10th sample: 1 000 per spectrum)
    sub = np.arange(0, 10000, 10)
    jobs = [(model, spectra[q], zs[sub]) for q in range(16)]
    with mp.get_context("spawn").Pool(min(16, os.cpu_count() or 1)) as pool:
        refs = pool.map(_zqso_oracle_one, jobs, chunksize=1)
    for q, (rll, _) in enumerate(refs):
        assert np.array_equal(np.isnan(ll[q][sub]), np.isnan(rll)), q
        assert H.ll_err(ll[q][sub], rll) < 1e-9, q
