#!/usr/bin/env python
"""
parity_sweep.py : parity of the CUDA catalogue engine on the workload bench.py measures (VERDICT r1 task 1).

  python tools/parity_sweep.py [--spectra 128] [--workers N] [--out profiles/parity_sweep_r02.json]

The first `--spectra` spectra of bench.py's configs[1] workload (synthetic.make_workload(Q, seed0 = 0): full size,
S = 10 000 DLA + 10 000 subDLA samples, max_dlas = 4, z_QSO 2.15 ... 5) go through
  * the device engine (dla_catalogue_process, per-sample arrays kept), and
  * the CPU oracle (oracle/dla_oracle.process_spectrum) on the host cores, one spectrum per worker,
and are compared at the tolerances of BASELINE.json's north_star.  The 20 spectra of
tests/golden/bench_sweep_S10000.npz are additionally compared with what the LIVE reference produced for them
(tests/golden/make_golden.py --only-sweep).  TEST INFRASTRUCTURE: the oracle is the checker, never the product.
"""
import argparse
import hashlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

LL_RTOL = 1e-9
EV_ATOL = 1e-6
S_SAMPLES, MAX_DLAS, NUM_LINES = 10000, 4, 3


def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def _ll_err(a, b):
    ok = ~np.isnan(b)
    if not np.any(ok):
        return 0.0
    return float(np.max(np.abs(a[ok] - b[ok]) / np.maximum(np.abs(b[ok]), 1.0)))


def _oracle_init():
    for var in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[var] = "1"
    try:
        from threadpoolctl import threadpool_limits

        globals()["_LIMIT"] = threadpool_limits(limits=1)
    except Exception:
        pass
    from oracle import dla_oracle  # noqa: F401


def _oracle_one(job):
    from oracle import dla_oracle

    model, dla, sub, counts, spec, z_qso = job
    wl, fl, nv, pm = spec
    out = dla_oracle.process_spectrum(model, dla, sub, counts, wl, fl, nv, pm, z_qso, MAX_DLAS, NUM_LINES, True)
    keep = ("sample_log_likelihoods_dla", "sample_log_likelihoods_lls", "base_sample_inds", "log_priors", "log_likelihoods",
            "log_posteriors", "model_posteriors", "p_dla", "p_no_dla", "MAP_z_dlas", "MAP_log_nhis", "min_z_dla", "max_z_dla")
    res = {k: out[k] for k in keep}
    res["num_pixels"] = int(out["prep"]["y"].shape[0])
    return res


def run_sweep(num_spectra=128, workers=None, batch_spectra=64, verbose=True):
    """Returns (report dict, list of failure strings)."""
    import multiprocessing as mp

    import __graft_entry__ as graft

    graft.build()
    from gpy_dla_detection_b200 import _lib, synthetic
    from gpy_dla_detection_b200.dla_samples import DLASamplesArrays
    from gpy_dla_detection_b200.null_gp import NullGP
    from gpy_dla_detection_b200.run_bayes_select import CatalogueProcessor
    from gpy_dla_detection_b200.subdla_samples import SubDLASamplesArrays

    _lib.init(0)
    workers = workers or max(1, min(os.cpu_count() or 1, 32))
    params, model, prior, dla, sub, z_qsos, spectra = synthetic.make_workload(num_spectra, 0, S_SAMPLES, NUM_LINES)

    # the oracle starts first (it takes minutes), the GPU works underneath
    for var in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[var] = "1"
    pool = mp.get_context("spawn").Pool(workers, initializer=_oracle_init)
    jobs = [(model, dla, sub, prior.less_ind(z_qsos[i]), spectra[i], float(z_qsos[i])) for i in range(num_spectra)]
    t0 = time.perf_counter()
    pending = pool.map_async(_oracle_one, jobs, chunksize=1)

    dla_s = DLASamplesArrays(params, prior, dla["offset_samples"], dla["log_nhi_samples"], dla["nhi_samples"])
    sub_s = SubDLASamplesArrays(params, prior, sub["offset_samples"], sub["log_nhi_samples"], sub["nhi_samples"],
                                sub["Z_lls"], sub["Z_dla"])
    proc = CatalogueProcessor(params, prior, model, dla_s, sub_s, MAX_DLAS, True, batch_spectra=batch_spectra)
    tg = time.perf_counter()
    out = proc.process(*proc.pack(spectra), z_qsos, keep_samples=True)
    gpu_s = time.perf_counter() - tg
    refs = pending.get()
    oracle_s = time.perf_counter() - t0
    pool.close()
    pool.join()

    failures = []
    rows = []
    worst = dict(ll=0.0, ll_lls=0.0, ev=0.0, post=0.0, p_dla=0.0)
    for q in range(num_spectra):
        ref = refs[q]
        ll, rl = out["sample_log_likelihoods_dla"][q], ref["sample_log_likelihoods_dla"]
        row = dict(index=q, z_qso=float(z_qsos[q]), n=int(out["num_pixels"][q]), status=int(out["status"][q]))
        row["nan_pattern_equal"] = bool(np.array_equal(np.isnan(ll), np.isnan(rl)))
        row["ll_err"] = _ll_err(ll, rl)
        row["ll_lls_err"] = _ll_err(out["sample_log_likelihoods_lls"][q], ref["sample_log_likelihoods_lls"])
        row["inds_equal"] = bool(np.array_equal(out["base_sample_inds"][q].T, ref["base_sample_inds"]))
        with np.errstate(invalid="ignore"):
            row["ev_err"] = float(np.nanmax(np.abs(out["log_likelihoods"][q] - ref["log_likelihoods"])))
            row["post_err"] = float(np.nanmax(np.abs(out["model_posteriors"][q] - ref["model_posteriors"])))
        row["p_dla_err"] = float(abs(out["p_dlas"][q] - ref["p_dla"]))
        gpu_model, ref_model = int(np.argmax(out["model_posteriors"][q])), int(np.argmax(ref["model_posteriors"]))
        row["map_model"] = gpu_model
        row["map_model_equal"] = gpu_model == ref_model
        # MAP DLA count as the reference's tests count it (tests/test_selection.py:443-446)
        cnt = lambda mp_, pn: int(np.argmax(np.concatenate([[pn], mp_[2:]])))  # noqa: E731
        row["dla_count"] = cnt(out["model_posteriors"][q], out["p_no_dlas"][q])
        row["dla_count_equal"] = row["dla_count"] == cnt(ref["model_posteriors"], ref["p_no_dla"])
        row["map_params_equal"] = bool(np.array_equal(out["MAP_z_dlas"][q], ref["MAP_z_dlas"], equal_nan=True)
                                       and np.array_equal(out["MAP_log_nhis"][q], ref["MAP_log_nhis"], equal_nan=True))
        row["z_range_equal"] = bool(out["min_z_dlas"][q] == ref["min_z_dla"] and out["max_z_dlas"][q] == ref["max_z_dla"])
        row["n_equal"] = row["n"] == ref["num_pixels"]
        rows.append(row)
        worst["ll"] = max(worst["ll"], row["ll_err"])
        worst["ll_lls"] = max(worst["ll_lls"], row["ll_lls_err"])
        worst["ev"] = max(worst["ev"], row["ev_err"])
        worst["post"] = max(worst["post"], row["post_err"])
        worst["p_dla"] = max(worst["p_dla"], row["p_dla_err"])
        for key, ok in (("nan_pattern_equal", row["nan_pattern_equal"]), ("ll_err", row["ll_err"] < LL_RTOL),
                        ("ll_lls_err", row["ll_lls_err"] < LL_RTOL), ("inds_equal", row["inds_equal"]),
                        ("ev_err", row["ev_err"] < EV_ATOL), ("post_err", row["post_err"] < EV_ATOL),
                        ("p_dla_err", row["p_dla_err"] < EV_ATOL), ("map_params_equal", row["map_params_equal"]),
                        ("z_range_equal", row["z_range_equal"]), ("n_equal", row["n_equal"]), ("status", row["status"] == 0)):
            if not ok:
                failures.append("spectrum %d (z=%.3f, n=%d): %s = %r" % (q, row["z_qso"], row["n"], key, row[key.replace("_equal", "_equal")]))
        if not (row["map_model_equal"] and row["dla_count_equal"]):
            # a flip is only excusable when the two candidates are degenerate: list it with their posterior gap
            srt = np.sort(ref["model_posteriors"])[::-1]
            failures.append("spectrum %d: MAP model %d vs oracle %d (oracle's top-two posterior gap %.3e)"
                            % (q, gpu_model, ref_model, srt[0] - srt[1]))

    # ---- live-reference goldens for the spectra that have them ---------------------------------------------
    golden_rows = []
    gpath = os.path.join(ROOT, "tests", "golden", "bench_sweep_S10000.npz")
    if os.path.exists(gpath):
        g = np.load(gpath)
        stride = int(g["stride"])
        margs = (model["rest_wavelengths"], model["mu"], model["M"], model["log_omega"], model["log_c_0"],
                 model["log_tau_0"], model["log_beta"])
        for j, q in enumerate(g["indices"]):
            q = int(q)
            if q >= num_spectra:
                continue
            assert abs(float(g["z_qso"][j]) - float(z_qsos[q])) == 0.0
            ll = out["sample_log_likelihoods_dla"][q]
            inds = np.ascontiguousarray(out["base_sample_inds"][q].T)
            gp = NullGP(params, prior, *margs)
            wl, fl, nv, pm = spectra[q]
            gp.set_data(params.emitted_wavelengths(wl, float(z_qsos[q])), fl, nv, pm, float(z_qsos[q]), build_model=False)
            r = dict(index=q,
                     ll_err=_ll_err(ll[::stride], g["ll_dla_strided"][j]),
                     ll_lls_err=_ll_err(out["sample_log_likelihoods_lls"][q][::stride], g["ll_lls_strided"][j]),
                     nan_pattern_equal=_sha(np.isnan(ll)) == str(g["ll_dla_nan_sha"][j]),
                     inds_equal=_sha(inds.astype(np.int32)) == str(g["base_inds_sha"][j]),
                     masks_equal=(_sha(gp.ind.astype(np.uint8)) == str(g["ind_sha"][j])
                                  and _sha(gp.ind_unmasked.astype(np.uint8)) == str(g["ind_unmasked_sha"][j])),
                     ev_err=float(np.max(np.abs(out["log_likelihoods"][q] - g["log_likelihoods"][j]))),
                     prior_err=float(np.max(np.abs(out["log_priors"][q] - g["log_priors"][j]))),
                     post_err=float(np.max(np.abs(out["model_posteriors"][q] - g["model_posteriors"][j]))),
                     p_dla_err=float(abs(out["p_dlas"][q] - g["p_dla"][j])),
                     map_model_equal=int(np.argmax(out["model_posteriors"][q])) == int(np.argmax(g["model_posteriors"][j])),
                     map_params_equal=bool(np.array_equal(out["MAP_z_dlas"][q], g["MAP_z_dlas"][j], equal_nan=True)
                                           and np.array_equal(out["MAP_log_nhis"][q], g["MAP_log_nhis"][j], equal_nan=True)),
                     z_range_equal=bool(out["min_z_dlas"][q] == g["min_z_dla"][j] and out["max_z_dlas"][q] == g["max_z_dla"][j]),
                     n_equal=int(out["num_pixels"][q]) == int(g["num_pixels"][j]))
            golden_rows.append(r)
            for key in ("nan_pattern_equal", "inds_equal", "masks_equal", "map_model_equal", "map_params_equal",
                        "z_range_equal", "n_equal"):
                if not r[key]:
                    failures.append("live-reference golden, spectrum %d: %s is False" % (q, key))
            for key, tol in (("ll_err", LL_RTOL), ("ll_lls_err", LL_RTOL), ("ev_err", EV_ATOL), ("prior_err", EV_ATOL),
                             ("post_err", EV_ATOL), ("p_dla_err", EV_ATOL)):
                if not r[key] < tol:
                    failures.append("live-reference golden, spectrum %d: %s = %.3e" % (q, key, r[key]))

    models = np.array([r["map_model"] for r in rows])
    report = dict(
        workload="synthetic.make_workload(%d, seed0=0): S=%d, max_dlas=%d, num_lines=%d" % (num_spectra, S_SAMPLES, MAX_DLAS, NUM_LINES),
        num_spectra=num_spectra, z_qso_min=float(np.min(z_qsos)), z_qso_max=float(np.max(z_qsos)),
        n_min=int(np.min(out["num_pixels"])), n_max=int(np.max(out["num_pixels"])),
        tolerances=dict(ll_rel=LL_RTOL, evidence_abs=EV_ATOL),
        worst_vs_oracle=worst,
        all_base_sample_inds_identical=bool(all(r["inds_equal"] for r in rows)),
        all_nan_patterns_identical=bool(all(r["nan_pattern_equal"] for r in rows)),
        all_map_models_identical=bool(all(r["map_model_equal"] and r["dla_count_equal"] for r in rows)),
        all_map_parameters_identical=bool(all(r["map_params_equal"] for r in rows)),
        map_model_histogram={str(m): int(np.sum(models == m)) for m in range(2 + MAX_DLAS)},
        live_reference_goldens=dict(count=len(golden_rows),
                                    worst_ll_err=max([r["ll_err"] for r in golden_rows], default=None),
                                    worst_ev_err=max([r["ev_err"] for r in golden_rows], default=None),
                                    worst_post_err=max([r["post_err"] for r in golden_rows], default=None),
                                    all_indices_masks_maps_identical=bool(all(
                                        r["inds_equal"] and r["masks_equal"] and r["map_model_equal"] and r["map_params_equal"]
                                        for r in golden_rows))),
        failures=failures,
        seconds=dict(gpu_process=gpu_s, oracle_pool=oracle_s, oracle_workers=workers),
        per_spectrum=rows,
        per_golden=golden_rows,
    )
    if verbose:
        print(json.dumps({k: v for k, v in report.items() if k not in ("per_spectrum", "per_golden")}, indent=1))
    return report, failures


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--spectra", type=int, default=128)
    ap.add_argument("--workers", type=int, default=None)
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    report, failures = run_sweep(args.spectra, args.workers, args.batch)
    if args.out:
        os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
        with open(args.out, "w") as f:
            json.dump(report, f, indent=1)
    return 1 if failures else 0


if __name__ == "__main__":
    sys.exit(main())
