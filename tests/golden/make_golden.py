#!/usr/bin/env python
"""
make_golden.py : generate tests/golden/*.npz by running the LIVE, UNMODIFIED reference
(/root/reference, imported with h5py/emcee stubbed - oracle/ref_loader.py) on seeded synthetic
inputs.  Run in the build container only (the GPU box has no /root/reference); the vectors
it writes are committed and are what pins the oracle and the CUDA path to the reference.

    python tests/golden/make_golden.py            # all fixtures (about 2 minutes)
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_loader  # noqa: E402

ref_loader.load_reference()
from gpy_dla_detection import voigt as rvoigt  # noqa: E402
from gpy_dla_detection.effective_optical_depth import effective_optical_depth as r_eod  # noqa: E402
from gpy_dla_detection.set_parameters import Parameters as RParameters  # noqa: E402
from gpy_dla_detection.null_gp import NullGP as RNullGP  # noqa: E402
from gpy_dla_detection.dla_gp import DLAGP as RDLAGP  # noqa: E402
from gpy_dla_detection.subdla_gp import SubDLAGP as RSubDLAGP  # noqa: E402
from gpy_dla_detection.bayesian_model_selection import BayesModelSelect as RBayes  # noqa: E402

from gpy_dla_detection_b200 import synthetic  # noqa: E402
from gpy_dla_detection_b200.set_parameters import Parameters  # noqa: E402


def golden_voigt():
    """voigt.voigt_absorption on a BOSS-like padded grid (a1) + effective optical depth (a2)."""
    loglam = 3.5523 + 1e-4 * np.arange(4650)
    wl = 10.0**loglam
    z_qso = 3.2
    sel = (wl / (1 + z_qso) >= 911.75) & (wl / (1 + z_qso) <= 1215.75)
    grid = wl[sel]
    cases = [(2.6, 20.3, 3, True), (3.05, 21.7, 3, True), (2.2, 19.6, 3, False), (2.9, 20.9, 5, True),
             (2.45, 22.4, 31, True), (3.1, 19.9, 31, False), (2.0, 20.0, 1, True)]
    out = {"wavelengths": grid, "cases": np.array(cases, dtype=np.float64)}
    for i, (zd, ln, nl, br) in enumerate(cases):
        out["profile_%d" % i] = rvoigt.voigt_absorption(grid, 10.0**ln, zd, num_lines=int(nl), broadening=bool(br))
    out["eod_kim"] = r_eod(grid, 3.65, 0.0023, z_qso, 31)
    out["eod_learned"] = r_eod(grid, 3.1, 0.0019, 2.9, 5)
    # Faddeeva real part at the arguments the profile visits (scipy wofz, the reference's own call)
    from scipy.special import wofz
    x = np.concatenate([np.linspace(0, 70, 1401), np.geomspace(70, 2.5e4, 400)])
    ys = rvoigt.gammas[[0, 1, 2, 4, 9, 30]] / (np.sqrt(2) * rvoigt.sigma)
    out["fadd_x"] = x
    out["fadd_y"] = ys
    out["fadd_re"] = np.stack([np.real(wofz(x + 1j * y)) for y in ys])
    # the reference's Lyman-series literals themselves (voigt.py:18-224)
    out["tables"] = np.stack([rvoigt.transition_wavelengths, rvoigt.oscillator_strengths, rvoigt.Gammas,
                              rvoigt.leading_constants, rvoigt.gammas])
    out["instrument_profile"] = rvoigt.instrument_profile
    out["sigma"] = rvoigt.sigma
    out["c"] = rvoigt.c
    np.savez_compressed(os.path.join(HERE, "voigt_golden.npz"), **out)
    print("voigt_golden.npz written")


def run_reference(S, z_qso, seed, max_dlas=4, num_lines=3, broadening=True):
    rp = RParameters(num_dla_samples=S, num_lines=num_lines)
    p = Parameters(num_dla_samples=S, num_lines=num_lines)
    model = synthetic.make_learned_model(0)
    prior = synthetic.SyntheticPrior(p)
    dla = synthetic.make_dla_sample_arrays(p)
    sub = synthetic.make_subdla_sample_arrays(p)
    wl, fl, nv, pm = synthetic.make_spectrum(model, z_qso, seed=seed)
    rest = rp.emitted_wavelengths(wl, z_qso)
    margs = (model["rest_wavelengths"], model["mu"], model["M"], model["log_omega"], model["log_c_0"],
             model["log_tau_0"], model["log_beta"])
    gp = RNullGP(rp, prior, *margs)
    dgp = RDLAGP(rp, prior, ref_loader.RefDLASamples(rp, dla), *margs, broadening=broadening)
    sgp = RSubDLAGP(rp, prior, ref_loader.RefDLASamples(rp, sub, True), *margs, broadening=broadening)
    for m in (gp, dgp, sgp):
        m.set_data(rest, fl, nv, pm, z_qso, build_model=True)
    np.random.seed(0)  # run_bayes_select.py:144
    bayes = RBayes([0, 1, max_dlas], 2)
    t0 = time.time()
    log_post = bayes.model_selection([gp, sgp, dgp], z_qso)
    dt = time.time() - t0
    try:
        map_z, map_n = dgp.maximum_a_posteriori()
    except ValueError:
        map_z = map_n = np.full((max_dlas, max_dlas), np.nan)
    # a few single-sample likelihoods and one this_dla_gp through the reference's own entry points
    zs = dgp.dla_samples.sample_z_dlas(dgp.this_wavelengths, z_qso)
    pick = np.array([0, 1, S // 3, S // 2, S - 1])
    single = np.array([dgp.sample_log_likelihood_k_dlas(np.array([zs[i]]), np.array([dla["nhi_samples"][i]])) for i in pick])
    pair = np.array([dgp.sample_log_likelihood_k_dlas(np.array([zs[i], zs[(i * 7 + 3) % S]]),
                                                      np.array([dla["nhi_samples"][i], dla["nhi_samples"][(i * 7 + 3) % S]]))
                     for i in pick])
    dmu, dM, dom = dgp.this_dla_gp(np.array([zs[pick[2]], zs[pick[3]]]),
                                   np.array([dla["nhi_samples"][pick[2]], dla["nhi_samples"][pick[3]]]))
    out = dict(
        S=S, z_qso=z_qso, seed=seed, max_dlas=max_dlas, num_lines=num_lines, broadening=broadening,
        wavelengths=wl, flux=fl, noise_variance=nv, pixel_mask=pm,
        x=gp.x, y=gp.y, v=gp.v, ind=gp.ind, ind_unmasked=gp.ind_unmasked,
        this_wavelengths=gp.this_wavelengths, unmasked_wavelengths=gp.unmasked_wavelengths,
        padded_wavelengths=gp.padded_wavelengths, this_mu=gp.this_mu, this_M=gp.this_M, this_omega2=gp.this_omega2,
        normalization_median=gp.normalization_median,
        log_priors=bayes.log_priors, log_likelihoods=bayes.log_likelihoods, log_posteriors=log_post,
        model_posteriors=bayes.model_posteriors, p_dla=bayes.p_dla, p_no_dla=bayes.p_no_dla,
        sample_log_likelihoods_dla=dgp.sample_log_likelihoods, base_sample_inds=dgp.base_sample_inds,
        sample_log_likelihoods_lls=sgp.sample_log_likelihoods[:, 0],
        MAP_z_dlas=map_z, MAP_log_nhis=map_n,
        min_z_dla=rp.min_z_dla(wl, z_qso), max_z_dla=rp.max_z_dla(wl, z_qso),
        sample_z_dlas=zs, pick=pick, single_ll=single, pair_ll=pair, this_dla_mu=dmu, this_dla_M=dM,
        this_dla_omega2=dom, prior_counts=np.array(prior.less_ind(z_qso), dtype=np.float64),
        reference_seconds=dt,
    )
    return out


def golden_spectra():
    cases = [
        ("spec_S256_z2p3", dict(S=256, z_qso=2.3, seed=11)),
        ("spec_S300_z3p0", dict(S=300, z_qso=3.0, seed=7)),
        ("spec_S256_z4p4", dict(S=256, z_qso=4.4, seed=23)),
        ("spec_S200_z2p8_lines5_nobroad", dict(S=200, z_qso=2.8, seed=5, num_lines=5, broadening=False)),
        ("spec_S200_z3p4_max2", dict(S=200, z_qso=3.4, seed=31, max_dlas=2)),
    ]
    for name, kw in cases:
        out = run_reference(**kw)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, "reference time %.1f s" % out["reference_seconds"], "p_dla", out["p_dla"])


def golden_full():
    """One spectrum at the full published size: S = 10 000, max_dlas = 4 (config 1 of BASELINE.json)."""
    out = run_reference(S=10000, z_qso=2.9, seed=3)
    # inputs are regenerated from the seeds by the tests; keep outputs only (float32 would lose parity)
    keep = ("S", "z_qso", "seed", "max_dlas", "num_lines", "broadening", "log_priors", "log_likelihoods",
            "log_posteriors", "model_posteriors", "p_dla", "sample_log_likelihoods_dla", "base_sample_inds",
            "sample_log_likelihoods_lls", "MAP_z_dlas", "MAP_log_nhis", "min_z_dla", "max_z_dla", "ind",
            "ind_unmasked", "this_mu", "this_omega2", "normalization_median", "prior_counts", "reference_seconds",
            "wavelengths", "flux", "noise_variance", "pixel_mask")
    np.savez_compressed(os.path.join(HERE, "spec_S10000_z2p9_full.npz"), **{k: out[k] for k in keep})
    print("full: reference time %.1f s" % out["reference_seconds"], "p_dla", out["p_dla"])


def golden_zqso():
    """ZGP (zqso_gp.py): inference_z_qso over z samples, set_data attributes and evidence at single redshifts."""
    from gpy_dla_detection.zqso_gp import ZGP as RZGP
    from gpy_dla_detection.zqso_set_parameters import ZParameters as RZParameters
    from gpy_dla_detection.zqso_samples import ZSamples as RZSamples

    model = synthetic.make_zqso_model(0)

    def build(num):
        rp = RZParameters(num_zqso_samples=num)
        gp = RZGP(rp, RZSamples(rp), model["rest_wavelengths"], model["mu"], model["M"], model["bluewards_mu"],
                  model["redwards_mu"], model["bluewards_sigma"], model["redwards_sigma"])
        return gp

    out = {}
    cases = [(2.3, 11), (3.4, 12), (4.9, 13)]
    out["cases"] = np.array(cases, dtype=np.float64)
    for c, (z_true, seed) in enumerate(cases):
        wl, fl, nv, pm = synthetic.make_zqso_spectrum(model, z_true, seed=seed)
        gp = build(96)
        gp.inference_z_qso(wl, fl, nv, pm)
        out["ll_%d" % c] = gp.sample_log_likelihoods
        out["z_map_%d" % c] = gp.z_map
        # a narrower prior volume around the truth (other arguments of inference_z_qso)
        gp.inference_z_qso(wl, fl, nv, pm, z_qso_min=z_true - 0.05, z_qso_max=z_true + 0.05)
        out["ll_narrow_%d" % c] = gp.sample_log_likelihoods
        out["z_map_narrow_%d" % c] = gp.z_map
        # attributes after set_data at one redshift + the three parts of the evidence
        gp.set_data(wl, fl, nv, pm, z_qso=z_true + 0.013, normalize=True, build_model=True)
        for k in ("x", "y", "v", "this_wavelengths", "this_mu", "this_M", "y_bw", "v_bw", "y_rw", "v_rw", "ind"):
            out["%s_%d" % (k, c)] = getattr(gp, k)
        out["evidence_%d" % c] = gp.log_model_evidence()
    # log_mvnpdf_iid known answers (zqso_gp.py:252-278)
    rng = np.random.default_rng(0)
    y, mu, d = rng.standard_normal(50), rng.standard_normal(50), 0.1 + rng.random(50)
    out["iid_y"], out["iid_mu"], out["iid_d"] = y, mu, d
    out["iid_value"] = RZGP.log_mvnpdf_iid(y, mu, d)
    np.savez_compressed(os.path.join(HERE, "zqso_golden.npz"), **out)
    print("zqso_golden.npz written")

    # the full published size: 10 000 z samples on one spectrum
    wl, fl, nv, pm = synthetic.make_zqso_spectrum(model, 2.75, seed=21)
    gp = build(10000)
    t0 = time.time()
    gp.inference_z_qso(wl, fl, nv, pm)
    dt = time.time() - t0
    np.savez_compressed(os.path.join(HERE, "zqso_full_S10000.npz"), z_true=2.75, seed=21,
                        sample_log_likelihoods=gp.sample_log_likelihoods, z_map=gp.z_map, reference_seconds=dt)
    print("zqso_full_S10000.npz written, reference time %.1f s, z_map %.4f" % (dt, gp.z_map))


def golden_mcmc():
    """log_posterior_mcmc.py free functions (the emcee target of DLAGP.run_mcmc) on one prepared spectrum."""
    from gpy_dla_detection import log_posterior_mcmc as rm

    S, z_qso, seed = 64, 3.0, 7
    rp = RParameters(num_dla_samples=S)
    p = Parameters(num_dla_samples=S)
    model = synthetic.make_learned_model(0)
    prior = synthetic.SyntheticPrior(p)
    dla = synthetic.make_dla_sample_arrays(p)
    wl, fl, nv, pm = synthetic.make_spectrum(model, z_qso, seed=seed)
    margs = (model["rest_wavelengths"], model["mu"], model["M"], model["log_omega"], model["log_c_0"],
             model["log_tau_0"], model["log_beta"])
    gp = RDLAGP(rp, prior, ref_loader.RefDLASamples(rp, dla), *margs)
    gp.set_data(rp.emitted_wavelengths(wl, z_qso), fl, nv, pm, z_qso, build_model=True)
    lo, hi = rp.min_z_dla(gp.this_wavelengths, z_qso), rp.max_z_dla(gp.this_wavelengths, z_qso)
    pdf = lambda x: 0.5 + 0.1 * (x - 20.0)  # noqa: E731  (any positive function of log N_HI)
    args = (gp.this_wavelengths, gp.y, gp.v, z_qso, lo, hi, 20.0, 23.0, pdf, gp.padded_wavelengths, gp.this_mu, gp.this_M,
            gp.this_omega2, gp.pixel_mask, gp.ind_unmasked, 3)
    rng = np.random.default_rng(4)
    thetas = np.stack([rng.uniform(lo - 0.05, hi + 0.05, 40), rng.uniform(19.8, 23.2, 40)], axis=1)
    post = np.array([rm.log_posterior(tuple(t), *args) for t in thetas])
    pair = rm.sample_log_likelihood_k_dlas(np.array([thetas[3, 0], thetas[5, 0]]), 10 ** np.array([20.4, 21.1]), *args[1:3],
                                           *args[9:])
    dmu, dM, dom = rm.this_dla_gp(np.array([thetas[3, 0]]), np.array([10**20.9]), *args[9:])
    np.savez_compressed(os.path.join(HERE, "mcmc_golden.npz"), S=S, z_qso=z_qso, seed=seed, thetas=thetas,
                        log_posterior=post, pair_ll=pair, this_dla_mu=dmu, this_dla_M=dM, this_dla_omega2=dom,
                        min_z_dla=lo, max_z_dla=hi)
    print("mcmc_golden.npz written;", int(np.isfinite(post).sum()), "of 40 thetas inside the prior")


def golden_lls():
    """voigt_lls.voigt_absorption and a DLAGP whose this_dla_gp uses it (the pattern of examples/gp_find_lls.py:159-224)."""
    from gpy_dla_detection import voigt_lls as rl

    loglam = 3.5523 + 1e-4 * np.arange(4650)
    wl_all = 10.0**loglam
    z_qso = 3.6
    sel = (wl_all / (1 + z_qso) >= 850.0) & (wl_all / (1 + z_qso) <= 1215.75)
    grid = wl_all[sel]  # reaches bluewards of the absorbers' Lyman limit
    cases = [(3.3, 17.5, 3, True), (3.5, 19.0, 4, True), (3.1, 20.6, 5, False), (3.55, 18.2, 31, True)]
    out = {"wavelengths": grid, "cases": np.array(cases, dtype=np.float64)}
    for i, (zl, ln, nl, br) in enumerate(cases):
        out["profile_%d" % i] = rl.voigt_absorption(grid, 10.0**ln, zl, num_lines=int(nl), broadening=bool(br))
        out["tau_%d" % i] = rl.tau_LLS_break(grid, 10.0**ln, zl)

    class RLLSGP(RDLAGP):
        def this_dla_gp(self, z_dlas, nhis):
            k_dlas = len(z_dlas)
            mask_ind = ~self.pixel_mask[self.ind_unmasked]
            absorption = rl.voigt_absorption(self.padded_wavelengths, z_lls=z_dlas[0], nhi=nhis[0],
                                             num_lines=self.params.num_lines)
            for j in range(1, k_dlas):
                absorption = absorption * rl.voigt_absorption(self.padded_wavelengths, z_lls=z_dlas[j], nhi=nhis[j],
                                                              num_lines=self.params.num_lines)
            absorption = absorption[mask_ind]
            return self.this_mu * absorption, self.this_M * absorption[:, None], self.this_omega2 * absorption**2

    S, z_q, seed = 128, 4.2, 17
    rp = RParameters(num_dla_samples=S, num_lines=4)
    p = Parameters(num_dla_samples=S, num_lines=4)
    model = synthetic.make_learned_model(0)
    prior = synthetic.SyntheticPrior(p)
    dla = synthetic.make_dla_sample_arrays(p)
    wl, fl, nv, pm = synthetic.make_spectrum(model, z_q, seed=seed)
    margs = (model["rest_wavelengths"], model["mu"], model["M"], model["log_omega"], model["log_c_0"],
             model["log_tau_0"], model["log_beta"])
    gp = RLLSGP(rp, prior, ref_loader.RefDLASamples(rp, dla), *margs)
    gp.set_data(rp.emitted_wavelengths(wl, z_q), fl, nv, pm, z_q, build_model=True)
    np.random.seed(0)
    ev = gp.log_model_evidences(3)
    out.update(S=S, z_qso=z_q, seed=seed, log_evidences=ev, sample_log_likelihoods=gp.sample_log_likelihoods,
               base_sample_inds=gp.base_sample_inds)
    # the extended model of examples/gp_find_lls.py:102,162-170: rest grid 850.75-1420.75 A (bluewards of the Lyman limit,
    # redwards of Ly-alpha), normalisation window 1425-1475 A, absorbers from log N = 17
    grid_kw = dict(min_lambda=850.75, max_lambda=1420.75, normalization_min_lambda=1425.0, normalization_max_lambda=1475.0)
    S2, z_q2, seed2 = 96, 3.9, 23
    rp2 = RParameters(num_dla_samples=S2, num_lines=4, **grid_kw)
    p2 = Parameters(num_dla_samples=S2, num_lines=4, **grid_kw)
    model2 = synthetic.make_learned_model(1, rest_min=850.75, rest_max=1420.75)
    prior2 = synthetic.SyntheticPrior(p2)
    lls2 = synthetic.make_lls_sample_arrays(S2)
    wl2, fl2, nv2, pm2 = synthetic.make_spectrum(model2, z_q2, seed=seed2, params=p2)
    margs2 = (model2["rest_wavelengths"], model2["mu"], model2["M"], model2["log_omega"], model2["log_c_0"],
              model2["log_tau_0"], model2["log_beta"])
    gp2 = RLLSGP(rp2, prior2, ref_loader.RefDLASamples(rp2, lls2), *margs2)
    gp2.set_data(rp2.emitted_wavelengths(wl2, z_q2), fl2, nv2, pm2, z_q2, build_model=True)
    np.random.seed(0)
    ev2 = gp2.log_model_evidences(2)
    out.update(ext_S=S2, ext_z_qso=z_q2, ext_seed=seed2, ext_log_evidences=ev2,
               ext_sample_log_likelihoods=gp2.sample_log_likelihoods, ext_base_sample_inds=gp2.base_sample_inds,
               ext_this_mu=gp2.this_mu, ext_this_omega2=gp2.this_omega2, ext_ind=gp2.ind, ext_x=gp2.x,
               ext_normalization_median=gp2.normalization_median)
    np.savez_compressed(os.path.join(HERE, "lls_golden.npz"), **out)
    print("lls_golden.npz written", ev, ev2, "n =", gp2.x.shape[0])


# 20 of the 128 spectra tools/parity_sweep.py runs: every 8th, plus the four highest redshifts of the draw
SWEEP_INDICES = tuple(sorted(set(range(0, 128, 8)) | set(int(i) for i in np.argsort(synthetic.sample_z_qsos(128, seed=12345))[-4:])))
SWEEP_STRIDE = 8                         # every 8th sample log-likelihood is kept


def _sha(a):
    import hashlib

    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def _sweep_one(i):
    """One spectrum of the bench workload (bench.py / synthetic.make_workload, seed0 = 0) through the live reference."""
    z_qsos = synthetic.sample_z_qsos(128, seed=12345)
    out = run_reference(S=10000, z_qso=float(z_qsos[i]), seed=i)
    ll = out["sample_log_likelihoods_dla"]
    inds = out["base_sample_inds"]
    return dict(
        index=i, z_qso=out["z_qso"], reference_seconds=out["reference_seconds"],
        log_priors=out["log_priors"], log_likelihoods=out["log_likelihoods"], log_posteriors=out["log_posteriors"],
        model_posteriors=out["model_posteriors"], p_dla=out["p_dla"], p_no_dla=out["p_no_dla"],
        MAP_z_dlas=out["MAP_z_dlas"], MAP_log_nhis=out["MAP_log_nhis"],
        min_z_dla=out["min_z_dla"], max_z_dla=out["max_z_dla"],
        ll_dla_strided=ll[::SWEEP_STRIDE], ll_lls_strided=out["sample_log_likelihoods_lls"][::SWEEP_STRIDE],
        ll_dla_nan_sha=_sha(np.isnan(ll)), ll_dla_nan_count=int(np.isnan(ll).sum()),
        base_inds_sha=_sha(inds.astype(np.int32)), base_inds_head=inds[:, :64].astype(np.int32),
        base_inds_tail=inds[:, -64:].astype(np.int32),
        ind_sha=_sha(out["ind"].astype(np.uint8)), ind_unmasked_sha=_sha(out["ind_unmasked"].astype(np.uint8)),
        num_pixels=int(out["x"].shape[0]),
    )


def golden_bench_sweep(workers=None):
    """
    Compact live-reference results for 16 spectra OF THE BENCH WORKLOAD at the full published size
    (S = 10 000, max_dlas = 4): evidences, posteriors, MAP arrays, every 8th sample log-likelihood,
    the NaN pattern's hash, and hash + head/tail of base_sample_inds (VERDICT r1 task 1).
    About 50 s per spectrum per core.
    """
    import multiprocessing as mp

    workers = workers or min(os.cpu_count() or 1, len(SWEEP_INDICES))
    with mp.get_context("fork").Pool(workers) as pool:
        rows = pool.map(_sweep_one, SWEEP_INDICES, chunksize=1)
    out = {"indices": np.array(SWEEP_INDICES), "stride": SWEEP_STRIDE}
    for key in rows[0]:
        if key == "index":
            continue
        vals = [r[key] for r in rows]
        out[key] = np.array(vals) if not isinstance(vals[0], str) else np.array(vals, dtype="U64")
    np.savez_compressed(os.path.join(HERE, "bench_sweep_S10000.npz"), **out)
    print("bench_sweep_S10000.npz written; reference seconds per spectrum: %.1f mean" % float(np.mean(out["reference_seconds"])))
    print("p_dla:", np.round(out["p_dla"], 4))


ZQSO_SWEEP_INDICES = (0, 1, 2, 3, 5, 8, 13, 21)  # spectra of bench.py's configs[4] workload (make_zqso_workload(Q, 0))


def _zqso_sweep_one(i):
    model, z_true, spectra = synthetic.make_zqso_workload(max(ZQSO_SWEEP_INDICES) + 1, 0)
    t0 = time.time()
    ll, z_map = ref_loader.run_reference_zqso(model, spectra[i], 10000)
    return dict(index=i, z_true=float(z_true[i]), z_map=float(z_map), ll_strided=ll[::SWEEP_STRIDE], ll_sha_nan=_sha(np.isnan(ll)),
                argmax=int(np.nanargmax(ll)), ll_max=float(np.nanmax(ll)), reference_seconds=time.time() - t0)


def golden_zqso_sweep(workers=None):
    """
    ZGP.inference_z_qso of the live reference (10 000 z samples) on 8 spectra OF THE BENCH WORKLOAD of configs[4]:
    every 8th sample log-likelihood, the NaN pattern's hash, the argmax and z_MAP.  About 30-60 s per spectrum per core.
    """
    import contextlib
    import io
    import multiprocessing as mp

    workers = workers or min(os.cpu_count() or 1, len(ZQSO_SWEEP_INDICES))
    with contextlib.redirect_stdout(io.StringIO()):
        with mp.get_context("fork").Pool(workers) as pool:
            rows = pool.map(_zqso_sweep_one, ZQSO_SWEEP_INDICES, chunksize=1)
    out = {"indices": np.array(ZQSO_SWEEP_INDICES), "stride": SWEEP_STRIDE}
    for key in rows[0]:
        if key == "index":
            continue
        vals = [r[key] for r in rows]
        out[key] = np.array(vals) if not isinstance(vals[0], str) else np.array(vals, dtype="U64")
    np.savez_compressed(os.path.join(HERE, "zqso_bench_sweep_S10000.npz"), **out)
    print("zqso_bench_sweep_S10000.npz written; z_true", np.round(out["z_true"], 3), "z_map", np.round(out["z_map"], 3),
          "seconds", np.round(out["reference_seconds"], 1))


if __name__ == "__main__":
    if "--only-zqso-sweep" in sys.argv:
        golden_zqso_sweep()
        sys.exit(0)
    if "--only-sweep" in sys.argv:
        golden_bench_sweep()
        sys.exit(0)
    if "--only-lls" in sys.argv:
        golden_lls()
        sys.exit(0)
    if "--only-mcmc" in sys.argv:
        golden_mcmc()
        sys.exit(0)
    if "--only-zqso" in sys.argv:
        golden_zqso()
        sys.exit(0)
    if "--only-voigt" in sys.argv:
        golden_voigt()
        sys.exit(0)
    golden_voigt()
    golden_spectra()
    golden_zqso()
    golden_mcmc()
    golden_lls()
    if "--no-full" not in sys.argv:
        golden_full()
