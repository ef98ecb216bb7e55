#!/usr/bin/env python
"""Bring-up script: run each C-ABI entry on the GPU and print the error against the oracle."""
import os, sys, time, traceback
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import numpy as np
from scipy.special import wofz
import __graft_entry__ as g
g.build()
from gpy_dla_detection_b200 import _lib, synthetic, voigt, effective_optical_depth as eod
from gpy_dla_detection_b200.set_parameters import Parameters
from gpy_dla_detection_b200.null_gp import NullGP
from gpy_dla_detection_b200.dla_gp import DLAGP
from gpy_dla_detection_b200.subdla_gp import SubDLAGP
from gpy_dla_detection_b200.dla_samples import DLASamplesArrays
from gpy_dla_detection_b200.subdla_samples import SubDLASamplesArrays
from oracle import dla_oracle as O

_lib.init(0)
def step(name, fn):
    t = time.time()
    try:
        msg = fn()
        print("[ok ] %-28s %s  (%.2fs)" % (name, msg, time.time() - t), flush=True)
    except Exception:
        print("[ERR] %-28s" % name, flush=True)
        traceback.print_exc()

S = int(os.environ.get("DEV_S", "600"))
params = Parameters(num_dla_samples=S)
model = synthetic.make_learned_model(0)
prior = synthetic.SyntheticPrior(params)
d = synthetic.make_dla_sample_arrays(params); s = synthetic.make_subdla_sample_arrays(params)
z_qso = 3.0
wl, fl, nv, pm = synthetic.make_spectrum(model, z_qso, seed=7)
rest = params.emitted_wavelengths(wl, z_qso)
prep = O.prepare_spectrum(model, rest, fl, nv, pm, z_qso)
margs = (model["rest_wavelengths"], model["mu"], model["M"], model["log_omega"], model["log_c_0"], model["log_tau_0"], model["log_beta"])

def t_fadd():
    x = np.concatenate([np.linspace(0, 70, 20001), np.geomspace(64, 3e4, 2000)])
    worst = 0
    for y in (4.7e-4, 1.2e-4, 7e-8):
        a = voigt.faddeeva_re(x, y); b = np.real(wofz(x + 1j * y))
        worst = max(worst, np.max(np.abs(a - b) / b))
    return "max rel vs scipy wofz %.2e" % worst
step("faddeeva", t_fadd)

def t_voigt():
    out = []
    for br in (True, False):
        for (zd, ln, nl) in ((2.5, 20.3, 3), (2.9, 21.5, 5), (2.2, 19.6, 31)):
            a = voigt.voigt_absorption(prep["padded_wavelengths"], 10**ln, zd, nl, br)
            b = O.voigt_absorption(prep["padded_wavelengths"], 10**ln, zd, nl, br)
            out.append(np.max(np.abs(a - b)))
    return "max abs err " + " ".join("%.1e" % v for v in out)
step("voigt_absorption", t_voigt)

def t_eod():
    a = eod.effective_optical_depth(prep["this_wavelengths"], 3.65, 0.0023, z_qso, 31)
    b = O.effective_optical_depth(prep["this_wavelengths"], 3.65, 0.0023, z_qso, 31)
    return "max rel %.2e" % np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300))
step("effective_optical_depth", t_eod)

def t_mvn():
    y = np.array([1., 2.]); mu = np.array([1., 2.]); M = np.array([[2., 3, 1], [1, 2, 4]]); dd = np.ones(2) * 2
    a = NullGP.log_mvnpdf_low_rank(y, mu, M, dd); b = O.log_mvnpdf_low_rank(y, mu, M, dd)
    return "kat %.12f vs %.12f" % (a, b)
step("log_mvnpdf_low_rank", t_mvn)

gp = NullGP(params, prior, *margs)
def t_prep():
    gp.set_data(rest, fl, nv, pm, z_qso, build_model=True)
    msgs = []
    msgs.append("masks eq %s %s" % (np.array_equal(gp.ind, prep["ind"]), np.array_equal(gp.ind_unmasked, prep["ind_unmasked"])))
    for k in ("x", "y", "v", "this_wavelengths", "unmasked_wavelengths", "padded_wavelengths", "this_mu", "this_M", "this_omega2"):
        a = getattr(gp, k); b = prep[k]
        msgs.append("%s %.1e" % (k, np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300))))
    msgs.append("median %.3e" % abs(gp.normalization_median - prep["normalization_median"]))
    return " | ".join(msgs)
step("set_data", t_prep)

def t_null():
    a = gp.log_model_evidence(); b = O.null_log_model_evidence(prep)
    return "%.10f vs %.10f rel %.2e" % (a, b, abs(a - b) / abs(b))
step("null evidence", t_null)

dla = DLAGP(params, prior, DLASamplesArrays(params, prior, d["offset_samples"], d["log_nhi_samples"], d["nhi_samples"]), *margs)
sub = SubDLAGP(params, prior, SubDLASamplesArrays(params, prior, s["offset_samples"], s["log_nhi_samples"], s["nhi_samples"], s["Z_lls"], s["Z_dla"]), *margs)
def t_ll():
    dla.set_data(rest, fl, nv, pm, z_qso, build_model=True)
    zs = dla.dla_samples.sample_z_dlas(dla.this_wavelengths, z_qso)
    msgs = []
    for kd in (1, 2, 4):
        rng = np.random.default_rng(kd)
        idx = rng.integers(0, S, size=(200, kd))
        zz = zs[idx]; nn = d["nhi_samples"][idx]
        a = dla.sample_log_likelihoods_batch(zz, nn)
        b = np.array([O.sample_log_likelihood_k_dlas(prep, zz[i], nn[i]) for i in range(60)])
        msgs.append("k=%d rel %.2e" % (kd, np.max(np.abs(a[:60] - b) / np.abs(b))))
    ab = dla.this_dla_gp(zz[0], nn[0]); 
    aa = O.absorption_k_dlas(prep, zz[0], nn[0])
    msgs.append("this_dla_gp mu %.1e" % np.max(np.abs(ab[0] - prep["this_mu"] * aa)))
    return " | ".join(msgs)
step("sample_log_likelihoods", t_ll)

def t_resample():
    rng = np.random.default_rng(5)
    ok = []
    for n in (7, 100, 129, 1000, 10000, 30000):
        W = np.exp(-rng.exponential(8.0, n)); W[rng.random(n) < 0.1] = 0.0
        U = rng.random(n)
        out = np.empty(n, dtype=np.int32)
        _lib.check(_lib.load_library().dla_resample_indices(_lib.dptr(W), _lib.dptr(U), n, _lib.iptr(out)))
        ok.append(np.array_equal(out, O.resample_indices(W.copy(), U)))
    return "bit-identical to numpy: %s" % ok
step("resample_indices", t_resample)

def t_evid():
    np.random.seed(0)
    ev = dla.log_model_evidences(4)
    U = np.random.RandomState(0).random_sample((3, S))
    ref = O.log_model_evidences(prep, d["offset_samples"], d["nhi_samples"], 4, U)
    sl, so = dla.sample_log_likelihoods, ref["sample_log_likelihoods"]
    return "ev abs %.2e | ll rel %.2e | nan pattern %s | base inds eq %s (%d diff)" % (
        np.max(np.abs(ev - ref["log_likelihoods"])), np.nanmax(np.abs(sl - so) / np.abs(so)),
        np.array_equal(np.isnan(sl), np.isnan(so)), np.array_equal(dla.base_sample_inds, ref["base_sample_inds"]),
        np.sum(dla.base_sample_inds != ref["base_sample_inds"]))
step("log_model_evidences", t_evid)

def t_sub():
    sub.set_data(rest, fl, nv, pm, z_qso, build_model=True)
    ev = sub.log_model_evidences(1)
    ref = O.log_model_evidences(prep, s["offset_samples"], s["nhi_samples"], 1, None)
    return "ev abs %.2e" % np.max(np.abs(ev - ref["log_likelihoods"]))
step("subdla evidences", t_sub)

step("smoke", lambda: (g.smoke(), "done")[1])
print("kernel launches:", _lib.load_library().dla_kernel_launch_count())
