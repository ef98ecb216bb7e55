"""
GPU parity tests of the MCMC posterior hook: gpy_dla_detection_b200.log_posterior_mcmc
against golden values from the live reference's log_posterior_mcmc.py (tests/golden/make_golden.py).
"""
import numpy as np
import pytest

from gpy_dla_detection_b200 import synthetic
from tests import helpers as H

pytestmark = pytest.mark.gpu


def _args(g):
    st = H.Setup(int(g["S"]))
    _, _, dla = st.gp_objects()
    z_qso = float(g["z_qso"])
    wl, fl, nv, pm = synthetic.make_spectrum(st.model, z_qso, seed=int(g["seed"]))
    dla.set_data(wl / (1 + z_qso), fl, nv, pm, z_qso, build_model=True)
    pdf = lambda x: 0.5 + 0.1 * (x - 20.0)  # noqa: E731
    return dla, (dla.this_wavelengths, dla.y, dla.v, z_qso, float(g["min_z_dla"]), float(g["max_z_dla"]), 20.0, 23.0, pdf,
                 dla.padded_wavelengths, dla.this_mu, dla.this_M, dla.this_omega2, dla.pixel_mask, dla.ind_unmasked, 3)


def test_log_posterior_golden(gpu):
    from gpy_dla_detection_b200 import log_posterior_mcmc as M

    g = H.golden("mcmc_golden.npz")
    dla, args = _args(g)
    ref = g["log_posterior"]
    one_by_one = np.array([M.log_posterior(tuple(t), *args) for t in g["thetas"][:12]])
    batch = M.log_posteriors(g["thetas"], *args)
    assert np.array_equal(np.isfinite(batch), np.isfinite(ref))  # same prior support (strict inequalities)
    ok = np.isfinite(ref)
    assert H.ll_err(batch[ok], ref[ok]) < 1e-9
    assert np.array_equal(one_by_one, batch[:12])
    pair = M.sample_log_likelihood_k_dlas(np.array([g["thetas"][3, 0], g["thetas"][5, 0]]), 10 ** np.array([20.4, 21.1]),
                                          *args[1:3], *args[9:])
    assert abs(pair - float(g["pair_ll"])) < 1e-9 * abs(float(g["pair_ll"]))
    dmu, dM, dom = M.this_dla_gp(np.array([g["thetas"][3, 0]]), np.array([10**20.9]), *args[9:])
    assert np.max(np.abs(dmu - g["this_dla_mu"])) < 1e-13 and np.max(np.abs(dM - g["this_dla_M"])) < 1e-13
    assert np.max(np.abs(dom - g["this_dla_omega2"])) < 1e-13
    # the tuple DLAGP hands to emcee has the reference's layout
    a = dla.mcmc_log_posterior_args()
    assert len(a) == 16 and a[15] == 3 and a[9] is dla.padded_wavelengths
    assert np.isfinite(M.log_posteriors(g["thetas"][ok][:4], *a)).all()


def test_run_mcmc_needs_emcee(gpu):
    g = H.golden("mcmc_golden.npz")
    dla, _ = _args(g)
    try:
        import emcee  # noqa: F401
    except ImportError:
        with pytest.raises(ImportError):
            dla.run_mcmc(8, nsamples=2)
