#!/usr/bin/env python
"""
bench_class_api.py : wall time of ONE spectrum through the drop-in class API (NullGP / SubDLAGP / DLAGP set_data x 3 +
BayesModelSelect.model_selection + maximum_a_posteriori), the per-spectrum loop body of run_bayes_select.py:141-230,
at the published size (S = 10 000, max_dlas = 4).  The batched catalogue engine is the throughput path (bench.py);
this is the latency of the reference-shaped per-object path.
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g  # noqa: E402

g.build()
from gpy_dla_detection_b200 import _lib, synthetic  # noqa: E402
from gpy_dla_detection_b200.bayesian_model_selection import BayesModelSelect  # noqa: E402
from gpy_dla_detection_b200.dla_gp import DLAGP  # noqa: E402
from gpy_dla_detection_b200.dla_samples import DLASamplesArrays  # noqa: E402
from gpy_dla_detection_b200.null_gp import NullGP  # noqa: E402
from gpy_dla_detection_b200.subdla_gp import SubDLAGP  # noqa: E402
from gpy_dla_detection_b200.subdla_samples import SubDLASamplesArrays  # noqa: E402

_lib.init(0)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 32
params, model, prior, dla, sub, z_qsos, spectra = synthetic.make_workload(N, 0)
margs = (model["rest_wavelengths"], model["mu"], model["M"], model["log_omega"], model["log_c_0"], model["log_tau_0"],
         model["log_beta"])
d = DLASamplesArrays(params, prior, dla["offset_samples"], dla["log_nhi_samples"], dla["nhi_samples"])
s = SubDLASamplesArrays(params, prior, sub["offset_samples"], sub["log_nhi_samples"], sub["nhi_samples"], sub["Z_lls"], sub["Z_dla"])
gp, sgp, dgp = NullGP(params, prior, *margs), SubDLAGP(params, prior, s, *margs), DLAGP(params, prior, d, *margs)
bayes = BayesModelSelect([0, 1, 4], 2)
times = []
for i in range(N):
    wl, fl, nv, pm = spectra[i]
    t0 = time.perf_counter()
    np.random.seed(0)
    rest = params.emitted_wavelengths(wl, z_qsos[i])
    for m in (gp, sgp, dgp):
        m.set_data(rest, fl, nv, pm, z_qsos[i], build_model=True)
    bayes.model_selection([gp, sgp, dgp], z_qsos[i])
    try:
        dgp.maximum_a_posteriori()
    except ValueError:
        pass
    times.append(time.perf_counter() - t0)
t = np.array(times[2:])
print(json.dumps({"path": "class API, one spectrum at a time (S=10000, max_dlas=4)", "spectra": int(t.size),
                  "ms_per_spectrum_median": 1e3 * float(np.median(t)), "ms_per_spectrum_mean": 1e3 * float(np.mean(t)),
                  "spectra_per_s": float(t.size / np.sum(t))}))
