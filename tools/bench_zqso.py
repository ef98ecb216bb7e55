#!/usr/bin/env python
"""Throughput of the zQSO sweep (BASELINE.json configs[4] shape: 10 000 z samples per spectrum) - secondary workload."""
import json, os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import __graft_entry__ as g
g.build()
from gpy_dla_detection_b200 import _lib, synthetic
from gpy_dla_detection_b200.zqso_gp import ZGP
from gpy_dla_detection_b200.zqso_samples import ZSamples
from gpy_dla_detection_b200.zqso_set_parameters import ZParameters

_lib.init(0)
Q = int(sys.argv[1]) if len(sys.argv) > 1 else 64
S = 10000
model = synthetic.make_zqso_model(0)
p = ZParameters(num_zqso_samples=S)
gp = ZGP(p, ZSamples(p), model["rest_wavelengths"], model["mu"], model["M"], model["bluewards_mu"], model["redwards_mu"],
         model["bluewards_sigma"], model["redwards_sigma"])
zq = synthetic.sample_z_qsos(Q, seed=5)
spectra = [synthetic.make_zqso_spectrum(model, float(z), seed=i) for i, z in enumerate(zq)]
zs = ZSamples(p).sample_z_qsos()
lib = _lib.load_library()
for it in range(3):
    t0 = time.perf_counter()
    out = gp.inference_z_qsos(spectra, zs, keep_samples=False)
    wall = time.perf_counter() - t0
    kms = lib.dla_last_kernel_ms()
err = np.abs(out["z_map"] - zq)
print(json.dumps({"workload": "zqso sweep, %d spectra x %d redshift samples" % (Q, S), "kernel_ms": kms, "wall_ms": 1e3 * wall,
                  "spectra_per_s_kernel": Q / (kms * 1e-3), "spectra_per_s_e2e": Q / wall,
                  "frac_within_0.05_of_truth": float(np.mean(err < 0.05))}))
