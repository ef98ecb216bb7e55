#!/usr/bin/env python
"""Small end-to-end case for compute-sanitizer: one short spectrum batch through every kernel of the path."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import __graft_entry__ as g
g.build()
from gpy_dla_detection_b200 import _lib, synthetic
from tests import helpers as H

_lib.init(0)
st = H.Setup(96)
z_qsos = np.array([2.3, 3.2, 4.6])
spectra = [synthetic.make_spectrum(st.model, z, seed=10 + i) for i, z in enumerate(z_qsos)]
proc = st.catalogue(3, True, batch_spectra=2)
out = proc.process(*proc.pack(spectra), z_qsos, keep_samples=True)
print("catalogue ok", out["p_dlas"], out["status"])
gp, sub, dla = st.gp_objects()
wl, fl, nv, pm = spectra[1]
for m in (gp, sub, dla):
    m.set_data(wl / (1 + 3.2), fl, nv, pm, 3.2)
np.random.seed(0)
print("class api ok", gp.log_model_evidence(), sub.log_model_evidences(1), dla.log_model_evidences(3))
from tests.test_gpu_zqso import build_zgp
model, zgp = build_zgp(24)
zgp.inference_z_qso(*synthetic.make_zqso_spectrum(model, 3.0, seed=1))
print("zqso ok", zgp.z_map)
