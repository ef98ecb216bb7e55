#!/usr/bin/env python
"""Soak test of the catalogue path: the same 1 000-spectrum workload REPEATS times; every output array of every
repeat must be bit-identical to the first (a race or a lost barrier phase in the kernels shows up as a difference or
a hang).  usage: soak.py [REPEATS] [SPECTRA]"""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import __graft_entry__ as g
g.build()
from gpy_dla_detection_b200 import _lib, synthetic
from gpy_dla_detection_b200.run_bayes_select import CatalogueProcessor
from gpy_dla_detection_b200.dla_samples import DLASamplesArrays
from gpy_dla_detection_b200.subdla_samples import SubDLASamplesArrays

repeats = int(sys.argv[1]) if len(sys.argv) > 1 else 10
Q = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
_lib.init(0)
params, model, prior, dla, sub, z_qsos, spectra = synthetic.make_workload(Q)
dla_s = DLASamplesArrays(params, prior, dla["offset_samples"], dla["log_nhi_samples"], dla["nhi_samples"])
sub_s = SubDLASamplesArrays(params, prior, sub["offset_samples"], sub["log_nhi_samples"], sub["nhi_samples"],
                            sub["Z_lls"], sub["Z_dla"])
proc = CatalogueProcessor(params, prior, model, dla_s, sub_s, 4, True, batch_spectra=128)
offsets, wl, fl, nv, pm = proc.pack(spectra)
first = None
t0 = time.perf_counter()
for r in range(repeats):
    out = proc.process(offsets, wl, fl, nv, pm, z_qsos, keep_samples=(Q <= 64))
    if first is None:
        first = {k: np.array(v, copy=True) for k, v in out.items() if isinstance(v, np.ndarray)}
        continue
    for k, v in first.items():
        if not np.array_equal(v, out[k], equal_nan=True):
            bad = np.argwhere(~((v == out[k]) | (np.isnan(v) & np.isnan(out[k]))))
            raise SystemExit("repeat %d: %s differs at %d entries, first %s" % (r, k, len(bad), bad[:3].tolist()))
print("soak ok: %d repeats x %d spectra bit-identical (%d arrays), %.1f s" % (repeats, Q, len(first), time.perf_counter() - t0))
