"""
CPU tests of the host-side mirror of the reference interface: parameters, priors, sample
objects, model-selection arithmetic, catalogue packing, and the multi-GPU sharding + gather
(world_size 2 over gloo).  No compute kernel is called here.
"""
import os
import subprocess
import sys

import numpy as np
import pytest

from gpy_dla_detection_b200 import synthetic
from gpy_dla_detection_b200.run_bayes_select import CatalogueProcessor, log_priors_for, shard_range
from gpy_dla_detection_b200.set_parameters import Parameters
from oracle import dla_oracle as O
from tests import helpers as H

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_parameters_defaults_match_reference():
    """set_parameters.py:21-102"""
    p = Parameters()
    assert (p.min_lambda, p.max_lambda, p.dlambda, p.k) == (911.75, 1215.75, 0.25, 20)
    assert (p.num_dla_samples, p.num_lines, p.num_forest_lines, p.width) == (10000, 3, 31, 3)
    assert p.pixel_spacing == 1e-4 and p.lya_wavelength == 1215.6701 and p.lyman_limit == 911.7633
    assert p.normalization_min_lambda == 1310 and p.normalization_max_lambda == 1325
    assert p.kms_to_z(3000.0) == 3000.0 * 1000 / 299792458
    assert Parameters(num_dla_samples=77, num_lines=31).num_dla_samples == 77


def test_z_dla_range_matches_oracle_and_golden():
    p = Parameters()
    for path in H.small_spectrum_fixtures():
        g = np.load(path)
        z = float(g["z_qso"])
        assert p.min_z_dla(g["wavelengths"], z) == float(g["min_z_dla"]) == O.z_dla_range(g["wavelengths"], z)[0]
        assert p.max_z_dla(g["wavelengths"], z) == float(g["max_z_dla"]) == O.z_dla_range(g["wavelengths"], z)[1]


def test_shard_range_partitions():
    for n in (0, 1, 7, 8, 1000, 160000):
        for w in (1, 2, 3, 4, 8):
            ranges = [shard_range(n, r, w) for r in range(w)]
            assert ranges[0][0] == 0 and ranges[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
            sizes = [b - a for a, b in ranges]
            assert max(sizes) - min(sizes) <= 1


def test_log_priors_vectorised_equals_per_spectrum():
    """dla_gp.py:398-426, subdla_gp.py:311-346 through the vectorised host path"""
    p = Parameters()
    prior = synthetic.SyntheticPrior(p)
    sub = synthetic.make_subdla_sample_arrays(Parameters(num_dla_samples=64))
    z = np.array([2.2, 2.5, 3.0, 3.9, 5.1])
    got = log_priors_for(prior, z, 4, sub["Z_lls"], sub["Z_dla"])
    assert got.shape == (5, 6) and np.all(np.isnan(got[:, 0]))
    for q, zq in enumerate(z):
        m, n = prior.less_ind(zq)
        assert np.array_equal(got[q, 2:], O.dla_log_priors(m, n, 4))
        assert np.allclose(got[q, 1], O.dla_log_priors(m, n, 1, sub["Z_lls"] / sub["Z_dla"])[0], rtol=1e-15)

    class Duck:  # a prior without the vectorisable attributes goes through less_ind
        def less_ind(self, zq):
            return prior.less_ind(zq)

    assert np.array_equal(log_priors_for(Duck(), z, 4, sub["Z_lls"], sub["Z_dla"]), got, equal_nan=True)


def test_model_selection_arithmetic_matches_golden():
    """bayesian_model_selection.py:48-149 as restated by the oracle vs the live reference's numbers"""
    for path in H.small_spectrum_fixtures():
        g = np.load(path)
        md = int(g["max_dlas"])
        ll = g["log_likelihoods"]
        sel = O.model_selection(g["log_priors"][1:2], g["log_priors"][2:], ll[0], ll[1:2], ll[2:])
        assert np.max(np.abs(sel["log_posteriors"] - g["log_posteriors"])) < 1e-12
        assert np.max(np.abs(sel["model_posteriors"] - g["model_posteriors"])) < 1e-12
        assert abs(sel["p_dla"] - float(g["p_dla"])) < 1e-12 and len(ll) == 2 + md


def test_sample_objects_expose_the_reference_attributes():
    st = H.Setup(128)
    d, s = st.sample_objects()
    assert d.offset_samples.shape == d.log_nhi_samples.shape == d.nhi_samples.shape == (128,)
    assert np.array_equal(d.nhi_samples, 10.0 ** d.log_nhi_samples)
    assert s._Z_lls > 0 and s._Z_dla > 0
    wl = 10 ** (3.6 + 1e-4 * np.arange(1000))
    z = d.sample_z_dlas(wl, 2.9)
    lo, hi = st.params.min_z_dla(wl, 2.9), st.params.max_z_dla(wl, 2.9)
    assert np.array_equal(z, lo + (hi - lo) * d.offset_samples)  # dla_samples.py:94-104
    assert np.array_equal(s.sample_z_lls(wl, 2.9), lo + (hi - lo) * s.offset_samples)
    assert np.all((d.log_nhi_samples >= 20) & (d.log_nhi_samples <= 23.1))
    assert np.all((s.log_nhi_samples >= 19.5) & (s.log_nhi_samples <= 20))


def test_pack_is_ragged_and_lossless():
    model = synthetic.make_learned_model(0)
    a = synthetic.make_spectrum(model, 2.5, seed=1)
    b = tuple(x[100:3000] for x in synthetic.make_spectrum(model, 3.5, seed=2))
    offsets, wl, fl, nv, pm = CatalogueProcessor.pack([a, b])
    assert list(offsets) == [0, len(a[0]), len(a[0]) + len(b[0])]
    assert offsets.dtype == np.int64 and pm.dtype == np.uint8 and wl.dtype == np.float64
    assert np.array_equal(wl[offsets[1]:], b[0]) and np.array_equal(pm[: offsets[1]].astype(bool), a[3])
    assert np.array_equal(nv[: offsets[1]], a[2], equal_nan=True)


def test_synthetic_workload_is_seeded():
    model = synthetic.make_learned_model(0)
    assert model["M"].shape == (1217, 20) and model["rest_wavelengths"][0] == 911.75
    a = synthetic.make_spectrum(model, 2.7, seed=5)
    b = synthetic.make_spectrum(model, 2.7, seed=5)
    for x, y in zip(a, b):
        assert np.array_equal(x, y, equal_nan=True)
    assert np.array_equal(synthetic.sample_z_qsos(10, seed=3), synthetic.sample_z_qsos(10, seed=3))


_WORKER = r"""
import os, sys
sys.path.insert(0, {root!r})
import numpy as np
import torch.distributed as dist
from gpy_dla_detection_b200.run_bayes_select import process_qso_sharded, shard_range

dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
Q = {Q}
seen = []

def fake_process(qso_list, z_list, read_spec, max_dlas, broadening, **kw):
    # stands in for the per-GPU engine: results are a function of the global spectrum index only
    seen.extend(qso_list)
    idx = np.array(qso_list, dtype=np.int64)
    return dict(p_dlas=np.sin(idx.astype(np.float64)), z_qsos=np.asarray(z_list, dtype=np.float64),
                MAP_z_dlas=np.tile(idx[:, None, None].astype(np.float64), (1, max_dlas, max_dlas)),
                base_sample_inds=np.tile(idx[:, None, None].astype(np.int32), (1, 5, max_dlas - 1)),
                note="not an array")

out = process_qso_sharded(list(range(Q)), [2.0 + 0.01 * i for i in range(Q)], None, 4, True, process_fn=fake_process)
a, b = shard_range(Q, rank, world)
assert seen == list(range(a, b)), (rank, seen)
if rank == 0:
    idx = np.arange(Q)
    assert sorted(out) == ["MAP_z_dlas", "base_sample_inds", "p_dlas", "z_qsos"]
    assert np.array_equal(out["p_dlas"], np.sin(idx.astype(np.float64)))
    assert out["MAP_z_dlas"].shape == (Q, 4, 4) and np.array_equal(out["MAP_z_dlas"][:, 0, 0], idx)
    assert out["base_sample_inds"].dtype == np.int32 and np.array_equal(out["base_sample_inds"][:, 2, 1], idx)
    print("GATHER_OK", Q, world)
else:
    assert out is None
dist.barrier()
dist.destroy_process_group()
"""


@pytest.mark.parametrize("Q", [7, 1])
def test_sharded_catalogue_gather_world_size_2_gloo(tmp_path, Q):
    """two ranks over gloo: block partition of the spectrum list, gather in spectrum order on rank 0"""
    script = tmp_path / "worker.py"
    script.write_text(_WORKER.format(root=ROOT, Q=Q))
    port = 29500 + (os.getpid() % 2000) + Q
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
           "127.0.0.1", "--master-port", str(port), str(script)]
    env = dict(os.environ, OMP_NUM_THREADS="1")
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=240, env=env)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
    assert "GATHER_OK %d 2" % Q in res.stdout


def test_patch_reference_installs_and_restores():
    """interop.patch_reference swaps the hot-path methods of the reference's classes and undo() puts them back (no GPU needed)."""
    from oracle import ref_loader

    if not ref_loader.reference_available():
        pytest.skip("reference package not present")
    pkg = ref_loader.load_reference()
    from gpy_dla_detection import dla_gp as rd, null_gp as rn, subdla_gp as rs, voigt as rv

    from gpy_dla_detection_b200 import interop, null_gp, voigt

    before = {(c, n): c.__dict__.get(n) for c in (rn.NullGP, rd.DLAGP, rs.SubDLAGP)
              for n in ("set_data", "log_model_evidence", "log_model_evidences", "this_dla_gp", "log_mvnpdf_low_rank")}
    ref_voigt = rv.voigt_absorption
    undo = interop.patch_reference(pkg)
    try:
        assert rn.NullGP.set_data is null_gp.NullGP.set_data
        assert isinstance(rn.NullGP.__dict__["log_mvnpdf_low_rank"], staticmethod)
        assert rd.DLAGP.__dict__["log_model_evidences"] is not before[(rd.DLAGP, "log_model_evidences")]
        assert rv.voigt_absorption is voigt.voigt_absorption and rd.voigt_absorption is voigt.voigt_absorption
        assert issubclass(rd.DLAGP, rn.NullGP)  # the reference's hierarchy is untouched: its isinstance checks hold
    finally:
        undo()
    for (c, n), old in before.items():
        assert c.__dict__.get(n) is old, (c, n)
    assert rv.voigt_absorption is ref_voigt and rd.voigt_absorption is ref_voigt


def test_bench_reference_arm_line_and_config_identity():
    """
    `bench.py --impl reference` (the CPU arm the driver divides by): one JSON line with the contract's keys, the
    live reference as implementation when it is staged, and a `config` object identical to the GPU arm's.
    """
    import json

    env = dict(os.environ, OMP_NUM_THREADS="1")
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--ref-fraction", "200"], capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-3000:]
    line = json.loads(res.stdout.strip().splitlines()[-1])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["unit"] == "spectra/s" and line["value"] > 0
    assert line["e2e"] == {"value": line["value"], "unit": "spectra/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = line["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and 1 <= cb["cores"] <= 16 and cb["sample_fraction"] == 1.0 / 200
    from oracle import ref_loader

    assert cb["kind"] == ("reference" if ref_loader.reference_available() else "port")
    sys.path.insert(0, ROOT)
    import bench

    assert line["config"] == bench.workload_config(1, bench.CONFIGS[1]["spectra"])  # what the GPU arm prints
    assert line["metric"] == bench.CONFIGS[1]["metric"]
