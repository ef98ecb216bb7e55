"""
CPU tests of the catalogue plumbing either side of the device engine:
the preloaded ragged store, the chunked writer with resume, the merge to the reference's output
file, and the chunk-dealing multi-process mode over gloo.  The device engine is replaced by a
deterministic stand-in (`FakeProcessor`) - the GPU counterpart is tests/test_gpu_catalogue_io.py.
"""
import json
import os
import subprocess
import sys
import types

import numpy as np
import pytest

from gpy_dla_detection_b200 import catalogue_io, preload, read_spec, synthetic
from gpy_dla_detection_b200.run_bayes_select import process_qso
from gpy_dla_detection_b200.set_parameters import Parameters

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _spectra(Q, seed=0):
    rng = np.random.default_rng(seed)
    out = []
    for q in range(Q):
        n = int(rng.integers(5, 40))
        wl = np.sort(rng.uniform(3600, 10000, n))
        fl = rng.normal(size=n)
        nv = rng.uniform(0.1, 1.0, n)
        pm = rng.random(n) < 0.1
        nv[pm] = np.nan
        out.append((wl, fl, nv, pm))
    return out


class FakeProcessor:
    """Stands in for CatalogueProcessor.process: results are a deterministic function of the chunk's inputs."""

    def __init__(self, max_dlas=4, S=6, fail_after=None):
        self.max_dlas, self.S, self.calls, self.fail_after = max_dlas, S, 0, fail_after

    def process(self, offsets, wl, fl, nv, pm, z_qsos, keep_samples=False):
        if self.fail_after is not None and self.calls >= self.fail_after:
            raise RuntimeError("simulated crash")
        self.calls += 1
        Q, md, S = len(z_qsos), self.max_dlas, self.S
        tot = np.array([np.nansum(fl[offsets[q]:offsets[q + 1]]) + pm[offsets[q]:offsets[q + 1]].sum() for q in range(Q)])
        z = np.asarray(z_qsos)
        m = 2 + md
        out = dict(
            min_z_dlas=z - 1.0, max_z_dlas=z - 0.1,
            log_priors=np.outer(tot, np.arange(m)), log_likelihoods=np.outer(z, np.arange(m)) - 3.0,
            log_posteriors=np.outer(tot + z, np.ones(m)), model_posteriors=np.full((Q, m), 1.0 / m),
            p_dlas=np.tanh(tot), p_no_dlas=1 - np.tanh(tot),
            MAP_z_dlas=np.tile(z[:, None, None], (1, md, md)), MAP_log_nhis=np.tile(tot[:, None, None], (1, md, md)),
            num_pixels=np.diff(offsets).astype(np.int32), status=np.zeros(Q, dtype=np.int32),
        )
        for name in ("log_priors", "log_likelihoods", "log_posteriors"):
            out[name + "_no_dla"], out[name + "_lls"], out[name + "_dla"] = out[name][:, 0], out[name][:, 1], out[name][:, -md:]
        if keep_samples:
            out["sample_log_likelihoods_dla"] = np.tile(tot[:, None, None], (1, S, md)) + np.arange(S)[None, :, None]
            out["sample_log_likelihoods_lls"] = np.tile(z[:, None], (1, S))
            out["base_sample_inds"] = np.tile(np.arange(S, dtype=np.int32)[None, :, None], (Q, 1, md - 1))
        return out


def test_preload_roundtrip_and_views(tmp_path):
    spectra = _spectra(11)
    names = ["spec-%d.fits" % i for i in range(11)]
    z = 2.0 + 0.1 * np.arange(11)
    lookup = dict(zip(names, spectra))
    store = preload.preload(names, lookup.__getitem__, str(tmp_path / "store"), z_qso_list=z)
    assert len(store) == 11 and store.qso_list == names and np.array_equal(store.z_qsos, z)
    for i in (0, 5, 10):
        for a, b in zip(store.spectrum(i), spectra[i]):
            assert np.array_equal(a, b, equal_nan=True)
        assert store.spectrum(i)[3].dtype == bool
    offsets, wl, fl, nv, pm = store.chunk(3, 8)
    assert offsets[0] == 0 and len(offsets) == 6 and offsets.dtype == np.int64 and pm.dtype == np.uint8
    assert np.array_equal(wl, np.concatenate([s[0] for s in spectra[3:8]]))
    assert np.array_equal(nv, np.concatenate([s[2] for s in spectra[3:8]]), equal_nan=True)
    assert isinstance(wl, np.memmap)  # a view of the mapped file, not a copy
    view = store.view(3, 8)
    assert len(view) == 5 and view.qso_list == names[3:8]
    assert np.array_equal(view.chunk(1, 3)[1], store.chunk(4, 6)[1])
    # reopening validates the files against the metadata
    again = preload.PreloadedSpectra(str(tmp_path / "store"))
    assert np.array_equal(again.offsets, store.offsets)
    os.remove(str(tmp_path / "store" / "meta.json"))
    with pytest.raises(FileNotFoundError):
        preload.PreloadedSpectra(str(tmp_path / "store"))


def test_preload_rejects_ragged_reader(tmp_path):
    with pytest.raises(ValueError):
        preload.preload(["a"], lambda _: (np.arange(3.0), np.arange(4.0), np.arange(3.0), np.zeros(3, bool)),
                        str(tmp_path / "bad"))


def test_read_spec_column_arithmetic():
    """read_spec.py:49-69: 10**loglam, NaN variance where ivar == 0, BRIGHTSKY bit 24 of and_mask"""
    loglam = np.array([3.6, 3.6001, 3.6002, 3.6003])
    ivar = np.array([4.0, 0.0, 0.25, 1.0])
    and_mask = np.array([0, 0, 1 << 24, 1 << 3])
    wl, fl, nv, pm = read_spec.arrays_from_boss_columns(loglam, np.arange(4.0), ivar, and_mask)
    assert np.array_equal(wl, 10**loglam)
    assert np.array_equal(nv, np.array([0.25, np.nan, 4.0, 1.0]), equal_nan=True)
    assert pm.tolist() == [False, True, True, False]


def _run(tmp_path, Q, out_name, **kw):
    spectra = _spectra(Q, seed=3)
    names = list(range(Q))
    z = 2.0 + 0.01 * np.arange(Q)
    return process_qso(names, z, lambda i: spectra[i], 4, True, params=Parameters(num_dla_samples=6),
                       out_dir=str(tmp_path / out_name), chunk_spectra=4, **kw)


def test_chunked_run_kill_and_resume_is_byte_identical(tmp_path):
    Q = 18  # 5 chunks of 4, 4, 4, 4, 2
    one_shot = _run(tmp_path, Q, "a", processor=FakeProcessor(), keep_samples=True)
    with pytest.raises(RuntimeError, match="simulated crash"):
        _run(tmp_path, Q, "b", processor=FakeProcessor(fail_after=3), keep_samples=True)
    man = json.load(open(tmp_path / "b" / "manifest.json"))
    assert sorted(man["chunks"]) == ["0", "1", "2"]
    survivor = FakeProcessor()
    resumed = _run(tmp_path, Q, "b", processor=survivor, keep_samples=True)
    assert survivor.calls == 2  # only the two missing chunks were recomputed
    a = catalogue_io.load_catalogue(one_shot["output_file"])
    b = catalogue_io.load_catalogue(resumed["output_file"])
    assert sorted(a) == sorted(b)
    for k in a:
        assert a[k].dtype == b[k].dtype and a[k].tobytes() == b[k].tobytes(), k
    # the merged file has the reference's dataset names, shapes and scalars (run_bayes_select.py:248-295)
    for name in ("prior_z_qso_increase", "k", "normalization_min_lambda", "normalization_max_lambda", "min_z_cut",
                 "max_z_cut", "num_dla_samples", "num_lines", "num_forest_lines", "min_z_dlas", "max_z_dlas",
                 "log_priors_no_dla", "log_priors_lls", "log_priors_dla", "log_likelihoods_no_dla", "log_likelihoods_lls",
                 "log_likelihoods_dla", "log_posteriors_no_dla", "log_posteriors_lls", "log_posteriors_dla", "MAP_z_dlas",
                 "MAP_log_nhis", "p_dlas", "p_no_dlas", "model_posteriors", "z_qsos", "qso_list"):
        assert name in a, name
    assert a["num_dla_samples"] == 6 and a["k"] == 20 and a["log_priors_dla"].shape == (Q, 4)
    assert a["model_posteriors"].shape == (Q, 6) and a["MAP_z_dlas"].shape == (Q, 4, 4) and a["qso_list"].shape == (Q,)
    # per-sample arrays: merged chunk by chunk into memory-mapped .npy files next to the .npz
    base = one_shot["output_file"][:-4]
    sl = np.load(base + ".sample_log_likelihoods_dla.npy", mmap_mode="r")
    bi = np.load(base + ".base_sample_inds.npy", mmap_mode="r")
    assert sl.shape == (Q, 6, 4) and bi.shape == (Q, 6, 3) and bi.dtype == np.int32
    assert np.array_equal(np.load(str(tmp_path / "b" / "processed_qsos_multi_meanflux.sample_log_likelihoods_dla.npy")), sl)
    # the returned dictionary carries the per-quasar arrays only
    assert "sample_log_likelihoods_dla" not in resumed and resumed["p_dlas"].shape == (Q,)


def test_resume_refuses_a_different_run(tmp_path):
    _run(tmp_path, 9, "c", processor=FakeProcessor())
    with pytest.raises(ValueError, match="cannot resume"):
        _run(tmp_path, 10, "c", processor=FakeProcessor())
    with pytest.raises(ValueError, match="cannot resume"):
        _run(tmp_path, 9, "c", processor=FakeProcessor(), keep_samples=True)
    # resume=False starts over
    p = FakeProcessor()
    _run(tmp_path, 10, "c", processor=p, resume=False)
    assert p.calls == 3


def test_resume_recomputes_a_truncated_chunk(tmp_path):
    _run(tmp_path, 9, "d", processor=FakeProcessor())
    with open(tmp_path / "d" / "chunk_000001.npz", "r+b") as f:
        f.truncate(100)
    p = FakeProcessor()
    out = _run(tmp_path, 9, "d", processor=p)
    assert p.calls == 1 and out["p_dlas"].shape == (9,)


def test_in_memory_run_matches_chunked_run(tmp_path):
    Q = 10
    spectra = _spectra(Q, seed=3)
    z = 2.0 + 0.01 * np.arange(Q)
    mem = process_qso(list(range(Q)), z, lambda i: spectra[i], 4, True, params=Parameters(num_dla_samples=6),
                      processor=FakeProcessor(), chunk_spectra=3, keep_samples=True)
    disk = _run(tmp_path, Q, "e", processor=FakeProcessor())
    for k in catalogue_io.PER_QUASAR:
        assert np.array_equal(mem[k], disk[k]), k
    assert mem["sample_log_likelihoods_dla"].shape == (Q, 6, 4) and np.array_equal(mem["z_qsos"], z)
    # preloaded store instead of a reader: same results
    store = preload.preload(list(range(Q)), lambda i: spectra[i], str(tmp_path / "store"), z_qso_list=z)
    pre = process_qso(store.qso_list, store.z_qsos, None, 4, True, params=Parameters(num_dla_samples=6),
                      processor=FakeProcessor(), chunk_spectra=3, preloaded=store)
    for k in catalogue_io.PER_QUASAR:
        assert np.array_equal(pre[k], disk[k]), k


def test_merge_writes_hdf5_through_h5py_when_available(tmp_path, monkeypatch):
    """h5py is absent from the image: a recording stand-in checks the calls the HDF5 branch makes."""
    created = {}

    class FakeDataset:
        def __init__(self, shape, dtype):
            self.arr = np.zeros(shape, dtype=dtype)

        def __setitem__(self, key, val):
            self.arr[key] = val

    class FakeFile:
        def __init__(self, name, mode):
            created["name"], created["mode"] = name, mode

        def __enter__(self):
            return self

        def __exit__(self, *a):
            return False

        def create_dataset(self, name, data=None, shape=None, dtype=None):
            if data is not None:
                created[name] = np.asarray(data)
                return None
            ds = FakeDataset(shape, dtype)
            created[name] = ds.arr
            return ds

    fake = types.ModuleType("h5py")
    fake.File = FakeFile
    fake.string_dtype = lambda encoding="utf-8": object
    monkeypatch.setitem(sys.modules, "h5py", fake)
    out = _run(tmp_path, 9, "f", processor=FakeProcessor(), keep_samples=True)
    assert out["output_file"].endswith("processed_qsos_multi_meanflux.h5") and created["mode"] == "w"
    assert created["p_dlas"].shape == (9,) and created["qso_list"].shape == (9,)
    assert created["sample_log_likelihoods_dla"].shape == (9, 6, 4) and created["base_sample_inds"].dtype == np.int32
    ref = _run(tmp_path, 9, "g", processor=FakeProcessor(), keep_samples=True)
    assert np.array_equal(created["log_posteriors_dla"], ref["log_posteriors_dla"])


_WORKER = r"""
import os, sys
sys.path.insert(0, {root!r})
sys.path.insert(0, os.path.join({root!r}, "tests"))
import numpy as np
import torch.distributed as dist
from gpy_dla_detection_b200.run_bayes_select import process_qso_sharded
from gpy_dla_detection_b200.set_parameters import Parameters
from gpy_dla_detection_b200 import catalogue_io
from test_catalogue_io import FakeProcessor, _spectra

dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
Q = 18
spectra = _spectra(Q, seed=3)
z = 2.0 + 0.01 * np.arange(Q)
proc = FakeProcessor()
out = process_qso_sharded(list(range(Q)), z, lambda i: spectra[i], 4, True, params=Parameters(num_dla_samples=6),
                          out_dir={out!r}, chunk_spectra=4, processor=proc, keep_samples=True)
print("RANK", rank, "CHUNKS", proc.calls)
if rank == 0:
    single = catalogue_io.load_catalogue({single!r})
    merged = catalogue_io.load_catalogue(out["output_file"])
    assert sorted(single) == sorted(merged)
    for k in single:
        assert single[k].tobytes() == merged[k].tobytes(), k
    print("SHARDED_OK")
else:
    assert out is None
dist.barrier()
dist.destroy_process_group()
"""


def test_sharded_chunk_dealing_world_size_2_gloo(tmp_path):
    """two ranks over gloo deal the chunks round-robin, rank 0 merges: identical to the single-process file"""
    single = _run(tmp_path, 18, "single", processor=FakeProcessor(), keep_samples=True)
    script = tmp_path / "worker.py"
    script.write_text(_WORKER.format(root=ROOT, out=str(tmp_path / "sharded"), single=single["output_file"]))
    port = 31500 + (os.getpid() % 2000)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
           "127.0.0.1", "--master-port", str(port), str(script)]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=240, env=dict(os.environ, OMP_NUM_THREADS="1"))
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
    assert "SHARDED_OK" in res.stdout
    assert "RANK 0 CHUNKS 3" in res.stdout and "RANK 1 CHUNKS 2" in res.stdout
