"""
GPU tests of the catalogue plumbing around the device engine (e): chunked run + merge == one-shot
process, preloaded store == reader, resume, and the sharded run (one process per GPU, NCCL only for the barriers /
gather) == the single-GPU run bit for bit.  The 2-GPU test skips itself when fewer than two GPUs are visible.
"""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from tests import helpers as H

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PER_SAMPLE = ("sample_log_likelihoods_dla", "sample_log_likelihoods_lls", "base_sample_inds")


def _workload(Q, S=128):
    from gpy_dla_detection_b200 import synthetic

    st = H.Setup(S)
    z = synthetic.sample_z_qsos(Q, seed=99)
    spectra = [synthetic.make_spectrum(st.model, float(z[i]), seed=700 + i) for i in range(Q)]
    spectra[3] = tuple(a[200:4000] for a in spectra[3])  # ragged
    return st, z, spectra


def _kwargs(st):
    d, s = st.sample_objects()
    return dict(params=st.params, prior=st.prior, model=st.model, dla_samples=d, subdla_samples=s, batch_spectra=4)


def test_chunked_run_equals_one_shot_and_resumes(gpu, tmp_path):
    from gpy_dla_detection_b200 import catalogue_io, preload
    from gpy_dla_detection_b200.run_bayes_select import process_qso

    Q = 11
    st, z, spectra = _workload(Q)
    names = ["spec-%02d" % i for i in range(Q)]
    lookup = dict(zip(names, spectra))
    proc = st.catalogue(4, True, batch_spectra=4)
    one = proc.process(*proc.pack(spectra), z, keep_samples=True)

    out = process_qso(names, z, lookup.__getitem__, 4, True, keep_samples=True, chunk_spectra=3,
                      out_dir=str(tmp_path / "run"), **_kwargs(st))
    merged = catalogue_io.load_catalogue(out["output_file"])
    for k in catalogue_io.PER_QUASAR:
        assert np.array_equal(merged[k], one[k], equal_nan=True), k
    base = out["output_file"][:-4]
    for k in PER_SAMPLE:
        assert np.array_equal(np.load(base + "." + k + ".npy"), one[k], equal_nan=True), k
    assert list(merged["qso_list"]) == names and np.array_equal(merged["z_qsos"], z)
    assert int(merged["num_dla_samples"]) == 128 and int(merged["num_lines"]) == 3

    # lose a chunk, resume: only that chunk is recomputed and the merged file is byte-identical
    before = {k: merged[k].tobytes() for k in merged}
    os.remove(str(tmp_path / "run" / "chunk_000002.npz"))
    lib = gpu.load_library()
    launches0 = lib.dla_kernel_launch_count()
    again = process_qso(names, z, lookup.__getitem__, 4, True, keep_samples=True, chunk_spectra=3,
                        out_dir=str(tmp_path / "run"), **_kwargs(st))
    one_chunk_launches = lib.dla_kernel_launch_count() - launches0
    assert 0 < one_chunk_launches < 60  # one chunk of 3 spectra = one batch
    merged2 = catalogue_io.load_catalogue(again["output_file"])
    assert {k: merged2[k].tobytes() for k in merged2} == before

    # the memory-mapped preloaded store feeds the engine the same bytes
    store = preload.preload(names, lookup.__getitem__, str(tmp_path / "store"), z_qso_list=z)
    pre = process_qso(store.qso_list, store.z_qsos, None, 4, True, keep_samples=True, chunk_spectra=4, preloaded=store,
                      **_kwargs(st))
    for k in catalogue_io.PER_QUASAR + PER_SAMPLE:
        assert np.array_equal(pre[k], one[k], equal_nan=True), k


def test_pipeline_many_batches_equals_single_batch(gpu):
    """the two-deep batch pipeline (batch = 2: nine batches in flight one after the other) vs one resident batch"""
    Q = 17
    st, z, spectra = _workload(Q, S=96)
    big = st.catalogue(4, True, batch_spectra=32)
    small = st.catalogue(4, True, batch_spectra=2)
    a = big.process(*big.pack(spectra), z, keep_samples=True)
    for _ in range(2):  # the second pass reuses the slots' staging buffers
        b = small.process(*small.pack(spectra), z, keep_samples=True)
        for k in a:
            assert np.array_equal(a[k], b[k], equal_nan=True), k
    small.stage(*small.pack(spectra), z)
    c = small.run_staged(keep_samples=False)
    for k in c:
        assert np.array_equal(a[k], c[k], equal_nan=True), k
    # evaluation accounting: every sample of levels >= 1 is either evaluated or left out by the separation mask, and the
    # mask count is exactly the number of NaNs the reference would have written over computed values
    tm = small.last_timing()
    S, md = 96, 4
    assert np.all(a["status"] == 0)
    assert tm["evaluations"] + tm["evaluations_masked"] == Q * (5 * S + 1)
    assert tm["evaluations_masked"] == int(np.isnan(a["sample_log_likelihoods_dla"][:, :, 1:]).sum())
    assert tm["evaluations_masked"] > 0 and tm["likelihood_flops"] > 0


_WORKER = r"""
import os, sys
sys.path.insert(0, {root!r})
import numpy as np
import torch
import torch.distributed as dist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
from gpy_dla_detection_b200 import _lib, catalogue_io
from gpy_dla_detection_b200.run_bayes_select import process_qso_sharded
from tests.test_gpu_catalogue_io import _workload, _kwargs
_lib.init(local)
Q = 13
st, z, spectra = _workload(Q)
names = list(range(Q))
single = np.load({single!r})
# (a) chunks dealt round-robin, every rank writes its own chunk files, rank 0 merges
out = process_qso_sharded(names, z, lambda i: spectra[i], 4, True, out_dir={out!r}, chunk_spectra=2, keep_samples=True,
                          **_kwargs(st))
if rank == 0:
    merged = catalogue_io.load_catalogue(out["output_file"])
    for k in catalogue_io.PER_QUASAR:
        assert np.array_equal(merged[k], single[k], equal_nan=True), k
    base = out["output_file"][:-4]
    assert np.array_equal(np.load(base + ".base_sample_inds.npy"), single["base_sample_inds"])
# (b) block partition + per-array NCCL gather
out = process_qso_sharded(names, z, lambda i: spectra[i], 4, True, keep_samples=True, **_kwargs(st))
if rank == 0:
    for k in catalogue_io.PER_QUASAR + ("sample_log_likelihoods_dla", "base_sample_inds"):
        assert np.array_equal(out[k], single[k], equal_nan=True), k
    print("SHARDED_GPU_OK", world)
else:
    assert out is None
dist.barrier()
dist.destroy_process_group()
"""


def test_sharded_catalogue_on_two_gpus_equals_single_gpu(gpu, tmp_path):
    if gpu.load_library().dla_device_count() < 2:
        pytest.skip("needs two visible GPUs")
    Q = 13
    st, z, spectra = _workload(Q)
    proc = st.catalogue(4, True, batch_spectra=4)
    one = proc.process(*proc.pack(spectra), z, keep_samples=True)
    np.savez(str(tmp_path / "single.npz"), **one)
    script = tmp_path / "worker.py"
    script.write_text(_WORKER.format(root=ROOT, single=str(tmp_path / "single.npz"), out=str(tmp_path / "sharded")))
    port = 33500 + (os.getpid() % 2000)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), str(script)]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
    assert "SHARDED_GPU_OK 2" in res.stdout
