#!/usr/bin/env python
"""
run_config2.py : BASELINE.json configs[2] - a DR12Q-sized catalogue through the REAL sharded catalogue path
(process_qso_sharded: preloaded ragged store -> chunks dealt to the ranks -> device engine -> chunk files -> merge).

  python tools/run_config2.py --spectra 16000                                     (1 GPU)
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29611 \
         tools/run_config2.py --spectra 160000 --unique 2000

The catalogue has `--spectra` entries; to keep host-side synthesis short it cycles over `--unique` distinct synthetic
spectra (spectrum i of the store is synthetic spectrum i mod unique), which changes nothing for the engine: every entry
is read from the store, uploaded and processed.  Rank 0 prints one JSON line: wall seconds of the whole path including
the barriers, the merge and the file write, spectra/s, and - with --check - whether the merged per-quasar arrays equal a
single-process run of the first `--check` spectra bit for bit.
"""
import argparse
import json
import os
import shutil
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--spectra", type=int, default=16000)
    ap.add_argument("--unique", type=int, default=1000)
    ap.add_argument("--chunk", type=int, default=1024)
    ap.add_argument("--batch", type=int, default=128)
    ap.add_argument("--check", type=int, default=0, help="compare the first N merged rows with a single-process run")
    ap.add_argument("--dir", default="/tmp/dla_config2")
    args = ap.parse_args()

    import torch
    import torch.distributed as dist

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import __graft_entry__ as graft

    if rank == 0:
        graft.build()
    if world > 1:
        dist.barrier()
    from gpy_dla_detection_b200 import _lib, catalogue_io, preload, synthetic
    from gpy_dla_detection_b200.dla_samples import DLASamplesArrays
    from gpy_dla_detection_b200.run_bayes_select import CatalogueProcessor, process_qso, process_qso_sharded
    from gpy_dla_detection_b200.subdla_samples import SubDLASamplesArrays

    _lib.init(local)
    Q, U = args.spectra, min(args.unique, args.spectra)
    params, model, prior, dla, sub, z_u, spectra_u = synthetic.make_workload(U, 0)
    dla_s = DLASamplesArrays(params, prior, dla["offset_samples"], dla["log_nhi_samples"], dla["nhi_samples"])
    sub_s = SubDLASamplesArrays(params, prior, sub["offset_samples"], sub["log_nhi_samples"], sub["nhi_samples"],
                                sub["Z_lls"], sub["Z_dla"])
    store_dir, out_dir = os.path.join(args.dir, "store"), os.path.join(args.dir, "out")
    names = ["synthetic-%06d" % i for i in range(Q)]
    z_all = z_u[np.arange(Q) % U]
    if rank == 0:
        shutil.rmtree(args.dir, ignore_errors=True)
        t0 = time.perf_counter()
        preload.preload(names, lambda name: spectra_u[int(name[-6:]) % U], store_dir, z_qso_list=z_all)
        preload_s = time.perf_counter() - t0
    if world > 1:
        dist.barrier()
    store = preload.PreloadedSpectra(store_dir)
    kw = dict(params=params, prior=prior, model=model, dla_samples=dla_s, subdla_samples=sub_s, batch_spectra=args.batch)
    proc = CatalogueProcessor(params, prior, model, dla_s, sub_s, 4, True, batch_spectra=args.batch)
    # warm the engine (allocations, first-launch costs) on a few spectra outside the timed region
    proc.process(*store.chunk(0, min(Q, args.batch)), z_all[: min(Q, args.batch)])
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = process_qso_sharded(names, z_all, None, 4, True, preloaded=store, out_dir=out_dir, chunk_spectra=args.chunk,
                              processor=proc, **kw)
    if world > 1:
        dist.barrier()
    wall = time.perf_counter() - t0
    if rank == 0:
        line = {"workload": "configs[2]: %d-spectrum catalogue (%d distinct synthetic spectra), max_dlas=4, S=10000, "
                            "sharded over %d GPU(s) by process_qso_sharded with the chunked writer" % (Q, U, world),
                "n_gpus": world, "spectra": Q, "chunk_spectra": args.chunk, "batch_spectra": args.batch,
                "wall_s_including_merge": wall, "spectra_per_s": Q / wall, "preload_s": preload_s,
                "store_bytes": int(store.offsets[-1]) * 25, "output_file": out["output_file"],
                "p_dla_gt_0.9": int(np.sum(out["p_dlas"] > 0.9)), "status_ok": int(np.sum(out["status"] == 0))}
        if args.check:
            n = min(args.check, Q)
            ref = process_qso(names[:n], z_all[:n], None, 4, True, preloaded=store.view(0, n), processor=proc, **kw)
            same = all(np.array_equal(out[k][:n], ref[k], equal_nan=True) for k in catalogue_io.PER_QUASAR)
            # rows of repeated spectra must repeat bit for bit as well
            rep = all(np.array_equal(out[k][:U][: Q - U], out[k][U:2 * U][: Q - U], equal_nan=True)
                      for k in catalogue_io.PER_QUASAR) if Q >= 2 * U else None
            line.update(first_rows_equal_single_process=bool(same), repeated_spectra_repeat=rep)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
