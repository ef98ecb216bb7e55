"""
catalogue_io.py : chunked catalogue writer with resume, and the merge to one file.

Reference: run_bayes_select.process_qso keeps every result array of the whole catalogue in RAM and
writes `processed_qsos_multi_meanflux.h5` once at the very end (run_bayes_select.py:248-295); long
runs are split by hand into SLURM array jobs and re-united by CDDF_analysis/sbatch_reunion.py:13-63
(concatenate every per-quasar dataset along the quasar axis, copy the scalars from the first piece).
Here the catalogue streams through the device engine in chunks of C spectra; every finished chunk is
one `chunk_%06d.npz` plus a line in `manifest.json`, so a run that dies at spectrum 150 000 restarts
at the first missing chunk, and `merge()` produces the reference's file:

  * dataset names and shapes exactly as run_bayes_select.py:250-295 (parameter scalars
    prior_z_qso_increase, k, normalization_min/max_lambda, min/max_z_cut, num_dla_samples, num_lines,
    num_forest_lines; per-quasar arrays; z_qsos; qso_list) - readable by CDDF_analysis/qso_loader.py;
  * HDF5 when `h5py` imports, else `<name>.npz` with the same keys (h5py is absent from the build image);
  * the per-sample arrays (sample_log_likelihoods_dla (Q,S,max), base_sample_inds (Q,S,max-1),
    51 GB at Q = 160k) are opt-in, live only in their chunk files, and are merged chunk by chunk
    into a resizable HDF5 dataset / a memory-mapped .npy - never one array in RAM.
"""
import hashlib
import json
import os
import time
from typing import Dict, List, Optional, Sequence

import numpy as np

MANIFEST = "manifest.json"
MANIFEST_VERSION = 1

# per-quasar datasets of run_bayes_select.py:262-288 (+ two diagnostics of this engine)
PER_QUASAR = (
    "min_z_dlas", "max_z_dlas",
    "log_priors_no_dla", "log_priors_lls", "log_priors_dla",
    "log_likelihoods_no_dla", "log_likelihoods_lls", "log_likelihoods_dla",
    "log_posteriors_no_dla", "log_posteriors_lls", "log_posteriors_dla",
    "MAP_z_dlas", "MAP_log_nhis", "p_dlas", "p_no_dlas", "model_posteriors",
    "num_pixels", "status",
)
PER_SAMPLE = ("sample_log_likelihoods_dla", "base_sample_inds", "sample_log_likelihoods_lls")
# scalars of run_bayes_select.py:250-262
PARAM_SCALARS = ("prior_z_qso_increase", "k", "normalization_min_lambda", "normalization_max_lambda", "min_z_cut",
                 "max_z_cut", "num_dla_samples", "num_lines", "num_forest_lines")


def _fingerprint(qso_list: Sequence, z_qsos: np.ndarray) -> str:
    h = hashlib.sha256()
    for q in qso_list:
        h.update(str(q).encode("utf-8"))
        h.update(b"\0")
    h.update(np.ascontiguousarray(z_qsos, dtype=np.float64).tobytes())
    return h.hexdigest()


def _atomic_json(path: str, obj) -> None:
    tmp = path + ".tmp"
    with open(tmp, "w") as f:
        json.dump(obj, f, indent=1, sort_keys=True)
        f.flush()
        os.fsync(f.fileno())
    os.replace(tmp, path)


class ChunkedCatalogueWriter:
    """
    One directory per catalogue run.  `chunks()` lists the (index, start, stop) still to do,
    `write_chunk` stores a finished chunk atomically, `merge` builds the reference's output file.
    """

    def __init__(self, out_dir: str, qso_list: Sequence, z_qso_list: Sequence[float], params, max_dlas: int,
                 chunk_spectra: int = 4096, keep_samples: bool = False, resume: bool = True):
        self.out_dir = out_dir
        self.qso_list = [str(q) for q in qso_list]
        self.z_qsos = np.asarray(z_qso_list, dtype=np.float64)
        if self.z_qsos.shape != (len(self.qso_list),):
            raise ValueError("qso_list and z_qso_list must have the same length")
        self.num_quasars = len(self.qso_list)
        self.chunk_spectra = int(chunk_spectra)
        if self.chunk_spectra < 1:
            raise ValueError("chunk_spectra must be positive")
        self.max_dlas = int(max_dlas)
        self.keep_samples = bool(keep_samples)
        self.param_scalars = {name: (getattr(params, name).item() if hasattr(getattr(params, name), "item")
                                     else getattr(params, name)) for name in PARAM_SCALARS}
        os.makedirs(out_dir, exist_ok=True)
        header = {
            "version": MANIFEST_VERSION,
            "num_quasars": self.num_quasars,
            "chunk_spectra": self.chunk_spectra,
            "max_dlas": self.max_dlas,
            "keep_samples": self.keep_samples,
            "params": self.param_scalars,
            "inputs_sha256": _fingerprint(self.qso_list, self.z_qsos),
        }
        path = os.path.join(out_dir, MANIFEST)
        if os.path.exists(path) and resume:
            with open(path) as f:
                self.manifest = json.load(f)
            for key, val in header.items():
                if self.manifest.get(key) != val:
                    raise ValueError("cannot resume in %s: manifest %s = %r, this run has %r"
                                     % (out_dir, key, self.manifest.get(key), val))
            # a chunk listed in the manifest must still be on disk with the recorded size
            for idx, rec in list(self.manifest["chunks"].items()):
                full = os.path.join(out_dir, rec["file"])
                if not os.path.exists(full) or os.path.getsize(full) != rec["bytes"]:
                    del self.manifest["chunks"][idx]
        else:
            self.manifest = dict(header, chunks={})
            _atomic_json(path, self.manifest)

    # -- plan ---------------------------------------------------------------------------------------
    @property
    def num_chunks(self) -> int:
        return (self.num_quasars + self.chunk_spectra - 1) // self.chunk_spectra

    def chunk_range(self, idx: int):
        start = idx * self.chunk_spectra
        return start, min(start + self.chunk_spectra, self.num_quasars)

    def is_done(self, idx: int) -> bool:
        return str(idx) in self.manifest["chunks"]

    def chunks(self, rank: int = 0, world_size: int = 1):
        """(index, start, stop) of the unfinished chunks this rank owns (round-robin over ranks)."""
        return [(i,) + self.chunk_range(i) for i in range(self.num_chunks)
                if i % world_size == rank and not self.is_done(i)]

    # -- write --------------------------------------------------------------------------------------
    def write_chunk(self, idx: int, results: Dict[str, np.ndarray], update_manifest: bool = True) -> str:
        start, stop = self.chunk_range(idx)
        arrays = {}
        for name in PER_QUASAR + (PER_SAMPLE if self.keep_samples else ()):
            if name not in results:
                if name in PER_SAMPLE and (name != "base_sample_inds" or self.max_dlas > 1):
                    raise KeyError("keep_samples run without %s in the chunk results" % name)
                continue
            a = np.asarray(results[name])
            if a.shape[0] != stop - start:
                raise ValueError("chunk %d: %s has %d rows for %d spectra" % (idx, name, a.shape[0], stop - start))
            arrays[name] = a
        fname = "chunk_%06d.npz" % idx
        full = os.path.join(self.out_dir, fname)
        tmp = full + ".tmp"
        with open(tmp, "wb") as f:
            np.savez(f, start=start, stop=stop, **arrays)
            f.flush()
            os.fsync(f.fileno())
        os.replace(tmp, full)
        rec = {"file": fname, "start": start, "stop": stop, "bytes": os.path.getsize(full), "time": time.time()}
        if update_manifest:
            self.manifest["chunks"][str(idx)] = rec
            _atomic_json(os.path.join(self.out_dir, MANIFEST), self.manifest)
        return full

    def adopt_chunks_on_disk(self) -> None:
        """
        Multi-process runs: every rank writes its own chunk files (update_manifest=False) and rank 0
        records them here after a barrier, so the manifest has a single writer.
        """
        for idx in range(self.num_chunks):
            if self.is_done(idx):
                continue
            fname = "chunk_%06d.npz" % idx
            full = os.path.join(self.out_dir, fname)
            if os.path.exists(full):
                start, stop = self.chunk_range(idx)
                self.manifest["chunks"][str(idx)] = {"file": fname, "start": start, "stop": stop,
                                                     "bytes": os.path.getsize(full), "time": os.path.getmtime(full)}
        _atomic_json(os.path.join(self.out_dir, MANIFEST), self.manifest)

    # -- merge (CDDF_analysis/sbatch_reunion.py:13-63) -------------------------------------------------
    def complete(self) -> bool:
        return all(self.is_done(i) for i in range(self.num_chunks))

    def load_merged(self) -> Dict[str, np.ndarray]:
        """The per-quasar arrays of the whole catalogue in RAM (small: ~0.5 KB per quasar); no per-sample arrays."""
        if not self.complete():
            missing = [i for i in range(self.num_chunks) if not self.is_done(i)]
            raise RuntimeError("catalogue incomplete: %d chunk(s) missing, first %d" % (len(missing), missing[0]))
        parts: Dict[str, List[np.ndarray]] = {}
        for idx in range(self.num_chunks):
            with np.load(os.path.join(self.out_dir, self.manifest["chunks"][str(idx)]["file"])) as z:
                for name in PER_QUASAR:
                    if name in z.files:
                        parts.setdefault(name, []).append(z[name])
        out = {name: np.concatenate(v, axis=0) for name, v in parts.items()}
        for name, val in self.param_scalars.items():
            out[name] = np.asarray(val)
        out["z_qsos"] = self.z_qsos.copy()
        return out

    def merge(self, filename: Optional[str] = None, force_npz: bool = False) -> str:
        """
        Write the reference's output file.  Returns the path written: `<filename>` (HDF5) when h5py is
        importable, else `<filename minus .h5>.npz` (+ one memory-mapped `.npy` per per-sample array).
        """
        filename = filename or os.path.join(self.out_dir, "processed_qsos_multi_meanflux.h5")
        merged = self.load_merged()
        h5py = None
        if not force_npz:
            try:
                import h5py  # noqa: F811
            except ImportError:
                h5py = None
        sample_names = [n for n in PER_SAMPLE if self.keep_samples and (n != "base_sample_inds" or self.max_dlas > 1)]

        def chunk_arrays(name):
            for idx in range(self.num_chunks):
                rec = self.manifest["chunks"][str(idx)]
                with np.load(os.path.join(self.out_dir, rec["file"])) as z:
                    yield rec["start"], rec["stop"], z[name]

        if h5py is not None:
            with h5py.File(filename, "w") as f:
                for name, arr in merged.items():
                    f.create_dataset(name, data=arr)
                f.create_dataset("qso_list", data=np.array(self.qso_list, h5py.string_dtype(encoding="utf-8")))
                for name in sample_names:
                    dset = None
                    for start, stop, arr in chunk_arrays(name):
                        if dset is None:
                            dset = f.create_dataset(name, shape=(self.num_quasars,) + arr.shape[1:], dtype=arr.dtype)
                        dset[start:stop] = arr
            return filename
        base = filename[:-3] if filename.endswith(".h5") else filename
        out_path = base + ".npz"
        merged["qso_list"] = np.array(self.qso_list, dtype=np.str_)
        tmp = out_path + ".tmp"
        with open(tmp, "wb") as f:
            np.savez(f, **merged)
        os.replace(tmp, out_path)
        for name in sample_names:
            mm = None
            for start, stop, arr in chunk_arrays(name):
                if mm is None:
                    mm = np.lib.format.open_memmap(base + "." + name + ".npy", mode="w+", dtype=arr.dtype,
                                                   shape=(self.num_quasars,) + arr.shape[1:])
                mm[start:stop] = arr
            if mm is not None:
                mm.flush()
                del mm
        return out_path


def load_catalogue(path: str) -> Dict[str, np.ndarray]:
    """Read a file written by `merge` (either flavour) into a dictionary of arrays."""
    if path.endswith(".npz"):
        with np.load(path, allow_pickle=False) as z:
            return {k: z[k] for k in z.files}
    import h5py

    with h5py.File(path, "r") as f:
        return {k: f[k][()] for k in f.keys()}
