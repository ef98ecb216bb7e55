"""
preload.py : ragged on-disk store of a spectrum list.

Reference: run_bayes_select.process_qso reads one FITS file per loop iteration through
read_spec.read_spec (read_spec.py:22-71, run_bayes_select.py:141-146); the MATLAB pipeline this
was translated from reads every file ONCE into `preloaded_qsos.mat` (preload_qsos.m:56-67:
cell arrays all_wavelengths / all_flux / all_noise_variance / all_pixel_mask) and all later
stages start from that file.  At > 1 000 spectra/s per GPU the per-spectrum FITS parse is the
bottleneck, so the device engine is fed from the same kind of store:

    <path>/offsets.i64          (Q + 1) int64   pixel offsets of every spectrum
    <path>/wavelengths.f64      flat float64    observed wavelengths
    <path>/flux.f64             flat float64
    <path>/noise_variance.f64   flat float64    NaN where ivar == 0 (read_spec.py:55-58)
    <path>/pixel_mask.u8        flat uint8      1 = bad pixel
    <path>/z_qsos.f64           (Q) float64     optional
    <path>/meta.json            {"num_spectra", "num_pixels", "qso_list", "version"}

Flat raw files: `preload` appends spectrum after spectrum without knowing the total size, and
`PreloadedSpectra` maps them with np.memmap, so a 160 000-spectrum catalogue (18.6 GB) is never
resident in host RAM - `chunk(a, b)` hands the engine zero-copy views of the mapped pages.
"""
import json
import os
from typing import Callable, Iterable, List, Optional, Sequence, Tuple

import numpy as np

VERSION = 1
_FILES = (("wavelengths", "wavelengths.f64", np.float64), ("flux", "flux.f64", np.float64),
          ("noise_variance", "noise_variance.f64", np.float64), ("pixel_mask", "pixel_mask.u8", np.uint8))


def preload(qso_list: Sequence, read_spec: Callable, path: str, z_qso_list: Optional[Sequence[float]] = None,
            progress: Optional[Callable[[int], None]] = None) -> "PreloadedSpectra":
    """
    Read every item of `qso_list` once with `read_spec(item) -> (wavelengths, flux, noise_variance,
    pixel_mask)` and append it to the store at `path` (a directory, created).  Returns the opened store.
    """
    os.makedirs(path, exist_ok=True)
    handles = {name: open(os.path.join(path, fname), "wb") for name, fname, _ in _FILES}
    offsets = [0]
    try:
        for i, item in enumerate(qso_list):
            wl, fl, nv, pm = read_spec(item)
            n = len(wl)
            if not (len(fl) == n and len(nv) == n and len(pm) == n):
                raise ValueError("read_spec returned arrays of different lengths for item %r" % (item,))
            handles["wavelengths"].write(np.ascontiguousarray(wl, dtype=np.float64).tobytes())
            handles["flux"].write(np.ascontiguousarray(fl, dtype=np.float64).tobytes())
            handles["noise_variance"].write(np.ascontiguousarray(nv, dtype=np.float64).tobytes())
            handles["pixel_mask"].write(np.ascontiguousarray(np.asarray(pm).astype(np.uint8)).tobytes())
            offsets.append(offsets[-1] + n)
            if progress is not None:
                progress(i)
    finally:
        for h in handles.values():
            h.close()
    np.asarray(offsets, dtype=np.int64).tofile(os.path.join(path, "offsets.i64"))
    if z_qso_list is not None:
        z = np.asarray(z_qso_list, dtype=np.float64)
        if z.shape != (len(offsets) - 1,):
            raise ValueError("z_qso_list must have one entry per spectrum")
        z.tofile(os.path.join(path, "z_qsos.f64"))
    meta = {"version": VERSION, "num_spectra": len(offsets) - 1, "num_pixels": int(offsets[-1]),
            "qso_list": [str(q) for q in qso_list]}
    tmp = os.path.join(path, "meta.json.tmp")
    with open(tmp, "w") as f:
        json.dump(meta, f)
    os.replace(tmp, os.path.join(path, "meta.json"))  # written last: a store without meta.json is incomplete
    return PreloadedSpectra(path)


class PreloadedSpectra:
    """Memory-mapped view of a store written by `preload`."""

    def __init__(self, path: str):
        meta_path = os.path.join(path, "meta.json")
        if not os.path.exists(meta_path):
            raise FileNotFoundError("%s is not a complete preloaded store (no meta.json)" % path)
        with open(meta_path) as f:
            self.meta = json.load(f)
        if self.meta.get("version") != VERSION:
            raise ValueError("unsupported preloaded-store version %r" % self.meta.get("version"))
        self.path = path
        self.qso_list: List[str] = self.meta["qso_list"]
        self.offsets = np.fromfile(os.path.join(path, "offsets.i64"), dtype=np.int64)
        Q, total = self.meta["num_spectra"], self.meta["num_pixels"]
        if self.offsets.shape != (Q + 1,) or self.offsets[0] != 0 or self.offsets[-1] != total:
            raise ValueError("offsets.i64 does not match meta.json")
        for name, fname, dtype in _FILES:
            full = os.path.join(path, fname)
            if os.path.getsize(full) != total * np.dtype(dtype).itemsize:
                raise ValueError("%s has the wrong size for %d pixels" % (fname, total))
            setattr(self, name, np.memmap(full, dtype=dtype, mode="r", shape=(total,)) if total else np.zeros(0, dtype))
        zp = os.path.join(path, "z_qsos.f64")
        self.z_qsos = np.fromfile(zp, dtype=np.float64) if os.path.exists(zp) else None

    def __len__(self) -> int:
        return self.meta["num_spectra"]

    def chunk(self, start: int, stop: int) -> Tuple[np.ndarray, np.ndarray, np.ndarray, np.ndarray, np.ndarray]:
        """(offsets rebased to 0, wavelengths, flux, noise_variance, pixel_mask) of spectra [start, stop): views."""
        a, b = int(self.offsets[start]), int(self.offsets[stop])
        return (self.offsets[start:stop + 1] - a, self.wavelengths[a:b], self.flux[a:b], self.noise_variance[a:b],
                self.pixel_mask[a:b])

    def spectrum(self, i: int) -> Tuple[np.ndarray, np.ndarray, np.ndarray, np.ndarray]:
        """The return value of read_spec for spectrum i (pixel_mask as bool)."""
        a, b = int(self.offsets[i]), int(self.offsets[i + 1])
        return (np.array(self.wavelengths[a:b]), np.array(self.flux[a:b]), np.array(self.noise_variance[a:b]),
                np.array(self.pixel_mask[a:b]).astype(bool))

    def __iter__(self) -> Iterable:
        return (self.spectrum(i) for i in range(len(self)))

    def view(self, start: int, stop: int) -> "StoreView":
        """Spectra [start, stop) as a store of their own (what one rank of a sharded run sees)."""
        return StoreView(self, start, stop)


class StoreView:
    """A contiguous range of a PreloadedSpectra with the same read interface."""

    def __init__(self, store: PreloadedSpectra, start: int, stop: int):
        if not (0 <= start <= stop <= len(store)):
            raise IndexError("view [%d, %d) outside a store of %d spectra" % (start, stop, len(store)))
        self.store, self.start, self.stop = store, start, stop
        self.qso_list = store.qso_list[start:stop]
        self.z_qsos = None if store.z_qsos is None else store.z_qsos[start:stop]

    def __len__(self) -> int:
        return self.stop - self.start

    def chunk(self, start: int, stop: int):
        return self.store.chunk(self.start + start, self.start + stop)

    def spectrum(self, i: int):
        return self.store.spectrum(self.start + i)
