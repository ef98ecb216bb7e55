"""
run_bayes_select.py : catalogue driver - Bayesian model selection for many spectra.

Mirrors the reference's run_bayes_select.process_qso (run_bayes_select.py:32-295): for every
spectrum, null / subDLA / DLA(1..max_dlas) evidences, priors, posteriors, MAP absorber
parameters, collected into (num_quasars, ...) arrays with the reference's dataset names.
The per-spectrum Python loop of the reference is replaced by the batched device engine
(`dla_catalogue_*` of the C-ABI): a batch of spectra is resident on the GPU and every
stage is one launch over the batch.  Spectra are independent, so a catalogue is sharded
over GPUs by giving each rank a slice of the spectrum list (see `shard_range`).

RNG parity: the reference re-seeds NumPy's global MT19937 with 0 for every spectrum
(run_bayes_select.py:144) and the DLA model draws (max_dlas - 1) x S uniforms through
np.random.choice, so every spectrum sees the same uniforms; they are generated once here
with RandomState(0) and handed to the device.
"""
import ctypes
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import _lib
from .null_gp import _Handle
from .set_parameters import Parameters


def shard_range(num_items: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous block partition of [0, num_items) over ranks (spectra are independent)."""
    base, rem = divmod(num_items, world_size)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def log_priors_for(prior, z_qsos: np.ndarray, max_dlas: int, Z_lls: float, Z_dla: float) -> np.ndarray:
    """
    (Q, 2 + max_dlas) log model priors [nan, subDLA, DLA 1..max] from the prior catalogue
    (dla_gp.py:398-426, subdla_gp.py:311-346); entry 0 is completed on the device.
    """
    z_qsos = np.asarray(z_qsos, dtype=np.float64)
    Q = z_qsos.shape[0]
    counts = np.empty((Q, 2))
    if hasattr(prior, "z_qsos") and hasattr(prior, "dla_ind") and hasattr(prior, "params"):
        # vectorised less_ind (model_priors.py:142-157): strict < on z_qso + prior_z_qso_increase
        order = np.argsort(prior.z_qsos, kind="stable")
        zs = np.asarray(prior.z_qsos)[order]
        cum = np.concatenate([[0], np.cumsum(np.asarray(prior.dla_ind)[order])])
        pos = np.searchsorted(zs, z_qsos + prior.params.prior_z_qso_increase, side="left")
        counts[:, 0] = cum[pos]
        counts[:, 1] = pos
    else:
        for q in range(Q):
            counts[q] = prior.less_ind(z_qsos[q])
    out = np.full((Q, 2 + max_dlas), np.nan)
    with np.errstate(divide="ignore", invalid="ignore"):
        ratio = counts[:, 0] / counts[:, 1]  # 0 / 0 for a quasar below every catalogue entry -> NaN priors
    p = ratio[:, None] ** np.arange(1, max_dlas + 1)[None, :]
    for i in range(max_dlas - 1):
        p[:, i] = p[:, i] - p[:, i + 1]
    with np.errstate(divide="ignore", invalid="ignore"):
        out[:, 2:] = np.log(p)
        out[:, 1] = np.log(Z_lls / Z_dla * ratio)
    return out


class CatalogueProcessor:
    """
    Device-resident engine for `process_qso`: holds the learned model, the QMC samples and
    the batch workspace on one GPU.
    """

    def __init__(
        self,
        params: Parameters,
        prior,
        model: Dict[str, np.ndarray],
        dla_samples,
        subdla_samples,
        max_dlas: int = 4,
        broadening: bool = True,
        min_z_separation: float = 3000.0,
        batch_spectra: int = 64,
        prev_tau_0: float = 0.0023,
        prev_beta: float = 3.65,
    ):
        self.params = params
        self.prior = prior
        self.max_dlas = int(max_dlas)
        self.dla_samples = dla_samples
        self.subdla_samples = subdla_samples
        self.S = int(params.num_dla_samples)
        lib = _lib.load_library()

        rest = _lib.f64(model["rest_wavelengths"])
        mu, M, lo = _lib.f64(model["mu"]), _lib.f64(model["M"]), _lib.f64(model["log_omega"])
        if params.min_lambda < rest[0] or params.max_lambda > rest[-1]:
            raise ValueError("modelling range exceeds the learned model's rest wavelength grid")
        mptr = ctypes.c_void_p()
        _lib.check(
            lib.dla_model_create(
                _lib.dptr(rest), _lib.dptr(mu), _lib.dptr(M), _lib.dptr(lo), rest.shape[0], M.shape[1],
                float(model["log_c_0"]), float(model["log_tau_0"]), float(model["log_beta"]),
                float(prev_tau_0), float(prev_beta), ctypes.byref(mptr),
            )
        )
        self._model = _Handle(mptr, "dla_model_destroy")

        # same draws as np.random.seed(0) + (max_dlas - 1) calls of np.random.choice per spectrum
        self.uniforms = np.random.RandomState(0).random_sample((max(self.max_dlas - 1, 1), self.S))
        ps = _lib.params_struct(params, broadening, params.kms_to_z(min_z_separation))
        cfg = _lib.CatalogueConfigStruct(self.S, self.max_dlas, int(batch_spectra), 1)
        d_off, d_lognhi, d_nhi = (
            _lib.f64(dla_samples.offset_samples), _lib.f64(dla_samples.log_nhi_samples), _lib.f64(dla_samples.nhi_samples)
        )
        s_off, s_nhi = _lib.f64(subdla_samples.offset_samples), _lib.f64(subdla_samples.nhi_samples)
        for a in (d_off, d_lognhi, d_nhi, s_off, s_nhi):
            assert a.shape == (self.S,)  # dla_gp.py:112-119
        cptr = ctypes.c_void_p()
        _lib.check(
            lib.dla_catalogue_create(
                self._model.ptr, ctypes.byref(ps), ctypes.byref(cfg), _lib.dptr(d_off), _lib.dptr(d_lognhi),
                _lib.dptr(d_nhi), _lib.dptr(s_off), _lib.dptr(s_nhi), _lib.dptr(self.uniforms), ctypes.byref(cptr),
            )
        )
        self._cat = _Handle(cptr, "dla_catalogue_destroy")
        self._staged = None

    # -- input packing -------------------------------------------------------------------------------
    @staticmethod
    def pack(spectra: Sequence[Tuple[np.ndarray, np.ndarray, np.ndarray, np.ndarray]]):
        """List of (wavelengths, flux, noise_variance, pixel_mask) -> ragged contiguous arrays."""
        lengths = np.array([len(s[0]) for s in spectra], dtype=np.int64)
        offsets = np.concatenate([[0], np.cumsum(lengths)]).astype(np.int64)
        wl = np.concatenate([np.asarray(s[0], dtype=np.float64) for s in spectra])
        fl = np.concatenate([np.asarray(s[1], dtype=np.float64) for s in spectra])
        nv = np.concatenate([np.asarray(s[2], dtype=np.float64) for s in spectra])
        pm = np.concatenate([np.asarray(s[3]).astype(np.uint8) for s in spectra])
        return offsets, wl, fl, nv, pm

    def _priors(self, z_qsos: np.ndarray) -> np.ndarray:
        return log_priors_for(self.prior, z_qsos, self.max_dlas, self.subdla_samples._Z_lls, self.subdla_samples._Z_dla)

    def _alloc_outputs(self, Q: int, keep_samples: bool):
        md, S, m = self.max_dlas, self.S, 2 + self.max_dlas
        out = dict(
            min_z_dlas=np.full((Q,), np.nan),
            max_z_dlas=np.full((Q,), np.nan),
            log_priors=np.full((Q, m), np.nan),
            log_likelihoods=np.full((Q, m), np.nan),
            log_posteriors=np.full((Q, m), np.nan),
            model_posteriors=np.full((Q, m), np.nan),
            p_dlas=np.full((Q,), np.nan),
            p_no_dlas=np.full((Q,), np.nan),
            MAP_z_dlas=np.full((Q, md, md), np.nan),
            MAP_log_nhis=np.full((Q, md, md), np.nan),
            num_pixels=np.zeros((Q,), dtype=np.int32),
            status=np.zeros((Q,), dtype=np.int32),
        )
        if keep_samples:
            out["sample_log_likelihoods_dla"] = np.full((Q, S, md), np.nan)
            out["sample_log_likelihoods_lls"] = np.full((Q, S), np.nan)
            out["base_sample_inds"] = np.zeros((Q, S, max(md - 1, 1)), dtype=np.int32)[:, :, : md - 1]
            out["base_sample_inds"] = np.ascontiguousarray(out["base_sample_inds"])
        st = _lib.CatalogueOutputsStruct()
        for name, _ in _lib.CatalogueOutputsStruct._fields_:
            arr = out.get(name)
            if arr is None or arr.size == 0:
                continue
            setattr(st, name, _lib.iptr(arr) if arr.dtype == np.int32 else _lib.dptr(arr))
        return out, st

    # -- the three ways to run ---------------------------------------------------------------------
    def process(self, offsets, wl, fl, nv, pm, z_qsos, keep_samples: bool = False) -> Dict[str, np.ndarray]:
        """Host buffers in, host arrays out (H2D and D2H inside the call)."""
        z_qsos = _lib.f64(z_qsos)
        Q = z_qsos.shape[0]
        pri = _lib.f64(self._priors(z_qsos))
        out, st = self._alloc_outputs(Q, keep_samples)
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        _lib.check(
            _lib.load_library().dla_catalogue_process(
                self._cat.ptr, Q, offsets.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)), _lib.dptr(wl), _lib.dptr(fl),
                _lib.dptr(nv), _lib.bptr(pm), _lib.dptr(z_qsos), _lib.dptr(pri), ctypes.byref(st),
            )
        )
        return self._finish(out)

    def stage(self, offsets, wl, fl, nv, pm, z_qsos) -> None:
        """Upload a set of spectra once; `run_staged` then measures the device path alone."""
        z_qsos = _lib.f64(z_qsos)
        pri = _lib.f64(self._priors(z_qsos))
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        _lib.check(
            _lib.load_library().dla_catalogue_stage(
                self._cat.ptr, z_qsos.shape[0], offsets.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)), _lib.dptr(wl),
                _lib.dptr(fl), _lib.dptr(nv), _lib.bptr(pm), _lib.dptr(z_qsos), _lib.dptr(pri),
            )
        )
        self._staged = z_qsos.shape[0]

    def run_staged(self, keep_samples: bool = False) -> Dict[str, np.ndarray]:
        assert self._staged, "call stage() first"
        out, st = self._alloc_outputs(self._staged, keep_samples)
        _lib.check(_lib.load_library().dla_catalogue_run_staged(self._cat.ptr, ctypes.byref(st)))
        return self._finish(out)

    def last_timing(self) -> Dict[str, float]:
        total, gram, voigt, flops = ctypes.c_double(), ctypes.c_double(), ctypes.c_double(), ctypes.c_double()
        launches = ctypes.c_longlong()
        _lib.check(
            _lib.load_library().dla_catalogue_last_timing(
                self._cat.ptr, ctypes.byref(total), ctypes.byref(gram), ctypes.byref(voigt), ctypes.byref(launches),
                ctypes.byref(flops),
            )
        )
        ev, mk = ctypes.c_longlong(), ctypes.c_longlong()
        _lib.check(_lib.load_library().dla_catalogue_last_counts(self._cat.ptr, ctypes.byref(ev), ctypes.byref(mk)))
        return dict(total_ms=total.value, likelihood_ms=gram.value, voigt_ms=voigt.value, launches=launches.value,
                    likelihood_flops=flops.value, evaluations=ev.value, evaluations_masked=mk.value)

    def _finish(self, out: Dict[str, np.ndarray]) -> Dict[str, np.ndarray]:
        """Split the (Q, 2+max) arrays into the reference's dataset names (run_bayes_select.py:197-209)."""
        md = self.max_dlas
        for name in ("log_priors", "log_likelihoods", "log_posteriors"):
            a = out[name]
            out[name + "_no_dla"] = a[:, 0]
            out[name + "_lls"] = a[:, 1]
            out[name + "_dla"] = a[:, -md:]
        return out


def _chunk_arrays(source, start: int, stop: int, read_spec: Callable, qso_list):
    """(offsets, wl, fl, nv, pm) of spectra [start, stop): views of a preloaded store, or read + pack."""
    if source is not None:
        return source.chunk(start, stop)
    return CatalogueProcessor.pack([read_spec(item) for item in qso_list[start:stop]])


def process_qso(
    qso_list: List,
    z_qso_list: List,
    read_spec: Optional[Callable] = None,
    max_dlas: int = 4,
    broadening: bool = True,
    plot_figures: bool = False,
    *,
    params: Optional[Parameters] = None,
    prior=None,
    model: Dict[str, np.ndarray] = None,
    dla_samples=None,
    subdla_samples=None,
    keep_samples: bool = False,
    batch_spectra: int = 128,
    chunk_spectra: int = 4096,
    out_dir: Optional[str] = None,
    out_filename: Optional[str] = None,
    resume: bool = True,
    preloaded=None,
    rank: int = 0,
    world_size: int = 1,
    processor: Optional["CatalogueProcessor"] = None,
    writer=None,
) -> Dict[str, np.ndarray]:
    """
    Process every spectrum of `qso_list` (run_bayes_select.py:32-295).

    `read_spec(item)` returns (wavelengths, flux, noise_variance, pixel_mask) as the reference's
    readers do (default: read_spec.read_spec, as in the reference); alternatively pass
    `preloaded=` a preload.PreloadedSpectra store (qso_list may then be the store's own list).
    The learned model, prior and sample objects are passed in (the reference loads them from .mat
    files that are not part of this repository).

    The catalogue streams through the device engine in chunks of `chunk_spectra` spectra: each chunk
    is read (or mapped), processed and - when `out_dir` is given - written as one chunk file with a
    manifest, so an interrupted run resumes at the first missing chunk (`resume=True`) and host
    memory holds one chunk of spectra, never the catalogue.  With `out_dir`, the reference's output
    file `processed_qsos_multi_meanflux.h5` (dataset names of run_bayes_select.py:248-295; `.npz`
    with the same keys when h5py is not installed) is written at the end and its path is returned as
    `out["output_file"]`.

    Returns the per-quasar arrays under the reference's dataset names.  The per-sample arrays
    (`sample_log_likelihoods_dla`, `base_sample_inds`, `sample_log_likelihoods_lls`: 0.3 MB per
    quasar, 51 GB at 160k) are produced only with `keep_samples=True`; they are returned in memory
    when no `out_dir` is given and otherwise live in the chunk files / the merged file only.
    """
    if plot_figures:
        raise NotImplementedError("plotting is outside the hot path")
    from .catalogue_io import ChunkedCatalogueWriter

    if read_spec is None and preloaded is None:
        from .read_spec import read_spec as _default_read_spec  # the reference's default argument

        read_spec = _default_read_spec
    params = params or Parameters()
    z_all = np.asarray(z_qso_list, dtype=np.float64)
    Q = len(qso_list)
    if z_all.shape != (Q,):
        raise ValueError("qso_list and z_qso_list must have the same length")
    if preloaded is not None and len(preloaded) != Q:
        raise ValueError("the preloaded store holds %d spectra, qso_list has %d" % (len(preloaded), Q))
    proc = processor or CatalogueProcessor(params, prior, model, dla_samples, subdla_samples, max_dlas, broadening,
                                           batch_spectra=batch_spectra)
    if writer is None and out_dir is not None:
        writer = ChunkedCatalogueWriter(out_dir, qso_list, z_all, params, max_dlas, chunk_spectra, keep_samples, resume)
    if writer is not None:
        todo = writer.chunks(rank, world_size)
    else:
        nchunks = (Q + chunk_spectra - 1) // chunk_spectra
        todo = [(i, i * chunk_spectra, min((i + 1) * chunk_spectra, Q)) for i in range(nchunks) if i % world_size == rank]

    in_memory: Dict[str, List[np.ndarray]] = {}
    for idx, start, stop in todo:
        offsets, wl, fl, nv, pm = _chunk_arrays(preloaded, start, stop, read_spec, qso_list)
        res = proc.process(offsets, wl, fl, nv, pm, z_all[start:stop], keep_samples=keep_samples)
        if writer is not None:
            writer.write_chunk(idx, res, update_manifest=(world_size == 1))
        else:
            for k, v in res.items():
                in_memory.setdefault(k, []).append(v)
    if writer is None:
        out = {k: np.concatenate(v, axis=0) for k, v in in_memory.items()}
        if world_size == 1:
            out["z_qsos"] = z_all
        return out
    if world_size > 1:
        return {"writer": writer}  # process_qso_sharded merges after the barrier
    out = writer.load_merged()
    out["output_file"] = writer.merge(out_filename)
    return out


# ---------------------------------------------------------------------------------------------------
# multi-GPU: spectra are independent, so a catalogue is block-partitioned over the ranks
# (one process per GPU); there is no collective on the data path, only a final gather of the
# (num_quasars, ...) result arrays on rank 0 - the in-box replacement of the reference's SLURM
# job array + sbatch_reunion merge (slurm/submit_gp_find_lls.sh, CDDF_analysis/sbatch_reunion.py:13-63).
# ---------------------------------------------------------------------------------------------------
def _dist_device(group=None):
    """Tensors of a collective live on the GPU for NCCL and on the host for gloo."""
    import torch
    import torch.distributed as dist

    return torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")


def gather_results(local: Dict[str, np.ndarray], num_items: int, rank: int, world_size: int,
                   group=None) -> Optional[Dict[str, np.ndarray]]:
    """
    Concatenate the per-rank result dictionaries in spectrum order on rank 0 (None elsewhere).
    Every array whose leading dimension is the rank's shard length is gathered; the shards
    must follow `shard_range`.  A rank with an empty shard passes an empty dictionary.

    One `torch.distributed.gather` per array into preallocated tensors (no pickling of the data):
    rank 0 - whose block-partition shard is never empty - announces the array names, dtypes and
    trailing shapes; every rank pads its shard to the longest shard and rank 0 trims.
    """
    if world_size == 1:
        return local
    import torch
    import torch.distributed as dist

    start, stop = shard_range(num_items, rank, world_size)
    mine = {k: np.ascontiguousarray(v) for k, v in local.items()
            if stop > start and isinstance(v, np.ndarray) and v.ndim >= 1 and v.shape[0] == stop - start}
    spec = [[(k, mine[k].dtype.str, tuple(mine[k].shape[1:])) for k in sorted(mine)]] if rank == 0 else [None]
    dist.broadcast_object_list(spec, src=0, group=group)  # names and shapes only: a few hundred bytes
    spec = spec[0]
    if stop > start and sorted(mine) != [k for k, _, _ in spec]:
        raise RuntimeError("rank %d returned a different set of result arrays than rank 0" % rank)
    dev = _dist_device(group)
    longest = shard_range(num_items, 0, world_size)[1]  # rank 0 holds a longest shard
    out = {} if rank == 0 else None
    for key, dtype, tail in spec:
        np_dtype = np.dtype(dtype)
        buf = np.zeros((longest,) + tuple(tail), dtype=np_dtype)
        if stop > start:
            if mine[key].dtype != np_dtype or tuple(mine[key].shape[1:]) != tuple(tail):
                raise RuntimeError("rank %d: array %s has dtype/shape %s%s, rank 0 has %s%s"
                                   % (rank, key, mine[key].dtype, mine[key].shape[1:], np_dtype, tuple(tail)))
            buf[: stop - start] = mine[key]
        t = torch.from_numpy(buf).to(dev)
        parts = [torch.empty_like(t) for _ in range(world_size)] if rank == 0 else None
        dist.gather(t, parts, dst=0, group=group)
        if rank == 0:
            pieces = []
            for r, part in enumerate(parts):
                a, b = shard_range(num_items, r, world_size)
                pieces.append(part[: b - a].cpu().numpy())
            out[key] = np.concatenate(pieces, axis=0)
    return out


def process_qso_sharded(
    qso_list: List,
    z_qso_list: List,
    read_spec: Optional[Callable] = None,
    max_dlas: int = 4,
    broadening: bool = True,
    *,
    rank: Optional[int] = None,
    world_size: Optional[int] = None,
    group=None,
    process_fn: Optional[Callable] = None,
    **kwargs,
) -> Optional[Dict[str, np.ndarray]]:
    """
    `process_qso` over the ranks of an initialised torch.distributed job (one rank per GPU); rank 0
    returns the whole catalogue (None on the other ranks).  Spectra are independent, so there is no
    collective on the data path.

    * without `out_dir`: rank r processes the spectra of shard_range(Q, r, world_size) and the
      per-quasar arrays are gathered on rank 0 (`gather_results`); per-sample arrays are gathered too
      when `keep_samples=True` - use `out_dir` for large catalogues;
    * with `out_dir=...`: the chunks of the run are dealt round-robin to the ranks, every rank writes its
      own chunk files, and after a barrier rank 0 records them in the manifest and writes the merged
      output file (the in-box replacement of the reference's SLURM job array + sbatch_reunion merge,
      slurm/submit_gp_find_lls.sh, CDDF_analysis/sbatch_reunion.py:13-63).  Resume works across a
      different number of ranks.

    `process_fn` (default `process_qso`) is called as process_fn(sub_list, sub_z, read_spec, max_dlas,
    broadening, **kwargs).
    """
    import os

    if rank is None:
        rank = int(os.environ.get("RANK", "0"))
    if world_size is None:
        world_size = int(os.environ.get("WORLD_SIZE", "1"))
    Q = len(qso_list)
    assert len(z_qso_list) == Q
    fn = process_fn or process_qso
    if kwargs.get("out_dir") is not None and process_fn is None:
        if world_size == 1:
            return fn(qso_list, z_qso_list, read_spec, max_dlas, broadening, **kwargs)
        import torch.distributed as dist
        from .catalogue_io import ChunkedCatalogueWriter

        def make_writer(resume):
            return ChunkedCatalogueWriter(kwargs["out_dir"], qso_list, z_qso_list, kwargs.get("params") or Parameters(),
                                          max_dlas, kwargs.get("chunk_spectra", 4096), kwargs.get("keep_samples", False),
                                          resume)

        # rank 0 creates (or validates) the manifest; the others read it after the barrier, so every rank
        # skips the same finished chunks
        writer = make_writer(kwargs.get("resume", True)) if rank == 0 else None
        dist.barrier(group=group)
        if rank != 0:
            writer = make_writer(True)
        fn(qso_list, z_qso_list, read_spec, max_dlas, broadening, rank=rank, world_size=world_size, writer=writer,
           **kwargs)
        dist.barrier(group=group)  # every rank has written its chunk files
        if rank != 0:
            return None
        writer.adopt_chunks_on_disk()
        out = writer.load_merged()
        out["output_file"] = writer.merge(kwargs.get("out_filename"))
        return out
    start, stop = shard_range(Q, rank, world_size)
    local = {}
    if stop > start:
        kw = dict(kwargs)
        if kw.get("preloaded") is not None:
            kw["preloaded"] = kw["preloaded"].view(start, stop)
        local = fn(qso_list[start:stop], z_qso_list[start:stop], read_spec, max_dlas, broadening, **kw)
        local.pop("z_qsos", None)
    out = gather_results(local, Q, rank, world_size, group)
    if out is not None and world_size > 1:
        out["z_qsos"] = np.asarray(z_qso_list, dtype=np.float64)
    return out
