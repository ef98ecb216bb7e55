#!/usr/bin/env python
"""
bench.py : spectra/sec of the per-spectrum Bayesian model-selection hot path.

  python bench.py --gpus N --steps K --warmup W [--config 1|3|4]     (N > 1: launched by torchrun)
  python bench.py --impl reference --steps K --warmup W [--config ...]   (CPU arm on the host cores)

--config selects the BASELINE.json workload (default 1, the configuration the metric is quoted on):
  1  configs[1]: 1 000 synthetic BOSS-like spectra per GPU per step, Ho-Bird-Garnett multi-DLA model selection,
     max_dlas = 4, 10 000 DLA + 10 000 subDLA QMC samples, num_lines = 3, k = 20
  3  configs[3]: full Lyman series (num_lines = 31, instrumental broadening), 30 000 + 30 000 samples, 256 spectra
  4  configs[4]: quasar-redshift estimation (ZGP), 10 000 z_QSO samples per spectrum, 1 000 spectra
(configs[2] is configs[1] sharded over the GPUs of a box: `--gpus N`, and tools/run_config2.py for the
 full catalogue path with the chunked writer.)

A step is one pass of the whole path over the step's spectra.
  value : spectra/s with the step's inputs already resident in HBM, timed by CUDA events on the library's
          stream, max over ranks, whole job.
  e2e   : the same through the public call with pinned HOST buffers (dla_catalogue_process /
          dla_zqso_inference): H2D of the spectra and D2H of the result arrays inside the timed region.
  roofline : FP64 tensor (DMMA) roofline of the dominant kernel: algorithmic flops / its
          CUDA-event time, against the FP64 peak measured on the same GPU in the same run.
  cpu_baseline / --impl reference : the reference's OWN classes (NullGP, SubDLAGP, DLAGP, BayesModelSelect, ZGP;
          unmodified files staged in oracle/_ref by oracle/make_ref.sh, kind "reference"), one spectrum per
          worker process over min(host cores, 16) workers, on the FIRST spectra of the GPU arm's list.  One
          spectrum costs the reference 35-65 s (config 1), so a step runs the reference with the first S/f of
          the QMC samples (cost is linear in S: 5 S + 1 likelihood evaluations per spectrum) and the value is
          scaled by 1/f; f is chosen so that the K + W steps fit the time budget and is stated in
          cpu_baseline.sample.  Falls back to the NumPy port (oracle/dla_oracle.py, kind "port") when oracle/_ref
          is absent.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

UNIT = "spectra/s"
MAX_CPU_WORKERS = 16

CONFIGS = {
    1: dict(
        label="configs[1]: synthetic BOSS-like spectra, max_dlas=4, 10k DLA + 10k subDLA samples, num_lines=3, k=20",
        metric="spectra/sec (max_dlas=4, 10k QMC samples)",
        kind="dla", S=10000, max_dlas=4, num_lines=3, spectra=1000, batch=128,
        ref_seconds=70.0,   # live reference, slowest of 16 spectra at full S (measured: 34-66 s per spectrum per core)
        port_seconds=16.0,
    ),
    3: dict(
        label="configs[3]: full Lyman series, num_lines=31 with instrumental broadening, max_dlas=4, "
              "30k DLA + 30k subDLA samples, k=20",
        metric="spectra/sec (max_dlas=4, 30k QMC samples, num_lines=31)",
        kind="dla", S=30000, max_dlas=4, num_lines=31, spectra=256, batch=64,
        ref_seconds=1100.0,  # 330 000 Voigt profiles of 31 lines (2 ms each) + 150 001 likelihoods per spectrum
        port_seconds=300.0,
    ),
    4: dict(
        label="configs[4]: quasar redshift estimation (ZGP), 10k z_QSO samples per spectrum, normalisation 1176-1256 A, k=20",
        metric="spectra/sec (ZGP, 10k z_QSO samples)",
        kind="zqso", S=10000, spectra=1000, batch=0,
        ref_seconds=36.0,    # 2.8 ms per redshift sample
        port_seconds=36.0,
    ),
}
FRACTIONS = (1, 2, 4, 5, 10, 20, 50, 100, 200)


def workload_config(cfg_id, spectra_per_gpu):
    """The `config` object of the JSON line: a description of the workload only, identical in both arms."""
    c = CONFIGS[cfg_id]
    out = {
        "workload": c["label"],
        "bench_config": cfg_id,
        "spectra_per_step_per_gpu": int(spectra_per_gpu),
        "num_samples": c["S"],
        "k": 20,
        "l2": "working set per batch (profile cache / spectra, GBs) exceeds L2; no flush needed",
        "parallelism": "spectra sharded over the GPUs, one process per GPU, no collective on the data path",
    }
    if c["kind"] == "dla":
        out.update(max_dlas=c["max_dlas"], num_lines=c["num_lines"])
    return out


# ----------------------------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """Sample SM clock and throttle reasons of one GPU every 100 ms through NVML."""

    REASONS = {
        0x8: "hw_slowdown",
        0x40: "hw_thermal_slowdown",
        0x20: "sw_thermal_slowdown",
        0x4: "sw_power_cap",
        0x80: "hw_power_brake_slowdown",
    }

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        while not self._stop_evt.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.handle, self.nv.NVML_CLOCK_SM))
                mask = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
                for bit, name in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop_evt.wait(0.1)

    def stop(self):
        self._stop_evt.set()
        if self.ok:
            self.join(timeout=2)
        return {
            "sm_mhz": float(np.median(self.samples)) if self.samples else None,
            "sm_max_mhz": self.max_mhz,
            "reasons": sorted(self.reasons),
        }


# ----------------------------------------------------------------------------------------------
# CPU arm: the reference's own classes (oracle/_ref) or the NumPy port, on the host cores
# ----------------------------------------------------------------------------------------------
_CPU_CTX = {}


def _cpu_init(num_threads):
    os.environ["OMP_NUM_THREADS"] = str(num_threads)
    os.environ["OPENBLAS_NUM_THREADS"] = str(num_threads)
    os.environ["MKL_NUM_THREADS"] = str(num_threads)
    try:
        from threadpoolctl import threadpool_limits

        _CPU_CTX["limit"] = threadpool_limits(limits=num_threads)
    except Exception:
        pass
    from oracle import dla_oracle, ref_loader  # noqa: F401  (imports + page-in happen here, untimed)

    if ref_loader.reference_available():
        ref_loader.load_reference()


def _cpu_one(job):
    """One spectrum through the CPU implementation; returns a scalar so that nothing large is pickled back."""
    import contextlib
    import io

    from oracle import dla_oracle, ref_loader, zqso_oracle

    kind, impl = job["kind"], job["impl"]
    if kind == "dla":
        if impl == "reference":
            out = ref_loader.run_reference_spectrum(job["model"], job["dla"], job["sub"], job["prior"], job["spectrum"],
                                                    job["z_qso"], job["S"], job["max_dlas"], job["num_lines"], True)
        else:
            wl, fl, nv, pm = job["spectrum"]
            out = dla_oracle.process_spectrum(job["model"], job["dla"], job["sub"], job["prior"].less_ind(job["z_qso"]),
                                              wl, fl, nv, pm, job["z_qso"], job["max_dlas"], job["num_lines"], True)
        return float(out["p_dla"])
    if impl == "reference":
        with contextlib.redirect_stdout(io.StringIO()):  # ZGP.inference_z_qso prints the MAP
            _, z_map = ref_loader.run_reference_zqso(job["model"], job["spectrum"], job["S"])
        return float(z_map)
    wl, fl, nv, pm = job["spectrum"]
    zs = np.linspace(2.14, 6.16, job["S"])  # zqso_samples.py:26-29 defaults
    res = zqso_oracle.inference_z_qso(job["model"], wl, fl, nv, pm, zs)
    return float(res[1] if isinstance(res, tuple) else res["z_map"])


class CpuArm:
    """A pool of single-threaded workers and the jobs of one step (the first `workers` spectra of rank 0's list)."""

    def __init__(self, cfg_id, fraction=None, steps_planned=1, budget_s=600.0, workers=None):
        import multiprocessing as mp

        from gpy_dla_detection_b200 import synthetic
        from oracle import ref_loader

        c = CONFIGS[cfg_id]
        self.cfg = c
        self.workers = workers or max(1, min(os.cpu_count() or 1, MAX_CPU_WORKERS))
        self.impl = "reference" if ref_loader.reference_available() else "port"
        per_spectrum = c["ref_seconds"] if self.impl == "reference" else c["port_seconds"]
        if fraction is None:
            fraction = FRACTIONS[-1]
            for f in FRACTIONS:
                if steps_planned * per_spectrum / f <= budget_s:
                    fraction = f
                    break
        n = self.workers
        if c["kind"] == "dla":
            self._inputs = synthetic.make_workload(n, 0, c["S"], c["num_lines"])
        else:
            self._inputs = synthetic.make_zqso_workload(n, 0)
        self.set_fraction(fraction)
        for var in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
            os.environ[var] = "1"  # inherited by the spawned workers before they import NumPy
        self.pool = mp.get_context("spawn").Pool(self.workers, initializer=_cpu_init, initargs=(1,))
        self.pool.map(abs, range(self.workers))  # workers up, imports done (untimed)

    def set_fraction(self, fraction):
        """(Re)build the step's jobs for 1/fraction of the samples of every spectrum."""
        c, n = self.cfg, self.workers
        self.fraction = int(fraction)
        self.S = max(c["S"] // self.fraction, 8)
        if c["kind"] == "dla":
            params, model, prior, dla, sub, z_qsos, spectra = self._inputs
            dla_f = {k: (v[: self.S] if np.ndim(v) else v) for k, v in dla.items()}
            sub_f = {k: (v[: self.S] if np.ndim(v) else v) for k, v in sub.items()}
            self.jobs = [dict(kind="dla", impl=self.impl, model=model, dla=dla_f, sub=sub_f, prior=prior,
                              spectrum=spectra[i], z_qso=float(z_qsos[i]), S=self.S, max_dlas=c["max_dlas"],
                              num_lines=c["num_lines"]) for i in range(n)]
        else:
            model, z_true, spectra = self._inputs
            self.jobs = [dict(kind="zqso", impl=self.impl, model=model, spectrum=spectra[i], S=self.S) for i in range(n)]

    def step(self):
        """One pass over the step's spectra; returns seconds."""
        t0 = time.perf_counter()
        self.pool.map(_cpu_one, self.jobs, chunksize=1)
        return time.perf_counter() - t0

    def rate(self, seconds):
        """spectra/s: `workers` spectra at 1/fraction of their samples each."""
        return self.workers / self.fraction / seconds

    def describe(self, seconds):
        c = self.cfg
        what = ("the reference's own classes (unmodified gpy_dla_detection from oracle/_ref)" if self.impl == "reference"
                else "the NumPy port oracle/dla_oracle.py (oracle/_ref not staged)")
        if self.fraction == 1:
            part = "full S=%d" % c["S"]
        else:
            what_samples = "QMC samples per model" if c["kind"] == "dla" else "z_QSO samples (same prior range, coarser grid)"
            part = ("%d of the %d %s (1/%d of the per-spectrum work, which is linear in the sample count; "
                    "value scaled by 1/%d)" % (self.S, c["S"], what_samples, self.fraction, self.fraction))
        return ("%d spectra per step = the first %d of the GPU arm's list, one per worker process, %s, %s; %.1f s per step"
                % (self.workers, self.workers, what, part, seconds))

    def close(self):
        self.pool.close()
        self.pool.join()


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    c = CONFIGS[args.config]
    arm = CpuArm(args.config, fraction=args.ref_fraction, steps_planned=args.steps + args.warmup, budget_s=args.ref_budget)
    times = []
    done = 0
    if args.ref_fraction is None and args.warmup >= 1:
        # The first warm-up step doubles as a calibration of THIS host: the static per-spectrum estimates come from the
        # build container, whose cores are half as fast as the B200 box's.  Use the largest share of the samples
        # (smallest divisor) with which the remaining steps still fit the budget.
        t_probe = arm.step()
        done = 1
        t_full = t_probe * arm.fraction
        remaining = args.steps + args.warmup - 1
        for f in FRACTIONS:
            if remaining * t_full / f <= args.ref_budget - t_probe:
                if f != arm.fraction:
                    arm.set_fraction(f)
                break
    for it in range(done, args.warmup + args.steps):
        dt = arm.step()
        if it >= args.warmup:
            times.append(dt)
    arm.close()
    total = float(np.sum(times))
    value = arm.rate(total / len(times))
    line = {
        "impl": "reference",
        "metric": c["metric"],
        "value": value,
        "unit": UNIT,
        "n_gpus": args.gpus,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": 1e3 * total / len(times),
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "f64",
        "data": "synthetic",
        "config": workload_config(args.config, args.spectra or c["spectra"]),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": arm.workers, "kind": arm.impl,
                         "sample": arm.describe(total / len(times)), "sample_fraction": 1.0 / arm.fraction},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ----------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------
def _dist_setup():
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    return rank, world, local_rank, barrier, max_over_ranks


def _pinned(a):
    import torch

    t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    return t, t.numpy()


def _traffic_from_profiles(name):
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture (not measured in this run)."""
    prof = os.path.join(ROOT, "profiles", name)
    if os.path.exists(prof):
        try:
            d = json.load(open(prof))
            return d.get("dram_bytes_per_launch"), "ncu --set full capture profiles/%s (%s); not measured in this run" % (
                name, d.get("capture", "see file"))
        except Exception:
            pass
    return None, "no capture committed"


def _cpu_baseline(args):
    if args.skip_cpu:
        return None
    arm = CpuArm(args.config, fraction=args.ref_fraction, steps_planned=1, budget_s=25.0)
    dt = arm.step()
    arm.close()
    return {"value": arm.rate(dt), "unit": UNIT, "cores": arm.workers, "kind": arm.impl, "sample": arm.describe(dt),
            "sample_fraction": 1.0 / arm.fraction}


def run_gpu_dla(args):
    import ctypes

    import torch.distributed as dist

    import __graft_entry__ as graft

    c = CONFIGS[args.config]
    rank, world, local_rank, barrier, max_over_ranks = _dist_setup()
    if rank == 0:
        graft.build()
    if world > 1:
        dist.barrier()

    from gpy_dla_detection_b200 import _lib, synthetic
    from gpy_dla_detection_b200.dla_samples import DLASamplesArrays
    from gpy_dla_detection_b200.run_bayes_select import CatalogueProcessor
    from gpy_dla_detection_b200.subdla_samples import SubDLASamplesArrays

    _lib.init(local_rank)
    lib = _lib.load_library()

    Q = args.spectra or c["spectra"]
    batch = args.batch or c["batch"]
    params, model, prior, dla, sub, z_qsos, spectra = synthetic.make_workload(Q, rank, c["S"], c["num_lines"])
    dla_s = DLASamplesArrays(params, prior, dla["offset_samples"], dla["log_nhi_samples"], dla["nhi_samples"])
    sub_s = SubDLASamplesArrays(params, prior, sub["offset_samples"], sub["log_nhi_samples"], sub["nhi_samples"],
                                sub["Z_lls"], sub["Z_dla"])
    proc = CatalogueProcessor(params, prior, model, dla_s, sub_s, c["max_dlas"], True, batch_spectra=batch)

    # pinned host buffers for the end-to-end leg
    offsets, wl, fl, nv, pm = proc.pack(spectra)
    keep = [_pinned(a) for a in (wl, fl, nv, pm)]
    wl_p, fl_p, nv_p, pm_p = (k[1] for k in keep)

    # ---- FP64 peaks on this GPU, now -----------------------------------------------------------
    dfma, dmma = ctypes.c_double(), ctypes.c_double()
    _lib.check(lib.dla_measure_fp64_peaks(ctypes.byref(dfma), ctypes.byref(dmma)))

    # ---- device-resident leg ---------------------------------------------------------------------
    proc.stage(offsets, wl_p, fl_p, nv_p, pm_p, z_qsos)
    for _ in range(args.warmup):
        proc.run_staged(keep_samples=False)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = lib.dla_kernel_launch_count()
    dev_ms, lik_ms, voigt_ms, lik_flops = 0.0, 0.0, 0.0, 0.0
    evals, evals_masked = 0, 0
    t_wall0 = time.perf_counter()
    for _ in range(args.steps):
        out = proc.run_staged(keep_samples=False)
        tm = proc.last_timing()
        dev_ms += tm["total_ms"]
        lik_ms += tm["likelihood_ms"]
        voigt_ms += tm["voigt_ms"]
        lik_flops += tm["likelihood_flops"]
        evals += tm["evaluations"]
        evals_masked += tm["evaluations_masked"]
    barrier()
    wall_ms = 1e3 * (time.perf_counter() - t_wall0)
    clocks = sampler.stop()
    launches = lib.dla_kernel_launch_count() - launches0
    dev_ms_max = max_over_ranks(dev_ms)
    value = world * Q * args.steps / (dev_ms_max * 1e-3)

    # ---- end-to-end leg: host buffers in, result arrays out -----------------------------------------
    for _ in range(min(args.warmup, 1)):
        proc.process(offsets, wl_p, fl_p, nv_p, pm_p, z_qsos, keep_samples=False)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        out = proc.process(offsets, wl_p, fl_p, nv_p, pm_p, z_qsos, keep_samples=False)
    barrier()
    e2e_ms = max_over_ranks(1e3 * (time.perf_counter() - t0))
    e2e_value = world * Q * args.steps / (e2e_ms * 1e-3)
    h2d = int(wl_p.nbytes + fl_p.nbytes + nv_p.nbytes + pm_p.nbytes + z_qsos.nbytes + Q * (2 + c["max_dlas"]) * 8)
    d2h = int(sum(v.nbytes for k, v in out.items() if isinstance(v, np.ndarray) and not k.endswith(("_no_dla", "_lls", "_dla"))))

    mean_pixels = float(np.mean(out["num_pixels"]))
    n_ok = int(np.sum(out["status"] == 0))

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel -------------------------------------------------------------
    achieved = lik_flops / (lik_ms * 1e-3) / 1e12 if lik_ms > 0 else 0.0
    peak = dmma.value
    traffic, traffic_source = _traffic_from_profiles("likelihood_kernel_dram_bytes.json") if args.config == 1 else (None, "no capture for this config")
    roofline = {
        "bound": "tensor",
        "kernel": "sample_likelihood_kernel (FP64 DMMA m8n8k4)",
        "achieved": achieved,
        "peak": peak,
        "unit": "TFLOP/s",
        "frac": achieved / peak if peak > 0 else None,
        "traffic": traffic,
        "traffic_source": traffic_source,
        "peak_source": "FP64 DMMA peak measured in this run by dla_measure_fp64_peaks (DFMA %.1f, DMMA %.1f TFLOP/s); "
                       "MEASURED_PEAKS.json has no FP64 entry" % (dfma.value, dmma.value),
        "kernel_share_of_step": lik_ms / dev_ms if dev_ms > 0 else None,
        "voigt_share_of_step": voigt_ms / dev_ms if dev_ms > 0 else None,
        "whole_step_frac_of_peak": lik_flops / (dev_ms * 1e-3) / 1e12 / peak if dev_ms > 0 and peak > 0 else None,
        "flops_per_evaluation": "472 n + 3.1e3 (SURVEY.md §8d), counted for the evaluations that are RUN: of the %d per "
                                "spectrum the reference computes, the level >= 1 samples its separation test overwrites "
                                "with NaN (dla_gp.py:164-177) are not evaluated" % (5 * c["S"] + 1),
        "evaluations_per_spectrum": evals / float(Q * args.steps),
        "evaluations_masked_per_spectrum": evals_masked / float(Q * args.steps),
    }

    line = {
        "metric": c["metric"],
        "value": value,
        "unit": UNIT,
        "n_gpus": world,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": dev_ms_max / args.steps,
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "f64",
        "data": "synthetic",
        "config": workload_config(args.config, Q),
        "engine": {"batch_spectra": batch, "mean_modelled_pixels": mean_pixels, "spectra_ok": n_ok,
                   "pipeline": "2-deep over batches: prep(i+1) ahead of compute(i), H2D(i+2) on a copy stream"},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": int(launches),
        "roofline": roofline,
        "cpu_baseline": _cpu_baseline(args),
        "wall_ms_per_step": wall_ms / args.steps,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def zqso_window_pixels(spectra, z_samples, prm):
    """sum over (spectrum, z) of the modelled pixels n_z (unmasked, inside the rest-frame window; +-1 pixel at the edges)."""
    total = 0.0
    for wl, fl, nv, pm in spectra:
        good = np.concatenate([[0], np.cumsum(~np.asarray(pm, dtype=bool))])
        lo = np.searchsorted(wl, np.maximum(prm.min_lambda * (1 + z_samples), wl[0]), side="right")
        hi = np.searchsorted(wl, np.minimum(prm.max_lambda * (1 + z_samples), wl[-1]), side="left")
        hi = np.maximum(hi, lo)
        total += float(np.sum(good[hi] - good[lo]))
    return total


def run_gpu_zqso(args):
    import ctypes

    import torch.distributed as dist

    import __graft_entry__ as graft

    c = CONFIGS[args.config]
    rank, world, local_rank, barrier, max_over_ranks = _dist_setup()
    if rank == 0:
        graft.build()
    if world > 1:
        dist.barrier()

    from gpy_dla_detection_b200 import _lib, synthetic
    from gpy_dla_detection_b200.zqso_gp import ZGP
    from gpy_dla_detection_b200.zqso_samples import ZSamples
    from gpy_dla_detection_b200.zqso_set_parameters import ZParameters

    _lib.init(local_rank)
    lib = _lib.load_library()
    Q = args.spectra or c["spectra"]
    model, z_true, spectra = synthetic.make_zqso_workload(Q, rank)
    p = ZParameters(num_zqso_samples=c["S"])
    gp = ZGP(p, ZSamples(p), model["rest_wavelengths"], model["mu"], model["M"], model["bluewards_mu"],
             model["redwards_mu"], model["bluewards_sigma"], model["redwards_sigma"])
    zs = np.ascontiguousarray(ZSamples(p).sample_z_qsos(), dtype=np.float64)
    packed = gp.pack(spectra)
    keep = [_pinned(a) for a in packed[1:]]
    packed = (packed[0],) + tuple(k[1] for k in keep)

    dfma, dmma = ctypes.c_double(), ctypes.c_double()
    _lib.check(lib.dla_measure_fp64_peaks(ctypes.byref(dfma), ctypes.byref(dmma)))
    flops_per_step = (472.0 + 42.0) * zqso_window_pixels(spectra, zs, p) + 3.1e3 * Q * c["S"]

    for _ in range(args.warmup):
        gp.inference_packed(packed, zs, keep_samples=False)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = lib.dla_kernel_launch_count()
    kern_ms = 0.0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        out = gp.inference_packed(packed, zs, keep_samples=False)
        kern_ms += gp.last_timing()["kernel_ms"]
    barrier()
    e2e_ms = max_over_ranks(1e3 * (time.perf_counter() - t0))
    clocks = sampler.stop()
    launches = lib.dla_kernel_launch_count() - launches0
    kern_ms_max = max_over_ranks(kern_ms)
    value = world * Q * args.steps / (kern_ms_max * 1e-3)
    e2e_value = world * Q * args.steps / (e2e_ms * 1e-3)
    h2d = int(sum(a.nbytes for a in packed) + zs.nbytes)
    d2h = int(out["z_map"].nbytes + out["map_index"].nbytes)
    hit = float(np.mean(np.abs(out["z_map"] - z_true) < 0.05))
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0
    achieved = flops_per_step * args.steps / (kern_ms * 1e-3) / 1e12
    peak = dmma.value
    traffic, traffic_source = _traffic_from_profiles("zqso_kernel_dram_bytes.json")
    line = {
        "metric": c["metric"],
        "value": value,
        "unit": UNIT,
        "n_gpus": world,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": kern_ms_max / args.steps,
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "f64",
        "data": "synthetic",
        "config": workload_config(args.config, Q),
        "engine": {"z_map_within_0.05_of_truth": hit, "mean_window_pixels": flops_per_step / 514.0 / Q / c["S"]},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": int(launches),
        "roofline": {
            "bound": "tensor",
            "kernel": "zqso_likelihood_kernel (FP64 DMMA m8n8k4)",
            "achieved": achieved,
            "peak": peak,
            "unit": "TFLOP/s",
            "frac": achieved / peak if peak > 0 else None,
            "traffic": traffic,
            "traffic_source": traffic_source,
            "peak_source": "FP64 DMMA peak measured in this run by dla_measure_fp64_peaks (DFMA %.1f, DMMA %.1f TFLOP/s)"
                           % (dfma.value, dmma.value),
            "flops_per_evaluation": "(472 + 42) n_z + 3.1e3 per redshift sample (SURVEY.md §8d: Gram + projection + "
                                    "per-pixel + interpolation of 21 columns); executed MMA work is 768 n_z (24 x 24 lower blocks)",
        },
        "cpu_baseline": _cpu_baseline(args),
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=1, choices=sorted(CONFIGS), help="BASELINE.json configs[] index")
    ap.add_argument("--spectra", type=int, default=0, help="spectra per step per GPU (default: the config's)")
    ap.add_argument("--batch", type=int, default=0, help="spectra resident per device batch (default: the config's)")
    ap.add_argument("--skip-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--ref-fraction", type=int, default=None,
                    help="CPU arm: run 1/N of the QMC samples per spectrum (default: largest share that fits --ref-budget)")
    ap.add_argument("--ref-budget", type=float, default=600.0, help="CPU arm: seconds the K + W steps may take")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    if CONFIGS[args.config]["kind"] == "zqso":
        return run_gpu_zqso(args)
    return run_gpu_dla(args)


if __name__ == "__main__":
    sys.exit(main())
