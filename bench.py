#!/usr/bin/env python
"""
bench.py : spectra/sec of the per-spectrum Bayesian model-selection hot path.

  python bench.py --gpus N --steps K --warmup W            (N > 1: launched by torchrun)
  python bench.py --impl reference --steps K --warmup W    (CPU arm: the oracle port on host cores)

Workload (BASELINE.json configs[1]): synthetic BOSS-like spectra, Ho-Bird-Garnett multi-DLA
model selection with max_dlas = 4, 10 000 DLA + 10 000 subDLA QMC samples, num_lines = 3,
synthetic learned model of the published shape (rest grid 911.75:0.25:1215.75, k = 20).
A step is one pass of the whole path (prepare, profiles, 5 x 10 000 + 1 likelihoods,
evidences, resampling, MAP, posteriors) over `--spectra` spectra per GPU.

  value : spectra/s with the step's inputs already resident in HBM (dla_catalogue_run_staged),
          timed by CUDA events on the library's stream, max over ranks, whole job.
  e2e   : the same through dla_catalogue_process with pinned HOST buffers: H2D of the spectra
          and D2H of the result arrays inside the timed region.
  roofline : FP64 tensor (DMMA) roofline of the dominant kernel, sample_likelihood_kernel:
          algorithmic flops (472 n + 3.1e3 per evaluation) / its CUDA-event time,
          against the FP64 peak measured on the same GPU in the same run.
  cpu_baseline : the NumPy oracle port timed on this box's host cores on a bounded sample.
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

METRIC = "spectra/sec (max_dlas=4, 10k QMC samples)"
UNIT = "spectra/s"
S_SAMPLES = 10000
MAX_DLAS = 4
NUM_LINES = 3


# ----------------------------------------------------------------------------------------------
# workload
# ----------------------------------------------------------------------------------------------
def make_workload(num_spectra, seed0=0):
    from gpy_dla_detection_b200 import synthetic
    from gpy_dla_detection_b200.set_parameters import Parameters

    params = Parameters(num_dla_samples=S_SAMPLES, num_lines=NUM_LINES)
    model = synthetic.make_learned_model(0)
    prior = synthetic.SyntheticPrior(params)
    dla = synthetic.make_dla_sample_arrays(params)
    sub = synthetic.make_subdla_sample_arrays(params)
    z_qsos = synthetic.sample_z_qsos(num_spectra, seed=12345 + seed0)
    spectra = [synthetic.make_spectrum(model, z_qsos[i], seed=seed0 * 1000003 + i) for i in range(num_spectra)]
    return params, model, prior, dla, sub, z_qsos, spectra


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
# CPU arm: the oracle port on the host cores
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
    from oracle import dla_oracle  # noqa: F401  (import + page-in happen here, untimed)


def _cpu_one(args):
    from oracle import dla_oracle

    model, dla, sub, counts, spec, z_qso = args
    wl, fl, nv, pm = spec
    out = dla_oracle.process_spectrum(model, dla, sub, counts, wl, fl, nv, pm, z_qso, MAX_DLAS, NUM_LINES, True)
    return float(out["p_dla"])


def cpu_sample_throughput(num_spectra, cores, seed0=7):
    """spectra/s of the oracle port: `num_spectra` spectra over a pool of `cores` single-thread workers."""
    import multiprocessing as mp

    params, model, prior, dla, sub, z_qsos, spectra = make_workload(num_spectra, seed0)
    jobs = [(model, dla, sub, prior.less_ind(z_qsos[i]), spectra[i], float(z_qsos[i])) for i in range(num_spectra)]
    ctx = mp.get_context("spawn")
    with ctx.Pool(cores, initializer=_cpu_init, initargs=(1,)) as pool:
        pool.map(abs, range(cores))  # workers up (imports done in the initializer), untimed
        t0 = time.perf_counter()
        pool.map(_cpu_one, jobs, chunksize=1)
        dt = time.perf_counter() - t0
    return num_spectra / dt, dt


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    per_step = cores  # one spectrum per core per step: a bounded sample of the 1000-spectrum workload
    times = []
    for it in range(args.warmup + args.steps):
        rate, dt = cpu_sample_throughput(per_step, cores, seed0=100 + it)
        if it >= args.warmup:
            times.append(dt)
    total = float(np.sum(times))
    value = per_step * len(times) / total
    sample = "%d spectra per step (one per core), full S=10000, max_dlas=4, oracle/dla_oracle.py" % per_step
    line = {
        "impl": "reference",
        "metric": METRIC,
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
        "config": {"workload": "configs[1]: synthetic BOSS-like spectra, max_dlas=4, 10k DLA + 10k subDLA samples, num_lines=3",
                   "spectra_per_step": per_step},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ----------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------
def run_gpu_arm(args):
    import torch
    import torch.distributed as dist

    import __graft_entry__ as graft

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    if rank == 0:
        graft.build()
    if world > 1:
        dist.barrier()

    from gpy_dla_detection_b200 import _lib
    from gpy_dla_detection_b200.run_bayes_select import CatalogueProcessor
    from gpy_dla_detection_b200.dla_samples import DLASamplesArrays
    from gpy_dla_detection_b200.subdla_samples import SubDLASamplesArrays

    _lib.init(local_rank)
    lib = _lib.load_library()

    Q = args.spectra
    params, model, prior, dla, sub, z_qsos, spectra = make_workload(Q, seed0=rank)
    dla_s = DLASamplesArrays(params, prior, dla["offset_samples"], dla["log_nhi_samples"], dla["nhi_samples"])
    sub_s = SubDLASamplesArrays(params, prior, sub["offset_samples"], sub["log_nhi_samples"], sub["nhi_samples"],
                                sub["Z_lls"], sub["Z_dla"])
    proc = CatalogueProcessor(params, prior, model, dla_s, sub_s, MAX_DLAS, True, batch_spectra=args.batch)

    # pinned host buffers for the end-to-end leg
    offsets, wl, fl, nv, pm = proc.pack(spectra)

    def pinned(a):
        t = torch.from_numpy(a).pin_memory()
        return t, t.numpy()

    keep = [pinned(a) for a in (wl, fl, nv, pm)]
    wl_p, fl_p, nv_p, pm_p = (k[1] for k in keep)

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

    # ---- FP64 peaks on this GPU, now -----------------------------------------------------------
    import ctypes

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
    t_wall0 = time.perf_counter()
    for _ in range(args.steps):
        out = proc.run_staged(keep_samples=False)
        tm = proc.last_timing()
        dev_ms += tm["total_ms"]
        lik_ms += tm["likelihood_ms"]
        voigt_ms += tm["voigt_ms"]
        lik_flops += tm["likelihood_flops"]
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
    h2d = int(wl_p.nbytes + fl_p.nbytes + nv_p.nbytes + pm_p.nbytes + z_qsos.nbytes + Q * (2 + MAX_DLAS) * 8)
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
    traffic = None
    prof = os.path.join(ROOT, "profiles", "likelihood_kernel_dram_bytes.json")
    if os.path.exists(prof):
        try:
            traffic = json.load(open(prof)).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    roofline = {
        "bound": "tensor",
        "kernel": "sample_likelihood_kernel (FP64 DMMA m8n8k4)",
        "achieved": achieved,
        "peak": peak,
        "unit": "TFLOP/s",
        "frac": achieved / peak if peak > 0 else None,
        "traffic": traffic,
        "peak_source": "FP64 DMMA peak measured in this run by dla_measure_fp64_peaks (DFMA %.1f, DMMA %.1f TFLOP/s); "
                       "MEASURED_PEAKS.json has no FP64 entry" % (dfma.value, dmma.value),
        "kernel_share_of_step": lik_ms / dev_ms if dev_ms > 0 else None,
        "voigt_share_of_step": voigt_ms / dev_ms if dev_ms > 0 else None,
        "flops_per_evaluation": "472 n + 3.1e3 (SURVEY.md §8d)",
    }

    # ---- CPU baseline on a bounded sample ---------------------------------------------------------------
    cpu = None
    if not args.skip_cpu and world >= 1:
        cores = os.cpu_count() or 1
        n_cpu = min(cores, 16)
        rate, dt = cpu_sample_throughput(n_cpu, n_cpu)
        cpu = {"value": rate, "unit": UNIT, "cores": n_cpu, "kind": "port",
               "sample": "%d spectra of the same workload (full S=10000, max_dlas=4), one per core, %.1f s" % (n_cpu, dt)}

    line = {
        "metric": METRIC,
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
        "config": {
            "workload": "configs[1]: synthetic BOSS-like spectra, max_dlas=4, 10k DLA + 10k subDLA samples, num_lines=3, k=20",
            "spectra_per_step_per_gpu": Q,
            "batch_spectra": args.batch,
            "mean_modelled_pixels": mean_pixels,
            "spectra_ok": n_ok,
            "l2": "working set per batch (profile cache, GBs) exceeds L2; no flush needed",
            "parallelism": "spectra sharded over %d GPU(s), no collective on the data path" % world,
        },
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": int(launches),
        "roofline": roofline,
        "cpu_baseline": cpu,
        "wall_ms_per_step": wall_ms / args.steps,
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
    ap.add_argument("--spectra", type=int, default=1000, help="spectra per step per GPU")
    ap.add_argument("--batch", type=int, default=128, help="spectra resident per device batch")
    ap.add_argument("--skip-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_gpu_arm(args)


if __name__ == "__main__":
    sys.exit(main())
