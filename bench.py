#!/usr/bin/env python
"""Benchmark of the B200 spherical-harmonic synthesis path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

A step is one pass of the hot path over one batch: BASELINE config 2, synthesis of 240 epochs of
degree-96 coefficients onto the 0.5 degree geographic grid with the water-height kernel
(62.2 M grid-points x epochs per GPU).  Epochs are independent, so every rank runs the same
per-GPU batch on its own synthetic epochs with no data-path collective ("weak" scaling; pass
--scaling strong to split the 240 epochs over the ranks instead).

Output: one JSON line (rank 0).  `value` is device-resident throughput, `e2e` goes through the
host-buffer C-ABI call (pinned host memory, H2D + D2H inside the timed region), `roofline`
describes the dominant kernel (FP64 tensor-core longitude contraction), `cpu_baseline` is the
numpy oracle port of the reference's to_grid timed on this box's host cores.
"""
import argparse
import json
import os
import statistics
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NMAX, DGRID, EPOCHS, KERNEL = 96, 0.5, 240, "ewh"
METRIC = "sh_synthesis_grid_points_x_epochs_per_s"
UNIT = "grid-pts*epochs/s"


def algorithmic_flops(nmax, nlat, nlon, epochs):
    """SURVEY 8(d) contract figure: direct two-stage synthesis, no symmetry / FFT credit."""
    L = nmax + 1
    K = L * L
    stage1 = 2.0 * epochs * nlat * K
    stage2 = 2.0 * (2 * L - 1) * epochs * nlat * nlon
    legendre = 5.0 * nlat * L * (L + 1) / 2
    return stage1, stage2, legendre


def synthetic_batch(nmax, epochs, first_epoch=0):
    from oracle import sh_oracle  # input generator only (shared with the tests)
    return np.stack([sh_oracle.synthetic_coefficients(nmax, first_epoch + e) for e in range(epochs)])


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index, period=0.01):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._halt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as exc:  # NVML missing: report that instead of inventing numbers
            self.err = str(exc)

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake_slowdown",
        }
        while not self._halt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(self.period)

    def stop(self):
        self._halt.set()
        self.join(timeout=2)

    def summary(self):
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "note": "no NVML samples"}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def bind_to_gpu_numa_node(index):
    """Run this rank on the host cores NVML lists as local to GPU `index`, so that first-touch places the
    pinned buffers on that NUMA node (8 ranks copying 0.5 GB per step each otherwise meet on one socket)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * w + b for w in range(words) for b in range(64) if (int(mask[w]) >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
    except Exception:
        pass


def cpu_reference_rate(epochs_sample, warm=1):
    """Oracle port of the reference's to_grid (same numpy work per call: Legendre table, factor
    scaling, trig table, L dgemms) on the host cores.  Returns (grid-pts*epochs/s, seconds, threads)."""
    from oracle import sh_oracle as orc
    grid = orc.geographic_grid(DGRID, DGRID)
    nlat, nlon = grid.shape
    batch = synthetic_batch(NMAX, epochs_sample + warm)
    for e in range(warm):
        orc.synthesis(batch[e], grid, KERNEL)
    t0 = time.perf_counter()
    for e in range(warm, warm + epochs_sample):
        orc.synthesis(batch[e], grid, KERNEL)
    dt = time.perf_counter() - t0
    threads = os.cpu_count()
    try:
        from threadpoolctl import threadpool_info
        blas = [p["num_threads"] for p in threadpool_info() if p.get("user_api") == "blas"]
        threads = max(blas) if blas else threads
    except Exception:
        pass
    return epochs_sample * nlat * nlon / dt, dt, threads


def run_reference(args, rank):
    """--impl reference: the reference's own CPU algorithm (numpy oracle port; the reference is
    pure Python and cannot travel to the GPU box) on the same config, metric and unit."""
    if rank != 0:
        return
    # bounded sample per step so that K steps end within a few minutes (0.2 s per epoch on 16 cores)
    sample = max(1, min(8, 120 // max(args.steps, 1)))
    rates, times = [], []
    for _ in range(min(max(args.warmup, 0), 3)):
        cpu_reference_rate(1, warm=1)
    for _ in range(max(args.steps, 1)):
        r, dt, threads = cpu_reference_rate(sample, warm=0)
        rates.append(r)
        times.append(dt)
    nlat, nlon = int(180 / DGRID), int(360 / DGRID)
    total_time = sum(times)
    value = len(times) * sample * nlat * nlon / total_time
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup,
        # a full step is 240 epochs; the sample is 8, the reference's cost is linear in epochs
        "ms_per_step": 1e3 * total_time / len(times) * (EPOCHS / sample),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "config2: degree-96 synthesis of 240 epochs -> 0.5deg GeographicGrid, ewh",
                   "nmax": NMAX, "grid": "geographic 0.5deg (360x720)", "epochs": EPOCHS, "kernel": KERNEL},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": "%d of 240 epochs per step (cost is linear in epochs: no cross-epoch reuse "
                                   "in the reference), numpy/OpenBLAS threads=%d, ~3/4 of the time is "
                                   "single-threaded numpy table building" % (sample, threads)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--e2e-steps", type=int, default=None, help="steps of the host-buffer loop (default min(steps, 30))")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist
    import grates_b200 as gb

    if world != args.gpus and world > 1:
        raise SystemExit("--gpus %d does not match WORLD_SIZE %d" % (args.gpus, world))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    distributed = world > 1
    full_affinity = os.sched_getaffinity(0)
    bind_to_gpu_numa_node(local_rank)          # pinned host buffers of the e2e loop land next to the GPU
    if distributed:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    # ---- workload ---------------------------------------------------------------------------------
    if args.scaling == "weak":
        epochs, first = EPOCHS, rank * EPOCHS
    else:
        per = (EPOCHS + world - 1) // world
        first = rank * per
        epochs = max(0, min(EPOCHS, first + per) - first)
    grid = gb.GeographicGrid(DGRID, DGRID)
    torch.cuda.synchronize()
    t_plan = time.perf_counter()
    plan = gb.get_plan(grid, NMAX, KERNEL, device=local_rank)      # host tables + upload, once per (grid, degree, kernel)
    torch.cuda.synchronize()
    t_plan = time.perf_counter() - t_plan
    nlat, nlon = plan.nlat, plan.nlon
    anm_host = synthetic_batch(NMAX, epochs, first)
    anm = torch.as_tensor(anm_host).to(dev)
    out = torch.empty((epochs, nlat, nlon), dtype=torch.float64, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)       # > 126 MB L2
    units_per_step = epochs * nlat * nlon
    lib = gb._lib.load()

    def barrier():
        if distributed:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up --------------------------------------------------------------------------------------
    for _ in range(max(args.warmup, 3)):
        plan.synthesis(anm, out=out)
    torch.cuda.synchronize()

    # ---- timed region: K steps, device-resident ---------------------------------------------------
    sampler = ClockSampler(local_rank)
    ev0 = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    ev1 = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    plan.set_profiling(args.steps)
    lib.gb_launch_count(1)
    barrier()
    sampler.start()
    t_wall0 = time.perf_counter()
    for s in range(args.steps):
        flush.zero_()                         # L2 flush between steps (not part of the step time)
        ev0[s].record()
        plan.synthesis(anm, out=out)
        ev1[s].record()
    barrier()
    t_wall = time.perf_counter() - t_wall0
    launches = int(lib.gb_launch_count(1))
    step_ms = [a.elapsed_time(b) for a, b in zip(ev0, ev1)]
    stages = plan.stage_times(args.steps)     # [steps, 3] pack / stage 1 / stage 2
    plan.set_profiling(0)
    my_time = sum(step_ms) * 1e-3
    t = torch.tensor([my_time], dtype=torch.float64, device=dev)
    total_units = torch.tensor([float(units_per_step)], dtype=torch.float64, device=dev)
    if distributed:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(total_units, op=dist.ReduceOp.SUM)
    max_time = float(t.item())
    value = float(total_units.item()) * args.steps / max_time

    # ---- end to end: host buffers through gb_synthesis_host ----------------------------------------
    e2e_steps = args.e2e_steps if args.e2e_steps is not None else min(args.steps, 30)
    pin_in = gb.PinnedArray(anm_host.shape)
    pin_in.array[...] = anm_host
    pin_out = gb.PinnedArray((epochs, nlat, nlon))
    for _ in range(2):
        plan.synthesis_host(pin_in.array, out=pin_out.array)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        plan.synthesis_host(pin_in.array, out=pin_out.array)
    torch.cuda.synchronize()
    e2e_time = time.perf_counter() - t0
    te = torch.tensor([e2e_time], dtype=torch.float64, device=dev)
    if distributed:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    sampler.stop()
    e2e_value = float(total_units.item()) * e2e_steps / float(te.item())
    checksum = float(np.abs(pin_out.array[0]).max()) if epochs else 0.0

    if rank == 0:
        # ---- roofline of the dominant kernel (Fourier stage 2, FP64 tensor pipe) ----------------------
        import ctypes
        c_mma, c_fma = ctypes.c_double(0), ctypes.c_double(0)
        gb._lib.check(lib.gb_probe_fp64_peak(local_rank, ctypes.byref(c_mma), ctypes.byref(c_fma)))
        peak = max(c_mma.value, c_fma.value)
        f1, f2, fl = algorithmic_flops(NMAX, nlat, nlon, epochs)
        s2_ms = float(np.mean(stages[:, 2])) if len(stages) else float("nan")
        s1_ms = float(np.mean(stages[:, 1])) if len(stages) else float("nan")
        pk_ms = float(np.mean(stages[:, 0])) if len(stages) else float("nan")
        achieved = f2 / (s2_ms * 1e-3) / 1e12
        # executed multiply-adds of the dominant kernel: the symmetric path contracts one quadrant of meridians
        sym = plan.symmetric
        f2_exec = f2 / 4.0 if sym else f2
        # the folded Legendre stage runs the recursion and its multiply-adds for the northern parallels only
        folded = plan.folded
        f1_exec = f1 / 2.0 if folded else f1
        step_mean_ms = 1e3 * my_time / args.steps
        traffic = None
        tfile = os.path.join(ROOT, "profiles", "stage2_traffic.json")
        if os.path.exists(tfile):
            try:
                traffic = json.load(open(tfile)).get("dram_bytes_per_launch")
            except Exception:
                traffic = None
        peaks_file = os.path.join(ROOT, "MEASURED_PEAKS.json")
        hbm = json.load(open(peaks_file)).get("hbm_gbs") if os.path.exists(peaks_file) else None
        bytes_step = 8.0 * units_per_step + 8.0 * epochs * (NMAX + 1) ** 2
        roofline = {
            "bound": "tensor",
            "kernel": ("gb_fourier_stage2_sym" if sym else "gb_fourier_stage2") + " (FP64 DMMA.8x8x4, SASS-verified)",
            "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "traffic": traffic,
            "peak_source": "live gb_probe_fp64_peak on this GPU (DMMA %.2f / DFMA %.2f TFLOP/s); MEASURED_PEAKS.json "
                           "has no FP64 figure" % (c_mma.value, c_fma.value),
            "algorithmic_flops_per_launch": f2, "kernel_ms": s2_ms,
            "executed_flops_per_launch": f2_exec, "executed_achieved": f2_exec / (s2_ms * 1e-3) / 1e12,
            "executed_frac": f2_exec / (s2_ms * 1e-3) / 1e12 / peak,
            "note": ("declared algorithmic shortcut: four-fold longitude symmetry of the grid (meridians symmetric "
                     "about 0 and under a half turn) -> the kernel executes 1/4 of the contract multiply-adds; "
                     "`achieved`/`frac` use the SURVEY 8(d) contract flops (direct contraction, no symmetry credit), "
                     "`executed_*` what the tensor pipe really did. GB_NO_SYMMETRY=1 runs the direct contraction. "
                     "Second declared shortcut (Legendre stage): parallels mirrored about the equator share one "
                     "recursion, gated on a measured hemisphere asymmetry of the reference's tables (GB_NO_FOLD=1 disables)."
                     if sym else "direct contraction (no symmetry shortcut active)"),
            "step": {"algorithmic_flops": f1 + f2 + fl, "ms": step_mean_ms,
                     "frac_of_fp64_peak": (f1 + f2 + fl) / (step_mean_ms * 1e-3) / 1e12 / peak,
                     "executed_flops": f1_exec + f2_exec + fl,
                     "executed_frac_of_fp64_peak": (f1_exec + f2_exec + fl) / (step_mean_ms * 1e-3) / 1e12 / peak,
                     "legendre_stage_folded_about_equator": bool(folded),
                     "kernel_ms": {"pack": pk_ms, "legendre_stage1": s1_ms, "fourier_stage2": s2_ms}},
            "hbm": {"algorithmic_bytes_per_step": bytes_step, "peak_gbs": hbm,
                    "achieved_gbs": bytes_step / (step_mean_ms * 1e-3) / 1e9,
                    "frac": (bytes_step / (step_mean_ms * 1e-3) / 1e9 / hbm) if hbm else None},
        }
        if args.no_cpu_baseline or world > 1:
            cpu = None
        else:
            os.sched_setaffinity(0, full_affinity)     # the CPU arm gets every host core
            rate, dt, threads = cpu_reference_rate(24)
            cpu = {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
                   "sample": "24 of 240 epochs after 1 warm-up call, %.1f s; oracle port of to_grid "
                             "(numpy+OpenBLAS, %d threads; the reference has no cross-epoch reuse so cost is "
                             "linear in epochs)" % (dt, threads)}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": 1e3 * max_time / args.steps, "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "config2: degree-96 synthesis of 240 epochs -> 0.5deg GeographicGrid, ewh",
                       "nmax": NMAX, "grid": "geographic 0.5deg (360x720)", "epochs_per_gpu": epochs,
                       "kernel": KERNEL, "parallelism": "epochs sharded, no data-path collective",
                       "l2": "256 MiB memset between steps; outputs 498 MB/step exceed the 126 MB L2",
                       "timing": "CUDA events per step on the launch stream, max over ranks of the summed step times",
                       "plan_create_s": t_plan},
            "clocks": sampler.summary(),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(anm_host.nbytes),
                    "d2h_bytes_per_step": int(pin_out.nbytes), "steps": e2e_steps,
                    "path": "SHPlan.synthesis_host -> gb_synthesis_host (pinned host buffers, chunked D2H overlap)",
                    "result_check_max_abs": checksum},
            "gpu_launches": launches,
            "wall_s_timed_region": t_wall,
            "roofline": roofline,
            "cpu_baseline": cpu,
        }
        print(json.dumps(line))
    pin_in.free()
    pin_out.free()
    if distributed:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
