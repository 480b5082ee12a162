#!/usr/bin/env python
"""bench.py -- the hot path's headline metric on B200: interpolant evaluations per second.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload NAME]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the evaluator over one batch of synthetic points: the table stays
resident, x (already in HBM) streams in, y streams out.  Default workload = BASELINE.json
configs[1]: Bessel J0 on [0,20], Chebyshev order 6, fp32 table (approx/32o6-p1.19e-06), 2^28
uniform random points per GPU.  Each rank owns its own 2^28-point contiguous shard of the global
array (weak scaling, no data-path collective); time = max over ranks of the device time.

Prints ONE JSON line (rank 0).  Keys beyond the base contract:
  e2e          same metric through the reference-facing C-ABI call on HOST buffers
               (ai_eval_host_*: pinned host memory, H2D + kernel + D2H inside the timed region)
  roofline     algorithmic bytes (2*sizeof(T) per point) / kernel time vs the measured HBM peak
  cpu_baseline the reference's own generated kernel (oracle/_ref) or the C port, on host cores
  clocks       SM clock / throttle reasons sampled (NVML) during the timed region

`--impl reference` times the reference's CPU implementation of the same path instead
(oracle/_ref = generate.gen_cheb's text compiled as C + OpenMP; else the C port of
Approximator.evaluate), all host threads, bounded sample per step.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))

import numpy as np  # noqa: E402

METRIC = "interpolant evals/sec"
UNIT = "Gpt/s"

# name -> (golden table, log2 points per GPU, description)
WORKLOADS = {
    "cfg2_j0_cheb_o6_f32": ("approx__32o6-p1.19209289551e-06", 28,
                            "BASELINE configs[1]: J0 [0,20] Chebyshev o6 fp32 (1.19e-6 table), 2^28 uniform random points per GPU"),
    "cfg3_j0_cheb_o13_f64": ("gen__cfg3_j0_cheb_o13_f64", 28,
                             "BASELINE configs[2]: J0 Chebyshev o13 fp64 (2.2e-14 table), 2^28 points per GPU"),
    "cfg1_sinsin_cheb_o3_f64": ("gen__cfg1_sinsin_cheb_o3_f64", 28, "configs[0] table at GPU size"),
    "cfg4_f1_leg_o7_f32": ("gen__cfg4_f1_leg_o7_f32", 28, "BASELINE configs[3]: deep skewed Legendre o7 tree, fp32"),
    "j0_cheb_o7_f64": ("approx__64o7-p2.22044604925e-14", 28, "J0 Chebyshev o7 fp64 2.2e-14 (64 leaves)"),
}
DEFAULT_WORKLOAD = "cfg2_j0_cheb_o6_f32"
CPU_SAMPLE_LOG2 = 26


NCU_SUMMARIES = {   # workload -> committed `ncu --set full` summary (tools/ncu_summary.py)
    "cfg2_j0_cheb_o6_f32": "r1_ncu_cfg2_cheb_o6_f32.json",
    "cfg3_j0_cheb_o13_f64": "r1_ncu_cfg3_cheb_o13_f64.json",
}


def ncu_traffic(workload, points_per_launch):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the eval kernel, from the committed
    ncu capture of the same workload and launch size (profiles/); None when there is none."""
    name = NCU_SUMMARIES.get(workload)
    path = os.path.join(ROOT, "profiles", name) if name else None
    if not path or not os.path.exists(path):
        return None
    with open(path) as fh:
        d = json.load(fh)
    if d.get("points_per_launch") != points_per_launch or not d.get("launches"):
        return None
    return d["launches"][0]["derived"]["dram_traffic_bytes"]


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            return json.load(fh), "measured"
    return {"hbm_gbs": 6650.0}, "fallback"


class ClockSampler(object):
    """SM clock and throttle reasons sampled while the timed region runs (NVML)."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4),
            "hw_power_brake": getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80),
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if mask & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop.wait(0.005)

    def __enter__(self):
        if self.nv is not None:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thread is not None:
            self._thread.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_reference_pass(golden_rec, x, y, threads):
    """One pass of the reference's CPU implementation over x.  Returns the kind used."""
    import oracle
    if oracle.ref_available() and golden_rec.name in oracle.ref_manifest():
        oracle.ref_eval(golden_rec.name, golden_rec.tree_1d, x, threads=threads, y=y)
        return "reference"
    out, _, _ = oracle.c_eval_tree(golden_rec.tree_1d, golden_rec.max_order, golden_rec.num_levels,
                                   golden_rec.basis, x, threads=threads, want_cond=False)
    y[:] = out
    return "port"


def make_cpu_sample(golden_rec, log2n):
    lo, hi = float(golden_rec.lower_bound), float(golden_rec.upper_bound)
    x = np.random.default_rng(1).uniform(lo, hi, 1 << log2n).astype(golden_rec.dtype)
    x = np.minimum(x, np.nextafter(golden_rec.dtype(hi), golden_rec.dtype(lo)))
    return x, np.empty_like(x)


def time_cpu_baseline(golden_rec, passes=3):
    threads = host_threads()
    x, y = make_cpu_sample(golden_rec, CPU_SAMPLE_LOG2)
    kind = cpu_reference_pass(golden_rec, x, y, threads)              # warm-up (page faults, OMP pool)
    best = float("inf")
    for _ in range(passes):
        t0 = time.perf_counter()
        cpu_reference_pass(golden_rec, x, y, threads)
        best = min(best, time.perf_counter() - t0)
    return {"value": x.size / best / 1e9, "unit": UNIT, "cores": threads, "kind": kind,
            "sample": "2^%d uniform random points of the same workload, best of %d passes, OpenMP static over %d threads"
                      % (CPU_SAMPLE_LOG2, passes, threads)}


def run_reference_arm(args, workload, golden_rec, desc):
    """--impl reference: the reference's CPU path, rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = host_threads()
    x, y = make_cpu_sample(golden_rec, CPU_SAMPLE_LOG2)
    kind = "port"
    for _ in range(max(1, args.warmup)):
        kind = cpu_reference_pass(golden_rec, x, y, threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_reference_pass(golden_rec, x, y, threads)
    dt = time.perf_counter() - t0
    value = x.size * args.steps / dt / 1e9
    sample = ("each step = one pass over 2^%d uniform random points of the workload on %d host threads "
              "(%s)" % (CPU_SAMPLE_LOG2, threads,
                        "reference's generate.gen_cheb kernel text compiled as C, oracle/_ref" if kind == "reference"
                        else "C port of Approximator.evaluate, oracle/ai_oracle.c"))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32" if golden_rec.dtype == np.float32 else "f64", "data": "synthetic",
        "config": {"workload": workload, "description": desc, "table": golden_rec.name,
                   "points_per_step": int(x.size)},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--log2-points", type=int, default=None, help="points per GPU (default: workload's)")
    ap.add_argument("--order", default="random", choices=["random", "sorted"],
                    help="point order: uniform random (run_test.py 'random' shuffle) or sorted linspace")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    args.warmup = max(3, args.warmup)

    import golden_io                               # fixture loader (tests/golden/golden_io.py)
    from adaptive_interpolation_b200 import tables
    table_name, log2n, desc = WORKLOADS[args.workload]
    if args.log2_points is not None:
        log2n = args.log2_points
    g = golden_io.load_golden(table_name)

    if args.impl == "reference":
        run_reference_arm(args, args.workload, g, desc)
        return

    import torch
    import torch.distributed as dist
    from adaptive_interpolation_b200 import sharding

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d" % (args.gpus, world))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device; there is no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    # NCCL may print to stdout (e.g. its version banner under NCCL_DEBUG=VERSION); the contract is
    # ONE JSON line there, so everything before that line goes to stderr at the descriptor level
    sys.stdout.flush()
    stdout_fd = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    flat = tables.flatten(g)
    tab = tables.DeviceTable(flat, local_rank)
    info = tab.info()
    n_local = 1 << log2n
    n_global = n_local * world
    lo_i, hi_i = sharding.shard_of(n_global, rank, world)
    assert hi_i - lo_i == n_local
    tdt = torch.float32 if flat.dtype == np.float32 else torch.float64
    esz = 4 if flat.dtype == np.float32 else 8

    # this rank's shard of the global synthetic array, generated on its own GPU
    gen = torch.Generator(device=dev)
    gen.manual_seed(1234 + rank)
    lb, ub = float(flat.breaks[0]), float(flat.breaks[-1])
    if args.order == "random":
        x = torch.rand(n_local, dtype=tdt, device=dev, generator=gen) * (ub - lb) + lb
    else:
        x = torch.linspace(0.0, 1.0, n_local + 1, dtype=torch.float64, device=dev)[:-1]
        x = ((x + rank) / world * (ub - lb) + lb).to(tdt)
    x = torch.clamp(x, max=float(np.nextafter(flat.dtype.type(ub), flat.dtype.type(lb)))).contiguous()
    y = torch.empty_like(x)
    stream = torch.cuda.current_stream(dev)

    def step():
        tab.eval_ptr(x.data_ptr(), y.data_ptr(), n_local, stream.cuda_stream)

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local_rank])
        torch.cuda.synchronize(dev)

    for _ in range(args.warmup):
        step()
    barrier()
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    stops = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clocks:
        e0.record(stream)
        for k in range(args.steps):
            starts[k].record(stream)
            step()
            stops[k].record(stream)
        e1.record(stream)
        barrier()
    total_ms = sharding.max_over_ranks(e0.elapsed_time(e1), dist if world > 1 else None, dev)
    kernel_ms = float(np.mean([a.elapsed_time(b) for a, b in zip(starts, stops)]))
    kernel_ms_min = float(np.min([a.elapsed_time(b) for a, b in zip(starts, stops)]))
    value = n_global * args.steps / (total_ms * 1e-3) / 1e9

    # a sample of the timed output, checked against the oracle inside the cpu_baseline leg below
    sl = slice(0, 1 << 14)
    xs_np, ys_np = x[sl].cpu().numpy(), y[sl].cpu().numpy()
    dev_units = None

    # ---- e2e: host buffers through ai_eval_host_* (pinned), copies inside the timed region
    e2e = None
    if not args.no_e2e:
        import ctypes
        from adaptive_interpolation_b200 import _cabi
        nbytes = n_local * esz
        hp = ctypes.c_void_p()
        # pin the host buffers from a CPU next to this rank's GPU (first-touch NUMA placement of the
        # pinned pages); best effort -- a container cpuset that excludes those CPUs keeps what it has
        affinity_before = os.sched_getaffinity(0) if hasattr(os, "sched_getaffinity") else None
        numa_note = "unchanged"
        try:
            import pynvml
            pynvml.nvmlInit()
            pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local_rank))
            numa_note = "thread bound to %d GPU-local CPUs while pinning" % len(os.sched_getaffinity(0))
        except Exception as e:                                          # pragma: no cover
            numa_note = "unchanged (%s)" % type(e).__name__
        _cabi.check(_cabi.lib().ai_host_alloc(ctypes.byref(hp), 2 * nbytes))
        try:
            hbuf = np.ctypeslib.as_array(ctypes.cast(hp, ctypes.POINTER(ctypes.c_byte)), (2 * nbytes,)).view(flat.dtype)
            hx, hy = hbuf[:n_local], hbuf[n_local:]
            hx[:] = x.cpu().numpy()
            hy[:] = 0                                                  # touch y from the same CPUs
            if affinity_before is not None:
                os.sched_setaffinity(0, affinity_before)
            e2e_steps = max(3, min(args.steps, 10))
            tab.eval_host(hx, hy)                                      # warm-up: allocates the pipeline
            tab.eval_host(hx, hy)
            barrier()
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                tab.eval_host(hx, hy)                                  # synchronous: y complete on return
            dt = time.perf_counter() - t0
            barrier()
            dt = sharding.max_over_ranks(dt, dist if world > 1 else None, dev)
            assert np.array_equal(hy[:4096], y[:4096].cpu().numpy())
            e2e = {"value": n_global * e2e_steps / dt / 1e9, "unit": UNIT,
                   "h2d_bytes_per_step": int(nbytes) * world, "d2h_bytes_per_step": int(nbytes) * world,
                   "steps": e2e_steps, "ms_per_step": dt / e2e_steps * 1e3,
                   "api": "ai_eval_host_%s (pinned host x/y, 32 MiB chunks, 3-slot H2D/kernel/D2H pipeline)"
                          % ("f32" if esz == 4 else "f64"),
                   "host_numa": numa_note}
        finally:
            if affinity_before is not None:
                os.sched_setaffinity(0, affinity_before)
            _cabi.check(_cabi.lib().ai_host_free(hp))

    peaks, peak_kind = measured_peaks()
    algo_bytes = 2.0 * esz * n_local
    achieved = algo_bytes / (kernel_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                "frac": achieved / peaks["hbm_gbs"], "traffic": ncu_traffic(args.workload, n_local),
                "algorithmic_bytes_per_launch": algo_bytes,
                "peak_source": "%s (MEASURED_PEAKS.json hbm_gbs, burst copy)" % peak_kind,
                "kernel": "ai::eval_kernel<%s,%d,%s>" % ({"chebyshev": "kCheb", "legendre": "kLeg", "monomial": "kMono"}[flat.basis],
                                                         flat.order, "float" if esz == 4 else "double"),
                "algorithmic_bytes_per_point": 2 * esz, "points_per_launch": n_local,
                "kernel_ms_avg": kernel_ms, "kernel_ms_min": kernel_ms_min,
                "flops_per_point": 8 + 4 * flat.order}
    try:
        import ctypes
        from adaptive_interpolation_b200 import _cabi
        tf = ctypes.c_double(0)
        _cabi.check(_cabi.lib().ai_measure_fma_peak(local_rank, 0 if esz == 4 else 1, ctypes.byref(tf)))
        roofline["fp_peak_tflops_measured"] = tf.value
        roofline["fp_frac"] = (8 + 4 * flat.order) * n_local / (kernel_ms * 1e-3) / 1e12 / tf.value
    except Exception as e:                                              # pragma: no cover
        roofline["fp_peak_error"] = str(e)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        # the only leg that touches oracle/: the reference's CPU kernel timed beside the GPU path,
        # and the oracle used as the checker of a sample of the timed GPU output
        cpu = time_cpu_baseline(g)
        from oracle import oracle_np
        y_o = oracle_np.evaluate(flat.breaks, flat.coef, flat.basis, flat.order, xs_np)
        cond = oracle_np.error_scale(flat.breaks, flat.coef, flat.basis, flat.order, xs_np)
        dev_units = float(np.max(np.abs(ys_np.astype(np.float64) - y_o.astype(np.float64))
                                 / (np.finfo(flat.dtype).eps * np.maximum(cond, 1e-300))))
        if dev_units > 4:
            raise SystemExit("bench output deviates from the oracle by %.1f units" % dev_units)

    if world > 1:
        dist.barrier(device_ids=[local_rank])
        dist.destroy_process_group()
    sys.stdout.flush()
    os.dup2(stdout_fd, 1)
    os.close(stdout_fd)
    if rank != 0:
        return
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32" if esz == 4 else "f64", "data": "synthetic",
        "config": {"workload": args.workload, "description": desc, "table": table_name,
                   "basis": flat.basis, "order": flat.order, "leaves": flat.n_leaves,
                   "points_per_gpu": n_local, "global_points": n_global, "point_order": args.order,
                   "sharding": "contiguous shards, table replicated, no collective",
                   "l2": "inputs (%.1f GiB x + %.1f GiB y per GPU) far exceed the 126 MB L2; no flush needed"
                         % (n_local * esz / 2 ** 30, n_local * esz / 2 ** 30),
                   "grid_cells": info["grid_cells"], "search_steps": info["search_steps"],
                   "table_image_bytes": info["image_bytes"], "blocks_per_sm": info["blocks_per_sm"]},
        "e2e": e2e, "gpu_launches": args.steps * world, "roofline": roofline, "cpu_baseline": cpu,
        "clocks": clocks.summary(), "max_deviation_units": dev_units,
    }
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
