#!/usr/bin/env python
"""bench.py -- the hot path's headline metric on B200: interpolant evaluations per second.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload NAME]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the evaluator over one batch of synthetic points: the table stays
resident, x (already in HBM) streams in, y streams out.  Headline workload = BASELINE.json
configs[1]: Bessel J0 on [0,20], Chebyshev order 6, fp32 table (approx/32o6-p1.19e-06), 2^28
uniform random points per GPU.  Each rank owns its own 2^28-point contiguous shard of the global
array (weak scaling, no data-path collective); time = max over ranks of the device time.

Prints ONE JSON line (rank 0).  Keys beyond the base contract:
  e2e          same metric through the reference-facing C-ABI call on HOST buffers
               (ai_eval_host_*: pinned host memory, H2D + kernel + D2H inside the timed region),
               with the plain-copy ceiling of the same bytes through the same pipeline beside it
  roofline     algorithmic bytes (2*sizeof(T) per point) / kernel time vs the measured HBM peak;
               roofline.sustained = the same launch repeated for >= 2 s with its clock record
  configs      every BASELINE.json config, as specified: cfg1 (2^20 linspace points, with the
               reference's Approximator.evaluate timed beside it), cfg2 (= the headline), cfg3
               (2^30 GLOBAL points split over the N GPUs: strong scaling), cfg4 (points clustered
               at the breakpoints, fp32 + fp64), cfg5 (every saved table at 2^30 global points:
               min / median / max roofline fraction)
  cpu_baseline the reference's own generated kernel (oracle/_ref) or the C port, on host cores,
               and b_ref: the unmodified reference's Approximator.evaluate on 1 core
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

# name -> (golden table, log2 points per GPU, point distribution, description)
WORKLOADS = {
    "cfg2_j0_cheb_o6_f32": ("approx__32o6-p1.19209289551e-06", 28, "uniform",
                            "BASELINE configs[1]: J0 [0,20] Chebyshev o6 fp32 (1.19e-6 table), 2^28 uniform random points per GPU"),
    "cfg3_j0_cheb_o13_f64": ("gen__cfg3_j0_cheb_o13_f64", 28, "uniform",
                             "BASELINE configs[2] table: J0 Chebyshev o13 fp64 (2.2e-14 table), 2^28 points per GPU"),
    "cfg1_sinsin_cheb_o3_f64": ("gen__cfg1_sinsin_cheb_o3_f64", 28, "uniform", "configs[0] table at GPU size"),
    "cfg4_f1_leg_o7_f32": ("gen__cfg4_f1_leg_o7_f32", 28, "clustered",
                           "BASELINE configs[3]: deep skewed Legendre o7 tree, fp32, points clustered at the breakpoints"),
    "cfg4_f1_leg_o7_f64": ("gen__cfg4_f1_leg_o7_f64", 28, "clustered",
                           "BASELINE configs[3]: deep skewed Legendre o7 tree, fp64, points clustered at the breakpoints"),
    "j0_cheb_o7_f64": ("approx__64o7-p2.22044604925e-14", 28, "uniform", "J0 Chebyshev o7 fp64 2.2e-14 (64 leaves)"),
    "j0_cheb_o5_f64_429": ("approximations__64o5-p2.22044604925e-14", 28, "uniform", "J0 Chebyshev o5 fp64 2.2e-14 (429 leaves)"),
    "large_j0_cheb_o3_f64": ("large__approx__64o3-p2.22044604925e-14", 28, "uniform",
                             "refit of the missing approx/64o3-p2.22e-14 (6872 leaves, L2-resident)"),
}
DEFAULT_WORKLOAD = "cfg2_j0_cheb_o6_f32"
CPU_SAMPLE_LOG2 = 26
CFG3_GLOBAL_LOG2 = 30           # BASELINE configs[2]: 2^30 points sharded over the GPUs
CFG5_GLOBAL_LOG2 = 30           # BASELINE configs[4]: every saved table at 2^30 points
SUSTAINED_SECONDS = 2.0
CONFIG_IDLE_SECONDS = 0.25      # idle gap before each per-config measurement (the power window recovers)

NCU_SUMMARIES = {   # workload -> committed `ncu --set full` summary (tools/ncu_summary.py)
    "cfg2_j0_cheb_o6_f32": "r3_ncu_cfg2.json",
    "cfg3_j0_cheb_o13_f64": "r3_ncu_cfg3.json",
}


def ncu_traffic(workload, points_per_launch):
    """(bytes, source): dram__bytes_read.sum + dram__bytes_write.sum per launch of the eval kernel
    from the committed ncu capture of the same workload and launch size (profiles/); (None, why)
    when there is none.  NOT measured in this run -- the source says which capture it is."""
    name = NCU_SUMMARIES.get(workload)
    path = os.path.join(ROOT, "profiles", name) if name else None
    if not path or not os.path.exists(path):
        return None, "no committed ncu capture for this workload"
    with open(path) as fh:
        d = json.load(fh)
    if d.get("points_per_launch") != points_per_launch or not d.get("launches"):
        return None, "committed capture profiles/%s is for another launch size" % name
    return (d["launches"][0]["derived"]["dram_traffic_bytes"],
            "profiles/%s: one `ncu --set full` capture of this kernel at this launch size, committed; "
            "not re-measured in this run" % name)


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            return json.load(fh), "measured"
    return {"hbm_gbs": 6650.0}, "fallback"


class ClockSampler(object):
    """SM clock and throttle reasons sampled while the timed regions run (NVML).  mark() returns
    the summary of the samples since the previous mark, so every config gets its own record."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], [], None
        self._stop = threading.Event()
        self._thread = None
        self._mark = 0
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
                mhz = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.samples.append(mhz)
                self.reasons.append(tuple(k for k, bit in names.items() if mask & bit))
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

    def _summary(self, lo, hi):
        s, r = self.samples[lo:hi], self.reasons[lo:hi]
        if not s:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(s)), "sm_mhz_min": float(np.min(s)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted({k for t in r for k in t}), "samples": len(s)}

    def mark(self):
        lo, self._mark = self._mark, len(self.samples)
        return self._summary(lo, self._mark)

    def summary(self):
        return self._summary(0, len(self.samples))


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


# ------------------------------------------------------------------ CPU legs (the only users of oracle/)
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


def reference_evaluate(golden_rec, x):
    """B-ref: the UNMODIFIED reference's Approximator.evaluate (approximator.py:273-296) from the
    installed copy in baseline/_ref (git-ignored, travels to the GPU box), on one core.  Returns
    (y, seconds), or (None, reason) when the install is not there."""
    ref_root = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.isdir(os.path.join(ref_root, "adaptive_interpolation")):
        return None, "baseline/_ref not installed"
    sys.path.insert(0, ref_root)
    try:
        import contextlib
        import io
        with contextlib.redirect_stdout(io.StringIO()):
            import adaptive_interpolation.adapt as ref_adapt
            import adaptive_interpolation.approximator as ref_app
            fitter = ref_adapt.Interpolant(lambda v: v, int(golden_rec.max_order), 1e-6, golden_rec.basis,
                                           "32" if golden_rec.dtype == np.float32 else "64")
            ap = ref_app.Approximator()
        ap.basis_function = fitter.basis_function
        ap.tree_1d = np.asarray(golden_rec.tree_1d, dtype=golden_rec.dtype)
        ap.max_order, ap.num_levels = int(golden_rec.max_order), int(golden_rec.num_levels)
        ap.basis, ap.optimizations = golden_rec.basis, []
        t0 = time.perf_counter()
        y = np.array(ap.evaluate(x))
        return y, time.perf_counter() - t0
    except Exception as e:                                          # pragma: no cover
        return None, "%s: %s" % (type(e).__name__, e)
    finally:
        sys.path.remove(ref_root)


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


# ------------------------------------------------------------------ synthetic points, generated on the device
def fill_points(torch, x, flat, kind, gen, rank=0, world=1):
    """Fill the device tensor x (table dtype) in place.
      uniform    rng.uniform(lb, ub), clipped below ub  
      sorted     this rank's slice of linspace(lb, ub)  (run_test.py:230-235)
      clustered  BASELINE configs[3] / SURVEY 8(d)4, the device port of
                 tests/golden/make_golden.py:make_points(clustered=True): each point is, with
                 probability 1/2, uniform on the domain, else breaks[k] +- |N(0,1)| * 2^-j with k
                 uniform over the breakpoints and j uniform in [10, 40]; every breakpoint and its
                 two nextafter neighbours are planted at random positions."""
    n = x.numel()
    dev = x.device
    lb, ub = float(flat.breaks[0]), float(flat.breaks[-1])
    below_ub = float(np.nextafter(flat.dtype.type(ub), flat.dtype.type(lb)))
    if kind == "uniform":
        x.uniform_(0.0, 1.0, generator=gen).mul_(ub - lb).add_(lb).clamp_(max=below_ub)
    elif kind == "sorted":
        step = 1 << 24
        for lo in range(0, n, step):
            hi = min(n, lo + step)
            t = torch.arange(lo, hi, dtype=torch.float64, device=dev)
            x[lo:hi] = (((t / n) + rank) / world * (ub - lb) + lb).to(x.dtype)
        x.clamp_(max=below_ub)
    elif kind == "clustered":
        b = torch.from_numpy(flat.breaks.astype(np.float64)).to(dev)
        step = 1 << 24                                   # bounds the fp64 temporaries
        for lo in range(0, n, step):
            m = min(n, lo + step) - lo
            u = torch.rand(m, dtype=torch.float64, device=dev, generator=gen) * (ub - lb) + lb
            k = torch.randint(0, b.numel(), (m,), device=dev, generator=gen)
            j = torch.randint(10, 41, (m,), device=dev, generator=gen)
            sgn = torch.randint(0, 2, (m,), device=dev, generator=gen).double() * 2.0 - 1.0
            mag = torch.randn(m, dtype=torch.float64, device=dev, generator=gen).abs_()
            c = b[k] + sgn * torch.ldexp(mag, -j)
            pick = torch.rand(m, dtype=torch.float32, device=dev, generator=gen) < 0.5
            x[lo:lo + m] = torch.where(pick, u, c).to(x.dtype)
            del u, k, j, sgn, mag, c, pick
        x.clamp_(max=below_ub)                           # make_points: r >= ub -> nextafter(ub, -inf)
        bt = torch.from_numpy(flat.breaks).to(dev)
        inf = torch.tensor(float("inf"), dtype=x.dtype, device=dev)
        special = torch.cat([bt, torch.nextafter(bt, -inf), torch.nextafter(bt, inf)])
        if special.numel() < n:
            pos = torch.randint(0, n, (special.numel(),), device=dev, generator=gen)
            x[pos] = special
    else:
        raise ValueError(kind)
    return x


class Arena(object):
    """Two device byte buffers sized for the largest config of this run, viewed per config in the
    table dtype: every config streams fresh x / y far larger than the 126 MB L2."""

    def __init__(self, torch, dev, nbytes):
        self.torch = torch
        # y starts a few KB off the allocation base so that x and y are not an exact power-of-two
        # distance apart (experiment hook AI_BENCH_Y_OFFSET, bytes, multiple of 256)
        self.yoff = int(os.environ.get("AI_BENCH_Y_OFFSET", "0"))
        self.xb = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        self.yb = torch.empty(nbytes + self.yoff, dtype=torch.uint8, device=dev)

    def views(self, np_dtype, n):
        tdt = self.torch.float32 if np_dtype == np.float32 else self.torch.float64
        esz = 4 if np_dtype == np.float32 else 8
        return self.xb[:n * esz].view(tdt), self.yb[self.yoff:self.yoff + n * esz].view(tdt)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--log2-points", type=int, default=None, help="points per GPU (default: workload's)")
    ap.add_argument("--order", default=None, choices=["random", "sorted", "clustered"],
                    help="point distribution (default: the workload's own: uniform random, or clustered for cfg4)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the per-BASELINE-config records")
    ap.add_argument("--no-sustained", action="store_true")
    args = ap.parse_args()
    args.warmup = max(3, args.warmup)

    import golden_io                               # fixture loader (tests/golden/golden_io.py)
    from adaptive_interpolation_b200 import tables
    table_name, log2n, dist_kind, desc = WORKLOADS[args.workload]
    if args.log2_points is not None:
        log2n = args.log2_points
    if args.order is not None:
        dist_kind = {"random": "uniform"}.get(args.order, args.order)
    g = golden_io.load_golden(table_name)

    if args.impl == "reference":
        run_reference_arm(args, args.workload, g, desc)
        return

    import ctypes
    import torch
    import torch.distributed as dist
    from adaptive_interpolation_b200 import _cabi, sharding

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
    dgroup = dist if world > 1 else None
    full_run = (args.workload == DEFAULT_WORKLOAD and args.log2_points is None and args.order is None)
    do_configs = full_run and not args.no_configs

    peaks, peak_kind = measured_peaks()
    hbm = peaks["hbm_gbs"]
    stream = torch.cuda.current_stream(dev)
    gen = torch.Generator(device=dev)
    gen.manual_seed(1234 + rank)

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local_rank])
        torch.cuda.synchronize(dev)

    def shard_points(n_global):
        lo_i, hi_i = sharding.shard_of(n_global, rank, world)
        return hi_i - lo_i

    n_local = 1 << log2n
    n_global = n_local * world
    assert shard_points(n_global) == n_local
    cfg_local = max(shard_points(1 << CFG3_GLOBAL_LOG2), shard_points(1 << CFG5_GLOBAL_LOG2)) if do_configs else 0
    flat = tables.flatten(g)
    esz = flat.dtype.itemsize
    arena = Arena(torch, dev, max(n_local * esz, cfg_local * 8, (1 << 28) * 8 if do_configs else 0))

    def time_table(tab, x, y, steps, warmup, per_launch=True):
        """warmup untimed launches, then `steps` timed ones bracketed by barrier + synchronize.
        Returns (total_ms max over ranks, mean launch ms, min launch ms)."""
        n = x.numel()
        for _ in range(warmup):
            tab.eval_ptr(x.data_ptr(), y.data_ptr(), n, stream.cuda_stream)
        barrier()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)] \
            if per_launch else []
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for k in range(steps):
            if per_launch:
                ev[k][0].record(stream)
            tab.eval_ptr(x.data_ptr(), y.data_ptr(), n, stream.cuda_stream)
            if per_launch:
                ev[k][1].record(stream)
        e1.record(stream)
        barrier()
        total = sharding.max_over_ranks(e0.elapsed_time(e1), dgroup, dev)
        if per_launch:
            ms = [a.elapsed_time(b) for a, b in ev]
            return total, float(np.mean(ms)), float(np.min(ms))
        return total, e0.elapsed_time(e1) / steps, None

    def frac_of_hbm(points, elem, ms):
        return 2.0 * elem * points / (ms * 1e-3) / 1e9 / hbm

    with ClockSampler(local_rank) as clocks:
        # ================================================================ headline workload
        tab = tables.DeviceTable(flat, local_rank)
        info = tab.info()
        x, y = arena.views(flat.dtype.type, n_local)
        fill_points(torch, x, flat, dist_kind, gen, rank, world)
        clocks.mark()
        total_ms, kernel_ms, kernel_ms_min = time_table(tab, x, y, args.steps, args.warmup)
        headline_clocks = clocks.mark()
        value = n_global * args.steps / (total_ms * 1e-3) / 1e9

        # a sample of the timed output, checked against the oracle inside the cpu_baseline leg below
        sl = slice(0, 1 << 14)
        xs_np, ys_np = x[sl].cpu().numpy(), y[sl].cpu().numpy()
        dev_units = None
        sustained = None

        # ---- e2e: host buffers through ai_eval_host_* (pinned), copies inside the timed region
        e2e = None
        if not args.no_e2e:
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
                dt = sharding.max_over_ranks(dt, dgroup, dev)
                assert np.array_equal(hy[:4096], y[:4096].cpu().numpy())
                # the ceiling: the same bytes through the same pipeline on all ranks at once, no kernel
                L = _cabi.lib()
                _cabi.check(L.ai_measure_host_roundtrip(tab.handle, hx.ctypes.data, hy.ctypes.data, n_local))
                barrier()
                t0 = time.perf_counter()
                for _ in range(e2e_steps):
                    _cabi.check(L.ai_measure_host_roundtrip(tab.handle, hx.ctypes.data, hy.ctypes.data, n_local))
                dt_copy = time.perf_counter() - t0
                barrier()
                dt_copy = sharding.max_over_ranks(dt_copy, dgroup, dev)
                e2e_value = n_global * e2e_steps / dt / 1e9
                ceiling_value = n_global * e2e_steps / dt_copy / 1e9
                e2e = {"value": e2e_value, "unit": UNIT,
                       "h2d_bytes_per_step": int(nbytes) * world, "d2h_bytes_per_step": int(nbytes) * world,
                       "steps": e2e_steps, "ms_per_step": dt / e2e_steps * 1e3,
                       "gbs_per_direction": nbytes * world * e2e_steps / dt / 1e9,
                       "copy_ceiling_gbs": nbytes * world * e2e_steps / dt_copy / 1e9,
                       "copy_ceiling_value": ceiling_value,
                       "frac_of_copy_ceiling": e2e_value / ceiling_value,
                       "copy_ceiling": "ai_measure_host_roundtrip: the same pinned bytes through the same chunks, slots and "
                                       "streams on all %d rank(s) concurrently with NO kernel between H2D and D2H; "
                                       "GB/s per direction, both directions running at once" % world,
                       "api": "ai_eval_host_%s (pinned host x/y, 32 MiB chunks, 3-slot H2D/kernel/D2H pipeline)"
                              % ("f32" if esz == 4 else "f64"),
                       "host_numa": numa_note}
            finally:
                if affinity_before is not None:
                    os.sched_setaffinity(0, affinity_before)
                _cabi.check(_cabi.lib().ai_host_free(hp))

        algo_bytes = 2.0 * esz * n_local
        achieved = algo_bytes / (kernel_ms * 1e-3) / 1e9
        traffic, traffic_source = ncu_traffic(args.workload, n_local)
        roofline = {"bound": "hbm", "achieved": achieved, "peak": hbm, "unit": "GB/s",
                    "frac": achieved / hbm, "traffic": traffic, "traffic_source": traffic_source,
                    "algorithmic_bytes_per_launch": algo_bytes,
                    "peak_source": "%s (MEASURED_PEAKS.json hbm_gbs, burst copy)" % peak_kind,
                    "kernel": "ai::eval_kernel<%s,%d,%s>" % ({"chebyshev": "kCheb", "legendre": "kLeg", "monomial": "kMono"}[flat.basis],
                                                             flat.order, "float" if esz == 4 else "double"),
                    "algorithmic_bytes_per_point": 2 * esz, "points_per_launch": n_local,
                    "kernel_ms_avg": kernel_ms, "kernel_ms_min": kernel_ms_min,
                    "flops_per_point": 8 + 4 * flat.order, "sustained": None}
        fma_peak = {}
        try:
            for dt_id, nm in ((0, "f32"), (1, "f64")):
                tf = ctypes.c_double(0)
                _cabi.check(_cabi.lib().ai_measure_fma_peak(local_rank, dt_id, ctypes.byref(tf)))
                fma_peak[nm] = tf.value
            mine = fma_peak["f32" if esz == 4 else "f64"]
            roofline["fp_peak_tflops_measured"] = mine
            roofline["fp_frac"] = (8 + 4 * flat.order) * n_local / (kernel_ms * 1e-3) / 1e12 / mine
        except Exception as e:                                              # pragma: no cover
            roofline["fp_peak_error"] = str(e)

        # ================================================================ every BASELINE config
        configs = []
        launches_configs = 0
        b_ref = None
        if do_configs:
            # (10 warm-up launches: after the idle gap the first few launches run below full clocks)
            csteps = max(5, min(args.steps, 20))

            def run_config(label, name, n_glob, kind, strong, steps=csteps):
                """One record: table `name`, n_glob GLOBAL points (this rank takes its contiguous shard)."""
                nonlocal launches_configs
                gg = golden_io.load_golden(name)
                ff = tables.flatten(gg)
                tt = tables.DeviceTable(ff, local_rank)
                nl = shard_points(n_glob)
                xx, yy = arena.views(ff.dtype.type, nl)
                fill_points(torch, xx, ff, kind, gen, rank, world)
                torch.cuda.synchronize(dev)
                time.sleep(CONFIG_IDLE_SECONDS)          # every config is a burst like the headline, not the tail of the previous one
                clocks.mark()
                # warm-up after the idle gap: the first launches run below full clocks; ten short launches,
                # three long ones (ten 4 ms launches would already eat into the board's power window)
                cwarm = 3 if nl * ff.dtype.itemsize >= (1 << 32) else 10
                tot, kms, kmin = time_table(tt, xx, yy, steps, cwarm)
                ck = clocks.mark()
                launches_configs += steps
                kms_max = sharding.max_over_ranks(kms, dgroup, dev)
                ti = tt.info()
                f_dt = "f32" if ff.dtype == np.float32 else "f64"
                rec = {"config": label, "workload": name, "points": n_glob, "points_per_gpu": nl,
                       "point_distribution": kind, "scaling": "strong" if strong else "weak",
                       "dtype": f_dt, "basis": ff.basis, "order": ff.order,
                       "leaves": ff.n_leaves, "search_steps": ti["search_steps"], "lookup_mode": ti["lookup_mode"],
                       "residency": ti["residency"],
                       "value": n_glob * steps / (tot * 1e-3) / 1e9, "unit": UNIT, "steps": steps,
                       "kernel_ms": kms_max, "kernel_ms_min": kmin,
                       "frac": frac_of_hbm(nl, ff.dtype.itemsize, kms_max),
                       "fp_frac": ((8 + 4 * ff.order) * nl / (kms_max * 1e-3) / 1e12 / fma_peak[f_dt]) if fma_peak else None,
                       "clocks": ck}
                tt.close()
                return rec

            # cfg1: the reference's own CPU-runnable case -- 2^20 linspace points, fp64 order 3
            n1 = 1 << 20
            g1 = golden_io.load_golden("gen__cfg1_sinsin_cheb_o3_f64")
            f1 = tables.flatten(g1)
            t1 = tables.DeviceTable(f1, local_rank)
            x1 = torch.linspace(0.0, 10.0, n1 + 1, dtype=torch.float64, device=dev)[:-1].contiguous()
            y1 = torch.empty_like(x1)
            clocks.mark()
            tot1, k1, k1min = time_table(t1, x1, y1, 50, 5)
            launches_configs += 50
            # the same launch 64 times inside ONE CUDA graph: GPU time per launch without the host's enqueue
            # gaps (a single launch between two events is bound by the Python / driver enqueue path)
            graph_ms, graph_err = None, ""
            try:
                side = torch.cuda.Stream(device=dev)
                side.wait_stream(stream)
                cg = torch.cuda.CUDAGraph()
                with torch.cuda.stream(side):
                    with torch.cuda.graph(cg, stream=side):
                        for _ in range(64):
                            t1.eval_ptr(x1.data_ptr(), y1.data_ptr(), n1, torch.cuda.current_stream().cuda_stream)
                stream.wait_stream(side)
                with torch.cuda.stream(stream):
                    for _ in range(3):
                        cg.replay()
                    g0, g1e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    g0.record(stream)
                    for _ in range(10):
                        cg.replay()
                    g1e.record(stream)
                torch.cuda.synchronize()
                graph_ms = g0.elapsed_time(g1e) / 640.0
                launches_configs += 13 * 64
                del cg
            except Exception as exc:                      # noqa: BLE001 -- an optional extra record, never fatal
                graph_ms = None
                graph_err = repr(exc)[:200]
            rec1 = {"config": "cfg1", "workload": "gen__cfg1_sinsin_cheb_o3_f64", "points": n1, "points_per_gpu": n1,
                    "point_distribution": "linspace(0, 10, 2^20, endpoint=False) (SURVEY 8(d)1); every rank evaluates the whole "
                                          "array (2^20 points are one launch of ~10 us: not sharded)",
                    "scaling": "replicated", "dtype": "f64", "basis": f1.basis, "order": f1.order, "leaves": f1.n_leaves,
                    "value": n1 / (k1 * 1e-3) / 1e9, "unit": UNIT, "steps": 50, "kernel_ms": k1, "kernel_ms_min": k1min,
                    "frac": frac_of_hbm(n1, 8, k1), "clocks": clocks.mark(),
                    "note": "launch-latency bound at this size (16 MiB of traffic); the same table at 2^28 points is in cfg5"}
            if graph_ms:
                rec1["in_cuda_graph"] = {"launches_per_graph": 64, "kernel_ms": graph_ms, "value": n1 / (graph_ms * 1e-3) / 1e9,
                                         "frac": frac_of_hbm(n1, 8, graph_ms),
                                         "note": "64 launches captured in one CUDA graph, 10 replays: per-launch time without "
                                                 "the host enqueue path"}
            else:
                rec1["in_cuda_graph"] = {"unavailable": graph_err}
            if rank == 0 and world == 1 and not args.no_cpu:
                x1_np = x1.cpu().numpy()
                y_ref, secs = reference_evaluate(g1, x1_np)
                if y_ref is None:
                    b_ref = {"unavailable": secs}
                else:
                    from oracle import oracle_np
                    cond = oracle_np.error_scale(f1.breaks, f1.coef, f1.basis, f1.order, x1_np)
                    units = float(np.max(np.abs(y1.cpu().numpy() - y_ref) / (np.finfo(np.float64).eps * np.maximum(cond, 1e-300))))
                    b_ref = {"value": n1 / secs / 1e9, "unit": UNIT, "cores": 1, "points": n1, "seconds": secs,
                             "impl": "Approximator.evaluate of the unmodified reference installed in baseline/_ref "
                                     "(approximator.py:273-296), BASELINE configs[0] as specified",
                             "gpu_vs_b_ref_max_units": units,
                             "gpu_speedup_kernel_only": secs / (k1 * 1e-3)}
                    if units > 4:
                        raise SystemExit("cfg1: GPU output deviates from the reference's evaluate by %.1f units" % units)
            configs.append(rec1)
            t1.close()
            del x1, y1

            # cfg2 = the headline
            configs.append({"config": "cfg2", "workload": table_name, "points": n_global, "points_per_gpu": n_local,
                            "point_distribution": dist_kind, "scaling": "weak", "dtype": "f32", "basis": flat.basis,
                            "order": flat.order, "leaves": flat.n_leaves, "value": value, "unit": UNIT, "steps": args.steps,
                            "kernel_ms": kernel_ms, "kernel_ms_min": kernel_ms_min, "frac": achieved / hbm,
                            "clocks": headline_clocks, "note": "the headline line above"})
            # the headline table on the reference's OWN input order: a sorted linspace (run_test.py:230-235)
            rec = run_config("cfg2-sorted", table_name, (1 << 28) * world, "sorted", False)
            rec["note"] = "BASELINE configs[1] table, points = this rank's slice of linspace(lb, ub) (leaf-uniform fast path)"
            configs.append(rec)
            # cfg3: 2^30 GLOBAL points over the N GPUs (strong scaling), fp64 order 13
            configs.append(run_config("cfg3", "gen__cfg3_j0_cheb_o13_f64", 1 << CFG3_GLOBAL_LOG2, "uniform", True))
            # cfg4: lookup stress on its own input distribution, both dtypes, 2^28 points per GPU
            for nm in ("gen__cfg4_f1_leg_o7_f32", "gen__cfg4_f1_leg_o7_f64"):
                configs.append(run_config("cfg4", nm, (1 << 28) * world, "clustered", False))
            # cfg5: every saved table (and the refitted missing ones) at 2^30 global points
            # (+ the monomial / Legendre tables the reference's fitter constructs for J0 on [0,20])
            gen5 = ("gen__j0_mono_o5_f64", "gen__j0_mono_o13_f64", "gen__j0_leg_o7_f32", "gen__j0_leg_o7_f64",
                    "gen__j0_leg_o6_remez_f64")
            names5 = [nm for nm in golden_io.golden_names()
                      if nm.startswith(("approx__", "approximations__", "large__")) or nm in gen5]
            rows = []
            for nm in names5:
                rec = run_config("cfg5", nm, 1 << CFG5_GLOBAL_LOG2, "uniform", True, steps=5)
                rows.append({k: rec[k] for k in ("workload", "dtype", "basis", "order", "leaves", "residency", "value",
                                                 "kernel_ms", "frac")})
                rows[-1]["sm_mhz"] = rec["clocks"].get("sm_mhz")
                rows[-1]["reasons"] = rec["clocks"].get("reasons")
            fr = np.array([r["frac"] for r in rows])
            lo7 = np.array([r["frac"] for r in rows if r["order"] <= 7 and r["workload"].startswith(("approx__", "approximations__"))])
            configs.append({"config": "cfg5", "workload": "all %d saved tables (approx/, approximations/) + %d large tables "
                            "refitted with the reference's fitter (.MISSING_LARGE_BLOBS and one more) + %d monomial / Legendre "
                            "tables fitted by the reference's fitter for J0 on [0,20] (SURVEY 8(d)5)"
                            % (sum(r["workload"].startswith(("approx__", "approximations__")) for r in rows),
                               sum(r["workload"].startswith("large__") for r in rows),
                               sum(r["workload"].startswith("gen__") for r in rows)),
                            "points": 1 << CFG5_GLOBAL_LOG2, "points_per_gpu": shard_points(1 << CFG5_GLOBAL_LOG2),
                            "point_distribution": "uniform", "scaling": "strong", "steps": 5,
                            "frac_min": float(fr.min()), "frac_median": float(np.median(fr)), "frac_max": float(fr.max()),
                            "frac_min_saved_order_le7": float(lo7.min()) if lo7.size else None,
                            "value_median": float(np.median([r["value"] for r in rows])), "unit": UNIT,
                            "clocks": {"sm_mhz_min_over_tables": min([r["sm_mhz"] for r in rows if r["sm_mhz"]] or [None]),
                                       "reasons": sorted({k for r in rows for k in (r["reasons"] or [])}),
                                       "sm_max_mhz": clocks.max_mhz},
                            "tables": rows})
        # ---- sustained: the same launch back to back for >= 2 s (north_star "sustains")
        if not args.no_sustained:
            fill_points(torch, x, flat, dist_kind, gen, rank, world)      # the arena was reused by the configs
            batch = 200
            first = last = None
            t_begin = time.perf_counter()
            launches = 0
            clocks.mark()
            # the batch count comes from the max-over-ranks step time, so every rank runs the same
            # number of (barrier-bracketed) batches
            n_batches = max(2, int(np.ceil(SUSTAINED_SECONDS / (total_ms / args.steps * 1e-3 * batch))))
            for _ in range(n_batches):
                _, ms, _ = time_table(tab, x, y, batch, 0, per_launch=False)
                launches += batch
                first = ms if first is None else first
                last = ms
            sc = clocks.mark()
            last = sharding.max_over_ranks(last, dgroup, dev)
            sustained = {"seconds": time.perf_counter() - t_begin, "launches": launches,
                         "kernel_ms_first_batch": first, "kernel_ms_last_batch": last,
                         "achieved": 2.0 * esz * n_local / (last * 1e-3) / 1e9,
                         "frac": frac_of_hbm(n_local, esz, last), "clocks": sc,
                         "note": "back-to-back launches of the headline kernel in batches of %d, no host sync inside a batch; "
                                 "frac is the LAST batch (sw_power_cap here is a note, not a rejection)" % batch}

        roofline["sustained"] = sustained
        all_clocks = clocks.summary()

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        # the only leg that touches oracle/: the reference's CPU kernel timed beside the GPU path,
        # and the oracle used as the checker of a sample of the timed GPU output
        cpu = time_cpu_baseline(g)
        cpu["b_ref"] = b_ref
        from oracle import oracle_np
        # B-np (SURVEY 8(d) CPU baseline ii): the vectorised numpy restatement of Approximator.evaluate,
        # one core, 2^24 uniform random points of the headline workload
        x_np24, _ = make_cpu_sample(g, 24)
        t0 = time.perf_counter()
        oracle_np.evaluate(flat.breaks, flat.coef, flat.basis, flat.order, x_np24)
        secs_np = time.perf_counter() - t0
        cpu["b_np"] = {"value": x_np24.size / secs_np / 1e9, "unit": UNIT, "cores": 1, "points": int(x_np24.size),
                       "seconds": secs_np, "impl": "oracle/oracle_np.py: vectorised numpy restatement of "
                                                   "Approximator.evaluate (approximator.py:273-296), one pass"}
        del x_np24
        y_o = oracle_np.evaluate(flat.breaks, flat.coef, flat.basis, flat.order, xs_np)
        cond = oracle_np.error_scale(flat.breaks, flat.coef, flat.basis, flat.order, xs_np)
        ok = np.isfinite(cond) & np.isfinite(y_o)
        dev_units = float(np.max(np.abs(ys_np[ok].astype(np.float64) - y_o[ok].astype(np.float64))
                                 / (np.finfo(flat.dtype).eps * np.maximum(cond[ok], 1e-300))))
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
    sustained_launches = sustained["launches"] if sustained else 0
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32" if esz == 4 else "f64", "data": "synthetic",
        "config": {"workload": args.workload, "description": desc, "table": table_name,
                   "basis": flat.basis, "order": flat.order, "leaves": flat.n_leaves,
                   "points_per_gpu": n_local, "global_points": n_global, "point_order": dist_kind,
                   "sharding": "contiguous shards, table replicated, no collective",
                   "l2": "inputs (%.1f GiB x + %.1f GiB y per GPU) far exceed the 126 MB L2; no flush needed"
                         % (n_local * esz / 2 ** 30, n_local * esz / 2 ** 30),
                   "grid_cells": info["grid_cells"], "search_steps": info["search_steps"],
                   "table_image_bytes": info["image_bytes"], "blocks_per_sm": info["blocks_per_sm"]},
        "e2e": e2e, "gpu_launches": args.steps * world,
        "gpu_launches_outside_timed_region": {"warmup": args.warmup * world, "sustained": sustained_launches * world,
                                              "configs": launches_configs * world},
        "roofline": roofline, "cpu_baseline": cpu,
        "clocks": headline_clocks if headline_clocks["samples"] else all_clocks, "clocks_whole_run": all_clocks,
        "max_deviation_units": dev_units, "configs": configs,
    }
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
