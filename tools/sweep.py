#!/usr/bin/env python
"""Kernel-time sweep on one GPU: copy ceilings, then every golden table (or a subset) with
random and sorted points.  Writes gpurun_out/sweep.json and prints a table.  Measurement aid
for profiles/; the contract line is bench.py's.

    python tools/sweep.py [--log2n 28] [--reps 20] [--tables all|bench|name,name] [--tag r1]
"""
import argparse
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import golden_io  # noqa: E402  (fixture loader, tests/golden/golden_io.py)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from adaptive_interpolation_b200 import _cabi, tables  # noqa: E402

BENCH_TABLES = ["approx__32o6-p1.19209289551e-06", "approx__32o3-p1.19209289551e-06",
                "approx__32o7-p1.19209289551e-06", "approx__32o13-p1.19209289551e-06",
                "approx__64o3-p1.19209289551e-06", "approx__64o6-p1.19209289551e-06",
                "approx__64o7-p2.22044604925e-14", "gen__cfg3_j0_cheb_o13_f64",
                "approximations__64o5-p2.22044604925e-14", "approximations__64o6-p2.22044604925e-14",
                "gen__cfg1_sinsin_cheb_o3_f64", "gen__cfg4_f1_leg_o7_f32", "gen__cfg4_f1_leg_o7_f64",
                "gen__j0_leg_o7_f32", "gen__j0_mono_o13_f64", "gen__j0_cheb_o10_f32"]


def time_ms(fn, reps, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for a, b in ev:
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    t = [a.elapsed_time(b) for a, b in ev]
    return float(np.median(t)), float(np.min(t))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log2n", type=int, default=28)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--tables", default="bench")
    ap.add_argument("--tag", default="sweep")
    ap.add_argument("--clustered", default="cfg4", choices=["cfg4", "all", "none"],
                    help="also time points clustered at the breakpoints: for the config-4 tables (default), all, none")
    args = ap.parse_args()
    n = 1 << args.log2n
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    hbm = json.load(open(peaks_path))["hbm_gbs"] if os.path.exists(peaks_path) else 6650.0
    out = {"n": n, "hbm_peak_gbs": hbm, "rows": []}
    s = torch.cuda.current_stream().cuda_stream
    L = _cabi.lib()

    a = torch.rand(n, dtype=torch.float32, device="cuda")
    b = torch.empty_like(a)
    med, mn = time_ms(lambda: b.copy_(a), args.reps)
    out["torch_copy_gbs"] = 8.0 * n / (mn * 1e-3) / 1e9
    med, mn = time_ms(lambda: _cabi.check(L.ai_copy_f32(a.data_ptr(), b.data_ptr(), n, 0, s)), args.reps)
    out["ai_copy_gbs"] = 8.0 * n / (mn * 1e-3) / 1e9
    for dt in (0, 1):
        tf = ctypes.c_double(0)
        _cabi.check(L.ai_measure_fma_peak(0, dt, ctypes.byref(tf)))
        out["fma_peak_tflops_%s" % ("f32" if dt == 0 else "f64")] = tf.value
    print("copy ceilings: torch %.0f GB/s, ai_copy %.0f GB/s (MEASURED_PEAKS %.0f); FMA peak f32 %.1f f64 %.1f TFLOP/s"
          % (out["torch_copy_gbs"], out["ai_copy_gbs"], hbm, out["fma_peak_tflops_f32"], out["fma_peak_tflops_f64"]))
    del a, b

    names = golden_io.golden_names() if args.tables == "all" else (BENCH_TABLES if args.tables == "bench" else args.tables.split(","))
    print("%-42s %-9s %3s %-4s %5s %2s %7s | %9s %8s %6s | %9s %8s %6s" % (
        "table", "basis", "n", "T", "L", "K", "smemB", "rand ms", "Gpt/s", "%hbm", "sort ms", "Gpt/s", "%hbm"))
    for name in names:
        if name.startswith("synth:"):
            # synth:<leaves>:<order>:<f32|f64>  -- uniform leaves on [0,16], Chebyshev fit of sin(3x)
            from numpy.polynomial import chebyshev as C
            _, nl, order, tname = name.split(":")
            nl, order, dt = int(nl), int(order), (np.float32 if tname == "f32" else np.float64)
            br = np.linspace(0.0, 16.0, nl + 1).astype(dt)
            nodes = np.cos(np.pi * (np.arange(order + 1) + 0.5) / (order + 1))
            mids, half = (br[:-1].astype(np.float64) + br[1:]) / 2, (br[1:].astype(np.float64) - br[:-1]) / 2
            vals = np.sin(3 * (mids[:, None] + half[:, None] * nodes[None, :]))
            co = np.stack([C.chebfit(nodes, v, order) for v in vals]).astype(dt)
            f = tables.FlatTable("chebyshev", order, dt, br, co)
        elif name.startswith("bisect:"):
            # bisect:<leaves>:<order>:<f32|f64>  -- a random bisection tree on [0,16] (depth <= 14: dyadic),
            # the shape the reference's adaptive fitter produces; random coefficients
            _, nl, order, tname = name.split(":")
            nl, order, dt = int(nl), int(order), (np.float32 if tname == "f32" else np.float64)
            rng = np.random.default_rng(nl)
            leaves = [(0, 0)]
            while len(leaves) < nl:
                j = int(rng.integers(0, len(leaves)))
                i, d = leaves[j]
                if d < 14:
                    leaves[j:j + 1] = [(2 * i, d + 1), (2 * i + 1, d + 1)]
            edges = sorted({i / 2.0 ** d for i, d in leaves} | {1.0})
            br = (16.0 * np.array(edges)).astype(dt)
            f = tables.FlatTable("chebyshev", order, dt, br, rng.standard_normal((len(br) - 1, order + 1)).astype(dt))
        else:
            g = golden_io.load_golden(name)
            f = tables.flatten(g)
        tab = tables.DeviceTable(f, 0)
        info = tab.info()
        tdt = torch.float32 if f.dtype == np.float32 else torch.float64
        esz = 4 if f.dtype == np.float32 else 8
        lb, ub = float(f.breaks[0]), float(f.breaks[-1])
        gen = torch.Generator(device="cuda")
        gen.manual_seed(7)
        x = torch.rand(n, dtype=tdt, device="cuda", generator=gen) * (ub - lb) + lb
        x = torch.clamp(x, max=float(np.nextafter(f.dtype.type(ub), f.dtype.type(lb))))
        y = torch.empty_like(x)
        row = {"table": name, "basis": f.basis, "order": f.order, "dtype": f.dtype.name, "leaves": f.n_leaves,
               "search_steps": info["search_steps"], "image_bytes": info["image_bytes"],
               "blocks_per_sm": info["blocks_per_sm"], "grid_cells": info["grid_cells"],
               "residency": info["residency"], "lookup_mode": info["lookup_mode"], "smem_bytes": info["smem_bytes"]}
        orders = ("random", "sorted", "clustered") if (args.clustered == "all" or name.startswith("gen__cfg4")) \
            and args.clustered != "none" else ("random", "sorted")
        for order in orders:
            if order == "sorted":
                x = torch.sort(x)[0]
            elif order == "clustered":
                # BASELINE configs[3]'s own input distribution (bench.fill_points, the device port of
                # tests/golden/make_golden.py:make_points(clustered=True))
                import bench
                bench.fill_points(torch, x, f, "clustered", gen)
            med, mn = time_ms(lambda: tab.eval_ptr(x.data_ptr(), y.data_ptr(), n, s), args.reps)
            row[order] = {"ms_median": med, "ms_min": mn, "gpts": n / (med * 1e-3) / 1e9,
                          "gbs": 2.0 * esz * n / (med * 1e-3) / 1e9, "frac_hbm": 2.0 * esz * n / (med * 1e-3) / 1e9 / hbm,
                          "tflops_ref_model": (8 + 4 * f.order) * n / (med * 1e-3) / 1e12}
        out["rows"].append(row)
        r, so = row["random"], row["sorted"]
        cl = row.get("clustered")
        print("%-42s %-9s %3d %-4s %5d %2d %7d | %9.3f %8.1f %6.1f | %9.3f %8.1f %6.1f%s" % (
            name[:42], f.basis, f.order, "f32" if esz == 4 else "f64", f.n_leaves, info["search_steps"],
            info["image_bytes"], r["ms_median"], r["gpts"], 100 * r["frac_hbm"], so["ms_median"], so["gpts"],
            100 * so["frac_hbm"],
            (" | clustered %9.3f %8.1f %6.1f" % (cl["ms_median"], cl["gpts"], 100 * cl["frac_hbm"])) if cl else ""))
        del x, y, tab
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "%s.json" % args.tag), "w") as fh:
        json.dump(out, fh, indent=1)


if __name__ == "__main__":
    main()
