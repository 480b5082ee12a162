#!/usr/bin/env python
"""The reference-facing call on HOST arrays (ai_eval_host_*): pinned buffers against ordinary (pageable)
numpy arrays -- what a caller of the reference's run_test.py-style binding actually passes.

    python tools/host_pageable_bench.py [--log2n 28]
"""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import golden_io  # noqa: E402

import numpy as np  # noqa: E402

from adaptive_interpolation_b200 import generate, tables  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log2n", type=int, default=28)
    ap.add_argument("--table", default="approx__32o6-p1.19209289551e-06")
    args = ap.parse_args()
    n = 1 << args.log2n
    f = tables.flatten(golden_io.load_golden(args.table))
    tab = tables.DeviceTable(f, 0)
    rng = np.random.default_rng(0)
    x = rng.uniform(float(f.breaks[0]), float(f.breaks[-1]), n).astype(f.dtype)
    x = np.minimum(x, np.nextafter(f.dtype.type(f.breaks[-1]), f.dtype.type(f.breaks[0])))
    y = np.empty_like(x)
    y[:] = 0                                         # touch the pages (first-touch faults are not the library's)
    xp = generate.pinned_empty(n, f.dtype)
    yp = generate.pinned_empty(n, f.dtype)
    xp[:] = x
    for label, xa, ya in (("pinned", xp, yp), ("pageable numpy", x, y)):
        tab.eval_host(xa, ya)                        # warm-up (pipeline buffers)
        best = 1e9
        for _ in range(3):
            t0 = time.perf_counter()
            tab.eval_host(xa, ya)
            best = min(best, time.perf_counter() - t0)
        print("%-16s %8.1f ms  %6.2f Gpt/s  %6.1f GB/s per direction" % (label, best * 1e3, n / best / 1e9,
                                                                      n * f.dtype.itemsize / best / 1e9), flush=True)
    assert np.array_equal(y, np.asarray(yp))
    tab.close()
    # the reference-facing Python call on a numpy array (generate.run -> the same pipeline, fresh y per call)
    # against what it did before: whole-array pageable copies through torch around one launch
    import torch
    g = golden_io.load_golden(args.table)
    generate.run(x[:1024], g)
    best = 1e9
    for _ in range(3):
        t0 = time.perf_counter()
        yr = generate.run(x, g)
        best = min(best, time.perf_counter() - t0)
    print("%-44s %8.1f ms  %6.2f Gpt/s" % ("generate.run(numpy x) -> numpy y", best * 1e3, n / best / 1e9), flush=True)
    assert np.array_equal(yr, y)
    dt = generate.device_table(g, 0)
    best = 1e9
    for _ in range(3):
        t0 = time.perf_counter()
        xd = torch.from_numpy(x).to("cuda")
        yd = torch.empty_like(xd)
        dt.eval_ptr(xd.data_ptr(), yd.data_ptr(), n, torch.cuda.current_stream().cuda_stream)
        yo = yd.cpu().numpy()
        best = min(best, time.perf_counter() - t0)
        del xd, yd
    print("%-44s %8.1f ms  %6.2f Gpt/s" % ("torch pageable copy + launch + copy back", best * 1e3, n / best / 1e9), flush=True)
    assert np.array_equal(yo, y)


if __name__ == "__main__":
    main()
