#!/usr/bin/env python
"""Evaluation kernel time when x and y cannot be 16-byte aligned together (x is a view that starts
one element into its buffer, y starts on a boundary): the whole call takes the kernel's scalar loop.

    python tools/misaligned_bench.py [--log2n 28]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import golden_io  # noqa: E402

import numpy as np  # noqa: E402
import torch  # noqa: E402

from adaptive_interpolation_b200 import tables  # noqa: E402
sys.path.insert(0, os.path.join(ROOT, "tools"))
from sum_bench import time_ms  # noqa: E402

TABLES = ["approx__32o6-p1.19209289551e-06", "approx__32o3-p1.19209289551e-06", "approx__64o6-p1.19209289551e-06",
          "gen__cfg3_j0_cheb_o13_f64", "gen__cfg4_f1_leg_o7_f32", "approximations__64o5-p2.22044604925e-14"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log2n", type=int, default=28)
    args = ap.parse_args()
    n = 1 << args.log2n
    s = torch.cuda.current_stream().cuda_stream
    print("%-44s %4s | %10s %8s | %14s %8s" % ("table", "T", "aligned ms", "Gpt/s", "x off by 1: ms", "Gpt/s"))
    for name in TABLES:
        f = tables.flatten(golden_io.load_golden(name))
        tab = tables.DeviceTable(f, 0)
        tdt = torch.float32 if f.dtype == np.float32 else torch.float64
        lo, hi = float(f.breaks[0]), float(f.breaks[-1])
        buf = (torch.rand(n + 4, device="cuda", dtype=tdt) * (hi - lo) + lo).clamp_(max=float(np.nextafter(f.dtype.type(hi), f.dtype.type(lo))))
        y = torch.empty(n, device="cuda", dtype=tdt)
        esz = f.dtype.itemsize
        ta = time_ms(lambda: tab.eval_ptr(buf.data_ptr(), y.data_ptr(), n, s))
        ya = y.clone()
        tm = time_ms(lambda: tab.eval_ptr(buf.data_ptr() + esz, y.data_ptr(), n, s))
        tab.eval_ptr(buf.data_ptr() + esz, ya.data_ptr() + 0, n, s)          # same points through the scalar loop
        tab.eval_ptr(buf.data_ptr() + esz, y.data_ptr(), n, s)
        torch.cuda.synchronize()
        assert torch.equal(ya.view(torch.int32 if esz == 4 else torch.int64), y.view(torch.int32 if esz == 4 else torch.int64))
        print("%-44s %4s | %10.3f %8.1f | %14.3f %8.1f" % (name[:44], "f32" if esz == 4 else "f64", ta, n / ta / 1e6, tm, n / tm / 1e6))
        tab.close()


if __name__ == "__main__":
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    main()
