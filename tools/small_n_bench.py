#!/usr/bin/env python
"""Kernel time at small n (BASELINE configs[0]: 2^20 points): launch + table staging + work."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import golden_io  # noqa: E402

import numpy as np  # noqa: E402
import torch  # noqa: E402

from adaptive_interpolation_b200 import tables  # noqa: E402
from sum_bench import time_ms  # noqa: E402


def main():
    s = torch.cuda.current_stream().cuda_stream
    names = sys.argv[1:] or ["gen__cfg1_sinsin_cheb_o3_f64", "approx__32o6-p1.19209289551e-06", "approx__64o7-p2.22044604925e-14"]
    for name in names:
        f = tables.flatten(golden_io.load_golden(name))
        tab = tables.DeviceTable(f, 0)
        info = tab.info()
        tdt = torch.float32 if f.dtype == np.float32 else torch.float64
        lo, hi = float(f.breaks[0]), float(f.breaks[-1])
        for log2n in (10, 14, 16, 18, 20, 22, 24):
            n = 1 << log2n
            x = torch.linspace(lo, hi, n + 1, device="cuda", dtype=tdt)[:n].contiguous()
            y = torch.empty_like(x)
            t = time_ms(lambda: tab.eval_ptr(x.data_ptr(), y.data_ptr(), n, s), reps=50, warm=10)
            xr = x[torch.randperm(n, device="cuda")].contiguous()
            tr = time_ms(lambda: tab.eval_ptr(xr.data_ptr(), y.data_ptr(), n, s), reps=50, warm=10)
            # the same launch 64 times in ONE CUDA graph: GPU time per launch without the host's enqueue gaps
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            g = torch.cuda.CUDAGraph()
            with torch.cuda.stream(side):
                with torch.cuda.graph(g, stream=side):
                    for _ in range(64):
                        tab.eval_ptr(xr.data_ptr(), y.data_ptr(), n, torch.cuda.current_stream().cuda_stream)
            torch.cuda.current_stream().wait_stream(side)
            tg = time_ms(lambda: g.replay(), reps=10, warm=3) / 64
            print("%-40s residency %2d smem %6d | n 2^%-2d | sorted %8.2f us | random %8.2f us | random, in a CUDA graph %7.2f us"
                  % (name[:40], info["residency"], info["smem_bytes"], log2n, 1e3 * t, 1e3 * tr, 1e3 * tg), flush=True)
        tab.close()


if __name__ == "__main__":
    main()
