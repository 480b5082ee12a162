#!/usr/bin/env python
"""End-to-end rate of generate.run_multi (generate.py:493-502, the reference-facing multi-GPU call)
from ONE process: a pinned host array sharded over all visible GPUs, one worker thread per device
through the library's chunked H2D / kernel / D2H pipeline.  Compare with `e2e` of the torchrun
contract line at the same N (one process per GPU).

    python tools/run_multi_e2e.py [--log2n-per-gpu 28] [--reps 3] [--tag run_multi]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import golden_io  # noqa: E402

import numpy as np  # noqa: E402
import torch  # noqa: E402

from adaptive_interpolation_b200 import approximator as app, generate  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log2n-per-gpu", type=int, default=28)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--tag", default="run_multi")
    args = ap.parse_args()
    ndev = torch.cuda.device_count()
    g = golden_io.load_golden("approx__32o6-p1.19209289551e-06")
    approx = app.Approximator.from_record(g)
    out = {"devices": ndev, "rows": []}
    for devices in sorted({1, 2, 4, ndev} & set(range(1, ndev + 1))):
        n = devices << args.log2n_per_gpu
        x = generate.pinned_empty(n, np.float32)
        y_pinned = generate.pinned_empty(n, np.float32)
        rng = np.random.default_rng(3)
        step = 1 << 24
        for lo in range(0, n, step):
            x[lo:lo + step] = rng.uniform(0, 20, min(step, n - lo)).astype(np.float32)
        np.minimum(x, np.nextafter(np.float32(20), np.float32(0)), out=x)
        devs = list(range(devices))
        generate.run_multi(x[:1 << 20], approx, devices=devs)                    # tables + pipelines
        generate.run_multi(x, approx, devices=devs, timing="wall", out=y_pinned)   # warm-up at full size
        best, best_pageable = float("inf"), float("inf")
        for _ in range(args.reps):
            sec, y = generate.run_multi(x, approx, devices=devs, timing="wall", out=y_pinned)
            best = min(best, sec)
        for _ in range(2):                                                          # default: a fresh pageable y per call
            sec, y = generate.run_multi(x, approx, devices=devs, timing="wall")
            best_pageable = min(best_pageable, sec)
        kernel, _ = generate.run_multi(x, approx, devices=devs, out=y_pinned)
        row = {"gpus": devices, "points": n, "wall_s_best": best, "e2e_gpts": n / best / 1e9,
               "e2e_gpts_pageable_out": n / best_pageable / 1e9,
               "gbs_per_direction": 4.0 * n / best / 1e9, "kernel_s_max_over_devices": kernel}
        out["rows"].append(row)
        print("run_multi  %d GPU(s)  %5.2f Gi points  wall %7.1f ms  %6.2f Gpt/s end to end, pinned x and y  (%5.1f GB/s per direction; "
              "kernels %.1f ms; with a fresh pageable y per call %.2f Gpt/s)"
              % (devices, n / 2 ** 30, best * 1e3, row["e2e_gpts"], row["gbs_per_direction"], kernel * 1e3,
                 row["e2e_gpts_pageable_out"]), flush=True)
        del x, y, y_pinned
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "%s.json" % args.tag), "w") as fh:
        json.dump(out, fh, indent=1)


if __name__ == "__main__":
    main()
