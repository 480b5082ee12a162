#!/usr/bin/env python
"""Several tables over the same x in one pass (ai_eval_multi_*): kernel
time of the fused launch against m back-to-back single-table launches, random and sorted points.

    python tools/multi_bench.py [--log2n 27] [--reps 15] [--tag multi]

Bytes per point: fused (1 + m) * sizeof(T), separate 2 * m * sizeof(T).
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import golden_io  # noqa: E402  (fixture loader, tests/golden/golden_io.py)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from adaptive_interpolation_b200 import tables  # noqa: E402
from tools.sweep import time_ms  # noqa: E402

GROUPS = {
    "cheb_o6_f32": ["approx__32o6-p1.19209289551e-06", "approximations__32o6-p1.19209289551e-06",
                    "approx__ci32o6-p1.19209289551e-06", "approx__32o6-p1.19209289551e-06"],
    "cheb_o6_f64": ["approx__64o6-p1.19209289551e-06", "approximations__64o6-p1.19209289551e-06",
                    "approx__ci64o6-p1.19209289551e-06", "approx__64o6-p1.19209289551e-06"],
    "cheb_o7_f64": ["approx__64o7-p1.19209289551e-06", "approx__ci64o7-p1.19209289551e-06"],
    "cheb_o13_f32": ["approx__32o13-p1.19209289551e-06", "approx__ci32o13-p1.19209289551e-06"],
    "cheb_o13_f64": ["gen__cfg3_j0_cheb_o13_f64", "gen__cfg3_j0_cheb_o13_f64"],
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log2n", type=int, default=27)
    ap.add_argument("--reps", type=int, default=15)
    ap.add_argument("--tag", default="multi")
    args = ap.parse_args()
    n = 1 << args.log2n
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    hbm = json.load(open(peaks_path))["hbm_gbs"] if os.path.exists(peaks_path) else 6650.0
    s = torch.cuda.current_stream().cuda_stream
    rows = []
    print("%-14s %2s %-7s | %9s %9s %6s | %9s %9s %6s | %5s" % (
        "group", "m", "points", "libry ms", "Gevals/s", "%hbm", "separ ms", "Gevals/s", "%hbm", "gain"))
    for gname, names in GROUPS.items():
        for m in range(2, len(names) + 1):
            tabs = [tables.DeviceTable(tables.flatten(golden_io.load_golden(nm)), 0) for nm in names[:m]]
            dt = tabs[0].flat.dtype
            tdt = torch.float32 if dt == np.float32 else torch.float64
            esz = 4 if dt == np.float32 else 8
            gen = torch.Generator(device="cuda")
            gen.manual_seed(7)
            x = torch.rand(n, dtype=tdt, device="cuda", generator=gen) * 20.0
            x = torch.clamp(x, max=float(np.nextafter(dt.type(20), dt.type(0))))
            ys = [torch.empty_like(x) for _ in tabs]
            yp = [y.data_ptr() for y in ys]
            for order in ("random", "sorted"):
                if order == "sorted":
                    x = torch.sort(x)[0]
                # the LIBRARY decides whether to fuse (records of more than 64 bytes take per-table launches)
                os.environ.pop("AI_B200_NO_FUSED_MULTI", None)
                fused = tables.DeviceTable.eval_multi_ptr(tabs, x.data_ptr(), yp, n, s)
                fm, _ = time_ms(lambda: tables.DeviceTable.eval_multi_ptr(tabs, x.data_ptr(), yp, n, s), args.reps)
                os.environ["AI_B200_NO_FUSED_MULTI"] = "1"
                assert not tables.DeviceTable.eval_multi_ptr(tabs, x.data_ptr(), yp, n, s)
                sm, _ = time_ms(lambda: tables.DeviceTable.eval_multi_ptr(tabs, x.data_ptr(), yp, n, s), args.reps)
                os.environ.pop("AI_B200_NO_FUSED_MULTI", None)
                row = {"group": gname, "m": m, "points": order, "n": n, "library_fused": bool(fused),
                       "fused_ms": fm, "fused_gevals": m * n / (fm * 1e-3) / 1e9,
                       "fused_frac_hbm": (1 + m) * esz * n / (fm * 1e-3) / 1e9 / hbm,
                       "separate_ms": sm, "separate_gevals": m * n / (sm * 1e-3) / 1e9,
                       "separate_frac_hbm": 2 * m * esz * n / (sm * 1e-3) / 1e9 / hbm}
                rows.append(row)
                if not fused:            # same launches on both sides: bytes per point are the separate ones
                    row["fused_frac_hbm"] = 2 * m * esz * n / (fm * 1e-3) / 1e9 / hbm
                print("%-14s %2d %-7s | %9.3f %9.1f %6.1f | %9.3f %9.1f %6.1f | %5.2f %s" % (
                    gname, m, order, fm, row["fused_gevals"], 100 * row["fused_frac_hbm"],
                    sm, row["separate_gevals"], 100 * row["separate_frac_hbm"], sm / fm,
                    "fused" if fused else "per-table launches (library's choice)"))
            del x, ys, tabs
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "%s.json" % args.tag), "w") as fh:
        json.dump({"n": n, "hbm_peak_gbs": hbm, "rows": rows}, fh, indent=1)


if __name__ == "__main__":
    main()
