#!/usr/bin/env python
"""Kernel time of the reduction variant (ai_eval_sum_*, the reference's no-"output" mode,
generate.py:284,392-394) next to the evaluation kernel on the same points.

    python tools/sum_bench.py [--log2n 28] [--tables name,name]
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

from adaptive_interpolation_b200 import tables  # noqa: E402

TABLES = ["approx__32o6-p1.19209289551e-06", "approx__32o3-p1.19209289551e-06", "approx__64o6-p1.19209289551e-06",
          "approx__64o7-p2.22044604925e-14", "gen__cfg3_j0_cheb_o13_f64", "gen__cfg4_f1_leg_o7_f32",
          "approximations__64o5-p2.22044604925e-14", "large__approx__64o3-p2.22044604925e-14"]


def time_ms(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for a, b in ev:
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    return float(np.median([a.elapsed_time(b) for a, b in ev]))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log2n", type=int, default=28)
    ap.add_argument("--tables", default=",".join(TABLES))
    ap.add_argument("--tag", default="sum_bench")
    args = ap.parse_args()
    n = 1 << args.log2n
    peaks = os.path.join(ROOT, "MEASURED_PEAKS.json")
    hbm = json.load(open(peaks))["hbm_gbs"] if os.path.exists(peaks) else 6536.4
    s = torch.cuda.current_stream().cuda_stream
    rows = []
    print("%-44s %4s | %9s %8s | %9s %8s %7s | %s" % ("table", "T", "eval ms", "Gpt/s", "sum ms", "Gpt/s", "%hbm(x)", "sum check"))
    for name in args.tables.split(","):
        f = tables.flatten(golden_io.load_golden(name))
        tab = tables.DeviceTable(f, 0)
        tdt = torch.float32 if f.dtype == np.float32 else torch.float64
        lo, hi = float(f.breaks[0]), float(f.breaks[-1])
        x = (torch.rand(n, device="cuda", dtype=tdt) * (hi - lo) + lo).clamp_(max=float(np.nextafter(f.dtype.type(hi), f.dtype.type(lo))))
        y = torch.empty_like(x)
        res = torch.zeros(1, dtype=torch.float64, device="cuda")
        te = time_ms(lambda: tab.eval_ptr(x.data_ptr(), y.data_ptr(), n, s))
        ts = time_ms(lambda: tab.sum_ptr(x.data_ptr(), res.data_ptr(), n, s))
        idx = torch.empty(n, dtype=torch.int32, device="cuda")
        ti = time_ms(lambda: tab.index_ptr(x.data_ptr(), idx.data_ptr(), n, s))
        del idx
        want = float(y.double().sum())
        got = float(res.item())
        rel = abs(got - want) / max(1.0, abs(want))
        esz = f.dtype.itemsize
        row = {"table": name, "dtype": "f32" if esz == 4 else "f64", "eval_ms": te, "sum_ms": ts,
               "eval_gpts": n / te / 1e6, "sum_gpts": n / ts / 1e6, "sum_frac_of_hbm_reading_x": n * esz / (ts * 1e-3) / 1e9 / hbm,
               "sum_rel_diff_vs_sum_of_eval": rel, "index_ms": ti, "index_gpts": n / ti / 1e6,
               "index_frac_of_hbm": n * (esz + 4) / (ti * 1e-3) / 1e9 / hbm}
        rows.append(row)
        print("%-44s %4s | %9.3f %8.1f | %9.3f %8.1f %7.1f | %.2e | index %7.3f ms %7.1f Gpt/s %5.1f%% of HBM"
              % (name[:44], row["dtype"], te, row["eval_gpts"], ts, row["sum_gpts"], 100 * row["sum_frac_of_hbm_reading_x"], rel,
                 ti, row["index_gpts"], 100 * row["index_frac_of_hbm"]))
        tab.close()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump({"n": n, "rows": rows}, open(os.path.join(ROOT, "gpurun_out", args.tag + ".json"), "w"), indent=1)


if __name__ == "__main__":
    main()
