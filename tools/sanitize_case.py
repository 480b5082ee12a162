#!/usr/bin/env python
"""Small driver for compute-sanitizer: every residency mode x lookup mode on a few tables,
special inputs included, plus index / sum / host entry points.  Run as
    compute-sanitizer --tool memcheck python tools/sanitize_case.py
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import golden_io  # noqa: E402  (fixture loader, tests/golden/golden_io.py)
TABLES = ["approx__32o6-p1.19209289551e-06", "approx__32o3-p1.19209289551e-06", "gen__cfg4_f1_leg_o7_f32",
          "gen__cfg4_f1_leg_o7_f64", "approximations__64o5-p2.22044604925e-14", "gen__j0_mono_o13_f64",
          "gen__cfg3_j0_cheb_o13_f64", "gen__line1_mono_o1_f64", "approx__64o13-p1.19209289551e-06",
          "approx__64o7-p2.22044604925e-14", "approx__32o7-p1.19209289551e-06",
          "large__approx__64o3-p2.22044604925e-15", "mx__cheb_o3_f64", "mx__leg_o13_f32", "trim__j0_cheb_o6_f64"]


def child():
    import numpy as np
    import torch
    from adaptive_interpolation_b200 import tables
    for name in TABLES:
        g = golden_io.load_golden(name)
        tab = tables.DeviceTable(tables.flatten(g), 0)
        info = tab.info()
        x = np.concatenate([g.x, np.sort(g.x[np.isfinite(g.x)])[:3000]])
        for n in (0, 1, 7, 4099, len(x)):
            xd = torch.from_numpy(np.ascontiguousarray(x[:n])).cuda()
            yd = torch.empty_like(xd)
            idx = torch.empty(n, dtype=torch.int32, device="cuda")
            res = torch.zeros(1, dtype=torch.float64, device="cuda")
            tab.eval_ptr(xd.data_ptr(), yd.data_ptr(), n, None)
            tab.index_ptr(xd.data_ptr(), idx.data_ptr(), n, None)
            tab.sum_ptr(xd.data_ptr(), res.data_ptr(), n, None)
            torch.cuda.synchronize()
        tab.eval_host(x)
        # several tables over the same x (fused when the layout allows it, else per-table launches)
        xd = torch.from_numpy(np.ascontiguousarray(x)).cuda()
        y0, y1 = torch.empty_like(xd), torch.empty_like(xd)
        tables.DeviceTable.eval_multi_ptr([tab, tab], xd.data_ptr(), [y0.data_ptr(), y1.data_ptr()], xd.numel(), None)
        torch.cuda.synchronize()
        print(name, "mode", os.environ.get("AI_B200_FORCE_MODE", "auto"), "cells", os.environ.get("AI_B200_MAX_CELLS", "-"),
              "residency", info["residency"],
              "lookup", info["lookup_mode"], "ok", flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "child":
        child()
    else:
        for mode in (None, "0", "1", "2", "3", "4", "9", "10", "nohdr", "nocell", "norank", "path", "nopath"):
            env = dict(os.environ)
            if mode == "nohdr":
                env["AI_B200_NO_PACKED_HDR"] = env["AI_B200_NO_UNIFORM_HDR"] = "1"
            elif mode == "nocell":
                env["AI_B200_NO_CELL_RECORDS"] = env["AI_B200_NO_UNIFORM_HDR"] = "1"
            elif mode == "norank":
                env["AI_B200_NO_RANK"] = "1"
            elif mode == "path":             # grid capped at two cells: every tree that is a path takes the path lookup
                env["AI_B200_MAX_CELLS"] = "2"
            elif mode == "nopath":           # ... and the bounded search otherwise
                env["AI_B200_MAX_CELLS"] = "2"
                env["AI_B200_NO_PATH"] = "1"
            elif mode is not None:
                env["AI_B200_FORCE_MODE"] = mode
            r = subprocess.run([sys.executable, __file__, "child"], env=env)
            if r.returncode != 0:
                sys.exit(r.returncode)
