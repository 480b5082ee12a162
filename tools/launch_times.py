#!/usr/bin/env python
"""Per-launch kernel times of N back-to-back launches (CUDA events): distribution and trend."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import golden_io  # noqa: E402  (fixture loader, tests/golden/golden_io.py)
import numpy as np, torch
from adaptive_interpolation_b200 import tables
g = golden_io.load_golden("approx__32o6-p1.19209289551e-06")
tab = tables.DeviceTable(tables.flatten(g), 0)
n = 1 << 28
x = torch.rand(n, device="cuda") * 20
x = torch.clamp(x, max=19.999998)
y = torch.empty_like(x)
s = torch.cuda.current_stream().cuda_stream
N = int(sys.argv[1]) if len(sys.argv) > 1 else 300
for _ in range(5):
    tab.eval_ptr(x.data_ptr(), y.data_ptr(), n, s)
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(N + 1)]
ev[0].record()
for k in range(N):
    tab.eval_ptr(x.data_ptr(), y.data_ptr(), n, s)
    ev[k + 1].record()
torch.cuda.synchronize()
t = np.array([ev[k].elapsed_time(ev[k + 1]) for k in range(N)])
print("min %.4f p10 %.4f median %.4f mean %.4f p90 %.4f max %.4f ms" % (t.min(), np.percentile(t, 10), np.median(t), t.mean(), np.percentile(t, 90), t.max()))
print("first 10:", np.round(t[:10], 4))
print("by 50s  :", [round(float(t[i:i + 50].mean()), 4) for i in range(0, N, 50)])

# the same experiment with a plain copy through the same load/store path, and torch's copy
import ctypes
from adaptive_interpolation_b200 import _cabi
L = _cabi.lib()
for name, fn in (("ai_copy_f32", lambda: _cabi.check(L.ai_copy_f32(x.data_ptr(), y.data_ptr(), n, 0, s))),
                 ("torch copy_", lambda: y.copy_(x))):
    torch.cuda.synchronize()
    import time; time.sleep(2.0)
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(N + 1)]
    ev[0].record()
    for k in range(N):
        fn()
        ev[k + 1].record()
    torch.cuda.synchronize()
    t = np.array([ev[k].elapsed_time(ev[k + 1]) for k in range(N)])
    print("%-12s min %.4f median %.4f mean %.4f max %.4f ms | by 50s: %s | GB/s first50 %.0f last50 %.0f" % (
        name, t.min(), np.median(t), t.mean(), t.max(), [round(float(t[i:i + 50].mean()), 4) for i in range(0, N, 50)],
        8.0 * n / t[:50].mean() / 1e6, 8.0 * n / t[-50:].mean() / 1e6))
