#!/usr/bin/env python
"""Trace SM clock / power (NVML) while the eval kernel runs back to back for ~0.3 s."""
import os, sys, threading, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import golden_io  # noqa: E402  (fixture loader, tests/golden/golden_io.py)
import numpy as np, torch, pynvml
from adaptive_interpolation_b200 import tables
order = sys.argv[1] if len(sys.argv) > 1 else "random"
pynvml.nvmlInit(); h = pynvml.nvmlDeviceGetHandleByIndex(0)
g = golden_io.load_golden("approx__32o6-p1.19209289551e-06")
tab = tables.DeviceTable(tables.flatten(g), 0)
n = 1 << 28
x = torch.clamp(torch.rand(n, device="cuda") * 20, max=19.999998)
if order == "sorted":
    x = torch.sort(x)[0]
y = torch.empty_like(x)
s = torch.cuda.current_stream().cuda_stream
samples = []; stop = threading.Event()
def sampler():
    t0 = time.perf_counter()
    while not stop.is_set():
        samples.append((time.perf_counter() - t0, pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM),
                        pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_MEM), pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0,
                        pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)))
        time.sleep(0.004)
N = 800
for _ in range(5): tab.eval_ptr(x.data_ptr(), y.data_ptr(), n, s)
torch.cuda.synchronize(); time.sleep(1.0)
th = threading.Thread(target=sampler); th.start()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(N + 1)]
ev[0].record()
for k in range(N):
    tab.eval_ptr(x.data_ptr(), y.data_ptr(), n, s); ev[k + 1].record()
torch.cuda.synchronize(); stop.set(); th.join()
t = np.array([ev[k].elapsed_time(ev[k + 1]) for k in range(N)])
print(order, "per-100 launch mean ms:", [round(float(t[i:i + 100].mean()), 4) for i in range(0, N, 100)])
for smp in samples[:: max(1, len(samples) // 24)]:
    print("t=%.3fs sm=%d mem=%d power=%.0fW reasons=0x%x" % smp)
