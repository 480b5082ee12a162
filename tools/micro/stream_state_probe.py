"""The pure-streaming (sorted-input) path has two timing states on this pool's boxes: the SAME
launch on the SAME buffers runs at 0.342-0.347 ms (94.5-96% of the copy peak) or at 0.360-0.363 ms
(90-91%), and the state flips with unrelated allocation events in the process (phases A-E below),
not with the kernel, the clocks (1965 MHz throughout, no throttle reason), an NVML polling thread
or the input values.  bench.py's `cfg2-sorted` record therefore reads 0.91 or 0.95 depending on
the run; the random-point path (SM-side co-limited) does not show it.
    python tools/micro/stream_state_probe.py        (profiles/r2_stream_state_probe.txt)
"""
import sys, time
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests/golden")
import numpy as np, torch, golden_io, bench
from adaptive_interpolation_b200 import tables
g = golden_io.load_golden("approx__32o6-p1.19209289551e-06")
f = tables.flatten(g); tab = tables.DeviceTable(f, 0)
n = 1 << 28
s = torch.cuda.current_stream().cuda_stream
x = torch.linspace(0.0, 1.0, n + 1, dtype=torch.float64, device="cuda")[:-1].mul_(20.0).float().contiguous()
x.clamp_(max=float(np.nextafter(np.float32(20), np.float32(0))))
y = torch.empty_like(x)
def batch(reps=50):
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for a, b in ev:
        a.record(); tab.eval_ptr(x.data_ptr(), y.data_ptr(), n, s); b.record()
    torch.cuda.synchronize()
    t = np.array([a.elapsed_time(b) for a, b in ev])
    return np.median(t), t.mean()
def phase(label, k=6):
    out = []
    for _ in range(k):
        out.append(batch()); time.sleep(0.2)
    print("%-34s medians %s means %s" % (label, np.round([o[0] for o in out], 4), np.round([o[1] for o in out], 4)), flush=True)
phase("A: nothing else")
with bench.ClockSampler(0) as cs:
    phase("B: ClockSampler thread running")
phase("C: sampler stopped")
z = x.clone()
phase("D: after a 1 GiB clone")
del z; torch.cuda.empty_cache()
phase("E: clone freed")
import pynvml
h = pynvml.nvmlDeviceGetHandleByIndex(0)
try:
    print("persistence", pynvml.nvmlDeviceGetPersistenceMode(h), "pstate", pynvml.nvmlDeviceGetPerformanceState(h))
except Exception as e:
    print("nvml query failed", e)
