"""Which event flips the sorted-input path between its two timing states (tools/micro/stream_state_probe.py)?
Phases: baseline; allocate 1 GiB untouched; memset it; free it; D2D copy into a new 1 GiB buffer; free.
Memory / SM clocks and power are sampled per phase."""
import subprocess, sys, time
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests/golden")
import numpy as np, torch, golden_io
from adaptive_interpolation_b200 import tables
g = golden_io.load_golden("approx__32o6-p1.19209289551e-06")
f = tables.flatten(g); tab = tables.DeviceTable(f, 0)
n = 1 << 28
s = torch.cuda.current_stream().cuda_stream
x = torch.linspace(0.0, 1.0, n + 1, dtype=torch.float64, device="cuda")[:-1].mul_(20.0).float().contiguous()
x.clamp_(max=float(np.nextafter(np.float32(20), np.float32(0))))
y = torch.empty_like(x)
print("x %x y %x" % (x.data_ptr(), y.data_ptr()))
def smi():
    return subprocess.run(["nvidia-smi", "--query-gpu=clocks.mem,clocks.sm,power.draw,temperature.gpu,temperature.memory",
                           "--format=csv,noheader"], capture_output=True, text=True).stdout.strip()
def batch(reps=50):
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for a, b in ev:
        a.record(); tab.eval_ptr(x.data_ptr(), y.data_ptr(), n, s); b.record()
    torch.cuda.synchronize()
    return float(np.median([a.elapsed_time(b) for a, b in ev]))
def phase(label, k=4):
    out = []
    for _ in range(k):
        out.append(batch()); time.sleep(0.2)
    print("%-44s medians %s | %s" % (label, np.round(out, 4), smi()), flush=True)
phase("A: baseline")
z = torch.empty(n, dtype=torch.float32, device="cuda")
print("z %x" % z.data_ptr())
phase("B: 1 GiB allocated, untouched")
z.zero_(); torch.cuda.synchronize()
phase("C: memset of that buffer")
del z; torch.cuda.empty_cache()
phase("D: freed")
z = torch.empty(n, dtype=torch.float32, device="cuda"); z.copy_(x); torch.cuda.synchronize()
print("z %x" % z.data_ptr())
phase("E: D2D copy x -> new 1 GiB buffer")
del z; torch.cuda.empty_cache()
phase("F: freed")
y2 = torch.empty_like(x)
print("y2 %x" % y2.data_ptr())
def batch2(reps=50):
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for a, b in ev:
        a.record(); tab.eval_ptr(x.data_ptr(), y2.data_ptr(), n, s); b.record()
    torch.cuda.synchronize()
    return float(np.median([a.elapsed_time(b) for a, b in ev]))
print("G: same x, a NEW y buffer                    medians", np.round([batch2() for _ in range(4)], 4), "|", smi())
phase("H: back to the first y")
