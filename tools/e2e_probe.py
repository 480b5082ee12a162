#!/usr/bin/env python
"""e2e experiment: host pipeline chunk/slot settings vs zero-copy (kernel reads/writes pinned host
memory directly over PCIe through UVA).  Prints Gpt/s and GB/s per direction."""
import ctypes, os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import golden_io  # noqa: E402  (fixture loader, tests/golden/golden_io.py)
import numpy as np

def one(n_log2=28):
    import torch
    from adaptive_interpolation_b200 import _cabi, tables
    g = golden_io.load_golden("approx__32o6-p1.19209289551e-06")
    tab = tables.DeviceTable(tables.flatten(g), 0)
    n = 1 << n_log2
    hp = ctypes.c_void_p()
    _cabi.check(_cabi.lib().ai_host_alloc(ctypes.byref(hp), 8 * n))
    buf = np.ctypeslib.as_array(ctypes.cast(hp, ctypes.POINTER(ctypes.c_byte)), (8 * n,)).view(np.float32)
    hx, hy = buf[:n], buf[n:]
    hx[:] = np.random.default_rng(0).uniform(0, 20, n).astype(np.float32)
    mode = os.environ.get("E2E_MODE", "pipe")
    def step():
        if mode == "pipe":
            tab.eval_host(hx, hy)
        else:
            tab.eval_ptr(hx.ctypes.data, hy.ctypes.data, n, None)
            torch.cuda.synchronize()
    step(); step()
    t0 = time.perf_counter()
    reps = 5
    for _ in range(reps):
        step()
    dt = (time.perf_counter() - t0) / reps
    print("%-10s chunk=%-4s slots=%-2s : %.2f ms/step  %.2f Gpt/s  %.1f GB/s per direction" % (
        mode, os.environ.get("AI_B200_HOST_CHUNK_MB", "32"), os.environ.get("AI_B200_HOST_SLOTS", "3"),
        dt * 1e3, n / dt / 1e9, 4 * n / dt / 1e9), flush=True)
    ref = tables.DeviceTable(tables.flatten(g), 0).eval_host(hx[:1 << 20].copy())[0]
    assert np.array_equal(ref, hy[:1 << 20])

if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "child":
        one()
    else:
        for env in ([{"E2E_MODE": "pipe", "AI_B200_HOST_CHUNK_MB": str(c), "AI_B200_HOST_SLOTS": str(s)}
                     for c in (8, 16, 32, 64, 128) for s in (2, 3, 4)] + [{"E2E_MODE": "zerocopy"}]):
            subprocess.run([sys.executable, __file__, "child"], env=dict(os.environ, **env))
