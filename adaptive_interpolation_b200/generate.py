"""Mirror of the reference's adaptive_interpolation/generate.py, served by the CUDA library.

Same names, argument meaning and error behaviour as the reference module:

    gen_mono / gen_cheb / gen_leg / gen_ispc (ap) -> str      generate.py:28,62,106,191
    build_code(approx, ispc=False)                            generate.py:408-477
    run(x, approx) -> y                                       generate.py:158-187
    run_single(x, ap) -> (seconds, y)                         generate.py:480-490
    run_multi(x, ap)  -> (seconds, y)                         generate.py:493-502

What changed underneath: the reference emits OpenCL / ISPC source text per approximation and
JIT-builds it (pyopencl) or shells out to ispc + g++ (run_test.py:46-122).  Here nothing is
generated: every (basis, order, dtype) is a precompiled sm_100a kernel template instance in
libai_b200.so, and "building" an approximation means flattening its table and staging it on the
GPU (ai_table_create).  The `gen_*` functions therefore return a short description of the
selected kernel; it is stored in approx.code exactly where the reference stores its source text.

x may be a numpy array (copied to the device and back OUTSIDE the timed region, as
generate.py:481-489 does), a torch CUDA tensor or any object with `__cuda_array_interface__`
(cupy, numba, pycuda: used in place without a copy; a torch CUDA tensor is returned).
No CUDA handles are ever stored on the Approximator, so write_to_file keeps working
(checkpoint row).
"""
from __future__ import absolute_import

import weakref

import numpy as np

from . import _cabi
from . import tables

with_cuda = True          # the reference's `with_pyopencl` flag (generate.py:17-22); see _backend()

_KERNEL_NAMES = {"chebyshev": "kCheb", "legendre": "kLeg", "monomial": "kMono"}


def _backend():
    """torch is the device-memory / stream plumbing.  Raises the reference's error type
    (ValueError, generate.py:187,477) when no usable backend exists -- never falls back."""
    try:
        import torch
    except ImportError as e:                                    # pragma: no cover
        raise ValueError("Function requires torch (device memory plumbing): %s" % e)
    try:
        _cabi.lib()
    except (ImportError, OSError) as e:
        raise ValueError("Function requires the CUDA library libai_b200.so: %s" % e)
    if not torch.cuda.is_available():
        raise ValueError("Function requires a CUDA device (B200 / sm_100a); none is visible.")
    return torch


# ------------------------------------------------------------------ "code generation"
def _describe(ap, basis):
    order = int(ap.max_order)
    if order < 0 or order > _cabi.MAX_ORDER:
        raise Exception("order %d outside the supported range 0..%d" % (order, _cabi.MAX_ORDER))
    dt = tables.table_dtype(ap)
    ctype = "float" if dt == np.float32 else "double"
    s = "// libai_b200.so (sm_100a): ai::eval_kernel<ai::%s, %d, %s>\n" % (_KERNEL_NAMES[basis], order, ctype)
    s += "// leaf lookup: exact power-of-two grid + bounded search over stored breakpoints\n"
    s += "// int ai_eval_%s(const ai_table_t*, const %s* x, %s* y, int64_t n, void* stream);\n" % (
        "f32" if dt == np.float32 else "f64", ctype, ctype)
    if getattr(ap, "size", None) is not None:
        s += "// size hint %d (the element count is a run-time argument here)\n" % int(ap.size)
    return s


def gen_mono(ap):
    return _describe(ap, "monomial")


def gen_cheb(ap):
    return _describe(ap, "chebyshev")


def gen_leg(ap):
    return _describe(ap, "legendre")


def gen_ispc(ap):
    """generate.py:191: the CPU (ISPC) generator.  Chebyshev only upstream
    (adaptive_interpolation.py:53-58); same kernel here."""
    if "none" in getattr(ap, "optimizations", []):
        return "y[0] = x[0]+10;"                       # generate.py:192-193 (ablation stub)
    return _describe(ap, ap.basis)


# ------------------------------------------------------------------ device tables
class _Entry(object):
    __slots__ = ("ref", "tables", "stamp")


_REGISTRY = {}


def _stamp(approx):
    t = np.asarray(approx.tree_1d)
    return (str(approx.basis), int(approx.max_order), str(t.dtype), t.shape, hash(t.tobytes()))


def device_table(approx, device=None):
    """The DeviceTable of `approx` on `device` (created on first use, cached off-object)."""
    torch = _backend()
    if device is None:
        device = torch.cuda.current_device()
    device = int(device)
    if isinstance(approx, tables.DeviceTable):
        return approx
    key = id(approx)
    ent = _REGISTRY.get(key)
    stamp = _stamp(approx)
    if ent is None or ent.ref() is not approx or ent.stamp != stamp:
        ent = _Entry()
        ent.tables = {}
        ent.stamp = stamp
        try:
            ent.ref = weakref.ref(approx)
            weakref.finalize(approx, _REGISTRY.pop, key, None)
        except TypeError:                  # object without weakref support: keep it alive instead
            ent.ref = (lambda obj=approx: obj)
        _REGISTRY[key] = ent
    if device not in ent.tables:
        ent.tables[device] = tables.DeviceTable(tables.flatten(approx), device)
    return ent.tables[device]


def build_code(approx, ispc=False):
    """generate.py:408-477.  ispc=True returns text (the C prototype that replaces the ISPC
    `export void eval(...)` header, generate.py:418-441).  Otherwise stages the table on the
    current GPU and returns (knl, queue, tree_dev), caching plain picklable tokens on
    approx.kernal / .queue / .tree_dev like generate.py:471-473."""
    if ispc:
        if approx.code is None:
            raise Exception("Need to generate code before building the kernel.")
        dt = tables.table_dtype(approx)
        c = "float" if dt == np.float32 else "double"
        header = 'extern "C" int ai_eval_%s(const ai_table_t* table, const %s x[], %s y[], int64_t n, void* stream)' % (
            "f32" if dt == np.float32 else "f64", c, c)
        code = str(approx.code).replace("\n", "\n\t")
        return header + "{\n\n" + code + "\n}"
    torch = _backend()
    dev = torch.cuda.current_device()
    tab = device_table(approx, dev)
    info = tab.info()
    knl = "ai::eval_kernel<%s,%d,%s>" % (_KERNEL_NAMES[tab.flat.basis], tab.flat.order,
                                         "float" if tab.flat.dtype == np.float32 else "double")
    approx.kernal = knl
    approx.queue = int(dev)
    approx.tree_dev = int(info["image_bytes"])
    return knl, int(dev), tab


def _as_device(torch, x, dtype, device):
    """-> (x_dev tensor, was_numpy).  dtype is the TABLE dtype: the reference kernel is declared
    in it (generate.py:172-174) and feeding another dtype there is undefined; we convert."""
    tdt = torch.float32 if dtype == np.float32 else torch.float64
    if not isinstance(x, torch.Tensor) and hasattr(x, "__cuda_array_interface__"):
        x = torch.as_tensor(x, device=device)       # cupy / numba / pycuda arrays: zero-copy view
    if isinstance(x, torch.Tensor):
        if not x.is_cuda:
            return x.to(device=device, dtype=tdt).contiguous().view(-1), False
        return x.to(dtype=tdt).contiguous().view(-1), False
    xa = np.ascontiguousarray(np.asarray(x), dtype=dtype).reshape(-1)
    return torch.from_numpy(xa).to(device), True


def _launch(torch, tab, x_dev, y_dev):
    stream = torch.cuda.current_stream(x_dev.device).cuda_stream
    tab.eval_ptr(x_dev.data_ptr(), y_dev.data_ptr(), x_dev.numel(), stream)


def run(x, approx):
    """generate.py:158-187: build and run in one call, untimed; returns y."""
    torch = _backend()
    dev = torch.cuda.current_device()
    tab = device_table(approx, dev)
    x_dev, was_numpy = _as_device(torch, x, tab.flat.dtype, torch.device("cuda", dev))
    y_dev = torch.empty_like(x_dev)
    _launch(torch, tab, x_dev, y_dev)
    torch.cuda.current_stream().synchronize()
    return y_dev.cpu().numpy() if was_numpy else y_dev


def run_single(x, ap):
    """generate.py:480-490: returns (seconds, y); the timed region is the kernel only
    (queue.finish(); t0; launch; queue.finish(); t1), measured here with CUDA events."""
    torch = _backend()
    dev = torch.cuda.current_device()
    tab = device_table(ap, dev)
    x_dev, was_numpy = _as_device(torch, x, tab.flat.dtype, torch.device("cuda", dev))
    y_dev = torch.empty_like(x_dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.current_stream().synchronize()
    e0.record()
    _launch(torch, tab, x_dev, y_dev)
    e1.record()
    e1.synchronize()
    seconds = e0.elapsed_time(e1) * 1e-3
    return seconds, (y_dev.cpu().numpy() if was_numpy else y_dev)


def run_multi(x, ap, devices=None):
    """generate.py:493-502 (the 'multi-threaded' launch).  Here: contiguous shards of x over the
    given GPUs of this process (default: all visible), table replicated, no collective; returns
    (max over devices of the kernel seconds, y)."""
    torch = _backend()
    from . import sharding
    if devices is None:
        devices = list(range(torch.cuda.device_count()))
    if isinstance(x, torch.Tensor) and x.is_cuda:
        return run_single(x, ap)
    dt = tables.table_dtype(ap)
    xa = np.ascontiguousarray(np.asarray(x), dtype=dt).reshape(-1)
    bounds = sharding.shard_bounds(xa.size, len(devices))
    parts = []
    for d, (lo, hi) in zip(devices, bounds):
        with torch.cuda.device(d):
            tab = device_table(ap, d)
            xd = torch.from_numpy(xa[lo:hi]).to(torch.device("cuda", d))
            yd = torch.empty_like(xd)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            parts.append((d, tab, xd, yd, e0, e1))
    for d, tab, xd, yd, e0, e1 in parts:
        torch.cuda.synchronize(d)
    for d, tab, xd, yd, e0, e1 in parts:            # launch everywhere before waiting anywhere
        with torch.cuda.device(d):
            e0.record()
            _launch(torch, tab, xd, yd)
            e1.record()
    seconds = 0.0
    out = np.empty_like(xa)
    for (d, tab, xd, yd, e0, e1), (lo, hi) in zip(parts, bounds):
        e1.synchronize()
        seconds = max(seconds, e0.elapsed_time(e1) * 1e-3)
        out[lo:hi] = yd.cpu().numpy()
    return seconds, out


def run_many(x, approxs):
    """Several approximations of one dtype over the SAME points in one pass (ai_eval_multi_*; no
    reference counterpart -- the reference re-reads x once per generated kernel).  Returns
    (seconds, [y_0, ..]): the timed region is the device work only, as in run_single.  The library
    fuses the tables into one kernel when they share basis and order and fit in shared memory
    together, else launches them back to back; the outputs are bit-identical to run() either way."""
    torch = _backend()
    approxs = list(approxs)
    if not 1 <= len(approxs) <= _cabi.MAX_MULTI:
        raise Exception("run_many takes 1..%d approximations, got %d" % (_cabi.MAX_MULTI, len(approxs)))
    dev = torch.cuda.current_device()
    tabs = [device_table(ap, dev) for ap in approxs]
    dt = tabs[0].flat.dtype
    if any(t.flat.dtype != dt for t in tabs):
        raise Exception("run_many needs approximations of one dtype")
    x_dev, was_numpy = _as_device(torch, x, dt, torch.device("cuda", dev))
    ys = [torch.empty_like(x_dev) for _ in tabs]
    stream = torch.cuda.current_stream(x_dev.device)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    stream.synchronize()
    e0.record()
    tables.DeviceTable.eval_multi_ptr(tabs, x_dev.data_ptr(), [y.data_ptr() for y in ys], x_dev.numel(),
                                      stream.cuda_stream)
    e1.record()
    e1.synchronize()
    seconds = e0.elapsed_time(e1) * 1e-3
    return seconds, [y.cpu().numpy() if was_numpy else y for y in ys]


def run_index(x, approx):
    """Leaf ordinal per point (bit-exact with the reference's BST walk, approximator.py:285-290)."""
    torch = _backend()
    dev = torch.cuda.current_device()
    tab = device_table(approx, dev)
    x_dev, was_numpy = _as_device(torch, x, tab.flat.dtype, torch.device("cuda", dev))
    idx = torch.empty(x_dev.numel(), dtype=torch.int32, device=x_dev.device)
    tab.index_ptr(x_dev.data_ptr(), idx.data_ptr(), x_dev.numel(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.current_stream().synchronize()
    return idx.cpu().numpy() if was_numpy else idx
