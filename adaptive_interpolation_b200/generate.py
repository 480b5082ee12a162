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
    __slots__ = ("ref", "tables", "stamp", "identity")


_REGISTRY = {}


def _identity(approx):
    """Cheap per-call check that `approx` still carries the table it was staged from: the identity,
    buffer address, shape and dtype of tree_1d plus basis and order.  The content hash (`_stamp`)
    is only recomputed when this changes -- hashing a 10^5-leaf tree_1d costs milliseconds, more than
    the kernel it would precede.  (A tree_1d edited IN PLACE needs generate.invalidate(approx).)"""
    t = approx.tree_1d
    addr = t.__array_interface__["data"][0] if isinstance(t, np.ndarray) else None
    return (id(t), addr, getattr(t, "shape", None), str(getattr(t, "dtype", "")), str(approx.basis),
            int(approx.max_order))


def _stamp(approx):
    t = np.asarray(approx.tree_1d)
    return (str(approx.basis), int(approx.max_order), str(t.dtype), t.shape, hash(t.tobytes()))


def invalidate(approx):
    """Drop the device tables cached for `approx` (after editing its tree_1d in place)."""
    _REGISTRY.pop(id(approx), None)


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
    ident = _identity(approx)
    if ent is not None and ent.ref() is approx and ent.identity != ident:
        # same object, different tree_1d array: staged tables stay only if the content is the same
        if ent.stamp == _stamp(approx):
            ent.identity = ident
        else:
            ent = None
    if ent is None or ent.ref() is not approx:
        ent = _Entry()
        ent.tables = {}
        ent.stamp = _stamp(approx)
        ent.identity = ident
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


def _device_of(torch, x):
    """The GPU a call runs on: the device of x when x already lives on one (torch CUDA tensor or
    any __cuda_array_interface__ object), else the current device.  The table is created (and
    cached) per device, and the stream is taken from the same device, so pointers, table and stream
    always agree."""
    if isinstance(x, torch.Tensor):
        if x.is_cuda:
            return int(x.device.index)
    elif hasattr(x, "__cuda_array_interface__"):
        return int(torch.as_tensor(x, device="cuda").device.index)
    return int(torch.cuda.current_device())


def _as_device(torch, x, dtype, device):
    """-> (x_dev tensor, was_numpy).  dtype is the TABLE dtype: the reference kernel is declared
    in it (generate.py:172-174) and feeding another dtype there is undefined; we convert."""
    tdt = torch.float32 if dtype == np.float32 else torch.float64
    if not isinstance(x, torch.Tensor) and hasattr(x, "__cuda_array_interface__"):
        x = torch.as_tensor(x, device="cuda")       # cupy / numba / pycuda arrays: zero-copy view, on ITS device
    if isinstance(x, torch.Tensor):
        if not x.is_cuda:
            return x.to(device=device, dtype=tdt).contiguous().view(-1), False
        return x.to(dtype=tdt).contiguous().view(-1), False
    xa = np.ascontiguousarray(np.asarray(x), dtype=dtype).reshape(-1)
    return torch.from_numpy(xa).to(device), True


def _launch(torch, tab, x_dev, y_dev):
    if x_dev.device.index != tab.device or y_dev.device.index != tab.device:
        raise ValueError("x / y live on cuda:%s but the table was staged on cuda:%d"
                         % (x_dev.device.index, tab.device))
    stream = torch.cuda.current_stream(x_dev.device).cuda_stream
    tab.eval_ptr(x_dev.data_ptr(), y_dev.data_ptr(), x_dev.numel(), stream)


def _is_host_array(torch, x):
    """numpy arrays, lists, scalars ...: anything that is neither a torch tensor nor a CUDA array."""
    return not isinstance(x, torch.Tensor) and not hasattr(x, "__cuda_array_interface__")


def _run_host(tab, x):
    """HOST input (what the reference's callers pass, generate.py:160-166 `cl_array.to_device`): the
    library's chunked H2D / kernel / D2H pipeline (ai_eval_host_*) -- pageable arrays are staged through
    pinned bounce buffers by a pool of copy threads, nothing of the array's full size is allocated on
    the GPU.  Returns (y, kernel seconds)."""
    xa = np.ascontiguousarray(np.asarray(x), dtype=tab.flat.dtype)
    y, ms = tab.eval_host(xa.reshape(-1))
    return y, ms * 1e-3


def run(x, approx):
    """generate.py:158-187: build and run in one call, untimed; returns y."""
    torch = _backend()
    dev = _device_of(torch, x)
    tab = device_table(approx, dev)
    if _is_host_array(torch, x):
        return _run_host(tab, x)[0]
    x_dev, was_numpy = _as_device(torch, x, tab.flat.dtype, torch.device("cuda", dev))
    y_dev = torch.empty_like(x_dev)
    _launch(torch, tab, x_dev, y_dev)
    torch.cuda.current_stream(x_dev.device).synchronize()
    return y_dev.cpu().numpy() if was_numpy else y_dev


def run_single(x, ap):
    """generate.py:480-490: returns (seconds, y); the timed region is the kernel only
    (queue.finish(); t0; launch; queue.finish(); t1), measured here with CUDA events."""
    torch = _backend()
    dev = _device_of(torch, x)
    tab = device_table(ap, dev)
    if _is_host_array(torch, x):
        y, seconds = _run_host(tab, x)            # seconds: the evaluation kernels only, summed over the chunks
        return seconds, y
    x_dev, was_numpy = _as_device(torch, x, tab.flat.dtype, torch.device("cuda", dev))
    y_dev = torch.empty_like(x_dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    stream = torch.cuda.current_stream(x_dev.device)
    stream.synchronize()
    e0.record(stream)
    _launch(torch, tab, x_dev, y_dev)
    e1.record(stream)
    e1.synchronize()
    seconds = e0.elapsed_time(e1) * 1e-3
    return seconds, (y_dev.cpu().numpy() if was_numpy else y_dev)


def pinned_empty(n, dtype):
    """A 1-D numpy array of n elements in page-locked host memory (ai_host_alloc).  Host arrays
    handed to run / run_multi / DeviceTable.eval_host move at full PCIe speed when they are pinned;
    pageable arrays are staged by the library's copy threads at roughly half that rate."""
    import ctypes
    dt = np.dtype(dtype)
    p = ctypes.c_void_p()
    nbytes = max(int(n), 1) * dt.itemsize
    _cabi.check(_cabi.lib().ai_host_alloc(ctypes.byref(p), nbytes))
    buf = (ctypes.c_byte * nbytes).from_address(p.value)
    arr = np.frombuffer(buf, dtype=dt, count=int(n))
    weakref.finalize(buf, _cabi.lib().ai_host_free, p)
    return arr


def run_multi(x, ap, devices=None, timing="kernel", out=None):
    """generate.py:493-502 (the 'multi-threaded' launch).  Here: contiguous shards of the HOST array
    x over the given GPUs of this process (default: all visible), table replicated, no collective.
    One worker thread per device pushes its shard through the library's pinned, chunked
    H2D / kernel / D2H pipeline (ai_eval_host_*; the GIL is released during the call), so the copies
    of all devices overlap.  Returns (seconds, y): seconds is the max over devices of the summed
    kernel time (`timing="kernel"`, the reference's kernel-only convention, generate.py:484-489) or
    the wall time of the whole call including every copy (`timing="wall"`).  `out` (optional): a
    contiguous host array of x's size in the table dtype to receive y -- a pinned one
    (`pinned_empty`) takes the D2H copies at full speed; by default a fresh pageable numpy array is
    allocated, as `y_dev.get()` does upstream."""
    torch = _backend()
    import threading
    import time
    from . import sharding
    if timing not in ("kernel", "wall"):
        raise Exception("timing must be 'kernel' or 'wall'")
    if devices is None:
        devices = list(range(torch.cuda.device_count()))
    if (isinstance(x, torch.Tensor) and x.is_cuda) or hasattr(x, "__cuda_array_interface__"):
        return run_single(x, ap)            # already resident on one device
    dt = tables.table_dtype(ap)
    xa = np.ascontiguousarray(np.asarray(x), dtype=dt).reshape(-1)
    if out is None:
        out = np.empty_like(xa)
    elif not (isinstance(out, np.ndarray) and out.dtype == dt and out.flags.c_contiguous and out.size == xa.size):
        raise Exception("out must be a contiguous %s array of x's size" % dt)
    else:
        out = out.reshape(-1)
    bounds = sharding.shard_bounds(xa.size, len(devices))
    tabs = [device_table(ap, d) for d in devices]          # created before the clock starts
    kernel_ms = [0.0] * len(devices)
    errors = []

    def worker(k):
        lo, hi = bounds[k]
        try:
            if hi > lo:
                _, kernel_ms[k] = tabs[k].eval_host(xa[lo:hi], out[lo:hi])
        except Exception as e:                              # re-raised in the caller's thread
            errors.append(e)

    t0 = time.perf_counter()
    threads = [threading.Thread(target=worker, args=(k,)) for k in range(len(devices))]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    wall = time.perf_counter() - t0
    if errors:
        raise errors[0]
    return (wall if timing == "wall" else max(kernel_ms) * 1e-3), out


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
    dev = _device_of(torch, x)
    tabs = [device_table(ap, dev) for ap in approxs]
    dt = tabs[0].flat.dtype
    if any(t.flat.dtype != dt for t in tabs):
        raise Exception("run_many needs approximations of one dtype")
    x_dev, was_numpy = _as_device(torch, x, dt, torch.device("cuda", dev))
    ys = [torch.empty_like(x_dev) for _ in tabs]
    stream = torch.cuda.current_stream(x_dev.device)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    stream.synchronize()
    e0.record(stream)
    tables.DeviceTable.eval_multi_ptr(tabs, x_dev.data_ptr(), [y.data_ptr() for y in ys], x_dev.numel(),
                                      stream.cuda_stream)
    e1.record(stream)
    e1.synchronize()
    seconds = e0.elapsed_time(e1) * 1e-3
    return seconds, [y.cpu().numpy() if was_numpy else y for y in ys]


def run_index(x, approx):
    """Leaf ordinal per point (bit-exact with the reference's BST walk, approximator.py:285-290)."""
    torch = _backend()
    dev = _device_of(torch, x)
    tab = device_table(approx, dev)
    x_dev, was_numpy = _as_device(torch, x, tab.flat.dtype, torch.device("cuda", dev))
    idx = torch.empty(x_dev.numel(), dtype=torch.int32, device=x_dev.device)
    stream = torch.cuda.current_stream(x_dev.device)
    tab.index_ptr(x_dev.data_ptr(), idx.data_ptr(), x_dev.numel(), stream.cuda_stream)
    stream.synchronize()
    return idx.cpu().numpy() if was_numpy else idx
