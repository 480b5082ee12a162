"""Host-side table handling: flatten the reference's Approximator data into the SoA form the
CUDA library consumes, and own the device tables.

Reference data this reads (never writes):
  adaptive_interpolation/approximator.py:118-136   tree_1d(): AoS node array, stride order+6,
        node = [mid, c0..cn, a, b, left_off, right_off], child links are ELEMENT offsets stored
        as table-dtype floats, leaves link to themselves (:196-199)
  adaptive_interpolation/approximator.py:211-216   interval_a / interval_b / coeff SoA lists
  run_test.py:242-246                              tables are cast to the table dtype before use

`flatten(approx)` works on any object carrying those attributes: the reference's own
Approximator (all three pickle schema versions), this package's stand-in
(approximator.py here), or any record with the same attributes.
"""
import ctypes
import weakref

import numpy as np

from . import _cabi

class FlatTable(object):
    """Sorted breakpoints + per-leaf coefficients, in the table dtype."""

    def __init__(self, basis, order, dtype, breaks, coef):
        self.basis = str(basis)
        self.order = int(order)
        self.dtype = np.dtype(dtype)
        self.breaks = np.ascontiguousarray(breaks, dtype=self.dtype)
        self.coef = np.ascontiguousarray(coef, dtype=self.dtype).reshape(len(self.breaks) - 1, self.order + 1)
        if self.basis not in _cabi.BASIS_IDS:
            raise Exception("Incorrect basis provided: monomial, chebyshev, legendre.")
        if self.dtype not in (np.dtype(np.float32), np.dtype(np.float64)):
            raise Exception("Incorrect data type specified")

    @property
    def n_leaves(self):
        return len(self.breaks) - 1

    @property
    def lower_bound(self):
        return self.breaks[0]

    @property
    def upper_bound(self):
        return self.breaks[-1]


def flatten_tree_1d(tree_1d, order):
    """Enumerate the leaves of a standard-layout tree_1d left to right.

    Follows the child links exactly as Approximator.evaluate does (approximator.py:285-290):
    offsets [n+4] / [n+5] cast with int().  Returns (a, b, coef[L, n+1]) in tree_1d's dtype."""
    t = np.asarray(tree_1d)
    n = int(order)
    stride = n + 6
    if t.ndim != 1 or len(t) == 0 or len(t) % stride:
        raise ValueError("tree_1d of length %d is not a whole number of stride-%d nodes" % (len(t), stride))
    a, b, coef = [], [], []
    stack = [0]
    seen = 0
    while stack:
        o = stack.pop()
        seen += 1
        if seen > len(t):
            raise ValueError("tree_1d child links form a cycle")
        if o < 0 or o + stride > len(t) or o % stride:
            raise ValueError("tree_1d child offset %d out of range" % o)
        lo, ro = int(t[o + n + 4]), int(t[o + n + 5])
        if lo == o or ro == o:
            a.append(t[o + n + 2])
            b.append(t[o + n + 3])
            coef.append(t[o + 1:o + n + 2])
        else:
            stack.append(ro)        # right first so that left pops first: left-to-right order
            stack.append(lo)
    a = np.array(a, dtype=t.dtype)
    b = np.array(b, dtype=t.dtype)
    if np.any(a[1:] != b[:-1]):
        k = int(np.nonzero(a[1:] != b[:-1])[0][0])
        raise ValueError("leaves %d and %d are not contiguous (%r != %r)" % (k, k + 1, b[k], a[k + 1]))
    return a, b, np.array(coef, dtype=t.dtype).reshape(len(a), n + 1)


def flatten_tree_1d_trim(tree_1d, order):
    """Enumerate the leaves of a `trim_data`-layout tree_1d (approximator.py:202-206, 232-233).

    Nodes are stored in preorder; a leaf is [mid, self, self, c0..cn, a, b] (both links equal), an
    inner node is [mid, left NODE index, right NODE index].  The reference's own walk reads those
    node indices as element offsets (approximator.py:275-290), so the links cannot be followed; a
    preorder scan visits the same leaves left to right.  Returns (a, b, coef[L, n+1])."""
    t = np.asarray(tree_1d)
    n = int(order)
    a, b, coef = [], [], []
    o, node, pending = 0, 0, []
    while o < len(t):
        if o + 3 > len(t):
            raise ValueError("trim_data tree_1d ends inside node %d" % node)
        if pending and pending[-1] == node:
            pending.pop()
        if t[o + 1] == t[o + 2]:
            if o + n + 6 > len(t):
                raise ValueError("trim_data tree_1d ends inside leaf node %d" % node)
            coef.append(t[o + 3:o + n + 4])
            a.append(t[o + n + 4])
            b.append(t[o + n + 5])
            o += n + 6
        else:
            if int(t[o + 1]) != node + 1:
                raise ValueError("trim_data node %d: left child %r is not the next node" % (node, t[o + 1]))
            pending.append(int(t[o + 2]))
            o += 3
        node += 1
    if pending or 2 * len(a) - 1 != node:
        raise ValueError("trim_data tree_1d is not a full binary tree (%d nodes, %d leaves)" % (node, len(a)))
    a = np.array(a, dtype=t.dtype)
    b = np.array(b, dtype=t.dtype)
    if np.any(a[1:] != b[:-1]):
        k = int(np.nonzero(a[1:] != b[:-1])[0][0])
        raise ValueError("leaves %d and %d are not contiguous (%r != %r)" % (k, k + 1, b[k], a[k + 1]))
    return a, b, np.array(coef, dtype=t.dtype).reshape(len(a), n + 1)


def flatten_nodes(root, basis, order, dtype):
    """Flatten the fitter's raw tree (adapt.py:14-48 Tree/Node objects, node.data = [mid, coeff,
    [a, b], err], adapt.py:233 / :381) straight to a FlatTable.  This is the route for deep or
    skewed trees that the reference's Approximator constructor cannot take (approximator.py:98,
    158, 178 -- SURVEY.md appendix C-2/C-3).  Children whose `data` was never set (the fitter
    returned early with accurate=False, adapt.py:206-227) make their parent the leaf, as the
    reference's own evaluate_tree effectively does (approximator.py:252-256)."""
    dt = np.dtype(dtype)
    a, b, coef = [], [], []
    stack = [root]
    while stack:
        node = stack.pop()
        left, right = getattr(node, "left", 0), getattr(node, "right", 0)
        is_leaf = (left == 0 or right == 0 or left is node or right is node
                   or not isinstance(getattr(left, "data", 0), (list, tuple))
                   or not isinstance(getattr(right, "data", 0), (list, tuple)))
        if is_leaf:
            data = node.data
            a.append(data[2][0])
            b.append(data[2][1])
            c = np.asarray(data[1], dtype=np.float64)
            if len(c) != int(order) + 1:
                raise ValueError("leaf has %d coefficients, expected order+1 = %d" % (len(c), int(order) + 1))
            coef.append(c)
        else:
            stack.append(right)
            stack.append(left)
    a = np.asarray(a, dtype=dt)
    b = np.asarray(b, dtype=dt)
    if np.any(a[1:] != b[:-1]):
        k = int(np.nonzero(a[1:] != b[:-1])[0][0])
        raise ValueError("leaves %d and %d are not contiguous (%r != %r)" % (k, k + 1, b[k], a[k + 1]))
    return FlatTable(basis, order, dt, np.concatenate([a, b[-1:]]), np.asarray(coef))


def decode_packed_map(approx):
    """Decode the reference's packed uniform-grid map into (leaf, first_cell, n_cells) per cell.

    Two encodings exist in the saved tables:
      V2/V3 `cmap`: leaf | l << D | L << 2D with D = max_level + 1   (approximator.py:155-188)
      V1 `ci*` files: `map` itself is int64 leaf | l << 20 | r << 40  (run_perf.py:315-319),
                      r = l + L
    Used to check that tables fitted for the "calc intervals" variant (generate.py:335-348)
    describe exactly the leaves this library flattens; the kernels do not read it."""
    cmap = np.asarray(getattr(approx, "cmap", []))
    if cmap.size:
        D = int(approx.num_levels)              # tree.max_level + 1
        mask = (1 << D) - 1
        c = cmap.astype(np.int64)
        return c & mask, (c >> D) & mask, c >> (2 * D)
    m = np.asarray(approx.map).astype(np.int64)
    if m.size and m.max() >= (1 << 20):
        mask = (1 << 20) - 1
        leaf, l, r = m & mask, (m >> 20) & mask, m >> 40
        return leaf, l, r - l
    raise ValueError("this table carries no packed map")


def table_dtype(approx):
    dt = getattr(approx, "dtype", None)
    if dt is None:
        dt = np.asarray(approx.tree_1d).dtype
    return np.dtype(dt)


def flatten(approx):
    """Approximator-like object -> FlatTable, cross-checked against its SoA arrays when present."""
    if isinstance(approx, FlatTable):
        return approx
    opts = list(getattr(approx, "optimizations", []) or [])
    dt = table_dtype(approx)
    n = int(approx.max_order)
    tree_1d = np.asarray(approx.tree_1d, dtype=dt)          # run_test.py:242 cast
    if "trim_data" in opts:
        a, b, coef = flatten_tree_1d_trim(tree_1d, n)
    else:
        a, b, coef = flatten_tree_1d(tree_1d, n)
    ia = getattr(approx, "interval_a", None)
    if ia is not None and len(ia) == len(a):
        ib, co = approx.interval_b, approx.coeff
        ok = (np.array_equal(np.asarray(ia, dtype=dt), a) and np.array_equal(np.asarray(ib, dtype=dt), b)
              and np.array_equal(np.asarray(co, dtype=dt).reshape(len(a), n + 1), coef))
        if not ok:
            raise ValueError("Approximator tree_1d and interval_a/interval_b/coeff disagree")
    breaks = np.concatenate([a, b[-1:]])
    return FlatTable(approx.basis, n, dt, breaks, coef)


def plan(flat, smem_cap_bytes=0):
    """Host-only dry run of the table construction (ai_table_plan): lookup grid, record layout and
    residency mode the library would choose, as a dict.  Needs no GPU."""
    info = _cabi.TableInfo()
    _cabi.check(_cabi.lib().ai_table_plan(
        _cabi.BASIS_IDS[flat.basis], flat.order,
        _cabi.DTYPE_F32 if flat.dtype == np.float32 else _cabi.DTYPE_F64,
        flat.n_leaves, flat.breaks.ctypes.data, flat.coef.ctypes.data, int(smem_cap_bytes), ctypes.byref(info)))
    return info.as_dict()


# ------------------------------------------------------------------------- device tables
class DeviceTable(object):
    """Owns one ai_table_t handle (a table resident on one GPU)."""

    def __init__(self, flat, device=0):
        self.flat = flat
        self.device = int(device)
        L = _cabi.lib()
        h = ctypes.c_void_p()
        _cabi.check(L.ai_table_create(
            _cabi.BASIS_IDS[flat.basis], flat.order,
            _cabi.DTYPE_F32 if flat.dtype == np.float32 else _cabi.DTYPE_F64,
            flat.n_leaves, flat.breaks.ctypes.data, flat.coef.ctypes.data, self.device, ctypes.byref(h)))
        self._h = h
        self._finalizer = weakref.finalize(self, L.ai_table_destroy, h)

    @classmethod
    def from_tree_1d(cls, basis, order, dtype, num_levels, tree_1d, device=0, layout=_cabi.TREE_LAYOUT_STANDARD):
        """Flattening done inside the library (ai_table_create_from_tree); layout
        _cabi.TREE_LAYOUT_TRIM for "trim_data" tables."""
        self = cls.__new__(cls)
        L = _cabi.lib()
        dt = np.dtype(dtype)
        t = np.ascontiguousarray(tree_1d, dtype=dt)
        h = ctypes.c_void_p()
        _cabi.check(L.ai_table_create_from_tree(
            _cabi.BASIS_IDS[basis], int(order), _cabi.DTYPE_F32 if dt == np.float32 else _cabi.DTYPE_F64,
            int(num_levels), int(layout), t.ctypes.data, len(t), int(device), ctypes.byref(h)))
        self._h = h
        self.device = int(device)
        self._finalizer = weakref.finalize(self, L.ai_table_destroy, h)
        info = self.info()
        nl = info["n_leaves"]
        breaks = np.empty(nl + 1, dtype=dt)
        coef = np.empty(nl * (int(order) + 1), dtype=dt)
        _cabi.check(L.ai_table_get_flat(h, breaks.ctypes.data, coef.ctypes.data))
        self.flat = FlatTable(basis, order, dt, breaks, coef)
        return self

    @property
    def handle(self):
        return self._h

    def info(self):
        info = _cabi.TableInfo()
        _cabi.check(_cabi.lib().ai_table_get_info(self._h, ctypes.byref(info)))
        return info.as_dict()

    def close(self):
        self._finalizer()

    # raw-pointer entry points (device pointers as ints, stream as int/None)
    def eval_ptr(self, x_ptr, y_ptr, n, stream=None):
        fn = _cabi.lib().ai_eval_f32 if self.flat.dtype == np.float32 else _cabi.lib().ai_eval_f64
        _cabi.check(fn(self._h, x_ptr, y_ptr, int(n), stream))

    def index_ptr(self, x_ptr, idx_ptr, n, stream=None):
        fn = _cabi.lib().ai_eval_index_f32 if self.flat.dtype == np.float32 else _cabi.lib().ai_eval_index_f64
        _cabi.check(fn(self._h, x_ptr, idx_ptr, int(n), stream))

    def sum_ptr(self, x_ptr, result_ptr, n, stream=None):
        fn = _cabi.lib().ai_eval_sum_f32 if self.flat.dtype == np.float32 else _cabi.lib().ai_eval_sum_f64
        _cabi.check(fn(self._h, x_ptr, result_ptr, int(n), stream))

    @staticmethod
    def eval_multi_ptr(tabs, x_ptr, y_ptrs, n, stream=None):
        """Several tables over the same x in one pass (ai_eval_multi_*).  Returns True when the
        fused kernel ran, False when the library issued one launch per table."""
        m = len(tabs)
        if m != len(y_ptrs):
            raise ValueError("one output pointer per table")
        f32 = tabs[0].flat.dtype == np.float32
        fn = _cabi.lib().ai_eval_multi_f32 if f32 else _cabi.lib().ai_eval_multi_f64
        hs = (ctypes.c_void_p * m)(*[t._h for t in tabs])
        ys = (ctypes.c_void_p * m)(*[int(p) for p in y_ptrs])
        fused = ctypes.c_int(0)
        _cabi.check(fn(hs, m, x_ptr, ys, int(n), stream, ctypes.byref(fused)))
        return bool(fused.value)

    def eval_host(self, x, y=None):
        """The reference-facing call on HOST arrays: H2D + kernel + D2H pipelined inside the
        library (ai_eval_host_*).  Returns (y, kernel_ms)."""
        x = np.ascontiguousarray(x, dtype=self.flat.dtype)
        if y is None:
            y = np.empty_like(x)
        if y.dtype != self.flat.dtype or not y.flags.c_contiguous or y.size != x.size:
            raise ValueError("y must be a contiguous %s array of x's size" % self.flat.dtype)
        ms = ctypes.c_double(0.0)
        fn = _cabi.lib().ai_eval_host_f32 if self.flat.dtype == np.float32 else _cabi.lib().ai_eval_host_f64
        _cabi.check(fn(self._h, x.ctypes.data, y.ctypes.data, x.size, ctypes.byref(ms)))
        return y, ms.value
