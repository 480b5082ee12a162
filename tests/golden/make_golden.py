#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/tables/ by RUNNING THE REFERENCE.

This script is the only thing in the repo that imports /root/reference.  It runs in the
build container (where the reference is mounted read-only); the GPU box has no reference,
so everything the tests/bench need from it is frozen here into small .npz files:

  * the table data of every saved approximation (27 pickles in approx/ and approximations/),
    exactly as stored on the reference's Approximator object (tree_1d, interval_a/b, coeff,
    map, cmap, ...): adaptive_interpolation/approximator.py:22-136,191-243
  * tables the BASELINE configs need that are not saved upstream, produced by the reference's
    own fitter adaptive_interpolation/adapt.py (make_interpolant / Interpolant.run_adapt)
  * for every table: input points x (table dtype), the reference's own evaluation
    y_ref = Approximator.evaluate(x) (approximator.py:273-296), the leaf ordinal idx_ref the
    reference's BST walk selects (approximator.py:285-290), and f_true (float64).

Usage (from the repo root, in the container):
    python tests/golden/make_golden.py            # writes tests/golden/tables/*.npz + MANIFEST.json

Nothing is written outside tests/golden/.  The reference is never modified; its known
defects on modern numpy are shimmed in-process.
"""
import io
import json
import os
import sys
import contextlib
import warnings

import numpy as np
import scipy.special as spec

REF = os.environ.get("AI_REFERENCE_ROOT", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "tables")

warnings.filterwarnings("ignore")
sys.path.insert(0, REF)
sys.setrecursionlimit(10000)

import adaptive_interpolation.adapt as ref_adapt            # noqa: E402
import adaptive_interpolation.approximator as ref_app      # noqa: E402
import adaptive_interpolation.adaptive_interpolation as ref_api  # noqa: E402


# ---------------------------------------------------------------- functions that were fitted
def f(x):
    """The function every saved pickle was fitted to (run_test.py / demo.py `f`): Bessel J0.
    The pickles reference it as __main__.f, so it must live at module level here."""
    return spec.jn(0, x)


def sinsin(x):
    return np.sin(np.sin(x))          # tests/performance_tests.py:133 (func1)


def cossin(x):
    return np.cos(np.sin(x))          # tests/performance_tests.py:134 (func2)


def poly4(x):
    return 4.123 * x**4 - 5.6 * x**3 - x**2 + 4.5      # tests/performance_tests.py:114


def poly6(x):
    return x**6 - 3 * x**5 - 2 * x**4 + x - 3          # tests/performance_tests.py:115


def line1(x):
    return 3 * x + 7                                   # tests/performance_tests.py:113


def f1(x0):
    """demo.py:39-56 piecewise/discontinuous demo function (restated so that scalar and
    array inputs both work; same branches and constants)."""
    x0 = np.atleast_1d(np.asarray(x0, dtype=np.float64))
    out = np.zeros_like(x0)
    for k, x in enumerate(x0):
        if x < 1:
            out[k] = 1 + x
        elif x < 2.02:
            out[k] = 1 + x**2
        elif x < 3.5:
            out[k] = -3 * np.log(x)
        elif x < 4.4:
            out[k] = np.exp(np.sqrt(x))
        elif x < 7.001:
            out[k] = 3
        elif x < 9.306:
            out[k] = np.sqrt(x**4.4) / 100.
        else:
            out[k] = x - 3
    return out


FUNCS = {"j0": f, "sinsin": sinsin, "cossin": cossin, "poly4": poly4, "poly6": poly6,
         "line1": line1, "f1": f1}


# ---------------------------------------------------------------- shims for reference defects
@contextlib.contextmanager
def linspace_int_shim():
    """adapt.py:182-188 passes float counts to np.linspace (TypeError on numpy >= 1.18)."""
    real = np.linspace

    def patched(start, stop, num=50, *a, **k):
        return real(start, stop, int(num), *a, **k)
    np.linspace = patched
    try:
        yield
    finally:
        np.linspace = real


@contextlib.contextmanager
def quiet():
    with contextlib.redirect_stdout(io.StringIO()):
        yield


def prune_dataless(node):
    """adapt.py:206-227: with accurate=False the fitter can return before setting node.data,
    leaving children whose data == 0.  Make the parent the leaf."""
    if node.left == 0 and node.right == 0:
        return
    bad = (not isinstance(node.left.data, list)) or (not isinstance(node.right.data, list))
    if bad:
        node.left, node.right = 0, 0
        return
    prune_dataless(node.left)
    prune_dataless(node.right)


def tree_stats(root):
    size, depth, stack = 0, 0, [root]
    while stack:
        n = stack.pop()
        size += 1
        depth = max(depth, n.level)
        if n.left != 0:
            stack.append(n.left)
            stack.append(n.right)
    return size, depth


def approximator_from_raw_tree(interp):
    """Drive the reference's own Approximator methods on an Interpolant whose tree the
    reference constructor cannot handle (skewed/deep trees trip approximator.py:98,158,178).
    Approximator(None) does nothing in __init__, so we set the attributes its methods read
    and call make_tree_array / tree_1d / (later) evaluate unchanged."""
    prune_dataless(interp.tree.root)
    size, depth = tree_stats(interp.tree.root)
    interp.tree.size, interp.tree.max_level = size, depth
    ap = ref_app.Approximator()
    ap.basis_function = interp.basis_function
    ap.tree = interp.tree
    ap.error = interp.allowed_error
    ap.upper_bound, ap.lower_bound = interp.upper_bound, interp.lower_bound
    ap.basis = interp.basis
    ap.optimizations = []
    ap.mid, ap.coeff, ap.interval_a, ap.interval_b, ap.map, ap.cmap = [], [], [], [], [], []
    ap.left, ap.right = [0] * size, [0] * size
    ap.curr_index, ap.leaf_index = -1, -1
    ap.max_order = int(interp.max_order)
    ap.num_levels = depth + 1
    ap.tree_vector = [None] * size
    ap.make_tree_array(interp.tree.root)
    ap.tree_1d = np.array(ap.tree_1d(), dtype=interp.dtype)
    ap.dtype = interp.dtype
    ap.dtype_name = "float" if interp.dtype == np.float32 else "double"
    ap.code, ap.size, ap.vector_width = 0, None, None
    # SoA arrays, taken straight from the leaves in order
    leaves = []

    def walk(n):
        if n.left == 0 and n.right == 0:
            leaves.append(n)
        else:
            walk(n.left)
            walk(n.right)
    walk(interp.tree.root)
    ap.interval_a = [n.data[2][0] for n in leaves]
    ap.interval_b = [n.data[2][1] for n in leaves]
    ap.coeff = [c for n in leaves for c in n.data[1]]
    ap.map, ap.cmap = np.zeros(0, np.int32), np.zeros(0, np.int32)
    return ap


# ---------------------------------------------------------------- reference evaluation
def ref_bst_leaf_offsets(ap, x):
    """The walk of approximator.py:285-290, returning the element offset of the selected node."""
    n = int(ap.max_order)
    child_index = n + 4
    t = ap.tree_1d
    out = np.empty(len(x), dtype=np.int64)
    for k, xn in enumerate(x):
        index = 0
        for _ in range(1, ap.num_levels):
            if t[index] > xn:
                index = int(t[index + child_index])
            else:
                index = int(t[index + child_index + 1])
        out[k] = index
    return out


def leaf_offsets_in_order(ap):
    """Element offsets of the leaves (self-looping nodes) in preorder = left-to-right order."""
    n = int(ap.max_order)
    stride = n + 6
    t = ap.tree_1d
    offs = []
    for node in range(len(t) // stride):
        o = node * stride
        if int(t[o + n + 4]) == o and int(t[o + n + 5]) == o:
            offs.append(o)
    return offs


def make_points(ap, breaks, rng, total, clustered=False):
    dt = ap.dtype
    lb, ub = dt(ap.lower_bound), dt(ap.upper_bound)
    b = np.asarray(breaks, dtype=dt)
    inf = dt(np.inf)
    special = [b, np.nextafter(b, -inf), np.nextafter(b, inf)]
    span = ub - lb
    special.append(np.array([lb - span, lb - dt(1e-3) * span, ub + dt(1e-3) * span, ub + span,
                             inf, -inf, np.nan, 0.0, -0.0], dtype=dt))
    pts = np.concatenate(special).astype(dt)
    nrand = max(total - len(pts), 256)
    if clustered:
        # BASELINE config 4: half uniform, half clustered geometrically close to breakpoints
        nu = nrand // 2
        u = rng.uniform(float(lb), float(ub), nu)
        k = rng.integers(0, len(b), nrand - nu)
        j = rng.integers(10, 41, nrand - nu)
        sgn = rng.choice([-1.0, 1.0], nrand - nu)
        c = b[k].astype(np.float64) + sgn * np.abs(rng.standard_normal(nrand - nu)) * 2.0 ** (-j)
        r = np.concatenate([u, c])
    else:
        r = rng.uniform(float(lb), float(ub), nrand)
    r = r.astype(dt)
    r = np.where(r >= ub, np.nextafter(ub, -inf), r)
    pts = np.concatenate([pts, r.astype(dt)])
    rng.shuffle(pts)
    return pts.astype(dt)


def freeze(name, ap, source, func_name, rng, total=2048, clustered=False, extra_x=None,
           allowed_error=None):
    n = int(ap.max_order)
    dt = np.dtype(ap.dtype).type
    tree_1d = np.asarray(ap.tree_1d, dtype=dt)
    leaf_offs = leaf_offsets_in_order(ap)
    a = np.array([tree_1d[o + n + 2] for o in leaf_offs], dtype=dt)
    b = np.array([tree_1d[o + n + 3] for o in leaf_offs], dtype=dt)
    # contiguity of the leaves -- a property every table must have
    assert np.all(a[1:] == b[:-1]), name
    assert a[0] == dt(ap.lower_bound) and b[-1] == dt(ap.upper_bound), name
    # cross-check against the SoA arrays the reference stores
    assert np.array_equal(np.asarray(ap.interval_a, dtype=dt), a), name
    assert np.array_equal(np.asarray(ap.interval_b, dtype=dt), b), name
    breaks = np.concatenate([a, b[-1:]])

    x = make_points(ap, breaks, rng, total, clustered)
    if extra_x is not None:
        x = np.concatenate([x, np.asarray(extra_x, dtype=dt)])
    with np.errstate(all="ignore"):
        y_ref = np.array(ap.evaluate(x))            # THE reference evaluation
    assert y_ref.dtype == dt, (name, y_ref.dtype)
    offs = ref_bst_leaf_offsets(ap, x)
    pos = {o: k for k, o in enumerate(leaf_offs)}
    idx_ref = np.array([pos[o] for o in offs], dtype=np.int32)
    with np.errstate(all="ignore"):
        f_true = np.asarray(FUNCS[func_name](x.astype(np.float64)), dtype=np.float64)
        dom = np.linspace(float(ap.lower_bound), float(ap.upper_bound), 200001)
        fmax = float(np.max(np.abs(FUNCS[func_name](dom))))

    err = allowed_error if allowed_error is not None else getattr(ap, "error", None)
    meta = dict(name=name, source=source, function=func_name, basis=str(ap.basis), order=n,
                num_levels=int(ap.num_levels), dtype=np.dtype(dt).name,
                lower_bound=float(ap.lower_bound), upper_bound=float(ap.upper_bound),
                optimizations=list(ap.optimizations), n_leaves=len(leaf_offs),
                allowed_error=None if err is None else float(err), fmax=fmax,
                dtype_name=str(getattr(ap, "dtype_name", "")), n_points=int(len(x)))
    cmap = np.asarray(getattr(ap, "cmap", []))
    np.savez_compressed(
        os.path.join(OUT, name + ".npz"),
        meta=np.array(json.dumps(meta)),
        tree_1d=tree_1d,
        interval_a=np.asarray(ap.interval_a, dtype=np.float64),
        interval_b=np.asarray(ap.interval_b, dtype=np.float64),
        coeff=np.asarray(ap.coeff, dtype=np.float64),
        map=np.asarray(ap.map).astype(np.int64),
        cmap=cmap.astype(np.int64) if cmap.size else np.zeros(0, np.int64),
        x=x, y_ref=y_ref, idx_ref=idx_ref, f_true=f_true)
    print("%-34s %-9s o%-2d %-7s L=%-4d levels=%-3d pts=%d" % (
        name, ap.basis, n, meta["dtype"], len(leaf_offs), ap.num_levels, len(x)))
    return meta


# ---------------------------------------------------------------- job lists
EPS64 = float(np.finfo(np.float64).eps)

GEN = [
    # name, (a, b, func, order, err, basis, adapt_type, dtype, accurate, opts)
    ("gen__cfg1_sinsin_cheb_o3_f64", (0, 10, "sinsin", 3, 1e-6, "chebyshev", "Remez", "64", True, ["arrays", "map"])),
    ("gen__cfg3_j0_cheb_o13_f64", (0, 20, "j0", 13, 100 * EPS64, "chebyshev", "Remez", "64", True, ["arrays", "map", "random"])),
    ("gen__j0_mono_o5_f64", (0, 20, "j0", 5, 1e-5, "monomial", "Trivial", "64", True, ["arrays", "map"])),
    ("gen__j0_mono_o13_f64", (0, 20, "j0", 13, 1e-9, "monomial", "Trivial", "64", True, ["arrays", "map"])),
    ("gen__j0_leg_o7_f32", (0, 20, "j0", 7, 1e-5, "legendre", "Trivial", "32", True, ["arrays", "map"])),
    ("gen__j0_leg_o7_f64", (0, 20, "j0", 7, 1e-9, "legendre", "Trivial", "64", True, ["arrays", "map"])),
    ("gen__j0_leg_o6_remez_f64", (0, 20, "j0", 6, 1e-8, "legendre", "Remez", "64", True, ["arrays", "map"])),
    ("gen__j0_cheb_o10_f32", (0, 20, "j0", 10, 1e-6, "chebyshev", "Trivial", "32", True, ["arrays", "map"])),
    # the reference's own accuracy/exactness tests (tests/performance_tests.py:112-162)
    ("gen__poly4_mono_o4_f64", (-10, 10, "poly4", 4, 1e-9, "monomial", "Trivial", "64", True, ["arrays", "map"])),
    ("gen__poly6_mono_o6_f64", (-10, 10, "poly6", 6, 1e-9, "monomial", "Trivial", "64", True, ["arrays", "map"])),
    ("gen__line1_mono_o1_f64", (-10, 10, "line1", 1, 1e-9, "monomial", "Trivial", "64", True, ["arrays", "map"])),
    ("gen__sinsin_mono_o10_f64", (-10, 10, "sinsin", 10, 1e-6, "monomial", "Trivial", "64", True, ["arrays", "map"])),
    ("gen__cossin_cheb_o10_f64", (-10, 10, "cossin", 10, 1e-9, "chebyshev", "Trivial", "64", True, ["arrays", "map"])),
    ("gen__sinsin_leg_o10_f64", (-10, 10, "sinsin", 10, 1e-9, "legendre", "Trivial", "64", True, ["arrays", "map"])),
    ("gen__sinsin_cheb_o2_f32", (-3, 3, "sinsin", 2, 1e-4, "chebyshev", "Trivial", "32", True, ["arrays", "map"])),
]


def matrix_jobs():
    """One reference-fitted table for EVERY (basis, order 3..13, dtype) the kernels are specialised
    on (BASELINE configs[4]: "32/64-bit, orders 3-13, monomial/Chebyshev/Legendre";
    adapt.py:97-139 defines all three bases for any order; tests/performance_tests.py:132-162 sweeps
    the bases).  Chebyshev / Legendre: J0 on [0,20]; monomial: sin(sin x) on [-2,2] (a monomial
    basis on the raw x of [0,20] is singular at these orders, adapt.py:213-219).  All 66 construct."""
    out = []
    for dtype, err in (("32", 1e-5), ("64", 1e-9)):
        for basis, short in (("chebyshev", "cheb"), ("legendre", "leg"), ("monomial", "mono")):
            for order in range(3, 14):
                fn, a, b = ("sinsin", -2, 2) if basis == "monomial" else ("j0", 0, 20)
                out.append(("mx__%s_o%d_f%s" % (short, order, dtype),
                            (a, b, fn, order, err, basis, "Trivial", dtype, True, ["arrays", "map"])))
    return out


# the reference's missing saved tables (/root/reference/.MISSING_LARGE_BLOBS), refitted with the
# recipe that made them (run_test.py:458-488 save_approximations: Chebyshev, Remez, dtype 64,
# J0 on [0,20] for approx/, [1,21] for approximations/).  The `ci` twins differ from their
# non-`ci` twin only in the "calc intervals" flag (packed map for the ISPC variant), which also
# caps the tree depth at 14 (adapt.py:69-72); they are refitted with that flag.
LARGE = [
    ("approx__64o3-p2.22044604925e-14", (0, 20, 3, 2.22044604925e-14, ["arrays", "map", "random"])),
    ("approx__64o3-p2.22044604925e-15", (0, 20, 3, 2.22044604925e-15, ["arrays", "map", "random"])),
    ("approx__ci64o3-p2.22044604925e-14", (0, 20, 3, 2.22044604925e-14, ["arrays", "map", "calc intervals", "random"])),
    ("approx__ci64o3-p2.22044604925e-15", (0, 20, 3, 2.22044604925e-15, ["arrays", "map", "calc intervals", "random"])),
    ("approximations__64o3-p2.22044604925e-14", (1, 21, 3, 2.22044604925e-14, ["arrays", "map", "calc intervals", "random"])),
    # not in .MISSING_LARGE_BLOBS: the same recipe one precision step further on the approximations/
    # domain, for a second reference-fitted table beyond 10^4 leaves
    ("approximations__64o3-p2.22044604925e-15", (1, 21, 3, 2.22044604925e-15, ["arrays", "map", "random"])),
]


def fit(a, b, fname, order, err, basis, atype, dtype, acc, opts):
    with linspace_int_shim(), quiet():
        return ref_api.make_interpolant(a, b, FUNCS[fname], order, err, basis, atype, dtype, acc, list(opts))


def job_pickle(d, fn, seed):
    ap = ref_api.load_from_file(os.path.join(REF, d, fn))
    perr = float(fn.split("-p")[1])
    return freeze("%s__%s" % (d, fn), ap, "pickle:%s/%s" % (d, fn), "j0", np.random.default_rng(seed),
                  allowed_error=perr)


def job_gen(name, spec, seed, total=2048):
    a, b, fname, order, err, basis, atype, dtype, acc, opts = spec
    try:
        ap = fit(*spec)
    except Exception as e:                      # a fit the reference cannot produce
        print("SKIP %s: %s: %s" % (name, type(e).__name__, str(e)[:80]))
        return None
    extra = None
    if name.startswith("gen__cfg1"):
        # BASELINE config 1: 2**20 linspace points; freeze a strided sample of them
        extra = np.linspace(0, 10, 2**20, endpoint=False)[::128]
    return freeze(name, ap, "make_interpolant%r" % (spec,), fname, np.random.default_rng(seed), total=total,
                  extra_x=extra, allowed_error=err)


def job_cfg4(name, dtype, err, seed):
    with linspace_int_shim(), quiet():
        it = ref_adapt.Interpolant(f1, 7, err, "legendre", dtype, False)
        it.run_adapt(0, 10, "Trivial")
    ap = approximator_from_raw_tree(it)
    return freeze(name, ap, "Interpolant(f1,7,%g,'legendre','%s',False).run_adapt(0,10,'Trivial')" % (err, dtype),
                  "f1", np.random.default_rng(seed), total=6144, clustered=True, allowed_error=err)


def job_large(name, spec, seed):
    """A missing saved table, refitted by the reference's fitter and flattened from its raw tree
    (the Approximator constructor trips approximator.py:178 on these; with
    "calc intervals" the fitter itself gives up past depth 14 -- recorded as a SKIP)."""
    a, b, order, err, opts = spec
    sys.setrecursionlimit(200000)
    try:
        with linspace_int_shim(), quiet():
            it = ref_adapt.Interpolant(f, order, err, "chebyshev", "64", True, optimizations=list(opts))
            it.run_adapt(a, b, "Remez")
    except Exception as e:
        print("SKIP %s: %s: %s" % (name, type(e).__name__, str(e)[:80].replace("\n", " ")))
        return None
    ap = approximator_from_raw_tree(it)
    ap.optimizations = list(opts)
    return freeze(name, ap, "refit of .MISSING_LARGE_BLOBS: Interpolant(j0,%d,%g,'chebyshev','64',True,%r).run_adapt(%g,%g,'Remez')"
                  % (order, err, opts, a, b), "j0", np.random.default_rng(seed), total=8192, allowed_error=err)


def job_trim(name, seed):
    """The `trim_data` tree layout (approximator.py:202-206,232-233): leaves
    [mid, self, self, c0..cn, a, b], inner nodes [mid, left NODE index, right NODE index].  The
    reference's tree_1d walk (approximator.py:275-290) reads those node indices as element offsets,
    so Approximator.evaluate is not usable on it; y_ref is Approximator.evaluate_tree
    (approximator.py:244-271), the walk over the node objects of the same tree (fp64 table: the
    node coefficients ARE the tree_1d values).  The BST leaf is restated on the node objects."""
    spec = (0, 20, "j0", 6, 1e-6, "chebyshev", "Remez", "64", True, ["arrays", "map", "trim_data"])
    ap = fit(*spec)
    rng = np.random.default_rng(seed)
    n = int(ap.max_order)
    dt = np.float64
    a = np.asarray(ap.interval_a, dtype=dt)
    b = np.asarray(ap.interval_b, dtype=dt)
    breaks = np.concatenate([a, b[-1:]])
    x = make_points(ap, breaks, rng, 2048)
    with np.errstate(all="ignore"):
        y_ref = np.array(ap.evaluate_tree(x), dtype=dt)
        f_true = np.asarray(FUNCS["j0"](x.astype(np.float64)), dtype=np.float64)
    idx_ref = np.empty(len(x), dtype=np.int32)
    leaves = []

    def walk(nd):
        if nd.left == 0 and nd.right == 0:
            leaves.append(nd)
        else:
            walk(nd.left)
            walk(nd.right)
    walk(ap.tree.root)
    pos = {id(nd): k for k, nd in enumerate(leaves)}
    for k, xn in enumerate(x):
        node = ap.tree.root
        for _ in range(1, int(ap.num_levels)):                 # approximator.py:250-256
            if node.data[0] > xn:
                node = node.left if (node.left != 0) else node
            else:
                node = node.right if (node.right != 0) else node
        idx_ref[k] = pos[id(node)]
    dom = np.linspace(0.0, 20.0, 200001)
    meta = dict(name=name, source="make_interpolant%r; y_ref = evaluate_tree" % (spec,), function="j0",
                basis="chebyshev", order=n, num_levels=int(ap.num_levels), dtype="float64",
                lower_bound=0.0, upper_bound=20.0, optimizations=list(ap.optimizations),
                n_leaves=len(leaves), allowed_error=1e-6, fmax=float(np.max(np.abs(f(dom)))),
                dtype_name="double", n_points=int(len(x)))
    np.savez_compressed(
        os.path.join(OUT, name + ".npz"), meta=np.array(json.dumps(meta)),
        tree_1d=np.asarray(ap.tree_1d, dtype=dt), interval_a=a, interval_b=b,
        coeff=np.asarray(ap.coeff, dtype=np.float64), map=np.asarray(ap.map).astype(np.int64),
        cmap=np.asarray(ap.cmap).astype(np.int64), x=x, y_ref=y_ref, idx_ref=idx_ref, f_true=f_true)
    print("%-34s trim_data layout, L=%d, tree_1d %d elements" % (name, len(leaves), len(ap.tree_1d)))
    return meta


def all_jobs():
    """[(name, thunk)] in a fixed order; every job has its own fixed seed."""
    jobs = []
    seed = 1000
    for d in ("approx", "approximations"):
        for fn in sorted(os.listdir(os.path.join(REF, d))):
            seed += 1
            jobs.append(("%s__%s" % (d, fn), lambda d=d, fn=fn, seed=seed: job_pickle(d, fn, seed)))
    for name, spec in GEN:
        seed += 1
        jobs.append((name, lambda name=name, spec=spec, seed=seed: job_gen(name, spec, seed)))
    for name, dtype, err in (("gen__cfg4_f1_leg_o7_f32", "32", 1e-4), ("gen__cfg4_f1_leg_o7_f64", "64", 1e-6)):
        seed += 1
        jobs.append((name, lambda name=name, dtype=dtype, err=err, seed=seed: job_cfg4(name, dtype, err, seed)))
    seed = 2000
    for name, spec in matrix_jobs():
        seed += 1
        jobs.append((name, lambda name=name, spec=spec, seed=seed: job_gen(name, spec, seed, total=1024)))
    jobs.append(("trim__j0_cheb_o6_f64", lambda: job_trim("trim__j0_cheb_o6_f64", 3001)))
    seed = 4000
    for name, spec in LARGE:
        seed += 1
        jobs.append(("large__" + name, lambda name=name, spec=spec, seed=seed: job_large("large__" + name, spec, seed)))
    return jobs


def _run_job_by_name(name):
    return dict(all_jobs())[name]()


def main():
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", nargs="*", default=None,
                    help="regenerate only the tables whose name starts with one of these prefixes "
                         "(the others and their MANIFEST entries are kept)")
    ap.add_argument("--list", action="store_true")
    ap.add_argument("--jobs", type=int, default=1, help="worker processes (every job has its own fixed seed)")
    args = ap.parse_args()
    os.makedirs(OUT, exist_ok=True)
    jobs = all_jobs()
    if args.list:
        print("\n".join(n for n, _ in jobs))
        return
    mpath = os.path.join(HERE, "MANIFEST.json")
    manifest = {}
    if args.only is not None and os.path.exists(mpath):
        with open(mpath) as fh:
            manifest = {m["name"]: m for m in json.load(fh)}
    selected = [(n, j) for n, j in jobs if args.only is None or any(n.startswith(p) for p in args.only)]
    if args.only is None:
        for fn in os.listdir(OUT):
            if fn.endswith(".npz"):
                os.remove(os.path.join(OUT, fn))
    for name, _ in selected:
        path = os.path.join(OUT, name + ".npz")
        if os.path.exists(path):
            os.remove(path)
        manifest.pop(name, None)
    if args.jobs > 1:
        import multiprocessing as mp
        with mp.get_context("fork").Pool(args.jobs) as pool:
            metas = pool.map(_run_job_by_name, [n for n, _ in selected], chunksize=1)
    else:
        metas = [job() for _, job in selected]
    for (name, _), meta in zip(selected, metas):
        if meta is not None:
            manifest[name] = meta
    order = {n: k for k, (n, _) in enumerate(jobs)}
    out = sorted(manifest.values(), key=lambda m: order.get(m["name"], 1 << 30))
    with open(mpath, "w") as fh:
        json.dump(out, fh, indent=1)
    print("wrote %d tables (%d in the manifest)" % (len(selected), len(out)))


if __name__ == "__main__":
    main()
