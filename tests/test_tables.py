"""Host logic: flattening the reference's table layouts, loading saved pickles, sharding.  CPU only."""
import io
import os
import pickle
import sys
import types

import numpy as np
import pytest

import adaptive_interpolation_b200.adaptive_interpolation as adapt_i
from adaptive_interpolation_b200 import approximator as app
from adaptive_interpolation_b200 import sharding, tables
from conftest import REFERENCE_ROOT, all_golden_names, golden, has_reference

NAMES = all_golden_names()


@pytest.mark.parametrize("name", NAMES)
def test_flatten_golden(name):
    g = golden(name)
    f = tables.flatten(g)
    assert f.n_leaves == g.meta["n_leaves"]
    assert f.breaks.dtype == g.dtype and f.coef.dtype == g.dtype
    assert np.all(np.diff(f.breaks) > 0)
    assert f.breaks[0] == g.lower_bound and f.breaks[-1] == g.upper_bound
    # run_test.py:242-246: SoA arrays cast to the table dtype are what the kernels consume
    assert np.array_equal(f.breaks[:-1], np.asarray(g.interval_a, dtype=g.dtype))
    assert np.array_equal(f.coef.ravel(), np.asarray(g.coeff, dtype=g.dtype))


def test_flatten_rejects_bad_trees():
    g = golden("approx__32o6-p1.19209289551e-06")
    t = g.tree_1d.copy()
    with pytest.raises(ValueError):
        tables.flatten_tree_1d(t[:-1], g.max_order)               # not a whole number of nodes
    bad = t.copy()
    bad[g.max_order + 4] = 10 ** 6                                 # child offset out of range
    with pytest.raises(ValueError):
        tables.flatten_tree_1d(bad, g.max_order)
    cyc = t.copy()
    stride = g.max_order + 6
    cyc[stride + g.max_order + 4] = 0                              # node 1's left child -> root
    with pytest.raises(ValueError):
        tables.flatten_tree_1d(cyc, g.max_order)


def test_flatten_detects_soa_mismatch():
    g = golden("approx__64o7-p1.19209289551e-06")
    ap = app.Approximator.from_record(g)
    ap.coeff = list(ap.coeff)
    ap.coeff[3] += 1.0
    with pytest.raises(ValueError):
        tables.flatten(ap)


def test_trim_data_layout_flattens_to_the_same_leaves():
    """approximator.py:202-206,232-233: a table built with optimizations=[..., "trim_data"] stores
    inner nodes as [mid, left node, right node] and leaves as [mid, self, self, c.., a, b].
    flatten() scans it in preorder; the result equals the SoA arrays the reference stored beside it."""
    g = golden("trim__j0_cheb_o6_f64")
    assert "trim_data" in g.optimizations
    f = tables.flatten(g)
    assert np.array_equal(f.breaks[:-1], np.asarray(g.interval_a)) and f.breaks[-1] == g.interval_b[-1]
    assert np.array_equal(f.coef.ravel(), np.asarray(g.coeff))
    n_inner = f.n_leaves - 1
    assert len(g.tree_1d) == 3 * n_inner + (g.max_order + 6) * f.n_leaves


def test_trim_data_layout_rejects_malformed_trees():
    g = golden("trim__j0_cheb_o6_f64")
    t = np.asarray(g.tree_1d).copy()
    with pytest.raises(ValueError):
        tables.flatten_tree_1d_trim(t[:-1], g.max_order)             # ends inside a leaf
    bad = t.copy()
    bad[1] = 5.0                                                     # root's left child is not node 1
    with pytest.raises(ValueError):
        tables.flatten_tree_1d_trim(bad, g.max_order)
    with pytest.raises(ValueError):
        tables.flatten_tree_1d_trim(t[3:], g.max_order)              # root removed: not one full tree


def test_single_leaf_and_order_one():
    g = golden("gen__line1_mono_o1_f64")
    f = tables.flatten(g)
    assert f.n_leaves == 1 and f.order == 1 and f.coef.shape == (1, 2)


def test_flatten_raw_node_tree_with_pruning():
    """adapt.py Tree/Node objects -> FlatTable, including a dataless pair of children
    (accurate=False early return, adapt.py:206-227): their parent becomes the leaf."""
    def node(a, b, c, children=None):
        n = app.Node()
        n.data = [(a + b) / 2.0, list(c), [a, b], 0.0]
        if children:
            n.left, n.right = children
        return n
    dead_l, dead_r = app.Node(), app.Node()                  # data == 0
    leaf0 = node(0.0, 1.0, [1, 2, 3])
    leaf1 = node(1.0, 1.5, [4, 5, 6], (dead_l, dead_r))      # children never fitted
    leaf2 = node(1.5, 2.0, [7, 8, 9])
    root = node(0.0, 2.0, [0, 0, 0], (leaf0, node(1.0, 2.0, [0, 0, 0], (leaf1, leaf2))))
    f = tables.flatten_nodes(root, "legendre", 2, np.float32)
    assert np.array_equal(f.breaks, np.array([0, 1, 1.5, 2], dtype=np.float32))
    assert np.array_equal(f.coef, np.array([[1, 2, 3], [4, 5, 6], [7, 8, 9]], dtype=np.float32))
    with pytest.raises(ValueError):
        tables.flatten_nodes(node(0.0, 1.0, [1, 2]), "legendre", 2, np.float32)      # wrong order


@pytest.mark.parametrize("name", [n for n in NAMES if n.startswith("approx")])
def test_packed_map_describes_the_flattened_leaves(name):
    """The reference's packed grid maps ("calc intervals", approximator.py:155-188 and the int64
    `map` of the ci* files) decode to exactly the leaves we flatten from tree_1d."""
    g = golden(name)
    try:
        leaf, l, n_cells = tables.decode_packed_map(g)
    except ValueError:
        pytest.skip("no packed map in this schema version")
    f = tables.flatten(g)
    cells = len(leaf)
    w = (float(g.upper_bound) - float(g.lower_bound)) / cells
    assert leaf.min() == 0 and leaf.max() == f.n_leaves - 1
    for c in range(cells):
        i = int(leaf[c])
        assert l[c] <= c < l[c] + n_cells[c]
        assert abs(float(f.breaks[i]) - (float(g.lower_bound) + l[c] * w)) <= 1e-12 * max(1.0, abs(float(f.breaks[i])))
        assert abs(float(f.breaks[i + 1]) - (float(g.lower_bound) + (l[c] + n_cells[c]) * w)) <= 1e-12 * max(1.0, abs(float(f.breaks[i + 1])))


def _fake_reference_modules():
    """Modules named like the reference's so a pickle in the reference's schema can be produced
    without the reference (class identity in a pickle is just module + qualname)."""
    mods = {}
    for modname, names in (("adaptive_interpolation", []),
                           ("adaptive_interpolation.approximator", ["Approximator"]),
                           ("adaptive_interpolation.adapt", ["Interpolant", "Tree", "Node"])):
        m = types.ModuleType(modname)
        for n in names:
            cls = type(n, (object,), {})
            cls.__module__ = modname
            if n == "Interpolant":
                def basis_function(self, *a):
                    return None
                cls.basis_function = basis_function
            setattr(m, n, cls)
        mods[modname] = m
    return mods


def test_load_pickle_in_reference_schema_without_reference(tmp_path):
    """A V1-schema pickle loads into the stand-in classes when neither
    the reference package nor __main__.f exists."""
    g = golden("approx__32o6-p1.19209289551e-06")
    mods = _fake_reference_modules()
    saved = {k: sys.modules.get(k) for k in mods}
    sys.modules.update(mods)
    try:
        A = mods["adaptive_interpolation.approximator"].Approximator
        I = mods["adaptive_interpolation.adapt"].Interpolant
        T = mods["adaptive_interpolation.adapt"].Tree
        N = mods["adaptive_interpolation.adapt"].Node
        ap, it, tree, node = A(), I(), T(), N()
        tree.root, tree.size, tree.max_level = node, 1, 0
        node.parent = node.left = node.right = 0
        node.level, node.data = 0, [0.5, [1.0], [0.0, 1.0], 0.0]
        it.tree = tree
        ap.basis_function = it.basis_function          # bound method, as approximator.py:26
        ap.tree = tree
        for k in ("basis", "max_order", "num_levels", "dtype", "dtype_name", "lower_bound", "upper_bound",
                  "tree_1d", "interval_a", "interval_b", "coeff", "map", "optimizations"):
            setattr(ap, k, getattr(g, k))
        ap.code, ap.size, ap.vector_width = 0, None, None
        blob = pickle.dumps(ap, pickle.HIGHEST_PROTOCOL)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    path = tmp_path / "32o6"
    path.write_bytes(blob)
    loaded = adapt_i.load_from_file(str(path))
    assert type(loaded).__name__ == "Approximator"
    f = tables.flatten(loaded)
    assert f.n_leaves == 15 and f.order == 6
    # and it can be written back (no device handles on the object)
    out = tmp_path / "again"
    adapt_i.write_to_file(str(out), loaded)
    assert tables.flatten(adapt_i.load_from_file(str(out))).n_leaves == 15


def test_load_from_file_error_message(tmp_path):
    p = tmp_path / "garbage"
    p.write_bytes(b"not a pickle")
    with pytest.raises(Exception) as ei:
        adapt_i.load_from_file(str(p))
    assert "Error loading instance from" in str(ei.value)          # adaptive_interpolation.py:104


@pytest.mark.skipif(not has_reference(), reason="reference tree not mounted")
def test_load_every_saved_pickle():
    """All 27 saved tables of the reference load (three schema versions) and flatten to exactly
    what the golden fixtures froze."""
    count = 0
    for d in ("approx", "approximations"):
        for fn in sorted(os.listdir(os.path.join(REFERENCE_ROOT, d))):
            ap = adapt_i.load_from_file(os.path.join(REFERENCE_ROOT, d, fn))
            f = tables.flatten(ap)
            g = tables.flatten(golden("%s__%s" % (d, fn)))
            assert np.array_equal(f.breaks, g.breaks) and np.array_equal(f.coef, g.coef)
            assert f.basis == g.basis and f.order == g.order and f.dtype == g.dtype
            # the pickled fitter tree (adapt.py Node objects) flattens to the same table
            fn_ = tables.flatten_nodes(ap.tree.root, ap.basis, ap.max_order, f.dtype)
            assert np.array_equal(fn_.breaks, f.breaks) and np.array_equal(fn_.coef, f.coef)
            count += 1
    assert count == 27


# ------------------------------------------------------------------ sharding
@pytest.mark.parametrize("n,w", [(0, 1), (1, 1), (5, 2), (2 ** 20, 8), (2 ** 30, 8), (2 ** 28 + 3, 4),
                                 (1000, 8), (7, 8)])
def test_shard_bounds_cover(n, w):
    b = sharding.shard_bounds(n, w)
    assert len(b) == w and b[0][0] == 0 and b[-1][1] == n
    for (lo, hi), (lo2, _) in zip(b, b[1:]):
        assert hi == lo2 and lo <= hi
    if n >= 512 * w:
        assert all(lo % 512 == 0 for lo, _ in b)               # shards keep 16-byte alignment
        sizes = [hi - lo for lo, hi in b]
        assert max(sizes) - min(sizes) <= 512 * w


def test_shard_bounds_errors():
    with pytest.raises(ValueError):
        sharding.shard_bounds(10, 0)
    with pytest.raises(ValueError):
        sharding.shard_bounds(-1, 2)
