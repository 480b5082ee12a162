"""Pin the oracle: the CPU restatements under oracle/ against the frozen outputs of the REAL
reference evaluator (Approximator.evaluate, approximator.py:273-296) for every golden table.
CPU only."""
import numpy as np
import pytest

import oracle
from oracle import oracle_np
from adaptive_interpolation_b200 import tables
from conftest import all_golden_names, golden, value_tolerance

NAMES = all_golden_names()
# oracle-vs-reference: both evaluate in the same operation order; only np.dot's BLAS summation
# order and numpy's pow differ -> a couple of eps of sum|c_i B_i| (measured max 2.1)
ORACLE_ULPS = 4
# tables whose tree_1d is in the standard layout the tree-walking restatements follow; the one
# trim_data fixture (whose links the reference's own walk cannot follow) is covered through
# tables.flatten + the numpy oracle below and in tests/test_tables.py
TREE_NAMES = [n for n in NAMES if not n.startswith("trim__")]


def leaf_ordinals(g, offsets):
    stride = g.max_order + 6
    nodes = len(g.tree_1d) // stride
    leaf_offs = [k * stride for k in range(nodes)
                 if int(g.tree_1d[k * stride + g.max_order + 4]) == k * stride]
    pos = {o: i for i, o in enumerate(leaf_offs)}
    return np.array([pos[int(o)] for o in offsets], dtype=np.int32)


@pytest.mark.parametrize("name", TREE_NAMES)
def test_c_oracle_matches_reference(name):
    g = golden(name)
    y, off, cond = oracle.c_eval_tree(g.tree_1d, g.max_order, g.num_levels, g.basis, g.x)
    assert np.array_equal(leaf_ordinals(g, off), g.idx_ref)          # leaf choice: bit-exact
    assert np.array_equal(np.isfinite(y), np.isfinite(g.y_ref))
    fin = np.isfinite(g.y_ref)
    err = np.abs(y[fin].astype(np.float64) - g.y_ref[fin].astype(np.float64))
    assert np.all(err <= value_tolerance(g.dtype, cond[fin], ORACLE_ULPS))


@pytest.mark.parametrize("name", NAMES)
def test_numpy_oracle_matches_reference(name):
    g = golden(name)
    f = tables.flatten(g)
    y, idx, cond = oracle_np.evaluate(f.breaks, f.coef, f.basis, f.order, g.x, want_cond=True)
    assert y.dtype == g.y_ref.dtype
    assert np.array_equal(idx, g.idx_ref)                            # searchsorted == BST walk
    fin = np.isfinite(g.y_ref)
    assert np.array_equal(np.isfinite(y), np.isfinite(g.y_ref))
    err = np.abs(y[fin].astype(np.float64) - g.y_ref[fin].astype(np.float64))
    assert np.all(err <= value_tolerance(g.dtype, cond[fin], ORACLE_ULPS))


@pytest.mark.parametrize("name", ["approx__32o3-p1.19209289551e-06", "gen__cfg4_f1_leg_o7_f64",
                                  "approximations__64o5-p2.22044604925e-14"])
def test_literal_walk_equals_searchsorted(name):
    """approximator.py:285-290 walked literally vs the closed form, on adversarial points."""
    g = golden(name)
    f = tables.flatten(g)
    b = f.breaks
    inf = g.dtype(np.inf)
    x = np.concatenate([b, np.nextafter(b, -inf), np.nextafter(b, inf),
                        np.array([-inf, inf, np.nan, 0.0, -0.0], dtype=g.dtype)])
    offs = oracle_np.walk_tree_1d(g.tree_1d, g.max_order, g.num_levels, x)
    assert np.array_equal(leaf_ordinals(g, offs), oracle_np.leaf_index(b, x))


@pytest.mark.parametrize("name", [n for n in TREE_NAMES if "cheb" in n or "approx" in n])
def test_reference_generated_kernel(name):
    """oracle/_ref: the reference's own gen_cheb text compiled as C (oracle/build_ref.py) agrees
    with the reference's Python evaluator.  Its rescale is (2./(b-a))*(x-a)-1 (generate.py:83),
    not adapt.py:130's, so the tolerance is wider (measured max 8.3)."""
    if not oracle.ref_available():
        pytest.skip("oracle/_ref not built (needs /root/reference; python oracle/build_ref.py)")
    g = golden(name)
    if name not in oracle.ref_manifest():
        pytest.skip("not a chebyshev table")
    y = oracle.ref_eval(name, g.tree_1d, g.x)
    _, _, cond = oracle.c_eval_tree(g.tree_1d, g.max_order, g.num_levels, g.basis, g.x)
    fin = np.isfinite(g.y_ref) & np.isfinite(y)
    err = np.abs(y[fin].astype(np.float64) - g.y_ref[fin].astype(np.float64))
    assert np.all(err <= value_tolerance(g.dtype, cond[fin], 32))


def test_reference_error_bound_holds_in_golden():
    """The frozen reference outputs respect the interpolants' stated error bounds (adapt.py:180-198:
    max|p - f| / max_domain|f| <= allowed_error), so the GPU tests can hold the GPU to them."""
    for name in NAMES:
        g = golden(name)
        if g.allowed_error is None or name.startswith("gen__cfg4"):
            continue          # cfg4 is built with accurate=False (depth cap may miss the bound)
        dom = (g.x >= g.lower_bound) & (g.x < g.upper_bound) & np.isfinite(g.f_true)
        rel = np.max(np.abs(g.y_ref[dom].astype(np.float64) - g.f_true[dom])) / g.fmax
        assert rel <= g.allowed_error * 1.5 + 8 * np.finfo(g.dtype).eps, (name, rel)


def test_parity_matrix_is_complete():
    """north_star: kernels specialised on basis x order 3..13 x dtype.  Every one of the 66
    combinations has a table fitted by the reference's own fitter and frozen with the reference's
    own evaluation (tests/golden/make_golden.py matrix_jobs), so every parametrised parity test
    (CPU oracle here, CUDA path in test_gpu_parity.py) runs on all of them."""
    have = set()
    for name in NAMES:
        g = golden(name)
        have.add((g.basis, int(g.max_order), np.dtype(g.dtype).name))
    want = {(b, o, d) for b in ("chebyshev", "legendre", "monomial") for o in range(3, 14)
            for d in ("float32", "float64")}
    assert not (want - have), sorted(want - have)
