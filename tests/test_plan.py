"""The table planner (host side of ai_table_create) through ai_table_plan: lookup grid, record
layout and residency mode for every golden table, without a GPU.  What the plan promises is
checked on the device by tests/test_gpu_parity.py; here the decisions themselves are pinned."""
import ctypes

import numpy as np
import pytest

from adaptive_interpolation_b200 import _cabi, build, tables
from conftest import all_golden_names, golden

NAMES = all_golden_names()
SAVED = [n for n in NAMES if n.startswith("approx")]


@pytest.fixture(scope="module", autouse=True)
def lib():
    build.build()
    return _cabi.lib()


def plan(name, **kw):
    return tables.plan(tables.flatten(golden(name)), **kw)


@pytest.mark.parametrize("name", NAMES)
def test_plan_is_consistent(name):
    f = tables.flatten(golden(name))
    info = tables.plan(f)
    assert info["n_leaves"] == f.n_leaves and info["order"] == f.order and info["device"] == -1
    assert info["lower_bound"] == float(f.breaks[0]) and info["upper_bound"] == float(f.breaks[-1])
    assert 1 <= info["grid_cells"] <= (1 << 20)
    assert info["smem_bytes"] <= 227 * 1024
    assert info["residency"] in (0, 1, 2, 3, 4, 9, 10, 11)
    if info["residency"] == 11:                        # eight 128-bit copies of headerless cell records
        assert info["cell_records"] == 1 and info["search_steps"] == 0 and info["smem_resident"] == 1
        assert info["bank_copies"] == 8
    if info["residency"] == 10:                        # compact L2-resident layout: headerless 32-byte cell records
        assert info["cell_records"] == 1 and info["search_steps"] == 0 and info["smem_resident"] == 0
        assert info["record_stride"] * f.dtype.itemsize % 32 == 0 and info["smem_bytes"] <= 64 * 1024
    if info["residency"] == 4:
        assert f.n_leaves == 2 and info["smem_bytes"] == 0
    if info["residency"] in (1, 3):
        assert info["bank_copies"] >= 2 and f.n_leaves > (32 if f.dtype == np.float32 else 16)
    if info["residency"] == 9:                         # 8 copies of 16-byte padded records, 128-bit gathers
        assert info["bank_copies"] == 8 and info["header_packed"] == 0 and info["cell_records"] == 0
        assert info["record_stride"] * f.dtype.itemsize % 16 == 0
        assert f.n_leaves > (32 if f.dtype == np.float32 else 16)
    if info["lookup_mode"] in (2, 3):                  # rank from a register bitmap / from rank words
        assert info["search_steps"] == 0
    hdr = 0 if (f.basis == "monomial" or info["residency"] in (10, 11)) else {0: 2, 1: 1, 2: 0}[info["header_packed"]]
    assert info["record_stride"] >= f.order + 1 + hdr


def test_saved_tables_plan():
    """All 27 saved approximations: exact grid (no search), staged in shared memory, one-word header."""
    for name in SAVED:
        info = plan(name)
        assert info["search_steps"] == 0 and info["smem_resident"] == 1, (name, info)
        assert info["header_packed"] in (1, 2) or info["residency"] in (9, 11), (name, info)
        assert info["image_bytes"] <= 48 * 1024, (name, info)
    assert plan("approx__32o6-p1.19209289551e-06")["lookup_mode"] == 2          # BASELINE configs[1]
    assert plan("approx__32o6-p1.19209289551e-06")["residency"] == 0
    assert plan("approx__32o3-p1.19209289551e-06")["residency"] == 1            # 62 leaves > 32 banks
    assert plan("approx__64o13-p1.19209289551e-06")["residency"] == 4           # two leaves: select
    big = plan("approximations__64o5-p2.22044604925e-14")                       # 429 leaves
    # full per-bank replication (16 copies) does not fit: 8 copies of headerless 48-byte cell records
    # (the table is dyadic), 128-bit loads, the header from 4-bit levels
    assert big["residency"] == 11 and big["cell_records"] == 1 and big["smem_bytes"] <= 227 * 1024


def test_config_tables_plan():
    cfg3 = plan("gen__cfg3_j0_cheb_o13_f64")
    assert cfg3["lookup_mode"] == 2 and cfg3["header_packed"] == 2 and cfg3["residency"] == 0   # 8 equal leaves
    for name, p, res in (("gen__cfg4_f1_leg_o7_f32", 20, 1), ("gen__cfg4_f1_leg_o7_f64", 42, 9)):
        info = plan(name)                               # deep skewed trees (depth 21 / 43): the path lookup,
        assert info["lookup_mode"] == 4 and info["search_steps"] == 0           # 8 rows below level 3
        assert info["grid_cells"] == 8 and info["grid_log2_scale"] == p and info["header_packed"] == 0
        assert info["residency"] == res                 # fp32: per-bank copies fit; fp64: 8 x 128-bit


def test_config4_tables_without_the_path_lookup(monkeypatch):
    monkeypatch.setenv("AI_B200_NO_PATH", "1")
    for name, k, res in (("gen__cfg4_f1_leg_o7_f32", 4, 1), ("gen__cfg4_f1_leg_o7_f64", 6, 9)):
        info = plan(name)                               # grid + bounded search
        assert info["search_steps"] == k and info["lookup_mode"] == 1 and info["header_packed"] == 0
        assert info["residency"] == res and info["grid_cells"] == 80


def test_smaller_shared_memory_changes_the_residency(monkeypatch):
    name = "approximations__64o5-p2.22044604925e-14"
    assert plan(name, smem_cap_bytes=160 * 1024)["bank_copies"] == 4
    assert plan(name, smem_cap_bytes=20 * 1024)["residency"] == 10              # does not fit: L2-resident, compact (dyadic)
    monkeypatch.setenv("AI_B200_NO_COMPACT", "1")
    assert plan(name, smem_cap_bytes=20 * 1024)["residency"] == 2               # ... or grid table + padded records
    assert plan(name, smem_cap_bytes=20 * 1024)["header_packed"] == 0           # global layout keeps [a, s]


def test_hooks(monkeypatch):
    name = "approx__32o6-p1.19209289551e-06"
    monkeypatch.setenv("AI_B200_NO_PACKED_HDR", "1")
    assert plan(name)["header_packed"] == 0
    monkeypatch.delenv("AI_B200_NO_PACKED_HDR")
    monkeypatch.setenv("AI_B200_NO_UNIFORM_HDR", "1")
    assert plan("gen__cfg3_j0_cheb_o13_f64")["header_packed"] == 1
    monkeypatch.delenv("AI_B200_NO_UNIFORM_HDR")
    monkeypatch.setenv("AI_B200_NO_RANK", "1")
    assert plan(name)["lookup_mode"] == 0
    monkeypatch.delenv("AI_B200_NO_RANK")
    monkeypatch.setenv("AI_B200_MAX_CELLS", "8")                                 # grid too coarse: the tree is a path
    info = plan(name)                                                            # below level 4 (16 rows)
    assert info["lookup_mode"] == 4 and info["search_steps"] == 0 and info["grid_cells"] <= 64
    monkeypatch.setenv("AI_B200_NO_PATH", "1")                                   # ... or a search follows
    info = plan(name)
    assert info["grid_cells"] <= 8 and info["search_steps"] >= 1 and info["lookup_mode"] == 1
    monkeypatch.delenv("AI_B200_NO_PATH")
    monkeypatch.delenv("AI_B200_MAX_CELLS")
    for mode in (0, 1, 2, 9, 10, 11):
        monkeypatch.setenv("AI_B200_FORCE_MODE", str(mode))
        assert plan(name)["residency"] == mode
    monkeypatch.delenv("AI_B200_FORCE_MODE")
    monkeypatch.setenv("AI_B200_NO_VEC_MODE", "1")
    assert plan("approximations__64o5-p2.22044604925e-14")["residency"] == 3     # the partial-copy layout it replaces
    monkeypatch.delenv("AI_B200_NO_VEC_MODE")
    monkeypatch.setenv("AI_B200_NO_COMPACT", "1")
    info = plan("approximations__64o5-p2.22044604925e-14")                       # eight copies WITH [a, s] and rank words
    assert info["residency"] == 9 and info["bank_copies"] == 8 and info["lookup_mode"] == 3


def test_large_tables_plan():
    """Beyond 65536 leaves the grid grows with the table (one L2 access instead of a long search)."""
    for n_leaves, k0 in ((131072, True), (100000, False)):
        br = np.linspace(0.0, 16.0, n_leaves + 1)
        f = tables.FlatTable("chebyshev", 3, np.float64, br, np.zeros((n_leaves, 4)))
        info = tables.plan(f)
        # 131072 equal leaves on [0, 16] are dyadic: compact cell records; 100000 are not: grid + search
        assert info["residency"] == (10 if k0 else 2) and info["smem_resident"] == 0
        assert (info["search_steps"] == 0) == k0 and info["search_steps"] <= 2
        assert info["grid_cells"] >= n_leaves


def test_rejects_malformed_tables(lib):
    info = _cabi.TableInfo()
    br = np.array([0.0, 2.0, 1.0, 3.0])
    co = np.zeros(3 * 4)
    assert lib.ai_table_plan(0, 3, 1, 3, br.ctypes.data, co.ctypes.data, 0, ctypes.byref(info)) == 2     # AI_ERR_TABLE
    assert b"strictly increasing" in lib.ai_last_error()
    br = np.array([0.0, np.inf])
    assert lib.ai_table_plan(0, 3, 1, 1, br.ctypes.data, co.ctypes.data, 0, ctypes.byref(info)) == 2
    assert lib.ai_table_plan(0, 3, 1, 0, br.ctypes.data, co.ctypes.data, 0, ctypes.byref(info)) == 2
    assert lib.ai_table_plan(0, 3, 1, 1, None, None, 0, ctypes.byref(info)) == 1                         # AI_ERR_INVALID
    assert lib.ai_table_plan(0, 3, 1, 1, br.ctypes.data, co.ctypes.data, 0, None) == 1
    assert lib.ai_table_plan(9, 3, 1, 1, br.ctypes.data, co.ctypes.data, 0, ctypes.byref(info)) == 1


def _random_breaks(rng, dtype, kind):
    L = int(rng.integers(1, 400))
    if kind == "dyadic":                     # what the reference's bisection produces
        lo = float(rng.integers(-50, 50))
        width = float(rng.choice([1.0, 3.0, 20.0, 0.375, 1e3]))
        cuts = {0.0, 1.0}
        for _ in range(L):
            c = sorted(cuts)
            k = int(rng.integers(0, len(c) - 1))
            if c[k + 1] - c[k] > 2.0 ** -18:
                cuts.add((c[k] + c[k + 1]) / 2)
        b = lo + width * np.array(sorted(cuts))
    elif kind == "path":                     # bisection towards a few points (config 4 style): deep and skewed
        lo = float(rng.integers(-50, 50))
        width = float(rng.choice([1.0, 10.0, 20.0, 0.375, 6.0]))
        depth = 22 if dtype == np.float32 and abs(lo) + width > 16 else (18 if dtype == np.float32 else 44)
        targets = rng.uniform(0, 1, int(rng.integers(1, 5)))
        cuts = {0.0, 1.0}
        for t in targets:
            a, c = 0.0, 1.0
            for _ in range(int(rng.integers(3, depth))):
                m = (a + c) / 2
                cuts.add(m)
                a, c = (m, c) if t >= m else (a, m)
        b = lo + width * np.array(sorted(cuts))
    elif kind == "uniform":
        b = np.linspace(float(rng.integers(-5, 5)), float(rng.integers(6, 40)), L + 1)
    else:
        lo, hi = sorted(rng.uniform(-100, 100, 2))
        b = np.concatenate([[lo], np.sort(rng.uniform(lo, hi, L - 1)), [hi]]) if L > 1 else np.array([lo, hi])
    b = np.unique(b.astype(dtype))
    return b if len(b) >= 2 else np.array([0, 1], dtype=dtype)


SEED = 0


@pytest.mark.parametrize("kind", ["dyadic", "uniform", "random", "path"])
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_planner_invariants_on_random_tables(kind, dtype):
    """Whatever the breakpoints, the plan fits the device and its flags are consistent."""
    rng = np.random.default_rng([["dyadic", "uniform", "random", "path"].index(kind), np.dtype(dtype).itemsize, SEED])
    seen = set()
    for _ in range(60):
        b = _random_breaks(rng, dtype, kind)
        L = len(b) - 1
        order = int(rng.integers(0, 14))
        basis = ["chebyshev", "legendre", "monomial"][int(rng.integers(0, 3))]
        f = tables.FlatTable(basis, order, dtype, b, rng.standard_normal((L, order + 1)).astype(dtype))
        info = tables.plan(f)
        seen.add((info["residency"], info["lookup_mode"], info["header_packed"], info["cell_records"]))
        assert info["n_leaves"] == L and info["smem_bytes"] <= 227 * 1024
        assert info["grid_cells"] >= 1 and info["search_steps"] >= 0
        if info["lookup_mode"] == 2 or info["cell_records"] or info["header_packed"] == 2:
            assert info["search_steps"] == 0                     # all need an exact grid
        if info["header_packed"]:
            assert basis != "monomial" and info["residency"] in (0, 1, 3, 4)    # (4: the image of a select table)
        if info["header_packed"] == 2:
            assert info["cell_records"] or L <= 32 or dtype == np.float64
        if info["residency"] == 4:
            assert L == 2
        if info["residency"] == 2:
            assert info["smem_bytes"] == 0 and info["smem_resident"] == 0
        if info["search_steps"] > 0:                                 # breakpoints off the grid
            assert info["lookup_mode"] == 1 and info["header_packed"] != 2 and not info["cell_records"]
        if info["lookup_mode"] == 4:                                 # path lookup: at most 64 rows, nothing else
            assert info["search_steps"] == 0 and info["grid_cells"] <= 64 and not info["cell_records"]
            assert info["header_packed"] != 2 and info["residency"] in (0, 1, 2, 3, 9)
    if kind == "path":
        assert any(m[1] == 4 for m in seen), seen
    assert len(seen) >= 2


def test_plan_rejects_a_nonsensical_leaf_count_with_a_status():
    """No C++ exception may cross the C ABI: n_leaves far beyond any table is AI_ERR_UNSUPPORTED
    before anything is allocated (the arrays behind the pointers are never read)."""
    import ctypes
    from adaptive_interpolation_b200 import _cabi
    L = _cabi.lib()
    one = np.zeros(4, dtype=np.float64)
    info = _cabi.TableInfo()
    rc = L.ai_table_plan(0, 3, 1, 1 << 40, one.ctypes.data, one.ctypes.data, 0, ctypes.byref(info))
    assert rc == 4 and b"leaves" in L.ai_last_error()
    rc = L.ai_table_plan(0, 3, 1, 0, one.ctypes.data, one.ctypes.data, 0, ctypes.byref(info))
    assert rc == 2


def test_refitted_large_tables_take_the_compact_layout(monkeypatch):
    """The reference's missing order-3 / 2.2e-14 and 2.2e-15 fits (6872-11680 leaves, refitted by its
    own fitter) are dyadic: exact grid of up to 2^20 fine cells, one headerless 32-byte record per
    coarse cell, 4-bit levels in shared memory.  With the layout disabled they fall back to grid +
    search over 16-byte padded records."""
    for name in NAMES:
        if not name.startswith("large__"):
            continue
        info = plan(name)
        assert info["residency"] == 10 and info["record_stride"] == 4 and info["search_steps"] == 0, (name, info)
        assert info["n_leaves"] > 4096
    monkeypatch.setenv("AI_B200_NO_COMPACT", "1")
    info = plan("large__approx__64o3-p2.22044604925e-14")
    assert info["residency"] == 2 and info["record_stride"] == 6
