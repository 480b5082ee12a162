"""Loader for the frozen reference fixtures under tests/golden/tables/ (TEST INFRASTRUCTURE).

A fixture is a table the reference saved or fitted plus the reference's own evaluation of it on
seeded points (made by tests/golden/make_golden.py, the only file that imports /root/reference).
Used by tests/, by bench.py and tools/ to obtain workload tables on the GPU box (where the
reference tree does not exist) and by __graft_entry__.smoke().  The product package never
imports this module.
"""
import json
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "tables")


class GoldenRecord(object):
    """A frozen reference table + the reference's own evaluation of it.  Carries the attributes of
    the reference Approximator that the hot path reads (approximator.py:22-136)."""

    def __init__(self, path):
        d = np.load(path, allow_pickle=False)
        self.meta = json.loads(str(d["meta"]))
        m = self.meta
        self.name = m["name"]
        self.basis = m["basis"]
        self.max_order = m["order"]
        self.num_levels = m["num_levels"]
        self.dtype = np.dtype(m["dtype"]).type
        self.dtype_name = m["dtype_name"]
        self.lower_bound = self.dtype(m["lower_bound"])
        self.upper_bound = self.dtype(m["upper_bound"])
        self.optimizations = list(m["optimizations"])
        self.allowed_error = m["allowed_error"]
        self.fmax = m["fmax"]
        self.tree_1d = d["tree_1d"]
        self.interval_a = list(d["interval_a"])
        self.interval_b = list(d["interval_b"])
        self.coeff = list(d["coeff"])
        self.map = d["map"]
        self.cmap = d["cmap"]
        self.code, self.size, self.vector_width = 0, None, None
        self.kernal = self.queue = self.tree_dev = self.x_dev = self.y_dev = None
        self.x, self.y_ref, self.idx_ref, self.f_true = d["x"], d["y_ref"], d["idx_ref"], d["f_true"]


def golden_names():
    if not os.path.isdir(GOLDEN_DIR):
        return []
    return sorted(f[:-4] for f in os.listdir(GOLDEN_DIR) if f.endswith(".npz"))


def load_golden(name):
    return GoldenRecord(os.path.join(GOLDEN_DIR, name + ".npz"))
