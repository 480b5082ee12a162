"""Loading the reference's saved approximations without the reference installed.

The saved tables (approx/, approximations/) are pickles of
adaptive_interpolation.approximator.Approximator (adaptive_interpolation.py:78-105).  They
reference four reference classes and one user function:
    adaptive_interpolation.approximator.Approximator
    adaptive_interpolation.adapt.Interpolant / Tree / Node      (approximator.py:26-27 keep the
                                                                  fitter's bound method and tree)
    __main__.f                                                   (the fitted function)
When the reference package is importable its own classes are used and nothing changes.  When it
is not (a GPU box with only this package), the unpickler maps them onto the plain data holders
below: same attribute names, no fitting code.  Evaluation of a stand-in goes to the GPU.
"""
import io
import pickle
import sys

import numpy as np


class Tree(object):
    """Data holder for adapt.py:14-24."""

    def __init__(self, root=0):
        self.root = root
        self.size = 0
        self.max_level = 0


class Node(object):
    """Data holder for adapt.py:28-48."""

    def __init__(self, parent=0, left=0, right=0):
        self.parent, self.left, self.right = parent, left, right
        self.level = 0
        self.data = 0

    def get_level(self):
        return self.level


class Interpolant(object):
    """Data holder for adapt.py:50-93.  The fitter (Remez / bisection) is not part of this
    package; make_interpolant delegates to the reference when it is importable."""

    def basis_function(self, x, n, basis, a, b):
        raise NotImplementedError(
            "Interpolant.basis_function belongs to the reference's CPU fitter; this package "
            "evaluates on the GPU (adaptive_interpolation_b200.generate.run / run_single)")


class Approximator(object):
    """Stand-in for approximator.py:20-114: the table attributes the hot path reads
    (tree_1d, max_order, num_levels, dtype, basis, bounds, interval_a/b, coeff, map, cmap,
    optimizations) plus the codegen state (code, size, vector_width).  Device handles are never
    stored here -- see generate._REGISTRY -- so write_to_file keeps working."""

    def __init__(self, my_adapt=None, optimizations=[]):
        self.optimizations = list(optimizations)
        self.code, self.size, self.vector_width = 0, None, None
        self.kernal = self.queue = self.x_dev = self.y_dev = self.tree_dev = None

    @classmethod
    def from_record(cls, rec):
        """Build from any Approximator-like record (e.g. tables.GoldenRecord)."""
        ap = cls(optimizations=getattr(rec, "optimizations", []))
        for name in ("basis", "max_order", "num_levels", "dtype", "dtype_name", "lower_bound",
                     "upper_bound", "tree_1d", "interval_a", "interval_b", "coeff", "map", "cmap"):
            if hasattr(rec, name):
                setattr(ap, name, getattr(rec, name))
        if hasattr(ap, "interval_a") and len(ap.interval_a):
            ap.intervals = list(ap.interval_a) + [ap.interval_b[-1]]
            ap.leaf_index = len(ap.interval_a) - 1
        return ap

    def evaluate(self, x):
        """Same call as approximator.py:273-296, served by the CUDA path (no Python loop)."""
        from . import generate
        return list(generate.run(np.asarray(x, dtype=self.dtype), self))


def _missing_function(*args, **kwargs):
    raise RuntimeError("the function this approximation was fitted to (__main__.f in the pickle) is "
                       "not defined in this process; it is only needed for re-fitting")


_STANDINS = {
    ("adaptive_interpolation.approximator", "Approximator"): Approximator,
    ("adaptive_interpolation.adapt", "Interpolant"): Interpolant,
    ("adaptive_interpolation.adapt", "Tree"): Tree,
    ("adaptive_interpolation.adapt", "Node"): Node,
}


class _Unpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if module == "__main__":
            main = sys.modules.get("__main__")
            if main is not None and hasattr(main, name):
                return getattr(main, name)
            return _missing_function
        if (module, name) in _STANDINS:
            try:
                return pickle.Unpickler.find_class(self, module, name)     # the reference's own class
            except (ImportError, AttributeError):
                return _STANDINS[(module, name)]
        if module.startswith("numpy.core") or module.startswith("numpy._core"):
            # numpy >= 2 moved numpy.core -> numpy._core; the saved pickles name numpy.core.  Try the
            # spelling this numpy prefers first and the other one on failure (numpy 1.x has no _core).
            new = module.replace("numpy.core", "numpy._core", 1)
            old = module.replace("numpy._core", "numpy.core", 1)
            first, second = (new, old) if int(np.__version__.split(".")[0]) >= 2 else (old, new)
            try:
                return pickle.Unpickler.find_class(self, first, name)
            except (ImportError, AttributeError):
                return pickle.Unpickler.find_class(self, second, name)
        return pickle.Unpickler.find_class(self, module, name)


def load(file_path):
    with open(file_path, "rb") as fh:
        return _Unpickler(io.BytesIO(fh.read())).load()
