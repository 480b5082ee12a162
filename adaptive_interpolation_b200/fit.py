"""GPU-assisted error estimation for the reference's adaptive fitter.

The fitter itself (bisection, Remez exchange) is the reference's CPU code and stays there.  Its
cost per tree node is dominated by `Interpolant.find_error` (adapt.py:180-198), which evaluates
the candidate polynomial at up to 5000 points of the node in a per-point Python loop
(`eval_coeff`, adapt.py:144-154) and re-evaluates the fitted function on the whole domain for the
normalisation.  That per-node evaluation is exactly this package's hot path with a one-leaf
table, so `accelerate(fitter)` swaps `find_error` on ONE fitter instance for a version that

  * evaluates the candidate on the GPU (ai_table_create with L = 1, ai_eval_*), and
  * computes max|f| over the whole domain once per fit instead of once per node (same value).

Opt-in only: `make_interpolant(..., optimizations=[..., "gpu find_error"])`.  The default fit is
untouched.  The error estimate differs from the reference's in the last bits (FMA evaluation), so
a split decision can differ only when an estimate sits within ~1e-15 of the threshold.
"""
import types

import numpy as np
import numpy.linalg as la

from . import _cabi, generate, tables

FLAG = "gpu find_error"


def _gpu_eval_single_leaf(basis, order, dtype, coeff, a, b, x):
    dt = np.dtype(dtype)
    flat = tables.FlatTable(basis, int(order), dt, np.array([a, b], dtype=dt), np.asarray(coeff, dtype=dt))
    tab = tables.DeviceTable(flat, generate._backend().cuda.current_device())
    try:
        return generate.run(np.asarray(x, dtype=dt), tab)
    finally:
        tab.close()


def accelerate(fitter):
    """Replace `fitter.find_error` (one reference Interpolant instance) with the GPU version."""
    original = fitter.find_error
    cache = {}

    def find_error(self, coeff, a, b, order):
        # sample counts and grids exactly as adapt.py:182-188
        n = int(min(5e3 * (b - a) + 10, 5e3))
        lb, ub = self.lower_bound, self.upper_bound
        key = (float(lb), float(ub))
        if key not in cache:
            full_x = np.linspace(lb, ub, int(5e3 * (ub - lb) + 10), dtype=self.dtype)
            cache[key] = la.norm(self.function(full_x), np.inf)
        x = np.linspace(a, b, n, dtype=self.dtype)
        # The two documented degenerate nodes go to the reference's own find_error: an interval that
        # is empty in the table dtype (no table can be built on it) and an order the kernels are not
        # instantiated for.  Everything else -- a missing library, a CUDA error, a shape bug -- raises.
        dt = np.dtype(self.dtype).type
        if not (dt(a) < dt(b)) or not (0 <= int(order) <= _cabi.MAX_ORDER):
            return original(coeff, a, b, order)
        approx = _gpu_eval_single_leaf(self.basis, order, self.dtype, coeff, a, b, x)
        max_abs_err = la.norm(approx - self.function(x), np.inf)
        return max_abs_err / cache[key]

    fitter.find_error = types.MethodType(find_error, fitter)
    return fitter
