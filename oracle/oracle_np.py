"""Vectorised numpy restatement of the reference evaluator.  TEST INFRASTRUCTURE ONLY.

Follows, in the table dtype and in the reference's operation order:
  approximator.py:285-290  leaf selection.  The BST walk (go left iff mid > x, else right; NaN
                           -> right; leaves self-loop) is equivalent to
                           i = clip(searchsorted(breaks, x, 'right') - 1, 0, L-1), NaN -> L-1
                           (verified against the literal walk in tests/test_oracle.py)
  adapt.py:130             t = (2*(x - a)/(b - a)) - 1
  adapt.py:111-120         Chebyshev   T0=1, T1=t, Ti = (2*t)*T(i-1) - T(i-2)
  adapt.py:97-108          "Legendre"  L0=1, L1=t, Li = ((2i-1)*t*L(i-1) + (i-1)*L(i-2))*(1./n)
  adapt.py:139             monomial    x**i on the raw x
  approximator.py:294-295  y = dot(c, basis)        (left-to-right here; BLAS upstream)
"""
import numpy as np


def leaf_index(breaks, x):
    """Leaf ordinal the reference's BST walk selects, for sorted breakpoints breaks[0..L]."""
    breaks = np.asarray(breaks)
    x = np.asarray(x, dtype=breaks.dtype)
    L = len(breaks) - 1
    i = np.searchsorted(breaks, x, side="right") - 1
    i = np.clip(i, 0, L - 1)
    i = np.where(np.isnan(x), L - 1, i)
    return i.astype(np.int32)


def walk_tree_1d(tree_1d, order, num_levels, x):
    """The literal walk of approximator.py:285-290 (pure Python; small inputs only).
    Returns the element offset of the selected node for each x."""
    t = np.asarray(tree_1d)
    child = int(order) + 4
    out = np.empty(len(x), dtype=np.int64)
    for k, xn in enumerate(np.asarray(x, dtype=t.dtype)):
        index = 0
        for _ in range(1, int(num_levels)):
            if t[index] > xn:
                index = int(t[index + child])
            else:
                index = int(t[index + child + 1])
        out[k] = index
    return out


def basis_values(basis, order, dtype, x, a, b):
    """B[i] arrays, i = 0..order, each rounded to `dtype` after every operation."""
    dt = np.dtype(dtype).type
    n = int(order)
    x = np.asarray(x, dtype=dt)
    if basis == "monomial":
        return [np.power(x, dt(i)).astype(dt) for i in range(n + 1)]
    t = (dt(2) * (x - a) / (b - a)) - dt(1)
    B = [np.ones_like(t)]
    if n == 0:
        return B
    B.append(t)
    if basis == "chebyshev":
        for i in range(2, n + 1):
            B.append(dt(2) * t * B[i - 1] - B[i - 2])
    elif basis == "legendre":
        invn = dt(1.0 / n)
        for i in range(2, n + 1):
            first = dt(2 * i - 1) * t * B[i - 1]
            second = dt(i - 1) * B[i - 2]
            B.append((first + second) * invn)
    else:
        raise ValueError(basis)
    return B


def basis_derivatives(basis, order, t):
    """dB_i/dt in float64 (differentiated recurrences); zero for the monomial basis, whose
    argument is the raw x and carries no rescale rounding."""
    n = int(order)
    t = np.asarray(t, dtype=np.float64)
    B = [np.ones_like(t), t]
    D = [np.zeros_like(t), np.ones_like(t)]
    if basis == "monomial":
        return [np.zeros_like(t) for _ in range(n + 1)]
    for i in range(2, n + 1):
        if basis == "chebyshev":
            B.append(2 * t * B[i - 1] - B[i - 2])
            D.append(2 * B[i - 1] + 2 * t * D[i - 1] - D[i - 2])
        else:
            B.append(((2 * i - 1) * t * B[i - 1] + (i - 1) * B[i - 2]) / n)
            D.append(((2 * i - 1) * (B[i - 1] + t * D[i - 1]) + (i - 1) * D[i - 2]) / n)
    return D[:n + 1]


def error_scale(breaks, coef, basis, order, x):
    """The scale (per point, float64) on which two correct evaluators may differ by O(eps):
         sum_i |c_i B_i(t)|            rounding inside the recurrence and the dot product
       + (1 + |t|) * sum_i |c_i B_i'(t)|   the rescaled argument t itself is only known to
                                           ~eps*(1+|t|): 2*(x-a)/(b-a)-1 (adapt.py:130) and
                                           (2./(b-a))*(x-a)-1 (generate.py:83) round differently
    For Chebyshev fits of J0 the first term alone vanishes at zeros of the function sitting at a
    leaf centre while the second does not; max|f| (the reference tests' scale,
    tests/generate_tests.py:27) is an upper-bound proxy for their sum."""
    breaks = np.asarray(breaks)
    dt = breaks.dtype.type
    n = int(order)
    coef = np.asarray(coef, dtype=dt).reshape(len(breaks) - 1, n + 1).astype(np.float64)
    x = np.asarray(x, dtype=dt)
    idx = leaf_index(breaks, x)
    a, b = breaks[idx].astype(np.float64), breaks[idx + 1].astype(np.float64)
    xx = x.astype(np.float64)
    with np.errstate(all="ignore"):
        if basis == "monomial":
            B = [xx ** i for i in range(n + 1)]
            D = [np.zeros_like(xx)] * (n + 1)
            t = np.zeros_like(xx)
        else:
            t = 2 * (xx - a) / (b - a) - 1
            B = [b_.astype(np.float64) for b_ in basis_values(basis, n, np.float64, t * 0 + xx, a, b)]
            D = basis_derivatives(basis, n, t)
        s0 = sum(np.abs(coef[idx, i] * B[i]) for i in range(n + 1))
        s1 = sum(np.abs(coef[idx, i] * D[i]) for i in range(n + 1))
        return s0 + (1 + np.abs(t)) * s1


def evaluate(breaks, coef, basis, order, x, want_cond=False):
    """y (and optionally idx, cond) for a flattened table.  coef has shape [L, order+1]."""
    breaks = np.asarray(breaks)
    dt = breaks.dtype.type
    coef = np.asarray(coef, dtype=dt).reshape(len(breaks) - 1, int(order) + 1)
    x = np.asarray(x, dtype=dt)
    idx = leaf_index(breaks, x)
    a, b = breaks[idx], breaks[idx + 1]
    with np.errstate(all="ignore"):
        B = basis_values(basis, order, dt, x, a, b)
        y = np.zeros_like(x)
        cond = np.zeros(x.shape, dtype=np.float64)
        for i in range(int(order) + 1):
            c = coef[idx, i]
            y = y + c * B[i]
            if want_cond:
                cond += np.abs(c.astype(np.float64) * B[i].astype(np.float64))
    if want_cond:
        return y, idx, cond
    return y
