"""Mirror of the reference facade adaptive_interpolation/adaptive_interpolation.py.

    make_interpolant(a, b, func, order, error, basis, adapt_type, dtype, accurate, optimizations)   :14
    generate_code(approx, size=None, vector_width=1, cpu=False)                                     :37
    write_to_file(file_path, approx) / load_from_file(file_path)                                    :78 / :91
    run_approximation(x, approx) -> (run_time, output)                                              :111
    run_code(x, approx, vectorized=None)      legacy spelling used by tests/generate_tests.py:26

Signatures, side effects (approx.vector_width / .size / .code) and error behaviour (bare
Exception with the reference's messages) are kept; generate_code / run_approximation are
re-pointed at the CUDA backend (generate.py in this package).  Fitting (make_interpolant) is the
reference's CPU code, unchanged and out of scope: it is delegated to the reference package when
that is importable.
"""
from __future__ import absolute_import

import contextlib
import pickle

import numpy as np

from . import approximator as app
from . import generate


@contextlib.contextmanager
def _numpy_compat():
    """adapt.py:182-188 passes float sample counts to np.linspace, a TypeError on numpy >= 1.18
    (SURVEY.md appendix C-1).  Cast them for the duration of a fit; the reference is untouched."""
    real = np.linspace

    def linspace(start, stop, num=50, *args, **kwargs):
        return real(start, stop, int(num), *args, **kwargs)
    np.linspace = linspace
    try:
        yield
    finally:
        np.linspace = real


def make_interpolant(a, b, func, order, error, basis="chebyshev",
                     adapt_type="Remez", dtype='64', accurate=True, optimizations=[]):
    """adaptive_interpolation.py:14-34.  The adaptive fit is the reference's own CPU code."""
    try:
        import adaptive_interpolation.adapt as ref_adapt
        import adaptive_interpolation.approximator as ref_app
    except ImportError:
        raise ImportError("make_interpolant runs the reference's CPU fitter (adaptive_interpolation.adapt), "
                          "which is not importable here; only evaluation is provided by this package")
    with _numpy_compat():
        my_adapt = ref_adapt.Interpolant(func, order, error, basis, dtype, accurate, optimizations=optimizations)
        my_adapt.run_adapt(a, b, adapt_type)
        approximation = ref_app.Approximator(my_adapt, optimizations=optimizations)
    dt = int(dtype)
    if dt <= 32:
        approximation.dtype_name = "float"
    elif dt <= 64:
        approximation.dtype_name = "double"
    elif dt <= 80:
        approximation.dtype_name = "long double"
    else:
        raise Exception("Incorrect data type specified")
    return approximation


def generate_code(approx, size=None, vector_width=1, cpu=False):
    """adaptive_interpolation.py:37-75: same validation, same side effects.  The returned
    "code" names the precompiled CUDA kernel that will evaluate this approximation."""
    approx.vector_width = vector_width
    approx.size = size
    if cpu:
        if size is None:
            raise Exception("The size must be known to generate cpu code.")
        if approx.basis != 'chebyshev':
            string_err = "ISPC generation only compatible with "
            string_err += "chebyshev interpolants currently."
            raise Exception(string_err)
        else:
            code = generate.gen_ispc(approx)
    else:
        if ((vector_width != 1) and (size is None)) \
                or ((vector_width == 1) and (size is not None)):
            string_err = "Both vector width and domain_size must be given "
            string_err += "if running single-threaded code."
            raise Exception(string_err)
        if approx.basis == 'monomial':
            code = generate.gen_mono(approx)
        elif approx.basis == 'chebyshev':
            code = generate.gen_cheb(approx)
        elif approx.basis == 'legendre':
            code = generate.gen_leg(approx)
        else:
            raise Exception("Incorrect basis provided: monomial, chebyshev, legendre.")
    approx.code = code
    return code


def write_to_file(file_path, approx):
    """adaptive_interpolation.py:78-88."""
    with open(file_path, 'wb') as f:
        pickle.dump(approx, f, pickle.HIGHEST_PROTOCOL)


def load_from_file(file_path):
    """adaptive_interpolation.py:91-105.  Also loads the saved tables when the reference package
    (and the `__main__.f` they reference) is absent -- see approximator.load."""
    try:
        return app.load(file_path)
    except Exception:
        str_err = "Error loading instance from {0}".format(file_path)
        raise Exception(str_err)


def run_approximation(x, approx):
    """adaptive_interpolation.py:111-123 -> (run_time, output)."""
    if approx.code == 0:
        string_err = "Approximator class does not have any associated "
        string_err += "code. Run a generate method to add code to the class."
        raise Exception(string_err)
    if approx.vector_width is None:
        output = generate.run(x, approx)
        run_time = None
    else:
        generate.build_code(approx)
        run_time, output = generate.run_single(x, approx)
    return run_time, output


def run_code(x, approx, vectorized=None):
    """Legacy entry point (tests/generate_tests.py:26, performance_tests.py:98-107): returns
    the output only."""
    return run_approximation(x, approx)[1]
