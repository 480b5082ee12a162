"""Drop-in facade with the reference's public entry points, served by the CUDA backend.

Entry points and where the reference defines them (adaptive_interpolation/adaptive_interpolation.py):

    make_interpolant(a, b, func, order, error, basis, adapt_type, dtype, accurate, optimizations)  :14
    generate_code(approx, size=None, vector_width=1, cpu=False) -> str                             :37
    write_to_file(file_path, approx) / load_from_file(file_path)                                   :78 / :91
    run_approximation(x, approx) -> (run_time, output)                                             :111
    run_code(x, approx, vectorized=None)   legacy spelling, tests/generate_tests.py:26

Contract kept: signatures, the attributes set on the Approximator (vector_width, size, code), the
bare `Exception` error convention and its message texts.  What differs is what happens underneath:
"code" names a precompiled sm_100a kernel instead of holding generated source, and evaluation runs
through libai_b200.so (see generate.py in this package).  The adaptive fit itself is the
reference's CPU code and stays there; make_interpolant only forwards to it.
"""
from __future__ import absolute_import

import contextlib
import pickle

import numpy as np

from . import approximator as _approximator
from . import fit as _fit
from . import generate

# --- message texts callers (and the reference's own tests) match on -------------------------------
_MSG_CPU_NEEDS_SIZE = "The size must be known to generate cpu code."
_MSG_ISPC_CHEB_ONLY = "ISPC generation only compatible with chebyshev interpolants currently."
_MSG_WIDTH_AND_SIZE = ("Both vector width and domain_size must be given "
                       "if running single-threaded code.")
_MSG_BAD_BASIS = "Incorrect basis provided: monomial, chebyshev, legendre."
_MSG_BAD_DTYPE = "Incorrect data type specified"
_MSG_NO_CODE = ("Approximator class does not have any associated "
                "code. Run a generate method to add code to the class.")
_MSG_LOAD_FAILED = "Error loading instance from {0}"

# basis name -> generator of the device "code" (reference: generate.gen_mono / gen_cheb / gen_leg)
_DEVICE_GENERATORS = {
    "monomial": generate.gen_mono,
    "chebyshev": generate.gen_cheb,
    "legendre": generate.gen_leg,
}
# upper bit width -> C type name stored in Approximator.dtype_name (reference :25-33)
_DTYPE_NAMES = ((32, "float"), (64, "double"), (80, "long double"))


def _dtype_name(dtype):
    bits = int(dtype)
    for limit, name in _DTYPE_NAMES:
        if bits <= limit:
            return name
    raise Exception(_MSG_BAD_DTYPE)


@contextlib.contextmanager
def _integer_linspace_counts():
    """The reference fitter hands float sample counts to np.linspace (adapt.py:182-188), which
    numpy >= 1.18 rejects.  While a fit runs, counts are truncated to int; nothing in the
    reference is modified."""
    original = np.linspace

    def linspace(start, stop, num=50, *args, **kwargs):
        return original(start, stop, int(num), *args, **kwargs)

    np.linspace = linspace
    try:
        yield
    finally:
        np.linspace = original


def make_interpolant(a, b, func, order, error, basis="chebyshev",
                     adapt_type="Remez", dtype='64', accurate=True, optimizations=[]):
    """Fit an adaptive interpolant of `func` on [a, b].  Pure delegation: the fit is the
    reference's Interpolant.run_adapt + Approximator constructor, run unchanged on the CPU.
    Passing "gpu find_error" in `optimizations` (the reference's free-form flag list) evaluates
    each node's candidate polynomial on the GPU instead of in the Python loop (fit.py)."""
    try:
        from adaptive_interpolation import adapt as reference_adapt
        from adaptive_interpolation import approximator as reference_approximator
    except ImportError:
        raise ImportError("make_interpolant runs the reference's CPU fitter (adaptive_interpolation.adapt), "
                          "which is not importable here; only evaluation is provided by this package")
    name = _dtype_name(dtype) if int(dtype) <= 80 else None
    with _integer_linspace_counts():
        fitter = reference_adapt.Interpolant(func, order, error, basis, dtype, accurate,
                                             optimizations=optimizations)
        if _fit.FLAG in optimizations:          # opt-in: GPU-evaluated find_error (fit.py)
            _fit.accelerate(fitter)
        fitter.run_adapt(a, b, adapt_type)
        table = reference_approximator.Approximator(fitter, optimizations=optimizations)
    table.dtype_name = name if name is not None else _dtype_name(dtype)
    return table


def generate_code(approx, size=None, vector_width=1, cpu=False):
    """Select the evaluator for `approx` and record it on the object.

    Mirrors the reference's argument rules: the CPU (ISPC) flavour needs a size and a Chebyshev
    basis; the device flavour needs `size` and a non-default `vector_width` together or not at
    all.  Returns the text stored in approx.code."""
    approx.vector_width, approx.size = vector_width, size
    if cpu:
        if size is None:
            raise Exception(_MSG_CPU_NEEDS_SIZE)
        if approx.basis != "chebyshev":
            raise Exception(_MSG_ISPC_CHEB_ONLY)
        text = generate.gen_ispc(approx)
    else:
        width_given, size_given = vector_width != 1, size is not None
        if width_given != size_given:
            raise Exception(_MSG_WIDTH_AND_SIZE)
        generator = _DEVICE_GENERATORS.get(approx.basis)
        if generator is None:
            raise Exception(_MSG_BAD_BASIS)
        text = generator(approx)
    approx.code = text
    return text


def write_to_file(file_path, approx):
    """Pickle the whole Approximator, as the reference does.  Works because this package never
    parks device handles on the object."""
    with open(file_path, "wb") as stream:
        pickle.dump(approx, stream, pickle.HIGHEST_PROTOCOL)


def load_from_file(file_path):
    """Unpickle a saved approximation.  Unlike the reference this also works when neither the
    reference package nor the fitted function (`__main__.f`) exists in the process -- see
    approximator.load.  Any failure surfaces as the reference's generic Exception."""
    try:
        return _approximator.load(file_path)
    except Exception:
        raise Exception(_MSG_LOAD_FAILED.format(file_path))


def run_approximation(x, approx):
    """Evaluate `approx` at x -> (run_time, output).  run_time is the kernel-only device time in
    seconds when generate_code recorded a vector width, else None (one-shot path)."""
    if approx.code == 0:
        raise Exception(_MSG_NO_CODE)
    if approx.vector_width is None:
        return None, generate.run(x, approx)
    generate.build_code(approx)
    return generate.run_single(x, approx)


def run_code(x, approx, vectorized=None):
    """Legacy entry point (tests/generate_tests.py:26, performance_tests.py:98-107): output only."""
    return run_approximation(x, approx)[1]
