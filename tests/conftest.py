import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")
if GOLDEN not in sys.path:
    sys.path.insert(0, GOLDEN)

import golden_io  # noqa: E402  (tests/golden/golden_io.py: the fixture loader)

REFERENCE_ROOT = "/root/reference"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def has_reference():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "adaptive_interpolation"))


@pytest.fixture(scope="session")
def golden_names():
    names = golden_io.golden_names()
    assert len(names) >= 44, "golden fixtures missing: run tests/golden/make_golden.py"
    return names


_cache = {}


def golden(name):
    if name not in _cache:
        _cache[name] = golden_io.load_golden(name)
    return _cache[name]


def all_golden_names():
    return golden_io.golden_names()


def value_tolerance(dtype, cond, ulps):
    """Allowed |y - y_ref|: `ulps` units of eps * sum_i |c_i B_i(x)| (the magnitude the rounding
    errors of ANY evaluation order live on; for Chebyshev it is ~ max|f|), plus a denormal floor."""
    eps = np.finfo(dtype).eps
    return ulps * eps * cond + np.finfo(dtype).tiny
