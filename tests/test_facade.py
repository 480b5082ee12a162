"""The reference-facing Python surface: same validation, side effects and error behaviour as
adaptive_interpolation/adaptive_interpolation.py:37-123.  CPU only."""
import numpy as np
import pytest
import torch

import adaptive_interpolation_b200.adaptive_interpolation as adapt_i
import adaptive_interpolation_b200.generate as generate
from adaptive_interpolation_b200 import approximator as app
from conftest import golden, has_reference


def make(name="approx__32o6-p1.19209289551e-06"):
    return app.Approximator.from_record(golden(name))


def test_generate_code_side_effects_and_text():
    ap = make()
    code = adapt_i.generate_code(ap)                            # defaults: size=None, vector_width=1
    assert ap.code == code and ap.vector_width == 1 and ap.size is None
    assert "eval_kernel<ai::kCheb, 6, float>" in code
    code = adapt_i.generate_code(ap, size=2 ** 10, vector_width=8)
    assert ap.size == 2 ** 10 and ap.vector_width == 8
    code = adapt_i.generate_code(ap, size=2 ** 10, vector_width=8, cpu=True)    # run_test.py:394
    assert "ai_eval_f32" in code
    text = generate.build_code(ap, ispc=True)                   # run_test.py:395
    assert text.startswith('extern "C" int ai_eval_f32(')


def test_generate_code_errors_match_reference():
    ap = make()
    with pytest.raises(Exception, match="The size must be known to generate cpu code."):
        adapt_i.generate_code(ap, cpu=True)
    with pytest.raises(Exception, match="Both vector width and domain_size must be given"):
        adapt_i.generate_code(ap, size=100)                     # vector_width == 1 with a size
    with pytest.raises(Exception, match="Both vector width and domain_size must be given"):
        adapt_i.generate_code(ap, vector_width=4)               # width without a size
    leg = make("gen__j0_leg_o7_f32")
    with pytest.raises(Exception, match="ISPC generation only compatible"):
        adapt_i.generate_code(leg, size=10, cpu=True)
    ap.basis = "fourier"
    with pytest.raises(Exception, match="Incorrect basis provided"):
        adapt_i.generate_code(ap)


def test_basis_dispatch():
    assert "kMono, 13, double" in adapt_i.generate_code(make("gen__j0_mono_o13_f64"))
    assert "kLeg, 7, float" in adapt_i.generate_code(make("gen__j0_leg_o7_f32"))
    assert "kCheb, 13, double" in adapt_i.generate_code(make("gen__cfg3_j0_cheb_o13_f64"))


def test_run_without_code_raises():
    ap = make()
    with pytest.raises(Exception, match="does not have any associated"):
        adapt_i.run_approximation(np.zeros(4, np.float32), ap)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU error path")
def test_no_gpu_raises_value_error_like_reference():
    """generate.py:187,477 raise ValueError when the backend is missing; there is no CPU fallback."""
    ap = make()
    adapt_i.generate_code(ap)
    with pytest.raises(ValueError, match="requires"):
        adapt_i.run_approximation(np.zeros(4, np.float32), ap)
    with pytest.raises(ValueError, match="requires"):
        generate.build_code(ap)
    with pytest.raises(ValueError, match="requires"):
        generate.run_single(np.zeros(4, np.float32), ap)


def test_pickle_roundtrip_keeps_working(tmp_path):
    ap = make()
    adapt_i.generate_code(ap, size=16, vector_width=4)
    p = tmp_path / "t"
    adapt_i.write_to_file(str(p), ap)
    back = adapt_i.load_from_file(str(p))
    assert back.code == ap.code and back.size == 16 and np.array_equal(back.tree_1d, ap.tree_1d)


@pytest.mark.skipif(not has_reference(), reason="reference tree not mounted")
def test_make_interpolant_delegates_to_reference(monkeypatch):
    import sys
    monkeypatch.syspath_prepend("/root/reference")
    for k in [k for k in sys.modules if k.startswith("adaptive_interpolation.") or k == "adaptive_interpolation"]:
        monkeypatch.delitem(sys.modules, k)
    import io, contextlib
    with contextlib.redirect_stdout(io.StringIO()):
        ap = adapt_i.make_interpolant(0, 10, lambda x: np.sin(np.sin(x)), 3, 1e-4, "chebyshev", "Trivial",
                                      "64", True, ["arrays", "map"])
    assert ap.dtype_name == "double" and ap.basis == "chebyshev" and len(ap.interval_a) > 4
    from adaptive_interpolation_b200 import tables
    f = tables.flatten(ap)
    assert f.breaks[0] == 0 and f.breaks[-1] == 10
    with pytest.raises(Exception, match="Incorrect data type specified"):
        with contextlib.redirect_stdout(io.StringIO()):
            adapt_i.make_interpolant(0, 1, np.sin, 3, 1e-3, "chebyshev", "Trivial", "128", True, ["arrays", "map"])
