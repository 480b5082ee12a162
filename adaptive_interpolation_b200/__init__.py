"""adaptive_interpolation_b200 -- B200-native evaluator for the batched interpolant-evaluation
hot path of jdoherty7/Adaptive_Interpolation, behind the reference's own Python surface.

    import adaptive_interpolation_b200.adaptive_interpolation as adapt_i    # drop-in for
    import adaptive_interpolation_b200.generate as generate                 # the reference modules
"""
from .adaptive_interpolation import (make_interpolant, generate_code, run_approximation, run_code,  # noqa: F401
                                     write_to_file, load_from_file)

__version__ = "0.1.0"
