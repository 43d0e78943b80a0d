"""Name shim: BASELINE.json speaks of "the undersampling path in
data/transform_utils.py"; in the reference tree that file only holds
normalisation helpers (data/transform_utils.py:1-54) and the undersampling
code lives in deep_med_lib.  The B200 implementation is in
:mod:`csmri_refinement_b200.undersampling`; this module re-exports it under
the name the spec uses."""
from .undersampling import (Undersample, cartesian_mask, cartesian_rows,  # noqa: F401
                            consume_noise_draws, undersample)
