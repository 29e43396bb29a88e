"""bayesian_inference_for_nn_b200 — B200-native particle-batched posterior inner loop behind the
Pyesian optimizer API (HMC / SVGD / BayesianModel.predict).  See DESIGN.md and include/pyesian_b200.h.
"""
from . import _lib
from ._lib import PyesianB200Error
from .keras_json import parse_model_json, make_sequential_json, ModelSpec

__all__ = ["PyesianB200Error", "parse_model_json", "make_sequential_json", "ModelSpec", "_lib"]
__version__ = "0.1.0"
