#!/usr/bin/env python
"""What the reference's GaussianPrior (Pyesian/distributions/GaussianPrior.py) builds for every accepted form of
(mean, rho) and how it fails for the others, recorded by running THE REFERENCE's class against the Keras stand-in of
tf_shim.py (a Flatten + Dense + Dense model, so that a parameter-less layer is part of it).

    python -B tests/golden/make_reference_prior_golden.py      # writes tests/golden/reference_prior.json
"""
import json
import os
import sys
import warnings
from unittest.mock import MagicMock

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import tf_shim  # noqa: E402

# (mean, rho) as Python literals: the test rebuilds them with eval-free JSON
FORMS = {"float": (0.5, 2.0), "int": (1, 3), "negative_scale": (0.0, -1.0),
         "per_layer": ([0.0, 0.25, -1.0], [1.0, 2.0, 0.5]),                 # indexed by model.layers position (Flatten = 0)
         "per_layer_int": ([0, 1, 2], [1, 1, 2]),
         "type_mismatch": (0.0, 1), "list_vs_float": ([0.0], 1.0), "strings": ("a", "b"), "mixed_list": ([0.0, 1], [1.0, 1]),
         "nested_lists": ([[], [[[0.0] * 3] * 4, [0.0] * 3], [[[0.0] * 2] * 3, [0.0] * 2]],
                          [[], [[[1.0] * 3] * 4, [1.0] * 3], [[[1.0] * 2] * 3, [1.0] * 2]])}


def nested_arrays(fill_mean, fill_rho, bad_shape=False):
    """the per-variable ("tensor") form: one array per trainable variable, [] for the parameter-less layer"""
    shapes = [[], [(4, 3), (3,)], [(3, 2), (2,)]]
    if bad_shape:
        shapes[2][0] = (2, 3)
    mk = lambda v: [[np.full(s, v, np.float32) for s in layer] for layer in shapes]
    return mk(fill_mean), mk(fill_rho)


def main():
    sys.dont_write_bytecode = True
    warnings.simplefilter("ignore")
    sys.modules["tensorflow"] = tf_shim.make_tf()
    sys.modules["tensorflow_probability"] = tf_shim.make_tfp()
    for name in ["wandb", "wandb.integration", "wandb.integration.keras", "tensorflow_datasets", "ucimlrepo", "matplotlib",
                 "matplotlib.pyplot", "scikitplot"]:
        sys.modules.setdefault(name, MagicMock())
    sys.path.insert(0, "/root/reference")
    import Pyesian.distributions  # noqa: F401
    GaussianPrior = sys.modules["Pyesian.distributions.GaussianPrior"].GaussianPrior
    model = tf_shim.Model(4, [(3, "relu", True), (2, "softmax", True)], leading_parameterless=1)     # Flatten, Dense, Dense
    out = {}
    forms = dict(FORMS)
    forms["nested_arrays"] = nested_arrays(0.5, 2.0)
    forms["nested_arrays_bad_shape"] = nested_arrays(0.5, 2.0, bad_shape=True)
    for name, (mean, rho) in forms.items():
        rec = {} if name.startswith("nested_arrays") else {"mean": mean, "rho": rho}
        try:
            priors = GaussianPrior(mean, rho).get_model_priors(model)
            if priors is None:
                rec["result"] = None                 # the nested ("tensor") form builds its list and returns nothing (:98)
            else:
                rec["result"] = [None if layer is None else [[d.loc.numpy().tolist() if d.loc.ndim else float(d.loc),
                                                              d.scale.numpy().tolist() if d.scale.ndim else float(d.scale)]
                                                             for d in layer] for layer in priors]
        except Exception as e:
            rec["error"] = [type(e).__name__, str(e)]
        out[name] = rec
        print(name, "->", "error " + rec["error"][1][:50] if "error" in rec else ("None" if rec["result"] is None else "ok"))
    with open(os.path.join(HERE, "reference_prior.json"), "w") as f:
        json.dump(out, f)


if __name__ == "__main__":
    main()
