#!/usr/bin/env python
"""C1 (HMC make_moons 2-50-2, 1600 rows, L=30) for ncu: S chains, a few iterations."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from bayesian_inference_for_nn_b200 import _lib, keras_json
from bayesian_inference_for_nn_b200.engine import Engine
from tools.bench_small import moons

S = int(os.environ.get("S", 4096))
X, y = moons(1600)
spec = keras_json.parse_model_json(keras_json.make_sequential_json(2, [50, 2], ["relu", "softmax"]))
eng = Engine(spec, device=0, seed=3)
eng.set_dataset(X, y, _lib.LOSS_SPARSE_CE)
eng.set_prior([0.0], [1.0], _lib.PRIOR_SCALAR)
eng.hmc_init(S, 0.005, 0.5, 30, _lib.HMC_REFERENCE)
for _ in range(3):
    d = eng.hmc_run(1, burning=False, sampling=True)
print("ms/iter", d["device_ms"], "grad-evals/s", S * 31 / d["device_ms"] * 1e3)
