#!/usr/bin/env python
"""Two canonical SVGD steps at BASELINE config C4 (4096 particles, 784-128-10, minibatch 1024 of 60000) for an ncu
launch list: `ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python tools/prof_c4.py`."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bayesian_inference_for_nn_b200 import _lib, keras_json  # noqa: E402
from bayesian_inference_for_nn_b200.engine import Engine  # noqa: E402

rng = np.random.default_rng(0)
S = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
sp = keras_json.parse_model_json(keras_json.make_sequential_json(784, [128, 10], ["relu", "softmax"]))
X = rng.random((60000, 784), dtype=np.float32)
y = rng.integers(0, 10, 60000).astype(np.int32)
eng = Engine(sp, seed=1)
eng.set_dataset(X, y, _lib.LOSS_SPARSE_CE)
eng.set_prior([0.0], [1.0], _lib.PRIOR_SCALAR)
eng.svgd_init(S, 0.01, _lib.SVGD_CANONICAL_MEDIAN)
if "PYB_SELECT_COMPACT" in os.environ:
    eng.set_option("select_compact", int(os.environ["PYB_SELECT_COMPACT"]))
for _ in range(int(os.environ.get("STEPS", 2))):
    eng.svgd_step(rng.permutation(60000)[:1024].astype(np.int32))
    print("step ms", eng.info("last_device_ms"))
