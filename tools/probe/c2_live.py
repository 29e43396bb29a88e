import sys, os, time, numpy as np
sys.path.insert(0, os.getcwd())
from bayesian_inference_for_nn_b200 import _lib, keras_json
from bayesian_inference_for_nn_b200.engine import Engine
rng = np.random.default_rng(0)
X = rng.standard_normal((1600, 2)).astype(np.float32); y = (X[:, 0] > 0).astype(np.int32)
for S in (10, 64):
    for fused in (1, 0):
        eng = Engine(keras_json.parse_model_json(keras_json.make_sequential_json(2, [50, 2], ["relu", "softmax"])), seed=1)
        eng.set_option("live_fused", fused)
        eng.set_dataset(X, y, _lib.LOSS_SPARSE_CE); eng.set_prior([0.0], [1.0], _lib.PRIOR_SCALAR)
        eng.svgd_init(S, 1e-3, _lib.SVGD_REFERENCE_LIVE)
        for _ in range(3): eng.svgd_step()
        l0 = eng.info("kernel_launches"); ms = []
        for _ in range(10):
            eng.svgd_step(); ms.append(eng.info("last_device_ms"))
        print(S, "fused" if fused else "per-particle launches", "ms/step %.3f" % np.mean(ms), "launches/step", (eng.info("kernel_launches") - l0) / 10, flush=True)
        eng.close()
