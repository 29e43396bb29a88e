#!/usr/bin/env python
"""Folder format of a stored posterior, written BY THE REFERENCE's own BayesianModel.store and
MultivariateNormalDiagPlusLowRank.store (Pyesian/nn/BayesianModel.py:177-203, distributions/
MultivariateNormalDiagPlusLowRank.py:11-16) on the TensorFlow stand-in, and the reverse direction checked on the spot: a
folder written by THIS repo's BayesianModel.store is loaded by the reference's BayesianModel.load and samples the same
parameters.  (Sampled's per-sample files are tf.io.serialize_tensor bytes — third-party — and are covered by the
TensorProto known-bytes test instead.)

    python -B tests/golden/make_reference_store_golden.py        # writes tests/golden/reference_store.json
"""
import json
import os
import sys
import tempfile
import warnings
from unittest.mock import MagicMock

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import tf_shim  # noqa: E402


def main():
    import torch
    sys.dont_write_bytecode = True
    warnings.simplefilter("ignore")
    sys.modules["tensorflow"] = tf_shim.make_tf()
    sys.modules["tensorflow_probability"] = tf_shim.make_tfp()
    for name in ["wandb", "wandb.integration", "wandb.integration.keras", "tensorflow_datasets", "ucimlrepo", "matplotlib",
                 "matplotlib.pyplot", "scikitplot"]:
        sys.modules.setdefault(name, MagicMock())
    sys.path.insert(0, "/root/reference")
    import Pyesian.nn  # noqa: F401
    import Pyesian.distributions  # noqa: F401
    RefBM = sys.modules["Pyesian.nn.BayesianModel"].BayesianModel
    RefLowRank = sys.modules["Pyesian.distributions.MultivariateNormalDiagPlusLowRank"].MultivariateNormalDiagPlusLowRank
    from bayesian_inference_for_nn_b200 import keras_json
    from bayesian_inference_for_nn_b200.distributions import MultivariateNormalDiagPlusLowRank as OurLowRank
    from bayesian_inference_for_nn_b200.nn import BayesianModel as OurBM

    js = keras_json.make_sequential_json(3, [4, 2], ["relu", "softmax"])
    rng = np.random.default_rng(0)
    sizes = [3 * 4 + 4, 4 * 2 + 2]
    params = [dict(mean=rng.normal(size=n).astype(np.float32), diag=rng.uniform(0.1, 0.2, n).astype(np.float32),
                   D=rng.normal(size=(n, 3)).astype(np.float32)) for n in sizes]
    tt = lambda a: tf_shim.TT(torch.as_tensor(a))

    # ---- written by the reference
    bm = RefBM(js)
    for layer, p in enumerate(params):
        bm.apply_distribution(RefLowRank(tt(p["mean"]), tt(p["diag"]), tt(p["D"])), layer, layer)
    files = {}
    with tempfile.TemporaryDirectory() as tmp:
        bm.store(tmp)
        for dp, _, fs in os.walk(tmp):
            for f in fs:
                files[os.path.relpath(os.path.join(dp, f), tmp)] = open(os.path.join(dp, f)).read()

    # ---- written by this repo, read by the reference
    ours = OurBM(js)
    for layer, p in enumerate(params):
        ours.apply_distribution(OurLowRank(p["mean"], p["diag"], p["D"]), layer, layer)
    with tempfile.TemporaryDirectory() as tmp:
        ours.store(tmp)
        back = RefBM.load(tmp)
        assert [list(iv) for iv in back._layers_dtbn_intervals] == [[0, 0], [1, 1]]
        for d, p in zip(back._distributions, params):
            assert type(d).__name__ == "MultivariateNormalDiagPlusLowRank"
            np.testing.assert_allclose(d._mean.numpy(), p["mean"], rtol=1e-6)
            np.testing.assert_allclose(d._diag.numpy(), p["diag"], rtol=1e-6)
            np.testing.assert_allclose(d._D.numpy(), p["D"], rtol=1e-6)
        reverse_ok = True
    out = {"model_json": js, "files": files, "params": [{k: v.tolist() for k, v in p.items()} for p in params],
           "reference_loaded_our_folder": reverse_ok}
    with open(os.path.join(HERE, "reference_store.json"), "w") as f:
        json.dump(out, f)
    print(sorted(files), "| reference loaded a folder written by this repo:", reverse_ok)
    print(files["layers_config.txt"].replace("\n", "\\n"))


if __name__ == "__main__":
    main()
