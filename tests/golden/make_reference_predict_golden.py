#!/usr/bin/env python
"""Golden posterior-predictive outputs produced by EXECUTING THE REFERENCE's Pyesian/nn/BayesianModel.py
(apply_distribution, _sample_weights, predict — unmodified) and Pyesian/distributions/Sampled.py on the torch-backed
TensorFlow stand-in of tf_shim.py: nb_samples weighted draws with `random` seeded, weights assigned variable by
variable in the flat order, forward, NaN -> 0, running sum / nb_samples.

    python -B tests/golden/make_reference_predict_golden.py      # writes tests/golden/reference_predict.npz
"""
import os
import random
import sys
import warnings
from unittest.mock import MagicMock

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import tf_shim  # noqa: E402


def main():
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
    BayesianModel = sys.modules["Pyesian.nn.BayesianModel"].BayesianModel
    Sampled = sys.modules["Pyesian.distributions.Sampled"].Sampled
    from bayesian_inference_for_nn_b200 import keras_json          # host-side JSON writer only
    import torch

    out = {}
    rng = np.random.default_rng(0)
    for name, (D, units, acts, n_stored, Nt, nb) in {"ce": (3, [6, 4], ["relu", "softmax"], 7, 25, 40),
                                                     "reg": (2, [5, 1], ["tanh", "linear"], 4, 11, 9)}.items():
        bm = BayesianModel(keras_json.make_sequential_json(D, units, acts))
        P = sum(int(np.prod(v.shape)) for v in bm._model.trainable_variables)
        W = rng.normal(0, 0.7, (n_stored, P)).astype(np.float32)
        if name == "reg":
            W[1, 0] = np.nan                                   # a NaN weight: its outputs are zeroed (BayesianModel.py:125)
        freq = rng.integers(1, 6, n_stored).tolist()
        bm.apply_distribution(Sampled([tf_shim.TT(torch.as_tensor(w)) for w in W], freq), 0, len(units) - 1)
        x = rng.normal(size=(Nt, D)).astype(np.float32)
        random.seed(7)
        samples, mean = bm.predict(tf_shim.TT(torch.as_tensor(x)), nb)
        random.seed(7)
        acc = np.cumsum(freq)
        tickets = [random.randint(1, int(acc[-1])) for _ in range(nb)]
        out.update({name + "_W": W, name + "_freq": np.asarray(freq), name + "_x": x, name + "_tickets": np.asarray(tickets),
                    name + "_samples": np.stack([s.numpy() for s in samples]), name + "_mean": mean.numpy()})
        print(name, "P =", P, "outputs", out[name + "_samples"].shape)
    np.savez_compressed(os.path.join(HERE, "reference_predict.npz"), **out)


if __name__ == "__main__":
    main()
