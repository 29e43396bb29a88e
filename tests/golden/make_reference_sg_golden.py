#!/usr/bin/env python
"""Golden SGLD / SWAG runs produced by EXECUTING THE REFERENCE's Pyesian/optimizers/SGLD.py and SWAG.py (compile, train's
schedule set-up, step, _init_*arrays — unmodified) on the torch-backed TensorFlow stand-in of tf_shim.py: minibatch
order, the Langevin / SGD update as written (noise drawn with stddev = lr and multiplied by lr again), the running
moments weighted by the step index, the deviation matrix (append, then overwrite the last column), the returned losses.

    python -B tests/golden/make_reference_sg_golden.py      # writes tests/golden/reference_sg.npz
"""
import os
import sys
import types
import warnings
from unittest.mock import MagicMock

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import tf_shim  # noqa: E402


def flat(model):
    return np.concatenate([v.numpy().reshape(-1) for v in model.trainable_variables]).astype(np.float32)


def per_layer(tensors):
    return np.concatenate([t.numpy() for t in tensors], axis=0)         # [P, cols]: layers stacked in model order


def main():
    sys.dont_write_bytecode = True
    warnings.simplefilter("ignore")
    sys.modules["tensorflow"] = tf_shim.make_tf()
    sys.modules["tensorflow_probability"] = tf_shim.make_tfp()
    for name in ["wandb", "wandb.integration", "wandb.integration.keras", "tensorflow_datasets", "ucimlrepo", "matplotlib",
                 "matplotlib.pyplot", "scikitplot"]:
        sys.modules.setdefault(name, MagicMock())
    sys.path.insert(0, "/root/reference")
    import Pyesian.optimizers  # noqa: F401
    from Pyesian.optimizers.hyperparameters import HyperParameters
    SGLD, SWAG = sys.modules["Pyesian.optimizers.SGLD"].SGLD, sys.modules["Pyesian.optimizers.SWAG"].SWAG
    from bayesian_inference_for_nn_b200 import keras_json          # host-side JSON writer only
    import torch

    D, units, acts, N, B, steps = 3, [5, 2], ["relu", "softmax"], 50, 16, 9
    rng = np.random.default_rng(5)
    X = rng.normal(size=(N, D)).astype(np.float32)
    y = rng.integers(0, 2, N).astype(np.int64)
    js = keras_json.make_sequential_json(D, units, acts)
    dataset = types.SimpleNamespace(training_dataset=lambda: tf_shim.ArrayData(X, y),
                                    loss=lambda reduction="auto": tf_shim.SparseCategoricalCrossentropy(reduction=reduction))
    out = {"X": X, "y": y, "meta": np.asarray([D, N, B, steps], dtype=np.int64)}

    # ---- SGLD
    torch.manual_seed(0)
    opt = SGLD()
    opt.compile(HyperParameters(batch_size=B, lr_upper=0.2, lr_lower=0.02, lr_gamma=0.55), js, dataset, verbose=False)
    opt._nb_iterations = steps
    opt._init_sgld_lr()                                     # what train() does before its loop (SGLD.py:124-126)
    out["sgld_theta0"] = flat(opt._base_model)
    tf_shim.RANDOM.rng = np.random.default_rng(11)
    th, zs, rets, lrs = [], [], [], []
    for s in range(steps):
        tf_shim.RANDOM.log.clear()
        lrs.append(float(opt._lr(opt._n)))
        rets.append(float(opt.step().numpy()))
        zs.append(np.concatenate([z.reshape(-1) for z in tf_shim.RANDOM.log]))
        th.append(flat(opt._base_model))
    out.update(sgld_theta=np.stack(th), sgld_z=np.stack(zs), sgld_ret=np.asarray(rets), sgld_lr=np.asarray(lrs),
               sgld_mean=per_layer(opt._mean)[:, 0], sgld_sq_mean=per_layer(opt._sq_mean)[:, 0],
               sgld_dev_cols=np.int64(opt._dev[0].shape[1]))

    # ---- SWAG (k = 3, frequency = 2: 5 moment updates in 9 steps, so the last column is overwritten twice)
    torch.manual_seed(1)
    start = tf_shim.model_from_json(js)
    opt = SWAG()
    opt.compile(HyperParameters(batch_size=B, lr=0.1, k=3, scale=0.5, frequency=2), js, dataset, verbose=False,
                starting_model=start)
    out["swag_theta0"] = flat(opt._base_model)
    assert np.array_equal(out["swag_theta0"], flat(start))
    th, rets = [], []
    for s in range(steps):
        rets.append(float(opt.step().numpy()))
        th.append(flat(opt._base_model))
    out.update(swag_theta=np.stack(th), swag_ret=np.asarray(rets), swag_mean=per_layer(opt._mean)[:, 0],
               swag_sq_mean=per_layer(opt._sq_mean)[:, 0], swag_dev=per_layer(opt._dev), swag_hyper=np.asarray([0.1, 3, 2]))
    np.savez_compressed(os.path.join(HERE, "reference_sg.npz"), **out)
    print("SGLD returns", np.round(rets[:2], 4), "dev cols", int(out["sgld_dev_cols"]), "| SWAG dev", out["swag_dev"].shape)


if __name__ == "__main__":
    main()
