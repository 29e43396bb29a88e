#!/usr/bin/env python
"""Golden SVGD steps produced by EXECUTING THE REFERENCE's Pyesian/optimizers/SVGD.py (compile_extra_components,
_init_particles, step, _svgd_gradients, rbf_kernel, unflatten_gradients, _pack_weights, _unpack_weights — unmodified) on
the torch-backed TensorFlow stand-in of tf_shim.py.  Pinned by the reference's own code: the sequential sweep over the
particles against the partly updated set, the float64 kernel with gamma = 1 and its autograd gradient, the broadcast of ONE
particle's gradient, phi / M, the per-particle legacy-Adam DESCENT step, float64 particle storage, the loss bookkeeping
(every 10 steps) and the validation pass.  Keras / Adam numerics come from their definitions (tf_shim.py).

    python -B tests/golden/make_reference_svgd_golden.py        # writes tests/golden/reference_svgd.npz
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


def load_reference():
    sys.dont_write_bytecode = True
    warnings.simplefilter("ignore")
    sys.modules["tensorflow"] = tf_shim.make_tf()
    sys.modules["tensorflow_probability"] = tf_shim.make_tfp()
    for name in ["wandb", "wandb.integration", "wandb.integration.keras", "tensorflow_datasets", "ucimlrepo", "matplotlib",
                 "matplotlib.pyplot", "scikitplot"]:
        sys.modules.setdefault(name, MagicMock())
    sys.path.insert(0, "/root/reference")
    import Pyesian.optimizers  # noqa: F401
    import Pyesian.distributions  # noqa: F401
    from Pyesian.optimizers.hyperparameters import HyperParameters
    return (sys.modules["Pyesian.optimizers.SVGD"].SVGD, sys.modules["Pyesian.distributions.GaussianPrior"].GaussianPrior,
            HyperParameters)


def run_case(SVGD, GaussianPrior, HyperParameters, name, D, units, acts, N, Nv, B, loss, M, lr, steps, seed, scale):
    from bayesian_inference_for_nn_b200 import keras_json          # host-side JSON writer only
    rng = np.random.default_rng(seed)
    X = rng.normal(size=(N, D)).astype(np.float32)
    Xv = rng.normal(size=(Nv, D)).astype(np.float32)
    if loss == "ce":
        y, yv = rng.integers(0, units[-1], N).astype(np.int64), rng.integers(0, units[-1], Nv).astype(np.int64)
        loss_cls = tf_shim.SparseCategoricalCrossentropy
    else:
        y, yv = rng.normal(size=(N, units[-1])).astype(np.float32), rng.normal(size=(Nv, units[-1])).astype(np.float32)
        loss_cls = tf_shim.MeanSquaredError
    data = tf_shim.ArrayData(X, y)
    dataset = types.SimpleNamespace(training_dataset=lambda: data, valid_data=tf_shim.ArrayData(Xv, yv), valid_size=Nv,
                                    loss=lambda reduction="auto": loss_cls(reduction=reduction))
    tf_shim.RANDOM.rng = np.random.default_rng(seed + 100)
    opt = SVGD()
    opt.compile(HyperParameters(batch_size=B, M=M, lr=lr), keras_json.make_sequential_json(D, units, acts), dataset,
                verbose=False, prior=GaussianPrior(0.0, scale))
    P = opt._particles.shape[1]
    before, after, ret = [], [], []
    for s in range(steps):
        before.append(opt._particles.copy())
        ret.append(float(opt.step().numpy()))
        after.append(opt._particles.copy())
    out = {name + "_X": X, name + "_y": y, name + "_Xv": Xv, name + "_yv": yv,
           name + "_before": np.stack(before), name + "_after": np.stack(after), name + "_ret": np.asarray(ret),
           name + "_train_losses": np.asarray([float(v.numpy()) for v in opt.train_losses]),
           name + "_valid_losses": np.asarray([float(v.numpy()) for v in opt.valid_losses]),
           name + "_meta": np.asarray([D, N, Nv, B, M, steps, P], dtype=np.int64), name + "_hyper": np.asarray([lr, scale])}
    ens, tl, vl = opt.result()
    out[name + "_result_weights"] = np.stack([np.concatenate([v.numpy().reshape(-1) for v in mdl.trainable_variables])
                                              for mdl in ens])
    print(name, "P =", P, "dtype", opt._particles.dtype, "losses", np.round(ret[:3], 4), "...", np.round(ret[-1], 4),
          "| recorded", len(opt.train_losses))
    return out


def main():
    SVGD, GaussianPrior, HyperParameters = load_reference()
    out = {}
    # close particles (prior scale 0.05): the kernel terms matter; 12 steps so that the 10-step bookkeeping fires once
    out.update(run_case(SVGD, GaussianPrior, HyperParameters, "ce", 2, [4, 2], ["relu", "softmax"], 48, 20, 16, "ce", 4,
                        1e-2, 12, seed=1, scale=0.05))
    # N(0,1) particles in a larger space: off-diagonal kernel entries underflow, phi_i = g_i / M
    out.update(run_case(SVGD, GaussianPrior, HyperParameters, "mse", 3, [6, 1], ["tanh", "linear"], 30, 10, 10, "mse", 3,
                        5e-3, 4, seed=2, scale=1.0))
    np.savez_compressed(os.path.join(HERE, "reference_svgd.npz"), **out)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
