#!/usr/bin/env python
"""The reference's SVGD on its own example (SVGD_classification.py:150-166: make_moons, Dense(64, relu)-Dense(2,
softmax), M = 10 particles, batch 64, lr = 1e-3, prior N(0, 1)), obtained by EXECUTING the reference's SVGD.compile /
train / step / result on the TensorFlow stand-in of tf_shim.py.  After the initial particles are drawn the live step is
deterministic (no randomness; the stand-in keeps the minibatch order fixed), so the golden is a TRAJECTORY: the initial
particles, every step's returned loss, the particles at checkpoints and at the end, the recorded train / validation
losses and the ensemble's test accuracy.

    python -B tests/golden/make_reference_svgd_moons_golden.py     # ~1 minute; writes tests/golden/reference_svgd_moons.npz
"""
import os
import sys
import time
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import tf_shim  # noqa: E402
from make_reference_svgd_golden import load_reference  # noqa: E402


def main():
    import sklearn.datasets
    import torch
    SVGD, GaussianPrior, HyperParameters = load_reference()
    from bayesian_inference_for_nn_b200 import keras_json          # host-side JSON writer only
    x, y = sklearn.datasets.make_moons(n_samples=2000, noise=0.2, random_state=0)
    perm = np.random.default_rng(0).permutation(2000)                    # Dataset: shuffle, 80 / 10 / 10 (Dataset.py:113-122)
    x, y = x[perm].astype(np.float32), y[perm].astype(np.int64)
    xtr, ytr, xte, yte, xva, yva = x[:1600], y[:1600], x[1600:1800], y[1600:1800], x[1800:], y[1800:]
    steps, M, B, lr = 400, 10, 64, 1e-3
    data = tf_shim.ArrayData(xtr, ytr)
    dataset = types.SimpleNamespace(training_dataset=lambda: data, valid_data=tf_shim.ArrayData(xva, yva), valid_size=200,
                                    loss=lambda reduction="auto": tf_shim.SparseCategoricalCrossentropy(reduction=reduction))
    tf_shim.RANDOM.rng = np.random.default_rng(4)
    torch.manual_seed(4)
    rets, checkpoints = [], {}

    class Recorder(SVGD):
        def step(self, *a, **kw):
            r = super().step(*a, **kw)
            rets.append(float(r.numpy()))
            if self._step in (1, 10, 50, 100, 200, steps):
                checkpoints[self._step] = self._particles.copy()
            return r

    opt = Recorder()
    opt.compile(HyperParameters(lr=lr, batch_size=B, M=M), keras_json.make_sequential_json(2, [64, 2], ["relu", "softmax"]),
                dataset, verbose=False, prior=GaussianPrior(0, 1))
    particles0 = opt._particles.copy()
    t0 = time.time()
    opt.train(steps)
    models, train_losses, valid_losses = opt.result()
    probs = np.mean([mdl(tf_shim.TT(torch.as_tensor(xte))).numpy() for mdl in models], axis=0)
    acc = float((probs.argmax(1) == yte).mean())
    out = {"x_train": xtr, "y_train": ytr, "x_test": xte, "y_test": yte, "x_valid": xva, "y_valid": yva,
           "hyper": np.asarray([steps, M, B, lr]), "particles0": particles0, "ret": np.asarray(rets),
           "train_losses": np.asarray([float(v.numpy()) for v in train_losses]),
           "valid_losses": np.asarray([float(v.numpy()) for v in valid_losses]), "accuracy": np.float64(acc)}
    for k, v in checkpoints.items():
        out["particles_%d" % k] = v
    np.savez_compressed(os.path.join(HERE, "reference_svgd_moons.npz"), **out)
    print("P = %d, loss %.4f -> %.4f, valid %.4f -> %.4f, ensemble test accuracy %.3f  (%.0f s)"
          % (particles0.shape[1], rets[0], rets[-1], out["valid_losses"][0], out["valid_losses"][-1], acc, time.time() - t0))


if __name__ == "__main__":
    main()
