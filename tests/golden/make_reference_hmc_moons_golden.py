#!/usr/bin/env python
"""Statistics of the reference's HMC on its own example (HMC_classification.py:36-66: make_moons 2000 points, noise 0.2,
Dense(50, relu)-Dense(2, softmax), epsilon = 0.005, m = 0.5, L = 30), obtained by EXECUTING the reference's HMC.train /
HMC.step / HMC.result and BayesianModel.predict on the TensorFlow stand-in of tf_shim.py — once with the prior the
script ships, GaussianPrior(0.0, -1.0), and once with GaussianPrior(0.0, 1.0).  The randomness is the stand-in's
(tf.random.normal queue) and Python's `random`, both seeded; the GPU test compares STATISTICS (accept rate, loss level,
test accuracy), not trajectories.

    python -B tests/golden/make_reference_hmc_moons_golden.py        # ~2 minutes; writes tests/golden/reference_hmc_moons.npz
"""
import os
import random
import sys
import time
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import tf_shim  # noqa: E402
from make_reference_hmc_golden import load_reference  # noqa: E402


def main():
    import sklearn.datasets
    import torch
    HMC, GaussianPrior, HyperParameters = load_reference()
    from bayesian_inference_for_nn_b200 import keras_json          # host-side JSON writer only
    x, y = sklearn.datasets.make_moons(n_samples=2000, noise=0.2, random_state=0)
    perm = np.random.default_rng(0).permutation(2000)                    # Dataset: shuffle, 80 / 10 / 10 (Dataset.py:113-122)
    x, y = x[perm].astype(np.float32), y[perm].astype(np.int64)
    xtr, ytr, xte, yte = x[:1600], y[:1600], x[1600:1800], y[1600:1800]
    js = keras_json.make_sequential_json(2, [50, 2], ["relu", "softmax"])
    n_iter = 150
    out = {"x_train": xtr, "y_train": ytr, "x_test": xte, "y_test": yte, "hyper": np.asarray([0.005, 0.5, 30, n_iter])}
    for name, rho in (("shipped_neg", -1.0), ("pos", 1.0)):
        data = tf_shim.ArrayData(xtr, ytr)
        dataset = types.SimpleNamespace(training_dataset=lambda: data,
                                        loss=lambda reduction="auto": tf_shim.SparseCategoricalCrossentropy(reduction=reduction))
        tf_shim.RANDOM.rng = np.random.default_rng(1)
        random.seed(1)
        torch.manual_seed(1)
        losses = []

        class Recorder(HMC):
            def step(self, *a, **kw):
                r = super().step(*a, **kw)
                losses.append(float(r.numpy()))
                return r

        opt = Recorder()
        opt.compile(HyperParameters(epsilon=0.005, m=0.5, L=30), js, dataset, verbose=False, prior=GaussianPrior(0.0, rho))
        t0 = time.time()
        opt.train(n_iter)                                                # 10 always-accept burn-in + n_iter sampling
        bm = opt.result()
        random.seed(2)
        _, preds = bm.predict(tf_shim.TT(torch.as_tensor(xte)), nb_samples=100)
        acc = float((preds.numpy().argmax(1) == yte).mean())
        dist = bm._distributions[0]
        out.update({name + "_losses": np.asarray(losses), name + "_accept_rate": np.float64(opt._accepted_runs / opt._total_runs),
                    name + "_accuracy": np.float64(acc), name + "_n_samples": np.int64(len(dist._samples)),
                    name + "_frequencies": np.asarray(dist._frequencies, dtype=np.int64)})
        print("%s: accept rate %.3f, %d samples, loss %.4f -> %.4f, test accuracy %.3f  (%.0f s)"
              % (name, opt._accepted_runs / opt._total_runs, len(dist._samples), losses[0], np.mean(losses[-30:]), acc,
                 time.time() - t0))
    np.savez_compressed(os.path.join(HERE, "reference_hmc_moons.npz"), **out)


if __name__ == "__main__":
    main()
