#!/usr/bin/env python
"""SVGD_mnist's model shape (784-128-10, BASELINE configs[3]) through the reference's own SVGD.step at a reduced scale
(5 particles, minibatch 256, 3 steps; prior scale 0.002 so that the particles are close enough for the gamma = 1 kernel
terms to matter), executed on the TensorFlow stand-in of tf_shim.py.  It lets the device's tensor-core gradient path for
128 hidden units (two particles per CTA pair in the dW1 GEMM, an odd particle count) be compared directly with what
Pyesian computes.  Inputs are re-created from their seeds by inputs(); only results are stored.

    python -B tests/golden/make_reference_wide_svgd_golden.py      # writes tests/golden/reference_wide_svgd.npz
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import tf_shim  # noqa: E402
from make_reference_svgd_golden import load_reference  # noqa: E402

D, H, C, N, B, M, STEPS, LR, SCALE = 784, 128, 10, 768, 256, 5, 3, 1e-3, 0.002
SHAPES = [(D, H), (H,), (H, C), (C,)]


def inputs():
    """every input of the run, re-created from its seed (the tests call this too; nothing here needs the reference)"""
    rng = np.random.default_rng(0)
    X = rng.random((N, D), dtype=np.float32)
    y = rng.integers(0, C, N).astype(np.int64)
    prng = np.random.default_rng(21)                 # the stand-in's normal queue: one prior draw per variable per particle
    p0 = np.stack([np.concatenate([(prng.standard_normal(s).astype(np.float32) * np.float32(SCALE)).reshape(-1) for s in SHAPES])
                   for _ in range(M)]).astype(np.float64)
    return dict(X=X, y=y, particles0=p0)


def main():
    SVGD, GaussianPrior, HyperParameters = load_reference()
    from bayesian_inference_for_nn_b200 import keras_json          # host-side JSON writer only
    inp = inputs()
    data = tf_shim.ArrayData(inp["X"], inp["y"])
    dataset = types.SimpleNamespace(training_dataset=lambda: data, valid_data=tf_shim.ArrayData(inp["X"][:8], inp["y"][:8]),
                                    valid_size=8,
                                    loss=lambda reduction="auto": tf_shim.SparseCategoricalCrossentropy(reduction=reduction))
    tf_shim.RANDOM.rng = np.random.default_rng(21)
    opt = SVGD()
    opt.compile(HyperParameters(batch_size=B, M=M, lr=LR), keras_json.make_sequential_json(D, [H, C], ["relu", "softmax"]),
                dataset, verbose=False, prior=GaussianPrior(0.0, SCALE))
    assert np.array_equal(opt._particles, inp["particles0"])
    rets, after = [], []
    for s in range(STEPS):
        rets.append(float(opt.step().numpy()))
        after.append(opt._particles.copy())
    d01 = float(np.sum((inp["particles0"][0] - inp["particles0"][1]) ** 2))
    print("P =", opt._particles.shape[1], "losses", np.round(rets, 5), "K_01 at start = exp(-%.3f) = %.3f" % (d01, np.exp(-d01)))
    np.savez_compressed(os.path.join(HERE, "reference_wide_svgd.npz"), ret=np.asarray(rets),
                        # the first legacy-Adam step is -lr * phi / (|phi| + eps'): its sign is the sign of -phi, element by element
                        step1_sign=np.sign(after[0] - inp["particles0"]).astype(np.int8),
                        step1_absmax=np.float64(np.abs(after[0] - inp["particles0"]).max()),
                        final=after[-1].astype(np.float32))


if __name__ == "__main__":
    main()
