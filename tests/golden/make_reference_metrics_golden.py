#!/usr/bin/env python
"""Golden vectors for Metrics.classification_uncertainty produced by running THE REFERENCE's own method
(Pyesian/visualisations/Metrics.py:344-375) in the build container.  TensorFlow is not installable here; the method uses
exactly six tf functions on small arrays — tf.stack, tf.reshape, tf.transpose, tf.matmul, tf.linalg.diag, tf.one_hot —
which are supplied by the NumPy one-liners below (their meaning is not in question; the arrays they return rebind on
`+=` like immutable tensors); the loop nest, the [C,1] - [C] broadcast inside the epistemic term, the accumulators
that are never reset between rows, the sum over draws and the division by the n_samples ARGUMENT all execute as written
in the reference.  The prediction cache (`_get_predictions`) and the dataset access (`_get_x_y`) are replaced on the
instance by functions that hand over the fixed per-draw probabilities.

    python -B tests/golden/make_reference_metrics_golden.py      # writes tests/golden/reference_metrics.npz
"""
import os
import sys
import types
import warnings
from unittest.mock import MagicMock

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


class Immutable(np.ndarray):
    """tf.Tensor is immutable: `acc += t` REBINDS acc to a new tensor, so the per-row values the reference appends to its
    lists stay distinct (a plain ndarray would be updated in place and every list entry would alias the final sum)."""

    def __iadd__(self, other):
        return np.add(self, other).view(Immutable)


def T(a):
    return np.asarray(a).view(Immutable)


def numpy_tf():
    tf = MagicMock()
    tf.stack = lambda values, axis=0: T(np.stack([np.asarray(v) for v in values], axis=axis))
    tf.reshape = lambda t, shape: T(np.reshape(np.asarray(t), shape))
    tf.transpose = lambda t: T(np.transpose(np.asarray(t)))
    tf.matmul = lambda a, b: T(np.matmul(np.asarray(a), np.asarray(b)))
    tf.one_hot = lambda idx, depth: T(np.eye(int(depth), dtype=np.float32)[int(idx)])
    tf.linalg = types.SimpleNamespace(diag=lambda v: T(np.diag(np.asarray(v))), matmul=tf.matmul)
    return tf


def main():
    sys.dont_write_bytecode = True
    warnings.simplefilter("ignore")
    sys.modules["tensorflow"] = numpy_tf()
    for name in ["tensorflow_probability", "wandb", "wandb.integration", "wandb.integration.keras", "tensorflow_datasets",
                 "ucimlrepo", "matplotlib", "matplotlib.pyplot", "scikitplot"]:
        sys.modules.setdefault(name, MagicMock())
    sys.path.insert(0, "/root/reference")
    import Pyesian.visualisations  # noqa: F401
    Metrics = sys.modules["Pyesian.visualisations.Metrics"].Metrics

    rng = np.random.default_rng(0)
    out = {}
    cases = [(1, 1, 2, 100), (3, 7, 2, 100), (5, 33, 4, 33), (4, 60, 10, 100), (2, 300, 3, 250)]
    for i, (n, N, C, n_samples_arg) in enumerate(cases):
        z = rng.normal(size=(n, N, C))
        probs = np.exp(z - z.max(-1, keepdims=True))
        probs = (probs / probs.sum(-1, keepdims=True)).astype(np.float32)
        y = rng.integers(0, C, N)
        m = Metrics.__new__(Metrics)
        m._dataset = types.SimpleNamespace(likelihood_model="Classification")
        m._get_x_y = lambda n_samples=100, data_type="test": (None, y)
        m._get_predictions = lambda inp, n_boundaries, y_true: ([probs[k] for k in range(n)], probs.mean(axis=0), y_true, inp)
        total, aleatoric, epistemic = m.classification_uncertainty(n_boundaries=n, n_samples=n_samples_arg)
        out["c%d_probs" % i], out["c%d_y" % i], out["c%d_arg" % i] = probs, y, np.int64(n_samples_arg)
        out["c%d_total" % i], out["c%d_aleatoric" % i], out["c%d_epistemic" % i] = (np.asarray(total), np.asarray(aleatoric),
                                                                                   np.asarray(epistemic))
    out["n_cases"] = np.int64(len(cases))
    np.savez_compressed(os.path.join(HERE, "reference_metrics.npz"), **out)
    print("wrote", len(cases), "cases; shapes", [out["c%d_total" % i].shape for i in range(len(cases))])


if __name__ == "__main__":
    main()
