#!/usr/bin/env python
"""The plotting grid of the reference (Pyesian/visualisations/Plotter.py:121-135, `_extract_grid_x`) produced by running
THE REFERENCE's method; the handful of tf functions it uses on small arrays (reduce_max / reduce_min, range, meshgrid,
stack, reshape, matmul, transpose) are NumPy one-liners here — tf.range as documented: ceil(|limit - start| / |delta|)
values start + i * delta in the input dtype.

    python -B tests/golden/make_reference_grid_golden.py        # writes tests/golden/reference_grid.npz
"""
import os
import sys
import types
import warnings
from unittest.mock import MagicMock

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def numpy_tf():
    tf = MagicMock()

    def tf_range(start, limit, delta):
        start, limit, delta = np.asarray(start), np.asarray(limit), np.asarray(delta)
        n = int(np.ceil(np.abs((limit - start) / delta)))
        return start + np.arange(n).astype(start.dtype) * delta

    tf.range = tf_range
    tf.math = types.SimpleNamespace(reduce_max=lambda x, axis=None: np.max(x, axis=axis),
                                    reduce_min=lambda x, axis=None: np.min(x, axis=axis))
    tf.meshgrid = lambda a, b, indexing="xy": np.meshgrid(a, b, indexing=indexing)
    tf.stack = lambda vals, axis=0: np.stack(vals, axis=axis)
    tf.reshape = lambda t, shape: np.reshape(t, shape)
    tf.transpose = lambda t: np.transpose(t)
    tf.linalg = types.SimpleNamespace(matmul=lambda a, b: np.matmul(a, b))
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
    Plotter = sys.modules["Pyesian.visualisations.Plotter"].Plotter
    rng = np.random.default_rng(0)
    out = {}
    cases = [(np.float64, 1e-2, 0.2, np.eye(2)), (np.float32, 5e-2, 0.0, np.eye(2)),
             (np.float64, 2e-2, 0.5, np.linalg.qr(rng.normal(size=(5, 5)))[0][:, :2])]
    for i, (dt, gran, zoom, base) in enumerate(cases):
        x = (rng.normal(size=(60, 2)) * [2.0, 0.5] + [1.0, -3.0]).astype(dt)
        p = Plotter.__new__(Plotter)
        dim1, dim2, grid = p._extract_grid_x(x, base.astype(dt), gran, zoom)
        out.update({"c%d_x" % i: x, "c%d_base" % i: base.astype(dt), "c%d_args" % i: np.asarray([gran, zoom]),
                    "c%d_dim1" % i: np.asarray(dim1), "c%d_dim2" % i: np.asarray(dim2), "c%d_grid" % i: np.asarray(grid)})
        print(i, np.asarray(dim1).shape, np.asarray(grid).shape, np.asarray(grid).dtype)
    out["n_cases"] = np.int64(len(cases))
    np.savez_compressed(os.path.join(HERE, "reference_grid.npz"), **out)


if __name__ == "__main__":
    main()
