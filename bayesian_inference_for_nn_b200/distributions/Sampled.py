"""Sampled — the empirical posterior HMC / SVGD return: flat weight vectors with integer multiplicities.

Behavioural mirror of Pyesian/distributions/Sampled.py:8-60.  A draw picks sample k with probability proportional to
its frequency, by the reference's own recipe — ``random.randint(1, total)`` looked up in the running totals with
``bisect_left`` (:29-32) — so a seeded ``random`` reproduces the reference's choices.  The constructor rejects the same
inputs with the same messages (:11-16, :24-25).  The folder format is the reference's (:34-60): ``info.json`` with
``size / n_samples / frequencies / dtypes`` and one TensorProto file ``samples/sample<i>.tf`` per sample.

Unlike the reference, which keeps a Python list of tensors, the samples live in ONE ``[n, P]`` float32 matrix: that is
what the device hands back and what ``BayesianModel.predict`` uploads once and gathers rows from.
"""
import json
import os
import random
from bisect import bisect_left
from itertools import accumulate

import numpy as np

from .Distribution import Distribution


def _as_matrix(samples):
    if isinstance(samples, np.ndarray):
        return np.ascontiguousarray(samples)
    return np.ascontiguousarray(np.stack([np.asarray(s) for s in samples]))


class Sampled(Distribution):
    def __init__(self, samples, frequencies):
        n = len(samples)
        if n == 0:
            raise ValueError("Can't have distribution Sampled with 0 samples")
        if n != len(frequencies):
            raise ValueError("Number of samples and list frequency do not have the same size")
        matrix = _as_matrix(samples)
        if matrix.ndim != 2:
            raise ValueError("Samples must have only one dimension")
        counts = [int(f) for f in frequencies]
        if not all(counts):
            raise ValueError("Samples frequencies can't sum up to zero")
        super().__init__(int(matrix.shape[1]))
        self._samples, self._n_samples, self._frequencies = matrix, n, counts
        self._acc_frequencies = list(accumulate(counts))          # running totals: the draw's lookup table

    # ---- drawing ---------------------------------------------------------------------------------------------------
    def sample_index(self) -> int:
        ticket = random.randint(1, self._acc_frequencies[-1])
        return bisect_left(self._acc_frequencies, ticket)

    def sample(self):
        return self._samples[self.sample_index()]

    # ---- whole-posterior views for the device path ----------------------------------------------------------------
    @property
    def samples(self):
        return self._samples

    @property
    def frequencies(self):
        return list(self._frequencies)

    # ---- persistence ----------------------------------------------------------------------------------------------
    def store(self, path: str):
        from ..nn.tensorproto import serialize_tensor
        meta = dict(size=self._size, n_samples=self._n_samples, frequencies=self._frequencies,
                    dtypes=[self._samples.dtype.name] * self._n_samples)
        with open(os.path.join(path, "info.json"), "w") as f:
            json.dump(meta, f)
        folder = os.path.join(path, "samples")
        os.makedirs(folder, exist_ok=True)
        for i, row in enumerate(self._samples):
            with open(os.path.join(folder, "sample%d.tf" % i), "wb") as f:
                f.write(serialize_tensor(row))

    @classmethod
    def load(cls, path: str) -> "Sampled":
        from ..nn.tensorproto import parse_tensor
        with open(os.path.join(path, "info.json")) as f:
            meta = json.load(f)
        folder = os.path.join(path, "samples")

        def read(i):
            with open(os.path.join(folder, "sample%d.tf" % i), "rb") as f:
                return parse_tensor(f.read())
        return cls(np.stack([read(i) for i in range(meta["n_samples"])]), meta["frequencies"])
