"""Sampled — empirical posterior: flat weight vectors with integer frequencies.

Mirrors Pyesian/distributions/Sampled.py:8-60: weighted draw through ``random.randint(1, total)``
+ ``bisect_left`` on the cumulative frequencies (:29-32) so that P(index) is proportional to its
frequency; the same validation errors (:11-16, :24-25).  Samples are kept as one ``[n, P]``
float32 matrix (what the device returns) instead of a list of tensors.
"""
import bisect
import json
import os
import random

import numpy as np

from .Distribution import Distribution


class Sampled(Distribution):
    def __init__(self, samples, frequencies):
        if len(samples) == 0:
            raise ValueError("Can't have distribution Sampled with 0 samples")
        if len(samples) != len(frequencies):
            raise ValueError("Number of samples and list frequency do not have the same size")
        mat = np.ascontiguousarray(np.stack([np.asarray(s) for s in samples]) if not isinstance(samples, np.ndarray)
                                   else samples)
        if mat.ndim != 2:
            raise ValueError("Samples must have only one dimension")
        super().__init__(int(mat.shape[1]))
        self._n_samples = int(mat.shape[0])
        self._samples = mat
        self._frequencies = [int(f) for f in frequencies]
        self._acc_frequencies = []
        acc = 0
        for f in self._frequencies:
            if f == 0:
                raise ValueError("Samples frequencies can't sum up to zero")
            acc += f
            self._acc_frequencies.append(acc)

    # reference semantics ------------------------------------------------------------------
    def sample_index(self) -> int:
        w = random.randint(1, self._acc_frequencies[-1])
        return bisect.bisect_left(self._acc_frequencies, w)

    def sample(self):
        return self._samples[self.sample_index()]

    # batched views used by BayesianModel.predict --------------------------------------------
    @property
    def samples(self):
        return self._samples

    @property
    def frequencies(self):
        return list(self._frequencies)

    # persistence (format of Sampled.store :34-48; the per-sample payload is a TensorProto) ----
    def store(self, path: str):
        from ..nn.tensorproto import serialize_tensor
        info = {"size": self._size, "n_samples": self._n_samples, "frequencies": self._frequencies,
                "dtypes": [str(self._samples.dtype.name)] * self._n_samples}
        with open(os.path.join(path, "info.json"), "w") as f:
            f.write(json.dumps(info))
        sdir = os.path.join(path, "samples")
        os.makedirs(sdir, exist_ok=True)
        for i in range(self._n_samples):
            with open(os.path.join(sdir, "sample%d.tf" % i), "wb") as f:
                f.write(serialize_tensor(self._samples[i]))

    @classmethod
    def load(cls, path: str) -> "Sampled":
        from ..nn.tensorproto import parse_tensor
        with open(os.path.join(path, "info.json"), "r") as f:
            info = json.load(f)
        sdir = os.path.join(path, "samples")
        rows = []
        for i in range(info["n_samples"]):
            with open(os.path.join(sdir, "sample%d.tf" % i), "rb") as f:
                rows.append(parse_tensor(f.read()))
        return Sampled(np.stack(rows), info["frequencies"])
