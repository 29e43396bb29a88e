"""MultivariateNormalDiagPlusLowRank — SWAG's posterior (Pyesian/distributions/MultivariateNormalDiagPlusLowRank.py:10-41).

``sample()`` follows :32-41 term by term: ``z1 ~ N(0, diag)`` with ``diag`` used as the standard deviation (SWAG.py:129
passes ``sq_mean - mean**2``, the variance — kept by the caller), ``z2 ~ N(0, I_k)``, result
``mean + z1 + D z2 * sqrt(1 / (2 (k - 1)))``.  ``store``/``load`` use the reference's ``distribution.json`` keys
(``mean``, ``D``, ``diag``; :11-24)."""
import json
import os
from math import sqrt

import numpy as np

from .Distribution import Distribution


class MultivariateNormalDiagPlusLowRank(Distribution):
    def __init__(self, mean, diag, D, rng=None):
        mean = np.asarray(mean, dtype=np.float32).reshape(-1)
        super().__init__(int(mean.shape[0]))
        self._mean = mean
        self._diag = np.asarray(diag, dtype=np.float32).reshape(-1)
        self._D = np.asarray(D, dtype=np.float32).reshape(mean.shape[0], -1)
        self._rng = rng if rng is not None else np.random.default_rng()

    def sample(self):
        k = self._D.shape[1]
        z1 = self._diag * self._rng.standard_normal(self._size).astype(np.float32)
        z2 = self._rng.standard_normal(k).astype(np.float32)
        with np.errstate(divide="ignore", invalid="ignore"):
            cov_mean = (self._D @ z2) * np.float32(sqrt(1 / (2 * (k - 1))) if k != 1 else np.inf)
        return self._mean + z1 + cov_mean

    def store(self, path: str):
        data = {"mean": self._mean.tolist(), "D": self._D.tolist(), "diag": self._diag.tolist()}
        with open(os.path.join(path, "distribution.json"), "w") as f:
            f.write(json.dumps(data))

    @classmethod
    def load(cls, path: str) -> "MultivariateNormalDiagPlusLowRank":
        with open(os.path.join(path, "distribution.json"), "r") as f:
            d = json.load(f)
        return MultivariateNormalDiagPlusLowRank(d["mean"], d["diag"], d["D"])
