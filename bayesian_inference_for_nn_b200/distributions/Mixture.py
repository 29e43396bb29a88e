"""Mixture — equal-weight mixture of same-sized distributions: the posterior of S independent SGLD / SWAG chains
(one component per chain; with one chain it is never used and ``result()`` has exactly the reference's structure)."""
import os

import numpy as np

from .Distribution import Distribution


class Mixture(Distribution):
    def __init__(self, components, rng=None):
        if len(components) == 0:
            raise ValueError("Can't have a Mixture with 0 components")
        if len({c.size() for c in components}) != 1:
            raise ValueError("Mixture components must have the same size")
        super().__init__(components[0].size())
        self._components = list(components)
        self._rng = rng if rng is not None else np.random.default_rng()

    @property
    def components(self):
        return list(self._components)

    def sample(self):
        return self._components[int(self._rng.integers(len(self._components)))].sample()

    def store(self, path: str):
        with open(os.path.join(path, "mixture.txt"), "w") as f:
            f.write("%d\n%s\n" % (len(self._components), self._components[0].__class__.__name__))
        for i, c in enumerate(self._components):
            os.makedirs(os.path.join(path, "component%d" % i), exist_ok=True)
            c.store(os.path.join(path, "component%d" % i))

    @classmethod
    def load(cls, path: str) -> "Mixture":
        from . import Normal, MultivariateNormalDiagPlusLowRank, Sampled
        reg = {"Normal": Normal, "MultivariateNormalDiagPlusLowRank": MultivariateNormalDiagPlusLowRank,
               "Sampled": Sampled}
        with open(os.path.join(path, "mixture.txt"), "r") as f:
            n, name = int(f.readline()), f.readline().strip()
        return Mixture([reg[name].load(os.path.join(path, "component%d" % i)) for i in range(n)])
