"""Mixture — equal-weight mixture of same-sized distributions: the posterior of S independent SGLD / SWAG chains
(one component per chain; with one chain it is never used and ``result()`` has exactly the reference's structure).

The reference applies one distribution per Dense layer (SGLD.py:147-165, SWAG.py:119-139) and ``BayesianModel`` draws
every layer interval on its own.  Independent chains are not exchangeable layer by layer (hidden units permute and
rescale between chains), so the per-layer mixtures of ONE posterior share a ``ChainSelector``: a drawn network takes
all its layers from the same chain."""
import os
import uuid

import numpy as np

from .Distribution import Distribution


class ChainSelector:
    """The component index shared by the per-layer mixtures of one multi-chain posterior.  ``BayesianModel`` calls
    ``new_draw()`` once per drawn weight vector; every linked ``Mixture.sample()`` then uses the same index."""
    _registry = {}

    def __init__(self, n, rng=None, group=None):
        self.n = int(n)
        self._rng = rng if rng is not None else np.random.default_rng()
        self.group = group if group is not None else uuid.uuid4().hex
        self.index = None

    def new_draw(self):
        self.index = int(self._rng.integers(self.n))
        return self.index

    @classmethod
    def for_group(cls, group, n):
        sel = cls._registry.get(group)
        if sel is None or sel.n != n:
            sel = cls._registry[group] = ChainSelector(n, group=group)
        return sel


class Mixture(Distribution):
    def __init__(self, components, rng=None, selector=None):
        if len(components) == 0:
            raise ValueError("Can't have a Mixture with 0 components")
        if len({c.size() for c in components}) != 1:
            raise ValueError("Mixture components must have the same size")
        if selector is not None and selector.n != len(components):
            raise ValueError("the selector and the mixture must have the same number of components")
        super().__init__(components[0].size())
        self._components = list(components)
        self._rng = rng if rng is not None else np.random.default_rng()
        self.selector = selector

    @property
    def components(self):
        return list(self._components)

    def sample(self):
        if self.selector is not None:
            k = self.selector.index if self.selector.index is not None else self.selector.new_draw()
            return self._components[k].sample()
        return self._components[int(self._rng.integers(len(self._components)))].sample()

    def store(self, path: str):
        with open(os.path.join(path, "mixture.txt"), "w") as f:
            f.write("%d\n%s\n" % (len(self._components), self._components[0].__class__.__name__))
            if self.selector is not None:
                f.write("group=%s\n" % self.selector.group)
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
            group = f.readline().strip()
        sel = ChainSelector.for_group(group[len("group="):], n) if group.startswith("group=") else None
        return Mixture([reg[name].load(os.path.join(path, "component%d" % i)) for i in range(n)], selector=sel)
