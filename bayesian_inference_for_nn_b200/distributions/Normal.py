"""Normal — independent Gaussian over a flat parameter vector; stands where the reference wraps
``tfp.distributions.Normal`` in ``TensorflowProbabilityDistribution`` (Pyesian/distributions/tf/
TensorflowProbabilityDistribution.py:9-59; built at SGLD.py:153-160 with ``scale = sq_mean - mean**2``, the VARIANCE
handed over as the standard deviation — kept by the caller, not here).  ``store``/``load`` use the JSON layout of the
reference's ``BaseSerializer`` (tf/BaseSerializer.py:19-34: ``{"type": "Normal", "params": {"loc": [...],
"scale": [...], ...}}``) so a folder written by either side loads on the other.  A NaN scale (negative "variance")
samples NaN, as tfp does; BayesianModel.predict zeroes NaN outputs (BayesianModel.py:125)."""
import json
import os

import numpy as np

from .Distribution import Distribution


class Normal(Distribution):
    def __init__(self, loc, scale, rng=None):
        loc = np.asarray(loc, dtype=np.float32).reshape(-1)
        scale = np.broadcast_to(np.asarray(scale, dtype=np.float32), loc.shape).copy()
        super().__init__(int(loc.shape[0]))
        self.loc, self.scale = loc, scale
        self._rng = rng if rng is not None else np.random.default_rng()

    def sample(self):
        z = self._rng.standard_normal(self._size).astype(np.float32)
        return self.loc + self.scale * z

    def store(self, path: str):
        data = {"type": "Normal", "params": {"loc": self.loc.tolist(), "scale": self.scale.tolist(),
                                             "validate_args": False, "allow_nan_stats": True, "name": "Normal"}}
        with open(os.path.join(path, "distribution.json"), "w") as f:
            f.write(json.dumps(data))

    @classmethod
    def load(cls, path: str) -> "Normal":
        with open(os.path.join(path, "distribution.json"), "r") as f:
            d = json.load(f)
        if d.get("type") != "Normal":
            raise ValueError("only tfp Normal distributions are supported, got %r" % (d.get("type"),))
        return Normal(d["params"]["loc"], d["params"]["scale"])


# the name BayesianModel.store writes into layers_config.txt for the reference's wrapper class
TensorflowProbabilityDistribution = Normal
