"""SWAG — SGD iterates summarised as a diagonal + low-rank Gaussian, ``n_chains`` runs per minibatch step on the device.

Drop-in for Pyesian/optimizers/SWAG.py:14-150.  Same surface: hyper-parameters ``batch_size, lr, k,
scale, frequency`` (:100-107), ``compile(..., starting_model=...)`` — a Keras model (anything with ``get_weights()``),
a list of weight arrays in ``get_weights()`` order, or a flat ``[P]`` / ``[n_chains, P]`` array — (:104-105, ``KeyError``
without it), ``step()`` returns the minibatch loss (:94), ``result()`` is a ``BayesianModel`` with one
``MultivariateNormalDiagPlusLowRank(mean, sq_mean - mean**2, sqrt(scale / (k - 1)) * dev)`` per weight layer (:119-139).
Reference arithmetic kept: plain SGD (:62-64), moments every ``frequency`` steps weighted by the STEP index (:72-80), the
deviation matrix grows to ``k`` columns and then only its last column is replaced (:83-89)."""
from math import sqrt

import numpy as np

from .. import _lib
from ..distributions import MultivariateNormalDiagPlusLowRank
from ._sgchains import StochasticGradientChains


class SWAG(StochasticGradientChains):
    KIND = _lib.SG_SWAG

    def _flatten_start(self, start):
        P = self._spec.n_params
        if hasattr(start, "get_weights"):
            start = start.get_weights()
        if isinstance(start, (list, tuple)):
            start = np.concatenate([np.asarray(w, dtype=np.float32).reshape(-1) for w in start])
        start = np.asarray(start, dtype=np.float32)
        if start.ndim == 1:
            start = start.reshape(1, -1)
        if start.shape[1] != P or start.shape[0] not in (1, self._n_chains):
            raise ValueError("starting_model has %s weights, the model needs %d" % (start.shape, P))
        return start

    def compile_extra_components(self, **kwargs):
        self._k = int(self._hyperparameters.k)
        self._frequency = int(self._hyperparameters.frequency)
        self._lr = self._hyperparameters.lr
        self._scale = self._hyperparameters.scale
        start = kwargs["starting_model"]
        self._batch_size = int(self._hyperparameters.batch_size)
        self._prepare()
        self._setup_engine(k_dev=self._k, frequency=self._frequency, theta0=self._flatten_start(start))

    def step(self, save_document_path=None):
        _, loss = self._engine.sg_step(self._lr, self._next_batch())
        self._write_loss(save_document_path, loss)
        self._n += 1
        return loss

    def result(self):
        st = self._engine.sg_state()
        mean, sq, dev = st["mean"], st["sq_mean"], st["dev"]           # dev [S, cols, P]
        f = np.float32(sqrt(self._scale / (self._k - 1)))
        return self._layer_posteriors(lambda lo, hi, s: MultivariateNormalDiagPlusLowRank(
            mean[s, lo:hi], sq[s, lo:hi] - mean[s, lo:hi] ** 2, f * dev[s, :, lo:hi].T))
