"""Shared host side of the S-batched stochastic-gradient optimizers (SGLD, SWAG): the minibatch order of
``Optimizer._dataset_setup`` (Optimizer.py:35-41: one shuffled pass per epoch, last batch partial, iterator restarted when
exhausted — SGLD.py:48-52, SWAG.py:46-50), the engine, and the per-layer posterior assembly of ``result()``."""
import numpy as np

from ..distributions import Mixture
from ..distributions.Mixture import ChainSelector
from ..engine import Engine
from ..keras_json import parse_model_json
from ..nn import BayesianModel
from .Optimizer import Optimizer


class StochasticGradientChains(Optimizer):
    KIND = None

    def __init__(self):
        super().__init__()
        self._engine = None
        self._n = None

    def _prepare(self):
        self._spec = parse_model_json(self._model_config)
        self._n_chains = int(self._hp("n_chains", 1))

    def _setup_engine(self, k_dev=0, frequency=1, theta0=None):
        # unseeded in the reference (tf.random / np.random without a library-level seed): a fresh seed per optimizer
        # unless one is given; self.seed reproduces the run
        seed = self._hp("seed", None)
        self.seed = int(np.random.SeedSequence().entropy & ((1 << 63) - 1)) if seed is None else int(seed)
        self._rng = np.random.default_rng(self.seed)
        self._engine = Engine(self._spec, device=int(self._hp("device", 0)), seed=self.seed)
        x, y = self._dataset.training_arrays()
        self._n_train = x.shape[0]
        self._engine.set_dataset(x, y, self._dataset.loss_kind, n_train=self._n_train)
        self._engine.sg_init(self._n_chains, self.KIND, k_dev=k_dev, frequency=frequency, theta0=theta0,
                             chain_offset=int(self._hp("chain_offset", 0)))
        self._epoch_batches = iter(())
        self._n = 0

    def _next_batch(self):
        b = next(self._epoch_batches, None)
        if b is None:
            perm = self._rng.permutation(self._n_train).astype(np.int32)
            self._epoch_batches = iter([perm[i:i + self._batch_size] for i in range(0, self._n_train, self._batch_size)])
            b = next(self._epoch_batches)
        return b

    def _write_loss(self, save_document_path, loss):
        if save_document_path is not None:
            with open(save_document_path, "a") as f:
                f.write(str(loss))

    def _layer_posteriors(self, make):
        """BayesianModel with one distribution per weight-carrying layer, applied on [idx, idx] like the reference
        (SGLD.py:147-165, SWAG.py:119-139).  ``make(lo, hi, chain)`` builds the distribution of one chain over the
        flat range of one layer; several chains become an equal-weight Mixture whose per-layer instances share ONE
        chain selector, so that a drawn network takes every layer from the same chain (chains are not exchangeable
        layer by layer)."""
        model = BayesianModel(self._model_config, device=int(self._hp("device", 0)))
        selector = ChainSelector(self._n_chains) if self._n_chains > 1 else None
        for d in self._spec.dense:
            lo, hi = self._spec.layer_param_range(d.keras_index, d.keras_index)
            comps = [make(lo, hi, s) for s in range(self._n_chains)]
            model.apply_distribution(comps[0] if len(comps) == 1 else Mixture(comps, selector=selector), d.keras_index,
                                     d.keras_index)
        return model

    @property
    def chains(self):
        """device state of every chain: theta, mean, sq_mean [S, P] (+ dev [S, cols, P] for SWAG), n"""
        return self._engine.sg_state()
