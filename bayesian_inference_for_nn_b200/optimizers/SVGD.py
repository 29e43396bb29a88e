"""SVGD — Stein variational gradient descent, all particles in one device batch.

Drop-in for Pyesian/optimizers/SVGD.py:45-251.  Same surface: hyper-parameters ``batch_size, M, lr``
(:220-225), ``compile(..., prior=...)`` (:223), generic ``train`` loop (Optimizer.py:94-137),
``step()`` returning the mean minibatch loss over particles (:125,141), train/valid loss histories
appended every 10 steps (:137-139), ``result()`` unpacking as ``(models, train_losses, valid_losses)``
(:244-249) — the returned object is also usable as the ``BayesianModel`` north_star asks for
(``result().predict(...)``, ``result().bayesian_model``).

``semantics="reference"`` (default) is the live code path: sequential sweep, gamma = 1, one
particle's gradient broadcast, legacy-Adam descent (:100-123).  ``semantics="canonical"`` is the
median-heuristic Stein update of ``baseline__kernel`` (:165-181) applied Jacobi-style with the prior
gradient included.  Minibatches: one shuffled pass per epoch without replacement, last
batch partial (Optimizer.py:35-41, SVGD.py:91-95).

``devices=[0, 1, ...]`` / ``n_devices=k`` (optional hyper-parameters): the M particles are sharded over several GPUs of
the box inside this process (multi.py; M must be a multiple of the device count); every device joins one NCCL
communicator and the exchange step of the update runs inside the library (csrc/svgd.cu).  ``seed`` absent: a fresh seed
per optimizer (``self.seed``), as the reference is unseeded.
"""
import numpy as np

from .. import _lib
from ..distributions import Sampled
from ..keras_json import parse_model_json
from ..multi import EngineGroup, devices_from
from ..nn import BayesianModel, ParticleModel
from .Optimizer import Optimizer


class SVGDResult(tuple):
    """``(models, train_losses, valid_losses)`` that also answers as a BayesianModel."""

    def __new__(cls, models, train_losses, valid_losses, bayesian_model):
        obj = super().__new__(cls, (models, train_losses, valid_losses))
        obj.bayesian_model = bayesian_model
        return obj

    def predict(self, x, nb_samples=None, **kw):
        n = nb_samples if nb_samples is not None else len(self[0])
        return self.bayesian_model.predict(x, n, **kw)

    def draw(self, nb_samples, mode="reference"):
        return self.bayesian_model.draw(nb_samples, mode)

    def classification_uncertainty(self, *a, **kw):
        return self.bayesian_model.classification_uncertainty(*a, **kw)

    @property
    def last_variance(self):
        return self.bayesian_model.last_variance

    def store(self, path):
        return self.bayesian_model.store(path)

    def sample_model(self):
        return self.bayesian_model.sample_model()


class SVGD(Optimizer):
    def __init__(self):
        super().__init__()
        self._step = 0
        self._M = None
        self.train_losses = []
        self.valid_losses = []
        self._engine = None
        self._valid_ready = False

    def compile_extra_components(self, **kwargs):
        self._batch_size = int(self._hyperparameters.batch_size)
        self._spec = parse_model_json(self._model_config)
        self._prior = kwargs["prior"]
        self._M = int(self._hyperparameters.M)
        self._lr = self._hyperparameters.lr
        sem = self._hp("semantics", "reference")
        self._semantics = (_lib.SVGD_CANONICAL_MEDIAN if sem in ("canonical", _lib.SVGD_CANONICAL_MEDIAN)
                           else _lib.SVGD_REFERENCE_LIVE)
        seed = self._hp("seed", None)
        self.seed = int(np.random.SeedSequence().entropy & ((1 << 63) - 1)) if seed is None else int(seed)
        self._rng = np.random.default_rng(self.seed)
        self._devices = devices_from(self._hp)
        R = len(self._devices)
        if self._M % R:
            raise ValueError("M = %d particles do not shard evenly over %d devices" % (self._M, R))
        self._group = EngineGroup(self._spec, self._devices, seed=self.seed)
        self._engine = self._group.engines[0]
        x, y = self._dataset.training_arrays()
        self._n_train = x.shape[0]
        lowered = self._prior.lower(self._spec)
        p0 = kwargs.get("particles0")
        p0 = None if p0 is None else np.asarray(p0, np.float64)
        Sl = self._M // R

        def setup(i, e):
            e.set_dataset(x, y, self._dataset.loss_kind, n_train=self._n_train)
            e.set_prior(*lowered)
        self._group.each(setup)
        self._group.svgd_join()
        self._group.each(lambda i, e: e.svgd_init(Sl, self._lr, self._semantics, offset=i * Sl,
                                                  particles0=None if p0 is None else p0[i * Sl:(i + 1) * Sl]))
        self._num_particles = self._spec.n_params
        self._epoch_batches = iter(())

    def _next_batch(self):
        """shuffle(cardinality).batch(batch_size): restart the pass when exhausted (SVGD.py:91-95)."""
        b = next(self._epoch_batches, None)
        if b is None:
            perm = self._rng.permutation(self._n_train).astype(np.int32)
            self._epoch_batches = iter([perm[i:i + self._batch_size] for i in range(0, self._n_train, self._batch_size)])
            b = next(self._epoch_batches)
        return b

    def _validation_loss(self):
        """mean over particles of the loss on the whole validation split (SVGD.py:126-129), evaluated on the device
        against the resident particles: the split is uploaded once, nothing but the scalar comes back."""
        if self._dataset.valid_size == 0:
            return 0.0
        if not self._valid_ready:
            xv, yv = self._dataset.split_arrays("valid")
            self._group.each(lambda i, e: e.svgd_set_validation(xv, yv))
            self._valid_ready = True
        return float(self._group.each(lambda i, e: e.svgd_validation_loss())[0])     # all-reduced: every rank holds it

    def step(self, save_document_path=None):
        self._step += 1
        batch = self._next_batch()
        total_loss = self._group.each(lambda i, e: e.svgd_step(batch))[0]            # all-reduced mean over all particles
        if self._step % 10 == 0:
            self.train_losses.append(total_loss)
            self.valid_losses.append(self._validation_loss())
        return total_loss

    @property
    def particles(self):
        return np.concatenate(self._group.each(lambda i, e: e.svgd_particles()))

    def result(self):
        parts = self.particles.astype(np.float32)
        bm = BayesianModel(self._model_config, device=self._devices[0])
        bm.apply_distribution(Sampled(parts, [1] * parts.shape[0]), 0, self._spec.n_keras_layers - 1)
        models = [ParticleModel(bm, parts[i]) for i in range(parts.shape[0])]
        return SVGDResult(models, self.train_losses, self.valid_losses, bm)
