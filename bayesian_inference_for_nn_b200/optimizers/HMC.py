"""HMC — Hamiltonian Monte Carlo, S chains at once on one B200.

Drop-in for Pyesian/optimizers/HMC.py:13-187.  Same surface: hyper-parameters ``m, L, epsilon``
(:53-55), ``compile(..., prior=GaussianPrior)`` (KeyError without ``prior``, :60), ``step(
save_document_path, sampling, burning)`` (:74), ``train(n)`` = 10 always-accept burn-in iterations +
n sampling iterations with the ``loss`` / ``accept_rate`` progress line (:106-126), ``result()`` ->
``BayesianModel`` holding one ``Sampled`` over all layers (:176-187).  Quirks kept: the
``nb_burn_epoch``/``nb_burn_epochs`` key mismatch (:61-62), save/loss-file arguments of ``train``
ignored (:106-126), chains start at the prior mean (:69-72).

Underneath, every iteration is a handful of kernel launches on the handle's stream with no host
sync (pyb_hmc_run); the reference's single chain becomes ``n_chains`` (optional hyper-parameter,
default 1) independent chains whose accepted states are pooled into the returned ``Sampled``.
Optional knobs (all absent => reference behaviour): ``n_chains``, ``seed`` (absent: a fresh 64-bit seed per
optimizer, kept in ``self.seed``), ``semantics`` ("reference" | "canonical"), ``device``, ``path`` ("auto" | "generic" |
"fused" | "tensor"), ``chain_offset`` (global id of the first local chain when chains are sharded over processes), and
``devices=[0, 1, ...]`` / ``n_devices=k``: the chains are sharded over several GPUs of the box inside this process
(multi.py) and ``result()`` pools them in global chain order — bit-identical to the one-device run.
"""
import numpy as np

from .. import _lib
from ..distributions import Sampled
from ..keras_json import parse_model_json
from ..multi import EngineGroup, devices_from
from ..nn import BayesianModel
from .Optimizer import Optimizer

_PATHS = {"auto": _lib.PATH_AUTO, "generic": _lib.PATH_GENERIC, "fused": _lib.PATH_FUSED_SMALL,
          "tensor": _lib.PATH_TENSOR}


class HMC(Optimizer):
    def __init__(self):
        super().__init__()
        self._nb_burn_epoch = 10
        self._engine = None
        self._spec = None
        self._epsilon = self._L = self._m = None
        self._total_runs = 0
        self._accepted_runs = 0
        self._current_loss = 0
        self.last_diag = None

    def compile_extra_components(self, **kwargs):
        self._m = self._hyperparameters.m
        self._L = self._hyperparameters.L
        self._epsilon = self._hyperparameters.epsilon
        self._spec = parse_model_json(self._model_config)
        prior = kwargs["prior"]
        if "nb_burn_epoch" in kwargs:
            self._nb_burn_epoch = kwargs["nb_burn_epochs"]
        self._n_chains = int(self._hp("n_chains", 1))
        sem = self._hp("semantics", "reference")
        self._semantics = _lib.HMC_CANONICAL if sem in ("canonical", _lib.HMC_CANONICAL) else _lib.HMC_REFERENCE
        # the reference is stochastic by default (unseeded tf.random.normal / random.random, HMC.py:88,171): without an
        # explicit seed every optimizer draws its own, so repeated runs are independent; self.seed reproduces a run
        seed = self._hp("seed", None)
        self.seed = int(np.random.SeedSequence().entropy & ((1 << 63) - 1)) if seed is None else int(seed)
        self._devices = devices_from(self._hp)
        self._engine = EngineGroup(self._spec, self._devices, seed=self.seed)
        path = self._hp("path", "auto")
        x, y = self._dataset.training_arrays()          # ONE full-dataset batch, frozen for the run (HMC.py:63-65)
        lowered = prior.lower(self._spec)

        def setup(i, e):
            e.set_option("path", _PATHS.get(path, path) if isinstance(path, str) else path)
            e.set_dataset(x, y, self._dataset.loss_kind, n_train=self._dataset.train_size)
            e.set_prior(*lowered)
        self._engine.each(setup)
        self._engine.hmc_init(self._n_chains, self._epsilon, self._m, int(self._L), self._semantics,
                              q0=kwargs.get("q0"), chain_offset=int(self._hp("chain_offset", 0)))

    # ---- one iteration (HMC.py:74-104) -----------------------------------------------------
    def step(self, save_document_path=None, sampling=True, burning=False):
        d = self._engine.hmc_run(1, burning=burning, sampling=sampling)
        self._book(d)
        return d["mean_loss"]

    def _book(self, d):
        self.last_diag = d
        self._total_runs += d["n_total"]
        self._accepted_runs += d["n_accepted"]
        self._current_loss = d["mean_loss"]

    def _phase(self, n, suffix, sampling, burning):
        """n iterations; with verbose on, the progress line is refreshed ~20 times per phase (each
        refresh is the only host sync), otherwise the whole phase is one C call."""
        chunk = max(1, n // 20) if self._verbose else max(1, n)
        done = 0
        while done < n:
            k = min(chunk, n - done)
            d = self._engine.hmc_run(k, burning=burning, sampling=sampling)
            self._book(d)
            done += k
            rate = self._accepted_runs / max(1, self._total_runs)
            self._print_progress(done / n, suffix=suffix, loss=d["mean_loss"], accept_rate=rate, bar_length=20)
        self._new_progress_line()

    def train(self, nb_iterations: int, loss_save_document_path: str = None, model_save_frequency: int = None,
              model_save_path: str = None):
        self._accepted_runs = self._total_runs = 0
        self._phase(int(self._nb_burn_epoch), "HMC - Burning", sampling=False, burning=True)
        self._accepted_runs = self._total_runs = 0
        self._engine.each(lambda i, e: e.hmc_reset_samples())
        self._phase(int(nb_iterations), "HMC - Sampling", sampling=True, burning=False)

    @property
    def accept_rate(self):
        return self._accepted_runs / max(1, self._total_runs)

    def result(self) -> BayesianModel:
        samples, freq, _chain = self._engine.hmc_samples()
        if samples.shape[0] == 0:
            q, _ = self._engine.hmc_state()          # no sampling iteration yet: current positions, weight 1
            samples, freq = q, np.ones(q.shape[0], np.int32)
        posterior = BayesianModel(self._model_config, device=self._devices[0])
        posterior.apply_distribution(Sampled(samples, freq.tolist()), 0, self._spec.n_keras_layers - 1)
        return posterior
