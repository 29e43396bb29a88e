"""SGLD — stochastic gradient Langevin dynamics, ``n_chains`` chains per minibatch step on the device.

Drop-in for Pyesian/optimizers/SGLD.py:14-166.  Same surface: hyper-parameters ``batch_size,
lr_upper, lr_lower, lr_gamma`` (:134-137), ``train(n)`` fixes the polynomial schedule ``lr(step) = a (b + step)^-gamma``
from ``n`` (:115-127), ``step()`` returns the running mean of the minibatch losses (:58,95), ``result()`` is a
``BayesianModel`` with one ``Normal(mean, sq_mean - mean**2)`` per weight layer (:146-165 — the variance is handed over as
the scale, kept).  Reference arithmetic kept: the noise is drawn with ``stddev = lr`` and multiplied by ``lr`` again
(:67-68), no prior term, moments updated on every step from step 0.  Weights start from the Keras defaults
(glorot-uniform kernels, zero biases: ``model_from_json``, :138) unless ``compile(..., theta0=...)`` gives them.

One ``pyb_sg_step`` call per step: minibatch gather by index, the shared minibatch gradient kernels for all chains,
and one fused pass for the Langevin update and both running moments.  ``n_chains`` (optional, default 1 = the
reference) runs independent chains on the same minibatches; their posteriors are pooled as a mixture."""
import numpy as np

from .. import _lib
from ..distributions import Normal
from ._sgchains import StochasticGradientChains


class SGLD(StochasticGradientChains):
    KIND = _lib.SG_SGLD

    def __init__(self):
        super().__init__()
        self._running_loss = None
        self._lr = None

    def compile_extra_components(self, **kwargs):
        self._batch_size = int(self._hyperparameters.batch_size)
        self._lr_upper = self._hyperparameters.lr_upper
        self._lr_lower = self._hyperparameters.lr_lower
        self._lr_gamma = self._hyperparameters.lr_gamma
        self._prepare()
        self._setup_engine(theta0=kwargs.get("theta0"))
        self._running_loss = 0

    def _init_sgld_lr(self):
        n = self._nb_iterations
        l_g = np.power(self._lr_lower, 1.0 / self._lr_gamma)
        u_g = np.power(self._lr_upper, 1.0 / self._lr_gamma)
        b = -(n * l_g) / (l_g - u_g)
        a = self._lr_upper * np.power(b, self._lr_gamma)
        self._lr = lambda step: a * np.power((b + step), -self._lr_gamma)

    def step(self, save_document_path=None, noise=None):
        if self._lr is None:
            raise TypeError("'NoneType' object is not callable")      # the reference's lr exists only after train()
        _, loss = self._engine.sg_step(self._lr(self._n), self._next_batch(), noise=noise)
        self._running_loss += loss
        self._write_loss(save_document_path, loss)
        self._n += 1
        return self._running_loss / self._n

    def train(self, nb_iterations: int, loss_save_document_path: str = None, model_save_frequency: int = None,
              model_save_path: str = None, weights_and_biases_log=False):
        self._nb_iterations = nb_iterations
        self._init_sgld_lr()
        super().train(nb_iterations, loss_save_document_path, model_save_frequency, model_save_path,
                      weights_and_biases_log)

    def result(self):
        st = self._engine.sg_state()
        mean, sq = st["mean"], st["sq_mean"]
        return self._layer_posteriors(lambda lo, hi, s: Normal(mean[s, lo:hi], sq[s, lo:hi] - mean[s, lo:hi] ** 2))
