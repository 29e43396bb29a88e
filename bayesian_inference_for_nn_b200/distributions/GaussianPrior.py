"""GaussianPrior — N(mean, rho) prior over every trainable variable.

Mirrors Pyesian/distributions/GaussianPrior.py: same-type check on (mean, rho) (:23-24); scalar
int/float, per-layer list, or per-variable nested list ("tensor") forms (:115-121); layers without
parameters (Flatten) get a ``None`` entry (:44-46).  The HMC path uses ``rho`` RAW as the standard
deviation (no softplus, :42-43), including the negative values the shipped scripts pass
(log of a negative scale makes the Hamiltonian NaN, reproduced on the device).

Instead of a list of ``tfp.distributions.Normal`` objects this build lowers the prior to the flat
per-element ``(mu[P], sigma[P])`` layout of the particle buffer.
"""
import numpy as np

from .. import _lib


class GaussianPrior:
    def __init__(self, mean, rho):
        if type(mean) != type(rho):
            raise Exception("mean and std dev must have the same type")
        self._mean = mean
        self._std_dev = rho

    # ---- lowering to the C ABI -------------------------------------------------------------
    def _kind(self):
        m = self._mean
        if isinstance(m, (int, float)):
            return "scalar"
        if isinstance(m, list) and m and (all(isinstance(v, int) for v in m) or all(isinstance(v, float) for v in m)):
            return "per_layer"
        if isinstance(m, list) and m and all(isinstance(v, list) for v in m):
            return "per_variable"
        raise Exception("mean and standard deviation should be an int, a float, a list or a tensor")

    def lower(self, spec):
        """-> (mean_array, sigma_array, pyb_prior_form) for ``pyb_set_prior_gaussian``."""
        kind = self._kind()
        if kind == "scalar":
            return (np.float32([self._mean]), np.float32([self._std_dev]), _lib.PRIOR_SCALAR)
        mu = np.zeros(spec.n_params, np.float32)
        sg = np.zeros(spec.n_params, np.float32)
        for (li, vi, off, shape) in spec.variables():
            n = int(np.prod(shape))
            if kind == "per_layer":
                mu[off:off + n] = self._mean[li]
                sg[off:off + n] = self._std_dev[li]
            else:
                m = np.asarray(self._mean[li][vi], dtype=np.float32)
                s = np.asarray(self._std_dev[li][vi], dtype=np.float32)
                if tuple(m.shape) != tuple(shape):
                    raise Exception("the shape of the mean tensor does not correspond to the shape of the model "
                                    "layer. Given shape: %s. Expected shape: %s" % (m.shape, shape))
                if tuple(s.shape) != tuple(shape):
                    raise Exception("the shape of the standard deviation tensor does not correspond to the shape of "
                                    "the model layer. Given shape: %s. Expected shape: %s" % (s.shape, shape))
                mu[off:off + n] = m.reshape(-1)
                sg[off:off + n] = s.reshape(-1)
        return mu, sg, _lib.PRIOR_PER_ELEMENT

    def get_model_priors(self, spec):
        """Reference-shaped view: one entry per ``model.layers`` element — ``None`` for layers without
        parameters, else a list of ``(mean, std)`` arrays per trainable variable."""
        mu, sg, form = self.lower(spec)
        if form == _lib.PRIOR_SCALAR:
            mu = np.full(spec.n_params, mu[0], np.float32)
            sg = np.full(spec.n_params, sg[0], np.float32)
        out = [None] * spec.n_keras_layers
        for (li, vi, off, shape) in spec.variables():
            n = int(np.prod(shape))
            if out[li] is None:
                out[li] = []
            out[li].append((mu[off:off + n].reshape(shape), sg[off:off + n].reshape(shape)))
        return out
