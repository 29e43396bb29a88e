"""A minimal TensorFlow / TensorFlow-Probability / Keras stand-in on top of torch (CPU, float32), used ONLY by the golden
generators in this folder to EXECUTE the reference's own optimizer code (Pyesian/optimizers/HMC.py, SVGD.py, ...) in the
build container, where TensorFlow cannot be installed.

What this pins and what it does not.  The reference's control flow and arithmetic *as written in its own files* run
unmodified: leapfrog schedule, kick counts, energies, accept rule, restore and bookkeeping logic, flatten order, the SVGD
sweep.  The third-party numerics underneath — Keras Dense / activations / losses, tfp.Normal.log_prob, tf.random, legacy
Adam — are supplied here from their published definitions (Dense: act(x W + b); SparseCategoricalCrossentropy on a
softmax output computed from the cached logits, Keras 2.15; MeanSquaredError: mean over the last axis, then the batch;
Normal.log_prob: -((x-mu)/sigma)^2/2 - log sigma - log(2 pi)/2; legacy Adam: TF's ResourceApplyAdam form) and autograd is
torch's.  Tensors are immutable (`a += b` rebinds), Variables are assigned in place, as in TF.
"""
import json
import math
import types

import numpy as np
import torch

torch.set_default_dtype(torch.float32)


def _t(x):
    if isinstance(x, TT):
        return x.t
    if isinstance(x, torch.Tensor):
        return x
    if isinstance(x, (int, float)):
        return torch.tensor(float(x))
    a = np.asarray(x)
    if a.dtype == np.float64:
        return torch.as_tensor(a)                      # float64 stays float64 (SVGD keeps its particles in a float64 array)
    return torch.as_tensor(a, dtype=torch.float32) if a.dtype.kind == "f" else torch.as_tensor(a)


class TT:
    """immutable tensor"""
    __array_ufunc__ = None          # numpy scalars (e.g. the SGLD learning rate) defer to the reflected operators below

    def __init__(self, t):
        self.t = t

    shape = property(lambda self: tuple(self.t.shape))
    dtype = property(lambda self: self.t.dtype)

    def numpy(self):
        return self.t.detach().cpu().numpy()

    def __len__(self):
        return self.t.shape[0]

    def __iter__(self):
        return (TT(r) for r in self.t)

    def __getitem__(self, k):
        return TT(self.t[k])

    def __add__(self, o): return TT(self.t + _t(o))
    def __radd__(self, o): return TT(_t(o) + self.t)
    def __sub__(self, o): return TT(self.t - _t(o))
    def __rsub__(self, o): return TT(_t(o) - self.t)
    def __mul__(self, o): return TT(self.t * _t(o))
    def __rmul__(self, o): return TT(_t(o) * self.t)
    def __truediv__(self, o): return TT(self.t / _t(o))
    def __rtruediv__(self, o): return TT(_t(o) / self.t)
    def __neg__(self): return TT(-self.t)
    def __pow__(self, o): return TT(self.t ** o)
    def __lt__(self, o): return bool((self.t < _t(o)).all())
    def __gt__(self, o): return bool((self.t > _t(o)).all())
    def __float__(self): return float(self.t)


class Variable(TT):
    def __init__(self, init):
        super().__init__(_t(init).detach().clone().requires_grad_(True))

    def assign(self, v):
        self.t = _t(v).detach().clone().requires_grad_(True)
        return self

    def assign_add(self, v):
        return self.assign(self.t.detach() + _t(v).detach())

    def assign_sub(self, v):
        return self.assign(self.t.detach() - _t(v).detach())


class GradientTape:
    def __init__(self, persistent=False):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

    def watch(self, what):
        for v in (what if isinstance(what, (list, tuple)) else [what]):
            if not v.t.requires_grad:
                v.t.requires_grad_(True)          # a watched plain tensor: gradients flow to it from here on

    def gradient(self, target, sources):
        single = not isinstance(sources, (list, tuple))
        srcs = [sources] if single else list(sources)
        if not _t(target).requires_grad:              # nothing watched feeds the target
            return None if single else [None] * len(srcs)
        grads = torch.autograd.grad(_t(target).sum(), [s.t for s in srcs], retain_graph=True, allow_unused=True)
        out = [None if g is None else TT(g) for g in grads]
        return out[0] if single else out


# ---- Keras stand-in -------------------------------------------------------------------------------------------------
_ACTS = {"linear": lambda z: z, "relu": torch.relu, "tanh": torch.tanh, "sigmoid": torch.sigmoid,
         "softmax": lambda z: torch.softmax(z, dim=-1)}


class Dense:
    def __init__(self, fan_in, units, activation, use_bias=True):
        self.units, self.activation_name, self.use_bias = units, activation, use_bias
        lim = math.sqrt(6.0 / (fan_in + units))
        self.kernel = Variable((torch.rand(fan_in, units) * 2 - 1) * lim)          # glorot_uniform, as Keras
        self.bias = Variable(torch.zeros(units)) if use_bias else None

    @property
    def trainable_variables(self):
        return [self.kernel] + ([self.bias] if self.use_bias else [])


class InputLike:
    trainable_variables = []


class Model:
    def __init__(self, in_dim, dense_specs, leading_parameterless=0):
        self.layers = [InputLike() for _ in range(leading_parameterless)]
        fan_in = in_dim
        for units, act, use_bias in dense_specs:
            self.layers.append(Dense(fan_in, units, act, use_bias))
            fan_in = units

    @property
    def trainable_variables(self):
        return [v for l in self.layers for v in l.trainable_variables]

    def to_json(self):
        return getattr(self, "_json", "{}")

    def get_weights(self):
        return [v.numpy().copy() for v in self.trainable_variables]

    def set_weights(self, weights):
        for v, w in zip(self.trainable_variables, weights):
            v.assign(torch.as_tensor(np.asarray(w, dtype=np.float32)))

    def clone(self):
        m = Model.__new__(Model)
        m.layers = []
        for l in self.layers:
            if isinstance(l, Dense):
                m.layers.append(Dense(l.kernel.shape[0], l.units, l.activation_name, l.use_bias))
            else:
                m.layers.append(InputLike())
        return m

    def __call__(self, x, training=False):
        a = _t(x).to(torch.float32)
        a = a.reshape(a.shape[0], -1)
        logits = None
        for l in self.layers:
            if isinstance(l, Dense):
                z = a @ l.kernel.t + (l.bias.t if l.use_bias else 0.0)
                logits = z
                a = _ACTS[l.activation_name](z)
        out = TT(a)
        out._keras_logits = logits if self.layers[-1].activation_name == "softmax" else None
        return out


def model_from_json(text):
    cfg = json.loads(text)["config"]
    layers = cfg["layers"] if isinstance(cfg, dict) else cfg
    specs, in_dim, leading = [], None, 0
    for l in layers:
        c = l["config"]
        shape = c.get("batch_input_shape") or c.get("batch_shape")
        if shape and in_dim is None:
            in_dim = int(np.prod([d for d in shape[1:]]))
        if l["class_name"] == "Dense":
            act = c.get("activation", "linear")
            specs.append((int(c["units"]), act if isinstance(act, str) else act["config"], bool(c.get("use_bias", True))))
        elif l["class_name"] in ("Flatten",) and not specs:
            leading += 1
    m = Model(in_dim, specs, leading)
    m._json = text                       # Keras round-trips the architecture JSON; the stand-in hands back what it was given
    return m


class SparseCategoricalCrossentropy:
    def __init__(self, reduction="auto", from_logits=False):
        self.reduction = reduction

    def __call__(self, y_true, y_pred):
        y = _t(y_true).reshape(-1).to(torch.int64)
        logits = getattr(y_pred, "_keras_logits", None)
        if logits is not None:                      # Keras 2.15 eager: cached logits of the softmax output
            per = torch.logsumexp(logits, dim=-1) - logits.gather(1, y[:, None])[:, 0]
        else:
            p = torch.clamp(_t(y_pred), 1e-7, 1 - 1e-7)
            per = -torch.log(p.gather(1, y[:, None])[:, 0])
        return TT(per if self.reduction == "none" else per.mean())


class MeanSquaredError:
    def __init__(self, reduction="auto"):
        self.reduction = reduction

    def __call__(self, y_true, y_pred):
        p = _t(y_pred)
        per = ((p - _t(y_true).to(torch.float32).reshape(p.shape)) ** 2).mean(dim=-1)
        return TT(per if self.reduction == "none" else per.mean())


# ---- tfp stand-in ---------------------------------------------------------------------------------------------------
class Normal:
    def __init__(self, loc, scale):
        self.loc, self.scale = _t(loc), _t(scale)

    def mean(self):
        return TT(self.loc * torch.ones_like(self.scale))

    def log_prob(self, x):
        z = (_t(x) - self.loc) / self.scale
        return TT(-0.5 * z * z - torch.log(self.scale) - 0.5 * math.log(2.0 * math.pi))

    def sample(self):
        return TT(self.loc + self.scale * torch.as_tensor(RANDOM.next_normal(tuple(self.scale.shape))))


# ---- injected randomness ----------------------------------------------------------------------------------------------
class _Random:
    """tf.random.normal draws come from a queue the generator fills, so that the goldens record them"""

    def __init__(self):
        self.rng = np.random.default_rng(0)
        self.log = []

    def next_normal(self, shape):
        z = self.rng.standard_normal(shape).astype(np.float32)
        self.log.append(z)
        return z


RANDOM = _Random()


def _random_normal(shape, mean=0.0, stddev=1.0, dtype=None):
    z = torch.as_tensor(RANDOM.next_normal(tuple(shape)))
    return TT(z * float(stddev) + float(mean))


# ---- legacy Adam (TF ResourceApplyAdam) ---------------------------------------------------------------------------------
class LegacyAdam:
    def __init__(self, learning_rate=0.001, beta_1=0.9, beta_2=0.999, epsilon=1e-7):
        self.lr, self.b1, self.b2, self.eps, self.t, self.slots = learning_rate, beta_1, beta_2, epsilon, 0, {}

    def apply_gradients(self, grads_and_vars):
        self.t += 1
        lr_t = self.lr * math.sqrt(1 - self.b2 ** self.t) / (1 - self.b1 ** self.t)
        for g, v in grads_and_vars:
            m, s = self.slots.get(id(v), (torch.zeros_like(v.t), torch.zeros_like(v.t)))
            g = _t(g).detach().to(torch.float32)
            m = self.b1 * m + (1 - self.b1) * g
            s = self.b2 * s + (1 - self.b2) * g * g
            self.slots[id(v)] = (m, s)
            v.assign(v.t.detach() - lr_t * m / (torch.sqrt(s) + self.eps))


# ---- the module objects --------------------------------------------------------------------------------------------
class NS(types.SimpleNamespace):
    """namespace whose unknown attributes are inert: the reference's other modules mention tf symbols at import time
    (annotations, base classes) that nothing executed here ever touches"""

    def __getattr__(self, name):
        from unittest.mock import MagicMock
        return MagicMock(name=name)


class _Module(types.ModuleType):
    def __getattr__(self, name):
        from unittest.mock import MagicMock
        return MagicMock(name=name)


def make_tf():
    tf = _Module("tensorflow")
    tf.Tensor, tf.Variable, tf.GradientTape = TT, Variable, GradientTape
    tf.float32, tf.float64 = torch.float32, torch.float64
    tf.zeros = lambda shape, dtype=None: TT(torch.zeros(tuple(shape)))
    tf.ones = lambda shape, dtype=None: TT(torch.ones(tuple(shape)))
    tf.zeros_like = lambda x, dtype=None: TT(torch.zeros_like(_t(x)))
    tf.ones_like = lambda x, dtype=None: TT(torch.ones_like(_t(x)))
    tf.constant = lambda v, dtype=None: TT(_t(v).to(torch.float32))
    tf.convert_to_tensor = lambda v, dtype=None: TT(_t(v))
    tf.cast = lambda v, dtype=None: TT(_t(v).to(torch.float32 if dtype in ("float32", torch.float32, None) else dtype))
    tf.float64 = torch.float64
    tf.identity = lambda v: TT(_t(v).detach().clone())
    tf.multiply = lambda a, b: TT(_t(a) * _t(b))
    tf.square = lambda a: TT(_t(a) ** 2)
    tf.reshape = lambda a, shape: TT(_t(a).reshape(tuple(shape) if not isinstance(shape, int) else (shape,)))
    tf.concat = lambda vals, axis=0: TT(torch.cat([_t(v) for v in vals], dim=axis))
    tf.stack = lambda vals, axis=0: TT(torch.stack([_t(v) for v in vals], dim=axis))
    tf.expand_dims = lambda a, axis: TT(_t(a).unsqueeze(axis))
    tf.reduce_sum = lambda a, axis=None: TT(_t(a).sum() if axis is None else _t(a).sum(dim=axis))
    tf.reduce_prod = lambda a: int(np.prod(tuple(a)))
    tf.exp = lambda a: TT(torch.exp(_t(a)))
    tf.repeat = lambda a, repeats, axis=0: TT(torch.repeat_interleave(_t(a), repeats, dim=axis))
    tf.matmul = lambda a, b: TT(_t(a) @ _t(b))
    tf.math = NS(reduce_sum=tf.reduce_sum, exp=tf.exp, square=tf.square,
                                    is_nan=lambda a: TT(torch.isnan(_t(a))))
    tf.where = lambda c, a, b: TT(torch.where(_t(c).bool(), _t(a), _t(b)))
    tf.size = lambda a: TT(torch.tensor(int(np.prod(tuple(a.shape)))))
    tf.random = NS(normal=_random_normal)
    tf.keras = NS(
        Model=Model, models=NS(Model=Model, model_from_json=model_from_json, clone_model=lambda m: m.clone()),
        losses=NS(SparseCategoricalCrossentropy=SparseCategoricalCrossentropy, MeanSquaredError=MeanSquaredError),
        optimizers=NS(legacy=NS(Adam=LegacyAdam)))
    tf.data = NS(Dataset=object)
    return tf


def make_tfp():
    tfp = _Module("tensorflow_probability")
    tfp.distributions = NS(Normal=Normal, Distribution=object)
    return tfp


class ArrayData:
    """what the reference asks of a tf.data.Dataset: cardinality().numpy().item(), shuffle, batch, iteration"""

    def __init__(self, x, y, bs=None):
        self.x, self.y, self.bs = x, y, bs

    def cardinality(self):
        n = len(self.x) if not self.bs else -(-len(self.x) // self.bs)
        return types.SimpleNamespace(numpy=lambda: np.int64(n))

    def shuffle(self, n):
        return self            # order kept: the goldens need reproducible minibatches

    def batch(self, bs):
        return ArrayData(self.x, self.y, int(bs.numpy()) if hasattr(bs, "numpy") else int(bs))

    def map(self, fn):
        return [fn(a, b) for a, b in zip(self.x, self.y)]

    def __iter__(self):
        for i in range(0, len(self.x), self.bs):
            yield TT(torch.as_tensor(self.x[i:i + self.bs])), TT(torch.as_tensor(self.y[i:i + self.bs]))
