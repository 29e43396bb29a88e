"""Weight-layout packer: Keras model JSON (``model.to_json()``) -> Dense-stack description and the
flat ``[P]`` parameter layout of the particle buffer.

Reference behaviour restated: ``tf.keras.models.model_from_json`` at HMC.py:56, SVGD.py:222,
BayesianModel.py:18; flat order = ``model.layers`` order -> ``trainable_variables`` order (kernel
``[in,out]`` then bias) -> C-order flatten (HMC.py:178-183, SVGD.py:159-160,230-239,
BayesianModel.py:73-77).  ``model.layers`` of a Sequential excludes the ``InputLayer``; a ``Flatten``
is a layer without parameters (its prior entry is ``None``, GaussianPrior.py:44-46).
"""
from __future__ import annotations

import json
from dataclasses import dataclass, field
from typing import List, Optional, Tuple

from . import _lib

_ACT = {"linear": _lib.ACT_LINEAR, None: _lib.ACT_LINEAR, "relu": _lib.ACT_RELU, "softmax": _lib.ACT_SOFTMAX,
        "tanh": _lib.ACT_TANH, "sigmoid": _lib.ACT_SIGMOID}


class UnsupportedModelError(ValueError):
    pass


@dataclass
class DenseLayer:
    units: int
    activation: int
    use_bias: bool
    fan_in: int
    w_off: int
    b_off: int            # -1 when use_bias is False
    keras_index: int      # index in model.layers
    name: str = ""


@dataclass
class ModelSpec:
    in_dim: int
    dense: List[DenseLayer]
    n_keras_layers: int                 # len(model.layers) in the reference
    keras_layer_kinds: List[str]        # class names of model.layers
    input_shape: Tuple[int, ...] = ()
    json_text: str = ""

    @property
    def n_params(self) -> int:
        if not self.dense:
            return 0
        last = self.dense[-1]
        return (last.b_off + last.units) if last.use_bias else (last.w_off + last.fan_in * last.units)

    @property
    def out_dim(self) -> int:
        return self.dense[-1].units

    def variables(self):
        """[(keras_layer_index, var_index, offset, shape)] in flat order."""
        out = []
        for d in self.dense:
            out.append((d.keras_index, 0, d.w_off, (d.fan_in, d.units)))
            if d.use_bias:
                out.append((d.keras_index, 1, d.b_off, (d.units,)))
        return out

    def layer_param_range(self, start_layer: int, end_layer: int):
        """flat [lo, hi) covered by model.layers[start_layer..end_layer] (both included)."""
        lo, hi = None, None
        for d in self.dense:
            if start_layer <= d.keras_index <= end_layer:
                a = d.w_off
                b = (d.b_off + d.units) if d.use_bias else (d.w_off + d.fan_in * d.units)
                lo = a if lo is None else min(lo, a)
                hi = b if hi is None else max(hi, b)
        return (0, 0) if lo is None else (lo, hi)


def _activation_name(cfg):
    a = cfg.get("activation", "linear")
    if isinstance(a, dict):          # Keras 3 serialises custom/activation objects as dicts
        a = a.get("config", a.get("class_name", "linear"))
        if isinstance(a, dict):
            a = a.get("name", "linear")
    return a


def _shape_from(cfg, layer):
    for key in ("batch_input_shape", "batch_shape"):
        if cfg.get(key):
            return [d for d in cfg[key][1:]]
    bc = layer.get("build_config") or {}
    if bc.get("input_shape"):
        return [d for d in bc["input_shape"][1:]]
    return None


def parse_model_json(model_config: str, in_dim: Optional[int] = None) -> ModelSpec:
    """Parse a Keras Sequential (or a linear Functional chain) of InputLayer/Flatten/Dense layers."""
    try:
        top = json.loads(model_config)
    except (TypeError, json.JSONDecodeError) as e:
        raise UnsupportedModelError("model_config is not valid JSON: %s" % e)
    if top.get("class_name") not in ("Sequential", "Functional", "Model"):
        raise UnsupportedModelError("only Sequential/Functional Dense stacks are supported, got %r"
                                    % top.get("class_name"))
    layers = top.get("config", {}).get("layers", [])
    cur_shape = None
    dense: List[DenseLayer] = []
    kinds: List[str] = []
    input_shape: Tuple[int, ...] = ()
    off = 0
    for layer in layers:
        cls = layer.get("class_name")
        cfg = layer.get("config", {})
        shp = _shape_from(cfg, layer)
        if cur_shape is None and shp is not None:
            cur_shape = list(shp)
            input_shape = tuple(int(d) for d in shp)
        if cls == "InputLayer":
            continue  # not part of model.layers for a Sequential
        if cur_shape is None:
            if in_dim is None:
                raise UnsupportedModelError(
                    "the model JSON carries no input shape (no batch_input_shape / InputLayer / build_config); "
                    "pass in_dim")
            cur_shape = [int(in_dim)]
            input_shape = (int(in_dim),)
        if cls == "Flatten":
            n = 1
            for d in cur_shape:
                n *= int(d)
            cur_shape = [n]
            kinds.append(cls)
        elif cls == "Dense":
            if len(cur_shape) != 1:
                raise UnsupportedModelError("Dense on a rank-%d input needs a Flatten first" % (len(cur_shape) + 1))
            act = _activation_name(cfg)
            if act not in _ACT:
                raise UnsupportedModelError("unsupported activation %r" % (act,))
            units = int(cfg["units"])
            use_bias = bool(cfg.get("use_bias", True))
            fan_in = int(cur_shape[0])
            w_off = off
            off += fan_in * units
            b_off = -1
            if use_bias:
                b_off = off
                off += units
            dense.append(DenseLayer(units, _ACT[act], use_bias, fan_in, w_off, b_off, len(kinds), cfg.get("name", "")))
            cur_shape = [units]
            kinds.append(cls)
        else:
            raise UnsupportedModelError("layer type %r is outside the Dense hot path" % cls)
    if not dense:
        raise UnsupportedModelError("the model has no Dense layer")
    for d in dense[:-1]:
        if d.activation == _lib.ACT_SOFTMAX:
            raise UnsupportedModelError("softmax is only supported on the output layer")
    first_in = dense[0].fan_in
    return ModelSpec(first_in, dense, len(kinds), kinds, input_shape, model_config)


def make_sequential_json(in_dim: int, units, activations, use_bias=None) -> str:
    """Helper for environments without Keras: emit a Keras-2.15-style Sequential JSON."""
    inv = {v: k for k, v in _ACT.items() if k}
    use_bias = use_bias or [True] * len(units)
    layers = [{"module": "keras.layers", "class_name": "InputLayer",
               "config": {"batch_input_shape": [None, int(in_dim)], "dtype": "float32", "sparse": False,
                          "ragged": False, "name": "dense_input"}, "registered_name": None}]
    fin = in_dim
    for i, (u, a, b) in enumerate(zip(units, activations, use_bias)):
        name = "dense" if i == 0 else "dense_%d" % i
        a = a if isinstance(a, str) else inv[a]
        layers.append({"module": "keras.layers", "class_name": "Dense",
                       "config": {"name": name, "trainable": True, "dtype": "float32", "units": int(u),
                                  "activation": a, "use_bias": bool(b)},
                       "registered_name": None, "build_config": {"input_shape": [None, int(fin)]}})
        fin = u
    return json.dumps({"class_name": "Sequential", "config": {"name": "sequential", "layers": layers},
                       "keras_version": "2.15.0", "backend": "tensorflow"})
