"""HyperParameters — attribute bag handed to ``Optimizer.compile``.

Mirrors Pyesian/optimizers/hyperparameters/HyperParameters.py:6-62: kwargs deep-copied, default
``batch_size=64`` (:17-18), ``AttributeError`` for a missing name (:21-24), and the GUI's text
format ``key value key value`` (:32-62).  HMC reads ``m, L, epsilon`` (HMC.py:53-55), SVGD reads
``batch_size, M, lr`` (SVGD.py:220-225).  Extra, optional knobs understood by this build (absent
=> reference behaviour): ``n_chains``, ``seed``, ``semantics``, ``device``, ``path``.
"""
import copy


class HyperParameters:
    def __init__(self, **kwargs):
        self._params = copy.deepcopy(kwargs)
        if "batch_size" not in kwargs:
            self._params["batch_size"] = 64
        self.connectors = "._-"

    def __getattr__(self, item):
        params = self.__dict__.get("_params", {})
        if item in params:
            return params[item]
        raise AttributeError("'HyperParameters' object has no attribute " + str(item))

    def get(self, item, default=None):
        return self._params.get(item, default)

    def from_file(self, fn):
        with open(fn, "r") as f:
            return self.parse(f.read())

    def parse(self, text: str):
        """Two-state scanner: a key is a run of [alnum . _ -], a value a run of [digit - .]."""
        keys, values = [], []
        key, val, reading_value = "", "", False
        for ch in text:
            if not reading_value:
                if ch.isalnum() or ch in self.connectors:
                    key += ch
                elif key:
                    keys.append(key)
                    key, reading_value = "", True
            else:
                if ch.isdigit() or ch in "-.":
                    val += ch
                elif val:
                    values.append(float(val))
                    val, reading_value = "", False
        if key:
            keys.append(key)
            values.extend([0.0] * (len(keys) - len(values)))
        elif val:
            values.append(float(val))
        for k, v in zip(keys, values):
            self._params[k] = v
        return self
