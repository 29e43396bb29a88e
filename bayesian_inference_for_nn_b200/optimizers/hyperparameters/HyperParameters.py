"""HyperParameters — attribute bag handed to ``Optimizer.compile``.

Mirrors Pyesian/optimizers/hyperparameters/HyperParameters.py:6-62: kwargs deep-copied, default
``batch_size=64`` (:17-18), ``AttributeError`` for a missing name (:21-24), and the GUI's text
format ``key value key value`` (:32-62).  HMC reads ``m, L, epsilon`` (HMC.py:53-55), SVGD reads
``batch_size, M, lr`` (SVGD.py:220-225).  Extra, optional knobs understood by this build (absent
=> reference behaviour): ``n_chains``, ``seed``, ``semantics``, ``device``, ``path``.
"""
import copy
import re


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

    _NUMBER = re.compile(r"[0-9.\-]+")

    def parse(self, text: str):
        """``name number name number ...`` as the GUI writes it (``static/hyperparams/*.txt``).  Same reading as the
        reference's character scanner (:32-62): a name is a run of alphanumerics and ``. _ -``; the character that ends
        it is dropped; whatever precedes the next run of digits / ``-`` / ``.`` is skipped; the character that ends a
        number is dropped too; a name at the very end of the text gets 0.0 (and so does every name still without a
        number then)."""
        names, numbers, pos, end = [], [], 0, len(text)
        while pos < end:
            m = self._name_at_or_after(text, pos)
            if m is None:
                break
            names.append(m.group())
            if m.end() == end:                                    # unterminated last name: pad with zeros
                numbers.extend([0.0] * (len(names) - len(numbers)))
                break
            v = self._NUMBER.search(text, m.end() + 1)
            if v is None:
                break
            numbers.append(float(v.group()))
            pos = v.end() + 1
        for i, name in enumerate(names):
            self._params[name] = numbers[i]        # a name left without a number is an IndexError, as in the reference
        return self

    def _name_at_or_after(self, text, pos):
        """first maximal run of name characters starting at or after pos (str.isalnum decides what a letter is)"""
        n = len(text)
        while pos < n and not (text[pos].isalnum() or text[pos] in self.connectors):
            pos += 1
        if pos == n:
            return None
        stop = pos
        while stop < n and (text[stop].isalnum() or text[stop] in self.connectors):
            stop += 1
        return _Span(text, pos, stop)


class _Span:
    def __init__(self, text, start, stop):
        self._text, self._start, self._stop = text, start, stop

    def group(self):
        return self._text[self._start:self._stop]

    def end(self):
        return self._stop
