"""Dataset — boundary mirror of Pyesian/datasets/Dataset.py (only what the hot path reads).

Kept: the shuffle + 80/10/10 ``take/skip`` split (:113-122), ``train_size/test_size/valid_size/size``,
``training_dataset()`` (:143-150), the ``loss(reduction)`` factory (:152-159), ``input_shape()``,
``train_data/valid_data/test_data`` objects that still answer ``.batch(n)``, ``.cardinality()`` and
iteration the way the scripts use them (e.g. ``next(iter(dataset.test_data.batch(n)))``,
HMC_classification.py:64-65).  Out of scope (host I/O): tfds names, image folders,
UCI ids — pass arrays, a DataFrame, the path of a CSV file (last ``target_dim`` columns are the
labels, :124-133), or a ``tf.data.Dataset`` (materialised once, lazily importing TensorFlow only in
that case).

Differences that matter for the device path: the split is materialised ONCE into contiguous NumPy
arrays (the reference re-shuffles on every iteration, SURVEY B-8), so the full training batch can be
uploaded to HBM in one copy (HMC.py:65 does ``next(iter(train.batch(N)))``).
"""
import numpy as np

from .. import _lib
from ..tensors import to_numpy


class ArrayDataset:
    """A tiny stand-in for the ``tf.data.Dataset`` objects the reference exposes."""

    def __init__(self, x, y, batch_size=None):
        self.x, self.y, self._bs = x, y, batch_size

    def cardinality(self):
        n = self.x.shape[0]
        if self._bs:
            n = (n + self._bs - 1) // self._bs
        return np.int64(n)

    def __len__(self):
        return int(self.cardinality())

    def batch(self, n):
        return ArrayDataset(self.x, self.y, int(n))

    def take(self, n):
        return ArrayDataset(self.x[:n], self.y[:n], self._bs)

    def skip(self, n):
        return ArrayDataset(self.x[n:], self.y[n:], self._bs)

    def shuffle(self, buffer_size=None, seed=None):
        perm = np.random.default_rng(seed).permutation(self.x.shape[0])
        return ArrayDataset(self.x[perm], self.y[perm], self._bs)

    def map(self, fn, **_):
        xs, ys = zip(*[fn(a, b) for a, b in zip(self.x, self.y)]) if self.x.shape[0] else ((), ())
        return ArrayDataset(np.asarray(xs), np.asarray(ys), self._bs)

    def cache(self):
        return self

    def prefetch(self, *_):
        return self

    def __iter__(self):
        if self._bs:
            for i in range(0, self.x.shape[0], self._bs):
                yield self.x[i:i + self._bs], self.y[i:i + self._bs]
        else:
            for a, b in zip(self.x, self.y):
                yield a, b


def _loss_kind_of(loss):
    name = loss if isinstance(loss, str) else getattr(loss, "__name__", type(loss).__name__)
    low = name.lower().replace("_", "")
    if "sparsecategoricalcrossentropy" in low:
        return _lib.LOSS_SPARSE_CE
    if "meansquarederror" in low or low == "mse":
        return _lib.LOSS_MSE
    raise ValueError("unsupported loss %r: the hot path covers SparseCategoricalCrossentropy and MeanSquaredError"
                     % (name,))


class _NumpyLoss:
    """Callable returned by ``Dataset.loss()`` when the loss was given by name (no Keras around).
    Host-side convenience for scripts/metrics; the training path never calls it."""

    def __init__(self, kind, reduction="auto"):
        self.kind, self.reduction = kind, reduction

    def __call__(self, y_true, y_pred):
        y_true, y_pred = to_numpy(y_true), to_numpy(y_pred, np.float32)
        if self.kind == _lib.LOSS_SPARSE_CE:
            p = np.clip(y_pred, 1e-7, 1 - 1e-7)
            per = -np.log(p[np.arange(p.shape[0]), y_true.reshape(-1).astype(np.int64)])
        else:
            per = ((y_pred - y_true.reshape(y_pred.shape)) ** 2).mean(axis=-1)
        return per if self.reduction == "none" else (per.sum() if self.reduction == "sum" else per.mean())


class Dataset:
    def __init__(self, dataset, loss, likelihoodModel="Classification", load_images=False, target_dim=1,
                 feature_normalisation=False, label_normalisation=False, train_proportion=0.8,
                 test_proportion=0.1, valid_proportion=0.1, seed=None):
        if train_proportion + test_proportion + valid_proportion != 1:
            raise ValueError("Dataset split test_proportions must sum up to 1")
        self._train_proportion, self._test_proportion, self._valid_proportion = (
            train_proportion, test_proportion, valid_proportion)
        self._loss = loss
        self.loss_kind = _loss_kind_of(loss)
        self.likelihood_model = likelihoodModel
        self.target_dim = target_dim
        self._label_mean = self._label_std = None
        self._load_images = load_images
        x, y = self._materialise(dataset)
        self._init_from_arrays(x, y, seed)
        if feature_normalisation:
            self.feature_normalisation()
        if label_normalisation:
            self.label_normalisation()

    # ---- ingestion -------------------------------------------------------------------------
    def _materialise(self, dataset):
        if isinstance(dataset, (tuple, list)) and len(dataset) == 2:
            return to_numpy(dataset[0]), to_numpy(dataset[1])
        if isinstance(dataset, ArrayDataset):
            return dataset.x, dataset.y
        if isinstance(dataset, str) and not self._load_images and dataset.lower().endswith(".csv"):
            import pandas as pd                      # _init_from_csv (:131-133): read_csv, then the DataFrame rule
            dataset = pd.read_csv(dataset)
        mod = type(dataset).__module__ or ""
        if mod.startswith("pandas"):
            return (dataset.iloc[:, :-self.target_dim].values, dataset.iloc[:, -self.target_dim:].values)
        if mod.startswith("tensorflow"):
            xs, ys = zip(*[(to_numpy(a), to_numpy(b)) for a, b in dataset])   # one pass, host side
            return np.stack(xs), np.stack(ys)
        raise ValueError("Unsupported dataset format")

    def _init_from_arrays(self, x, y, seed):
        x, y = np.asarray(x), np.asarray(y)
        if x.shape[0] != y.shape[0]:
            raise ValueError("features and labels disagree on the number of rows")
        perm = np.random.default_rng(seed).permutation(x.shape[0])   # dataset.shuffle(cardinality) (:114)
        x, y = x[perm], y[perm]
        self.size = int(x.shape[0])
        self.train_size = int(self._train_proportion * self.size)
        self.test_size = int(self._test_proportion * self.size)
        self.valid_size = int(self._valid_proportion * self.size)
        a, b = self.train_size, self.train_size + self.test_size
        self.train_data = ArrayDataset(x[:a], y[:a])
        self.test_data = ArrayDataset(x[a:b], y[a:b])
        self.valid_data = ArrayDataset(x[b:], y[b:])     # skip(test_size) keeps the remainder (:120)

    # ---- reference surface -----------------------------------------------------------------
    def training_dataset(self):
        return self.train_data

    def loss(self, reduction="auto"):
        if isinstance(self._loss, str):
            return _NumpyLoss(self.loss_kind, reduction)
        try:
            return self._loss(reduction=reduction)
        except TypeError:
            return _NumpyLoss(self.loss_kind, reduction)

    def input_shape(self):
        return tuple(self.train_data.x.shape[1:])

    def feature_normalisation(self):
        if self.likelihood_model == "Regression":
            mean = self.train_data.x.mean(axis=0)
            std = self.train_data.x.astype(np.float64).std(axis=0) + 1e-8
            f = lambda d: ArrayDataset(((d.x - mean) / std), d.y)
        else:
            f = lambda d: ArrayDataset(d.x.astype(np.float32) / 255, d.y)
        self.train_data, self.valid_data, self.test_data = f(self.train_data), f(self.valid_data), f(self.test_data)

    def label_normalisation(self):
        if self.likelihood_model == "Regression":
            n = max(1, int(self.train_data.y.shape[0] / 10))
            self._label_mean = self.train_data.y[:n].mean()
            self._label_std = self.train_data.y[:n].astype(np.float64).std()
            g = lambda d: ArrayDataset(d.x, (d.y - self._label_mean) / (self._label_std + 1e-8))
            self.train_data, self.valid_data, self.test_data = g(self.train_data), g(self.valid_data), g(self.test_data)

    # ---- device view -----------------------------------------------------------------------
    def training_arrays(self):
        """(X float32 [N, D_flat], y) laid out for ``pyb_set_dataset``."""
        x = np.ascontiguousarray(self.train_data.x, dtype=np.float32).reshape(self.train_size, -1)
        if self.loss_kind == _lib.LOSS_SPARSE_CE:
            y = np.ascontiguousarray(self.train_data.y).reshape(-1).astype(np.int32)
        else:
            y = np.ascontiguousarray(self.train_data.y, dtype=np.float32).reshape(self.train_size, -1)
        return x, y

    def split_arrays(self, which):
        d = {"train": self.train_data, "valid": self.valid_data, "test": self.test_data}[which]
        x = np.ascontiguousarray(d.x, dtype=np.float32).reshape(d.x.shape[0], -1)
        if self.loss_kind == _lib.LOSS_SPARSE_CE:
            y = np.ascontiguousarray(d.y).reshape(-1).astype(np.int32)
        else:
            y = np.ascontiguousarray(d.y, dtype=np.float32).reshape(d.x.shape[0], -1)
        return x, y
