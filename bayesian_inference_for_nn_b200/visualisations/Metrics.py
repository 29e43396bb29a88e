"""Metrics — performance / uncertainty metrics of a BayesianModel over a Dataset, predictions on the device.

Adapter for Pyesian/visualisations/Metrics.py:10-413.  Same surface and behaviour: the
prediction cache keyed on ``n_boundaries`` and the label shape (:27-45), ``summary`` (:47-76), the regression scores
(:81-196), the classification scores (:201-333) including the reference's swapped ``precision``/``recall`` calls
(:252 uses ``recall_score`` macro, :279 uses ``precision_score`` micro), ``ece`` with the *probabilities* handed to
``tfp.stats.expected_calibration_error`` as logits (:331), ``auroc`` (:377-402), ``_save`` (:404-411).

What runs where: every score is a few flops on an ``[n_samples, C]`` mean and stays on the host (scikit-learn, as in
the reference).  The two heavy parts are on the GPU: the ``n_boundaries`` forward passes (one ``pyb_predict`` call,
``BayesianModel.predict``) and ``classification_uncertainty`` (:344-375), a Python double loop over draws x rows
building C x C matrices in the reference and ONE ``pyb_predict_uncertainty`` call here, evaluated on the very same
weight draws the cached predictions came from.
"""
import os

import numpy as np

from ..tensors import to_numpy


def _softmax(z):
    z = z - z.max(axis=-1, keepdims=True)
    e = np.exp(z)
    return e / e.sum(axis=-1, keepdims=True)


def expected_calibration_error(num_bins, logits, labels_true):
    """tfp.stats.expected_calibration_error as Metrics.py:331 calls it (equal-width confidence bins over the softmax
    of ``logits``; sum over bins of count/N * |accuracy - mean confidence|)."""
    p = _softmax(np.asarray(logits, dtype=np.float64))
    pred, conf = p.argmax(axis=-1), p.max(axis=-1)
    correct = (pred == np.asarray(labels_true).reshape(-1)).astype(np.float64)
    bins = np.clip(np.floor(conf * num_bins).astype(np.int64), 0, num_bins - 1)
    count = np.bincount(bins, minlength=num_bins).astype(np.float64)
    acc = np.bincount(bins, weights=correct, minlength=num_bins)
    cs = np.bincount(bins, weights=conf, minlength=num_bins)
    nz = count > 0
    return float(np.sum(np.abs(acc[nz] - cs[nz])) / p.shape[0])


class Metrics:
    def __init__(self, model, dataset):
        self._model = model
        self._dataset = dataset
        self._nb_predictions = 0
        self._cached_samples = None
        self._cached_prediction = None
        self._cached_true_values = None
        self._cached_input = None
        self._cached_draws = None

    # ---- data + cached predictions (Metrics.py:27-45, 335-342) -----------------------------------------------
    def _get_x_y(self, n_samples=100, data_type="test"):
        d = self._dataset.valid_data
        if data_type == "test":
            d = self._dataset.test_data
        elif data_type == "train":
            d = self._dataset.train_data
        x, y_true = next(iter(d.batch(n_samples)))
        return x, y_true

    def _two_class(self, y_pred):
        if y_pred.shape[1] == 1 and self._dataset.likelihood_model == "Classification":
            return np.concatenate([1 - y_pred, y_pred], axis=1)
        return y_pred

    def _get_predictions(self, input, n_boundaries, y_true):
        y_true = to_numpy(y_true)
        if (self._nb_predictions == n_boundaries and self._cached_true_values is not None
                and y_true.shape == self._cached_true_values.shape):
            return (self._cached_samples, self._two_class(self._cached_prediction), self._cached_true_values,
                    self._cached_input)
        if self._cached_draws is not None:
            self._cached_draws.free()
        draws = self._model.draw(n_boundaries) if hasattr(self._model, "draw") else None
        if draws is not None:
            y_samples, y_pred = self._model.predict(input, n_boundaries, draws=draws)
        else:
            y_samples, y_pred = self._model.predict(input, n_boundaries)
        self._nb_predictions = n_boundaries
        self._cached_draws = draws
        self._cached_input = input
        self._cached_samples = y_samples
        self._cached_prediction = to_numpy(y_pred, np.float32)
        self._cached_true_values = y_true
        return y_samples, self._two_class(self._cached_prediction), y_true, input

    def _save(self, save_path, name, content):
        if save_path is not None:
            directory = os.path.join(save_path, "report")
            os.makedirs(directory, exist_ok=True)
            with open(os.path.join(directory, name), "w") as f:
                f.write(str(content))

    def _score(self, kind, n_boundaries, n_samples, data_type):
        classification = self._dataset.likelihood_model == "Classification"
        if kind == "regression" and classification:
            raise Exception("this metric could only be computed for regression")
        if kind == "classification" and not classification:
            raise Exception("this metric could only be computed for classification")
        input, y_true = self._get_x_y(n_samples=n_samples, data_type=data_type)
        return self._get_predictions(input, n_boundaries, y_true)

    def summary(self, n_boundaries=30, n_samples=100, data_type="test", save_path=None):
        kw = dict(n_boundaries=n_boundaries, n_samples=n_samples, data_type=data_type, save_path=save_path)
        if self._dataset.likelihood_model == "Regression":
            for f in (self.mse, self.rmse, self.mae, self.r2, self.log_likeliood):
                f(**kw)
        elif self._dataset.likelihood_model == "Classification":
            for f in (self.accuracy, self.recall, self.precision, self.f1_score, self.auroc, self.ece):
                f(**kw)
        else:
            print("Invalid loss function")

    # ---- regression (Metrics.py:81-196) ----------------------------------------------------------------------
    def _regression(self, name, label, fn, n_boundaries, n_samples, data_type, save_path):
        _, y_pred, y_true, _ = self._score("regression", n_boundaries, n_samples, data_type)
        res = fn(np.asarray(y_true, dtype=np.float64).reshape(y_pred.shape), np.asarray(y_pred, dtype=np.float64))
        self._save(save_path, name, res)
        print("{}: {}".format(label, res))
        return res

    def mse(self, n_boundaries=30, n_samples=100, data_type="test", save_path=None):
        import sklearn.metrics as skmet
        return self._regression("MSE", "MSE", skmet.mean_squared_error, n_boundaries, n_samples, data_type, save_path)

    def rmse(self, n_boundaries=30, n_samples=100, data_type="test", save_path=None):
        import sklearn.metrics as skmet
        return self._regression("RMSE", "RMSE", skmet.root_mean_squared_error, n_boundaries, n_samples, data_type,
                                save_path)

    def mae(self, n_boundaries=30, n_samples=100, data_type="test", save_path=None):
        import sklearn.metrics as skmet
        return self._regression("MAE", "MAE", skmet.mean_absolute_error, n_boundaries, n_samples, data_type, save_path)

    def r2(self, n_boundaries=30, n_samples=100, data_type="test", save_path=None):
        import sklearn.metrics as skmet
        return self._regression("R2", "R2 score", skmet.r2_score, n_boundaries, n_samples, data_type, save_path)

    def log_likeliood(self, n_boundaries=30, n_samples=100, data_type="test", save_path=None):
        """mean log N(y_pred; y_true, 1) (Metrics.py:193-194), in float32 like the reference."""
        def fn(y_true, y_pred):
            d = (y_pred - y_true).astype(np.float32)
            return float(np.mean(np.float32(-0.5) * d * d - np.float32(0.5 * np.log(2.0 * np.pi))))
        return self._regression("log_likelihood", "log likelihood", fn, n_boundaries, n_samples, data_type, save_path)

    # ---- classification (Metrics.py:201-333, 377-402) ----------------------------------------------------------
    def _labels(self, n_boundaries, n_samples, data_type):
        _, y_pred, y_true, _ = self._score("classification", n_boundaries, n_samples, data_type)
        return np.asarray(y_true).reshape(-1), y_pred

    def accuracy(self, n_boundaries=30, n_samples=100, data_type="test", save_path=None):
        import sklearn.metrics as skmet
        y_true, y_pred = self._labels(n_boundaries, n_samples, data_type)
        res = skmet.accuracy_score(y_true, y_pred.argmax(axis=1)) * 100
        self._save(save_path, "Accuracy", res)
        print("Accuracy: {}%".format(res))
        return res

    def precision(self, n_boundaries=30, n_samples=100, data_type="test", save_path=None):
        import sklearn.metrics as skmet
        y_true, y_pred = self._labels(n_boundaries, n_samples, data_type)
        res = skmet.recall_score(y_true, y_pred.argmax(axis=1), average="macro") * 100      # sic, Metrics.py:252
        self._save(save_path, "Precision", res)
        print("Precision: {}%".format(res))
        return res

    def recall(self, n_boundaries=30, n_samples=100, data_type="test", save_path=None):
        import sklearn.metrics as skmet
        y_true, y_pred = self._labels(n_boundaries, n_samples, data_type)
        res = skmet.precision_score(y_true, y_pred.argmax(axis=1), average="micro") * 100   # sic, Metrics.py:279
        self._save(save_path, "Recall", res)
        print("Recall: {}%".format(res))
        return res

    def f1_score(self, n_boundaries=30, n_samples=100, data_type="test", save_path=None):
        import sklearn.metrics as skmet
        y_true, y_pred = self._labels(n_boundaries, n_samples, data_type)
        res = skmet.f1_score(y_true, y_pred.argmax(axis=1), average="macro")
        self._save(save_path, "F1_score", res)
        print("F1 score: {}".format(res))
        return res

    def ece(self, n_boundaries=30, n_samples=100, data_type="test", save_path=None, n_bins=5):
        y_true, y_pred = self._labels(n_boundaries, n_samples, data_type)
        res = expected_calibration_error(n_bins, logits=y_pred, labels_true=y_true)
        self._save(save_path, "ECE", res)
        print("ECE: {}".format(res))
        return res

    def auroc(self, n_boundaries=10, n_samples=100, data_type="test", save_path=None, multi_class="ovr"):
        import sklearn.metrics as skmet
        if self._dataset.likelihood_model != "Classification":
            raise ValueError("ROC can only be plotted for Classification")
        y_true, y_pred = self._labels(n_boundaries, n_samples, data_type)
        one_hot = np.eye(y_pred.shape[1])[y_true.astype(np.int64)]
        res = skmet.roc_auc_score(one_hot, y_pred, average="micro", multi_class=multi_class)
        self._save(save_path, "AUROC", res)
        print("AUROC: {}".format(res))
        return res

    def classification_uncertainty(self, n_boundaries=30, n_samples=100, data_type="test", save_path=None,
                                   semantics="reference"):
        """-> (epistemics + aleatorics, aleatorics, epistemics), each [rows, C, C] (Metrics.py:344-375).
        ``semantics="canonical"`` returns the per-row decomposition with (p - onehot)(p - onehot)^T instead of what the
        reference's code computes (running sums over the rows, label-free broadcast epistemic term)."""
        if self._dataset.likelihood_model != "Classification":
            raise Exception("only for classification")
        input, y_true = self._get_x_y(n_samples=n_samples, data_type=data_type)
        self._get_predictions(input, n_boundaries, y_true)          # fills / re-uses the cache and its weight draws
        return self._model.classification_uncertainty(self._cached_input, self._cached_true_values, n_boundaries,
                                                      divisor=n_samples, semantics=semantics,
                                                      draws=self._cached_draws)
