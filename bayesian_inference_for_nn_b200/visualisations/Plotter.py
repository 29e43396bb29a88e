"""Plotter — the quantities behind Pyesian's plots, evaluated on the device; drawing is optional.

Adapter for Pyesian/visualisations/Plotter.py:14-400.  Every public method of the reference keeps
its name and arguments and now RETURNS the arrays it plots, so it is usable without a display:

* ``plot_decision_boundaries`` (:169-196, ``_plot_2d_decision_boundary`` :100-119): the grid of ``_extract_grid_x``
  (:121-135; ~10^4 points at the default granularity) pushed through ``n_boundaries`` weight draws — one
  ``pyb_predict`` call instead of ``n_boundaries`` eager forward passes;
* ``plot_uncertainty_area`` (:198-226, ``_plot_2d_uncertainty_area`` :54-78): the mask ``max_c mean_c < threshold``
  over the same grid;
* ``regression_uncertainty`` (:228-259): ``np.var`` over the per-draw outputs is the population variance the device
  already reduces (``BayesianModel.last_variance``);
* ``entropy`` (:348-374), ``confusion_matrix`` (:262-283), ``roc_one_vs_rest`` (:137-167),
  ``compare_prediction_to_target`` (:286-321), ``learning_diagnostics`` (:378-393).

matplotlib is imported lazily and only when it is installed: with it, the figures are drawn/saved like the
reference's (``report/plots/<name>.png``, :395-400); without it the methods just return their data.
"""
import os

import numpy as np

from ..tensors import to_numpy


def _plt():
    try:
        import matplotlib.pyplot as plt
        return plt
    except Exception:
        return None


class Plotter:
    def __init__(self, model, dataset):
        self._dataset = dataset
        self._model = model
        self._nb_predictions = 0
        self._cached_samples = None
        self._cached_prediction = None
        self._cached_true_values = None
        self._cached_input = None
        self._cached_data_type = None
        self._cached_variance = None

    # ---- data + cache (Plotter.py:32-52, 80-98) ----------------------------------------------------------------
    def _two_class(self, y_pred):
        if y_pred.shape[1] == 1 and self._dataset.likelihood_model == "Classification":
            return np.concatenate([1 - y_pred, y_pred], axis=1)
        return y_pred

    def _get_predictions(self, input, nb_boundaries, y_true, data_type):
        y_true = to_numpy(y_true)
        if (self._nb_predictions == nb_boundaries and self._cached_true_values is not None
                and y_true.shape == self._cached_true_values.shape and data_type == self._cached_data_type):
            return (self._cached_samples, self._two_class(self._cached_prediction), self._cached_true_values,
                    self._cached_input)
        y_samples, y_pred = self._model.predict(input, nb_boundaries)
        self._nb_predictions = nb_boundaries
        self._cached_data_type = data_type
        self._cached_input = input
        self._cached_samples = y_samples
        self._cached_prediction = to_numpy(y_pred, np.float32)
        self._cached_true_values = y_true
        self._cached_variance = getattr(self._model, "last_variance", None)
        return y_samples, self._two_class(self._cached_prediction), y_true, input

    def _get_x_y(self, n_samples=100, data_type="test"):
        d = self._dataset.valid_data
        if data_type == "test":
            d = self._dataset.test_data
        elif data_type == "train":
            d = self._dataset.train_data
        return next(iter(d.batch(n_samples)))

    def _extract_x_y_from_dataset(self, dimension=2, n_samples=100, data_type="test"):
        x, y = self._get_x_y(n_samples, data_type)
        x, y = to_numpy(x), to_numpy(y)
        if x.shape[1] > dimension:
            # the reference calls a non-existent tf.pca here (:94); the principal axes of the centred rows are what
            # it describes ("Will apply PCA to reduce to dimension")
            print("Will apply PCA to reduce to dimension ", dimension)
            _, _, vt = np.linalg.svd(x - x.mean(axis=0), full_matrices=False)
            base = vt[:dimension].T.astype(x.dtype)
            return x @ base, y, base
        elif x.shape[1] < dimension:
            raise ValueError("Dimension ", x.shape[1], " is inferior to given dimension")
        return x, y, np.eye(dimension, dtype=x.dtype)

    @staticmethod
    def _range(start, limit, delta, dtype):
        """tf.range for floats: ceil(|limit - start| / delta) values start + i * delta, in the input dtype."""
        start, limit, delta = dtype.type(start), dtype.type(limit), dtype.type(delta)
        n = int(np.ceil(np.abs((limit - start) / delta)))
        return start + np.arange(n).astype(dtype) * delta

    def _extract_grid_x(self, x, base_matrix, granularity, un_zoom_level):
        x = to_numpy(x)
        dt = x.dtype if x.dtype.kind == "f" else np.dtype(np.float32)
        mx, mn = x.max(axis=0), x.min(axis=0)
        size1, size2 = mx[0] - mn[0], mx[1] - mn[1]
        dim1 = self._range(mn[0] - (un_zoom_level / 2) * size1, mx[0] + (un_zoom_level / 2) * size1,
                           granularity * (mx[0] - mn[0] + un_zoom_level * size1), dt)
        dim2 = self._range(mn[1] - (un_zoom_level / 2) * size2, mx[1] + (un_zoom_level / 2) * size2,
                           granularity * (mx[1] - mn[1] + un_zoom_level * size2), dt)
        dim1, dim2 = np.meshgrid(dim1, dim2, indexing="ij")
        grid_x = np.stack([dim1.reshape(-1), dim2.reshape(-1)], axis=1)
        return dim1, dim2, grid_x @ np.asarray(base_matrix, dtype=dt).T

    def _save(self, save_path, name):
        plt = _plt()
        if plt is None:
            return
        directory = os.path.join(save_path, "report")
        plots = os.path.join(directory, "plots")
        os.makedirs(plots, exist_ok=True)
        plt.savefig(os.path.join(plots, name + ".png"))

    def _finish(self, save_path, name):
        plt = _plt()
        if plt is None:
            return
        self._save(save_path, name) if save_path else plt.show()

    # ---- grids on the device -------------------------------------------------------------------------------------
    def decision_boundary_grids(self, x, base_matrix, granularity=1e-2, n_boundaries=10, un_zoom_level=0.2):
        """(dim1, dim2, [n_boundaries, g1, g2] class-0 probability per weight draw): the surfaces whose 0.5 contour
        the reference draws (:108-114)."""
        dim1, dim2, grid = self._extract_grid_x(x, base_matrix, granularity, un_zoom_level)
        samples, _ = self._model.predict(grid, n_boundaries)
        return dim1, dim2, np.stack([np.asarray(s)[:, 0].reshape(dim1.shape) for s in samples])

    def uncertainty_area(self, x, base_matrix, granularity=1e-2, n_samples=100, uncertainty_threshold=0.8,
                         un_zoom_level=0.2):
        """(dim1, dim2, float32 mask [g1, g2]) with mask = 1 where max_c mean_c < threshold (:57-72)."""
        dim1, dim2, grid = self._extract_grid_x(x, base_matrix, granularity, un_zoom_level)
        _, predictions = self._model.predict(grid, n_samples)
        predictions = np.asarray(predictions)
        if predictions.shape[1] == 1:
            predictions = np.concatenate([1 - predictions, predictions], axis=1)
        mask = (predictions.max(axis=1) < uncertainty_threshold).astype(np.float32).reshape(dim1.shape)
        return dim1, dim2, mask

    def plot_decision_boundaries(self, dimension=2, granularity=1e-2, n_boundaries=30, n_samples=100, data_type="test",
                                 un_zoom_level=0.2, save_path=None):
        if self._dataset.likelihood_model != "Classification":
            raise ValueError("Decision boundary can only be plotted for Classification")
        x, y, base = self._extract_x_y_from_dataset(dimension=dimension, n_samples=n_samples, data_type=data_type)
        if dimension != 2:
            raise ValueError("Decision boundary can only be plotted in 2 dimensions")
        n_boundaries = 10      # the reference passes a literal 10 whatever the argument says (:187-188)
        dim1, dim2, surfaces = self.decision_boundary_grids(x, base, granularity, n_boundaries, un_zoom_level)
        plt = _plt()
        if plt is not None:
            y1 = np.asarray(y).reshape(-1)
            plt.figure(figsize=(8, 6))
            plt.scatter(x[y1 == 0][:, 0], x[y1 == 0][:, 1], marker="o", c="blue", label="Class 0")
            plt.scatter(x[y1 == 1][:, 0], x[y1 == 1][:, 1], marker="x", c="red", label="Class 1")
            for pred in surfaces:
                plt.contour(dim1, dim2, pred, [0.5], colors=["red"])
            plt.legend()
            plt.xlabel("Feature 1")
            plt.ylabel("Feature 2")
            plt.title("Multiple Decision Boundaries N=" + str(n_boundaries))
            self._finish(save_path, "decision_boundaries")
        return dim1, dim2, surfaces

    def plot_uncertainty_area(self, dimension=2, granularity=1e-2, n_samples=100, data_type="test",
                              uncertainty_threshold=0.8, un_zoom_level=0.2, save_path=None):
        if self._dataset.likelihood_model != "Classification":
            raise ValueError("Uncertainty area can only be plotted for Classification")
        x, y, base = self._extract_x_y_from_dataset(dimension=dimension, n_samples=n_samples, data_type=data_type)
        if dimension != 2:
            return None
        dim1, dim2, mask = self.uncertainty_area(x, base, granularity, n_samples, uncertainty_threshold, un_zoom_level)
        plt = _plt()
        if plt is not None:
            y1 = np.asarray(y).reshape(-1)
            for i in range(np.unique(y1).shape[0]):
                plt.scatter(x[y1 == i][:, 0], x[y1 == i][:, 1], marker="o", label="Class " + str(i))
            plt.contourf(dim1, dim2, mask, [0.9, 1.1], colors=["orange"], alpha=0.5)
            plt.xlabel("Feature 1")
            plt.ylabel("Feature 2")
            plt.legend()
            plt.title("Uncertainty area with threshold " + str(uncertainty_threshold))
            self._finish(save_path, "uncertainty_area")
        return dim1, dim2, mask

    # ---- per-row quantities ----------------------------------------------------------------------------------------
    def regression_uncertainty(self, n_boundaries=30, n_samples=100, data_type="test", save_path=None):
        """-> (pred_dev, err): mean deviation per row and mean sqrt(variance over the draws) per row (:241-246)."""
        if self._dataset.likelihood_model != "Regression":
            raise ValueError("regression uncertainty cannot be computed for other than regression problems")
        x, y_true = self._get_x_y(n_samples, data_type)
        y_samples, y_pred, y_true, x = self._get_predictions(x, n_boundaries, y_true, data_type)
        variance = self._cached_variance if self._cached_variance is not None else np.var(np.asarray(y_samples), axis=0)
        err = np.mean(np.sqrt(variance), axis=1)
        pred_dev = np.mean(np.asarray(y_pred) - np.asarray(y_true).reshape(y_pred.shape), axis=1)
        plt = _plt()
        if plt is not None:
            plt.figure(figsize=(10, 5))
            plt.hlines([0], 0, len(err))
            plt.plot(range(len(err)), pred_dev - err, label="Epistemic Lower", alpha=0.5)
            plt.scatter(range(len(err)), pred_dev, label="Averaged deviation", alpha=0.5, c="k")
            plt.plot(range(len(err)), pred_dev + err, label="Epistemic Upper", alpha=0.5)
            plt.legend()
            plt.title("Epistemic Uncertainty")
            plt.ylabel("Pred-True difference")
            self._finish(save_path, "epistemic_uncertainty")
        return pred_dev, err

    def entropy(self, n_boundaries=30, n_samples=100, data_type="test", save_path=None):
        """-> sorted per-row entropies -sum p log(p + 1e-5) of the mean prediction (:362-366)."""
        x, y_true = self._get_x_y(n_samples, data_type)
        _, y_pred, y_true, x = self._get_predictions(x, n_boundaries, y_true, data_type)
        if self._dataset.likelihood_model != "Classification":
            raise Exception("Entropy is only available for classification")
        p = np.asarray(y_pred)
        entropies = np.sort(np.nan_to_num(-np.sum(p * np.log(p + 1e-5), axis=1)))
        plt = _plt()
        if plt is not None:
            plt.plot(range(len(y_true)), entropies)
            plt.title("Entropies for each input")
            plt.xlabel("Sample Index")
            plt.ylabel("entropy")
            self._finish(save_path, "entropy")
        return entropies

    def confusion_matrix(self, n_boundaries=30, n_samples=100, data_type="test", save_path=None):
        """-> row-normalised confusion matrix of argmax(mean prediction) (:276-281)."""
        import sklearn.metrics as skmet
        if self._dataset.likelihood_model != "Classification":
            raise ValueError("Confusion matrix cannot be computed for other than classification problems")
        x, y_true = self._get_x_y(n_samples, data_type)
        _, y_pred, y_true, x = self._get_predictions(x, n_boundaries, y_true, data_type)
        labels = np.asarray(y_pred).argmax(axis=1)
        cm = skmet.confusion_matrix(np.asarray(y_true).reshape(labels.shape), labels, normalize="true")
        plt = _plt()
        if plt is not None:
            skmet.ConfusionMatrixDisplay(cm).plot()
            plt.title("Confusion Matrix")
            self._finish(save_path, "confusion_matrix")
        return cm

    def roc_one_vs_rest(self, n_samples=100, label_of_interest: int = 0, n_boundaries=10, data_type="test"):
        """-> (fpr, tpr, thresholds) of class ``label_of_interest`` against the rest (:137-167)."""
        import sklearn.metrics as skmet
        if self._dataset.likelihood_model != "Classification":
            raise ValueError("ROC can only be plotted for Classification")
        x, y_true = self._get_x_y(n_samples, data_type)
        _, y_pred, y_true, x = self._get_predictions(x, n_boundaries, y_true, data_type)
        onehot = (np.asarray(y_true).reshape(-1) == label_of_interest).astype(np.int64)
        fpr, tpr, thr = skmet.roc_curve(onehot, np.asarray(y_pred)[:, label_of_interest])
        plt = _plt()
        if plt is not None:
            skmet.RocCurveDisplay(fpr=fpr, tpr=tpr).plot()
            plt.title("One-vs-Rest ROC curve")
            plt.show()
        return fpr, tpr, thr

    def compare_prediction_to_target(self, n_boundaries=30, n_samples=100, data_type="test", save_path=None):
        """-> (y_true, prediction): regression means, or argmax labels for classification (:297-321)."""
        x, y_true = self._get_x_y(n_samples, data_type)
        _, y_pred, y_true, x = self._get_predictions(x, n_boundaries, y_true, data_type)
        y_pred = np.asarray(y_pred)
        if self._dataset.likelihood_model == "Regression":
            y_true = np.asarray(y_true).reshape(y_pred.shape)
            plt = _plt()
            if plt is not None and y_true.shape[1] == 1:
                plt.figure(figsize=(10, 5))
                plt.scatter(range(len(y_true)), y_true, label="True Values", alpha=0.5)
                plt.scatter(range(len(y_pred)), y_pred, label="Predicted Mean", alpha=0.5)
                plt.legend()
                plt.title("True vs Predicted Values")
                self._finish(save_path, "comparison_pred_true")
            return y_true, y_pred
        return np.asarray(y_true).reshape(-1), y_pred.argmax(axis=1)

    def learning_diagnostics(self, loss_file: str, save_path=None):
        if loss_file is None:
            return None
        losses = np.loadtxt(loss_file)
        plt = _plt()
        if plt is not None:
            plt.plot(losses)
            plt.title("Training Loss")
            plt.xlabel("Iterations")
            plt.ylabel("Loss")
            self._finish(save_path, "learning_diagnostics")
        return losses
