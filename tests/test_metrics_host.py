"""CPU tests for the Metrics / Plotter adapters: the oracle's restatement of
Metrics.classification_uncertainty against a literal transcription of the reference's loops (Metrics.py:344-375; the
oracle is additionally pinned to the output of the reference method itself in tests/test_reference_goldens.py),
the calibration error against a hand-computed case, and the host logic of both classes (cache, error behaviour,
reference quirks) with a stub model — no GPU calls."""
import numpy as np
import pytest

from Pyesian.datasets import Dataset
from Pyesian.visualisations import Metrics, Plotter
from bayesian_inference_for_nn_b200.visualisations.Metrics import expected_calibration_error


def literal_uncertainty(y_samples, y_true, n_samples):
    """the reference's loop nest, statement by statement, with NumPy in place of tf"""
    aleatorics = 0
    epistemics = 0
    for sample in y_samples:
        aleatoric = 0
        epistemic = 0
        nb_classes = sample.shape[1]
        aleatorics_tmp, epistemics_tmp = [], []
        for prediction, label in zip(sample, y_true):
            col = prediction.reshape(-1, 1)
            aleatoric = aleatoric + (np.diag(prediction) - col @ col.T)
            dev = col - np.eye(nb_classes)[label]       # [C,1] - [C] broadcasts to [C,C], as tf does at Metrics.py:362
            epistemic = epistemic + dev @ dev.T
            epistemics_tmp.append(epistemic)
            aleatorics_tmp.append(aleatoric)
        aleatorics = aleatorics + np.asarray(aleatorics_tmp)
        epistemics = epistemics + np.asarray(epistemics_tmp)
    epistemics = epistemics / n_samples
    aleatorics = aleatorics / n_samples
    return epistemics + aleatorics, aleatorics, epistemics


def probs(rng, n, N, C):
    z = rng.normal(size=(n, N, C))
    e = np.exp(z - z.max(-1, keepdims=True))
    return e / e.sum(-1, keepdims=True)


def test_oracle_uncertainty_matches_the_literal_loops(oracle):
    rng = np.random.default_rng(0)
    for (n, N, C) in [(1, 1, 2), (3, 7, 2), (5, 33, 4), (2, 300, 10)]:
        s = probs(rng, n, N, C)
        y = rng.integers(0, C, N)
        want = literal_uncertainty(list(s), y, 100)
        got = oracle.classification_uncertainty(s, y, n_samples_arg=100)
        for a, b in zip(got, want):
            np.testing.assert_allclose(a, b, rtol=1e-12, atol=1e-14)
    # the aleatoric part of the canonical form = first differences of the reference's running sums; weights = repeated draws
    s = probs(rng, 4, 9, 3)
    y = rng.integers(0, 3, 9)
    cum = oracle.classification_uncertainty(s, y, 9)[1]
    row = oracle.classification_uncertainty(s, y, 9, semantics="canonical")[1]
    np.testing.assert_allclose(np.diff(cum, axis=0, prepend=0 * cum[:1]), row, atol=1e-14)
    # the reference's epistemic term does not depend on the labels at all (its broadcast sums over every class)
    e1 = oracle.classification_uncertainty(s, y, 9)[2]
    e2 = oracle.classification_uncertainty(s, (y + 1) % 3, 9)[2]
    np.testing.assert_allclose(e1, e2, atol=1e-13)
    rep = np.concatenate([s, s[:1], s[:1]])
    a = oracle.classification_uncertainty(rep, y, 9)
    b = oracle.classification_uncertainty(s, y, 9, weights=[3, 1, 1, 1])
    np.testing.assert_allclose(a[0], b[0], atol=1e-13)


def test_uncertainty_identities(oracle):
    """diag(p) - p p^T has zero row sums and is PSD; the epistemic part is a sum of outer products."""
    rng = np.random.default_rng(1)
    s = probs(rng, 6, 20, 5)
    y = rng.integers(0, 5, 20)
    tot, al, ep = oracle.classification_uncertainty(s, y, 20, semantics="canonical")
    assert np.abs(al.sum(axis=-1)).max() < 1e-14
    assert min(np.linalg.eigvalsh(m).min() for m in al) > -1e-12
    assert min(np.linalg.eigvalsh(m).min() for m in ep) > -1e-12
    # one-unit output is widened to [1 - p, p]
    p1 = rng.uniform(size=(3, 8, 1))
    a = oracle.classification_uncertainty(p1, rng.integers(0, 2, 8), 8)
    assert a[0].shape == (8, 2, 2)


def test_expected_calibration_error(oracle):
    # two rows, 2 bins; softmax is applied to whatever is passed (the reference passes probabilities, Metrics.py:331)
    logits = np.log(np.array([[0.9, 0.1], [0.6, 0.4], [0.55, 0.45], [0.2, 0.8]]))
    y = np.array([0, 1, 0, 1])
    # confidences .9 .6 .55 .8, all in the upper bin of 2: acc = 3/4, conf = .7125
    want = abs(0.75 - 0.7125)
    assert abs(expected_calibration_error(2, logits, y) - want) < 1e-12
    assert abs(oracle.expected_calibration_error(2, logits, y) - want) < 1e-12
    # 5 bins: .9 and .8 fall in bin 4 (acc 1, conf .85); .6 and .55 in bin 2 (.55 -> floor(2.75)=2, .6 -> floor(3.0)=3)
    rng = np.random.default_rng(3)
    z = rng.normal(size=(200, 4))
    yy = rng.integers(0, 4, 200)
    assert abs(expected_calibration_error(5, z, yy) - oracle.expected_calibration_error(5, z, yy)) < 1e-12


class StubModel:
    """answers like BayesianModel without touching a GPU: fixed per-draw probabilities"""

    def __init__(self, C, seed=0, regression=False):
        self.C, self.rng, self.calls, self.regression = C, np.random.default_rng(seed), 0, regression
        self.last_variance = None

    def predict(self, x, nb_samples, **kw):
        self.calls += 1
        x = np.asarray(x)
        if self.regression:
            s = self.rng.normal(size=(nb_samples, x.shape[0], self.C)).astype(np.float32)
        else:
            s = probs(self.rng, nb_samples, x.shape[0], self.C).astype(np.float32)
        self.last_variance = s.var(axis=0)
        return [a for a in s], s.mean(axis=0)


def moons_dataset(n=400):
    rng = np.random.default_rng(0)
    x = rng.normal(size=(n, 2))
    y = (x[:, 0] + x[:, 1] > 0).astype(np.int64)
    return Dataset((x, y), "SparseCategoricalCrossentropy", "Classification", seed=0)


def test_metrics_cache_and_scores(capsys, tmp_path):
    ds = moons_dataset()
    model = StubModel(2)
    m = Metrics(model, ds)
    acc = m.accuracy(n_boundaries=7, n_samples=30)
    assert model.calls == 1 and 0 <= acc <= 100
    m.f1_score(n_boundaries=7, n_samples=30)
    m.ece(n_boundaries=7, n_samples=30)
    assert model.calls == 1                      # same n_boundaries and label shape => cached (Metrics.py:28-31)
    m.accuracy(n_boundaries=8, n_samples=30)
    assert model.calls == 2
    # the reference's swapped calls are kept: "precision" is macro recall, "recall" is micro precision
    import sklearn.metrics as skmet
    x, y = next(iter(ds.test_data.batch(30)))
    pred = m._cached_prediction.argmax(1)
    assert m.precision(n_boundaries=8, n_samples=30) == skmet.recall_score(y, pred, average="macro") * 100
    assert m.recall(n_boundaries=8, n_samples=30) == skmet.precision_score(y, pred, average="micro") * 100
    m.summary(n_boundaries=8, n_samples=30, save_path=str(tmp_path))
    assert sorted(p.name for p in (tmp_path / "report").iterdir()) == ["AUROC", "Accuracy", "ECE", "F1_score", "Precision",
                                                                      "Recall"]
    with pytest.raises(Exception):
        m.mse()
    capsys.readouterr()


def test_metrics_regression(capsys):
    rng = np.random.default_rng(0)
    x = rng.normal(size=(200, 3))
    y = x.sum(axis=1, keepdims=True)
    ds = Dataset((x, y), "MeanSquaredError", "Regression", seed=0)
    m = Metrics(StubModel(1, regression=True), ds)
    mse, rmse = m.mse(n_boundaries=5, n_samples=20), m.rmse(n_boundaries=5, n_samples=20)
    assert abs(rmse - np.sqrt(mse)) < 1e-9
    yt = m._cached_true_values.reshape(-1, 1)
    ll = m.log_likeliood(n_boundaries=5, n_samples=20)
    want = np.mean(-0.5 * (m._cached_prediction - yt) ** 2 - 0.5 * np.log(2 * np.pi))
    assert abs(ll - want) < 1e-5
    with pytest.raises(Exception):
        m.accuracy()
    with pytest.raises(Exception, match="only for classification"):
        m.classification_uncertainty()
    capsys.readouterr()


def test_plotter_grid_and_masks():
    ds = moons_dataset()
    model = StubModel(2)
    pl = Plotter(model, ds)
    x, y, base = pl._extract_x_y_from_dataset(2, 50, "test")
    np.testing.assert_array_equal(base, np.eye(2))
    dim1, dim2, grid = pl._extract_grid_x(x, base, 1e-2, 0.2)
    assert dim1.shape == dim2.shape and grid.shape == (dim1.size, 2) and 99 <= dim1.shape[0] <= 101
    # the grid spans the data range widened by un_zoom_level / 2 on either side (Plotter.py:121-131)
    span = x[:, 0].max() - x[:, 0].min()
    assert abs(dim1.min() - (x[:, 0].min() - 0.1 * span)) < 1e-12
    d1, d2, mask = pl.plot_uncertainty_area(n_samples=40, uncertainty_threshold=0.8)
    assert mask.shape == d1.shape and set(np.unique(mask)) <= {0.0, 1.0}
    d1, d2, surf = pl.plot_decision_boundaries(n_boundaries=30, n_samples=40)
    assert surf.shape == (10,) + d1.shape          # the reference draws 10 boundaries whatever is asked (:187-188)
    with pytest.raises(ValueError, match="2 dimensions"):
        pl.plot_decision_boundaries(dimension=1)
    ent = pl.entropy(n_boundaries=5, n_samples=40)
    assert ent.shape == (40,) and np.all(np.diff(ent) >= 0)
    cm = pl.confusion_matrix(n_boundaries=5, n_samples=40)
    assert cm.shape == (2, 2)
    with pytest.raises(ValueError):
        pl.regression_uncertainty()
    # more than 2 features: projected on the principal axes
    rng = np.random.default_rng(0)
    x5 = rng.normal(size=(300, 5)) * np.array([5, 3, 1, 0.1, 0.1])
    ds5 = Dataset((x5, (x5[:, 0] > 0).astype(np.int64)), "SparseCategoricalCrossentropy", "Classification", seed=0)
    xp, _, base5 = Plotter(StubModel(2), ds5)._extract_x_y_from_dataset(2, 100, "train")
    assert xp.shape == (100, 2) and base5.shape == (5, 2)
    np.testing.assert_allclose(base5.T @ base5, np.eye(2), atol=1e-12)


def test_plotter_regression_uncertainty():
    rng = np.random.default_rng(0)
    x = rng.normal(size=(200, 3))
    ds = Dataset((x, x.sum(axis=1, keepdims=True)), "MeanSquaredError", "Regression", seed=0)
    model = StubModel(1, regression=True)
    pl = Plotter(model, ds)
    dev, err = pl.regression_uncertainty(n_boundaries=6, n_samples=25)
    s = np.asarray(pl._cached_samples)
    np.testing.assert_allclose(err, np.mean(np.sqrt(np.var(s, axis=0)), axis=1), rtol=1e-6)
    np.testing.assert_allclose(dev, np.mean(s.mean(0) - pl._cached_true_values.reshape(-1, 1), axis=1), rtol=1e-5,
                               atol=1e-6)
