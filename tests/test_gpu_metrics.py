"""GPU parity of the Metrics / Plotter device work through the C ABI: pyb_predict_uncertainty
against the oracle's restatement of Metrics.classification_uncertainty (Metrics.py:344-375) on the same weight
samples, and the adapters driven the way the reference's scripts drive them."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from Pyesian.datasets import Dataset  # noqa: E402
from Pyesian.distributions import Sampled  # noqa: E402
from Pyesian.nn import BayesianModel  # noqa: E402
from Pyesian.visualisations import Metrics, Plotter  # noqa: E402
from bayesian_inference_for_nn_b200 import keras_json  # noqa: E402
from bayesian_inference_for_nn_b200.engine import Engine  # noqa: E402

TOL = 2e-5   # float32 forward pass against the float64 oracle; the matrices are sums of products of probabilities


def per_draw_outputs(O, spec_json, W, x):
    spec = keras_json.parse_model_json(spec_json)
    ospec = O.MLPSpec(spec.in_dim, [d.units for d in spec.dense], [d.activation for d in spec.dense],
                      [d.use_bias for d in spec.dense])
    return O.forward(ospec, W.astype(np.float64), x.astype(np.float64), dtype=np.float64)


@pytest.mark.parametrize("shape", [(2, [16, 3], ["relu", "softmax"], 9, 700),      # rows cross the 256-row scan segments
                                   (2, [50, 2], ["relu", "softmax"], 30, 100),     # the reference's moons net
                                   (3, [8, 1], ["tanh", "sigmoid"], 5, 300),       # one-unit output widened to 2 classes
                                   (784, [128, 10], ["relu", "softmax"], 12, 512)])  # tensor path forward
def test_uncertainty_matches_oracle(oracle, shape):
    D, units, acts, n, Nt = shape
    js = keras_json.make_sequential_json(D, units, acts)
    eng = Engine(keras_json.parse_model_json(js))
    rng = np.random.default_rng(0)
    W = rng.normal(0, 0.3 if D < 100 else 0.05, (n, eng.P)).astype(np.float32)
    x = rng.uniform(0, 1, (Nt, D)).astype(np.float32)
    Ce = 2 if units[-1] == 1 else units[-1]
    y = rng.integers(0, Ce, Nt)
    outs = per_draw_outputs(oracle, js, W, x)
    for semantics in ("reference", "canonical"):
        want = oracle.classification_uncertainty(outs, y, n_samples_arg=100, semantics=semantics)
        tot, al, ep, mean = eng.predict_uncertainty(W, x, y, semantics=semantics, divisor=100)
        for got, ref in zip((tot, al, ep), want):
            assert got.shape == (Nt, Ce, Ce)
            assert np.abs(got - ref).max() <= TOL * max(1.0, np.abs(ref).max())
        np.testing.assert_allclose(mean, outs.mean(axis=0), atol=2e-5)
    # integer multiplicities == repeated draws
    w = rng.integers(1, 4, n).astype(np.float32)
    a = eng.predict_uncertainty(W, x, y, weights=w, divisor=Nt)
    b = oracle.classification_uncertainty(outs, y, Nt, weights=w)
    assert np.abs(a[0] - b[0]).max() <= TOL * max(1.0, np.abs(b[0]).max())
    # device-resident samples give the same bits as host samples
    Wd = eng.device_array(W)
    c = eng.predict_uncertainty(Wd, x, y, weights=w, divisor=Nt)
    np.testing.assert_array_equal(a[0], c[0])
    with pytest.raises(Exception):
        eng.predict_uncertainty(W, x, np.full(Nt, Ce), divisor=Nt)     # label out of range


def moons(n, seed=0, noise=0.2):
    rng = np.random.default_rng(seed)
    n0 = n // 2
    t0, t1 = rng.uniform(0, np.pi, n0), rng.uniform(0, np.pi, n - n0)
    x = np.concatenate([np.stack([np.cos(t0), np.sin(t0)], 1), np.stack([1 - np.cos(t1), 0.5 - np.sin(t1)], 1)])
    y = np.concatenate([np.zeros(n0, np.int64), np.ones(n - n0, np.int64)])
    return x + rng.normal(0, noise, x.shape), y


def test_metrics_and_plotter_flow(oracle, capsys):
    js = keras_json.make_sequential_json(2, [50, 2], ["relu", "softmax"])
    spec = keras_json.parse_model_json(js)
    x, y = moons(1000)
    ds = Dataset((x, y), "SparseCategoricalCrossentropy", "Classification", seed=0)
    rng = np.random.default_rng(1)
    bm = BayesianModel(js)
    bm.apply_distribution(Sampled(rng.normal(0, 0.5, (40, spec.n_params)).astype(np.float32),
                                  rng.integers(1, 5, 40).tolist()), 0, spec.n_keras_layers - 1)
    m = Metrics(bm, ds)
    acc = m.accuracy(n_boundaries=25, n_samples=80)
    tot, al, ep = m.classification_uncertainty(n_boundaries=25, n_samples=80)
    # computed on the SAME weight draws as the cached predictions: the oracle on the cached per-draw outputs agrees
    want = oracle.classification_uncertainty(np.stack(m._cached_samples), m._cached_true_values, n_samples_arg=80)
    for got, ref in zip((tot, al, ep), want):
        assert got.shape == (80, 2, 2)
        assert np.abs(got - ref).max() <= TOL * max(1.0, np.abs(ref).max())
    assert 0 <= acc <= 100
    m.summary(n_boundaries=25, n_samples=80)
    # Plotter: ~10^4-point grid through 100 draws in one device call
    pl = Plotter(bm, ds)
    d1, d2, mask = pl.plot_uncertainty_area(n_samples=100, uncertainty_threshold=0.8)
    assert mask.shape == d1.shape and d1.size > 9000
    d1, d2, surf = pl.plot_decision_boundaries(n_samples=100)
    assert surf.shape == (10,) + d1.shape and np.all((surf >= 0) & (surf <= 1))
    # exact-mode mask equals the oracle's mask of the frequency-weighted mean away from the threshold
    d = bm._distributions[0]
    ospec = oracle.MLPSpec(2, [50, 2], [oracle.ACT_RELU, oracle.ACT_SOFTMAX], [True, True])
    _, _, grid = pl._extract_grid_x(*pl._extract_x_y_from_dataset(2, 100, "test")[::2], 1e-2, 0.2)
    mean, _ = oracle.predictive(ospec, d.samples, grid.astype(np.float32), weights=d.frequencies)
    got = bm.uncertainty_mask(grid, 0, 0.8, mode="exact")
    clear = np.abs(mean.max(axis=-1) - 0.8) > 1e-4
    np.testing.assert_array_equal(got[clear], oracle.uncertainty_mask(mean, 0.8)[clear])
    capsys.readouterr()
