"""GPU tests of the drop-in Python surface: the reference's own script flow
(HMC_classification.py:36-66, SVGD_classification.py:150-166, HMC_regression.py:26-60) must run
unchanged against the B200 build and produce a usable BayesianModel."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from Pyesian.datasets import Dataset  # noqa: E402
from Pyesian.distributions import GaussianPrior, Sampled  # noqa: E402
from Pyesian.nn import BayesianModel  # noqa: E402
from Pyesian.optimizers import HMC, SVGD  # noqa: E402
from Pyesian.optimizers.hyperparameters import HyperParameters  # noqa: E402
from bayesian_inference_for_nn_b200 import keras_json  # noqa: E402


def moons(n, seed=0, noise=0.2):
    rng = np.random.default_rng(seed)
    n0 = n // 2
    t0, t1 = rng.uniform(0, np.pi, n0), rng.uniform(0, np.pi, n - n0)
    x = np.concatenate([np.stack([np.cos(t0), np.sin(t0)], 1), np.stack([1 - np.cos(t1), 0.5 - np.sin(t1)], 1)])
    y = np.concatenate([np.zeros(n0, np.int64), np.ones(n - n0, np.int64)])
    return x + rng.normal(0, noise, x.shape), y


MOONS_JSON = keras_json.make_sequential_json(2, [50, 2], ["relu", "softmax"])


def test_hmc_script_flow_learns_moons(tmp_path):
    x, y = moons(2000)
    dataset = Dataset((x, y), "SparseCategoricalCrossentropy", "Classification", seed=0)
    opt = HMC()
    opt.compile(HyperParameters(epsilon=0.005, m=0.5, L=30, n_chains=8, seed=1), MOONS_JSON, dataset, verbose=False,
                prior=GaussianPrior(0.0, 1.0))
    with pytest.raises(Exception, match="Model Already compiled"):
        opt.compile(HyperParameters(epsilon=0.005, m=0.5, L=30), MOONS_JSON, dataset, prior=GaussianPrior(0.0, 1.0))
    opt.train(60)
    assert 0.05 < opt.accept_rate <= 1.0
    bm = opt.result()
    assert isinstance(bm, BayesianModel)
    x_test, y_true = next(iter(dataset.test_data.batch(dataset.test_size)))
    samples, preds = bm.predict(x_test, nb_samples=100)
    assert len(samples) == 100 and samples[0].shape == (200, 2) and preds.shape == (200, 2)
    np.testing.assert_allclose(np.mean(samples, axis=0), preds, rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(np.var(samples, axis=0), bm.last_variance, rtol=1e-3, atol=1e-6)
    acc = (preds.argmax(1) == y_true).mean()
    assert acc > 0.85, acc          # logs/HMC_classification_FULL.txt reaches 95-98 % with the same settings
    # exact mode is within Monte-Carlo error of the reference-style draw
    _, exact = bm.predict(x_test, nb_samples=0, mode="exact")
    assert np.abs(exact - preds).max() < 0.2
    assert bm.uncertainty_mask(x_test, 50, 0.7).shape == (200,)
    # store / load round trip in the reference's folder format
    bm.store(str(tmp_path / "m"))
    bm2 = BayesianModel.load(str(tmp_path / "m"))
    _, exact2 = bm2.predict(x_test, nb_samples=0, mode="exact")
    np.testing.assert_array_equal(exact, exact2)
    m = bm.sample_model()
    assert m.predict(x_test).shape == (200, 2) and [w.shape for w in m.get_weights()] == [(2, 50), (50,), (50, 2), (2,)]


def test_hmc_single_chain_step_and_negative_sigma_quirk():
    x, y = moons(500, seed=1)
    dataset = Dataset((x, y), "SparseCategoricalCrossentropy", "Classification", seed=0)
    opt = HMC()
    opt.compile(HyperParameters(epsilon=0.005, m=0.5, L=5), MOONS_JSON, dataset, verbose=False,
                prior=GaussianPrior(0.0, -1.0))        # what the shipped scripts pass (HMC_classification.py:50)
    loss = opt.step(sampling=False, burning=True)
    assert np.isfinite(loss)
    opt.train(5)
    assert opt.accept_rate == 0.0                       # NaN Hamiltonian: nothing accepted after burn-in
    bm = opt.result()
    d = bm._distributions[0]
    assert isinstance(d, Sampled) and d.frequencies == [6]


def test_hmc_regression_flow():
    rng = np.random.default_rng(0)
    x = rng.uniform(1, 20, (600, 1)).astype(np.float32)
    dataset = Dataset((x, 2 * x + 2), "MeanSquaredError", "Regression", seed=0)
    js = keras_json.make_sequential_json(1, [1, 1], ["linear", "linear"])
    opt = HMC()
    opt.compile(HyperParameters(epsilon=5e-4, m=1.0, L=20, n_chains=4), js, dataset, verbose=False,
                prior=GaussianPrior(0.0, 1.0))
    opt.train(30)
    bm = opt.result()
    xt, yt = next(iter(dataset.test_data.batch(dataset.test_size)))
    _, pred = bm.predict(xt, 20)
    assert pred.shape == (60, 1) and np.isfinite(pred).all()


@pytest.mark.parametrize("semantics", ["reference", "canonical"])
def test_svgd_script_flow(semantics):
    x, y = moons(2000, seed=2)
    dataset = Dataset((x, y), "SparseCategoricalCrossentropy", "Classification", seed=0)
    js = keras_json.make_sequential_json(2, [64, 2], ["relu", "softmax"])
    opt = SVGD()
    lr = 1e-2
    opt.compile(HyperParameters(lr=lr, batch_size=64, M=10, semantics=semantics, seed=0), js, dataset, verbose=False,
                prior=GaussianPrior(0, 1))
    first = opt.step()
    opt.train(150)
    models, train_losses, valid_losses = opt.result()
    assert len(models) == 10 and len(train_losses) == 15 and len(valid_losses) == 15
    assert train_losses[-1] < first
    xt, yt = next(iter(dataset.test_data.batch(dataset.test_size)))
    preds = np.mean([m.predict(xt) for m in models], axis=0)
    res = opt.result()
    _, mean = res.predict(xt, mode="exact")
    np.testing.assert_allclose(preds, mean, rtol=1e-4, atol=1e-5)
    assert (mean.argmax(1) == yt).mean() > 0.8


@pytest.mark.parametrize("shape", [(2, [50, 2], ["relu", "softmax"], 333, "ce"),       # generic forward
                                   (4, [8, 3], ["tanh", "linear"], 100, "mse"),
                                   (784, [128, 10], ["relu", "softmax"], 640, "ce")])  # tensor path (scratch gradient)
def test_svgd_validation_loss_on_device(oracle, shape):
    """SVGD.py:126-129: every particle's loss over the whole validation split, evaluated against the resident particles"""
    from bayesian_inference_for_nn_b200 import _lib
    from bayesian_inference_for_nn_b200.engine import Engine
    D, units, acts, Nv, kind = shape
    rng = np.random.default_rng(0)
    S = 6
    spec_o = oracle.MLPSpec(D, units, acts)
    eng = Engine(keras_json.parse_model_json(keras_json.make_sequential_json(D, units, acts)))
    loss = _lib.LOSS_SPARSE_CE if kind == "ce" else _lib.LOSS_MSE
    mk_y = (lambda n: rng.integers(0, units[-1], n).astype(np.int32)) if kind == "ce" else \
        (lambda n: rng.normal(size=(n, units[-1])).astype(np.float32))
    X, y = rng.uniform(0, 1, (256, D)).astype(np.float32), mk_y(256)
    eng.set_dataset(X, y, loss)
    eng.set_prior([0.0], [1.0], _lib.PRIOR_SCALAR)
    parts = rng.normal(0, 0.3 if D < 100 else 0.05, (S, spec_o.n_params))
    eng.svgd_init(S, 1e-3, _lib.SVGD_CANONICAL_MEDIAN, particles0=parts)
    with pytest.raises(_lib.PyesianB200Error):
        eng.svgd_validation_loss()                          # no validation set yet
    Xv, yv = rng.uniform(0, 1, (Nv, D)).astype(np.float32), mk_y(Nv)
    eng.svgd_set_validation(Xv, yv)
    want, _ = oracle.mean_loss_and_grad(spec_o, parts.astype(np.float32), Xv, yv, loss, dtype=np.float64, want_grad=False)
    mean, per = eng.svgd_validation_loss(per_particle=True)
    np.testing.assert_allclose(per, want, rtol=1e-4)
    assert abs(mean - want.mean()) < 1e-4 * abs(want.mean())
    eng.svgd_step(np.arange(64, dtype=np.int32))              # the scratch gradient must not disturb the next step
    before = eng.svgd_particles()
    eng.svgd_validation_loss()
    np.testing.assert_array_equal(eng.svgd_particles(), before)
    eng.close()


def _two_gpus():
    from bayesian_inference_for_nn_b200 import _lib
    return _lib.device_count() >= 2


@pytest.mark.parametrize("shape", ["moons", "wide"])
def test_hmc_multi_device_flow(shape):
    """HyperParameters(devices=[0, 1]): the chains are sharded over two GPUs inside ONE optimizer (one engine and one host
    thread per device, global chain ids in the Philox counters) and result() pools them in global chain order — the
    samples, their frequencies and the accept rate are those of the one-device run, bit for bit."""
    if not _two_gpus():
        pytest.skip("needs 2 GPUs")
    if shape == "moons":
        x, y = moons(600, seed=3)
        js, hp = MOONS_JSON, dict(epsilon=0.005, m=0.5, L=10, n_chains=7, seed=5)
    else:
        rng = np.random.default_rng(0)
        x, y = rng.random((640, 784)), rng.integers(0, 10, 640)
        js, hp = keras_json.make_sequential_json(784, [256, 10], ["relu", "softmax"]), dict(epsilon=1e-3, m=1.0, L=3, n_chains=5, seed=5)
    runs = []
    for devices in ([0], [0, 1]):
        dataset = Dataset((x, y), "SparseCategoricalCrossentropy", "Classification", seed=0)
        opt = HMC()
        opt.compile(HyperParameters(devices=devices, **hp), js, dataset, verbose=False, prior=GaussianPrior(0.0, 1.0))
        opt.train(6)
        bm = opt.result()
        d = bm._distributions[0]
        runs.append((d.samples.copy(), list(d.frequencies), opt.accept_rate))
    (s1, f1, a1), (s2, f2, a2) = runs
    assert f1 == f2 and a1 == a2
    np.testing.assert_array_equal(s1, s2)


def test_svgd_multi_device_flow():
    """SVGD with devices=[0, 1]: the particles are sharded over two GPUs behind the same optimizer object; losses and
    particles follow the one-device run (same minibatches, same initial particles)."""
    if not _two_gpus():
        pytest.skip("needs 2 GPUs")
    x, y = moons(800, seed=4)
    rng = np.random.default_rng(1)
    P = 2 * 50 + 50 + 50 * 2 + 2
    p0 = rng.normal(0, 0.3, (8, P))
    runs = []
    for devices in ([0], [0, 1]):
        dataset = Dataset((x, y), "SparseCategoricalCrossentropy", "Classification", seed=0)
        opt = SVGD()
        opt.compile(HyperParameters(lr=1e-2, batch_size=64, M=8, semantics="canonical", seed=2, devices=devices), MOONS_JSON,
                    dataset, verbose=False, prior=GaussianPrior(0.0, 1.0), particles0=p0)
        losses = [opt.step() for _ in range(12)]
        runs.append((np.asarray(losses), opt.particles.copy()))
        models, tl, vl = opt.result()
        assert len(models) == 8 and len(tl) == 1 and len(vl) == 1
    np.testing.assert_allclose(runs[1][0], runs[0][0], rtol=1e-4)
    assert np.abs(runs[1][1] - runs[0][1]).max() < 2e-3 * 1e-2 * 12 + 1e-6
    with pytest.raises(ValueError):
        SVGD().compile(HyperParameters(lr=1e-2, batch_size=64, M=7, devices=[0, 1]), MOONS_JSON,
                       Dataset((x, y), "SparseCategoricalCrossentropy", "Classification", seed=0), verbose=False,
                       prior=GaussianPrior(0.0, 1.0))
