"""CPU tests for the SGLD / SWAG widening: the oracle's restatement of the reference's update
arithmetic against a literal single-model transcription of SGLD.step / SWAG.step (SGLD.py:60-94, SWAG.py:58-91), the
learning-rate schedule's end points (SGLD.py:115-121), the posterior distributions (sampling formula, reference
file formats) and the optimizers' host-side error behaviour — no GPU calls."""
import json
import os

import numpy as np
import pytest

from Pyesian.distributions import Mixture, MultivariateNormalDiagPlusLowRank, Normal
from Pyesian.nn import BayesianModel
from bayesian_inference_for_nn_b200 import keras_json


def literal_run(O, spec, theta, batches, kind, lrs, zs, k, frequency):
    """one model, the reference's statements in order (per layer == per flat vector, the update is element-wise)"""
    f = np.float32
    theta = theta.astype(f).copy()
    mean, sq_mean = np.zeros_like(theta), np.zeros_like(theta)
    dev = np.zeros((theta.shape[0], 0), f)
    losses = []
    for n, ((Xb, yb), lr, z) in enumerate(zip(batches, lrs, zs)):
        loss, g = O.mean_loss_and_grad(spec, theta[None], Xb, yb, O.LOSS_SPARSE_CE, dtype=np.float32)
        g = g[0].astype(f)
        losses.append(loss[0])
        if kind == O.SG_SGLD:
            noise = f(lr) * z                                   # tf.random.normal(stddev=lr)
            theta = theta + (-f(lr)) * (g + noise)              # var.assign_add(-lr * (grad + noise))
            update = True
        else:
            theta = theta - f(lr) * g                           # var.assign_sub(lr * grad)
            update = n % frequency == 0
        if update:
            mean = (mean * f(n) + theta) / (f(n) + f(1.0))
            sq_mean = (sq_mean * f(n) + theta ** 2) / (f(n) + f(1.0))
            col = (theta - mean)[:, None]
            if kind == O.SG_SWAG and dev.shape[1] == k:
                dev = np.concatenate((dev[:, :k - 1], col), axis=1)
            else:
                dev = np.concatenate((dev, col), axis=1)
    return theta, mean, sq_mean, dev, np.array(losses)


@pytest.mark.parametrize("kind", ["sgld", "swag"])
def test_oracle_matches_the_literal_single_model_loop(oracle, kind):
    O = oracle
    rng = np.random.default_rng(0)
    spec = O.MLPSpec(3, [5, 2], ["relu", "softmax"])
    S, steps, k, freq = 3, 9, 4, 2
    theta0 = rng.normal(0, 0.3, (S, spec.n_params)).astype(np.float32)
    batches = [(rng.normal(size=(16, 3)).astype(np.float32), rng.integers(0, 2, 16)) for _ in range(steps)]
    lrs = [0.05 / (1 + i) for i in range(steps)]
    zs = rng.normal(size=(steps, S, spec.n_params)).astype(np.float32)
    kk = O.SG_SGLD if kind == "sgld" else O.SG_SWAG
    st = O.sg_init_state(theta0)
    losses = [O.sg_step(spec, st, Xb, yb, O.LOSS_SPARSE_CE, kk, lr, z=zs[i], k=k, frequency=freq)
              for i, ((Xb, yb), lr) in enumerate(zip(batches, lrs))]
    for s in range(S):
        th, mean, sq, dev, ls = literal_run(O, spec, theta0[s], batches, kk, lrs, zs[:, s], k, freq)
        np.testing.assert_array_equal(st.theta[s], th)
        np.testing.assert_array_equal(st.mean[s], mean)
        np.testing.assert_array_equal(st.sq_mean[s], sq)
        np.testing.assert_allclose(np.array(losses)[:, s], ls, rtol=1e-6)
        if kind == "swag":
            assert len(st.dev) == k == dev.shape[1]           # ceil(9 / 2) = 5 updates > k: the last column was replaced
            np.testing.assert_array_equal(np.stack([c[s] for c in st.dev], axis=1), dev)
    assert st.n == steps


def test_sgld_schedule_end_points(oracle):
    lr = oracle.sgld_lr_schedule(500, 1e-2, 1e-4, 0.55)
    assert abs(lr(0) - 1e-2) < 1e-12 and abs(lr(500) - 1e-4) < 1e-12
    assert all(lr(i) > lr(i + 1) for i in range(0, 500, 50))
    from Pyesian.optimizers import SGLD
    from Pyesian.optimizers.hyperparameters import HyperParameters
    opt = SGLD()
    opt._hyperparameters = HyperParameters(lr_upper=1e-2, lr_lower=1e-4, lr_gamma=0.55)
    opt._lr_upper, opt._lr_lower, opt._lr_gamma, opt._nb_iterations = 1e-2, 1e-4, 0.55, 500
    opt._init_sgld_lr()
    assert all(abs(opt._lr(i) - lr(i)) < 1e-15 for i in (0, 1, 17, 499, 500))


def test_glorot_init_restatement(oracle):
    spec = oracle.MLPSpec(20, [30, 4], ["relu", "softmax"])
    th = oracle.glorot_uniform_init(spec, seed=7, chain_ids=[0, 1, 5])
    (w0, b0, _, _), (w1, b1, _, _) = spec.offsets()[0]
    l0, l1 = np.sqrt(6 / 50), np.sqrt(6 / 34)
    k0, k1 = th[:, w0:w0 + 600], th[:, w1:w1 + 120]
    assert np.all(np.abs(k0) < l0) and np.all(np.abs(k1) < l1) and np.abs(k0).max() > 0.9 * l0
    assert np.all(th[:, b0:b0 + 30] == 0) and np.all(th[:, b1:b1 + 4] == 0)
    assert abs(k0.mean()) < 0.02 and abs(k0.var() - l0 ** 2 / 3) < 0.01
    assert not np.array_equal(th[0], th[1])
    np.testing.assert_array_equal(th[2], oracle.glorot_uniform_init(spec, 7, [5])[0])


def test_distributions_sampling_and_files(tmp_path):
    rng = np.random.default_rng(0)
    n = Normal(np.arange(4.0), [0.0, 1.0, 2.0, 0.5], rng=np.random.default_rng(1))
    draws = np.stack([n.sample() for _ in range(4000)])
    np.testing.assert_allclose(draws.mean(0), np.arange(4.0), atol=0.12)
    np.testing.assert_allclose(draws.std(0), [0.0, 1.0, 2.0, 0.5], atol=0.08)
    assert np.isnan(Normal([0.0], [np.nan]).sample()).all()          # negative "variance" scale: NaN like tfp
    os.makedirs(tmp_path / "n")
    n.store(str(tmp_path / "n"))
    blob = json.load(open(tmp_path / "n" / "distribution.json"))
    assert blob["type"] == "Normal" and set(blob["params"]) >= {"loc", "scale"}      # BaseSerializer layout
    np.testing.assert_array_equal(Normal.load(str(tmp_path / "n")).scale, n.scale)

    # mean + diag * z1 + D z2 sqrt(1 / (2 (k - 1)))   (MultivariateNormalDiagPlusLowRank.py:32-41)
    D = rng.normal(size=(6, 3)).astype(np.float32)
    m = MultivariateNormalDiagPlusLowRank(np.ones(6), np.full(6, 0.1), D, rng=np.random.default_rng(2))
    draws = np.stack([m.sample() for _ in range(20000)])
    want_cov = np.diag(np.full(6, 0.01)) + D @ D.T / (2 * (3 - 1))
    np.testing.assert_allclose(np.cov(draws.T), want_cov, atol=0.06)
    os.makedirs(tmp_path / "m")
    m.store(str(tmp_path / "m"))
    assert set(json.load(open(tmp_path / "m" / "distribution.json"))) == {"mean", "D", "diag"}
    m2 = MultivariateNormalDiagPlusLowRank.load(str(tmp_path / "m"))
    np.testing.assert_array_equal(m2._D, m._D)

    mix = Mixture([Normal(np.zeros(2), 0.0), Normal(np.ones(2), 0.0)], rng=np.random.default_rng(3))
    vals = {tuple(mix.sample()) for _ in range(50)}
    assert vals == {(0.0, 0.0), (1.0, 1.0)}
    with pytest.raises(ValueError):
        Mixture([Normal(np.zeros(2), 1.0), Normal(np.zeros(3), 1.0)])


def test_bayesian_model_store_load_with_layer_distributions(tmp_path):
    js = keras_json.make_sequential_json(2, [3, 2], ["relu", "softmax"])
    spec = keras_json.parse_model_json(js)
    bm = BayesianModel(js)
    for d in spec.dense:
        lo, hi = spec.layer_param_range(d.keras_index, d.keras_index)
        bm.apply_distribution(Mixture([Normal(np.full(hi - lo, 0.5), 0.0), Normal(np.full(hi - lo, 0.5), 0.0)])
                              if d.keras_index else Normal(np.full(hi - lo, -1.0), 0.0), d.keras_index, d.keras_index)
    w = bm._draw_flat()
    assert np.all(w[:9] == -1.0) and np.all(w[9:] == 0.5)
    bm.store(str(tmp_path / "bm"))
    names = open(tmp_path / "bm" / "layers_config.txt").read().split()
    assert "TensorflowProbabilityDistribution" in names and "Mixture" in names      # the reference's class name
    bm2 = BayesianModel.load(str(tmp_path / "bm"))
    np.testing.assert_array_equal(bm2._draw_flat(), w)


def test_multi_chain_posterior_draws_one_chain_for_every_layer(tmp_path):
    """n_chains > 1: the per-layer mixtures share one chain selector, so a drawn network equals ONE chain's layers
    (independent chains are not exchangeable layer by layer) — also after store / load."""
    from bayesian_inference_for_nn_b200.distributions.Mixture import ChainSelector
    js = keras_json.make_sequential_json(2, [3, 2], ["relu", "softmax"])
    spec = keras_json.parse_model_json(js)
    n_chains = 5
    chains = np.arange(n_chains, dtype=np.float32)[:, None] + np.zeros((n_chains, spec.n_params), np.float32)
    bm = BayesianModel(js)
    sel = ChainSelector(n_chains, rng=np.random.default_rng(0))
    for d in spec.dense:
        lo, hi = spec.layer_param_range(d.keras_index, d.keras_index)
        bm.apply_distribution(Mixture([Normal(chains[c, lo:hi], 0.0) for c in range(n_chains)], selector=sel),
                              d.keras_index, d.keras_index)
    seen = set()
    for _ in range(60):
        w = bm._draw_flat()
        assert np.all(w == w[0]), w                  # every layer from the same chain
        seen.add(int(w[0]))
    assert seen == set(range(n_chains))
    bm.store(str(tmp_path / "bm"))
    bm2 = BayesianModel.load(str(tmp_path / "bm"))
    for _ in range(30):
        w = bm2._draw_flat()
        assert np.all(w == w[0])
    # unlinked mixtures keep drawing per layer
    bm3 = BayesianModel(js)
    for d in spec.dense:
        lo, hi = spec.layer_param_range(d.keras_index, d.keras_index)
        bm3.apply_distribution(Mixture([Normal(chains[c, lo:hi], 0.0) for c in range(n_chains)],
                                       rng=np.random.default_rng(d.keras_index)), d.keras_index, d.keras_index)
    assert any(not np.all((w := bm3._draw_flat()) == w[0]) for _ in range(30))


def test_swag_requires_a_starting_model():
    from Pyesian.datasets import Dataset
    from Pyesian.optimizers import SWAG
    from Pyesian.optimizers.hyperparameters import HyperParameters
    js = keras_json.make_sequential_json(2, [3, 2], ["relu", "softmax"])
    ds = Dataset((np.zeros((10, 2)), np.zeros(10, np.int64)), "SparseCategoricalCrossentropy", "Classification")
    with pytest.raises(KeyError):
        SWAG().compile(HyperParameters(lr=0.1, k=3, scale=1.0, frequency=1), js, ds)
    with pytest.raises(AttributeError):
        SWAG().compile(HyperParameters(lr=0.1), js, ds, starting_model=np.zeros(17))
