"""GPU parity of the S-batched SGLD / SWAG chains through the C ABI against the oracle's float32
restatement of SGLD.step / SWAG.step (SGLD.py:46-95, SWAG.py:43-94) on identical minibatches and injected noise, the
device RNG against its Philox restatement, and the drop-in script flow."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from Pyesian.datasets import Dataset  # noqa: E402
from Pyesian.nn import BayesianModel  # noqa: E402
from Pyesian.optimizers import SGLD, SWAG  # noqa: E402
from Pyesian.optimizers.hyperparameters import HyperParameters  # noqa: E402
from bayesian_inference_for_nn_b200 import _lib, keras_json  # noqa: E402
from bayesian_inference_for_nn_b200.engine import Engine  # noqa: E402
from conftest import rel_err  # noqa: E402

SHAPES = [(2, [50, 2], ["relu", "softmax"], 400, 64, _lib.LOSS_SPARSE_CE),
          (5, [16, 8, 1], ["tanh", "relu", "linear"], 300, 50, _lib.LOSS_MSE),
          (784, [128, 10], ["relu", "softmax"], 2048, 512, _lib.LOSS_SPARSE_CE)]     # minibatch gradients on the tensor path


def setup(oracle, shape, S, seed=0):
    D, units, acts, N, B, loss = shape
    rng = np.random.default_rng(seed)
    spec_o = oracle.MLPSpec(D, units, acts)
    eng = Engine(keras_json.parse_model_json(keras_json.make_sequential_json(D, units, acts)), seed=11)
    X = rng.uniform(0, 1, (N, D)).astype(np.float32)
    y = rng.integers(0, units[-1], N).astype(np.int32) if loss == _lib.LOSS_SPARSE_CE else \
        rng.normal(size=(N, units[-1])).astype(np.float32)
    eng.set_dataset(X, y, loss)
    theta0 = rng.normal(0, 0.3 if D < 100 else 0.05, (S, spec_o.n_params)).astype(np.float32)
    return eng, spec_o, X, y, theta0, rng, B, loss


@pytest.mark.parametrize("shape", SHAPES)
def test_sgld_steps_match_oracle_with_injected_noise(oracle, shape):
    S, steps = 5, 6
    eng, spec, X, y, theta0, rng, B, loss = setup(oracle, shape, S)
    eng.sg_init(S, _lib.SG_SGLD, theta0=theta0)
    st = oracle.sg_init_state(theta0)
    lr = oracle.sgld_lr_schedule(steps, 1e-2, 1e-3, 0.55)
    for n in range(steps):
        idx = rng.permutation(X.shape[0])[:B].astype(np.int32)
        z = rng.standard_normal((S, spec.n_params)).astype(np.float32)
        want_loss = oracle.sg_step(spec, st, X[idx], y[idx], loss, oracle.SG_SGLD, lr(n), z=z)
        got_loss, got_mean = eng.sg_step(lr(n), idx, noise=z)
        np.testing.assert_allclose(got_loss, want_loss, rtol=1e-4)
        assert abs(got_mean - want_loss.astype(np.float64).mean()) < 1e-4 * abs(got_mean) + 1e-7
    got = eng.sg_state()
    assert got["n"] == steps and got["dev"] is None
    assert rel_err(got["theta"], st.theta) < 2e-5
    assert rel_err(got["mean"], st.mean) < 2e-5
    assert rel_err(got["sq_mean"], st.sq_mean) < 4e-5
    # the element pass itself is exact: the device moments are the float32 recurrences of the device's own iterates
    eng.close()


def test_update_pass_is_bit_exact_given_the_gradient(oracle):
    """with lr = 0 the parameters never move, so mean / sq_mean / dev must equal the oracle's recurrences bit for bit"""
    S, steps, k, freq = 3, 7, 3, 2
    eng, spec, X, y, theta0, rng, B, loss = setup(oracle, SHAPES[0], S)
    eng.sg_init(S, _lib.SG_SWAG, k_dev=k, frequency=freq, theta0=theta0)
    st = oracle.sg_init_state(theta0)
    for n in range(steps):
        idx = rng.permutation(X.shape[0])[:B].astype(np.int32)
        oracle.sg_step(spec, st, X[idx], y[idx], loss, oracle.SG_SWAG, 0.0, k=k, frequency=freq)
        eng.sg_step(0.0, idx)
    got = eng.sg_state()
    np.testing.assert_array_equal(got["theta"], theta0)
    np.testing.assert_array_equal(got["mean"], st.mean)
    np.testing.assert_array_equal(got["sq_mean"], st.sq_mean)
    np.testing.assert_array_equal(got["dev"], np.stack(st.dev, axis=1))
    eng.close()


@pytest.mark.parametrize("shape", SHAPES[:2])
def test_swag_steps_match_oracle(oracle, shape):
    S, steps, k, freq = 4, 11, 4, 2
    eng, spec, X, y, theta0, rng, B, loss = setup(oracle, shape, S, seed=1)
    eng.sg_init(S, _lib.SG_SWAG, k_dev=k, frequency=freq, theta0=theta0[:1])      # one starting model, broadcast
    st = oracle.sg_init_state(np.repeat(theta0[:1], S, axis=0))
    for n in range(steps):
        idx = rng.permutation(X.shape[0])[:B].astype(np.int32)
        want = oracle.sg_step(spec, st, X[idx], y[idx], loss, oracle.SG_SWAG, 0.05, k=k, frequency=freq)
        got, _ = eng.sg_step(0.05, idx)
        np.testing.assert_allclose(got, want, rtol=1e-4)
    got = eng.sg_state()
    assert got["dev"].shape == (S, k, spec.n_params)       # 6 moment updates > k: the last column was overwritten twice
    assert rel_err(got["theta"], st.theta) < 2e-5 and rel_err(got["mean"], st.mean) < 2e-5
    dev = np.stack(st.dev, axis=1)
    assert np.abs(got["dev"] - dev).max() < 2e-5 * max(1.0, np.abs(st.theta).max())
    # identical chains stay identical (same start, same minibatches, no noise)
    np.testing.assert_array_equal(got["theta"][0], got["theta"][-1])
    eng.close()


def test_device_rng_init_and_langevin_noise(oracle):
    S, off = 4, 3
    eng, spec, X, y, _, rng, B, loss = setup(oracle, SHAPES[0], S)
    eng.sg_init(S, _lib.SG_SGLD, chain_offset=off)
    th0 = eng.sg_state()["theta"]
    want0 = oracle.glorot_uniform_init(spec, 11, np.arange(off, off + S))
    np.testing.assert_allclose(th0, want0, rtol=0, atol=1e-7)
    st = oracle.sg_init_state(th0)
    for n in range(3):
        idx = rng.permutation(X.shape[0])[:B].astype(np.int32)
        z = oracle.philox_normals(11, np.arange(off, off + S), n, oracle.STREAM_SGLD, spec.n_params)
        oracle.sg_step(spec, st, X[idx], y[idx], loss, oracle.SG_SGLD, 0.3, z=z)
        eng.sg_step(0.3, idx)
    # lr = 0.3 makes the noise term (lr^2 z) dominate the update: a wrong stream / counter would show at 1e-1
    assert rel_err(eng.sg_state()["theta"], st.theta) < 1e-4
    eng.close()


def moons(n, seed=0, noise=0.2):
    rng = np.random.default_rng(seed)
    n0 = n // 2
    t0, t1 = rng.uniform(0, np.pi, n0), rng.uniform(0, np.pi, n - n0)
    x = np.concatenate([np.stack([np.cos(t0), np.sin(t0)], 1), np.stack([1 - np.cos(t1), 0.5 - np.sin(t1)], 1)])
    y = np.concatenate([np.zeros(n0, np.int64), np.ones(n - n0, np.int64)])
    return x + rng.normal(0, noise, x.shape), y


MOONS_JSON = keras_json.make_sequential_json(2, [50, 2], ["relu", "softmax"])


def test_sgld_script_flow(tmp_path):
    x, y = moons(2000)
    ds = Dataset((x, y), "SparseCategoricalCrossentropy", "Classification", seed=0)
    opt = SGLD()
    opt.compile(HyperParameters(batch_size=128, lr_upper=0.3, lr_lower=0.05, lr_gamma=0.55, n_chains=4, seed=0),
                MOONS_JSON, ds, verbose=False)
    with pytest.raises(TypeError):
        SGLD.step(opt)                      # the schedule exists only once train() fixed the horizon (SGLD.py:124-126)
    opt.train(1500, loss_save_document_path=str(tmp_path / "loss.txt"))
    bm = opt.result()
    assert isinstance(bm, BayesianModel) and len(bm._distributions) == 2
    xt, yt = next(iter(ds.test_data.batch(ds.test_size)))
    samples, preds = bm.predict(xt, nb_samples=50)
    assert len(samples) == 50 and preds.shape == (200, 2)
    assert (preds.argmax(1) == yt).mean() > 0.8
    assert (tmp_path / "loss.txt").exists()
    bm.store(str(tmp_path / "m"))
    assert BayesianModel.load(str(tmp_path / "m")).predict(xt, 5)[1].shape == (200, 2)


def test_swag_script_flow():
    x, y = moons(2000)
    ds = Dataset((x, y), "SparseCategoricalCrossentropy", "Classification", seed=0)
    spec = keras_json.parse_model_json(MOONS_JSON)
    rng = np.random.default_rng(0)
    start = [rng.uniform(-0.3, 0.3, (2, 50)).astype(np.float32), np.zeros(50, np.float32),
             rng.uniform(-0.3, 0.3, (50, 2)).astype(np.float32), np.zeros(2, np.float32)]      # get_weights() order
    opt = SWAG()
    opt.compile(HyperParameters(batch_size=128, lr=0.1, k=10, scale=0.5, frequency=5), MOONS_JSON, ds, verbose=False,
                starting_model=start)
    first = opt.step()
    opt.train(1200)
    st = opt.chains
    assert st["n"] == 1201 and st["dev"].shape == (1, 10, spec.n_params)
    bm = opt.result()
    xt, yt = next(iter(ds.test_data.batch(ds.test_size)))
    _, preds = bm.predict(xt, nb_samples=30)
    assert (preds.argmax(1) == yt).mean() > 0.8 and opt.step() < first


def test_errors_and_edge_cases(oracle):
    """status codes across the boundary; a partial last minibatch; full-dataset steps (batch_idx = NULL)"""
    from bayesian_inference_for_nn_b200._lib import PyesianB200Error
    eng, spec, X, y, theta0, rng, B, loss = setup(oracle, SHAPES[0], 2)
    with pytest.raises(PyesianB200Error) as e:
        eng.sg_step(0.1)
    assert e.value.code == -3                                   # PYB_ERR_STATE: step before init
    with pytest.raises(PyesianB200Error) as e:
        eng.sg_init(2, _lib.SG_SWAG, k_dev=1, frequency=1)
    assert e.value.code == -1                                   # SWAG needs k >= 2
    with pytest.raises(PyesianB200Error):
        eng.sg_init(2, 7)
    with pytest.raises(ValueError):
        eng.sg_init(3, _lib.SG_SGLD, theta0=theta0)              # 2 rows for 3 chains
    eng.sg_init(2, _lib.SG_SGLD, theta0=theta0)
    with pytest.raises(PyesianB200Error) as e:
        eng.sg_step(0.1, np.array([0, X.shape[0]], np.int32))
    assert e.value.code == -1                                   # row index out of range
    st = oracle.sg_init_state(theta0)
    z = np.zeros((2, spec.n_params), np.float32)
    for idx in (np.arange(7, dtype=np.int32), None):            # ragged 7-row batch, then the whole dataset
        Xb, yb = (X, y) if idx is None else (X[idx], y[idx])
        want = oracle.sg_step(spec, st, Xb, yb, loss, oracle.SG_SGLD, 0.05, z=z)
        got, _ = eng.sg_step(0.05, idx, noise=z)
        np.testing.assert_allclose(got, want, rtol=1e-4)
    assert rel_err(eng.sg_state()["theta"], st.theta) < 2e-5
    eng.close()


def test_sharded_sgld_chains_reproduce_the_unsharded_run(oracle):
    """chains never interact and the Philox counters use the GLOBAL chain id (chain_offset): two shards of 3 chains
    driven with the same minibatches give exactly the 6-chain run (first row, applied to the SG chains)"""
    S = 6
    eng, spec, X, y, _, rng, B, loss = setup(oracle, SHAPES[0], S)
    batches = [rng.permutation(X.shape[0])[:B].astype(np.int32) for _ in range(5)]
    eng.sg_init(S, _lib.SG_SGLD)
    for n, idx in enumerate(batches):
        eng.sg_step(0.05 / (n + 1), idx)
    whole = eng.sg_state()
    parts = []
    for r in range(2):
        e2 = Engine(keras_json.parse_model_json(keras_json.make_sequential_json(2, [50, 2], ["relu", "softmax"])), seed=11)
        e2.set_dataset(X, y, loss)
        e2.sg_init(S // 2, _lib.SG_SGLD, chain_offset=r * (S // 2))
        for n, idx in enumerate(batches):
            e2.sg_step(0.05 / (n + 1), idx)
        parts.append(e2.sg_state())
        e2.close()
    for k in ("theta", "mean", "sq_mean"):
        np.testing.assert_array_equal(np.concatenate([p[k] for p in parts]), whole[k])
    eng.close()
