"""CPU tests of the host-side mirror of the reference API (no GPU, no compute)."""
import json
import os
import random

import numpy as np
import pytest

from bayesian_inference_for_nn_b200 import _lib, keras_json
from bayesian_inference_for_nn_b200.datasets import ArrayDataset, Dataset
from bayesian_inference_for_nn_b200.distributions import GaussianPrior, Sampled
from bayesian_inference_for_nn_b200.nn import tensorproto
from bayesian_inference_for_nn_b200.optimizers.hyperparameters import HyperParameters
from bayesian_inference_for_nn_b200 import tensors

# Keras-2.15 Sequential JSON of the shape the GUI stores (InputLayer + Dense x2, build_config)
KERAS_215 = json.dumps({"class_name": "Sequential", "config": {"name": "sequential", "layers": [
    {"module": "keras.layers", "class_name": "InputLayer", "config": {"batch_input_shape": [None, 3], "dtype": "float32",
                                                                      "sparse": False, "ragged": False, "name": "dense_input"}},
    {"module": "keras.layers", "class_name": "Dense", "config": {"name": "dense", "units": 16, "activation": "relu",
                                                                  "use_bias": True, "batch_input_shape": [None, 3]},
     "build_config": {"input_shape": [None, 3]}},
    {"module": "keras.layers", "class_name": "Dense", "config": {"name": "dense_1", "units": 2, "activation": "relu",
                                                                  "use_bias": True}, "build_config": {"input_shape": [None, 16]}}]},
    "keras_version": "2.15.0", "backend": "tensorflow"})


def test_packer_layout_matches_keras_order():
    s = keras_json.parse_model_json(KERAS_215)
    assert (s.in_dim, s.n_params, s.n_keras_layers) == (3, 3 * 16 + 16 + 16 * 2 + 2, 2)
    assert [(d.w_off, d.b_off) for d in s.dense] == [(0, 48), (64, 96)]
    assert s.variables() == [(0, 0, 0, (3, 16)), (0, 1, 48, (16,)), (1, 0, 64, (16, 2)), (1, 1, 96, (2,))]
    assert s.layer_param_range(1, 1) == (64, 98) and s.layer_param_range(0, 1) == (0, 98)


def test_packer_flatten_and_no_bias_and_errors():
    mnist = json.dumps({"class_name": "Sequential", "config": {"layers": [
        {"class_name": "Flatten", "config": {"batch_input_shape": [None, 28, 28]}},
        {"class_name": "Dense", "config": {"units": 128, "activation": "relu"}},
        {"class_name": "Dense", "config": {"units": 10, "activation": "softmax", "use_bias": False}}]}})
    s = keras_json.parse_model_json(mnist)
    assert s.in_dim == 784 and s.n_keras_layers == 3 and s.keras_layer_kinds == ["Flatten", "Dense", "Dense"]
    assert s.n_params == 784 * 128 + 128 + 128 * 10 and s.dense[1].b_off == -1 and s.dense[0].keras_index == 1
    noshape = json.dumps({"class_name": "Sequential", "config": {"layers": [
        {"class_name": "Dense", "config": {"units": 4, "activation": "tanh"}}]}})
    with pytest.raises(keras_json.UnsupportedModelError):
        keras_json.parse_model_json(noshape)
    assert keras_json.parse_model_json(noshape, in_dim=6).n_params == 28
    conv = json.dumps({"class_name": "Sequential", "config": {"layers": [
        {"class_name": "Conv2D", "config": {"batch_input_shape": [None, 8, 8, 1]}}]}})
    with pytest.raises(keras_json.UnsupportedModelError):
        keras_json.parse_model_json(conv)
    with pytest.raises(keras_json.UnsupportedModelError):
        keras_json.parse_model_json("not json")
    rt = keras_json.parse_model_json(keras_json.make_sequential_json(2, [50, 2], ["relu", "softmax"]))
    assert rt.n_params == 252 and rt.dense[1].activation == _lib.ACT_SOFTMAX


def test_hyperparameters_surface():
    hp = HyperParameters(epsilon=0.005, m=0.5, L=30)
    assert (hp.epsilon, hp.m, hp.L, hp.batch_size) == (0.005, 0.5, 30, 64)
    with pytest.raises(AttributeError):
        hp.lr
    parsed = HyperParameters().parse("lr: 0.01\nbatch_size: 128\nk 10\nfrequency = -2.5")
    assert parsed.lr == 0.01 and parsed.batch_size == 128.0 and parsed.k == 10.0 and parsed.frequency == -2.5


def test_gaussian_prior_forms_and_errors():
    s = keras_json.parse_model_json(KERAS_215)
    with pytest.raises(Exception, match="same type"):
        GaussianPrior(0, 1.0)
    mu, sg, form = GaussianPrior(0.0, -1.0).lower(s)
    assert form == _lib.PRIOR_SCALAR and mu[0] == 0 and sg[0] == -1       # rho used RAW, sign kept
    mu, sg, form = GaussianPrior([0.0, 1.0], [1.0, 2.0]).lower(s)
    assert form == _lib.PRIOR_PER_ELEMENT and set(mu[:64]) == {0.0} and set(mu[64:]) == {1.0} and set(sg[64:]) == {2.0}
    tm = [[np.zeros((3, 16)), np.ones(16)], [np.full((16, 2), 2.0), np.full(2, 3.0)]]
    ts = [[np.ones((3, 16)), np.ones(16)], [np.ones((16, 2)), np.full(2, 0.5)]]
    mu, sg, form = GaussianPrior(tm, ts).lower(s)
    assert mu[47] == 0 and mu[48] == 1 and mu[64] == 2 and mu[97] == 3 and sg[97] == 0.5
    with pytest.raises(Exception, match="shape of the mean"):
        GaussianPrior([[np.zeros((2, 2)), np.ones(16)], tm[1]], ts).lower(s)
    with pytest.raises(Exception):
        GaussianPrior("a", "b").lower(s)
    pri = GaussianPrior(0.0, 1.0).get_model_priors(keras_json.parse_model_json(json.dumps(
        {"class_name": "Sequential", "config": {"layers": [
            {"class_name": "Flatten", "config": {"batch_input_shape": [None, 2, 2]}},
            {"class_name": "Dense", "config": {"units": 3}}]}})))
    assert pri[0] is None and len(pri[1]) == 2 and pri[1][0][0].shape == (4, 3)


def test_sampled_draw_is_frequency_weighted_and_validates():
    with pytest.raises(ValueError):
        Sampled([], [])
    with pytest.raises(ValueError):
        Sampled([np.zeros(2)], [1, 2])
    with pytest.raises(ValueError):
        Sampled([np.zeros(2), np.ones(2)], [1, 0])
    d = Sampled([np.zeros(3), np.ones(3), np.full(3, 2.0)], [1, 7, 2])
    random.seed(0)
    idx = np.array([d.sample_index() for _ in range(20000)])
    frac = np.bincount(idx, minlength=3) / idx.size
    np.testing.assert_allclose(frac, [0.1, 0.7, 0.2], atol=0.015)
    assert d.sample().shape == (3,) and d.size() == 3


def test_dataset_split_and_surface():
    x = np.arange(2000 * 2, dtype=np.float64).reshape(2000, 2)
    y = (np.arange(2000) % 2).astype(np.int64)
    ds = Dataset((x, y), "SparseCategoricalCrossentropy", "Classification", seed=0)
    assert (ds.size, ds.train_size, ds.test_size, ds.valid_size) == (2000, 1600, 200, 200)
    xb, yb = next(iter(ds.test_data.batch(ds.test_size)))
    assert xb.shape == (200, 2) and yb.shape == (200,)
    assert int(ds.training_dataset().cardinality()) == 1600 and ds.input_shape() == (2,)
    xt, yt = ds.training_arrays()
    assert xt.dtype == np.float32 and yt.dtype == np.int32 and xt.shape == (1600, 2)
    # the three splits partition the data
    allx = np.concatenate([ds.train_data.x, ds.test_data.x, ds.valid_data.x])
    assert sorted(allx[:, 0].tolist()) == x[:, 0].tolist()
    with pytest.raises(ValueError):
        Dataset((x, y), "SparseCategoricalCrossentropy", train_proportion=0.5)
    with pytest.raises(ValueError):
        Dataset((x, y), "Hinge")
    reg = Dataset((x[:, :1], 2 * x[:, :1] + 2), "MeanSquaredError", "Regression")
    assert reg.loss_kind == _lib.LOSS_MSE and reg.training_arrays()[1].shape == (1600, 1)
    assert abs(float(reg.loss()(np.zeros((4, 1)), np.ones((4, 1)))) - 1.0) < 1e-7

    class SparseCategoricalCrossentropy:      # a Keras-like loss class is recognised by name
        def __init__(self, reduction="auto"):
            self.reduction = reduction
    assert Dataset((x, y), SparseCategoricalCrossentropy).loss(reduction="none").reduction == "none"


def test_tensorproto_roundtrip_and_known_bytes():
    a = np.arange(6, dtype=np.float32)
    buf = tensorproto.serialize_tensor(a)
    # dtype DT_FLOAT(1), shape {dim{size:6}}, 24 content bytes
    assert buf[:2] == b"\x08\x01" and buf[2:8] == b"\x12\x04\x12\x02\x08\x06" and buf[8:10] == b"\x22\x18"
    np.testing.assert_array_equal(tensorproto.parse_tensor(buf), a)
    for arr in (np.random.default_rng(0).standard_normal((3, 300)), np.int32([[1, -2]]), np.int64([2 ** 40])):
        np.testing.assert_array_equal(tensorproto.parse_tensor(tensorproto.serialize_tensor(arr)), arr)


def test_sampled_store_load_roundtrip(tmp_path):
    d = Sampled(np.random.default_rng(1).standard_normal((4, 10)).astype(np.float32), [1, 2, 3, 4])
    d.store(str(tmp_path))
    info = json.load(open(os.path.join(tmp_path, "info.json")))
    assert info["n_samples"] == 4 and info["size"] == 10 and info["dtypes"] == ["float32"] * 4
    d2 = Sampled.load(str(tmp_path))
    np.testing.assert_array_equal(d2.samples, d.samples)
    assert d2.frequencies == [1, 2, 3, 4]


def test_dlpack_ingestion_host():
    torch = pytest.importorskip("torch")
    t = torch.arange(12, dtype=torch.float32).reshape(3, 4)
    a, mem, ptr = tensors.ingest(t, np.float32)
    assert mem == _lib.MEM_HOST and a.shape == (3, 4) and ptr == a.ctypes.data
    np.testing.assert_array_equal(a, t.numpy())
    cap = torch.utils.dlpack.to_dlpack(torch.arange(5, dtype=torch.float64))
    a, mem, _ = tensors.ingest(cap, np.float32)           # legacy capsule, cast on the host
    assert mem == _lib.MEM_HOST and a.dtype == np.float32 and a.tolist() == [0, 1, 2, 3, 4]
    with pytest.raises(ValueError):
        tensors.ingest(torch.utils.dlpack.to_dlpack(torch.zeros(4, 4)[:, ::2]), np.float32)   # not contiguous
    a, mem, _ = tensors.ingest([[1, 2], [3, 4]], np.int32)
    assert a.dtype == np.int32 and mem == _lib.MEM_HOST


def test_array_dataset_batching():
    d = ArrayDataset(np.arange(10).reshape(10, 1), np.arange(10))
    assert [b[0].shape[0] for b in d.batch(4)] == [4, 4, 2]
    assert int(d.batch(4).cardinality()) == 3 and int(d.skip(3).take(2).cardinality()) == 2


def test_compile_fails_loudly_without_gpu_and_only_once():
    from bayesian_inference_for_nn_b200.optimizers import HMC
    if _lib.device_count() > 0:
        pytest.skip("GPU present")
    x = np.random.default_rng(0).standard_normal((100, 2))
    ds = Dataset((x, (x[:, 0] > 0).astype(int)), "SparseCategoricalCrossentropy")
    opt = HMC()
    mj = keras_json.make_sequential_json(2, [5, 2], ["relu", "softmax"])
    with pytest.raises(KeyError):       # prior kwarg is mandatory (HMC.py:60)
        HMC().compile(HyperParameters(epsilon=0.01, m=1, L=2), mj, ds)
    with pytest.raises(_lib.PyesianB200Error, match="no CPU fallback"):
        opt.compile(HyperParameters(epsilon=0.01, m=1, L=2), mj, ds, prior=GaussianPrior(0.0, 1.0))
    with pytest.raises(Exception, match="Model Already compiled"):
        opt.compile(HyperParameters(epsilon=0.01, m=1, L=2), mj, ds, prior=GaussianPrior(0.0, 1.0))
    with pytest.raises(AttributeError):
        HMC().compile(HyperParameters(epsilon=0.01), mj, ds, prior=GaussianPrior(0.0, 1.0))


def test_hyperparameters_parse_matches_the_reference_goldens():
    """tests/golden/hyperparams_parse.json was produced by importing the REFERENCE's HyperParameters.py (pure Python) in the
    build container (tests/golden/make_hyperparams_golden.py): same parameters, same failure modes, case by case."""
    import json
    import os
    from conftest import GOLDEN
    from Pyesian.optimizers.hyperparameters import HyperParameters
    cases = json.load(open(os.path.join(GOLDEN, "hyperparams_parse.json")))
    assert len(cases) > 400
    for c in cases:
        try:
            got = ("ok", dict(HyperParameters(seed_kw=3).parse(c["text"])._params))
        except Exception as e:
            got = ("err", type(e).__name__)
        want = ("ok", c["params"]) if c["status"] == "ok" else ("err", c["error"])
        assert got == want, (c["text"], got, want)
    assert any(c["status"] == "err" for c in cases)            # a name without a number is an IndexError, kept


def test_dataset_from_a_csv_file_and_a_dataframe(tmp_path):
    """Dataset.py:124-133: the last `target_dim` columns are the labels; a CSV path goes through read_csv."""
    pd = pytest.importorskip("pandas")
    rng = np.random.default_rng(0)
    df = pd.DataFrame({"a": rng.random(50), "b": rng.random(50), "t": rng.random(50)})
    path = tmp_path / "reg.csv"
    df.to_csv(path, index=False)
    from_df = Dataset(df, "MeanSquaredError", "Regression", seed=3)
    from_csv = Dataset(str(path), "MeanSquaredError", "Regression", seed=3)
    assert from_csv.size == 50 and from_csv.input_shape() == (2,)
    np.testing.assert_allclose(from_csv.train_data.x, from_df.train_data.x, rtol=1e-12)
    np.testing.assert_allclose(from_csv.train_data.y, from_df.train_data.y, rtol=1e-12)
    assert from_csv.train_data.y.shape == (40, 1)
    two = Dataset(df, "MeanSquaredError", "Regression", target_dim=2)
    assert two.input_shape() == (1,) and two.train_data.y.shape == (40, 2)
    with pytest.raises(ValueError):
        Dataset("no_such_builder", "MeanSquaredError", "Regression")
