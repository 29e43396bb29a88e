"""pyb_set_dataset: a caller that re-submits the SAME data every step (bench.py's end-to-end leg, a training loop that
passes its dataset each time) keeps what was derived from it — the split / sliced operands of the tensor path and the
carried HMC evaluation; a different dataset replaces everything.  Labels are validated before any state changes."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from bayesian_inference_for_nn_b200 import _lib, keras_json  # noqa: E402
from bayesian_inference_for_nn_b200.engine import Engine  # noqa: E402


def _engine(D=784, H=256, C=10):
    return Engine(keras_json.parse_model_json(keras_json.make_sequential_json(D, [H, C], ["relu", "softmax"])), seed=3)


@pytest.mark.parametrize("shape", [(784, 256, 10, 640), (2, 50, 2, 300)])
def test_identical_resubmission_keeps_operands_and_carry(shape):
    D, H, C, N = shape
    rng = np.random.default_rng(1)
    X = rng.random((N, D)).astype(np.float32)
    y = rng.integers(0, C, N).astype(np.int32)
    runs = []
    for resubmit in (False, True):
        eng = _engine(D, H, C)
        eng.set_dataset(X, y, _lib.LOSS_SPARSE_CE)
        eng.set_prior([0.0], [1.0], _lib.PRIOR_SCALAR)
        eng.hmc_init(3, 1e-3, 1.0, 4, _lib.HMC_REFERENCE)
        eng.hmc_run(2, burning=True, sampling=False)
        evals = []
        for _ in range(3):
            if resubmit:
                eng.set_dataset(X.copy(), y.copy(), _lib.LOSS_SPARSE_CE)
            evals.append(eng.hmc_run(1, burning=False, sampling=True)["grad_evals"])
        runs.append((eng.hmc_state(), evals, int(eng.info("dataset_kept")), int(eng.info("dataset_uploads"))))
        eng.close()
    (s0, e0, k0, u0), (s1, e1, k1, u1) = runs
    assert (k0, u0) == (0, 1) and (k1, u1) == (3, 4)
    assert e0 == e1                                   # the carried evaluation survives an identical re-upload
    np.testing.assert_array_equal(s0[0], s1[0])       # and the chains are bit-identical
    np.testing.assert_array_equal(s0[1], s1[1])


def test_a_different_dataset_replaces_the_resident_one():
    D, H, C, N = 784, 256, 10, 384
    rng = np.random.default_rng(2)
    X = rng.random((N, D)).astype(np.float32)
    y = rng.integers(0, C, N).astype(np.int32)
    q = (rng.standard_normal((2, D * H + H + H * C + C)) * 0.05).astype(np.float32)
    eng = _engine()
    eng.set_dataset(X, y, _lib.LOSS_SPARSE_CE)
    eng.set_prior([0.0], [1.0], _lib.PRIOR_SCALAR)
    U0, _, g0 = eng.hmc_eval(q)
    X2 = X.copy()
    X2[17, 5] += 0.25                                 # one element differs
    eng.set_dataset(X2, y, _lib.LOSS_SPARSE_CE)
    assert int(eng.info("dataset_kept")) == 0
    U1, _, g1 = eng.hmc_eval(q)
    fresh = _engine()
    fresh.set_dataset(X2, y, _lib.LOSS_SPARSE_CE)
    fresh.set_prior([0.0], [1.0], _lib.PRIOR_SCALAR)
    U2, _, g2 = fresh.hmc_eval(q)
    np.testing.assert_array_equal(U1, U2)
    np.testing.assert_array_equal(g1, g2)
    assert np.any(g1 != g0)
    y2 = y.copy()
    y2[3] = (y2[3] + 1) % C                           # same X, one label differs
    eng.set_dataset(X2, y2, _lib.LOSS_SPARSE_CE)
    assert int(eng.info("dataset_kept")) == 0
    eng.set_dataset(X2, y2, _lib.LOSS_SPARSE_CE)
    assert int(eng.info("dataset_kept")) == 1


def test_bad_labels_are_rejected_before_any_state_changes():
    torch = pytest.importorskip("torch")
    D, H, C, N = 784, 256, 10, 256
    rng = np.random.default_rng(4)
    X = rng.random((N, D)).astype(np.float32)
    y = rng.integers(0, C, N).astype(np.int32)
    q = (rng.standard_normal((1, D * H + H + H * C + C)) * 0.05).astype(np.float32)
    eng = _engine()
    eng.set_dataset(X, y, _lib.LOSS_SPARSE_CE)
    eng.set_prior([0.0], [1.0], _lib.PRIOR_SCALAR)
    U0, _, g0 = eng.hmc_eval(q)
    bad = y.copy()
    bad[5] = C
    for Xb, yb in ((X[:128], bad[:128]), (torch.from_numpy(X[:128]).cuda(), torch.from_numpy(bad[:128]).cuda())):
        with pytest.raises(_lib.PyesianB200Error):
            eng.set_dataset(Xb, yb, _lib.LOSS_SPARSE_CE)
        U1, _, g1 = eng.hmc_eval(q)                   # the resident dataset (and its N) is untouched
        np.testing.assert_array_equal(U0, U1)
        np.testing.assert_array_equal(g0, g1)
