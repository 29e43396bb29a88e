"""GPU tests at BASELINE.json's FULL data size (60000 x 784, 784-256-10) through size-independent
properties, where the oracle would take minutes:
  * additivity: the summed negative log-likelihood and its gradient over the whole dataset equal the
    sums over two disjoint halves;
  * the tcgen05 path and the fp32 SIMT path agree on the same device;
  * chains are independent: evaluating a chain alone or inside a batch gives the same numbers;
  * a device-resident dataset handed over as a DLPack capsule (zero-copy) equals the host upload."""
import numpy as np
import pytest

from conftest import rel_err

pytestmark = pytest.mark.gpu

from bayesian_inference_for_nn_b200 import _lib, keras_json  # noqa: E402
from bayesian_inference_for_nn_b200.engine import Engine  # noqa: E402

D, H, C, N, S = 784, 256, 10, 60000, 4


@pytest.fixture(scope="module")
def data():
    rng = np.random.default_rng(0)
    X = rng.random((N, D), dtype=np.float32)
    y = rng.integers(0, C, N).astype(np.int32)
    P = D * H + H + H * C + C
    q = (rng.standard_normal((S, P)) * 0.03).astype(np.float32)
    return X, y, q


def make(X, y, path=_lib.PATH_AUTO, act="relu", tc_i8=None):
    eng = Engine(keras_json.parse_model_json(keras_json.make_sequential_json(D, [H, C], [act, "softmax"])))
    eng.set_option("path", path)
    if tc_i8 is not None:
        eng.set_option("tc_i8", tc_i8)
    eng.set_dataset(X, y, _lib.LOSS_SPARSE_CE)
    eng.set_prior([0.0], [1.0], _lib.PRIOR_SCALAR)
    return eng


def nll_and_grad(eng, q, n_rows):
    """sum of the per-row NLL and its gradient (prior removed): U - prior = n*loss."""
    U, loss, g = eng.hmc_eval(q)
    prior_g = q.astype(np.float64)                       # d/dq of 1/2 q^2 with N(0,1)
    return loss.astype(np.float64) * n_rows, g.astype(np.float64) - prior_g


@pytest.mark.parametrize("tc_i8", [0, 2])
def test_full_size_additivity_and_path_agreement(data, tc_i8):
    """tc_i8 = 0: bf16x3 products; 2: int8 slices.  The slices carry 16-bit FIXED point per W1 column, so the
    pre-activations differ by ~3e-5 relative between the whole dataset and its halves (different centring means) and
    ~2e-5 of the 15 M relu masks flip: the additivity bound for the gradient is the flip noise there (each flip moves
    this random-label gradient by ~1e-5 of its norm), the smooth part is held to 1e-4 by the kink-aware test below."""
    X, y, q = data
    full = make(X, y, tc_i8=tc_i8)
    nll, g = nll_and_grad(full, q, N)
    assert int(full.info("path_used")) == _lib.PATH_TENSOR
    cut = 29952                                           # not a multiple of the 128/256-row tiles on purpose
    a = make(X[:cut], y[:cut], tc_i8=tc_i8)
    b = make(X[cut:], y[cut:], tc_i8=tc_i8)
    nll_a, g_a = nll_and_grad(a, q, cut)
    nll_b, g_b = nll_and_grad(b, q, N - cut)
    np.testing.assert_allclose(nll_a + nll_b, nll, rtol=2e-5)
    print("additivity tc_i8=%d: %s" % (tc_i8, ["%.1e" % rel_err(g_a[s] + g_b[s], g[s]) for s in range(S)]))
    for s in range(S):
        assert rel_err(g_a[s] + g_b[s], g[s]) < (1e-4 if tc_i8 == 0 else 1e-3)
    # fp32 SIMT path on the same inputs.  Its pre-activations differ from the tensor path's in the last bits, so
    # among the 15 M (row, unit) pairs of a chain a few relu masks flip (|z1| ~ 1e-7); one flip moves this
    # random-label gradient by ~1/sqrt(N*H/2) = 3.6e-4 of its norm, hence the looser bound here.  The smooth-activation
    # test below checks the full-size arithmetic against float64 at the real tolerance.
    gen = make(X, y, _lib.PATH_GENERIC)
    nll_g, g_g = nll_and_grad(gen, q, N)
    np.testing.assert_allclose(nll_g, nll, rtol=2e-5)
    for s in range(S):
        assert rel_err(g_g[s], g[s]) < 2e-3
    # chain independence: chain 2 alone == chain 2 inside the batch
    U_all, _, g_all = full.hmc_eval(q)
    U_one, _, g_one = full.hmc_eval(q[2:3])
    np.testing.assert_allclose(U_one[0], U_all[2], rtol=1e-6)
    assert rel_err(g_one[0], g_all[2]) < 1e-6


def test_full_size_gradient_against_float64_oracle(data, oracle):
    """One chain, all 60000 rows, tanh hidden layer (no kinks): log-prob and gradient of both device paths against
    the float64 oracle at the parity tolerance.  This is the test that exposed the truncating fp32 accumulation
    of the tensor cores over 60000-long reductions (now split-K, see TcGemmParams)."""
    O = oracle
    X, y, q = data
    spec = O.MLPSpec(D, [H, C], ["tanh", "softmax"])
    loss64, g64 = O.mean_loss_and_grad(spec, q[:1], X, y, O.LOSS_SPARSE_CE, np.float64)
    for path in (_lib.PATH_TENSOR, _lib.PATH_GENERIC):
        eng = make(X, y, path, act="tanh")
        nll, g = nll_and_grad(eng, q[:1], N)
        assert int(eng.info("path_used")) == path
        assert abs(nll[0] - loss64[0] * N) < 2e-5 * abs(nll[0])
        assert rel_err(g[0], g64[0] * N) < 1e-4, (path, rel_err(g[0], g64[0] * N))


@pytest.mark.parametrize("tc_i8", [0, 2])
def test_full_size_relu_gradient_against_float64_oracle_with_the_device_mask(data, oracle, tc_i8):
    """The HEADLINE configuration itself (60000 x 784, 784-256-10, relu) against the float64 oracle at 1e-4.  relu' is
    discontinuous, so any two correct implementations disagree about the units whose pre-activation lies within their
    rounding of zero (float32 rounding for bf16x3, ~3e-5 relative for the int8 slices); the comparison is therefore
    kink-aware: the oracle evaluates loss and gradient with the relu mask the DEVICE used (pyb_debug_relu_mask), and the
    test bounds how far from zero the float64 pre-activations of the disagreeing units are."""
    O = oracle
    X, y, q = data
    spec = O.MLPSpec(D, [H, C], ["relu", "softmax"])
    eng = make(X, y, _lib.PATH_TENSOR, tc_i8=tc_i8)
    nll, g = nll_and_grad(eng, q[:1], N)
    mask = np.empty((N, H), np.uint8)
    _lib.check(_lib.load().pyb_debug_relu_mask(eng.h, 0, mask.ctypes.data))
    z1 = X.astype(np.float64) @ q[0, :D * H].astype(np.float64).reshape(D, H) + q[0, D * H:D * H + H].astype(np.float64)
    flips = (z1 > 0) != mask.astype(bool)
    print("tc_i8=%d: %d of %d relu masks differ from float64, max |z1| among them %.2e (rms z1 %.2e)"
          % (tc_i8, int(flips.sum()), N * H, float(np.abs(z1[flips]).max()) if flips.any() else 0.0, float(np.sqrt((z1 ** 2).mean()))))
    assert flips.mean() < (1e-5 if tc_i8 == 0 else 2e-4)
    if flips.any():
        assert np.abs(z1[flips]).max() < (3e-5 if tc_i8 == 0 else 4e-4) * np.sqrt((z1 ** 2).mean())
    loss64, g64 = O.mean_loss_and_grad(spec, q[:1], X, y, O.LOSS_SPARSE_CE, np.float64, relu_masks={0: mask.astype(bool)[None]})
    err = rel_err(g[0], g64[0] * N)
    print("tc_i8=%d: full-size relu gradient vs float64 oracle with the device mask: %.2e, nll %.2e"
          % (tc_i8, err, abs(nll[0] - loss64[0] * N) / abs(nll[0])))
    assert abs(nll[0] - loss64[0] * N) < 2e-5 * abs(nll[0])
    assert err < 1e-4


def test_full_size_hmc_energy_bookkeeping(data):
    """One full-size iteration for a handful of chains: the Hamiltonian bookkeeping is self-consistent
    (U0 of iteration k+1 equals U1 of an accepted iteration k; K1 equals the kinetic energy of the returned
    momentum) and the textbook integrator nearly conserves energy at a small step size."""
    X, y, q = data
    eng = make(X, y)
    eng.hmc_init(S, 2e-5, 1.0, 5, _lib.HMC_CANONICAL, q0=q)
    eng.hmc_run(1, burning=True, sampling=False)
    first = eng.hmc_last()
    _, p = eng.hmc_state()
    np.testing.assert_allclose(first["K1"], (p.astype(np.float64) ** 2).sum(1) / 2, rtol=1e-5)
    assert np.all(np.abs(first["log_alpha"]) < 1.0)
    eng.hmc_run(1, burning=True, sampling=False)
    second = eng.hmc_last()
    np.testing.assert_allclose(second["U0"], first["U1"], rtol=1e-6)


def test_device_resident_dataset_through_dlpack(data):
    torch = pytest.importorskip("torch")
    X, y, q = data
    n = 4096
    host = make(X[:n], y[:n])
    U_h, l_h, g_h = host.hmc_eval(q[:2])
    Xd = torch.from_numpy(X[:n]).cuda()
    yd = torch.from_numpy(y[:n]).cuda()
    dev = Engine(keras_json.parse_model_json(keras_json.make_sequential_json(D, [H, C], ["relu", "softmax"])))
    dev.set_dataset(Xd, yd, _lib.LOSS_SPARSE_CE)          # __dlpack__ -> device pointers, no host staging
    dev.set_prior([0.0], [1.0], _lib.PRIOR_SCALAR)
    U_d, l_d, g_d = dev.hmc_eval(q[:2])
    np.testing.assert_array_equal(U_h, U_d)
    np.testing.assert_array_equal(g_h, g_d)
