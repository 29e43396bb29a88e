"""CPU tests of the oracle itself: it must reproduce the frozen fixtures, agree with an independent
autograd implementation (torch, test-only), and encode the reference's quirks."""
import math

import numpy as np
import pytest

from conftest import load_golden, rel_err, spec_from_golden

torch = pytest.importorskip("torch")


def test_philox_known_answers(oracle):
    # Random123 kat_vectors, philox4x32-10
    kat = [([0, 0, 0, 0], [0, 0], [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
           ([0xffffffff] * 4, [0xffffffff] * 2, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
           ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0],
            [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1])]
    for c, k, want in kat:
        got = oracle.philox4x32_10(np.array(c, np.uint32), np.array(k, np.uint32))
        assert [int(v) for v in got] == want


def test_philox_streams_match_fixture(oracle):
    g = load_golden("philox")
    z = oracle.philox_normals(int(g["seed"]), g["chains"], int(g["iteration"]), oracle.STREAM_MOMENTUM, 37)
    np.testing.assert_array_equal(z, g["normals"])
    np.testing.assert_array_equal(oracle.philox_uniforms(int(g["seed"]), g["chains"], int(g["iteration"])), g["uniforms"])
    big = oracle.philox_normals(1, np.arange(8), 0, 0, 40000)
    assert abs(big.mean()) < 0.01 and abs(big.std() - 1) < 0.01


def _torch_loss(spec, theta, X, y, loss_kind, O):
    t = torch.tensor(theta, dtype=torch.float64, requires_grad=True)
    a = torch.tensor(X, dtype=torch.float64)[None].expand(theta.shape[0], -1, -1)
    offs, _ = spec.offsets()
    for l, (w, b, fi, fo) in enumerate(offs):
        z = a @ t[:, w:w + fi * fo].reshape(-1, fi, fo)
        if b >= 0:
            z = z + t[:, None, b:b + fo]
        act = spec.acts[l]
        a = {O.ACT_LINEAR: lambda v: v, O.ACT_SOFTMAX: lambda v: v, O.ACT_RELU: torch.relu, O.ACT_TANH: torch.tanh,
             O.ACT_SIGMOID: torch.sigmoid}[act](z)
    if loss_kind == O.LOSS_SPARSE_CE:
        yt = torch.tensor(np.asarray(y).reshape(-1), dtype=torch.long)
        loss = torch.stack([torch.nn.functional.cross_entropy(a[s], yt) for s in range(a.shape[0])])
    else:
        yt = torch.tensor(np.asarray(y, np.float64)).reshape(X.shape[0], -1)
        loss = ((a - yt[None]) ** 2).mean(dim=2).mean(dim=1)
    loss.sum().backward()
    return loss.detach().numpy(), t.grad.numpy()


@pytest.mark.parametrize("name", ["hmc_c1_mini", "hmc_c3_mini", "hmc_regression", "hmc_deep_mse"])
def test_gradients_match_autograd(oracle, name):
    O = oracle
    g = load_golden(name)
    spec = spec_from_golden(g, O)
    loss, grad = O.mean_loss_and_grad(spec, g["q"], g["X"], g["y"], int(g["loss_kind"]), np.float64)
    tl, tg = _torch_loss(spec, g["q"], g["X"], g["y"], int(g["loss_kind"]), O)
    np.testing.assert_allclose(loss, tl, rtol=1e-10)
    assert rel_err(grad, tg) < 1e-10


@pytest.mark.parametrize("name", ["hmc_c1_mini", "hmc_c1_canonical", "hmc_c3_mini", "hmc_regression", "hmc_deep_mse"])
def test_hmc_fixture_reproduces(oracle, name):
    O = oracle
    g = load_golden(name)
    spec = spec_from_golden(g, O)
    mu, sg = O.expand_prior(spec, 0.0, float(g["sigma"]))
    prob = O.Problem(spec, g["X"], g["y"], int(g["loss_kind"]), mu, sg)
    r64 = O.hmc_iteration(prob, g["q"], g["p"], g["u"], float(g["eps"]), float(g["m"]), int(g["L"]), False,
                          int(g["semantics"]), np.float64)
    np.testing.assert_allclose(r64["log_alpha"], g["log_alpha"], rtol=1e-9, atol=1e-9)
    np.testing.assert_array_equal(r64["accept"], g["accept"])
    assert rel_err(r64["qL"], g["qL"]) < 1e-6
    # float32 (what the reference computes in) stays within the parity budget of the float64 fixture
    r32 = O.hmc_iteration(prob, g["q"], g["p"], g["u"], float(g["eps"]), float(g["m"]), int(g["L"]), False,
                          int(g["semantics"]), np.float32)
    assert rel_err(r32["qL"], g["qL"]) < 1e-4 and rel_err(r32["pL"], g["pL"]) < 1e-3
    assert np.all(np.abs(r32["U0"] - g["U0"]) <= 1e-4 * np.abs(g["U0"]))


def _tiny_problem(O, sigma=1.0, seed=0, N=64):
    rng = np.random.default_rng(seed)
    spec = O.MLPSpec(2, [8, 2], ["relu", "softmax"])
    X = rng.standard_normal((N, 2)).astype(np.float32)
    y = (X[:, 0] * X[:, 1] > 0).astype(np.int32)
    mu, sg = O.expand_prior(spec, 0.0, sigma)
    return spec, O.Problem(spec, X, y, O.LOSS_SPARSE_CE, mu, sg), rng


def test_reference_leapfrog_has_L_plus_one_kicks(oracle):
    """HMC.py:83-87: L full kicks + two half kicks; canonical has L-1 full kicks."""
    O = oracle
    spec, prob, rng = _tiny_problem(O)
    q = rng.standard_normal((1, spec.n_params)).astype(np.float32) * 0.1
    p = rng.standard_normal((1, spec.n_params)).astype(np.float32)
    eps, m, L = 1e-3, 1.0, 1
    ref = O.hmc_iteration(prob, q, p, [0.5], eps, m, L, True, O.HMC_REFERENCE, np.float64)
    can = O.hmc_iteration(prob, q, p, [0.5], eps, m, L, True, O.HMC_CANONICAL, np.float64)
    _, _, g0 = O.potential(prob, q, np.float64)
    q1 = q + eps / m * (p - eps / 2 * g0)
    _, _, g1 = O.potential(prob, q1, np.float64)
    np.testing.assert_allclose(ref["pL"], p - eps / 2 * g0 - 1.5 * eps * g1, rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(can["pL"], p - eps / 2 * g0 - 0.5 * eps * g1, rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(ref["qL"], q1, rtol=1e-12)


def test_negative_sigma_gives_nan_hamiltonian_and_rejects(oracle):
    """SURVEY B-1: GaussianPrior(0,-1) => log(scale) NaN => every post-burn-in proposal rejected,
    while burn-in (always accept) still moves."""
    O = oracle
    spec, prob, rng = _tiny_problem(O, sigma=-1.0)
    q = np.zeros((2, spec.n_params), np.float32)
    p = rng.standard_normal((2, spec.n_params)).astype(np.float32)
    r = O.hmc_iteration(prob, q, p, [0.0, 0.0], 1e-3, 1.0, 3, False)
    assert np.all(np.isnan(r["log_alpha"])) and not r["accept"].any()
    np.testing.assert_array_equal(r["q"], q)
    rb = O.hmc_iteration(prob, q, p, [0.9, 0.9], 1e-3, 1.0, 3, True)
    assert rb["accept"].all() and np.abs(rb["q"] - q).max() > 0


def test_sample_bookkeeping(oracle):
    """HMC.py:75-77, 92-104: first sampling call seeds [q], accept appends, reject bumps the last."""
    O = oracle
    book = O.SampleBook(2)
    q0 = np.array([[0.0], [10.0]])
    book.record(q0, np.array([[1.0], [10.0]]), np.array([True, False]))
    book.record(np.array([[1.0], [10.0]]), np.array([[1.0], [11.0]]), np.array([False, True]))
    assert [float(s[0]) for s in book.samples[0]] == [0.0, 1.0] and book.freq[0] == [1, 2]
    assert [float(s[0]) for s in book.samples[1]] == [10.0, 11.0] and book.freq[1] == [2, 1]
    assert O.sampled_draw_index([1, 3, 4], 1) == 0 and O.sampled_draw_index([1, 3, 4], 2) == 1
    assert O.sampled_draw_index([1, 3, 4], 3) == 1 and O.sampled_draw_index([1, 3, 4], 4) == 2


def test_median_kernel_matches_scipy(oracle):
    sp = pytest.importorskip("scipy.spatial.distance")
    rng = np.random.default_rng(3)
    X = rng.standard_normal((9, 20))
    K, dxkxy, h = oracle.median_kernel(X)
    d2 = sp.squareform(sp.pdist(X)) ** 2
    h_ref = math.sqrt(0.5 * np.median(d2) / math.log(X.shape[0] + 1))
    assert abs(h - h_ref) < 1e-12
    K_ref = np.exp(-d2 / h_ref ** 2 / 2)
    np.testing.assert_allclose(K, K_ref, rtol=1e-10)
    ref = -K_ref @ X
    s = K_ref.sum(axis=1)
    for i in range(X.shape[1]):
        ref[:, i] = ref[:, i] + X[:, i] * s
    np.testing.assert_allclose(dxkxy, ref / h_ref ** 2, rtol=1e-9, atol=1e-12)


def test_live_kernel_gradient_matches_autograd(oracle):
    """SVGD._svgd_gradients :54-68: grad_kernel = -1/2 d(sum K)/dX = 2 sum_k K_ik (x_i-x_k)."""
    rng = np.random.default_rng(4)
    X = torch.tensor(rng.standard_normal((5, 7)) * 0.3, dtype=torch.float64, requires_grad=True)
    diff = X[:, None, :] - X[None, :, :]
    K = torch.exp(-(diff ** 2).sum(-1))
    K.sum().backward()
    want = (-X.grad / 2).numpy()
    Xn = X.detach().numpy()
    Kn = K.detach().numpy()
    got = np.stack([2 * (Kn[i][:, None] * (Xn[i][None] - Xn)).sum(0) for i in range(5)])
    np.testing.assert_allclose(got, want, rtol=1e-10, atol=1e-12)


def test_svgd_fixture_reproduces(oracle):
    O = oracle
    g = load_golden("svgd_mini")
    spec = O.MLPSpec(2, [50, 2], ["relu", "softmax"])
    P = spec.n_params
    am, av = np.zeros((6, P), np.float32), np.zeros((6, P), np.float32)
    parts = g["particles0"].copy()
    for t, ix in enumerate(g["idx"], 1):
        parts, am, av, loss, phi = O.svgd_live_step(spec, parts, g["X"][ix], g["y"][ix], O.LOSS_SPARSE_CE, am, av, t,
                                                    float(g["lr"]))
    np.testing.assert_allclose(parts, g["live_particles"], rtol=1e-6, atol=1e-7)
    phi, h, K = O.svgd_phi_canonical(g["particles0"], g["G"])
    np.testing.assert_allclose(phi, g["phi_hook"], rtol=1e-5, atol=1e-7)
    assert abs(h - float(g["h_hook"])) < 1e-9


def test_adam_legacy_first_step_is_lr_sign(oracle):
    th, m, v = oracle.adam_legacy(np.zeros(3, np.float32), np.float32([1, -2, 0.5]), np.zeros(3, np.float32),
                                  np.zeros(3, np.float32), 1, 0.01)
    np.testing.assert_allclose(th, [-0.01, 0.01, -0.01], rtol=1e-5)


def test_predictive_fixture_and_mask(oracle):
    O = oracle
    g = load_golden("predict_mini")
    spec = O.MLPSpec(2, [50, 2], ["relu", "softmax"])
    mean, var = O.predictive(spec, g["W"], g["x"], g["freq"], np.float32)
    np.testing.assert_allclose(mean, g["mean"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(var, g["var"], rtol=1e-3, atol=1e-6)
    # weighted == unweighted on the frequency-expanded sample list (Sampled semantics)
    Wrep = np.repeat(g["W"], g["freq"], axis=0)
    mean2, var2 = O.predictive(spec, Wrep, g["x"], None, np.float64)
    np.testing.assert_allclose(mean2, g["mean"], rtol=1e-9)
    np.testing.assert_allclose(var2, g["var"], rtol=1e-7, atol=1e-12)
    np.testing.assert_array_equal(O.uncertainty_mask(g["mean"], 0.7), g["mask"])
