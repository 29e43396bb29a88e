"""GPU tests of the tcgen05/TMA path (bf16x3 split products): the raw GEMM kernel against float64
NumPy, then log-prob/gradient/trajectory parity of the MNIST-width HMC path against the oracle and
against the generic fp32 SIMT path on the same inputs."""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import rel_err

pytestmark = pytest.mark.gpu

from bayesian_inference_for_nn_b200 import _lib, keras_json  # noqa: E402
from bayesian_inference_for_nn_b200.engine import Engine  # noqa: E402


def engine(D, H, Cc, act="relu", out_act="softmax", seed=0):
    js = keras_json.make_sequential_json(D, [H, Cc], [act, out_act])
    return Engine(keras_json.parse_model_json(js), seed=seed)


def debug_gemm(eng, A, B):
    fn = _lib.load().pyb_debug_tc_gemm
    A = np.ascontiguousarray(A, np.float32)
    B = np.ascontiguousarray(B, np.float32)
    D = np.empty((A.shape[0], B.shape[0]), np.float32)
    _lib.check(fn(eng.h, A.ctypes.data, B.ctypes.data, A.shape[0], B.shape[0], A.shape[1], D.ctypes.data))
    return D


@pytest.mark.parametrize("pair", [0, 1])
@pytest.mark.parametrize("M,Nn,K", [(128, 256, 32), (128, 256, 784), (300, 64, 96), (785, 256, 1000), (1000, 128, 40),
                                     (2500, 16, 2048), (4096, 256, 800)])
def test_split_bf16_gemm_matches_float64(M, Nn, K, pair):
    """pair=1: the CTA-pair (cta_group::2, double-buffered TMEM) kernel where it applies."""
    rng = np.random.default_rng(M + Nn + K)
    A = rng.standard_normal((M, K)).astype(np.float32)
    B = (rng.standard_normal((Nn, K)) * 0.3).astype(np.float32)
    eng = engine(64, 32, 4)
    eng.set_option("tc_pair", pair)
    D = debug_gemm(eng, A, B)
    want = A.astype(np.float64) @ B.astype(np.float64).T
    scale = np.sqrt((A.astype(np.float64) ** 2).sum(1))[:, None] * np.sqrt((B.astype(np.float64) ** 2).sum(1))[None]
    err = np.abs(D - want) / scale
    assert np.isfinite(D).all()
    assert err.max() < 3e-5, err.max()        # ~2^-16 per product, error relative to |a||b|
    assert rel_err(D, want) < 2e-5


def problem(oracle, D, H, Cc, N, S, seed, act="relu", loss="ce", q_scale=0.05):
    O = oracle
    rng = np.random.default_rng(seed)
    out_act = "softmax" if loss == "ce" else "linear"
    spec = O.MLPSpec(D, [H, Cc], [act, out_act])
    X = rng.random((N, D)).astype(np.float32)
    if loss == "ce":
        y, kind = rng.integers(0, Cc, N).astype(np.int32), O.LOSS_SPARSE_CE
    else:
        y, kind = rng.standard_normal((N, Cc)).astype(np.float32), O.LOSS_MSE
    q = (rng.standard_normal((S, spec.n_params)) * q_scale).astype(np.float32)
    if act == "relu":
        q = move_off_relu_kinks(q, X, D, H)
    mu, sg = O.expand_prior(spec, 0.0, 1.0)
    return spec, O.Problem(spec, X, y, kind, mu, sg), q, out_act, rng


def move_off_relu_kinks(q, X, D, H, margin=1e-3):
    """relu'(z) = 1[z>0] is discontinuous: a pre-activation within rounding distance of 0 flips the
    mask between ANY two fp32 implementations (and a single flipped unit moves a random-label
    gradient by ~1e-3 relative), so the 1e-4 gradient tolerance is only meaningful away from the
    kinks.  Shift each hidden unit's bias into the middle of a gap of its pre-activations so that
    min |z1| >= margin for every (row, unit)."""
    q = q.copy()
    X64 = X.astype(np.float64)
    for s in range(q.shape[0]):
        W = q[s, :D * H].astype(np.float64).reshape(D, H)
        b = q[s, D * H:D * H + H].astype(np.float64)
        Z = np.sort(X64 @ W + b, axis=0)                       # [N, H]
        for h in range(H):
            z = Z[:, h]
            edges = np.concatenate([[z[0] - 1.0], z, [z[-1] + 1.0]])
            gaps = edges[1:] - edges[:-1]
            mids = 0.5 * (edges[1:] + edges[:-1])
            ok = np.where(gaps > 4 * margin)[0]
            m = mids[ok[np.argmin(np.abs(mids[ok]))]]          # admissible gap centre closest to zero
            b[h] -= m
        q[s, D * H:D * H + H] = b.astype(np.float32)
        Zc = X64 @ W + q[s, D * H:D * H + H].astype(np.float64)
        assert np.abs(Zc).min() > 0.5 * margin
    return q


CASES = [(784, 256, 10, 512, 3, "relu", "ce"), (784, 128, 10, 300, 2, "relu", "ce"), (64, 32, 4, 1000, 5, "tanh", "ce"),
         (128, 64, 3, 257, 2, "sigmoid", "mse"), (784, 256, 10, 128, 1, "relu", "ce"),
         # the fused layer-1 GEMM + layer-2 epilogue (relu, H = 128/256) in every class padding CP = 4, 8, 12, 16,
         # both losses, ragged row counts (last tile partly / wholly zero fill) and an odd chain count
         (64, 128, 2, 129, 3, "relu", "ce"), (96, 256, 7, 1000, 2, "relu", "ce"), (64, 128, 16, 640, 2, "relu", "ce"),
         (128, 256, 3, 385, 5, "relu", "mse"), (784, 256, 12, 257, 1, "relu", "ce"),
         # 128 hidden units: dW1 on the dual hidden-major kernel with two CHAINS per CTA pair; odd chain count (the last
         # pair has one partner) and more than 8192 rows (split-K partial sums per chain)
         (784, 128, 10, 9000, 3, "relu", "ce"), (320, 128, 4, 640, 5, "relu", "mse")]


@pytest.mark.parametrize("D,H,Cc,N,S,act,loss", CASES)
def test_tensor_path_logprob_and_gradient(oracle, D, H, Cc, N, S, act, loss):
    O = oracle
    spec, prob, q, out_act, _ = problem(O, D, H, Cc, N, S, seed=D + H + N, act=act, loss=loss)
    U64, loss64, g64 = O.potential(prob, q, np.float64)
    eng = engine(D, H, Cc, act, out_act)
    eng.set_dataset(prob.X, prob.y, prob.loss_kind)
    eng.set_prior([0.0], [1.0], _lib.PRIOR_SCALAR)
    eng.set_option("path", _lib.PATH_TENSOR)
    U, ls, g = eng.hmc_eval(q)
    assert int(eng.info("path_used")) == _lib.PATH_TENSOR
    np.testing.assert_allclose(U, U64, rtol=1e-4)
    np.testing.assert_allclose(ls, loss64, rtol=1e-4)
    for s in range(S):
        assert rel_err(g[s], g64[s]) < 1e-4, (s, rel_err(g[s], g64[s]))
    # and against the fp32 SIMT path on the same device
    eng.set_option("path", _lib.PATH_GENERIC)
    Ug, lg, gg = eng.hmc_eval(q)
    np.testing.assert_allclose(U, Ug, rtol=1e-4)
    for s in range(S):
        assert rel_err(g[s], gg[s]) < 1e-4
    # chain batching inside the tensor path must not change anything
    eng.set_option("path", _lib.PATH_TENSOR)
    eng.set_option("chain_batch", 1)
    eng.set_dataset(prob.X, prob.y, prob.loss_kind)     # re-upload also re-derives the split operands
    U1, _, g1 = eng.hmc_eval(q)
    np.testing.assert_array_equal(U, U1)
    np.testing.assert_array_equal(g, g1)


@pytest.mark.parametrize("H,Cc,N,S", [(256, 10, 700, 4), (128, 5, 300, 3)])
def test_fused_epilogue_agrees_with_the_unfused_kernels(oracle, H, Cc, N, S):
    """tc_fuse=0 runs G1, k_layer2 and the dW2 GEMM as separate kernels, tc_dual=0 the single-accumulator dW1 GEMM:
    same arithmetic in a different summation order, so the results agree far inside the parity budget."""
    O = oracle
    spec, prob, q, out_act, _ = problem(O, 784, H, Cc, N, S, seed=H + N)
    eng = engine(784, H, Cc, "relu", out_act)
    eng.set_option("tc_i8", 0)                       # this is an A/B of the bf16x3 kernels (the int8 slices: test_gpu_i8.py)
    eng.set_dataset(prob.X, prob.y, prob.loss_kind)
    eng.set_prior([0.0], [1.0], _lib.PRIOR_SCALAR)
    eng.set_option("path", _lib.PATH_TENSOR)
    U, ls, g = eng.hmc_eval(q)
    for key in ("tc_fuse", "tc_dual"):
        eng.set_option(key, 0)
        eng.set_dataset(prob.X, prob.y, prob.loss_kind)
        U0, ls0, g0 = eng.hmc_eval(q)
        eng.set_option(key, 1)
        np.testing.assert_allclose(U0, U, rtol=2e-6)
        np.testing.assert_allclose(ls0, ls, rtol=2e-6)
        for s in range(S):
            assert rel_err(g0[s], g[s]) < 1e-5, (key, s, rel_err(g0[s], g[s]))
    # repeatable bit for bit
    eng.set_dataset(prob.X, prob.y, prob.loss_kind)
    U2, _, g2 = eng.hmc_eval(q)
    np.testing.assert_array_equal(U, U2)
    np.testing.assert_array_equal(g, g2)


def test_tensor_path_hmc_iteration_matches_oracle(oracle):
    O = oracle
    D, H, Cc, N, S, L, eps = 784, 256, 10, 640, 3, 4, 1e-3
    spec, prob, q, out_act, rng = problem(O, D, H, Cc, N, S, seed=11)
    p = rng.standard_normal((S, spec.n_params)).astype(np.float32)
    u = np.float32([0.0, 0.999, 0.5])
    want = O.hmc_iteration(prob, q, p, u, eps, 1.0, L, False, O.HMC_REFERENCE, np.float64)
    eng = engine(D, H, Cc)
    eng.set_dataset(prob.X, prob.y, prob.loss_kind)
    eng.set_prior([0.0], [1.0], _lib.PRIOR_SCALAR)
    eng.hmc_init(S, eps, 1.0, L, _lib.HMC_REFERENCE, q0=q)        # path AUTO must pick the tensor path
    eng.hmc_inject(p=p, u=u)
    d = eng.hmc_run(1, burning=False, sampling=True)
    assert int(eng.info("path_used")) == _lib.PATH_TENSOR
    last = eng.hmc_last()
    for k in ("U0", "K0", "U1", "K1"):
        np.testing.assert_allclose(last[k], want[k], rtol=1e-4, err_msg=k)
    lu = np.log(np.maximum(u.astype(np.float64), 1e-300))
    decisive = np.abs(want["log_alpha"] - lu) > 1e-5 * np.maximum(1.0, np.abs(want["U0"]))
    np.testing.assert_array_equal(last["accept"][decisive], want["accept"][decisive])
    qd, pd = eng.hmc_state()
    for s in range(S):
        if last["accept"][s] == want["accept"][s]:
            assert rel_err(qd[s], want["q"][s]) < 1e-3
            assert rel_err(pd[s], want["pL"][s]) < 1e-3
    assert d["grad_evals"] == S * (L + 1)


def test_h128_chain_pair_dw1_agrees_with_the_feature_major_kernel(oracle):
    """A/B of the two dW1 kernels for 128 hidden units ("tc_h128_pairs"): same products, same k order"""
    spec, prob, q, out_act, _ = problem(oracle, 784, 128, 10, 1024, 7, seed=3, act="relu", loss="ce")
    eng = engine(784, 128, 10, "relu", out_act)
    eng.set_dataset(prob.X, prob.y, prob.loss_kind)
    eng.set_prior([0.0], [1.0], _lib.PRIOR_SCALAR)
    eng.set_option("path", _lib.PATH_TENSOR)
    _, _, g_pairs = eng.hmc_eval(q)
    eng.set_option("tc_h128_pairs", 0)
    _, _, g_fm = eng.hmc_eval(q)
    for s in range(7):
        assert rel_err(g_pairs[s], g_fm[s]) < 2e-6


def test_tensor_path_reversibility_canonical(oracle):
    """Size-independent property: the textbook leapfrog is time-reversible.  Integrate L steps,
    negate the momentum, integrate L more: the chain returns to its start (to fp32 rounding)."""
    O = oracle
    D, H, Cc, N, S, L, eps = 784, 256, 10, 2048, 4, 5, 5e-4
    spec, prob, q, out_act, rng = problem(O, D, H, Cc, N, S, seed=3)
    p = rng.standard_normal((S, spec.n_params)).astype(np.float32)
    eng = engine(D, H, Cc)
    eng.set_dataset(prob.X, prob.y, prob.loss_kind)
    eng.set_prior([0.0], [1.0], _lib.PRIOR_SCALAR)
    eng.hmc_init(S, eps, 1.0, L, _lib.HMC_CANONICAL, q0=q)
    eng.hmc_inject(p=p)
    eng.hmc_run(1, burning=True, sampling=False)
    q1, p1 = eng.hmc_state()
    e0 = eng.hmc_last()
    assert np.all(np.abs(e0["log_alpha"]) < 0.5)                    # energy nearly conserved
    eng.hmc_inject(p=-p1)
    eng.hmc_run(1, burning=True, sampling=False)
    q2, p2 = eng.hmc_state()
    assert rel_err(q2, q) < 1e-5 and rel_err(-p2, p) < 1e-4
    assert np.abs(q1 - q).max() > 0


def test_predictive_on_tensor_path(oracle):
    """BayesianModel.predict shape (config C5 in small): tensor-core forward for every weight sample."""
    O = oracle
    D, H, Cc, Nt, n = 784, 256, 10, 300, 5
    rng = np.random.default_rng(21)
    spec = O.MLPSpec(D, [H, Cc], ["relu", "softmax"])
    W = (rng.standard_normal((n, spec.n_params)) * 0.05).astype(np.float32)
    x = rng.random((Nt, D)).astype(np.float32)
    freq = np.float32([1, 2, 1, 3, 1])
    mean64, var64 = O.predictive(spec, W, x, freq, np.float64)
    eng = engine(D, H, Cc)
    mean, var, allo = eng.predict(W, x, weights=freq, want_all=True)
    assert int(eng.info("path_used")) == _lib.PATH_TENSOR
    np.testing.assert_allclose(mean, mean64, rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(var, var64, rtol=2e-3, atol=1e-7)
    np.testing.assert_allclose(allo, O.forward(spec, W, x, np.float64), rtol=1e-4, atol=1e-6)
    eng.set_option("path", _lib.PATH_GENERIC)
    mean_g, _, _ = eng.predict(W, x, weights=freq)
    assert int(eng.info("path_used")) == _lib.PATH_GENERIC
    np.testing.assert_allclose(mean, mean_g, rtol=1e-4, atol=1e-6)
    # weight samples and inputs already resident in HBM (DLPack): read in place, same numbers bit for bit;
    # the unfused kernels (tc_fuse=0: G1 + k_layer2_fwd through A1^T) agree with the fused forward epilogue
    torch = pytest.importorskip("torch")
    eng.set_option("path", _lib.PATH_TENSOR)
    mean_d, var_d, all_d = eng.predict(torch.from_numpy(W).cuda(), torch.from_numpy(x).cuda(), weights=freq, want_all=True)
    np.testing.assert_array_equal(mean_d, mean)
    np.testing.assert_array_equal(all_d, allo)
    eng.set_option("tc_fuse", 0)
    mean_u, _, all_u = eng.predict(W, x, weights=freq, want_all=True)
    np.testing.assert_allclose(all_u, allo, rtol=2e-5, atol=1e-7)      # the unfused path carries a1 as bf16 hi + lo (2^-17)
    np.testing.assert_allclose(mean_u, mean, rtol=2e-5, atol=1e-7)


@pytest.mark.parametrize("sem", [_lib.SVGD_CANONICAL_MEDIAN, _lib.SVGD_REFERENCE_LIVE])
def test_svgd_minibatch_gradients_on_tensor_path(oracle, sem):
    """SVGD_mnist shape in small (784-128-10, minibatch from a resident pool): the per-particle gradients
    come from the tcgen05 path on a re-split minibatch; the Stein update is checked against the oracle."""
    O = oracle
    D, H, Cc, N, S, B = 784, 128, 10, 600, 4, 256
    spec, prob, q, out_act, rng = problem(O, D, H, Cc, N, S, seed=5)
    idx = [rng.permutation(N)[:B].astype(np.int32) for _ in range(2)]
    parts = q.astype(np.float64)
    am, av = np.zeros((S, spec.n_params), np.float32), np.zeros((S, spec.n_params), np.float32)
    want = parts.copy()
    losses = []
    for t, ix in enumerate(idx, 1):
        if sem == _lib.SVGD_CANONICAL_MEDIAN:
            want, am, av, loss, _, _ = O.svgd_canonical_step(prob, want, prob.X[ix], prob.y[ix], am, av, t, 1e-3)
        else:
            want, am, av, loss, _ = O.svgd_live_step(spec, want, prob.X[ix], prob.y[ix], prob.loss_kind, am, av, t, 1e-3)
        losses.append(loss)
    eng = engine(D, H, Cc)
    eng.set_dataset(prob.X, prob.y, prob.loss_kind)
    eng.set_prior([0.0], [1.0], _lib.PRIOR_SCALAR)
    eng.svgd_init(S, 1e-3, sem, particles0=parts)
    got_losses = [eng.svgd_step(ix) for ix in idx]
    assert int(eng.info("path_used")) == _lib.PATH_TENSOR
    np.testing.assert_allclose(got_losses, losses, rtol=1e-4)
    # Adam's first steps are lr * phi / (|phi| + 1e-7): for the handful of coordinates with |phi| ~ 1e-7 (expected
    # among 4e5) the step is ill-conditioned in phi, so bound the bulk tightly and the worst case by Adam's own bound
    diff = np.abs(eng.svgd_particles() - want)
    assert np.quantile(diff, 0.9995) < 2e-3 * 1e-3 * 2 + 1e-6
    assert diff.max() <= 2.1 * 1e-3 * len(idx)


@pytest.mark.parametrize("S", [256, 384, 1024])
def test_svgd_canonical_phi_on_tensor_cores(oracle, S):
    """Large particle sets: Gram matrix (bf16x3 GEMM; S >= 512: upper tile triangle + mirror) -> exact median bandwidth
    -> K Y contraction (bf16x3 GEMM) against the float64 oracle of SVGD.baseline__kernel."""
    rng = np.random.default_rng(S)
    eng = engine(64, 32, 4)                      # P = 2212
    P = eng.P
    X = (rng.standard_normal((S, P)) * 0.3).astype(np.float32).astype(np.float64)
    G = rng.standard_normal((S, P)).astype(np.float32)
    want, h_ref, _ = oracle.svgd_phi_canonical(X, G)
    phi, h = eng.svgd_phi(X, G, _lib.SVGD_CANONICAL_MEDIAN)
    assert abs(h - h_ref) < 1e-5 * h_ref
    assert rel_err(phi, want) < 1e-4
    eng.set_option("path", _lib.PATH_GENERIC)    # float64 SIMT kernels on the same inputs
    phi_g, h_g = eng.svgd_phi(X, G, _lib.SVGD_CANONICAL_MEDIAN)
    assert abs(h_g - h_ref) < 1e-9 * h_ref and rel_err(phi_g, want) < 1e-5
