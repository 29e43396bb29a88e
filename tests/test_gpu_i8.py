"""GPU tests of the int8-slice operand scheme (csrc/tc_i8.cuh; `tc_i8` option): the forward GEMM (mode 1) and the dW1 GEMM
(mode 2) on `tcgen05.mma kind::i8` with exact int32 accumulation, against the float64 oracle at the 1e-4 parity budget
and against the bf16x3 kernels on the same inputs."""
import numpy as np
import pytest

from conftest import rel_err
from test_gpu_tensor import engine, problem

pytestmark = pytest.mark.gpu

from bayesian_inference_for_nn_b200 import _lib  # noqa: E402

# (D, H, C, N, S, loss): relu hidden layer of 128 / 256 units (the fused kernel); mode 2 needs H = 256 and the
# cross-entropy (else the dW1 GEMM stays on bf16x3, which the test then exercises together with the int8 forward GEMM)
CASES = [(784, 256, 10, 512, 3, "ce"), (784, 256, 10, 128, 1, "ce"), (96, 256, 7, 1000, 2, "ce"), (784, 256, 12, 257, 1, "ce"),
         (784, 128, 10, 300, 2, "ce"), (64, 128, 2, 129, 3, "ce"), (128, 256, 3, 385, 5, "mse"), (784, 256, 10, 9000, 2, "ce"),
         (200, 256, 4, 640, 150, "ce")]


@pytest.mark.parametrize("mode", [1, 2])
@pytest.mark.parametrize("D,H,Cc,N,S,loss", CASES)
def test_int8_slices_logprob_and_gradient(oracle, D, H, Cc, N, S, loss, mode):
    O = oracle
    spec, prob, q, out_act, _ = problem(O, D, H, Cc, N, S, seed=D + H + N, act="relu", loss=loss)
    U64, loss64, g64 = O.potential(prob, q, np.float64)
    eng = engine(D, H, Cc, "relu", out_act)
    eng.set_option("tc_i8", mode)
    eng.set_dataset(prob.X, prob.y, prob.loss_kind)
    eng.set_prior([0.0], [1.0], _lib.PRIOR_SCALAR)
    eng.set_option("path", _lib.PATH_TENSOR)
    U, ls, g = eng.hmc_eval(q)
    assert int(eng.info("path_used")) == _lib.PATH_TENSOR
    errs = [rel_err(g[s], g64[s]) for s in range(S)]
    print("int8 slices mode %d, %d-%d-%d N=%d S=%d %s: gradient rel err max %.2e, loss rel err max %.2e"
          % (mode, D, H, Cc, N, S, loss, max(errs), np.max(np.abs(ls - loss64) / np.abs(loss64))))
    np.testing.assert_allclose(U, U64, rtol=1e-4)
    np.testing.assert_allclose(ls, loss64, rtol=1e-4)
    assert max(errs) < 1e-4, errs
    # against bf16x3 on the same device
    eng.set_option("tc_i8", 0)
    U0, _, g0 = eng.hmc_eval(q)
    np.testing.assert_allclose(U, U0, rtol=1e-4)
    assert max(rel_err(g[s], g0[s]) for s in range(S)) < 1e-4
    # repeatable bit for bit, and chain batching changes nothing
    eng.set_option("tc_i8", mode)
    eng.set_option("chain_batch", 1)
    U1, _, g1 = eng.hmc_eval(q)
    np.testing.assert_array_equal(U, U1)
    np.testing.assert_array_equal(g, g1)


@pytest.mark.parametrize("mode", [1, 2])
@pytest.mark.parametrize("D,H,Cc,N,S,loss", [(784, 256, 10, 1000, 5, "ce"), (96, 128, 16, 640, 3, "ce"), (128, 256, 3, 385, 150, "mse"),
                                             (784, 256, 12, 257, 2, "ce")])
def test_mma_epilogue_agrees_with_the_simt_epilogue(oracle, D, H, Cc, N, S, loss, mode):
    """tc_epi_mma = 1 (option, default 0; tc_fused_mma.cuh: the two layer-2 products of the fused kernel's epilogue on mma.sync with
    bf16 hi/lo fragments) against the FP32-SIMT epilogue of tc_g1_layer2_fused<.., I8>: the same layer-1 accumulators, the
    layer-2 products in a different arithmetic (2^-16 per product), so the results agree far inside the parity budget —
    and the relu masks, which depend on layer 1 only, are identical."""
    spec, prob, q, out_act, _ = problem(oracle, D, H, Cc, N, S, seed=D + N, act="relu", loss=loss)
    eng = engine(D, H, Cc, "relu", out_act)
    eng.set_option("tc_i8", mode)
    eng.set_dataset(prob.X, prob.y, prob.loss_kind)
    eng.set_prior([0.0], [1.0], _lib.PRIOR_SCALAR)
    eng.set_option("path", _lib.PATH_TENSOR)
    U2, l2, g2 = eng.hmc_eval(q)
    m2 = np.empty((N, H), np.uint8)
    _lib.check(_lib.load().pyb_debug_relu_mask(eng.h, 0, m2.ctypes.data))
    eng.set_option("tc_epi_mma", 0)
    U1, l1, g1 = eng.hmc_eval(q)
    m1 = np.empty((N, H), np.uint8)
    _lib.check(_lib.load().pyb_debug_relu_mask(eng.h, 0, m1.ctypes.data))
    np.testing.assert_array_equal(m1, m2)
    np.testing.assert_allclose(l2, l1, rtol=2e-6)
    err = max(rel_err(g2[s], g1[s]) for s in range(S))
    print("mma epilogue vs simt epilogue, mode %d %d-%d-%d: gradient %.2e" % (mode, D, H, Cc, err))
    assert err < 3e-5


def test_int8_slices_zero_weights_and_unnormalised_data(oracle):
    """All chains of the reference start at W = 0 (HMC.py:69-72): every scale of the scheme degenerates there (zero W1
    columns, constant W2 rows); and data with very different feature ranges, negative values and constant columns (the
    forward operand is scaled per data row after centring, the backward operand per feature)."""
    from test_gpu_tensor import move_off_relu_kinks
    O = oracle
    D, H, Cc, N, S = 784, 256, 10, 400, 3
    spec, prob, q, out_act, rng = problem(O, D, H, Cc, N, S, seed=5)
    X = ((prob.X - 0.3) * (10.0 ** rng.uniform(-2, 2, D))).astype(np.float32)
    X[:, 7] = 0.0
    X[:, 9] = 2.5
    q = move_off_relu_kinks((q * 0.02).astype(np.float32), X, D, H)
    q[0] = 0.0
    prob = O.Problem(spec, X, prob.y, prob.loss_kind, prob.mu, prob.sigma)
    U64, loss64, g64 = O.potential(prob, q, np.float64)
    eng = engine(D, H, Cc, "relu", out_act)
    eng.set_option("tc_i8", 2)
    eng.set_dataset(prob.X, prob.y, prob.loss_kind)
    eng.set_prior([0.0], [1.0], _lib.PRIOR_SCALAR)
    eng.set_option("path", _lib.PATH_TENSOR)
    U, ls, g = eng.hmc_eval(q)
    eng.set_option("tc_i8", 0)
    _, _, g0 = eng.hmc_eval(q)
    print("unnormalised data: int8 vs float64 %s, bf16x3 vs float64 %s" % (["%.1e" % rel_err(g[s], g64[s]) for s in range(S)],
                                                                        ["%.1e" % rel_err(g0[s], g64[s]) for s in range(S)]))
    np.testing.assert_allclose(ls, loss64, rtol=1e-4)
    for s in range(S):
        assert rel_err(g[s], g64[s]) < 1e-4


def test_int8_slices_hmc_iteration_matches_oracle(oracle):
    O = oracle
    D, H, Cc, N, S, L, eps = 784, 256, 10, 640, 3, 4, 1e-3
    spec, prob, q, out_act, rng = problem(O, D, H, Cc, N, S, seed=11)
    p = rng.standard_normal((S, spec.n_params)).astype(np.float32)
    u = np.float32([0.0, 0.999, 0.5])
    want = O.hmc_iteration(prob, q, p, u, eps, 1.0, L, False, O.HMC_REFERENCE, np.float64)
    eng = engine(D, H, Cc)
    eng.set_option("tc_i8", 2)
    eng.set_dataset(prob.X, prob.y, prob.loss_kind)
    eng.set_prior([0.0], [1.0], _lib.PRIOR_SCALAR)
    eng.hmc_init(S, eps, 1.0, L, _lib.HMC_REFERENCE, q0=q)
    eng.hmc_inject(p=p, u=u)
    eng.hmc_run(1, burning=False, sampling=True)
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


def test_int8_slices_predictive(oracle):
    O = oracle
    D, H, Cc, Nt, n = 784, 256, 10, 300, 5
    rng = np.random.default_rng(21)
    spec = O.MLPSpec(D, [H, Cc], ["relu", "softmax"])
    W = (rng.standard_normal((n, spec.n_params)) * 0.05).astype(np.float32)
    x = rng.random((Nt, D)).astype(np.float32)
    freq = np.float32([1, 2, 1, 3, 1])
    mean64, var64 = O.predictive(spec, W, x, freq, np.float64)
    eng = engine(D, H, Cc)
    eng.set_option("tc_i8", 1)
    mean, var, allo = eng.predict(W, x, weights=freq, want_all=True)
    assert int(eng.info("path_used")) == _lib.PATH_TENSOR
    np.testing.assert_allclose(mean, mean64, rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(allo, O.forward(spec, W, x, np.float64), rtol=1e-4, atol=1e-6)


def test_int8_slices_on_a_converging_chain_and_the_loss_guard(oracle):
    """Where the slices stop holding the 1e-4 budget, and what the library does about it.  The dW1 operand dZ1 is sliced
    against an A-PRIORI scale per hidden unit (the bound (max_c W2 - min_c W2) * N_train / N on |dZ1|) and W1 against its
    column maximum: 16-bit FIXED point, an error relative to the largest magnitude.  Random weights do not stress that; a
    converging chain does — more and more rows are confidently right, a few rows carry dZ1 (heavy tails), and the gradient
    becomes a small difference of large per-row terms.  Measured (float64 emulation and device alike):
    err ~ 1.3e-5 / rms_rows(1 - p_y): 3e-5 at loss 0.65, 5e-5 at loss 0.2, 2e-4 at loss 0.07, 7e-4 at loss 0.02 — where
    bf16x3 itself is at 1.1e-4, because the comparison is that ill-conditioned.
    A student is trained on teacher labels (float64 Adam on the host) and the device gradient is compared with the float64
    oracle under the device's own relu mask (kink-aware) at two stages:
      * loss ~ 0.65 (75 % accuracy): the default (tc_i8 = -1) runs on slices and is inside 1e-4;
      * loss ~ 0.02 (100 %): the loss guard (min chain loss < tc_i8_min_loss = 0.35) answers on bf16x3; forcing the slices
        is outside the budget (documented bound below), which is why the guard exists."""
    O = oracle
    D, H, C, N = 784, 256, 10, 2048
    rng = np.random.default_rng(1)
    X = rng.random((N, D))
    y = (X @ rng.normal(0, 1, (D, C))).argmax(1)
    W1, b1 = rng.normal(0, 0.05, (D, H)), np.zeros(H)
    W2, b2 = rng.normal(0, 0.05, (H, C)), np.zeros(C)
    m = [np.zeros_like(v) for v in (W1, b1, W2, b2)]
    v2 = [np.zeros_like(v) for v in (W1, b1, W2, b2)]
    spec = O.MLPSpec(D, [H, C], ["relu", "softmax"])
    Xf, yi = X.astype(np.float32), y.astype(np.int32)
    eng = engine(D, H, C, "relu", "softmax")
    eng.set_dataset(Xf, yi, _lib.LOSS_SPARSE_CE)
    eng.set_prior([0.0], [1.0], _lib.PRIOR_SCALAR)
    eng.set_option("path", _lib.PATH_TENSOR)

    def device_vs_oracle(q, mode):
        eng.set_option("tc_i8", mode)
        _, ls, g = eng.hmc_eval(q)
        mask = np.empty((N, H), np.uint8)
        _lib.check(_lib.load().pyb_debug_relu_mask(eng.h, 0, mask.ctypes.data))
        loss64, g64 = O.mean_loss_and_grad(spec, q, Xf, yi, O.LOSS_SPARSE_CE, np.float64, relu_masks={0: mask.astype(bool)[None]})
        g_like = g[0].astype(np.float64) - q[0].astype(np.float64)  # remove the N(0, 1) prior term
        return rel_err(g_like, g64[0] * N), abs(ls[0] - loss64[0]) / abs(loss64[0]), int(eng.info("tc_split"))

    seen = {}
    for t in range(1, 301):
        A1 = np.maximum(X @ W1 + b1, 0)
        Z2 = A1 @ W2 + b2
        Pm = np.exp(Z2 - Z2.max(1, keepdims=True))
        Pm /= Pm.sum(1, keepdims=True)
        dZ2 = (Pm - np.eye(C)[y]) / N
        dZ1 = (dZ2 @ W2.T) * (A1 > 0)
        for i, (w, g) in enumerate(zip((W1, b1, W2, b2), (X.T @ dZ1, dZ1.sum(0), A1.T @ dZ2, dZ2.sum(0)))):
            m[i] = 0.9 * m[i] + 0.1 * g
            v2[i] = 0.999 * v2[i] + 0.001 * g * g
            w -= 3e-3 * (m[i] / (1 - 0.9 ** t)) / (np.sqrt(v2[i] / (1 - 0.999 ** t)) + 1e-8)
        if t in (100, 300):
            A1 = np.maximum(X @ W1 + b1, 0)
            Z2 = A1 @ W2 + b2
            Pm = np.exp(Z2 - Z2.max(1, keepdims=True))
            Pm /= Pm.sum(1, keepdims=True)
            q = np.concatenate([W1.ravel(), b1, W2.ravel(), b2]).astype(np.float32)[None]
            stats = (float((Z2.argmax(1) == y).mean()), float(-np.log(Pm[np.arange(N), y]).mean()),
                     float(np.sqrt(((1.0 - Pm[np.arange(N), y]) ** 2).mean())))
            res = {mode: device_vs_oracle(q, mode) for mode in (-1, 2, 1, 0)}
            seen[t] = (stats, res)
            print("step %d: accuracy %.3f, loss %.3f, rms(1 - p_y) %.3f; gradient vs float64 (device mask): default %.2e (split %d), "
                  "forced int8 slices %.2e, slices in the forward GEMM only %.2e, bf16x3 %.2e; loss %.1e"
                  % ((t,) + stats + (res[-1][0], res[-1][2], res[2][0], res[1][0], res[0][0], res[2][1])))
    (st100, r100), (st300, r300) = seen[100], seen[300]
    assert st100[1] > 0.35 and st300[1] < 0.1 and st300[0] > 0.97
    # still fitting: the default runs on slices, inside the budget
    assert r100[-1][2] == 2 and r100[-1][0] < 1e-4 and r100[2][0] < 1e-4 and r100[0][0] < 1e-4
    # converged: the guard tripped and the default is the bf16x3 answer (ill-conditioned for every scheme: a looser bound)
    assert int(eng.info("i8_guard_trips")) == 1 and r300[-1][2] == 0
    assert r300[-1][0] == r300[0][0] and r300[0][0] < 3e-4
    assert 1e-4 < r300[2][0] < 5e-3          # forced slices: outside 1e-4 here (documented)
    assert all(v[1] < 1e-4 for v in r300.values())
