"""GPU tests straight against the goldens produced by EXECUTING THE REFERENCE's own code (tests/golden/reference_*.npz,
see tests/test_reference_goldens.py and DESIGN.md §2): the device, driven through the C ABI with the same positions,
momenta, uniforms, minibatches and noise, must reproduce what Pyesian's HMC.py / SVGD.py / SGLD.py / SWAG.py /
BayesianModel.py / Metrics.py computed — no oracle in between."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from bayesian_inference_for_nn_b200 import _lib, keras_json  # noqa: E402
from bayesian_inference_for_nn_b200.engine import Engine  # noqa: E402
from conftest import GOLDEN  # noqa: E402


def make(D, units, acts, seed=0):
    return Engine(keras_json.parse_model_json(keras_json.make_sequential_json(D, units, acts)), seed=seed)


HMC_CASES = {"ce": ([5, 2], ["relu", "softmax"], _lib.LOSS_SPARSE_CE, ([0.0], [1.0], _lib.PRIOR_SCALAR)),
             "neg": ([5, 2], ["relu", "softmax"], _lib.LOSS_SPARSE_CE, ([0.0], [-1.0], _lib.PRIOR_SCALAR)),
             "mse": ([4, 1], ["tanh", "linear"], _lib.LOSS_MSE, ([0.0, 0.0, 0.5, 0.5], [1.0, 1.0, 2.0, 2.0], _lib.PRIOR_PER_VARIABLE))}


@pytest.mark.parametrize("path", ["generic", "auto"])
@pytest.mark.parametrize("name", sorted(HMC_CASES))
def test_hmc_iterations_reproduce_the_reference_run(name, path):
    g = np.load(os.path.join(GOLDEN, "reference_hmc.npz"))
    units, acts, loss, prior = HMC_CASES[name]
    D, N, L, n_burn, n_samp, P = (int(v) for v in g[name + "_meta"])
    eps, m = (float(v) for v in g[name + "_hyper"])
    eng = make(D, units, acts)
    eng.set_option("path", {"generic": _lib.PATH_GENERIC, "auto": _lib.PATH_AUTO}[path])
    eng.set_dataset(g[name + "_X"], g[name + "_y"], loss)
    eng.set_prior(*prior)
    for it in range(n_burn + n_samp):
        burning = bool(g[name + "_burning"][it])
        eng.hmc_init(1, eps, m, L, _lib.HMC_REFERENCE, q0=g[name + "_q_before"][it][None])
        eng.hmc_inject(p=g[name + "_p"][it][None], u=np.float32([g[name + "_u"][it]]))
        eng.hmc_run(1, burning=burning, sampling=not burning)
        last = eng.hmc_last()
        for k in ("K0", "U0", "K1", "U1"):
            want = g[name + "_" + k][it]
            if np.isnan(want):
                assert np.isnan(last[k][0]), (it, k)
            else:
                assert abs(last[k][0] - want) <= 2e-4 * max(1.0, abs(want)), (it, k, last[k][0], want)
        la = float(last["log_alpha"][0])
        if burning or np.isnan(la) or abs(la - np.log(max(g[name + "_u"][it], 1e-300))) > 1e-3:
            assert int(last["accept"][0]) == int(g[name + "_accepted"][it]), (it, la)
            q, _ = eng.hmc_state()
            np.testing.assert_allclose(q[0], g[name + "_q_after"][it], rtol=1e-4, atol=5e-6)
            assert abs(last["loss"][0] - g[name + "_ret_loss"][it]) <= 2e-5 * max(1.0, abs(g[name + "_ret_loss"][it]))
    eng.close()


def test_a_continued_chain_reproduces_the_reference_samples():
    """the whole sampling phase of the 'ce' run as ONE chain on the device (carried evaluation, bookkeeping): the Sampled
    contents equal the reference's"""
    g = np.load(os.path.join(GOLDEN, "reference_hmc.npz"))
    name = "ce"
    units, acts, loss, prior = HMC_CASES[name]
    D, N, L, n_burn, n_samp, P = (int(v) for v in g[name + "_meta"])
    eps, m = (float(v) for v in g[name + "_hyper"])
    eng = make(D, units, acts)
    eng.set_dataset(g[name + "_X"], g[name + "_y"], loss)
    eng.set_prior(*prior)
    eng.hmc_init(1, eps, m, L, _lib.HMC_REFERENCE, q0=g[name + "_q_before"][n_burn][None])
    for it in range(n_burn, n_burn + n_samp):
        eng.hmc_inject(p=g[name + "_p"][it][None], u=np.float32([g[name + "_u"][it]]))
        eng.hmc_run(1, burning=False, sampling=True)
        la = float(eng.hmc_last()["log_alpha"][0])
        if abs(la - np.log(max(g[name + "_u"][it], 1e-300))) < 1e-3:
            pytest.skip("a decision of the golden run sits inside the tolerance band")
    samples, freq, chain = eng.hmc_samples()
    assert freq.tolist() == g[name + "_frequencies"].tolist()
    np.testing.assert_allclose(samples, g[name + "_samples"], rtol=2e-4, atol=1e-5)
    eng.close()


@pytest.mark.parametrize("name,units,acts,loss", [("ce", [4, 2], ["relu", "softmax"], _lib.LOSS_SPARSE_CE),
                                                  ("mse", [6, 1], ["tanh", "linear"], _lib.LOSS_MSE)])
def test_svgd_live_steps_reproduce_the_reference_run(name, units, acts, loss):
    g = np.load(os.path.join(GOLDEN, "reference_svgd.npz"))
    D, N, Nv, B, M, steps, P = (int(v) for v in g[name + "_meta"])
    lr, scale = (float(v) for v in g[name + "_hyper"])
    eng = make(D, units, acts)
    eng.set_dataset(g[name + "_X"], g[name + "_y"], loss)
    eng.set_prior([0.0], [scale], _lib.PRIOR_SCALAR)
    eng.svgd_init(M, lr, _lib.SVGD_REFERENCE_LIVE, particles0=g[name + "_before"][0])
    eng.svgd_set_validation(g[name + "_Xv"], g[name + "_yv"])
    n_batches = -(-N // B)
    for s in range(steps):
        b = s % n_batches
        loss_s = eng.svgd_step(np.arange(b * B, min((b + 1) * B, N), dtype=np.int32))
        want, before = g[name + "_after"][s], g[name + "_before"][s]
        got = eng.svgd_particles()
        assert np.abs(got - want).max() <= 5e-6 + 2e-4 * np.abs(want - before).max(), (s, np.abs(got - want).max())
        assert abs(loss_s - g[name + "_ret"][s]) <= 2e-5 * max(1.0, abs(loss_s))
        if s + 1 == 10:
            assert abs(eng.svgd_validation_loss() - g[name + "_valid_losses"][0]) <= 1e-4 * max(1.0, g[name + "_valid_losses"][0])
    eng.close()


@pytest.mark.parametrize("name,units,acts", [("ce", [6, 4], ["relu", "softmax"]), ("reg", [5, 1], ["tanh", "linear"])])
def test_predict_reproduces_the_reference_run(name, units, acts):
    import bisect
    g = np.load(os.path.join(GOLDEN, "reference_predict.npz"))
    W, freq, x = g[name + "_W"], g[name + "_freq"], g[name + "_x"]
    acc = list(np.cumsum(freq))
    idx = [bisect.bisect_left(acc, int(t)) for t in g[name + "_tickets"]]          # Sampled.sample (Sampled.py:29-32)
    eng = make(x.shape[1], units, acts)
    mean, var, allo = eng.predict(W[idx], x, want_all=True)
    np.testing.assert_allclose(allo, g[name + "_samples"], rtol=1e-4, atol=5e-6)
    np.testing.assert_allclose(mean, g[name + "_mean"], rtol=1e-4, atol=5e-6)
    uniq, counts = np.unique(idx, return_counts=True)                              # what BayesianModel.predict sends
    mean_w, _, _ = eng.predict(W[uniq], x, weights=counts.astype(np.float32))
    np.testing.assert_allclose(mean_w, g[name + "_mean"], rtol=1e-4, atol=5e-6)
    eng.close()


def test_sgld_and_swag_steps_reproduce_the_reference_run():
    g = np.load(os.path.join(GOLDEN, "reference_sg.npz"))
    D, N, B, steps = (int(v) for v in g["meta"])
    n_batches = -(-N // B)
    idx = lambda s: np.arange((s % n_batches) * B, min((s % n_batches + 1) * B, N), dtype=np.int32)
    eng = make(D, [5, 2], ["relu", "softmax"])
    eng.set_dataset(g["X"], g["y"], _lib.LOSS_SPARSE_CE)
    eng.sg_init(1, _lib.SG_SGLD, theta0=g["sgld_theta0"][None])
    running = 0.0
    for s in range(steps):
        _, loss = eng.sg_step(float(g["sgld_lr"][s]), idx(s), noise=g["sgld_z"][s][None].astype(np.float32))
        running += loss
        assert abs(running / (s + 1) - g["sgld_ret"][s]) < 1e-5
        np.testing.assert_allclose(eng.sg_state()["theta"][0], g["sgld_theta"][s], rtol=1e-4, atol=1e-6)
    st = eng.sg_state()
    np.testing.assert_allclose(st["mean"][0], g["sgld_mean"], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(st["sq_mean"][0], g["sgld_sq_mean"], rtol=2e-4, atol=1e-6)
    lr, k, freq = float(g["swag_hyper"][0]), int(g["swag_hyper"][1]), int(g["swag_hyper"][2])
    eng.sg_init(1, _lib.SG_SWAG, k_dev=k, frequency=freq, theta0=g["swag_theta0"][None])
    for s in range(steps):
        _, loss = eng.sg_step(lr, idx(s))
        assert abs(loss - g["swag_ret"][s]) < 1e-5
        np.testing.assert_allclose(eng.sg_state()["theta"][0], g["swag_theta"][s], rtol=1e-4, atol=1e-6)
    st = eng.sg_state()
    np.testing.assert_allclose(st["mean"][0], g["swag_mean"], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(st["dev"][0].T, g["swag_dev"], rtol=5e-4, atol=5e-6)
    eng.close()


def test_classification_uncertainty_reproduces_the_reference_run():
    """the goldens hold per-draw probabilities, not weights: a one-layer softmax model over n*C inputs whose k-th weight
    sample picks block k of x = [log p_1 | ... | log p_n] makes the device's forward pass return exactly those"""
    g = np.load(os.path.join(GOLDEN, "reference_metrics.npz"))
    for i in range(int(g["n_cases"])):
        probs, y, arg = g["c%d_probs" % i], g["c%d_y" % i], int(g["c%d_arg" % i])
        n, N, C = probs.shape
        eng = make(n * C, [C], ["softmax"])
        x = np.concatenate([np.log(probs[k]) for k in range(n)], axis=1).astype(np.float32)      # [N, n*C]
        W = np.zeros((n, n * C * C + C), np.float32)
        for k in range(n):
            kern = np.zeros((n * C, C), np.float32)
            kern[k * C:(k + 1) * C] = np.eye(C)
            W[k, :n * C * C] = kern.reshape(-1)
        tot, al, ep, mean = eng.predict_uncertainty(W, x, y, semantics="reference", divisor=arg)
        np.testing.assert_allclose(mean, probs.mean(axis=0), rtol=1e-4, atol=1e-6)
        for got, key in ((tot, "total"), (al, "aleatoric"), (ep, "epistemic")):
            want = g["c%d_%s" % (i, key)]
            assert np.abs(got - want).max() <= 5e-5 * max(1.0, np.abs(want).max()), (i, key)
        eng.close()


@pytest.mark.parametrize("name,sigma", [("shipped_neg", -1.0), ("pos", 1.0)])
def test_hmc_on_make_moons_behaves_like_the_reference_run(name, sigma):
    """BASELINE configs[0] (HMC_classification.py: make_moons, 2-50-2, epsilon 0.005, m 0.5, L 30): the reference's own
    HMC.train / result / BayesianModel.predict were executed on the TensorFlow stand-in (reference_hmc_moons.npz: accept
    rate 0.893 and 95.5 % test accuracy with sigma = +1; accept rate 0 and 96 % with the shipped sigma = -1, whose NaN
    Hamiltonian rejects everything after the always-accept burn-in).  The device's chains (own Philox randomness, so only
    statistics are comparable) must land in the same place."""
    g = np.load(os.path.join(GOLDEN, "reference_hmc_moons.npz"))
    eps, m, L, n_iter = float(g["hyper"][0]), float(g["hyper"][1]), int(g["hyper"][2]), int(g["hyper"][3])
    S = 32
    eng = make(2, [50, 2], ["relu", "softmax"], seed=3)
    eng.set_dataset(g["x_train"], g["y_train"], _lib.LOSS_SPARSE_CE)
    eng.set_prior([0.0], [sigma], _lib.PRIOR_SCALAR)
    eng.hmc_init(S, eps, m, L, _lib.HMC_REFERENCE)
    burn = eng.hmc_run(10, burning=True, sampling=False)          # HMC.train: 10 always-accept iterations first (:106-113)
    assert burn["accept_rate"] == 1.0
    d = eng.hmc_run(n_iter, burning=False, sampling=True)
    ref_rate, ref_acc = float(g[name + "_accept_rate"]), float(g[name + "_accuracy"])
    samples, freq, chain = eng.hmc_samples()
    if sigma < 0:
        assert ref_rate == 0.0 and d["n_accepted"] == 0 and d["n_nan"] == S * n_iter
        assert samples.shape[0] == S and np.all(freq == n_iter + 1)                # the reference run: 1 sample, weight 151
        assert int(g[name + "_n_samples"]) == 1 and int(g[name + "_frequencies"][0]) == n_iter + 1
    else:
        assert abs(d["accept_rate"] - ref_rate) < 0.08, (d["accept_rate"], ref_rate)
        # mean returned loss over the last 30 iterations of the reference chain vs the device's chains
        assert abs(d["mean_loss"] - float(np.mean(g[name + "_losses"][10:]))) < 0.03
    mean, _, _ = eng.predict(samples, g["x_test"], weights=freq.astype(np.float32))
    acc = float((mean.argmax(1) == g["y_test"]).mean())
    assert acc > 0.9 and abs(acc - ref_acc) < 0.06, (acc, ref_acc)
    eng.close()


def test_svgd_on_make_moons_follows_the_reference_trajectory():
    """BASELINE configs[1] as the reference ships it (2-64-2, M = 10, batch 64, lr 1e-3): 400 steps of the reference's own
    SVGD.train executed on the TensorFlow stand-in (reference_svgd_moons.npz).  The device, started from the same
    particles and fed the same minibatches, follows the run: returned losses, particles at the checkpoints, validation
    loss and the ensemble's test accuracy."""
    g = np.load(os.path.join(GOLDEN, "reference_svgd_moons.npz"))
    steps, M, B, lr = int(g["hyper"][0]), int(g["hyper"][1]), int(g["hyper"][2]), float(g["hyper"][3])
    X, y = g["x_train"], g["y_train"]
    nb = -(-X.shape[0] // B)
    eng = make(2, [64, 2], ["relu", "softmax"])
    eng.set_dataset(X, y, _lib.LOSS_SPARSE_CE)
    eng.set_prior([0.0], [1.0], _lib.PRIOR_SCALAR)
    eng.svgd_init(M, lr, _lib.SVGD_REFERENCE_LIVE, particles0=g["particles0"])
    eng.svgd_set_validation(g["x_valid"], g["y_valid"])
    for s in range(steps):
        b = s % nb
        loss = eng.svgd_step(np.arange(b * B, min((b + 1) * B, X.shape[0]), dtype=np.int32))
        assert abs(loss - g["ret"][s]) <= 5e-4 * max(1.0, abs(loss)), (s, loss, g["ret"][s])
        key = "particles_%d" % (s + 1)
        if key in g.files:
            moved = np.abs(g[key] - g["particles0"]).max()
            got = eng.svgd_particles()
            assert np.abs(got - g[key]).max() <= 2e-2 * moved + 1e-6, (s + 1, np.abs(got - g[key]).max(), moved)
    assert abs(eng.svgd_validation_loss() - g["valid_losses"][-1]) < 5e-3
    mean, _, _ = eng.predict(eng.svgd_particles().astype(np.float32), g["x_test"])
    assert abs(float((mean.argmax(1) == g["y_test"]).mean()) - float(g["accuracy"])) <= 0.01
    eng.close()


def test_tensor_path_reproduces_the_reference_at_the_headline_width():
    """784-256-10: the reference's HMC.step (5 iterations, 2048 rows) and BayesianModel.predict (20 draws, 1024 rows)
    executed on the TensorFlow stand-in (reference_wide.npz) against the device's TENSOR-CORE path (bf16x3 split GEMMs,
    fused layer-2 epilogue), one continued chain with the carried evaluation."""
    from test_reference_goldens import _wide_inputs
    import bisect
    g = np.load(os.path.join(GOLDEN, "reference_wide.npz"))
    inp = _wide_inputs()
    eps, m, L = float(g["hmc_hyper"][0]), float(g["hmc_hyper"][1]), int(g["hmc_hyper"][2])
    eng = make(784, [256, 10], ["relu", "softmax"])
    eng.set_dataset(inp["X"], inp["y"], _lib.LOSS_SPARSE_CE)
    eng.set_prior([0.0], [1.0], _lib.PRIOR_SCALAR)
    eng.hmc_init(1, eps, m, L, _lib.HMC_REFERENCE)                       # starts at the prior mean, like the reference
    for it in range(5):
        eng.hmc_inject(p=inp["p"][it][None], u=np.float32([g["hmc_u"][it]]))
        burning = bool(g["hmc_burning"][it])
        eng.hmc_run(1, burning=burning, sampling=not burning)
        assert int(eng.info("path_used")) == _lib.PATH_TENSOR
        last = eng.hmc_last()
        for k in ("K0", "U0", "K1", "U1"):
            assert abs(last[k][0] - g["hmc_" + k][it]) <= 1e-4 * abs(g["hmc_" + k][it]), (it, k, last[k][0], g["hmc_" + k][it])
        assert int(last["accept"][0]) == int(g["hmc_accepted"][it])
        # the energies are ~2e5 and float32 on both sides (resolution 0.016): what decides is their difference
        ref_la = (g["hmc_K0"][it] + g["hmc_U0"][it]) - g["hmc_K1"][it] - g["hmc_U1"][it]
        assert abs(float(last["log_alpha"][0]) - ref_la) < 0.1, (it, float(last["log_alpha"][0]), ref_la)
        q, _ = eng.hmc_state()
        assert abs(np.linalg.norm(q.astype(np.float64)) - g["hmc_q_norms"][it]) <= 1e-4 * g["hmc_q_norms"][it]
    assert np.linalg.norm(q[0] - g["hmc_q_final"]) <= 1e-3 * np.linalg.norm(g["hmc_q_final"])
    acc = list(np.cumsum(inp["freq"]))
    idx = [bisect.bisect_left(acc, int(t)) for t in g["pred_tickets"]]
    mean, _, allo = eng.predict(inp["W"][idx], inp["x"], want_all=True)
    assert int(eng.info("path_used")) == _lib.PATH_TENSOR
    np.testing.assert_allclose(allo, g["pred_samples"], rtol=2e-4, atol=5e-6)
    np.testing.assert_allclose(mean, g["pred_mean"], rtol=2e-4, atol=5e-6)
    eng.close()


def test_tensor_path_reproduces_the_reference_svgd_steps_at_the_mnist_width():
    """784-128-10 (SVGD_mnist's model), 5 close particles, minibatch 256, 3 live steps of the reference's own SVGD.step
    (reference_wide_svgd.npz) against the device: minibatch gradients on the tensor-core path for 128 hidden units (dW1
    with two particles per CTA pair and an odd particle count), float64 gamma = 1 kernel terms, legacy-Adam descent."""
    from test_reference_goldens import _wide_svgd_inputs
    g = np.load(os.path.join(GOLDEN, "reference_wide_svgd.npz"))
    inp, ns = _wide_svgd_inputs()
    B, M, lr = ns["B"], ns["M"], ns["LR"]
    eng = make(784, [128, 10], ["relu", "softmax"])
    eng.set_dataset(inp["X"], inp["y"], _lib.LOSS_SPARSE_CE)
    eng.set_prior([0.0], [ns["SCALE"]], _lib.PRIOR_SCALAR)
    eng.svgd_init(M, lr, _lib.SVGD_REFERENCE_LIVE, particles0=inp["particles0"])
    for s in range(ns["STEPS"]):
        loss = eng.svgd_step(np.arange(s * B, (s + 1) * B, dtype=np.int32))
        assert int(eng.info("path_used")) == _lib.PATH_TENSOR
        assert abs(loss - g["ret"][s]) < 5e-6, (s, loss, g["ret"][s])
        if s == 0:
            step = eng.svgd_particles() - inp["particles0"]
            agree = np.mean(np.sign(step) == g["step1_sign"])        # sign of -phi, element by element
            assert agree > 0.999, agree
            assert abs(np.abs(step).max() - float(g["step1_absmax"])) < 1e-5
    got = eng.svgd_particles()
    moved = np.abs(g["final"] - inp["particles0"]).max()
    assert np.abs(got - g["final"]).max() <= 5e-2 * moved, (np.abs(got - g["final"]).max(), moved)
    eng.close()
