"""Generate the golden fixtures under tests/golden/ from the CPU oracle (float64 arithmetic on
float32 inputs).  Run from the repo root:  python tests/golden/make_golden.py

The reference cannot be imported in this container (TensorFlow/TFP absent) and its own tests hold
no vectors, so these fixtures freeze the ORACLE's restatement (parity unpinned — see the oracle
header).  They protect against drift of the oracle itself and give the GPU tests seed-independent
targets.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
import pyesian_oracle as O  # noqa: E402


def moons(n, rng, noise=0.2):
    """make_moons-like two interleaving half circles (no sklearn needed)."""
    n0 = n // 2
    t0, t1 = rng.uniform(0, np.pi, n0), rng.uniform(0, np.pi, n - n0)
    x = np.concatenate([np.stack([np.cos(t0), np.sin(t0)], 1), np.stack([1 - np.cos(t1), 0.5 - np.sin(t1)], 1)])
    y = np.concatenate([np.zeros(n0, np.int32), np.ones(n - n0, np.int32)])
    x = x + rng.normal(0, noise, x.shape)
    perm = rng.permutation(n)
    return x[perm].astype(np.float32), y[perm]


def hmc_case(name, spec, X, y, loss_kind, S, L, eps, m, sigma, seed, semantics=O.HMC_REFERENCE, q_scale=0.3):
    rng = np.random.default_rng(seed)
    P = spec.n_params
    mu, sg = O.expand_prior(spec, 0.0, sigma)
    prob = O.Problem(spec, X, y, loss_kind, mu, sg)
    q = (rng.standard_normal((S, P)) * q_scale).astype(np.float32)
    pstd = m if semantics == O.HMC_REFERENCE else np.sqrt(m)
    p = (rng.standard_normal((S, P)) * pstd).astype(np.float32)
    u = rng.random(S).astype(np.float32)
    U, loss, g = O.potential(prob, q, np.float64)
    r = O.hmc_iteration(prob, q, p, u, eps, m, L, False, semantics, np.float64)
    # choose the uniforms so that the batch mixes accepts and rejects with a clear margin
    with np.errstate(over="ignore"):
        alpha = np.exp(np.minimum(r["log_alpha"], 50.0))
    u = np.minimum(0.999, alpha * np.where(np.arange(S) % 2 == 0, 0.5, 2.0)).astype(np.float32)
    r = O.hmc_iteration(prob, q, p, u, eps, m, L, False, semantics, np.float64)
    np.savez_compressed(
        os.path.join(HERE, name + ".npz"),
        in_dim=spec.in_dim, units=np.int32(spec.units), acts=np.int32(spec.acts), use_bias=np.int32(spec.use_bias), X=X, y=y, loss_kind=loss_kind,
        S=S, L=L, eps=eps, m=m, sigma=sigma, semantics=semantics, q=q, p=p, u=u,
        U=U, loss=loss, grad=g.astype(np.float32), U0=r["U0"], K0=r["K0"], U1=r["U1"], K1=r["K1"],
        log_alpha=r["log_alpha"], accept=r["accept"], q_out=r["q"].astype(np.float32),
        qL=r["qL"].astype(np.float32), pL=r["pL"].astype(np.float32), loss0=r["loss0"], loss1=r["loss1"])
    print(name, "accept", r["accept"], "log_alpha", r["log_alpha"])


def main():
    rng = np.random.default_rng(20260418)
    # C1-shaped mini: 2-50-2 on moons
    X, y = moons(200, rng)
    c1 = O.MLPSpec(2, [50, 2], ["relu", "softmax"])
    hmc_case("hmc_c1_mini", c1, X, y, O.LOSS_SPARSE_CE, S=4, L=5, eps=0.005, m=0.5, sigma=1.0, seed=1)
    hmc_case("hmc_c1_canonical", c1, X, y, O.LOSS_SPARSE_CE, S=3, L=4, eps=0.01, m=2.0, sigma=1.0, seed=2,
             semantics=O.HMC_CANONICAL)
    # C3-shaped mini: 784-32-10 on MNIST-shaped uniforms
    Xm = rng.random((96, 784)).astype(np.float32)
    ym = rng.integers(0, 10, 96).astype(np.int32)
    c3 = O.MLPSpec(784, [32, 10], ["relu", "softmax"])
    hmc_case("hmc_c3_mini", c3, Xm, ym, O.LOSS_SPARSE_CE, S=2, L=3, eps=1e-3, m=1.0, sigma=1.0, seed=3, q_scale=0.05)
    # regression 1-1-1 linear (HMC_regression.py:36-39)
    xr = rng.uniform(1, 20, (64, 1)).astype(np.float32)
    yr = (2 * xr + 2).astype(np.float32)
    reg = O.MLPSpec(1, [1, 1], ["linear", "linear"])
    hmc_case("hmc_regression", reg, xr, yr, O.LOSS_MSE, S=3, L=6, eps=2e-5, m=1.0, sigma=1.0, seed=4, q_scale=1.0)
    # deeper stack with every activation, MSE
    deep = O.MLPSpec(5, [7, 6, 4, 3], ["tanh", "sigmoid", "relu", "linear"], [True, False, True, True])
    Xd = rng.standard_normal((40, 5)).astype(np.float32)
    yd = rng.standard_normal((40, 3)).astype(np.float32)
    hmc_case("hmc_deep_mse", deep, Xd, yd, O.LOSS_MSE, S=3, L=3, eps=1e-3, m=1.0, sigma=2.0, seed=5, q_scale=0.7)

    # SVGD: 6 particles on the moons mini, minibatch 32
    P = c1.n_params
    parts = (rng.standard_normal((6, P)) * 0.2).astype(np.float32).astype(np.float64)
    idx = [rng.permutation(200)[:32].astype(np.int32) for _ in range(2)]
    am, av = np.zeros((6, P), np.float32), np.zeros((6, P), np.float32)
    pl = parts.copy()
    live_losses, live_phis = [], []
    for t, ix in enumerate(idx, 1):
        pl, am, av, loss, phi = O.svgd_live_step(c1, pl, X[ix], y[ix], O.LOSS_SPARSE_CE, am, av, t, 1e-2)
        live_losses.append(loss)
        live_phis.append(phi)
    mu, sg = O.expand_prior(c1, 0.0, 1.0)
    prob = O.Problem(c1, X, y, O.LOSS_SPARSE_CE, mu, sg)
    am, av = np.zeros((6, P), np.float32), np.zeros((6, P), np.float32)
    pc = parts.copy()
    can_losses, can_phis, can_h = [], [], []
    for t, ix in enumerate(idx, 1):
        pc, am, av, loss, phi, h = O.svgd_canonical_step(prob, pc, X[ix], y[ix], am, av, t, 1e-2)
        can_losses.append(loss)
        can_phis.append(phi)
        can_h.append(h)
    G = rng.standard_normal((6, P)).astype(np.float32)
    phi_hook, h_hook, K_hook = O.svgd_phi_canonical(parts, G)
    np.savez_compressed(os.path.join(HERE, "svgd_mini.npz"), X=X, y=y, particles0=parts, idx=np.stack(idx), lr=1e-2,
                        live_particles=pl, live_losses=live_losses, live_phi_last=live_phis[-1],
                        can_particles=pc, can_losses=can_losses, can_phi_last=can_phis[-1], can_h=can_h,
                        G=G, phi_hook=phi_hook.astype(np.float32), h_hook=h_hook, K_hook=K_hook)
    print("svgd live losses", live_losses, "canonical", can_losses, "h", can_h)

    # predictive: 7 weight samples with frequencies on a 50-point grid
    W = (rng.standard_normal((7, P)) * 0.5).astype(np.float32)
    freq = np.int32([1, 3, 1, 2, 5, 1, 1])
    xg = rng.uniform(-2, 3, (50, 2)).astype(np.float32)
    mean, var = O.predictive(c1, W, xg, freq, np.float64)
    mean1, var1 = O.predictive(c1, W, xg, None, np.float64)
    np.savez_compressed(os.path.join(HERE, "predict_mini.npz"), W=W, freq=freq, x=xg, mean=mean, var=var,
                        mean_unweighted=mean1, var_unweighted=var1, mask=O.uncertainty_mask(mean, 0.7))

    # Philox stream samples (device RNG restatement)
    z = O.philox_normals(0x1234ABCD5678, [0, 1, 1000], 7, O.STREAM_MOMENTUM, 37)
    uu = O.philox_uniforms(0x1234ABCD5678, [0, 1, 1000], 7)
    np.savez_compressed(os.path.join(HERE, "philox.npz"), seed=np.uint64(0x1234ABCD5678), chains=np.int32([0, 1, 1000]),
                        iteration=7, normals=z, uniforms=uu)


if __name__ == "__main__":
    main()
