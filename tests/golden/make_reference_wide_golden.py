#!/usr/bin/env python
"""The headline model shape (784-256-10, BASELINE configs[2] and [4]) through the reference's own code at reduced row
counts, so that the tensor-core path of the device is compared DIRECTLY with what Pyesian computes:

  * HMC: the reference's HMC.step (2 always-accept burn-in iterations to leave the all-zero start, then 3 sampling
    iterations; L = 3, epsilon = 2e-3, m = 1, prior N(0,1), 2048 rows) — energies, decisions, end points;
  * predictive: the reference's BayesianModel.predict for 12 stored weight samples over 1024 rows — per-draw outputs, mean.

Executed on the TensorFlow stand-in of tf_shim.py (torch CPU float32).

    python -B tests/golden/make_reference_wide_golden.py        # ~1 minute; writes tests/golden/reference_wide.npz
"""
import os
import random
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import tf_shim  # noqa: E402
from make_reference_hmc_golden import flat, load_reference  # noqa: E402


D, H, C, N, N_STORED, NT = 784, 256, 10, 2048, 12, 1024
SHAPES = [(D, H), (H,), (H, C), (C,)]


def inputs():
    """every input of the run, re-created from its seed (the test calls this too; nothing here needs the reference)"""
    rng = np.random.default_rng(0)
    X = rng.random((N, D), dtype=np.float32)
    y = rng.integers(0, C, N).astype(np.int64)
    P = sum(int(np.prod(s)) for s in SHAPES)
    W = (rng.standard_normal((N_STORED, P)) * 0.05).astype(np.float32)
    freq = rng.integers(1, 4, N_STORED)
    x = rng.random((NT, D), dtype=np.float32)
    prng = np.random.default_rng(9)                  # the tf.random.normal queue of the HMC run (m = 1)
    p = [np.concatenate([prng.standard_normal(s).astype(np.float32).reshape(-1) for s in SHAPES]) for _ in range(5)]
    return dict(X=X, y=y, W=W, freq=freq, x=x, p=p)


def main():
    import torch
    HMC, GaussianPrior, HyperParameters = load_reference()
    BayesianModel = sys.modules["Pyesian.nn.BayesianModel"].BayesianModel
    Sampled = sys.modules["Pyesian.distributions.Sampled"].Sampled
    from bayesian_inference_for_nn_b200 import keras_json          # host-side JSON writer only
    L, eps, m = 3, 2e-3, 1.0
    js = keras_json.make_sequential_json(D, [H, C], ["relu", "softmax"])
    rng = np.random.default_rng(0)
    X = rng.random((N, D), dtype=np.float32)
    y = rng.integers(0, C, N).astype(np.int64)
    data = tf_shim.ArrayData(X, y)
    dataset = types.SimpleNamespace(training_dataset=lambda: data,
                                    loss=lambda reduction="auto": tf_shim.SparseCategoricalCrossentropy(reduction=reduction))

    class Recorder(HMC):
        trace = None

        def _kinetic_energy(self):
            v = super()._kinetic_energy()
            self.trace.append(("K", float(v.numpy().reshape(-1)[0])))
            return v

        def _potential_energy(self):
            u, l = super()._potential_energy()
            self.trace.append(("U", float(u.numpy().reshape(-1)[0]), float(l.numpy())))
            return u, l

    opt = Recorder()
    opt.trace = []
    opt.compile(HyperParameters(epsilon=eps, m=m, L=L), js, dataset, verbose=False, prior=GaussianPrior(0.0, 1.0))
    tf_shim.RANDOM.rng = np.random.default_rng(9)
    rec = {k: [] for k in ("q_before", "p", "u", "burning", "K0", "U0", "K1", "U1", "loss1", "accepted", "q_after")}
    for it in range(5):
        burning = it < 2
        rec["q_before"].append(flat(opt._model))
        random.seed(50 + it)
        state = random.getstate()
        u = random.random()
        random.setstate(state)
        tf_shim.RANDOM.log.clear()
        opt.trace.clear()
        acc0 = opt._accepted_runs
        opt.step(sampling=not burning, burning=burning)
        ks = [t for t in opt.trace if t[0] == "K"]
        us = [t for t in opt.trace if t[0] == "U"]
        rec["p"].append(np.concatenate([z.reshape(-1) for z in tf_shim.RANDOM.log]).astype(np.float32) * np.float32(m))
        rec["u"].append(u); rec["burning"].append(burning)
        rec["K0"].append(ks[0][1]); rec["K1"].append(ks[1][1]); rec["U0"].append(us[0][1]); rec["U1"].append(us[-1][1])
        rec["loss1"].append(us[-1][2]); rec["accepted"].append(opt._accepted_runs - acc0)
        rec["q_after"].append(flat(opt._model))
        print("HMC iteration", it, "U0 %.3f U1 %.3f K0 %.3f K1 %.3f accepted %d" % (us[0][1], us[-1][1], ks[0][1], ks[1][1],
                                                                                rec["accepted"][-1]))
    # small fixture: the inputs are re-created from the seeds by inputs() below; only results are stored
    out = {"hmc_" + k: np.asarray(rec[k]) for k in ("u", "burning", "K0", "U0", "K1", "U1", "loss1", "accepted")}
    out["hmc_q_norms"] = np.asarray([np.linalg.norm(q.astype(np.float64)) for q in rec["q_after"]])
    out["hmc_q_final"] = rec["q_after"][-1]
    out["hmc_hyper"] = np.asarray([eps, m, L])
    chk = inputs()
    assert np.array_equal(chk["X"], X) and np.array_equal(chk["y"], y) and all(np.array_equal(a, b) for a, b in zip(chk["p"], rec["p"]))

    # ---- predictive at the same width
    n_stored, Nt, nb = N_STORED, NT, 20
    bm = BayesianModel(js)
    P = out["hmc_q_final"].shape[0]
    W = (rng.standard_normal((n_stored, P)) * 0.05).astype(np.float32)
    freq = rng.integers(1, 4, n_stored).tolist()
    bm.apply_distribution(Sampled([tf_shim.TT(torch.as_tensor(w)) for w in W], freq), 0, 1)
    x = rng.random((Nt, D), dtype=np.float32)
    random.seed(11)
    samples, mean = bm.predict(tf_shim.TT(torch.as_tensor(x)), nb)
    random.seed(11)
    acc = np.cumsum(freq)
    assert np.array_equal(chk["W"], W) and np.array_equal(chk["x"], x) and chk["freq"].tolist() == freq
    out.update(pred_tickets=np.asarray([random.randint(1, int(acc[-1])) for _ in range(nb)]),
               pred_samples=np.stack([s.numpy() for s in samples]), pred_mean=mean.numpy())
    np.savez_compressed(os.path.join(HERE, "reference_wide.npz"), **out)
    print("predictive outputs", out["pred_samples"].shape)


if __name__ == "__main__":
    main()
