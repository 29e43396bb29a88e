#!/usr/bin/env python
"""Golden trajectories produced by EXECUTING THE REFERENCE's Pyesian/optimizers/HMC.py (compile_extra_components, step,
_step_p, _step_q, _kinetic_energy, _potential_energy, _sample_kinetic_energy, _snapshot_q, result — unmodified) together
with its GaussianPrior, Sampled and Optimizer classes, on the torch-backed TensorFlow stand-in of tf_shim.py (TensorFlow is
not installable in the build container).  See tf_shim.py for exactly what that pins: the reference's own control flow and
arithmetic (leapfrog schedule and kick counts, momentum scale, both energies, the Metropolis rule with random.random(),
restore and sample/frequency bookkeeping, the flatten order of result()); Keras / tfp numerics come from their definitions.

    python -B tests/golden/make_reference_hmc_golden.py        # writes tests/golden/reference_hmc.npz
"""
import os
import random
import sys
import types
import warnings
from unittest.mock import MagicMock

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import tf_shim  # noqa: E402


def load_reference():
    sys.dont_write_bytecode = True
    warnings.simplefilter("ignore")
    sys.modules["tensorflow"] = tf_shim.make_tf()
    sys.modules["tensorflow_probability"] = tf_shim.make_tfp()
    for name in ["wandb", "wandb.integration", "wandb.integration.keras", "tensorflow_datasets", "ucimlrepo", "matplotlib",
                 "matplotlib.pyplot", "scikitplot"]:
        sys.modules.setdefault(name, MagicMock())
    sys.path.insert(0, "/root/reference")
    import Pyesian.optimizers  # noqa: F401
    import Pyesian.distributions  # noqa: F401
    from Pyesian.optimizers.hyperparameters import HyperParameters
    return (sys.modules["Pyesian.optimizers.HMC"].HMC, sys.modules["Pyesian.distributions.GaussianPrior"].GaussianPrior,
            HyperParameters)


def flat(model):
    return np.concatenate([v.numpy().reshape(-1) for l in model.layers for v in l.trainable_variables]).astype(np.float32)


def run_case(HMC, GaussianPrior, HyperParameters, name, D, units, acts, N, loss, prior, m, L, eps, n_burn, n_samp, seed):
    from bayesian_inference_for_nn_b200 import keras_json          # host-side JSON writer only (no oracle, no device)
    rng = np.random.default_rng(seed)
    X = rng.normal(size=(N, D)).astype(np.float32)
    if loss == "ce":
        y = rng.integers(0, units[-1], N).astype(np.int64)
        loss_cls = tf_shim.SparseCategoricalCrossentropy
    else:
        y = rng.normal(size=(N, units[-1])).astype(np.float32)
        loss_cls = tf_shim.MeanSquaredError
    data = tf_shim.ArrayData(X, y)
    dataset = types.SimpleNamespace(training_dataset=lambda: data, loss=lambda reduction="auto": loss_cls(reduction=reduction))

    class Recorder(HMC):                    # records what the reference's own methods return; changes nothing
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
    opt.compile(HyperParameters(epsilon=eps, m=m, L=L), keras_json.make_sequential_json(D, units, acts), dataset,
                verbose=False, prior=GaussianPrior(*prior))
    P = flat(opt._model).shape[0]
    rec = {k: [] for k in ("q_before", "p", "u", "burning", "K0", "U0", "loss0", "K1", "U1", "loss1", "ret_loss", "accepted",
                           "q_after")}
    tf_shim.RANDOM.rng = np.random.default_rng(seed + 100)
    for it in range(n_burn + n_samp):
        burning = it < n_burn
        if it == n_burn:                                   # HMC.train resets the books between the phases (:115-118)
            opt._accepted_runs = opt._total_runs = 0
            opt._frequency, opt._samples = [], []
        rec["q_before"].append(flat(opt._model))
        random.seed(1000 * seed + it)
        state = random.getstate()
        u = random.random()
        random.setstate(state)
        tf_shim.RANDOM.log.clear()
        opt.trace.clear()
        acc0 = opt._accepted_runs
        ret = opt.step(sampling=not burning, burning=burning)
        ks = [t for t in opt.trace if t[0] == "K"]
        us = [t for t in opt.trace if t[0] == "U"]
        assert len(ks) == 2 and len(us) == L + 4, (len(ks), len(us))       # U0, L+2 kicks, U1
        rec["p"].append(np.concatenate([z.reshape(-1) for z in tf_shim.RANDOM.log]).astype(np.float32) * np.float32(m))
        rec["u"].append(u)
        rec["burning"].append(burning)
        rec["K0"].append(ks[0][1]); rec["K1"].append(ks[1][1])
        rec["U0"].append(us[0][1]); rec["loss0"].append(us[0][2])
        rec["U1"].append(us[-1][1]); rec["loss1"].append(us[-1][2])
        rec["ret_loss"].append(float(ret.numpy()))
        rec["accepted"].append(opt._accepted_runs - acc0)
        rec["q_after"].append(flat(opt._model))
    bm = opt.result()
    dist = bm._distributions[0]
    out = {name + "_" + k: np.asarray(v) for k, v in rec.items()}
    out[name + "_samples"] = np.stack([s.numpy() for s in dist._samples]).astype(np.float32)
    out[name + "_frequencies"] = np.asarray(dist._frequencies, dtype=np.int64)
    out[name + "_intervals"] = np.asarray(bm._layers_dtbn_intervals, dtype=np.int64)
    out[name + "_X"], out[name + "_y"] = X, y
    out[name + "_meta"] = np.asarray([D, N, L, n_burn, n_samp, P], dtype=np.int64)
    out[name + "_hyper"] = np.asarray([eps, m], dtype=np.float64)
    print(name, "P =", P, "accepted:", rec["accepted"], "freq:", dist._frequencies)
    return out


def main():
    HMC, GaussianPrior, HyperParameters = load_reference()
    out = {}
    # make_moons-like classification, m != 1 (momentum scale and kinetic energy use m differently), mixed accept/reject
    out.update(run_case(HMC, GaussianPrior, HyperParameters, "ce", 2, [5, 2], ["relu", "softmax"], 40, "ce", (0.0, 1.0),
                        0.5, 3, 0.12, 2, 7, seed=1))
    # the shipped scripts' prior GaussianPrior(0.0, -1.0): NaN Hamiltonian => nothing accepted after burn-in
    out.update(run_case(HMC, GaussianPrior, HyperParameters, "neg", 2, [5, 2], ["relu", "softmax"], 40, "ce", (0.0, -1.0),
                        0.5, 3, 0.01, 2, 3, seed=2))
    # regression, per-layer list prior, tanh hidden layer, m = 1
    out.update(run_case(HMC, GaussianPrior, HyperParameters, "mse", 3, [4, 1], ["tanh", "linear"], 30, "mse",
                        ([0.0, 0.5], [1.0, 2.0]), 1.0, 2, 0.08, 1, 5, seed=3))
    np.savez_compressed(os.path.join(HERE, "reference_hmc.npz"), **out)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
