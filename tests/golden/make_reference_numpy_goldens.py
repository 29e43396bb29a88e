#!/usr/bin/env python
"""Golden vectors produced BY THE REFERENCE's own code for the parts of the hot path that are plain NumPy / SciPy /
Python inside it.  TensorFlow, TFP, wandb, tfds, ucimlrepo, matplotlib and scikitplot are not installed here, so they are
replaced by inert stubs ONLY to let `import Pyesian...` succeed; nothing stubbed is ever called — the functions below
touch numpy, scipy, math, random and bisect alone:

  * SVGD.baseline__kernel            (Pyesian/optimizers/SVGD.py:165-181)  -> median-heuristic K and dxkxy   [§8 a10]
  * SGLD._init_sgld_lr               (Pyesian/optimizers/SGLD.py:115-121)  -> learning-rate schedule          [§8 f4]
  * Sampled.__init__ / Sampled.sample (Pyesian/distributions/Sampled.py:9-32) -> frequency-weighted draws      [§8 a6, a13]
  * BayesianModel.apply_distribution (Pyesian/nn/BayesianModel.py:25-48)   -> interval bookkeeping / errors   [§8 b]

Run in the build container (where /root/reference exists), with `python -B` so that nothing is written next to the
reference sources:

    python -B tests/golden/make_reference_numpy_goldens.py        # writes tests/golden/reference_numpy.npz / .json
"""
import json
import os
import random
import sys
import warnings
from unittest.mock import MagicMock

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
STUBBED = ["tensorflow", "tensorflow_probability", "wandb", "wandb.integration", "wandb.integration.keras",
           "tensorflow_datasets", "ucimlrepo", "matplotlib", "matplotlib.pyplot", "scikitplot"]


def load_reference():
    sys.dont_write_bytecode = True
    warnings.simplefilter("ignore")
    for name in STUBBED:
        sys.modules.setdefault(name, MagicMock())
    sys.path.insert(0, "/root/reference")
    import Pyesian.optimizers  # noqa: F401
    import Pyesian.distributions  # noqa: F401
    import Pyesian.nn  # noqa: F401
    return (sys.modules["Pyesian.optimizers.SVGD"].SVGD, sys.modules["Pyesian.optimizers.SGLD"].SGLD,
            sys.modules["Pyesian.distributions.Sampled"].Sampled, sys.modules["Pyesian.nn.BayesianModel"].BayesianModel)


class Bag:
    pass


def main():
    SVGD, SGLD, Sampled, BayesianModel = load_reference()
    rng = np.random.default_rng(0)
    arrays, meta = {}, {}

    # ---- median-heuristic kernel: particle sets of several sizes / scales, plus a fixed bandwidth
    kernel_cases = []
    for i, (M, P, scale) in enumerate([(2, 3, 1.0), (5, 7, 1.0), (10, 322, 1.0), (64, 252, 0.3), (33, 40, 5.0), (7, 1, 2.0)]):
        X = rng.normal(0, scale, (M, P))
        me = Bag()
        me._particles = X
        K, dxkxy = SVGD.baseline__kernel(me)
        arrays["kern%d_X" % i], arrays["kern%d_K" % i], arrays["kern%d_dxkxy" % i] = X, K, dxkxy
        K2, d2 = SVGD.baseline__kernel(me, h=1.7)
        arrays["kern%d_K_h17" % i], arrays["kern%d_dxkxy_h17" % i] = K2, d2
        kernel_cases.append(i)
    meta["kernel_cases"] = kernel_cases

    # ---- SGLD schedule
    sched = []
    for (n, up, lo, gam) in [(500, 1e-2, 1e-4, 0.55), (1000, 0.3, 0.05, 0.55), (50, 1e-3, 1e-5, 0.9), (7, 0.5, 0.4, 0.51)]:
        me = Bag()
        me._nb_iterations, me._lr_upper, me._lr_lower, me._lr_gamma = n, up, lo, gam
        SGLD._init_sgld_lr(me)
        steps = sorted({0, 1, 2, n // 3, n // 2, n - 1, n})
        sched.append({"n": n, "lr_upper": up, "lr_lower": lo, "lr_gamma": gam, "steps": steps,
                      "lr": [float(me._lr(s)) for s in steps]})
    meta["sgld_schedule"] = sched

    # ---- Sampled: weighted draws under a seeded `random`, and the constructor's error behaviour
    draws = []
    for seed, freqs in [(0, [1, 5, 2]), (1, [3]), (2, [1] * 9), (3, [10, 1, 1, 10, 2, 7])]:
        s = Sampled([np.full(4, float(k)) for k in range(len(freqs))], list(freqs))
        random.seed(seed)
        draws.append({"seed": seed, "frequencies": list(freqs), "picked": [int(s.sample()[0]) for _ in range(200)]})
    meta["sampled_draws"] = draws
    errs = {}
    for name, args in [("length_mismatch", ([np.zeros(2)], [1, 2])), ("zero_frequency", ([np.zeros(2), np.ones(2)], [1, 0])),
                       ("two_dimensional", ([np.zeros((2, 2))], [1]))]:
        try:
            Sampled(*args)
            errs[name] = None
        except Exception as e:
            errs[name] = [type(e).__name__, str(e)]
    meta["sampled_errors"] = errs

    # ---- BayesianModel.apply_distribution: interval list after a sequence of calls (model_from_json is stubbed, so the
    # layer count is injected; the method itself is plain Python)
    seqs = []
    for n_layers, calls in [(3, [(0, 2)]), (3, [(0, 0), (1, 1), (2, 2)]), (4, [(2, 3), (0, 1)]), (4, [(1, 1), (0, 0), (3, 3), (2, 2)]),
                            (2, [(1, 0)]), (2, [(0, 2)]), (2, [(-1, 0)])]:
        bm = BayesianModel.__new__(BayesianModel)
        bm._n_layers, bm._layers_dtbn_intervals, bm._distributions = n_layers, [], []
        rec = {"n_layers": n_layers, "calls": [list(c) for c in calls]}
        try:
            for j, (a, b) in enumerate(calls):
                bm.apply_distribution("dist%d" % j, a, b)
            rec["intervals"], rec["distributions"] = [list(iv) for iv in bm._layers_dtbn_intervals], list(bm._distributions)
        except Exception as e:
            rec["error"] = [type(e).__name__, str(e)]
        seqs.append(rec)
    meta["apply_distribution"] = seqs

    np.savez_compressed(os.path.join(HERE, "reference_numpy.npz"), **arrays)
    with open(os.path.join(HERE, "reference_numpy.json"), "w") as f:
        json.dump(meta, f, indent=1)
    print("wrote", len(arrays), "arrays;", {k: len(v) for k, v in meta.items()})


if __name__ == "__main__":
    main()
