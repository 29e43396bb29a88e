#!/usr/bin/env python
"""Small-width measurements (BASELINE configs C1/C2): HMC make_moons 2-50-2 (N=1600, L=30,
eps=0.005, m=0.5) for S chains, and one SVGD step for 64 particles.  Prints one JSON line per case.
Roofline: 1.6 MFLOP and 23 232 algorithmic bytes per grad-eval; FP32 SIMT peak
148 SM x 128 lanes x 2 x 1.965 GHz = 74.4 TFLOP/s."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bayesian_inference_for_nn_b200 import _lib, keras_json  # noqa: E402
from bayesian_inference_for_nn_b200.engine import Engine  # noqa: E402


def moons(n, seed=0, noise=0.2):
    rng = np.random.default_rng(seed)
    n0 = n // 2
    t0, t1 = rng.uniform(0, np.pi, n0), rng.uniform(0, np.pi, n - n0)
    x = np.concatenate([np.stack([np.cos(t0), np.sin(t0)], 1), np.stack([1 - np.cos(t1), 0.5 - np.sin(t1)], 1)])
    y = np.concatenate([np.zeros(n0, np.int32), np.ones(n - n0, np.int32)])
    return (x + rng.normal(0, noise, x.shape)).astype(np.float32), y


def main():
    X, y = moons(1600)
    spec = keras_json.parse_model_json(keras_json.make_sequential_json(2, [50, 2], ["relu", "softmax"]))
    peaks = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json"))) if os.path.exists(
        os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}
    for path, pname in ((_lib.PATH_AUTO, "fused_small"), (_lib.PATH_GENERIC, "generic")):
        for S in (1, 1024, 16384, 131072):
            if pname == "generic" and S > 16384:
                continue
            eng = Engine(spec, seed=1)
            eng.set_option("path", path)
            eng.set_dataset(X, y, _lib.LOSS_SPARSE_CE)
            eng.set_prior([0.0], [1.0], _lib.PRIOR_SCALAR)
            L = 30
            eng.hmc_init(S, 0.005, 0.5, L)
            eng.hmc_run(3, burning=True, sampling=False)
            iters = 20 if S <= 1024 else (5 if S <= 16384 else 2)
            t0 = time.perf_counter()
            d = eng.hmc_run(iters, burning=False, sampling=False)
            wall = time.perf_counter() - t0
            evals = S * L * iters
            rate = evals / (d["device_ms"] / 1e3)
            print(json.dumps({"case": "C1 HMC moons 2-50-2 N=1600 L=30", "path": pname, "chains": S, "iters": iters,
                              "grad_evals_per_s": rate, "ms_per_iteration": d["device_ms"] / iters,
                              "wall_ms_per_iteration": 1e3 * wall / iters, "accept_rate": d["accept_rate"],
                              "launches_per_iteration": d["kernel_launches"] / iters,
                              "fp32_tflops_algorithmic": rate * 1.6e6 / 1e12, "frac_of_fp32_simt_peak_74.4": rate * 1.6e6 / 74.4e12,
                              "algorithmic_GBps": rate * 23232 / 1e9, "frac_of_measured_hbm": rate * 23232 / 1e9 / peaks["hbm_gbs"]}),
                  flush=True)
            eng.close()
    # C2: SVGD, 64 particles, full batch
    for sem, name in ((_lib.SVGD_REFERENCE_LIVE, "reference_live"), (_lib.SVGD_CANONICAL_MEDIAN, "canonical_median")):
        eng = Engine(spec, seed=1)
        eng.set_dataset(X, y, _lib.LOSS_SPARSE_CE)
        eng.set_prior([0.0], [1.0], _lib.PRIOR_SCALAR)
        eng.svgd_init(64, 1e-3, sem)
        for _ in range(3):
            eng.svgd_step(None)
        t0 = time.perf_counter()
        n = 20
        ms = 0.0
        for _ in range(n):
            eng.svgd_step(None)
            ms += eng.info("last_device_ms")
        wall = time.perf_counter() - t0
        print(json.dumps({"case": "C2 SVGD moons 64 particles full batch", "semantics": name, "steps_per_s": n / (ms / 1e3),
                          "device_ms_per_step": ms / n, "wall_ms_per_step": 1e3 * wall / n,
                          "particle_grad_evals_per_s": 64 * n / (ms / 1e3)}), flush=True)
        eng.close()


if __name__ == "__main__":
    main()
