#!/usr/bin/env python
"""Measurements of the SURVEY §8(f) rows built after the headline path: S-batched SGLD / SWAG steps
(784-128-10, minibatch 1024 of a 60000-row pool, 1024 chains; and the reference's own make_moons shape) and
Metrics.classification_uncertainty (1000 weight samples x 10000x784 rows, 10 classes) — with the oracle's
reference-shaped CPU loop timed beside each on a bounded sample.  One JSON line per case."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from bayesian_inference_for_nn_b200 import _lib, keras_json  # noqa: E402
from bayesian_inference_for_nn_b200.engine import Engine  # noqa: E402


def spec(D, H, C):
    return keras_json.parse_model_json(keras_json.make_sequential_json(D, [H, C], ["relu", "softmax"]))


def sg_chains(D, H, C, N, B, S, steps=20):
    import pyesian_oracle as O
    rng = np.random.default_rng(0)
    sp = spec(D, H, C)
    X = rng.random((N, D), dtype=np.float32)
    y = rng.integers(0, C, N).astype(np.int32)
    P = sp.n_params
    for kind, name in ((_lib.SG_SGLD, "SGLD"), (_lib.SG_SWAG, "SWAG k=20 frequency=1")):
        eng = Engine(sp, seed=1)
        eng.set_dataset(X, y, _lib.LOSS_SPARSE_CE)
        eng.sg_init(S, kind, k_dev=20, frequency=1)
        idx = [rng.permutation(N)[:B].astype(np.int32) for _ in range(steps + 3)]
        for ix in idx[:3]:
            eng.sg_step(1e-3, ix)
        ms = []
        t0 = time.perf_counter()
        for ix in idx[3:]:
            eng.sg_step(1e-3, ix)
            ms.append(eng.info("last_device_ms"))
        wall = (time.perf_counter() - t0) / steps
        ms = float(np.median(ms))
        flops = S * 6.0 * B * (D * H + H * C)
        # CPU: the reference's loop shape = ONE model per step (SGLD.py:46-95), NumPy float32 on the host cores
        so = O.MLPSpec(D, [H, C], ["relu", "softmax"])
        st = O.sg_init_state(rng.normal(0, 0.05, (1, P)).astype(np.float32))
        z = rng.standard_normal((1, P)).astype(np.float32)
        kk = O.SG_SGLD if kind == _lib.SG_SGLD else O.SG_SWAG
        O.sg_step(so, st, X[idx[0]], y[idx[0]], O.LOSS_SPARSE_CE, kk, 1e-3, z=z, k=20)
        t0 = time.perf_counter()
        n_cpu = 0
        while time.perf_counter() - t0 < 3.0:
            O.sg_step(so, st, X[idx[n_cpu % len(idx)]], y[idx[n_cpu % len(idx)]], O.LOSS_SPARSE_CE, kk, 1e-3, z=z, k=20)
            n_cpu += 1
        cpu_steps_per_s = n_cpu / (time.perf_counter() - t0)
        print(json.dumps({"case": "%s %d-%d-%d, minibatch %d of %d, %d chains" % (name, D, H, C, B, N, S),
                          "device_ms_per_step": ms, "wall_ms_per_step_through_c_abi": 1e3 * wall,
                          "chain_steps_per_s": S * 1e3 / ms, "algorithmic_tflops_gradients": flops / (ms / 1e3) / 1e12,
                          "update_pass_bytes": (28 if kind == _lib.SG_SGLD else 32) * S * P,
                          "grad_path": int(eng.info("path_used")),
                          "cpu_baseline": {"value": cpu_steps_per_s, "unit": "chain-steps/s", "cores": os.cpu_count(),
                                           "kind": "port", "sample": "one model, %d steps, numpy fp32" % n_cpu}}),
              flush=True)
        eng.close()


def uncertainty():
    import pyesian_oracle as O
    rng = np.random.default_rng(0)
    n, Nt = 1000, 10000
    sp = spec(784, 256, 10)
    W = (rng.standard_normal((n, sp.n_params)) * 0.05).astype(np.float32)
    x = rng.random((Nt, 784), dtype=np.float32)
    y = rng.integers(0, 10, Nt)
    eng = Engine(sp)
    Wd, xd = eng.device_array(W), eng.device_array(x)
    eng.predict_uncertainty(Wd, xd, y)
    eng.predict(Wd, xd)
    base = eng.info("last_device_ms")
    eng.predict_uncertainty(Wd, xd, y)
    ms = eng.info("last_device_ms")
    # CPU: the reference's double loop over draws x rows (Metrics.py:350-366) on a bounded sample, vectorised per draw
    probs = rng.random((4, Nt, 10))
    t0 = time.perf_counter()
    O.classification_uncertainty(probs, y, Nt)
    cpu = (time.perf_counter() - t0) / 4
    print(json.dumps({"case": "classification_uncertainty: 1000 weight samples x 10000x784 rows, 784-256-10",
                      "device_ms_total": ms, "device_ms_forward_only": base, "device_ms_uncertainty_part": ms - base,
                      "per_draw_outputs_bytes": n * Nt * 10 * 4,
                      "uncertainty_part_GBps_of_outputs_read": n * Nt * 10 * 4 / ((ms - base) / 1e3) / 1e9 if ms > base else None,
                      "cpu_baseline": {"value": cpu * n * 1e3, "unit": "ms for the 1000 draws (matrices only, no forward)",
                                       "cores": os.cpu_count(), "kind": "port", "sample": "4 draws, numpy fp64"}}),
          flush=True)
    eng.close()


if __name__ == "__main__":
    sg_chains(784, 128, 10, 60000, 1024, 1024)
    sg_chains(2, 50, 2, 1600, 128, 4096)
    uncertainty()
