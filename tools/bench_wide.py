#!/usr/bin/env python
"""Wide-shape measurements beside the headline bench: BASELINE config C5 (posterior predictive,
n=1000 weight samples x 10000x784 test rows, 784-256-10) and config C4 (SVGD, 784-128-10, minibatch
1024 from a 60000-row pool) for a range of particle counts.  One JSON line per case."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bayesian_inference_for_nn_b200 import _lib, keras_json  # noqa: E402
from bayesian_inference_for_nn_b200.engine import Engine  # noqa: E402


def spec(D, H, C):
    return keras_json.parse_model_json(keras_json.make_sequential_json(D, [H, C], ["relu", "softmax"]))


def predictive():
    rng = np.random.default_rng(0)
    n, Nt = 1000, 10000
    sp = spec(784, 256, 10)
    W = (rng.standard_normal((n, sp.n_params)) * 0.05).astype(np.float32)
    x = rng.random((Nt, 784), dtype=np.float32)
    for path, name in ((_lib.PATH_AUTO, "tensor"), (_lib.PATH_GENERIC, "generic")):
        eng = Engine(sp)
        eng.set_option("path", path)
        eng.predict(W, x)                           # warm-up at full size (buffers grow once)
        t0 = time.perf_counter()
        eng.predict(W, x)
        wall = time.perf_counter() - t0
        ms = eng.info("last_device_ms")
        flops = 2.0 * Nt * (784 * 256 + 256 * 10) * n
        print(json.dumps({"case": "C5 predictive n=1000 x 10000x784, 784-256-10 (mean/var only)", "path": name,
                          "device_ms": ms, "wall_ms_incl_h2d_of_weights": 1e3 * wall, "samples_per_s": n / (ms / 1e3),
                          "rows_x_samples_per_s": n * Nt / (ms / 1e3), "algorithmic_tflops": flops / (ms / 1e3) / 1e12}),
              flush=True)
        if name == "tensor":
            try:                                    # the same call with the samples and inputs already resident in HBM
                import torch
                Wd, xd = torch.from_numpy(W).cuda(), torch.from_numpy(x).cuda()
                eng.predict(Wd, xd)
                eng.predict(Wd, xd)
                ms = eng.info("last_device_ms")
                print(json.dumps({"case": "C5 predictive n=1000 x 10000x784, 784-256-10 (mean/var only)",
                                  "path": "tensor, weights and inputs resident in HBM (DLPack)", "device_ms": ms,
                                  "samples_per_s": n / (ms / 1e3), "rows_x_samples_per_s": n * Nt / (ms / 1e3),
                                  "algorithmic_tflops": flops / (ms / 1e3) / 1e12}), flush=True)
            except ImportError:
                pass
        eng.close()


def svgd(S_list):
    rng = np.random.default_rng(0)
    N, B = 60000, 1024
    sp = spec(784, 128, 10)
    X = rng.random((N, 784), dtype=np.float32)
    y = rng.integers(0, 10, N).astype(np.int32)
    for sem, name in ((_lib.SVGD_CANONICAL_MEDIAN, "canonical_median"), (_lib.SVGD_REFERENCE_LIVE, "reference_live")):
        for S in S_list:
            if name == "reference_live" and S > 256:
                continue                          # the live sweep is sequential in S by construction
            eng = Engine(sp, seed=1)
            eng.set_dataset(X, y, _lib.LOSS_SPARSE_CE)
            eng.set_prior([0.0], [1.0], _lib.PRIOR_SCALAR)
            eng.svgd_init(S, 0.01, sem)
            idx = [rng.permutation(N)[:B].astype(np.int32) for _ in range(4)]
            eng.svgd_step(idx[0])
            ms = 0.0
            for ix in idx[1:]:
                eng.svgd_step(ix)
                ms += eng.info("last_device_ms")
            ms /= 3
            P = sp.n_params
            flops = S * 6.0 * B * (784 * 128 + 128 * 10) + (4.0 * S * S * P if name == "canonical_median" else 0.0)
            print(json.dumps({"case": "C4 SVGD 784-128-10, minibatch 1024 of 60000", "semantics": name, "particles": S,
                              "device_ms_per_step": ms, "steps_per_s": 1e3 / ms, "particle_grad_evals_per_s": S * 1e3 / ms,
                              "algorithmic_tflops": flops / (ms / 1e3) / 1e12, "grad_path": int(eng.info("path_used"))}),
                  flush=True)
            eng.close()


if __name__ == "__main__":
    predictive()
    svgd([64, 512, 4096] if "--full" in sys.argv else [64, 512])
