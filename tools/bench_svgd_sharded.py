#!/usr/bin/env python
"""C4 (SVGD, 784-128-10, minibatch 1024 of a 60000-row pool) with the particles sharded over the GPUs of one box:
one process per GPU; the Stein phase is sharded over the parameters (gradient / particle all-to-all, all-reduced Gram
matrix).  The phase split is measured with CUDA events on the main stream: "wait_*" is what of an exchange or of
the side-stream reduction / median chain is NOT hidden behind compute.
    python tools/bench_svgd_sharded.py [--particles 4096] [--world 1,2,4,8] [--steps 5]
One JSON line per world size: device ms per step (max over ranks) and the speed-up over one GPU."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def worker(rank, world, uid, S_total, steps, out):
    from bayesian_inference_for_nn_b200 import _lib, keras_json
    from bayesian_inference_for_nn_b200.engine import Engine
    rng = np.random.default_rng(0)
    N, B = 60000, 1024
    X = rng.random((N, 784), dtype=np.float32)
    y = rng.integers(0, 10, N).astype(np.int32)
    idx = [rng.permutation(N)[:B].astype(np.int32) for _ in range(steps + 2)]
    sp = keras_json.parse_model_json(keras_json.make_sequential_json(784, [128, 10], ["relu", "softmax"]))
    eng = Engine(sp, device=rank, seed=1)
    eng.set_dataset(X, y, _lib.LOSS_SPARSE_CE)
    eng.set_prior([0.0], [1.0], _lib.PRIOR_SCALAR)
    Sl = S_total // world
    if world > 1:
        eng.svgd_set_comm(rank, world, uid)
        eng.set_option("svgd_p2p", int(os.environ.get("PYB_SVGD_P2P", "1")))
        eng.set_option("svgd_halves", int(os.environ.get("PYB_SVGD_HALVES", "0")))
        eng.set_option("svgd_gram_sync", int(os.environ.get("PYB_SVGD_GRAM_SYNC", "0")))
    eng.svgd_init(Sl, 0.01, _lib.SVGD_CANONICAL_MEDIAN, offset=rank * Sl)
    ms, loss, phases = [], [], []
    eng.set_option("profile", 1)
    for k, ix in enumerate(idx):
        loss.append(eng.svgd_step(ix))
        if k >= 2:
            ms.append(eng.info("last_device_ms"))
            phases.append([eng.info("svgd_phase_ms_%d" % j) for j in range(7)] if world > 1 else [0.0] * 7)
    out.put((rank, float(np.mean(ms)), loss[-1], np.mean(phases, axis=0).tolist(), int(eng.info("svgd_p2p")) if world > 1 else 0))
    eng.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--particles", type=int, default=4096)
    ap.add_argument("--world", default="1,2")
    ap.add_argument("--steps", type=int, default=5)
    a = ap.parse_args()
    import multiprocessing as mp
    from bayesian_inference_for_nn_b200 import _lib
    ctx = mp.get_context("spawn")
    base = None
    for world in [int(w) for w in a.world.split(",")]:
        if world > _lib.device_count():
            continue
        uid = _lib.nccl_unique_id() if world > 1 else None
        q = ctx.Queue()
        procs = [ctx.Process(target=worker, args=(r, world, uid, a.particles, a.steps, q)) for r in range(world)]
        for p in procs:
            p.start()
        res = [q.get(timeout=600) for _ in range(world)]
        for p in procs:
            p.join(timeout=60)
        ms = max(r[1] for r in res)
        base = base or ms
        print(json.dumps({"case": "C4 SVGD canonical_median 784-128-10, minibatch 1024", "particles": a.particles,
                          "n_gpus": world, "device_ms_per_step": ms, "particle_grad_evals_per_s": a.particles * 1e3 / ms,
                          "speedup_vs_1gpu": base / ms, "mean_loss_last": res[0][2],
                          "exchange": "peer-memory stores (own kernels)" if all(r[4] for r in res) else "nccl send/recv",
                          "phase_ms_max_over_ranks": dict(zip(["gram_partial", "gradients", "wait_reduced_kernel_matrix", "wait_gradient_all_to_all",
                                                               "ky_adam", "unpack_own_slice", "wait_particle_exchange"],
                                                              np.max([r[3] for r in res], axis=0).round(3).tolist()))}), flush=True)


if __name__ == "__main__":
    main()
