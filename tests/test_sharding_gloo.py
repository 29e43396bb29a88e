"""World-size-2 CPU test (gloo) of the multi-rank host logic: chain ranges, diagnostic reduction
(counts summed, device time = max over ranks), and the oracle-level property the sharding relies on:
chains are independent and the RNG is keyed by the GLOBAL chain id, so shards reproduce the whole."""
import os
import socket

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402

from bayesian_inference_for_nn_b200.sharding import (combine_predictive_moments, reduce_hmc_diag, shard_range,  # noqa: E402
                                                     shard_weight_samples)


def test_shard_range_partitions():
    for total in (1, 7, 8, 1024, 1025):
        for world in (1, 2, 3, 8):
            spans = [shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    import pyesian_oracle as O
    S, P, it, seed = 6, 37, 3, 99
    lo, hi = shard_range(S, rank, world)
    # the device RNG restated: momenta of this rank's chains under their GLOBAL ids
    p_local = O.philox_normals(seed, np.arange(lo, hi), it, O.STREAM_MOMENTUM, P)
    gathered = [None] * world
    dist.all_gather_object(gathered, p_local)
    diag = {"n_accepted": 10 * (rank + 1), "n_total": 20, "n_nan": rank, "grad_evals": 100, "kernel_launches": 7,
            "mean_loss": 0.5 + rank, "accept_rate": 0.0, "device_ms": 3.0 + 2.0 * rank}
    from dist_helpers import torch_all_reduce
    red = reduce_hmc_diag(diag, torch_all_reduce(dist, "sum"), torch_all_reduce(dist, "max"))
    dist.barrier()
    if rank == 0:
        out.put((np.concatenate(gathered), red))
    dist.destroy_process_group()


def test_two_rank_sharding_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got, red = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    import pyesian_oracle as O
    whole = O.philox_normals(99, np.arange(6), 3, O.STREAM_MOMENTUM, 37)
    np.testing.assert_array_equal(got, whole)            # shards == unsharded stream
    assert red["n_accepted"] == 30 and red["n_total"] == 40 and red["n_nan"] == 1 and red["grad_evals"] == 200
    assert red["device_ms"] == 5.0                       # max over ranks, never the sum
    assert abs(red["mean_loss"] - 1.0) < 1e-12 and abs(red["accept_rate"] - 0.75) < 1e-12


def _pred_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    import pyesian_oracle as O
    spec = O.MLPSpec(3, [6, 4], ["relu", "softmax"])
    rng = np.random.default_rng(5)                      # every rank builds the same problem
    W = rng.normal(0, 0.5, (7, spec.n_params)).astype(np.float32)
    freq = rng.integers(1, 5, 7).astype(np.float32)
    x = rng.normal(size=(11, 3)).astype(np.float32)
    Wl, fl = shard_weight_samples(W, freq, rank, world)           # 4 + 3 samples
    mean_l, var_l = O.predictive(spec, Wl, x, weights=fl)         # what this rank's device call would return
    from dist_helpers import torch_all_reduce
    mean, var, wsum = combine_predictive_moments(mean_l, var_l, float(fl.sum()), torch_all_reduce(dist, "sum"))
    dist.barrier()
    if rank == 1:                                        # any rank holds the full answer
        out.put((mean, var, wsum))
    dist.destroy_process_group()


def test_sharded_predictive_moments_gloo():
    """weight samples split over 2 ranks: combining the per-rank moments reproduces the unsharded predictive"""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_pred_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    mean, var, wsum = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    import pyesian_oracle as O
    spec = O.MLPSpec(3, [6, 4], ["relu", "softmax"])
    rng = np.random.default_rng(5)
    W = rng.normal(0, 0.5, (7, spec.n_params)).astype(np.float32)
    freq = rng.integers(1, 5, 7).astype(np.float32)
    x = rng.normal(size=(11, 3)).astype(np.float32)
    want_mean, want_var = O.predictive(spec, W, x, weights=freq)
    assert wsum == float(freq.sum())
    np.testing.assert_allclose(mean, want_mean, rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(var, want_var, rtol=1e-4, atol=1e-7)
    # single process: identity
    m1, v1, w1 = combine_predictive_moments(want_mean, want_var, 3.0, None)
    np.testing.assert_allclose(m1, want_mean, rtol=1e-6)
    np.testing.assert_allclose(v1, want_var, rtol=1e-4, atol=1e-8)
