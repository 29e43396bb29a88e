"""Chain / particle sharding across ranks (one process per GPU).

HMC chains never interact (the reference runs exactly one, HMC.py:37,66-67), so they shard with NO
data-path collective: rank r owns global chains [lo, hi) and passes `chain_offset=lo` to
`pyb_hmc_init` so the Philox counters — and therefore every trajectory — are those of the unsharded
run.  Only scalar diagnostics are reduced, and only when they are read.
"""
from __future__ import annotations


def shard_range(total: int, rank: int, world: int):
    """Contiguous, balanced split of `total` items: the first (total % world) ranks get one extra."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, extra = divmod(int(total), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def reduce_hmc_diag(diag: dict, dist=None, device=None) -> dict:
    """Combine per-rank `pyb_hmc_run` diagnostics: counts are summed, device time is the max over
    ranks, mean loss is weighted by the number of chain-iterations.  `dist` = torch.distributed
    (any backend) or None for a single process."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return dict(diag)
    import torch
    kw = {"device": device} if device is not None else {}
    sums = torch.tensor([diag["n_accepted"], diag["n_total"], diag["n_nan"], diag["grad_evals"],
                         diag["kernel_launches"], diag["mean_loss"] * diag["n_total"]], dtype=torch.float64, **kw)
    mx = torch.tensor([diag["device_ms"]], dtype=torch.float64, **kw)
    dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    s = sums.tolist()
    return {"n_accepted": int(s[0]), "n_total": int(s[1]), "n_nan": int(s[2]), "grad_evals": int(s[3]),
            "kernel_launches": int(s[4]), "mean_loss": s[5] / max(1.0, s[1]),
            "accept_rate": s[0] / max(1.0, s[1]), "device_ms": float(mx.item())}


def shard_weight_samples(W, weights, rank: int, world: int):
    """This rank's contiguous share of the weight samples of a posterior-predictive call (and of their integer
    multiplicities, if any): the partition `Engine.set_comm(..., predict_sharded=True)` + `Engine.predict` expects."""
    lo, hi = shard_range(len(W), rank, world)
    return W[lo:hi], (None if weights is None else weights[lo:hi])


def combine_predictive_moments(mean, var, wsum, dist=None):
    """Host-side equivalent of the device all-reduce, for ranks WITHOUT a shared NCCL communicator: each rank passes
    the mean / population variance / total weight of its own samples and gets the moments over all samples
    (sum of w, sum of w*o and sum of w*o^2 are additive).  `dist` = torch.distributed (any backend) or None."""
    import numpy as np
    mean, var = np.asarray(mean, np.float64), np.asarray(var, np.float64)
    s = np.stack([np.full_like(mean, float(wsum)), wsum * mean, wsum * (var + mean * mean)])
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        import torch
        t = torch.from_numpy(s)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        s = t.numpy()
    m = s[1] / s[0]
    return m.astype(np.float32), np.maximum(s[2] / s[0] - m * m, 0.0).astype(np.float32), float(s[0].flat[0])
