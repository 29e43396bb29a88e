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


def reduce_hmc_diag(diag: dict, all_reduce_sum=None, all_reduce_max=None) -> dict:
    """Combine per-rank `pyb_hmc_run` diagnostics: counts are summed, device time is the max over ranks, mean loss is
    weighted by the number of chain-iterations.  The caller brings its collective — `all_reduce_sum(a)` /
    `all_reduce_max(a)` take and return a float64 NumPy array (tests/dist_helpers.py wraps `torch.distributed`; the
    package itself depends on no tensor library) — or nothing for a single process."""
    import numpy as np
    if all_reduce_sum is None:
        return dict(diag)
    s = np.asarray(all_reduce_sum(np.array([diag["n_accepted"], diag["n_total"], diag["n_nan"], diag["grad_evals"],
                                            diag["kernel_launches"], diag["mean_loss"] * diag["n_total"]], np.float64)))
    mx = np.asarray((all_reduce_max or all_reduce_sum)(np.array([diag["device_ms"]], np.float64)))
    return {"n_accepted": int(s[0]), "n_total": int(s[1]), "n_nan": int(s[2]), "grad_evals": int(s[3]),
            "kernel_launches": int(s[4]), "mean_loss": float(s[5]) / max(1.0, float(s[1])),
            "accept_rate": float(s[0]) / max(1.0, float(s[1])), "device_ms": float(mx[0])}


def shard_weight_samples(W, weights, rank: int, world: int):
    """This rank's contiguous share of the weight samples of a posterior-predictive call (and of their integer
    multiplicities, if any): the partition `Engine.set_comm(..., predict_sharded=True)` + `Engine.predict` expects."""
    lo, hi = shard_range(len(W), rank, world)
    return W[lo:hi], (None if weights is None else weights[lo:hi])


def combine_predictive_moments(mean, var, wsum, all_reduce_sum=None):
    """Host-side equivalent of the device all-reduce, for ranks WITHOUT a shared NCCL communicator: each rank passes
    the mean / population variance / total weight of its own samples and gets the moments over all samples
    (sum of w, sum of w*o and sum of w*o^2 are additive).  `all_reduce_sum(a)`: float64 NumPy array in and out (see
    `reduce_hmc_diag`), or None for a single process."""
    import numpy as np
    mean, var = np.asarray(mean, np.float64), np.asarray(var, np.float64)
    s = np.stack([np.full_like(mean, float(wsum)), wsum * mean, wsum * (var + mean * mean)])
    if all_reduce_sum is not None:
        s = np.asarray(all_reduce_sum(s))
    m = s[1] / s[0]
    return m.astype(np.float32), np.maximum(s[2] / s[0] - m * m, 0.0).astype(np.float32), float(s[0].flat[0])
