"""Several GPUs of one box behind ONE optimizer object.

The reference runs a single chain / a Python loop over particles on one device (HMC.py:45-72, SVGD.py:219-228).  Here
``HyperParameters(devices=[0, 1, ...])`` (or ``n_devices=k``) makes ``compile`` create one ``Engine`` per device in this
process, each driven by its own host thread (every C call releases the GIL, so the devices run concurrently):

* HMC chains never interact: device i owns the global chains [lo_i, hi_i) (``sharding.shard_range``) and is initialised
  with ``chain_offset=lo_i``, so the Philox counters — and therefore every trajectory and every sample — are those of the
  one-device run; ``result()`` pools the per-device samples in global chain order.
* SVGD particles: each engine joins one NCCL communicator (``pyb_svgd_set_comm``) and the exchange step runs inside the
  library (svgd.cu); the engines must hold the same number of particles.
"""
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from . import _lib
from .engine import Engine
from .sharding import shard_range


def devices_from(hp_get):
    """``devices`` (list of ordinals) or ``n_devices`` (first k devices) or ``device`` (one) -> list of ordinals"""
    devs = hp_get("devices", None)
    if devs is None:
        n = hp_get("n_devices", None)
        devs = list(range(int(n))) if n is not None else [int(hp_get("device", 0))]
    devs = [int(d) for d in devs]
    if len(devs) == 0 or len(set(devs)) != len(devs):
        raise ValueError("devices must be a non-empty list of distinct device ordinals")
    if len(devs) > 1:                       # (one device: pyb_create reports a bad ordinal / a missing GPU itself)
        have = _lib.device_count()
        if max(devs) >= have:
            raise ValueError("device %d requested, %d visible" % (max(devs), have))
    return devs


class EngineGroup:
    """One Engine per device; ``each(fn)`` runs ``fn(i, engine)`` on every engine concurrently and returns the results in
    device order (an exception in any thread is re-raised here)."""

    def __init__(self, spec, devices, seed=0):
        self.devices = list(devices)
        self.engines = [Engine(spec, device=d, seed=seed) for d in self.devices]
        self._pool = ThreadPoolExecutor(max_workers=len(self.engines)) if len(self.engines) > 1 else None

    def __len__(self):
        return len(self.engines)

    def each(self, fn):
        if self._pool is None:
            return [fn(0, self.engines[0])]
        futs = [self._pool.submit(fn, i, e) for i, e in enumerate(self.engines)]
        return [f.result() for f in futs]

    def close(self):
        for e in self.engines:
            e.close()
        if self._pool is not None:
            self._pool.shutdown(wait=True)
            self._pool = None

    # ---- HMC ------------------------------------------------------------------------------------------------------
    def hmc_init(self, n_chains, eps, m, L, semantics, q0=None, chain_offset=0):
        self.ranges = [shard_range(n_chains, i, len(self)) for i in range(len(self))]
        if any(hi == lo for lo, hi in self.ranges):
            raise ValueError("n_chains must be at least the number of devices")
        q0 = None if q0 is None else np.asarray(q0, np.float32)

        def init(i, e):
            lo, hi = self.ranges[i]
            q = None if q0 is None else (q0[lo:hi] if q0.shape[0] == n_chains else q0)
            e.hmc_init(hi - lo, eps, m, L, semantics, q0=q, chain_offset=chain_offset + lo)
        self.each(init)

    def hmc_run(self, n, burning, sampling):
        ds = self.each(lambda i, e: e.hmc_run(n, burning=burning, sampling=sampling))
        return merge_hmc_diag(ds)

    def hmc_samples(self):
        parts = self.each(lambda i, e: e.hmc_samples())
        return tuple(np.concatenate([p[k] for p in parts]) for k in range(3))

    def hmc_state(self):
        parts = self.each(lambda i, e: e.hmc_state())
        return np.concatenate([p[0] for p in parts]), np.concatenate([p[1] for p in parts])

    # ---- SVGD -----------------------------------------------------------------------------------------------------
    def svgd_join(self):
        """one NCCL communicator over the group's engines (a no-op for one device)"""
        if len(self) > 1:
            uid = _lib.nccl_unique_id()
            self.each(lambda i, e: e.svgd_set_comm(i, len(self), uid))


def merge_hmc_diag(ds):
    """per-device ``pyb_hmc_run`` diagnostics -> one: counts summed, mean loss weighted by chain-iterations, device time
    the maximum (the devices ran side by side)"""
    if len(ds) == 1:
        return ds[0]
    out = dict(ds[0])
    for k in ("n_accepted", "n_total", "n_nan", "grad_evals", "kernel_launches"):
        out[k] = sum(d[k] for d in ds)
    tot = max(1, out["n_total"])
    out["mean_loss"] = sum(d["mean_loss"] * d["n_total"] for d in ds) / tot
    out["accept_rate"] = out["n_accepted"] / tot
    out["device_ms"] = max(d["device_ms"] for d in ds)
    return out
