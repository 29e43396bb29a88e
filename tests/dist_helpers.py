"""torch.distributed collectives as the NumPy-in / NumPy-out callables `bayesian_inference_for_nn_b200.sharding` takes
(the package itself imports no tensor library)."""
import numpy as np


def torch_all_reduce(dist, op="sum", device=None):
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return None
    import torch

    def fn(a):
        t = torch.from_numpy(np.ascontiguousarray(a, np.float64))
        if device is not None:
            t = t.to(device)
        dist.all_reduce(t, op=dist.ReduceOp.SUM if op == "sum" else dist.ReduceOp.MAX)
        return t.cpu().numpy()
    return fn
