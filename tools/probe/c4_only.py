import sys, os
sys.path.insert(0, os.getcwd())
sys.path.insert(0, os.path.join(os.getcwd(), "tools"))
import bench_wide
from bayesian_inference_for_nn_b200 import _lib
_orig = bench_wide.svgd
def only_canonical(S_list):
    import numpy as np, json
    rng = np.random.default_rng(0)
    N, B = 60000, 1024
    sp = bench_wide.spec(784, 128, 10)
    X = rng.random((N, 784), dtype=np.float32)
    y = rng.integers(0, 10, N).astype(np.int32)
    for S in S_list:
        eng = bench_wide.Engine(sp, seed=1)
        eng.set_dataset(X, y, _lib.LOSS_SPARSE_CE)
        eng.set_prior([0.0], [1.0], _lib.PRIOR_SCALAR)
        eng.svgd_init(S, 0.01, _lib.SVGD_CANONICAL_MEDIAN)
        for _ in range(2):
            eng.svgd_step(rng.permutation(N)[:B].astype(np.int32))
            print(S, eng.info("last_device_ms"), flush=True)
        eng.close()
only_canonical([4096])
