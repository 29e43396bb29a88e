#!/usr/bin/env python
"""One pyb_predict_uncertainty call at the C5 shape (1000 weight samples x 10000x784 rows, 784-256-10) and three SWAG
steps (784-128-10, minibatch 1024, 1024 chains) for ncu: launch list
(`ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv`) or one
`--set full` capture (`-k regex:k_uncert_s2 -s 3 -c 1`, `-k regex:k_sg_update -s 1 -c 1`).  Run from the repo root."""
import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, "tools")
import numpy as np
from bayesian_inference_for_nn_b200 import _lib
from bayesian_inference_for_nn_b200.engine import Engine
import bench_next
rng = np.random.default_rng(0)
sp = bench_next.spec(784, 256, 10)
eng = Engine(sp)
W = eng.device_array((rng.standard_normal((1000, sp.n_params)) * 0.05).astype(np.float32))
x = eng.device_array(rng.random((10000, 784), dtype=np.float32))
y = rng.integers(0, 10, 10000)
eng.predict_uncertainty(W, x, y)
eng.predict_uncertainty(W, x, y)
sp2 = bench_next.spec(784, 128, 10)
e2 = Engine(sp2, seed=1)
X = rng.random((60000, 784), dtype=np.float32); yy = rng.integers(0, 10, 60000).astype(np.int32)
e2.set_dataset(X, yy, _lib.LOSS_SPARSE_CE)
e2.sg_init(1024, _lib.SG_SWAG, k_dev=20, frequency=1)
for i in range(3):
    e2.sg_step(1e-3, rng.permutation(60000)[:1024].astype(np.int32))
