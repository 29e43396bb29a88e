#!/usr/bin/env python
"""Where the fused int8 forward kernel's time goes (option "tc_timeline"): the C3 workload (784-256-10, 60000 rows, 1024
chains), a few HMC iterations, then the per-item cycle sums the kernel recorded in its last launch:
  issuer (one thread per CTA pair): wait for the epilogue to drain TMEM | k loop | of which waiting for TMA fills
  epilogue warp 0: top of item | wait for accumulators | phase A (TMEM held) | logits exchange + loss + phase B"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bench
from bayesian_inference_for_nn_b200 import _lib, keras_json
from bayesian_inference_for_nn_b200.engine import Engine

S = int(os.environ.get("S", 1024))
X, y = bench.synth(60000)
spec = keras_json.parse_model_json(keras_json.make_sequential_json(784, [256, 10], ["relu", "softmax"]))
em = 1
for flags in [int(v) for v in os.environ.get("FLAGS", "0").split(",")]:
    eng = Engine(spec, device=0, seed=1234)
    eng.set_option("tc_epi_mma", em)
    eng.set_option("tc_i8", 2)
    eng.set_option("tc_timeline", 1 | flags)
    for kv in sys.argv[1:]:
        k, v = kv.split("=")
        eng.set_option(k, float(v))
    eng.set_dataset(X, y, _lib.LOSS_SPARSE_CE)
    eng.set_prior([0.0], [1.0], _lib.PRIOR_SCALAR)
    eng.hmc_init(S, 2e-5, 1.0, 20, _lib.HMC_REFERENCE)
    eng.hmc_run(2, burning=True, sampling=False)
    d = eng.hmc_run(1, burning=False, sampling=True)
    t = [eng.info("tc_timeline_%d" % k) for k in range(10)]
    print("flags=%d split=%d ms/iter %.0f | issuer: wait_tmem %.0f kloop %.0f (fill wait %.0f) items/cluster %.1f | "
          "epilogue: top %.0f wait_acc %.0f phaseA %.0f exchange %.0f softmax %.0f phaseB %.0f  (cycles per item)" %
          (flags, int(eng.info("tc_split")), d["device_ms"], t[0], t[1], t[2], t[3], t[4], t[5], t[6], t[8], t[9], t[7]), flush=True)
    eng.close()
