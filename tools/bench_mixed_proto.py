#!/usr/bin/env python
"""Per-launch device time of the 1-CTA split GEMM kernel in its two operand schemes on the same shape (one item per SM,
two 128-row tiles per item, N = 256, K = 3200): bf16x3 (6 MMA slots per 32 K-elements) against the prototype
fp16 + 2x e4m3 (4 slots).  Both move the same stage bytes; times come from the CUDA events the
library records around each launch ("profile" option).  Usage: python tools/bench_mixed_proto.py [reps=20] [--i8]      (--i8 adds the int8 two-slice prototype, item 4b)"""
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
reps = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 20
os.environ["PYB_DEBUG_GEMM_REPS"] = str(reps + 1)

from bayesian_inference_for_nn_b200 import _lib, keras_json  # noqa: E402
from bayesian_inference_for_nn_b200.engine import Engine  # noqa: E402
from test_gpu_tensor import mixed_operands  # noqa: E402  (host-side operand preparation of the prototype)

eng = Engine(keras_json.parse_model_json(keras_json.make_sequential_json(64, [32, 4], ["relu", "softmax"])))
sms = int(eng.info("sm_count"))
M, Nn, K, base = sms * 256, 256, 3200, 1024
rng = np.random.default_rng(0)
A0 = rng.standard_normal((base, K)).astype(np.float32)
B = (rng.standard_normal((Nn, K)) * 0.3).astype(np.float32)
tile = lambda x: np.ascontiguousarray(np.tile(x, ((M + base - 1) // base, 1))[:M])
lib = _lib.load()
out = {"shape": [M, Nn, K], "reps": reps, "sm_count": sms}
D = np.empty((M, Nn), np.float32)


def timed(call):
    eng.set_option("profile", 1)
    call()
    ms, n = eng.info("prof_ms"), eng.info("prof_launches")
    eng.set_option("profile", 0)
    return ms / n, int(n)


# bf16x3 on the 1-CTA kernel
A = tile(A0)
fn = lib.pyb_debug_tc_gemm
fn.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]
fn.restype = C.c_int
eng.set_option("tc_pair", 0)
ms, n = timed(lambda: _lib.check(fn(eng.h, A.ctypes.data, B.ctypes.data, M, Nn, K, D.ctypes.data)))
want = A0.astype(np.float64) @ B.astype(np.float64).T
rel = lambda d: float(np.linalg.norm(d[:base] - want) / np.linalg.norm(want))
out["bf16x3"] = {"ms_per_launch": ms, "launches": n, "tflops_algorithmic": 2.0 * M * Nn * K / ms / 1e9, "rel_err": rel(D)}
del A

# fp16 + 2x e4m3
a16, _, _, a8, sa = mixed_operands(A0)
b16, _, _, b8, sb = mixed_operands(B)
a16s = tile((a16.astype(np.float64) * 32.0).astype(np.float16).view(np.uint16))
b16s = np.ascontiguousarray((b16.astype(np.float64) * 64.0).astype(np.float16).view(np.uint16))
a8 = tile(a8)
fm = lib.pyb_debug_tc_gemm_mixed
fm.argtypes = [C.c_void_p] * 5 + [C.c_int32] * 3 + [C.c_float, C.c_void_p]
fm.restype = C.c_int
osc = np.float32(2.0 ** -11 / (sa * sb))
ms, n = timed(lambda: _lib.check(fm(eng.h, a16s.ctypes.data, a8.ctypes.data, b16s.ctypes.data, b8.ctypes.data, M, Nn, K,
                                    osc, D.ctypes.data)))
out["fp16_2xe4m3"] = {"ms_per_launch": ms, "launches": n, "tflops_algorithmic": 2.0 * M * Nn * K / ms / 1e9, "rel_err": rel(D)}
out["speedup"] = out["bf16x3"]["ms_per_launch"] / out["fp16_2xe4m3"]["ms_per_launch"]

if "--i8" in sys.argv:      # int8 two-slice prototype
    from test_gpu_tensor import int8_slices  # noqa: E402
    ah, al, s_a = int8_slices(A0)
    bh, bl, s_b = int8_slices(B)
    ah, al = tile(ah), tile(al)
    fi = lib.pyb_debug_tc_gemm_i8
    fi.argtypes = [C.c_void_p] * 5 + [C.c_int32] * 3 + [C.c_void_p]
    fi.restype = C.c_int
    ms, n = timed(lambda: _lib.check(fi(eng.h, ah.ctypes.data, al.ctypes.data, bh.ctypes.data, bl.ctypes.data, M, Nn, K,
                                        D.ctypes.data)))
    got = D[:base].astype(np.float64) * s_a * s_b.T / 127.0 ** 2
    out["int8_2slices"] = {"ms_per_launch": ms, "launches": n, "tflops_algorithmic": 2.0 * M * Nn * K / ms / 1e9,
                           "rel_err": float(np.linalg.norm(got - want) / np.linalg.norm(want))}
    out["speedup_int8"] = out["bf16x3"]["ms_per_launch"] / ms
print(json.dumps(out))
