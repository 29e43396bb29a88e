"""CPU checks of the C-ABI shared library: it loads, exports every symbol the header declares, and
fails loudly (no CPU fallback) when there is no GPU.  No compute calls are made here."""
import ctypes
import os
import re

import pytest

from conftest import ROOT

LIB = os.path.join(ROOT, "bayesian_inference_for_nn_b200", "libpyesian_b200.so")
HDR = os.path.join(ROOT, "include", "pyesian_b200.h")


def _declared():
    text = open(HDR).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pyb_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(LIB):
        from bayesian_inference_for_nn_b200.build import build
        build()
    return ctypes.CDLL(LIB)


def test_header_symbols_are_exported(lib):
    names = _declared()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), "missing export %s" % n


def test_ctypes_signatures_cover_the_header():
    from bayesian_inference_for_nn_b200 import _lib
    assert sorted(list(_lib.SIGNATURES) + ["pyb_last_error"]) == _declared()


def test_version_and_no_cpu_fallback(lib):
    from bayesian_inference_for_nn_b200 import _lib
    assert lib.pyb_version() == 1
    L = _lib.load()
    if _lib.device_count() > 0:
        pytest.skip("a GPU is present; the no-device failure mode is covered on CPU boxes")
    units = (ctypes.c_int32 * 1)(2)
    acts = (ctypes.c_int32 * 1)(0)
    bias = (ctypes.c_int32 * 1)(1)
    desc = _lib.ModelDesc(1, 2, units, acts, bias)
    h = ctypes.c_void_p()
    rc = L.pyb_create(ctypes.byref(desc), 0, 0, ctypes.byref(h))
    assert rc == -2  # PYB_ERR_CUDA
    assert b"no CPU fallback" in L.pyb_last_error()
    with pytest.raises(_lib.PyesianB200Error):
        _lib.check(rc)


def test_null_arguments_are_rejected(lib):
    from bayesian_inference_for_nn_b200 import _lib
    L = _lib.load()
    assert L.pyb_create(None, 0, 0, None) == -1
    assert L.pyb_param_count(None, None) == -1
    assert L.pyb_hmc_run(None, 1, 0, 1, None) == -1
    assert L.pyb_destroy(None) == 0


def test_product_code_never_touches_the_oracle_or_the_test_stand_ins():
    """the oracle, the goldens and the TensorFlow stand-in are test infrastructure: nothing under the package (or the
    Pyesian alias, bench.py's b200 arm aside) may import them, and there is no CPU fallback module to route through"""
    import ast
    banned = {"pyesian_oracle", "tf_shim", "oracle", "tests", "conftest"}
    for top in ("bayesian_inference_for_nn_b200", "Pyesian"):
        for dp, _, files in os.walk(os.path.join(ROOT, top)):
            for f in files:
                if not f.endswith(".py"):
                    continue
                tree = ast.parse(open(os.path.join(dp, f)).read())
                for node in ast.walk(tree):
                    names = []
                    if isinstance(node, ast.Import):
                        names = [a.name for a in node.names]
                    elif isinstance(node, ast.ImportFrom) and node.level == 0:
                        names = [node.module or ""]
                    for n in names:
                        assert n.split(".")[0] not in banned, (os.path.join(dp, f), n)
    srcs = [f for f in os.listdir(os.path.join(ROOT, "bayesian_inference_for_nn_b200", "csrc")) if f.endswith((".cu", ".cuh"))]
    assert len(srcs) >= 15 and not any("cpu" in f.lower() or "fallback" in f.lower() for f in srcs)
