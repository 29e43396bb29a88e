"""Tensor ingestion for the C ABI: NumPy buffers, ``__dlpack__`` producers, legacy DLPack capsules
and ``tf.Tensor`` (through ``tf.experimental.dlpack.to_dlpack``, imported lazily — TensorFlow is
never a hard dependency).  Host tensors are passed as host pointers; CUDA tensors are passed as
device pointers (``PYB_MEM_DEVICE``) with zero copies on the Python side.

The reference keeps everything as eager ``tf.Tensor``/``tf.Variable`` objects (HMC.py:65,
SVGD.py:91-96); DLPack is the exchange format BASELINE.json's north_star names.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib

kDLCPU, kDLCUDA, kDLCUDAHost, kDLCUDAManaged = 1, 2, 3, 13
_DL_CODE = {0: "i", 1: "u", 2: "f"}


class _DLDevice(C.Structure):
    _fields_ = [("device_type", C.c_int32), ("device_id", C.c_int32)]


class _DLDataType(C.Structure):
    _fields_ = [("code", C.c_uint8), ("bits", C.c_uint8), ("lanes", C.c_uint16)]


class _DLTensor(C.Structure):
    _fields_ = [("data", C.c_void_p), ("device", _DLDevice), ("ndim", C.c_int32), ("dtype", _DLDataType),
                ("shape", C.POINTER(C.c_int64)), ("strides", C.POINTER(C.c_int64)), ("byte_offset", C.c_uint64)]


class _DLManagedTensor(C.Structure):
    _fields_ = [("dl_tensor", _DLTensor), ("manager_ctx", C.c_void_p), ("deleter", C.c_void_p)]


C.pythonapi.PyCapsule_GetPointer.restype = C.c_void_p
C.pythonapi.PyCapsule_GetPointer.argtypes = [C.py_object, C.c_char_p]
C.pythonapi.PyCapsule_IsValid.restype = C.c_int
C.pythonapi.PyCapsule_IsValid.argtypes = [C.py_object, C.c_char_p]


class DeviceView:
    """A borrowed view of a CUDA tensor: shape + raw device pointer.  Keeps the capsule (and so the
    producer's memory) alive; the capsule's own destructor runs the DLPack deleter."""

    def __init__(self, capsule, owner, shape, ptr, dtype):
        self._capsule, self._owner = capsule, owner
        self.shape, self.ptr, self.dtype = tuple(shape), ptr, dtype


def _is_capsule(obj):
    return type(obj).__name__ == "PyCapsule"


def _from_capsule(capsule, owner, want_dtype):
    if not C.pythonapi.PyCapsule_IsValid(capsule, b"dltensor"):
        raise ValueError("expected an unconsumed DLPack capsule named 'dltensor'")
    mt = C.cast(C.pythonapi.PyCapsule_GetPointer(capsule, b"dltensor"), C.POINTER(_DLManagedTensor)).contents
    t = mt.dl_tensor
    shape = [t.shape[i] for i in range(t.ndim)]
    kind = _DL_CODE.get(t.dtype.code)
    if kind is None or t.dtype.lanes != 1:
        raise TypeError("unsupported DLPack dtype code %d" % t.dtype.code)
    dt = np.dtype("%s%d" % (kind, t.dtype.bits // 8))
    if t.strides:
        exp = 1
        for i in range(t.ndim - 1, -1, -1):
            if shape[i] != 1 and t.strides[i] != exp:
                raise ValueError("DLPack tensor must be C-contiguous")
            exp *= shape[i]
    ptr = (t.data or 0) + t.byte_offset
    if t.device.device_type in (kDLCUDA, kDLCUDAManaged):
        if dt != np.dtype(want_dtype):
            raise TypeError("device tensor has dtype %s, the C ABI needs %s (cast it on the producer side)"
                            % (dt, np.dtype(want_dtype)))
        return DeviceView(capsule, owner, shape, ptr, dt), _lib.MEM_DEVICE, ptr
    if t.device.device_type in (kDLCPU, kDLCUDAHost):
        n = int(np.prod(shape)) if shape else 1
        buf = (C.c_char * (n * dt.itemsize)).from_address(ptr)
        # own copy: the capsule (and the producer's buffer) may be released right after this call
        arr = np.array(np.frombuffer(buf, dtype=dt).reshape(shape), dtype=want_dtype, copy=True, order="C")
        return arr, _lib.MEM_HOST, arr.ctypes.data
    raise TypeError("unsupported DLPack device type %d" % t.device.device_type)


def ingest(obj, want_dtype):
    """-> (holder with .shape, mem_kind, raw pointer).  The holder must stay referenced for the
    duration of the C call."""
    if isinstance(obj, np.ndarray) or isinstance(obj, (list, tuple)) or np.isscalar(obj):
        a = np.ascontiguousarray(obj, dtype=want_dtype)
        return a, _lib.MEM_HOST, a.ctypes.data
    if _is_capsule(obj):
        return _from_capsule(obj, None, want_dtype)
    if type(obj).__name__ == "DeviceArray" and hasattr(obj, "ptr"):      # engine.DeviceArray: already in HBM
        if np.dtype(obj.dtype) != np.dtype(want_dtype):
            raise TypeError("device array has dtype %s, expected %s" % (np.dtype(obj.dtype), np.dtype(want_dtype)))
        return obj, _lib.MEM_DEVICE, obj.ptr
    mod = type(obj).__module__ or ""
    if mod.startswith("tensorflow"):
        import tensorflow as tf  # lazy: only when the caller already handed us a tf.Tensor
        if obj.dtype != tf.as_dtype(np.dtype(want_dtype)):
            obj = tf.cast(obj, tf.as_dtype(np.dtype(want_dtype)))
        return _from_capsule(tf.experimental.dlpack.to_dlpack(obj), obj, want_dtype)
    if hasattr(obj, "__dlpack__"):
        dev = obj.__dlpack_device__() if hasattr(obj, "__dlpack_device__") else (kDLCPU, 0)
        if int(dev[0]) in (kDLCPU, kDLCUDAHost):
            try:
                a = np.from_dlpack(obj)
                a = np.ascontiguousarray(a, dtype=want_dtype)
                return a, _lib.MEM_HOST, a.ctypes.data
            except Exception:
                pass
        return _from_capsule(obj.__dlpack__(), obj, want_dtype)
    if hasattr(obj, "numpy"):
        a = np.ascontiguousarray(obj.numpy(), dtype=want_dtype)
        return a, _lib.MEM_HOST, a.ctypes.data
    a = np.ascontiguousarray(np.asarray(obj), dtype=want_dtype)
    return a, _lib.MEM_HOST, a.ctypes.data


def to_numpy(obj, dtype=None):
    """Host NumPy view/copy of anything array-like (used for labels, test inputs, results)."""
    if isinstance(obj, np.ndarray):
        return obj if dtype is None else obj.astype(dtype, copy=False)
    if hasattr(obj, "numpy"):
        a = obj.numpy()
    elif hasattr(obj, "__dlpack__") and not _is_capsule(obj):
        try:
            a = np.from_dlpack(obj)
        except Exception:
            a = np.asarray(obj)
    else:
        a = np.asarray(obj)
    return a if dtype is None else a.astype(dtype, copy=False)
