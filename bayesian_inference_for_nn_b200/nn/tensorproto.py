"""Hand-rolled TensorFlow ``TensorProto`` wire format — just enough for ``tf.io.serialize_tensor`` /
``tf.io.parse_tensor`` interoperability of the ``Sampled`` store (Sampled.py:45-48, :57-59), so that
posteriors saved by this build load in the reference and vice versa, without TensorFlow installed.

TensorProto:      1 dtype (varint)  2 tensor_shape (message)  4 tensor_content (bytes)
TensorShapeProto: 2 dim (repeated message)        Dim: 1 size (varint int64)
"""
import numpy as np

_DT = {np.dtype("float32"): 1, np.dtype("float64"): 2, np.dtype("int32"): 3, np.dtype("int64"): 9}
_DT_INV = {v: k for k, v in _DT.items()}


def _varint(n: int) -> bytes:
    n &= (1 << 64) - 1
    out = bytearray()
    while True:
        b = n & 0x7F
        n >>= 7
        out.append(b | (0x80 if n else 0))
        if not n:
            return bytes(out)


def _read_varint(buf, pos):
    shift = val = 0
    while True:
        b = buf[pos]
        pos += 1
        val |= (b & 0x7F) << shift
        if not b & 0x80:
            return val, pos
        shift += 7


def _field(num, wire, payload: bytes) -> bytes:
    key = _varint((num << 3) | wire)
    return key + (payload if wire == 0 else _varint(len(payload)) + payload)


def serialize_tensor(a) -> bytes:
    a = np.ascontiguousarray(a)
    if a.dtype not in _DT:
        raise TypeError("unsupported dtype %s" % a.dtype)
    shape = b"".join(_field(2, 2, _field(1, 0, _varint(int(d)))) for d in a.shape)
    return _field(1, 0, _varint(_DT[a.dtype])) + _field(2, 2, shape) + _field(4, 2, a.tobytes())


def _fields(buf):
    pos = 0
    while pos < len(buf):
        key, pos = _read_varint(buf, pos)
        num, wire = key >> 3, key & 7
        if wire == 0:
            val, pos = _read_varint(buf, pos)
        elif wire == 2:
            n, pos = _read_varint(buf, pos)
            val = buf[pos:pos + n]
            pos += n
        elif wire == 5:
            val = buf[pos:pos + 4]
            pos += 4
        elif wire == 1:
            val = buf[pos:pos + 8]
            pos += 8
        else:
            raise ValueError("unsupported wire type %d" % wire)
        yield num, wire, val


def parse_tensor(buf: bytes) -> np.ndarray:
    dtype, shape, content, floats = None, [], None, []
    for num, wire, val in _fields(bytes(buf)):
        if num == 1:
            dtype = _DT_INV[val]
        elif num == 2:
            for n2, _, dim in _fields(val):
                if n2 == 2:
                    size = 0
                    for n3, _, v in _fields(dim):
                        if n3 == 1:
                            size = v
                    shape.append(size)
        elif num == 4:
            content = val
        elif num == 5:   # float_val (packed or not) — small tensors written by other producers
            floats.append(val)
    if dtype is None:
        raise ValueError("TensorProto without dtype")
    if content is not None:
        return np.frombuffer(content, dtype=dtype).reshape(shape).copy()
    if floats:
        raw = b"".join(floats)
        arr = np.frombuffer(raw, dtype=np.float32)
        n = int(np.prod(shape)) if shape else 1
        if arr.size == 1 and n > 1:
            arr = np.full(n, arr[0], np.float32)
        return arr.reshape(shape).astype(dtype)
    return np.zeros(shape, dtype)
