"""Dependency-free ONNX protobuf *writer* (wire format only).

Used by :mod:`mimic3_b200.synth_voice` to emit a ``generator.onnx`` with the
initializer names / shapes of a Mimic 3 VITS voice so
that the C++ weight reader (``csrc/onnx_reader.cc``) and the oracle's Python
reader (``oracle/onnx_min.py``) can be exercised without the ``onnx`` package
(absent in this image) and without network access to MycroftAI/mimic3-voices.

Field numbers follow onnx.proto (ModelProto.graph=7, GraphProto.node=1,
initializer=5, input=11, output=12, TensorProto.dims=1/data_type=2/name=8/
raw_data=9, NodeProto.input=1/output=2/name=3/op_type=4/attribute=5).
"""
from __future__ import annotations

import struct
from typing import Iterable, List, Optional, Sequence, Tuple

import numpy as np

FLOAT = 1
INT64 = 7


def _varint(v: int) -> bytes:
    if v < 0:
        v += 1 << 64
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _tag(field: int, wire: int) -> bytes:
    return _varint((field << 3) | wire)


def f_varint(field: int, v: int) -> bytes:
    return _tag(field, 0) + _varint(v)


def f_bytes(field: int, b: bytes) -> bytes:
    return _tag(field, 2) + _varint(len(b)) + b


def f_str(field: int, s: str) -> bytes:
    return f_bytes(field, s.encode("utf-8"))


def f_float(field: int, v: float) -> bytes:
    return _tag(field, 5) + struct.pack("<f", v)


def tensor_proto(name: str, arr: np.ndarray, *, packed_dims: bool = False,
                 use_float_data: bool = False) -> bytes:
    """TensorProto. ``packed_dims`` / ``use_float_data`` exercise the alternative
    encodings a real exporter may choose (readers must accept both)."""
    arr = np.ascontiguousarray(arr)
    if arr.dtype == np.float32:
        dt = FLOAT
    elif arr.dtype == np.int64:
        dt = INT64
    else:
        raise TypeError(f"unsupported dtype {arr.dtype}")
    out = bytearray()
    if packed_dims and arr.ndim:
        out += f_bytes(1, b"".join(_varint(int(d)) for d in arr.shape))
    else:
        for d in arr.shape:
            out += f_varint(1, int(d))
    out += f_varint(2, dt)
    if use_float_data and dt == FLOAT:
        out += f_bytes(4, arr.astype("<f4").tobytes())  # packed repeated float
    elif use_float_data and dt == INT64:
        out += f_bytes(7, b"".join(_varint(int(v)) for v in arr.reshape(-1)))
    out += f_str(8, name)
    if not use_float_data:
        out += f_bytes(9, arr.astype(arr.dtype.newbyteorder("<")).tobytes())
    return bytes(out)


def attr_ints(name: str, vals: Sequence[int]) -> bytes:
    out = f_str(1, name)
    for v in vals:
        out += f_varint(8, int(v))
    out += f_varint(20, 7)  # AttributeType.INTS
    return out


def attr_int(name: str, v: int) -> bytes:
    return f_str(1, name) + f_varint(3, int(v)) + f_varint(20, 2)


def attr_float(name: str, v: float) -> bytes:
    return f_str(1, name) + f_float(2, v) + f_varint(20, 1)


def attr_tensor(name: str, tensor: bytes) -> bytes:
    return f_str(1, name) + f_bytes(5, tensor) + f_varint(20, 4)


def node_proto(op_type: str, inputs: Iterable[str], outputs: Iterable[str],
               name: str = "", attrs: Optional[List[bytes]] = None) -> bytes:
    out = bytearray()
    for i in inputs:
        out += f_str(1, i)
    for o in outputs:
        out += f_str(2, o)
    if name:
        out += f_str(3, name)
    out += f_str(4, op_type)
    for a in attrs or []:
        out += f_bytes(5, a)
    return bytes(out)


def value_info(name: str, elem_type: int, dims: Sequence[object]) -> bytes:
    shape = bytearray()
    for d in dims:
        if isinstance(d, str):
            shape += f_bytes(1, f_str(2, d))
        else:
            shape += f_bytes(1, f_varint(1, int(d)))
    tensor_type = f_varint(1, elem_type) + f_bytes(2, bytes(shape))
    type_proto = f_bytes(1, tensor_type)
    return f_str(1, name) + f_bytes(2, type_proto)


def model_proto(nodes: List[bytes], initializers: List[bytes],
                inputs: List[bytes], outputs: List[bytes],
                producer: str = "mimic3_b200.synth_voice", opset: int = 15,
                graph_name: str = "torch-jit-export") -> bytes:
    g = bytearray()
    for n in nodes:
        g += f_bytes(1, n)
    g += f_str(2, graph_name)
    for t in initializers:
        g += f_bytes(5, t)
    for i in inputs:
        g += f_bytes(11, i)
    for o in outputs:
        g += f_bytes(12, o)
    m = bytearray()
    m += f_varint(1, 7)  # ir_version
    m += f_str(2, producer)
    m += f_bytes(7, bytes(g))
    m += f_bytes(8, f_str(1, "") + f_varint(2, opset))
    return bytes(m)
