"""TEST INFRASTRUCTURE (oracle side) -- dependency-free ONNX protobuf *reader*.

Independent of the product's C++ reader (``mimic3_b200/csrc/onnx_reader.cc``):
the oracle must not share parsing code with the thing it checks.  Field numbers
per onnx.proto.  Only what a VITS ``generator.onnx``
needs: initializers (FLOAT / INT64; raw_data, float_data, int64_data), node
list (op_type, inputs, outputs, name), Constant-node tensors.
"""
from __future__ import annotations

import struct
from typing import Dict, Iterator, List, Tuple

import numpy as np


def _read_varint(buf: memoryview, pos: int) -> Tuple[int, int]:
    result = 0
    shift = 0
    while True:
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not b & 0x80:
            return result, pos
        shift += 7


def _fields(buf: memoryview) -> Iterator[Tuple[int, int, object]]:
    pos = 0
    n = len(buf)
    while pos < n:
        key, pos = _read_varint(buf, pos)
        field, wire = key >> 3, key & 7
        if wire == 0:
            v, pos = _read_varint(buf, pos)
            yield field, wire, v
        elif wire == 1:
            yield field, wire, bytes(buf[pos:pos + 8])
            pos += 8
        elif wire == 2:
            ln, pos = _read_varint(buf, pos)
            yield field, wire, buf[pos:pos + ln]
            pos += ln
        elif wire == 5:
            yield field, wire, bytes(buf[pos:pos + 4])
            pos += 4
        else:
            raise ValueError(f"unsupported wire type {wire}")


def _signed(v: int) -> int:
    return v - (1 << 64) if v >= 1 << 63 else v


def parse_tensor(buf: memoryview) -> Tuple[str, np.ndarray]:
    dims: List[int] = []
    dtype = 0
    name = ""
    raw = None
    floats: List[bytes] = []
    int64s: List[int] = []
    for field, wire, v in _fields(buf):
        if field == 1:
            if wire == 0:
                dims.append(_signed(v))
            else:  # packed
                p = 0
                while p < len(v):
                    d, p = _read_varint(v, p)
                    dims.append(_signed(d))
        elif field == 2:
            dtype = v
        elif field == 4:
            floats.append(bytes(v) if wire == 2 else v)
        elif field == 7:
            if wire == 0:
                int64s.append(_signed(v))
            else:
                p = 0
                while p < len(v):
                    d, p = _read_varint(v, p)
                    int64s.append(_signed(d))
        elif field == 8:
            name = bytes(v).decode("utf-8")
        elif field == 9:
            raw = bytes(v)
    if dtype == 1:
        if raw is not None:
            arr = np.frombuffer(raw, dtype="<f4")
        else:
            arr = np.frombuffer(b"".join(floats), dtype="<f4")
    elif dtype == 7:
        if raw is not None:
            arr = np.frombuffer(raw, dtype="<i8")
        else:
            arr = np.array(int64s, dtype=np.int64)
    else:
        return name, None  # other dtypes are not needed for weights
    return name, arr.reshape(dims).copy()


class OnnxModel:
    def __init__(self) -> None:
        self.initializers: Dict[str, np.ndarray] = {}
        self.nodes: List[dict] = []
        self.inputs: List[str] = []
        self.outputs: List[str] = []


def _parse_node(buf: memoryview) -> dict:
    node = {"input": [], "output": [], "name": "", "op_type": "", "tensors": {}}
    for field, wire, v in _fields(buf):
        if field == 1:
            node["input"].append(bytes(v).decode())
        elif field == 2:
            node["output"].append(bytes(v).decode())
        elif field == 3:
            node["name"] = bytes(v).decode()
        elif field == 4:
            node["op_type"] = bytes(v).decode()
        elif field == 5:
            aname = ""
            t = None
            for f2, w2, v2 in _fields(v):
                if f2 == 1:
                    aname = bytes(v2).decode()
                elif f2 == 5:
                    t = parse_tensor(v2)[1]
            if t is not None:
                node["tensors"][aname] = t
    return node


def load(path: str) -> OnnxModel:
    data = memoryview(open(path, "rb").read())
    m = OnnxModel()
    for field, wire, v in _fields(data):
        if field != 7:
            continue
        for f2, w2, v2 in _fields(v):
            if f2 == 5:
                name, arr = parse_tensor(v2)
                if arr is not None:
                    m.initializers[name] = arr
            elif f2 == 1:
                m.nodes.append(_parse_node(v2))
            elif f2 in (11, 12):
                for f3, w3, v3 in _fields(v2):
                    if f3 == 1:
                        (m.inputs if f2 == 11 else m.outputs).append(bytes(v3).decode())
    return m


def named_parameters(path: str) -> Dict[str, np.ndarray]:
    """Initializers keyed by PyTorch module path, resolving the three export styles
    (plain names; ``weight_g``/``weight_v``; anonymous conv weights found through
    the Conv node that also consumes ``<module>.bias``)."""
    m = load(path)
    P = dict(m.initializers)
    for node in m.nodes:  # Constant nodes may hold folded tensors
        if node["op_type"] == "Constant" and "value" in node["tensors"] and node["output"]:
            P.setdefault(node["output"][0], node["tensors"]["value"])
    for name in list(P):
        if name.endswith(".weight_g"):
            base = name[: -len(".weight_g")]
            g = P[name].astype(np.float32)
            v = P[base + ".weight_v"].astype(np.float32)
            norm = np.sqrt((v.astype(np.float64) ** 2).sum(axis=tuple(range(1, v.ndim)), keepdims=True))
            P[base + ".weight"] = (g * (v / norm.astype(np.float32))).astype(np.float32)
    for node in m.nodes:
        if node["op_type"] in ("Conv", "ConvTranspose") and len(node["input"]) >= 3:
            w, b = node["input"][1], node["input"][2]
            if b.endswith(".bias") and w in P:
                P.setdefault(b[: -len(".bias")] + ".weight", P[w])
    return P
