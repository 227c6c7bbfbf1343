"""Minimal ONNX reader for checkpoint ingestion: pulls the weight tensors out of a
``kokoro-v1.0.onnx`` / ``model.onnx`` file (the file ``OrtKoko::new`` loads, ort_koko.rs:31-35, ort_base.rs:27-33)
without the ``onnx`` package -- just the protobuf wire format.

What it reads (field numbers from onnx.proto3): ``ModelProto.graph`` (7) -> ``GraphProto.initializer`` (5) and
``GraphProto.node`` (1); every ``TensorProto`` (dims 1, data_type 2, float_data 4, int32_data 5, int64_data 7,
name 8, raw_data 9, double_data 10) as a numpy array, fp16 / bf16 / double / int8 payloads included; ``Constant``
nodes' ``value`` tensors (exporters fold weight-norm into constants); and for each node its op type, name, inputs
and outputs, so that a caller can resolve anonymised initialisers (``onnx::MatMul_123``, ``onnx::LSTM_456``)
through the graph.  Helpers: ONNX ``LSTM`` weights ([dirs, 4H, *] in gate order i,o,f,c) -> PyTorch
``weight_ih_l0`` / ``weight_hh_l0`` / ``bias_*`` (+ ``_reverse``) in gate order i,f,g,o; ``dequantize`` for the
fp16 / int8 variants the reference can download (hf_cache.rs:135-144).

Status: the reader, the LSTM re-ordering and the name-preserving part of the mapping are tested on files written
by the small protobuf WRITER in tests/test_onnx_init.py (no ``.onnx`` exists in this environment, no network).
Resolving the anonymised MatMul / LSTM initialisers of the real export through its topology is NOT verified and
``convert_onnx`` reports those names instead of guessing.
"""
from __future__ import annotations

import struct
from collections import OrderedDict
from typing import Dict, List, Mapping, Tuple

import numpy as np

_DT = {1: np.float32, 2: np.uint8, 3: np.int8, 4: np.uint16, 5: np.int16, 6: np.int32, 7: np.int64, 9: np.bool_,
       10: np.float16, 11: np.float64, 12: np.uint32, 13: np.uint64}
_BF16 = 16


def _varint(buf: memoryview, p: int) -> Tuple[int, int]:
    v, shift = 0, 0
    while True:
        b = buf[p]
        p += 1
        v |= (b & 0x7F) << shift
        if b < 0x80:
            return v, p
        shift += 7
        if shift > 70:
            raise ValueError("malformed varint")


def fields(buf: memoryview):
    """Yield (field_number, wire_type, value) for one message; value is an int (varint / fixed) or a memoryview."""
    p, n = 0, len(buf)
    while p < n:
        key, p = _varint(buf, p)
        fno, wt = key >> 3, key & 7
        if wt == 0:
            v, p = _varint(buf, p)
        elif wt == 1:
            v = struct.unpack_from("<Q", buf, p)[0]
            p += 8
        elif wt == 2:
            ln, p = _varint(buf, p)
            if p + ln > n:
                raise ValueError("truncated length-delimited field")
            v = buf[p:p + ln]
            p += ln
        elif wt == 5:
            v = struct.unpack_from("<I", buf, p)[0]
            p += 4
        else:
            raise ValueError(f"unsupported wire type {wt}")
        yield fno, wt, v


def _packed_varints(v, wt) -> List[int]:
    if wt == 0:
        return [v]
    out, p = [], 0
    while p < len(v):
        x, p = _varint(v, p)
        out.append(x)
    return out


def _signed64(x: int) -> int:
    return x - (1 << 64) if x >= (1 << 63) else x


def parse_tensor(buf: memoryview) -> Tuple[str, np.ndarray]:
    dims: List[int] = []
    dtype, name, raw = 1, "", None
    f32: List[bytes] = []
    f64: List[bytes] = []
    i32: List[int] = []
    i64: List[int] = []
    external = False
    for fno, wt, v in fields(buf):
        if fno == 1:
            dims += [_signed64(x) for x in _packed_varints(v, wt)]
        elif fno == 2:
            dtype = v
        elif fno == 4:
            f32.append(bytes(v) if wt == 2 else struct.pack("<I", v))
        elif fno == 5:
            i32 += _packed_varints(v, wt)
        elif fno == 7:
            i64 += _packed_varints(v, wt)
        elif fno == 8:
            name = bytes(v).decode("utf-8")
        elif fno == 9:
            raw = bytes(v)
        elif fno == 10:
            f64.append(bytes(v) if wt == 2 else struct.pack("<Q", v))
        elif fno == 14 and v == 1:
            external = True
    if external:
        raise ValueError(f"tensor {name!r} uses external data, which this reader does not follow")
    shape = tuple(dims)
    if dtype == _BF16:
        src = np.frombuffer(raw, dtype="<u2") if raw is not None else np.asarray(i32, dtype=np.uint16)
        arr = (src.astype(np.uint32) << 16).view(np.float32)
    else:
        if dtype not in _DT:
            raise ValueError(f"tensor {name!r}: unsupported data_type {dtype}")
        npdt = np.dtype(_DT[dtype])
        if raw is not None:
            arr = np.frombuffer(raw, dtype=npdt.newbyteorder("<")).astype(npdt)
        elif f32:
            arr = np.frombuffer(b"".join(f32), dtype="<f4").astype(npdt)
        elif f64:
            arr = np.frombuffer(b"".join(f64), dtype="<f8").astype(npdt)
        elif i64:
            arr = np.asarray([_signed64(x) for x in i64], dtype=np.int64).astype(npdt)
        elif dtype == 10:       # fp16 stored as bit patterns in int32_data
            arr = np.asarray(i32, dtype=np.uint16).view(np.float16)
        else:
            arr = np.asarray([x - (1 << 32) if x >= (1 << 31) and npdt.kind == "i" else x for x in i32]).astype(npdt)
    n = int(np.prod(shape)) if shape else 1
    if arr.size != n:
        raise ValueError(f"tensor {name!r}: {arr.size} values for shape {shape}")
    return name, arr.reshape(shape)


class Node:
    __slots__ = ("op_type", "name", "inputs", "outputs", "tensors")

    def __init__(self):
        self.op_type, self.name = "", ""
        self.inputs: List[str] = []
        self.outputs: List[str] = []
        self.tensors: Dict[str, np.ndarray] = {}    # attribute name -> tensor (Constant.value)


def _parse_node(buf: memoryview) -> Node:
    nd = Node()
    for fno, wt, v in fields(buf):
        if fno == 1:
            nd.inputs.append(bytes(v).decode())
        elif fno == 2:
            nd.outputs.append(bytes(v).decode())
        elif fno == 3:
            nd.name = bytes(v).decode()
        elif fno == 4:
            nd.op_type = bytes(v).decode()
        elif fno == 5:                              # AttributeProto: name 1, t 5
            aname, t = "", None
            for f2, w2, v2 in fields(v):
                if f2 == 1:
                    aname = bytes(v2).decode()
                elif f2 == 5 and w2 == 2:
                    t = v2
            if t is not None:
                nd.tensors[aname] = parse_tensor(t)[1]
    return nd


def read_model(path: str) -> Tuple["OrderedDict[str, np.ndarray]", List[Node]]:
    """(initialisers + Constant values by output name, nodes in graph order)."""
    with open(path, "rb") as f:
        data = memoryview(f.read())
    graph = None
    for fno, wt, v in fields(data):
        if fno == 7 and wt == 2:
            graph = v
    if graph is None:
        raise ValueError(f"{path}: no GraphProto (field 7) -- not an ONNX model")
    tensors: "OrderedDict[str, np.ndarray]" = OrderedDict()
    nodes: List[Node] = []
    for fno, wt, v in fields(graph):
        if fno == 5 and wt == 2:
            name, arr = parse_tensor(v)
            tensors[name] = arr
        elif fno == 1 and wt == 2:
            nd = _parse_node(v)
            nodes.append(nd)
            if nd.op_type == "Constant" and "value" in nd.tensors and nd.outputs:
                tensors[nd.outputs[0]] = nd.tensors["value"]
    return tensors, nodes


def read_initializers(path: str) -> "OrderedDict[str, np.ndarray]":
    return read_model(path)[0]


def dequantize(arr: np.ndarray, scale=None, zero_point=None) -> np.ndarray:
    """fp16 / bf16 / double -> fp32; int8 / uint8 with (scale, zero_point) -> fp32 (DequantizeLinear)."""
    if arr.dtype in (np.int8, np.uint8) and scale is not None:
        zp = 0 if zero_point is None else np.asarray(zero_point, dtype=np.float32)
        return ((arr.astype(np.float32) - zp) * np.asarray(scale, dtype=np.float32)).astype(np.float32)
    return np.ascontiguousarray(arr, dtype=np.float32)


def onnx_lstm_to_torch(W: np.ndarray, R: np.ndarray, B: np.ndarray = None) -> Dict[str, np.ndarray]:
    """ONNX LSTM operands (W [D,4H,I], R [D,4H,H], B [D,8H]; gate order i,o,f,c) -> PyTorch nn.LSTM parameters
    (gate order i,f,g,o); direction 1 gets the ``_reverse`` suffix."""
    D, H4, _ = W.shape
    H = H4 // 4
    perm = np.concatenate([np.arange(0, H), np.arange(2 * H, 3 * H), np.arange(3 * H, 4 * H), np.arange(H, 2 * H)])
    out: Dict[str, np.ndarray] = {}
    for d in range(D):
        sfx = "_l0" + ("_reverse" if d == 1 else "")
        out["weight_ih" + sfx] = np.ascontiguousarray(W[d][perm], dtype=np.float32)
        out["weight_hh" + sfx] = np.ascontiguousarray(R[d][perm], dtype=np.float32)
        if B is not None:
            out["bias_ih" + sfx] = np.ascontiguousarray(B[d][:H4][perm], dtype=np.float32)
            out["bias_hh" + sfx] = np.ascontiguousarray(B[d][H4:][perm], dtype=np.float32)
        else:
            out["bias_ih" + sfx] = np.zeros(H4, np.float32)
            out["bias_hh" + sfx] = np.zeros(H4, np.float32)
    return out


def state_dict_from_onnx(path: str) -> Tuple[Dict[str, np.ndarray], List[str]]:
    """Everything that can be recovered by NAME: initialisers whose names are state-dict names are taken as they
    are (fp16 -> fp32); LSTM nodes whose W operand is a named ``...weight_ih_l0`` initialiser, or whose node name
    carries the module path (``/text_encoder/lstm/LSTM``), are converted back to PyTorch layout under that module.
    Returns (flat state dict, names of float initialisers it could not place)."""
    tensors, nodes = read_model(path)
    flat: Dict[str, np.ndarray] = {}
    used = set()
    for nd in nodes:
        if nd.op_type != "LSTM" or len(nd.inputs) < 3:
            continue
        W, R = tensors.get(nd.inputs[1]), tensors.get(nd.inputs[2])
        B = tensors.get(nd.inputs[3]) if len(nd.inputs) > 3 and nd.inputs[3] else None
        if W is None or R is None:
            continue
        mod = nd.name.strip("/").rsplit("/", 1)[0].replace("/", ".") if "/" in nd.name else ""
        if not mod:
            continue
        for k, v in onnx_lstm_to_torch(dequantize(W), dequantize(R), None if B is None else dequantize(B)).items():
            flat[f"{mod}.{k}"] = v
        used.update(nd.inputs[1:4])
    unplaced: List[str] = []
    for name, arr in tensors.items():
        if name in used or arr.dtype.kind != "f":
            continue
        if "::" in name or name.startswith("/") or name.isdigit():
            unplaced.append(name)
            continue
        flat[name] = dequantize(arr)
    return flat, unplaced


def convert_onnx(src: str, dst: str) -> int:
    """``.onnx`` -> KKXW when every tensor the backend needs can be recovered by name; otherwise raises with the
    list of what is missing and which anonymised initialisers were left over (see the module docstring)."""
    from .convert import convert_state_dict
    from .weightfile import write_weights
    flat, unplaced = state_dict_from_onnx(src)
    try:
        tensors = convert_state_dict(flat)
    except KeyError as e:
        raise KeyError(f"{e.args[0]}; {len(unplaced)} anonymised initialisers not resolved "
                       f"({', '.join(unplaced[:4])}{' ...' if len(unplaced) > 4 else ''})") from None
    write_weights(dst, tensors)
    return len(tensors)
