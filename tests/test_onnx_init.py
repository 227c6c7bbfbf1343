"""ONNX weight ingestion without the onnx package (kokorox_b200/onnx_init.py).  No .onnx file
exists in this environment, so the files are produced by the small protobuf WRITER below (field numbers from
onnx.proto3) -- an independent implementation of the wire format the reader parses."""
import struct

import numpy as np
import pytest

from kokorox_b200 import onnx_init
from kokorox_b200.weightfile import random_weights, read_weights, weight_specs


# ---- protobuf writer ---------------------------------------------------------------------------
def vint(x):
    x &= (1 << 64) - 1
    out = bytearray()
    while True:
        b = x & 0x7F
        x >>= 7
        out.append(b | (0x80 if x else 0))
        if not x:
            return bytes(out)


def fld(no, payload):            # length-delimited
    return vint(no << 3 | 2) + vint(len(payload)) + payload


def fvar(no, x):
    return vint(no << 3) + vint(x)


def tensor(name, arr, how="raw", dtype_code=None):
    codes = {np.dtype(np.float32): 1, np.dtype(np.int64): 7, np.dtype(np.float16): 10, np.dtype(np.float64): 11,
             np.dtype(np.int8): 3, np.dtype(np.int32): 6}
    msg = b""
    if how == "unpacked_dims":
        for d in arr.shape:
            msg += fvar(1, d)
    else:
        msg += fld(1, b"".join(vint(d) for d in arr.shape))
    msg += fvar(2, dtype_code or codes[arr.dtype])
    if how in ("raw", "unpacked_dims"):
        msg += fld(9, arr.astype(arr.dtype.newbyteorder("<")).tobytes())
    elif how == "float_data":
        msg += fld(4, arr.astype("<f4").tobytes())
    elif how == "int64_data":
        msg += fld(7, b"".join(vint(int(v)) for v in arr.reshape(-1)))
    elif how == "double_data":
        msg += fld(10, arr.astype("<f8").tobytes())
    elif how == "int32_data":
        msg += fld(5, b"".join(vint(int(v)) for v in arr.reshape(-1)))
    msg += fld(8, name.encode())
    return msg


def node(op, name, inputs, outputs, const=None):
    msg = b"".join(fld(1, i.encode()) for i in inputs) + b"".join(fld(2, o.encode()) for o in outputs)
    msg += fld(3, name.encode()) + fld(4, op.encode())
    if const is not None:
        msg += fld(5, fld(1, b"value") + fvar(20, 4) + fld(5, const))
    return msg


def model(initializers, nodes=()):
    graph = b"".join(fld(1, n) for n in nodes) + fld(2, b"main_graph") + b"".join(fld(5, t) for t in initializers)
    return fvar(1, 8) + fld(2, b"pytorch") + fld(7, graph) + fld(8, fld(1, b"") + fvar(2, 17))


# ---- tests -------------------------------------------------------------------------------------
def test_tensor_encodings(tmp_path):
    rng = np.random.default_rng(0)
    a = rng.standard_normal((3, 4, 5)).astype(np.float32)
    h = rng.standard_normal((7, 2)).astype(np.float16)
    i = np.array([[-5, 0, 2 ** 40], [1, -1, 7]], dtype=np.int64)
    d = rng.standard_normal(6)
    q = rng.integers(-128, 128, size=(4, 4)).astype(np.int8)
    bf = (a.view(np.uint32) >> 16).astype(np.uint16)
    inits = [tensor("a_raw", a), tensor("a_fd", a, "float_data"), tensor("a_unpacked", a, "unpacked_dims"),
             tensor("h", h), tensor("i_raw", i), tensor("i_data", i, "int64_data"), tensor("d", d, "double_data"),
             tensor("q", q), tensor("scalar", np.float32(2.5).reshape(())),
             tensor("bf", bf.view(np.float16), dtype_code=16),
             tensor("h_i32", h.view(np.uint16).astype(np.int32), "int32_data", dtype_code=10)]
    c = rng.standard_normal((2, 2)).astype(np.float32)
    p = tmp_path / "m.onnx"
    p.write_bytes(model(inits, [node("Constant", "/c", [], ["/c_output_0"], tensor("", c))]))
    t, nodes = onnx_init.read_model(str(p))
    for k in ("a_raw", "a_fd", "a_unpacked"):
        assert t[k].dtype == np.float32 and np.array_equal(t[k], a)
    assert t["h"].dtype == np.float16 and np.array_equal(t["h"], h)
    assert np.array_equal(t["h_i32"], h)
    assert np.array_equal(t["i_raw"], i) and np.array_equal(t["i_data"], i)
    assert np.array_equal(t["d"], d) and np.array_equal(t["q"], q)
    assert t["scalar"].shape == () and float(t["scalar"]) == 2.5
    assert np.array_equal(t["bf"], (bf.astype(np.uint32) << 16).view(np.float32))
    assert np.array_equal(t["/c_output_0"], c) and nodes[0].op_type == "Constant"
    assert np.array_equal(onnx_init.dequantize(q, 0.5, 3), (q.astype(np.float32) - 3) * 0.5)
    with pytest.raises(ValueError):
        onnx_init.read_model(__file__)


def to_onnx_lstm(sd, prefix, H=256):
    """PyTorch bi-LSTM parameters (gate order i,f,g,o) -> ONNX LSTM operands W, R, B (gate order i,o,f,c), the
    way torch.onnx lays them out."""
    def reorder(w):
        i, f, g, o = (w[k * H:(k + 1) * H] for k in range(4))
        return np.concatenate([i, o, f, g])
    W = np.stack([reorder(sd[f"{prefix}.weight_ih_l0{s}"]) for s in ("", "_reverse")])
    R = np.stack([reorder(sd[f"{prefix}.weight_hh_l0{s}"]) for s in ("", "_reverse")])
    B = np.stack([np.concatenate([reorder(sd[f"{prefix}.bias_ih_l0{s}"]), reorder(sd[f"{prefix}.bias_hh_l0{s}"])])
                  for s in ("", "_reverse")])
    return W, R, B


def test_lstm_gate_order_against_torch():
    # the re-ordered parameters must drive torch.nn.LSTM to the output of a hand-written ONNX-order LSTM cell
    import torch
    H, I, T = 8, 5, 6
    rng = np.random.default_rng(1)
    W = rng.standard_normal((1, 4 * H, I)).astype(np.float32)
    R = rng.standard_normal((1, 4 * H, H)).astype(np.float32) * 0.3
    B = rng.standard_normal((1, 8 * H)).astype(np.float32) * 0.1
    x = rng.standard_normal((T, I)).astype(np.float32)
    h, c, want = np.zeros(H), np.zeros(H), []
    sig = lambda z: 1 / (1 + np.exp(-z))
    for t in range(T):                                   # ONNX LSTM definition, gates i, o, f, c
        z = W[0] @ x[t] + R[0] @ h + B[0][:4 * H] + B[0][4 * H:]
        i, o, f, g = sig(z[:H]), sig(z[H:2 * H]), sig(z[2 * H:3 * H]), np.tanh(z[3 * H:])
        c = f * c + i * g
        h = o * np.tanh(c)
        want.append(h.copy())
    p = onnx_init.onnx_lstm_to_torch(W, R, B)
    lstm = torch.nn.LSTM(I, H)
    with torch.no_grad():
        for k, v in p.items():
            getattr(lstm, k).copy_(torch.from_numpy(v))
        got = lstm(torch.from_numpy(x)[:, None])[0][:, 0].numpy()
    np.testing.assert_allclose(got, np.stack(want), rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("half", [False, True])
def test_onnx_to_kkxw_round_trip(tmp_path, half):
    ref = random_weights(1234)
    kinds = {n: k for n, _s, k in weight_specs()}
    lstm_mods = sorted({n.rsplit(".", 1)[0] for n in ref if ".weight_ih_l0" in n and not n.endswith("_reverse")})
    enc = (lambda a: a.astype(np.float16)) if half else (lambda a: a)
    inits, nodes, skip = [], [], set()
    for j, mod in enumerate(lstm_mods):
        W, R, B = to_onnx_lstm(ref, mod)
        names = [f"onnx::LSTM_{9000 + 3 * j + k}" for k in range(3)]
        inits += [tensor(nm, enc(a)) for nm, a in zip(names, (W, R, B))]
        nodes.append(node("LSTM", "/" + mod.replace(".", "/") + "/LSTM", ["x"] + names, ["y"]))
        skip.update(n for n in ref if n.startswith(mod + ".") and "_l0" in n)
    inits += [tensor(n, enc(w)) for n, w in ref.items() if n not in skip]
    inits.append(tensor("onnx::MatMul_1", np.zeros((2, 2), np.float32)))        # an anonymised leftover
    inits.append(tensor("shape_const", np.array([1, -1], dtype=np.int64)))
    p = tmp_path / "kokoro.onnx"
    p.write_bytes(model(inits, nodes))
    flat, unplaced = onnx_init.state_dict_from_onnx(str(p))
    assert unplaced == ["onnx::MatMul_1"] and "shape_const" not in flat
    out = tmp_path / "k.kkxw"
    assert onnx_init.convert_onnx(str(p), str(out)) == len(ref)
    back = read_weights(str(out))
    for n, w in ref.items():
        if half:
            np.testing.assert_array_equal(back[n], w.astype(np.float16).astype(np.float32), err_msg=n)
        else:
            np.testing.assert_array_equal(back[n], w, err_msg=n)
    assert kinds  # (spec table loaded)


def test_unresolved_model_reports_what_is_missing(tmp_path):
    ref = random_weights(1234)
    inits = [tensor(n, w) for n, w in list(ref.items())[:10]] + [tensor("onnx::MatMul_77", np.zeros((3, 3), np.float32))]
    p = tmp_path / "partial.onnx"
    p.write_bytes(model(inits))
    with pytest.raises(KeyError, match="anonymised initialisers not resolved"):
        onnx_init.convert_onnx(str(p), str(tmp_path / "x.kkxw"))
