"""Flat weight file ("KKXW") for the B200 Kokoro backend.

This is the checkpoint the C library loads (``kkx_create(weights_path, ...)``); it plays the role
that ``kokoro-v1.0.onnx`` plays for the reference's ``OrtKoko::new(model_path)``
(/root/reference/kokorox/src/onn/ort_koko.rs:31-35, ort_base.rs:14-39).

Layout (little endian):
    0   8 bytes  magic  b"KKXW0001"
    8   u32      n_tensors
    12  u32      header_bytes (offset of the data blob, 256-aligned)
    16  entries: u16 name_len, name bytes, u32 dtype (0 = f32), u32 ndim, u32 dims[ndim],
                 u64 offset (from blob start, 64-aligned), u64 nbytes
    blob

Tensor names follow the upstream Kokoro state-dict keys without the ``module.``
prefix and with weight-norm folded (``weight_g``/``weight_v`` -> ``weight``).
"""
from __future__ import annotations

import struct
from collections import OrderedDict
from typing import Dict, List, Tuple

import numpy as np

MAGIC = b"KKXW0001"


def write_weights(path: str, tensors: "OrderedDict[str, np.ndarray]") -> None:
    entries = []
    off = 0
    for name, arr in tensors.items():
        a = np.ascontiguousarray(arr, dtype=np.float32)
        entries.append((name.encode(), a, off))
        off += (a.nbytes + 63) // 64 * 64
    hdr = bytearray()
    for nm, a, o in entries:
        hdr += struct.pack("<H", len(nm)) + nm
        hdr += struct.pack("<II", 0, a.ndim)
        hdr += struct.pack("<%dI" % a.ndim, *a.shape)
        hdr += struct.pack("<QQ", o, a.nbytes)
    header_bytes = (16 + len(hdr) + 255) // 256 * 256
    with open(path, "wb") as f:
        f.write(MAGIC)
        f.write(struct.pack("<II", len(entries), header_bytes))
        f.write(hdr)
        f.write(b"\0" * (header_bytes - 16 - len(hdr)))
        for nm, a, o in entries:
            assert f.tell() == header_bytes + o
            f.write(a.tobytes())
            pad = (a.nbytes + 63) // 64 * 64 - a.nbytes
            if pad:
                f.write(b"\0" * pad)


def read_weights(path: str) -> "OrderedDict[str, np.ndarray]":
    with open(path, "rb") as f:
        buf = f.read()
    if buf[:8] != MAGIC:
        raise ValueError(f"{path}: not a KKXW0001 weight file")
    n, header_bytes = struct.unpack_from("<II", buf, 8)
    p = 16
    out: "OrderedDict[str, np.ndarray]" = OrderedDict()
    for _ in range(n):
        (ln,) = struct.unpack_from("<H", buf, p); p += 2
        name = buf[p:p + ln].decode(); p += ln
        dtype, ndim = struct.unpack_from("<II", buf, p); p += 8
        dims = struct.unpack_from("<%dI" % ndim, buf, p); p += 4 * ndim
        off, nbytes = struct.unpack_from("<QQ", buf, p); p += 16
        if dtype != 0:
            raise ValueError(f"{name}: unsupported dtype {dtype}")
        out[name] = np.frombuffer(buf, dtype=np.float32, count=nbytes // 4,
                                  offset=header_bytes + off).reshape(dims)
    return out


# --------------------------------------------------------------------------- architecture spec
def _lstm(prefix: str, n_in: int, H: int = 256) -> List[Tuple[str, tuple, str]]:
    out = []
    for sfx in ("", "_reverse"):
        out += [(f"{prefix}.weight_ih_l0{sfx}", (4 * H, n_in), "lstm"),
                (f"{prefix}.weight_hh_l0{sfx}", (4 * H, H), "lstm"),
                (f"{prefix}.bias_ih_l0{sfx}", (4 * H,), "lstm"),
                (f"{prefix}.bias_hh_l0{sfx}", (4 * H,), "lstm")]
    return out


def _blk(prefix: str, ci: int, co: int, up: bool = False):
    out = [(f"{prefix}.conv1.weight", (co, ci, 3), "wn_conv"), (f"{prefix}.conv1.bias", (co,), "bias"),
           (f"{prefix}.conv2.weight", (co, co, 3), "wn_conv"), (f"{prefix}.conv2.bias", (co,), "bias"),
           (f"{prefix}.norm1.fc.weight", (2 * ci, 128), "style_fc"), (f"{prefix}.norm1.fc.bias", (2 * ci,), "style_b"),
           (f"{prefix}.norm2.fc.weight", (2 * co, 128), "style_fc"), (f"{prefix}.norm2.fc.bias", (2 * co,), "style_b")]
    if ci != co:
        out.append((f"{prefix}.conv1x1.weight", (co, ci, 1), "wn_conv"))
    if up:
        out += [(f"{prefix}.pool.weight", (ci, 1, 3), "wn_pool"), (f"{prefix}.pool.bias", (ci,), "bias")]
    return out


def _arb(prefix: str, c: int, k: int):
    out = []
    for j in range(3):
        out += [(f"{prefix}.convs1.{j}.weight", (c, c, k), "wn_conv"), (f"{prefix}.convs1.{j}.bias", (c,), "bias"),
                (f"{prefix}.convs2.{j}.weight", (c, c, k), "wn_conv"), (f"{prefix}.convs2.{j}.bias", (c,), "bias"),
                (f"{prefix}.adain1.{j}.fc.weight", (2 * c, 128), "style_fc"), (f"{prefix}.adain1.{j}.fc.bias", (2 * c,), "style_b"),
                (f"{prefix}.adain2.{j}.fc.weight", (2 * c, 128), "style_fc"), (f"{prefix}.adain2.{j}.fc.bias", (2 * c,), "style_b"),
                (f"{prefix}.alpha1.{j}", (1, c, 1), "alpha"), (f"{prefix}.alpha2.{j}", (1, c, 1), "alpha")]
    return out


def weight_specs() -> List[Tuple[str, tuple, str]]:
    """(name, shape, kind) for every tensor of Kokoro-82M, folded."""
    s: List[Tuple[str, tuple, str]] = []
    L = "bert.encoder.albert_layer_groups.0.albert_layers.0."
    s += [("bert.embeddings.word_embeddings.weight", (178, 128), "emb"),
          ("bert.embeddings.position_embeddings.weight", (512, 128), "emb"),
          ("bert.embeddings.token_type_embeddings.weight", (2, 128), "emb"),
          ("bert.embeddings.LayerNorm.weight", (128,), "ln_w"), ("bert.embeddings.LayerNorm.bias", (128,), "ln_b"),
          ("bert.encoder.embedding_hidden_mapping_in.weight", (768, 128), "linear"),
          ("bert.encoder.embedding_hidden_mapping_in.bias", (768,), "bias")]
    for nm in ("query", "key", "value", "dense"):
        s += [(L + f"attention.{nm}.weight", (768, 768), "linear"), (L + f"attention.{nm}.bias", (768,), "bias")]
    s += [(L + "attention.LayerNorm.weight", (768,), "ln_w"), (L + "attention.LayerNorm.bias", (768,), "ln_b"),
          (L + "ffn.weight", (2048, 768), "linear"), (L + "ffn.bias", (2048,), "bias"),
          (L + "ffn_output.weight", (768, 2048), "linear"), (L + "ffn_output.bias", (768,), "bias"),
          (L + "full_layer_layer_norm.weight", (768,), "ln_w"), (L + "full_layer_layer_norm.bias", (768,), "ln_b"),
          ("bert.pooler.weight", (768, 768), "linear"), ("bert.pooler.bias", (768,), "bias"),
          ("bert_encoder.weight", (512, 768), "linear"), ("bert_encoder.bias", (512,), "bias")]
    # text encoder
    s.append(("text_encoder.embedding.weight", (178, 512), "emb"))
    for i in range(3):
        s += [(f"text_encoder.cnn.{i}.0.weight", (512, 512, 5), "wn_conv"), (f"text_encoder.cnn.{i}.0.bias", (512,), "bias"),
              (f"text_encoder.cnn.{i}.1.gamma", (512,), "ln_w"), (f"text_encoder.cnn.{i}.1.beta", (512,), "ln_b")]
    s += _lstm("text_encoder.lstm", 512)
    # predictor
    for i in range(3):
        s += _lstm(f"predictor.text_encoder.lstms.{2 * i}", 640)
        s += [(f"predictor.text_encoder.lstms.{2 * i + 1}.fc.weight", (1024, 128), "style_fc"),
              (f"predictor.text_encoder.lstms.{2 * i + 1}.fc.bias", (1024,), "style_b")]
    s += _lstm("predictor.lstm", 640)
    s += [("predictor.duration_proj.linear_layer.weight", (50, 512), "dur_w"),
          ("predictor.duration_proj.linear_layer.bias", (50,), "dur_b")]
    s += _lstm("predictor.shared", 640)
    for br in ("F0", "N"):
        s += _blk(f"predictor.{br}.0", 512, 512)
        s += _blk(f"predictor.{br}.1", 512, 256, up=True)
        s += _blk(f"predictor.{br}.2", 256, 256)
    s += [("predictor.F0_proj.weight", (1, 256, 1), "f0_proj_w"), ("predictor.F0_proj.bias", (1,), "f0_proj_b"),
          ("predictor.N_proj.weight", (1, 256, 1), "linear"), ("predictor.N_proj.bias", (1,), "bias")]
    # decoder
    s += _blk("decoder.encode", 514, 1024)
    for i in range(3):
        s += _blk(f"decoder.decode.{i}", 1090, 1024)
    s += _blk("decoder.decode.3", 1090, 512, up=True)
    s += [("decoder.F0_conv.weight", (1, 1, 3), "wn_conv"), ("decoder.F0_conv.bias", (1,), "bias"),
          ("decoder.N_conv.weight", (1, 1, 3), "wn_conv"), ("decoder.N_conv.bias", (1,), "bias"),
          ("decoder.asr_res.0.weight", (64, 512, 1), "wn_conv"), ("decoder.asr_res.0.bias", (64,), "bias")]
    G = "decoder.generator."
    s += [(G + "m_source.l_linear.weight", (1, 9), "linear"), (G + "m_source.l_linear.bias", (1,), "bias"),
          (G + "noise_convs.0.weight", (256, 22, 12), "conv"), (G + "noise_convs.0.bias", (256,), "bias"),
          (G + "noise_convs.1.weight", (128, 22, 1), "conv"), (G + "noise_convs.1.bias", (128,), "bias")]
    s += _arb(G + "noise_res.0", 256, 7)
    s += _arb(G + "noise_res.1", 128, 11)
    s += [(G + "ups.0.weight", (512, 256, 20), "wn_convT"), (G + "ups.0.bias", (256,), "bias"),
          (G + "ups.1.weight", (256, 128, 12), "wn_convT"), (G + "ups.1.bias", (128,), "bias")]
    for i, c in enumerate((256, 128)):
        for j, k in enumerate((3, 7, 11)):
            s += _arb(G + f"resblocks.{i * 3 + j}", c, k)
    s += [(G + "conv_post.weight", (22, 128, 7), "post_w"), (G + "conv_post.bias", (22,), "bias")]
    return s


def param_count_unfolded() -> Dict[str, int]:
    """Parameter totals per top-level group with weight-norm un-folded (g + v), for the
    self-check against the published totals."""
    tot: Dict[str, int] = {}
    for name, shape, kind in weight_specs():
        n = int(np.prod(shape))
        if kind in ("wn_conv", "wn_pool", "wn_convT", "post_w"):
            n += shape[0]          # weight_g has one entry per dim-0 slice
        g = name.split(".")[0]
        tot[g] = tot.get(g, 0) + n
    return tot


def random_weights(seed: int = 1234) -> "OrderedDict[str, np.ndarray]":
    """Deterministic random-init recipe -- used when no checkpoint is present.

    N(0, (g/sqrt(fan_in))^2) weights; non-trivial LayerNorm / Snake parameters so that channel
    indexing bugs show up; duration and F0 heads biased so that durations are ragged 1..8 frames
    per token and F0 crosses the voiced threshold.
    """
    rng = np.random.default_rng(seed)
    out: "OrderedDict[str, np.ndarray]" = OrderedDict()
    for name, shape, kind in weight_specs():
        fan_in = int(np.prod(shape[1:])) if len(shape) > 1 else 1
        if kind == "emb":
            a = rng.standard_normal(shape) * 0.5
        elif kind in ("linear", "conv", "wn_conv"):
            gain = 1.0
            # keep token identity alive through 12 shared ALBERT layers: peaky attention,
            # residual-dominated sub-layer outputs
            if "attention.query" in name or "attention.key" in name:
                gain = 2.0
            elif "attention.dense" in name:
                gain = 0.08
            elif "ffn_output" in name:
                gain = 0.2
            a = rng.standard_normal(shape) * (gain / math_sqrt(fan_in))
        elif kind == "wn_convT":
            # ConvTranspose1d [Ci,Co,k] stride s: each output sums Ci*k/s products
            s = 10 if shape[2] == 20 else 6
            a = rng.standard_normal(shape) * (1.0 / math_sqrt(shape[0] * shape[2] / s))
        elif kind == "wn_pool":
            a = 0.6 + 0.2 * rng.standard_normal(shape)
        elif kind == "post_w":
            a = rng.standard_normal(shape) * (0.12 / math_sqrt(fan_in))
        elif kind == "lstm":
            a = rng.uniform(-1.0, 1.0, shape) * (1.0 / 16.0)
            if "weight_ih" in name:
                a = rng.standard_normal(shape) * (1.0 / math_sqrt(fan_in))
        elif kind == "style_fc":
            a = rng.standard_normal(shape) * (0.3 / (0.15 * math_sqrt(128)))
        elif kind == "style_b":
            a = rng.standard_normal(shape) * 0.05
        elif kind == "bias":
            a = rng.standard_normal(shape) * (0.01 if "albert_layers" in name else 0.05)
        elif kind == "ln_w":
            # the ALBERT layer is applied 12x with the same LN: keep its affine close to identity
            a = 1.0 + (0.02 if "albert_layers" in name else 0.1) * rng.standard_normal(shape)
        elif kind == "ln_b":
            a = (0.01 if "albert_layers" in name else 0.1) * rng.standard_normal(shape)
        elif kind == "alpha":
            a = np.clip(1.0 + 0.3 * rng.standard_normal(shape), 0.3, 2.0)
        elif kind == "dur_w":
            # 50 independent directions with a large gain: a few of the 50 sigmoids sit near 0.5
            # for any given token, so the sum is ragged (1..6) while its utterance mean is stable
            a = rng.standard_normal(shape) * (12.0 / math_sqrt(fan_in))
        elif kind == "dur_b":
            a = -7.0 + 0.3 * rng.standard_normal(shape)
        elif kind == "f0_proj_w":
            a = rng.standard_normal(shape) * (60.0 / math_sqrt(fan_in))
        elif kind == "f0_proj_b":
            a = np.full(shape, 90.0)
        else:
            raise KeyError(kind)
        out[name] = a.astype(np.float32)
    return out


def math_sqrt(x: float) -> float:
    return float(np.sqrt(x))
