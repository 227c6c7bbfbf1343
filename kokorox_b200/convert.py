"""Checkpoint ingestion: an upstream Kokoro-82M PyTorch state dict -> the KKXW file that
``kkx_create`` loads (the role ``kokoro-v1.0.onnx`` plays for ``OrtKoko::new``, ort_koko.rs:31-35).

    python -m kokorox_b200.convert kokoro-v1_0.pth kokoro.kkxw

Accepted input: the upstream ``kokoro-v1_0.pth`` layout -- a dict of five module state dicts (``bert``,
``bert_encoder``, ``predictor``, ``text_encoder``, ``decoder``) -- or one flat state dict; keys may carry a
``module.`` prefix (DataParallel); weight-norm may be stored old style (``weight_g`` / ``weight_v``) or as
parametrizations (``parametrizations.weight.original0`` / ``original1``).  Weight-norm is folded
(``w = g * v / ||v||``, the norm over every dim but 0, which is also what ``weight_norm(ConvTranspose1d)`` uses
upstream), every tensor the backend needs (``weightfile.weight_specs``) must be present with the expected shape;
unknown extra tensors (``position_ids``, pooler variants ...) are ignored.  No checkpoint exists in this
environment (no network): the round trip is tested on a synthetic un-folded state dict (tests/test_convert.py).
"""
from __future__ import annotations

import sys
from collections import OrderedDict
from typing import Dict, Mapping

import numpy as np

from .weightfile import weight_specs, write_weights

_TOP = ("bert", "bert_encoder", "predictor", "text_encoder", "decoder")


def _np(t) -> np.ndarray:
    if hasattr(t, "detach"):
        t = t.detach().cpu().float().numpy()
    return np.ascontiguousarray(np.asarray(t, dtype=np.float32))


def flatten_state_dict(sd: Mapping) -> Dict[str, np.ndarray]:
    """Nested {module: state_dict} or flat -> flat {name: array} without ``module.`` prefixes."""
    flat: Dict[str, np.ndarray] = {}
    nested = all(isinstance(v, Mapping) for v in sd.values()) and any(k in _TOP for k in sd)
    items = []
    if nested:
        for top, sub in sd.items():
            for k, v in sub.items():
                k = k[7:] if k.startswith("module.") else k
                items.append((f"{top}.{k}", v))
    else:
        for k, v in sd.items():
            items.append((k[7:] if k.startswith("module.") else k, v))
    for k, v in items:
        flat[k.replace(".module.", ".")] = _np(v)
    return flat


def fold_weight_norm(flat: Dict[str, np.ndarray]) -> Dict[str, np.ndarray]:
    out = dict(flat)
    pairs = []
    for k in flat:
        if k.endswith(".weight_g"):
            pairs.append((k[:-len("_g")], k, k[:-len("_g")] + "_v"))
        elif k.endswith(".parametrizations.weight.original0"):
            base = k[:-len(".parametrizations.weight.original0")]
            pairs.append((base + ".weight", k, base + ".parametrizations.weight.original1"))
    for wname, gk, vk in pairs:
        if vk not in flat:
            raise KeyError(f"weight-norm pair incomplete: {gk} without {vk}")
        g, v = flat[gk].astype(np.float64), flat[vk].astype(np.float64)
        axes = tuple(range(1, v.ndim))
        norm = np.sqrt((v * v).sum(axis=axes, keepdims=True))
        out[wname] = (v * (g.reshape(norm.shape) / norm)).astype(np.float32)
        out.pop(gk, None)
        out.pop(vk, None)
    return out


def convert_state_dict(sd: Mapping) -> "OrderedDict[str, np.ndarray]":
    flat = fold_weight_norm(flatten_state_dict(sd))
    tensors: "OrderedDict[str, np.ndarray]" = OrderedDict()
    missing, bad = [], []
    for name, shape, _kind in weight_specs():
        if name not in flat:
            missing.append(name)
            continue
        a = flat[name]
        if tuple(a.shape) != tuple(shape):
            if int(np.prod(a.shape)) == int(np.prod(shape)):
                a = a.reshape(shape)                      # e.g. alpha stored as [C] instead of [1,C,1]
            else:
                bad.append(f"{name}: {tuple(a.shape)} != {tuple(shape)}")
                continue
        tensors[name] = np.ascontiguousarray(a, dtype=np.float32)
    if missing or bad:
        raise KeyError("checkpoint does not match Kokoro-82M: missing %d tensors (%s%s)%s" % (
            len(missing), ", ".join(missing[:5]), " ..." if len(missing) > 5 else "",
            ("; shape mismatches: " + "; ".join(bad[:5])) if bad else ""))
    return tensors


def convert_file(src: str, dst: str) -> int:
    import torch
    sd = torch.load(src, map_location="cpu", weights_only=True)
    if isinstance(sd, Mapping) and "net" in sd and isinstance(sd["net"], Mapping):
        sd = sd["net"]                                     # StyleTTS2-style training checkpoints
    tensors = convert_state_dict(sd)
    write_weights(dst, tensors)
    return len(tensors)


if __name__ == "__main__":
    if len(sys.argv) != 3:
        sys.exit("usage: python -m kokorox_b200.convert <kokoro-v1_0.pth> <out.kkxw>")
    n = convert_file(sys.argv[1], sys.argv[2])
    print(f"wrote {n} tensors to {sys.argv[2]}")
