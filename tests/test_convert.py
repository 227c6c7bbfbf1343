"""Checkpoint ingestion (kokorox_b200/convert.py): an un-folded, prefixed, nested state dict in the
upstream kokoro-v1_0.pth layout must convert to exactly the tensors the backend loads."""
import numpy as np
import pytest

from kokorox_b200.convert import convert_state_dict
from kokorox_b200.weightfile import random_weights, read_weights, weight_specs, write_weights


def _unfold(w, rng, style):
    """weight -> (g, v) with an arbitrary positive rescaling of v, stored old style or as parametrizations."""
    v = (w.astype(np.float64) * rng.uniform(0.5, 2.0)).astype(np.float32)
    axes = tuple(range(1, w.ndim))
    g = np.sqrt((w.astype(np.float64) ** 2).sum(axis=axes, keepdims=True)).astype(np.float32)
    return g, v


@pytest.mark.parametrize("style", ["old", "parametrizations"])
def test_state_dict_round_trip(style, tmp_path):
    import torch
    rng = np.random.default_rng(0)
    ref = random_weights(1234)
    kinds = {n: k for n, _s, k in weight_specs()}
    nested = {}
    for name, w in ref.items():
        top, rest = name.split(".", 1)
        sub = nested.setdefault(top, {})
        key = "module." + rest
        if kinds[name] in ("wn_conv", "wn_pool", "wn_convT", "post_w"):
            g, v = _unfold(w, rng, style)
            if style == "old":
                sub[key + "_g"] = torch.from_numpy(g)
                sub[key + "_v"] = torch.from_numpy(v)
            else:
                base = key[:-len(".weight")]
                sub[base + ".parametrizations.weight.original0"] = torch.from_numpy(g)
                sub[base + ".parametrizations.weight.original1"] = torch.from_numpy(v)
        else:
            sub[key] = torch.from_numpy(w.copy())
    nested["bert"]["module.embeddings.position_ids"] = torch.arange(512)[None]     # an extra upstream buffer
    got = convert_state_dict(nested)
    assert list(got) == [n for n, _s, _k in weight_specs()]
    for name, w in ref.items():
        np.testing.assert_allclose(got[name], w, rtol=2e-6, atol=1e-7, err_msg=name)
    p = tmp_path / "w.kkxw"
    write_weights(str(p), got)
    back = read_weights(str(p))
    assert all(np.array_equal(back[n], got[n]) for n in got)


def test_missing_tensor_is_an_error():
    ref = dict(random_weights(1234))
    ref.pop("predictor.duration_proj.linear_layer.weight")
    with pytest.raises(KeyError):
        convert_state_dict(ref)
