"""CPU tests: the C-ABI library builds, loads, exports every symbol include/*.h declares, and
fails loudly (no CPU fallback) when there is no GPU."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from kokorox_b200 import build
    path = build.build()
    return ctypes.CDLL(path)


def declared_symbols():
    syms = []
    for h in ("kkx.h", "kkx_test.h"):
        src = open(os.path.join(ROOT, "include", h)).read()
        syms += re.findall(r"KKX_API[^;(]*?\b(kkx_\w+)\s*\(", src)
    return sorted(set(syms))


def test_header_declares_the_boundary():
    syms = declared_symbols()
    for s in ("kkx_init", "kkx_create", "kkx_destroy", "kkx_last_error", "kkx_infer", "kkx_infer_batch",
              "kkx_release", "kkx_stage_batch", "kkx_run_staged", "kkx_fetch_staged"):
        assert s in syms


def test_library_exports_every_declared_symbol(lib):
    for s in declared_symbols():
        assert hasattr(lib, s), f"libkkx.so does not export {s}"


def test_no_torch_types_in_the_abi():
    src = open(os.path.join(ROOT, "include", "kkx.h")).read()
    assert "torch" not in src and "at::" not in src and "std::" not in src


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(_has_gpu(), reason="checks the no-GPU failure mode")
def test_fails_loudly_without_gpu(lib, tmp_path):
    lib.kkx_last_error.restype = ctypes.c_char_p
    lib.kkx_last_error.argtypes = [ctypes.c_void_p]
    assert lib.kkx_init() == -4
    ctx = ctypes.c_void_p()
    rc = lib.kkx_create(b"/nonexistent.kkxw", 0, ctypes.byref(ctx))
    assert rc == -4 and not ctx.value
    assert b"no CPU fallback" in lib.kkx_last_error(None)
    from kokorox_b200.onn import B200Koko, KkxError
    with pytest.raises(KkxError):
        B200Koko.new("/nonexistent.kkxw")


def test_python_shim_mirrors_reference_names():
    # OrtKoko::new / OrtKoko::infer / init_ort (ort_koko.rs:31-42, mod.rs:19-49)
    from kokorox_b200 import onn
    assert callable(onn.init_ort) and hasattr(onn.B200Koko, "new") and hasattr(onn.B200Koko, "infer")


def test_parse_style_name_mirrors_mix_styles():
    # TTSKoko::mix_styles parsing (koko.rs:1255-1295): single voice, weighted mixes (weight * 0.1 in f32, not
    # renormalised), skipped malformed parts, the reference's error cases
    import numpy as np
    import pytest
    from kokorox_b200.onn import KkxError, parse_style_name
    ids = {"af_sky": 0, "af_nicole": 1, "am_echo": 2}
    assert parse_style_name("af_sky", ids) == ([0], [np.float32(1.0)])
    v, p = parse_style_name("af_sky.4+af_nicole.5", ids)
    assert v == [0, 1] and p == [np.float32(4.0) * np.float32(0.1), np.float32(5.0) * np.float32(0.1)]
    v, p = parse_style_name("af_sky.4+am_echo", ids)              # part without a weight is skipped
    assert v == [0]
    v, p = parse_style_name("af_sky.x+am_echo.2", ids)            # non-numeric weight is skipped
    assert v == [2]
    for bad in ("nobody", "nobody.4+af_sky.5", "af_sky+am_echo"):
        with pytest.raises(KkxError):
            parse_style_name(bad, ids)


def test_output_containers_match_the_reference_encoders():
    # websocket lib.rs:696-736 (encode_audio: PCM16 WAV + base64) and utils/wav.rs:19-50 (streaming float WAV);
    # host-side byte work in libkkx, no GPU involved
    import base64
    import struct
    from kokorox_b200.onn import encode_audio, wav_header
    rng = np.random.default_rng(3)
    for n in (0, 1, 2, 3, 4, 5, 7, 600, 24001):
        pcm = rng.integers(-32768, 32768, size=n).astype(np.int16)
        data = pcm.astype("<i2").tobytes()
        want = (b"RIFF" + struct.pack("<I", 36 + len(data)) + b"WAVE" + b"fmt " +
                struct.pack("<IHHIIHH", 16, 1, 1, 24000, 24000 * 2, 2, 16) + b"data" + struct.pack("<I", len(data)) + data)
        got = encode_audio(pcm)
        assert got == base64.b64encode(want).decode(), n
        assert wav_header(n) == want[:44]
    h = wav_header(None, 24000, 1)
    assert h == (b"RIFF" + b"\xff" * 4 + b"WAVE" + b"fmt " + struct.pack("<IHHIIHH", 16, 3, 1, 24000, 24000 * 4, 4, 32) +
                 b"data" + b"\xff" * 4)
