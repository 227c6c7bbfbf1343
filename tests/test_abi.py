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


def test_every_session_option_is_documented():
    """Every key kkx_set_option accepts (csrc/capi.cu) is described in the public header and in INTEGRATION.md's option
    table: a maintainer must be able to find what a switch does without reading the library."""
    capi = open(os.path.join(ROOT, "kokorox_b200", "csrc", "capi.cu")).read()
    body = capi[capi.index("KKX_API int kkx_set_option"):]
    body = body[:body.index("KKX_API", 10)]
    keys = set(re.findall(r'k == "(\w+)"', body))
    assert {"precision", "coalesce", "max_frames", "gemm_pair", "lstm_fast_gates", "ups_phase_loop"} <= keys
    header = open(os.path.join(ROOT, "include", "kkx.h")).read()
    integ = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    missing_h = sorted(k for k in keys if f'"{k}"' not in header)
    missing_i = sorted(k for k in keys if f"`{k}`" not in integ)
    assert not missing_h, f"not described in include/kkx.h: {missing_h}"
    assert not missing_i, f"not in INTEGRATION.md's option table: {missing_i}"


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


@pytest.mark.parametrize("header", ["kkx.h", "kkx_test.h"])
def test_headers_are_plain_c(header):
    # the boundary is bound from Rust FFI / cgo / ctypes: the headers must be valid C99 (and C++11) on their own
    import shutil
    import subprocess
    if not shutil.which("gcc"):
        pytest.skip("no gcc")
    inc = os.path.join(ROOT, "include")
    for comp, lang, std in (("gcc", "c", "-std=c99"), ("g++", "c++", "-std=c++11")):
        if not shutil.which(comp):
            continue
        r = subprocess.run([comp, std, "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-I", inc, "-x", lang, "-"],
                           input=f'#include "{header}"\n', text=True, capture_output=True)
        assert r.returncode == 0, r.stderr


def test_c_program_links_against_the_library(lib, tmp_path):
    # a C translation unit that uses only include/kkx.h builds and links against libkkx.so, calls the host-side
    # entry points and gets the documented "no device / bad argument" codes (no compute without a GPU)
    import shutil
    import subprocess
    if not shutil.which("gcc"):
        pytest.skip("no gcc")
    from kokorox_b200 import build
    so = build.build()
    src = tmp_path / "t.c"
    src.write_text(r"""
#include <stdio.h>
#include <string.h>
#include "kkx.h"
int main(void) {
  unsigned char h[44];
  char b64[128];
  int16_t pcm[3] = {0, 1, -1};
  if (kkx_wav_header_pcm16(h, 3, 24000) != 44 || memcmp(h, "RIFF", 4) != 0) return 1;
  if (kkx_wav_header_f32_stream(h, 1, 24000) != 44 || h[20] != 3) return 2;
  if (kkx_encode_wav16_base64(pcm, 3, 24000, NULL, 0) != 68) return 3;          /* (44 + 6 + 2) / 3 * 4 */
  if (kkx_encode_wav16_base64(pcm, 3, 24000, b64, sizeof b64) != 68 || strlen(b64) != 68) return 4;
  if (kkx_infer(NULL, NULL, 0, NULL, 1.0f, NULL, NULL, NULL) >= 0) return 5;     /* null ctx -> error code */
  printf("%s\n", kkx_version());
  return 0;
}
""")
    exe = tmp_path / "t"
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                        so, "-Wl,-rpath," + os.path.dirname(so)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.startswith("kkx "), (r.returncode, r.stdout, r.stderr)
