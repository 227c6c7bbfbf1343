"""A/B of the LSTM recurrence kernels on the shapes of a B = 64 x 510 step: libm gate functions (variant 0) against
SFU gate functions (variant 1).  Device time per launch from kkx_test_lstm_batch_v (CUDA events, `reps` launches)."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from kokorox_b200 import build as b  # noqa: E402

lib = C.CDLL(os.environ.get("KKX_LIB", os.path.join(b.OUT_DIR, "libkkx.so")))
lib.kkx_test_last_error.restype = C.c_char_p
fp = lambda a: a.ctypes.data_as(C.POINTER(C.c_float))  # noqa: E731
ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int))  # noqa: E731
rng = np.random.default_rng(0)
whh = (rng.standard_normal((2, 256, 1024)) * 0.06).astype(np.float32)
for name, B, lo, hi in (("tokens 64 x 510", 64, 510, 511), ("frames 40 x ~1200", 40, 1150, 1300), ("frames 24 x ~1200", 24, 1150, 1300),
                        ("8 x 510", 8, 510, 511), ("1 x 510", 1, 510, 511)):
    lens = rng.integers(lo, hi, B).astype(np.int32)
    offs = (np.concatenate([[0], np.cumsum(lens[:-1] + 32)]) + 32).astype(np.int32)
    rows = int(offs[-1] + lens[-1] + 32)
    xproj = (rng.standard_normal((rows, 2048)) * 0.7).astype(np.float32)
    res = {}
    for variant in (0, 1, 0, 1):
        out = np.zeros((rows, 512), np.float32)
        ms = C.c_float(0)
        rc = lib.kkx_test_lstm_batch_v(0, fp(xproj), fp(whh), B, ip(offs), ip(lens), rows, variant, 5, fp(out), C.byref(ms))
        assert rc == 0, lib.kkx_test_last_error()
        res.setdefault(variant, []).append((ms.value, out))
    d = res[0][0][1] - res[1][0][1]
    rel = float(np.sqrt((d ** 2).sum() / (res[0][0][1] ** 2).sum()))
    steps = int(lens.max())
    print(f"{name:20s} libm gates {min(t for t, _ in res[0]):7.3f} ms  SFU gates {min(t for t, _ in res[1]):7.3f} ms  "
          f"({min(t for t, _ in res[1]) * 1e6 / steps:6.0f} ns/step)  rel-L2 between them {rel:.2e}, max abs {np.abs(d).max():.2e}", flush=True)
