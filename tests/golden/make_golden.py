"""Generates tests/golden/*.npz from the oracle (run in the build container, CPU only).

The reference holds no golden audio or durations; these vectors pin the
ORACLE against itself across environments (dev container vs GPU box) and are the fixtures the
CUDA path is compared with.  Inputs: the reference's own example token sequences
(ort_koko.rs:46, tokenize.rs:124-126) and the cfg0 synthetic utterance.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests.conftest import (REF_EXAMPLE_IDS, REF_TOKENIZE_IDS, ensure_weights, make_noise,  # noqa: E402
                            synth_case)
from kokorox_b200.weightfile import read_weights  # noqa: E402
from oracle.kokoro_ref import KokoroOracle  # noqa: E402

KEEP = ("dur_float", "F0", "N", "har_source")


def main():
    o = KokoroOracle(read_weights(ensure_weights()))
    cases = {
        "ref_example": (np.asarray(REF_EXAMPLE_IDS, dtype=np.int64), synth_case(1, 0, 100)[1], 1.0),
        "ref_tokenize": (np.asarray(REF_TOKENIZE_IDS, dtype=np.int64), synth_case(1, 0, 101)[1], 1.2),
        "cfg0": (*synth_case(50, 0, 100), 1.0),
    }
    for name, (ids, style, speed) in cases.items():
        noise = make_noise(50 * len(ids), seed=7)
        r = o.forward(ids, style, speed, noise=noise, stages=True)
        out = {"tokens": ids, "style": style, "speed": np.float32(speed), "noise_seed": np.int64(7),
               "noise_frames_max": np.int64(50 * len(ids)), "pred_dur": r["pred_dur"].astype(np.int32),
               "audio": r["audio"].astype(np.float32)}
        for k in KEEP:
            out["stage." + k] = np.asarray(r["stages"][k], dtype=np.float32)
        out["stage.conv_post_rms"] = np.float32(np.sqrt(np.mean(r["stages"]["conv_post"] ** 2)))
        path = os.path.join(os.path.dirname(__file__), name + ".npz")
        np.savez_compressed(path, **out)
        print(name, "T =", r["T"], "dur =", r["pred_dur"][:12], "->", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
