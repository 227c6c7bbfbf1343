"""The reference's own arithmetic, when it can be had: ONNX Runtime CPU EP on the real kokoro ONNX file.

TEST INFRASTRUCTURE ONLY (same rule as oracle/kokoro_ref.py: only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this).

The reference computes a waveform by handing ``input_ids`` [1,N] i64, ``style`` [1,256] f32 and ``speed`` [1] f32 to
an ORT session built from the downloaded model file (/root/reference/kokorox/src/onn/ort_koko.rs:56-79,
ort_base.rs:27-33; model file kokorox/src/tts/koko.rs:57, utils/hf_cache.rs:8-10,135-144).  Neither ``onnxruntime``
nor a model file exists in the build container or on the GPU boxes (no network), so this leg normally reports
``absent`` -- it costs nothing then.  On a machine that has both it is the only thing that can pin parity to the
reference itself: SURVEY.md 8c / BASELINE.md section 2 promise it, and bench.py prints its result under ``"ort"``.

What is comparable against ORT: the graph has ONE output, the waveform; per-token durations are not
observable, only sum(dur) = len / 600.  The SineGen noise inside the graph cannot be injected, so waveform distances
are noise-limited: reported are sum(dur) equality, max-abs, relative L2 and a log-mel distance.
"""
from __future__ import annotations

import os
from typing import Dict, Optional, Tuple

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CANDIDATES = (
    lambda: os.environ.get("KKX_ONNX", ""),
    lambda: os.path.expanduser("~/.cache/huggingface/kokoro/v1.0/model.onnx"),      # hf_cache.rs:88-118
    lambda: os.path.join(ROOT, "checkpoints", "kokoro-v1.0.onnx"),                   # koko.rs:57 file name
    lambda: os.path.join(ROOT, "checkpoints", "model.onnx"),
)


def find_model() -> Optional[str]:
    for c in CANDIDATES:
        p = c()
        if p and os.path.isfile(p):
            return p
    return None


def available() -> Tuple[bool, str]:
    """(usable, reason).  Usable only when BOTH the onnxruntime package and a model file resolve."""
    try:
        import onnxruntime  # noqa: F401
    except Exception as e:  # noqa: BLE001
        return False, f"onnxruntime not importable ({type(e).__name__})"
    p = find_model()
    if p is None:
        return False, "no model file ($KKX_ONNX, ~/.cache/huggingface/kokoro/v1.0/model.onnx, checkpoints/kokoro-v1.0.onnx)"
    return True, p


class OrtReference:
    """ORT CPU EP session on the reference's model file, called with the reference's input names and shapes."""

    def __init__(self, model_path: Optional[str] = None, threads: Optional[int] = None):
        import onnxruntime as ort
        self.path = model_path or find_model()
        if not self.path:
            raise FileNotFoundError("no kokoro ONNX model file found")
        so = ort.SessionOptions()
        if threads:
            so.intra_op_num_threads = int(threads)
        self.sess = ort.InferenceSession(self.path, sess_options=so, providers=["CPUExecutionProvider"])
        names = [i.name for i in self.sess.get_inputs()]
        # ort_koko.rs:71-75 feeds "input_ids" (older exports: "tokens"), "style", "speed"
        self.tok_name = "input_ids" if "input_ids" in names else ("tokens" if "tokens" in names else names[0])
        self.threads = threads

    def infer(self, tokens, style, speed: float = 1.0) -> np.ndarray:
        feeds = {self.tok_name: np.asarray(tokens, np.int64).reshape(1, -1),
                 "style": np.asarray(style, np.float32).reshape(1, 256),
                 "speed": np.asarray([speed], np.float32)}
        out = self.sess.run(None, feeds)[0]
        return np.asarray(out, np.float32).reshape(-1)


def log_mel_distance(a: np.ndarray, b: np.ndarray, sr: int = 24000, n_fft: int = 1024, hop: int = 256,
                     n_mels: int = 80) -> float:
    """Mean absolute difference (dB) of log-mel spectrograms, time-aligned by truncation."""
    def mel_fb():
        f = np.linspace(0, sr / 2, n_fft // 2 + 1)
        m = np.linspace(0, 2595 * np.log10(1 + sr / 2 / 700), n_mels + 2)
        hz = 700 * (10 ** (m / 2595) - 1)
        fb = np.zeros((n_mels, len(f)))
        for i in range(n_mels):
            lo, c, hi = hz[i], hz[i + 1], hz[i + 2]
            fb[i] = np.clip(np.minimum((f - lo) / max(c - lo, 1e-9), (hi - f) / max(hi - c, 1e-9)), 0, None)
        return fb

    def spec(x):
        n = max(1 + (len(x) - n_fft) // hop, 1)
        x = np.pad(x, (0, max(0, n_fft - len(x))))
        fr = np.stack([x[i * hop:i * hop + n_fft] for i in range(n)]) * np.hanning(n_fft)
        p = np.abs(np.fft.rfft(fr, axis=1)) ** 2
        return 10 * np.log10(p @ mel_fb().T + 1e-10)
    sa, sb = spec(np.asarray(a, np.float64)), spec(np.asarray(b, np.float64))
    n = min(len(sa), len(sb))
    return float(np.abs(sa[:n] - sb[:n]).mean())


def compare(cuda_audio: np.ndarray, ort_audio: np.ndarray) -> Dict[str, float]:
    """The distances SURVEY 8c lists.  Equal sum(dur) <=> equal length (600 samples per frame)."""
    n = min(len(cuda_audio), len(ort_audio))
    a, b = np.asarray(cuda_audio[:n], np.float64), np.asarray(ort_audio[:n], np.float64)
    return {
        "frames_cuda": len(cuda_audio) // 600, "frames_ort": len(ort_audio) // 600,
        "sum_dur_equal": bool(len(cuda_audio) == len(ort_audio)),
        "max_abs": float(np.abs(a - b).max()) if n else 0.0,
        "rel_l2": float(np.sqrt(((a - b) ** 2).sum() / max((b ** 2).sum(), 1e-30))) if n else 0.0,
        "log_mel_db": log_mel_distance(cuda_audio, ort_audio),
    }
