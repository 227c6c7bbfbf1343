"""Synthetic inputs and the random-init checkpoint shared by the bench, the smoke test, the tools and the
test-suite.  No checkpoint ships with the reference and there is no network, so the CUDA path and
the CPU oracle both load the deterministic random-init file this module writes on first use."""
from __future__ import annotations

import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WEIGHTS = os.environ.get("KKX_WEIGHTS", os.path.join(ROOT, "weights", "kokoro_random_1234.kkxw"))

# ort_koko.rs:46 -- the example input in the reference's own comment
REF_EXAMPLE_IDS = [0, 56, 51, 142, 156, 69, 63, 3, 16, 61, 4, 16, 156, 51, 4, 16, 62, 77, 156, 51, 86, 5, 0]
# tokenize.rs:124-126 -- "$h@l'oU, w'3:ld!$"
REF_TOKENIZE_IDS = [0, 50, 83, 54, 156, 57, 135, 3, 16, 65, 156, 87, 158, 54, 46, 5, 0]


def ensure_weights(path: str = WEIGHTS, seed: int = 1234) -> str:
    """Random-init Kokoro-82M weights (SURVEY.md 8d recipe); written once, atomically (several ranks may race)."""
    from .weightfile import random_weights, write_weights
    if not os.path.exists(path):
        os.makedirs(os.path.dirname(path), exist_ok=True)
        tmp = path + ".tmp%d" % os.getpid()
        write_weights(tmp, random_weights(seed))
        os.replace(tmp, path)
    return path


def ensure_second_weights() -> str:
    """A second, different weight set (seed 4321) standing in for the v1.1-zh model of BASELINE configs[4]:
    same architecture, different parameters, its own device-resident WeightSet (TTSManager keeps one TTSKoko per
    language, koko.rs:66-143)."""
    return ensure_weights(os.path.join(os.path.dirname(WEIGHTS), "kokoro_random_4321.kkxw"), seed=4321)


def synth_case(n_tokens: int, seed: int, style_seed: int):
    """One synthetic utterance: ids ~ U{1..177} wrapped in the 0 pads of koko.rs:1168-1173, style ~ N(0, 0.15^2)."""
    rng = np.random.default_rng(seed)
    ids = np.concatenate([[0], rng.integers(1, 178, n_tokens), [0]]).astype(np.int64)
    style = np.random.default_rng(style_seed).normal(0, 0.15, 256).astype(np.float32)
    return ids, style


def synth_batch(B: int, n_tokens: int, rank: int = 0):
    """BASELINE configs[2]: B utterances of n_tokens ids (seeds 1000+b), styles (seeds 2000+b), speed 1.0."""
    toks, styles = [], []
    for b in range(B):
        ids, st = synth_case(n_tokens, 1000 + rank * 4096 + b, 2000 + rank * 4096 + b)
        toks.append(ids)
        styles.append(st)
    return toks, np.stack(styles), np.ones(B, np.float32)


def make_noise(n_frames_max: int, seed: int = 7) -> np.ndarray:
    """Explicit SineGen noise [600 * frames * 9] for parity runs (oracle and CUDA path read the same buffer)."""
    return np.random.default_rng(seed).standard_normal(600 * n_frames_max * 9).astype(np.float32)


def synth_requests_cfg4(n: int = 4096, seed: int = 4, n_voices: int = 54, zh_share: float = 0.10):
    """BASELINE configs[4]: n requests, N ~ U{10..510} tokens, a voice out of n_voices, speed ~ U[0.8, 1.3],
    zh_share of them flagged for the second model.  Returns (tokens list, voice ids, speeds, zh flags)."""
    rng = np.random.default_rng(seed)
    lens = rng.integers(10, 511, size=n)
    toks = [np.concatenate([[0], rng.integers(1, 178, int(k)), [0]]).astype(np.int64) for k in lens]
    voices = rng.integers(0, n_voices, size=n)
    speeds = rng.uniform(0.8, 1.3, size=n).astype(np.float32)
    zh = rng.random(n) < zh_share
    return toks, voices, speeds, zh


def synth_voice_table(n_voices: int = 54, seed: int = 54) -> np.ndarray:
    """[V, 511, 256] f32 synthetic voice table (the shape TTSKoko::load_voices builds, koko.rs:1308-1334)."""
    return (np.random.default_rng(seed).standard_normal((n_voices, 511, 256)) * 0.15).astype(np.float32)
