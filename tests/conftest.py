import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WEIGHTS = os.environ.get("KKX_WEIGHTS", os.path.join(ROOT, "weights", "kokoro_random_1234.kkxw"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def ensure_weights(path: str = WEIGHTS) -> str:
    """Random-init Kokoro-82M weights (SURVEY.md 8d recipe, seed 1234); no checkpoint ships with
    the reference and there is no network, so both the oracle and the CUDA path load this file."""
    from kokorox_b200.weightfile import random_weights, write_weights
    if not os.path.exists(path):
        os.makedirs(os.path.dirname(path), exist_ok=True)
        tmp = path + ".tmp%d" % os.getpid()
        write_weights(tmp, random_weights(1234))
        os.replace(tmp, path)
    return path


@pytest.fixture(scope="session")
def weights_path():
    return ensure_weights()


@pytest.fixture(scope="session")
def weights(weights_path):
    from kokorox_b200.weightfile import read_weights
    return read_weights(weights_path)


@pytest.fixture(scope="session")
def oracle(weights):
    from oracle.kokoro_ref import KokoroOracle
    return KokoroOracle(weights)


def synth_case(n_tokens: int, seed: int, style_seed: int):
    """Synthetic utterance: ids ~ U{1..177} wrapped in the 0 pads of koko.rs:1168-1173,
    style ~ N(0, 0.15^2)."""
    rng = np.random.default_rng(seed)
    ids = np.concatenate([[0], rng.integers(1, 178, n_tokens), [0]]).astype(np.int64)
    style = np.random.default_rng(style_seed).normal(0, 0.15, 256).astype(np.float32)
    return ids, style


def make_noise(n_frames_max: int, seed: int = 7) -> np.ndarray:
    return np.random.default_rng(seed).standard_normal(600 * n_frames_max * 9).astype(np.float32)


# ort_koko.rs:46 -- the example input in the reference's own comment
REF_EXAMPLE_IDS = [0, 56, 51, 142, 156, 69, 63, 3, 16, 61, 4, 16, 156, 51, 4, 16, 62, 77, 156, 51, 86, 5, 0]
# tokenize.rs:124-126 -- "$h@l'oU, w'3:ld!$"
REF_TOKENIZE_IDS = [0, 50, 83, 54, 156, 57, 135, 3, 16, 65, 156, 87, 158, 54, 46, 5, 0]
