import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from kokorox_b200.synth import (REF_EXAMPLE_IDS, REF_TOKENIZE_IDS, WEIGHTS, ensure_weights, make_noise,  # noqa: E402,F401
                                 synth_case)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


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
