import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def cfg1_clip():
    """BASELINE cfg1: 10 s, 16 kHz, 8-partial sine mix."""
    import numpy as np
    sr, n = 16000, 160000
    midi = [45, 52, 57, 60, 64, 69, 76, 81]
    ph = np.random.default_rng(0).uniform(0, 2 * np.pi, 8)
    t = np.arange(n)
    y = sum((0.5 / (k + 1)) * np.sin(2 * np.pi * 440.0 * 2 ** ((m - 69) / 12) * t / sr + ph[k])
            for k, m in enumerate(midi))
    return y.astype(np.float32), sr


@pytest.fixture(scope="session")
def cfg1():
    return cfg1_clip()
