import os
import sys
import types

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run by the driver with -m gpu)")


def make_hp(**overrides):
    """Reference yaml defaults (neural_speech/hparams/audio.yaml:6-18) as a plain namespace."""
    cfg = dict(num_mels=80, num_freq=1025, sample_rate=20000, frame_length_ms=50, frame_shift_ms=12.5,
               preemphasis=0.97, min_level_db=100, ref_level_db=20, max_iters=300, griffin_lim_iters=60,
               power=1.5, silence_threshold=0.1)
    cfg.update(overrides)
    return types.SimpleNamespace(**cfg)


def speechlike(n, seed, sr=20000):
    rng = np.random.RandomState(seed)
    t = np.arange(n) / sr
    f0 = rng.uniform(100, 250)
    x = np.zeros(n)
    for k in range(1, 31):
        x += np.sin(2 * np.pi * f0 * k * t + rng.uniform(0, 2 * np.pi)) / k
    x *= 0.6 + 0.4 * np.sin(2 * np.pi * 3 * t)
    x += 10 ** (-50 / 20) * rng.randn(n)
    x *= 0.9 / np.max(np.abs(x))
    return x.astype(np.float32)


@pytest.fixture(scope="session")
def golden():
    path = os.path.join(ROOT, "tests", "golden", "reference_audio.npz")
    return dict(np.load(path))


def trim_signals():
    """the clips of tests/golden/make_golden_process.py (same seeds, same code)"""
    out = {}
    out["quiet_ends"] = np.concatenate([0.002 * speechlike(9000, 1), speechlike(30000, 2), 0.001 * speechlike(12000, 3)]).astype(np.float32)
    out["pauses"] = np.concatenate([np.zeros(5000, np.float32), speechlike(15000, 4), np.zeros(7000, np.float32), speechlike(800, 5),
                                    np.zeros(6000, np.float32), speechlike(14000, 6), np.zeros(3000, np.float32)]).astype(np.float32)
    out["loud"] = speechlike(20000, 7)
    out["short"] = speechlike(1500, 8)
    out["silence"] = np.zeros(8000, np.float32)
    return out


@pytest.fixture(scope="session")
def golden_process():
    return dict(np.load(os.path.join(ROOT, "tests", "golden", "reference_process.npz")))


@pytest.fixture(scope="session")
def golden_tf():
    return dict(np.load(os.path.join(ROOT, "tests", "golden", "reference_tf_twin.npz")))
