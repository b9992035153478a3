import os
import sys
import wave

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def load_wav(name):
    with wave.open(os.path.join(GOLDEN, name), "rb") as f:
        return np.frombuffer(f.readframes(f.getnframes()), dtype="<i2").copy()


@pytest.fixture(scope="session")
def test_pcm():
    return load_wav("test.wav")


@pytest.fixture(scope="session")
def noise_pcm():
    return load_wav("noise.wav")


@pytest.fixture(scope="session")
def library_path():
    from koala_b200 import _build
    return _build.build()          # compiles only if sources are newer than the .so


@pytest.fixture(scope="session")
def shipped_model_path():
    from koala_b200 import default_model_path
    return default_model_path()


@pytest.fixture(scope="session")
def random_model_path(tmp_path_factory):
    from koala_b200 import spec
    p = str(tmp_path_factory.mktemp("model") / "random.kpv")
    spec.save_model(p, spec.random_model())
    return p


def synth_pcm(n_streams, n_frames, seed=0x4B4F414C):
    """Synthetic 16 kHz PCM of SURVEY.md section 8d: half noise-only streams, half speech-like + noise."""
    rng = np.random.default_rng(seed)
    n = n_frames * 256
    t = np.arange(n) / 16000.0
    out = np.empty((n_streams, n), np.float64)
    for s in range(n_streams):
        noise = rng.standard_normal(n) * 760.0
        if s % 2 == 0:
            out[s] = noise
        else:
            f0 = rng.uniform(100, 250)
            harm = sum(np.sin(2 * np.pi * f0 * k * t + rng.uniform(0, 6.28)) / k for k in range(1, 12))
            env = 0.5 * (1 + np.sin(2 * np.pi * 4.0 * t + rng.uniform(0, 6.28)))
            sp = harm * env
            out[s] = sp / (np.sqrt(np.mean(sp ** 2)) + 1e-9) * 2030.0 + noise
    return np.clip(np.rint(out), -32768, 32767).astype(np.int16).reshape(n_streams, n_frames, 256)
