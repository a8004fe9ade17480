"""pytest configuration: registers the `gpu` marker and puts the repo + package on sys.path.

`-m "not gpu"` : oracle vs the reference-generated golden vectors, host logic, C-ABI symbol check (no GPU needed).
`-m gpu`       : the parity tests proper -- CUDA path through the C ABI vs the oracle / goldens on a B200.
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "mp3-steganography-lib_b200")
GOLDEN = os.path.join(ROOT, "tests", "golden")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run by the driver with -m gpu)")


@pytest.fixture(scope="session")
def built():
    """Build (if stale) the CUDA library and the oracle once per session."""
    import __graft_entry__ as ge
    ge.build()
    return ge


@pytest.fixture(scope="session")
def oracle(built):
    from oracle import oracle as O
    return O


@pytest.fixture(scope="session")
def handle(built):
    from mp3stego_b200 import _lib
    h = _lib.Handle(0)
    yield h
    h.close()


def golden_path(name):
    return os.path.join(GOLDEN, name)


def load_npz(name):
    return np.load(golden_path(name), allow_pickle=False)


SYNTH_CASES = ["s11_128_plain", "s11_128_hide", "s12_320_plain", "s12_320_hide", "s13_64_hide",
               "quiet_128_plain", "quiet_128_hide"]


def synth_wav(seed, n_frames, sr=44100):
    """SURVEY.md 8(d) tone+noise clip (same generator as tests/golden/make_golden.py)."""
    rng = np.random.default_rng(seed)
    n = n_frames * 1152
    t = np.arange(n) / sr
    f = rng.uniform(100, 5000, size=2)
    x = np.stack([0.4 * np.sin(2 * np.pi * f[c] * t) + 0.05 * rng.standard_normal(n) for c in range(2)], axis=1)
    return (x * 32767).astype(np.int16)
