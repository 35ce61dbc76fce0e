import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


SMALL = dict(H=200, W=200, K=(277.7775, 277.7775, 100.0, 100.0))


@pytest.fixture(scope="session")
def small_seq():
    """6 keyframes, 2 objects, 200x200 — the oracle finishes an iteration on it in well under a second."""
    from ro_map_b200 import synthetic as syn
    return syn.make_sequence(n_frames=6, n_objects=2, seed=1337, **SMALL)


@pytest.fixture(scope="session")
def oracle():
    from oracle import mon_oracle
    mon_oracle.build()
    return mon_oracle


def uniform_open_closed(rng: np.random.Generator, shape) -> np.ndarray:
    """(0,1] like curandGenerateUniform."""
    return (1.0 - rng.random(shape, dtype=np.float32)).astype(np.float32)
