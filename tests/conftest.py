"""pytest configuration: marker registration and shared fixtures/helpers."""

from __future__ import annotations

import os
import sys

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

GOLDEN = os.path.join(REPO, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_golden(name: str):
    return np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)


def golden_subunits(g) -> dict:
    """Rebuild the subunit dict of a golden case, preserving key and member order."""
    flat, offs, keys = g["su_flat"], g["su_offsets"], g["su_keys"]
    return {
        int(k): [int(a) for a in flat[offs[i] : offs[i + 1]]] for i, k in enumerate(keys)
    }


def golden_bonds(g) -> tuple:
    return tuple(tuple(int(x) for x in b) for b in g["long_bonds"])


@pytest.fixture(scope="session")
def golden_dir() -> str:
    return GOLDEN
