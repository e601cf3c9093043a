"""pytest configuration: markers and shared paths.

`-m "not gpu"` runs on the CPU-only build container (oracle vs golden vectors,
host logic, C-ABI symbol checks, gloo world_size-2 sharding logic);
`-m gpu` runs on a B200 and calls the CUDA path through the C-ABI.
"""
from __future__ import annotations

import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


@pytest.fixture(scope="session")
def golden_dir() -> Path:
    return ROOT / "tests" / "golden"
