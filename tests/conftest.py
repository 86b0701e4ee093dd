import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

REFERENCE_DIR = Path("/root/reference")  # only exists in the build container, never on the GPU box


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def pytest_collection_modifyitems(config, items):
    # GPU tests are only meaningful on the GPU box; skip them when no device is visible so that a
    # plain `pytest tests/` on the CPU container stays green.
    import torch

    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device visible")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return Path(__file__).resolve().parent / "golden"


@pytest.fixture(scope="session")
def have_reference():
    return (REFERENCE_DIR / "src" / "feature_extraction.py").exists()
