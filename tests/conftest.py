import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
sys.path.insert(0, str(ROOT / "oracle"))  # oracle_build: the checker's build recipes


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "slow: long-running CPU test")


def _has_gpu() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


# GPU tests run from the bottom of the stack up: the kernels against the oracle through the C ABI
# first (the parity tests proper), then the full-size properties, then model_t on the engine, the
# reference's own sources on the engine, and the multi-GPU runs last -- a run stopped at its first
# failure (-x) has then covered everything below the layer that failed.
GPU_ORDER = ["test_gpu_parity", "test_gpu_fullsize", "test_gpu_model", "test_gpu_reference_sources",
             "test_gpu_multi", "test_gpu_zz_partition_exchange"]


def _gpu_rank(item) -> int:
    name = Path(str(item.fspath)).stem
    return GPU_ORDER.index(name) if name in GPU_ORDER else -1  # everything else keeps its place in front


def pytest_collection_modifyitems(config, items):
    items.sort(key=_gpu_rank)  # stable: the order inside a file is kept
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return ROOT / "tests" / "golden"
