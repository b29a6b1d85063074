"""Shared pytest configuration: the ``gpu`` marker and golden-fixture loaders."""

from __future__ import annotations

import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
GOLDEN = ROOT / "tests" / "golden"
for p in (str(ROOT), str(ROOT / "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config: pytest.Config) -> None:
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config: pytest.Config, items: list[pytest.Item]) -> None:
    import torch

    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def balls7() -> dict:
    return dict(np.load(GOLDEN / "balls7.npz"))


@pytest.fixture(scope="session")
def xyz60() -> dict:
    return dict(np.load(GOLDEN / "xyz60.npz"))


@pytest.fixture(scope="session")
def ref_lens() -> dict:
    return dict(np.load(GOLDEN / "ref_lens.npz"))


@pytest.fixture(scope="session")
def ref_soap() -> dict:
    return dict(np.load(GOLDEN / "ref_soap.npz"))


@pytest.fixture(scope="session")
def numba_cases() -> dict:
    return dict(np.load(GOLDEN / "numba_cases.npz"))
