import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200, sm_100a); run with -m gpu")


def pytest_collection_modifyitems(config, items):
    """GPU tests must never silently pass on a box without a GPU: they are skipped
    unless selected with -m gpu, and then they hard-fail if CUDA is missing."""
    markexpr = config.getoption("-m") or ""
    if "gpu" in markexpr and "not gpu" not in markexpr:
        return
    skip = pytest.mark.skip(reason="GPU test: select with -m gpu")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def cuda_device():
    import torch

    assert torch.cuda.is_available(), "-m gpu tests need a CUDA device"
    from sam_quantization_b200 import _lib

    _lib.device_check(torch.device("cuda:0"))
    return torch.device("cuda:0")


@pytest.fixture
def samq_env(monkeypatch):
    """Set / unset a SAMQ_* developer switch AND make the library re-read its configuration (the
    switches are resolved once at load, not per call); everything is restored afterwards."""
    from sam_quantization_b200 import _lib

    class Env:
        def set(self, name, value):
            monkeypatch.setenv(name, value)
            _lib.reload_config()

        def unset(self, name):
            monkeypatch.delenv(name, raising=False)
            _lib.reload_config()

    yield Env()
    monkeypatch.undo()
    _lib.reload_config()
