import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    from oracle_lib import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def hw():
    """the product package (builds libhw1f.so in-tree if it is stale or missing)"""
    import __graft_entry__ as g
    g.build_engine()
    import hw1f_b200
    return hw1f_b200


@pytest.fixture(scope="session", params=["decomposed", "reference_order"])
def engine(hw, request):
    """every GPU parity test runs in both simulation modes (include/hw1f.h: HW1F_MODE_*)"""
    eng = hw.Engine(device=0)
    eng.set_mode(hw._ffi.MODE_DECOMPOSED if request.param == "decomposed" else hw._ffi.MODE_REFERENCE_ORDER)
    yield eng
    eng.close()


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")
