import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    path = os.path.join(ROOT, "tests", "golden", "krylov_golden.pt")
    return torch.load(path, weights_only=False)


@pytest.fixture(scope="session")
def golden_generalized():
    """the reference's davidson on generalized problems and wide start blocks (oracle/gen_golden_generalized.py)"""
    path = os.path.join(ROOT, "tests", "golden", "generalized_golden.pt")
    return torch.load(path, weights_only=False)["davidson_generalized"]


@pytest.fixture(scope="session", autouse=True)
def _build_extension():
    # the product fails loudly without its CUDA extension; build it (nvcc cross-compiles on CPU boxes)
    from xitorch_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        from xitorch_b200.csrc.build import build
        build()


@pytest.fixture(scope="session")
def emu_lib(tmp_path_factory):
    """the engines' sources as a host build (tools/emu_engine), built once per test session"""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import emu_engine_lib
    lib = emu_engine_lib.build(str(tmp_path_factory.mktemp("emu_engine")))
    if lib is None:
        pytest.skip("g++ not available")
    return lib
