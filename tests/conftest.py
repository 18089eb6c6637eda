"""pytest configuration: registers the `gpu` marker and shared fixtures.

-m "not gpu": oracle vs golden vectors / zlib / compiled reference, host logic,
              symbol exports of the C-ABI library, SIMT-emulator kernel logic.
-m gpu:       parity tests proper, through the C ABI of libjdeflate.so on a B200.
"""
import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
for _p in (str(ROOT), str(ROOT / "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run on the GPU box")


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _have_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def jd():
    """The product library (fails loudly when it was not built)."""
    from jdeflate_b200 import api
    return api.load()


@pytest.fixture(scope="session")
def ref():
    """The compiled reference (oracle/_ref), prebuilt in the build container."""
    from jdeflate_b200 import api
    p = ROOT / "oracle" / "_ref" / "libjdeflate_ref.so"
    if not p.exists():
        pytest.skip("oracle/_ref/libjdeflate_ref.so not built")
    return api.JDeflateLib(p)


@pytest.fixture(scope="session")
def emu():
    """TEST INFRASTRUCTURE: the same kernel + host sources compiled for the CPU SIMT
    emulator (tests/simt).  Lets `-m "not gpu"` exercise kernel and host logic through
    the very same C ABI; it is never shipped and never used by the product."""
    from jdeflate_b200 import api
    from jdeflate_b200.build import build_emu
    return api.JDeflateLib(build_emu())


@pytest.fixture(scope="session", params=["emu", pytest.param("gpu", marks=pytest.mark.gpu)])
def lib(request):
    """Library under test: the SIMT-emulator build on CPU, the product build on a B200."""
    return request.getfixturevalue("emu" if request.param == "emu" else "jd")


@pytest.fixture(scope="session")
def oracle():
    from support import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def corpus():
    from support import Corpus
    return Corpus()


@pytest.fixture(scope="session")
def golden():
    import json
    return json.loads((ROOT / "tests" / "golden" / "golden.json").read_text())
