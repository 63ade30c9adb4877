import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden():
    import gzip, json
    with gzip.open(os.path.join(ROOT, "tests", "golden", "golden.json.gz"), "rt") as f:
        return json.load(f)


@pytest.fixture(scope="session")
def golden_edge():
    import gzip, json
    with gzip.open(os.path.join(ROOT, "tests", "golden", "golden_edge.json.gz"), "rt") as f:
        return json.load(f)


@pytest.fixture(scope="session")
def emu_finder():
    """CPU single-stepper of the kernel bodies (tests/emu) -- logic tests only, never the product."""
    from common import build_emu, EMU_LIB
    from csa_b200.api import RotationFinder
    build_emu()
    rf = RotationFinder(lib_path=EMU_LIB)
    yield rf
    rf.close()


@pytest.fixture(scope="session")
def gpu_finder():
    """the product: libcsa_gpu.so on cuda:0, through the C ABI"""
    from csa_b200.api import RotationFinder
    rf = RotationFinder(device=0)
    yield rf
    rf.close()
