import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "sc-gameengine_b200"))
sys.path.insert(0, str(ROOT / "tests"))
sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu under gpurun)")


@pytest.fixture(scope="session")
def port():
    import oracle_bind
    return oracle_bind.port_lib()


@pytest.fixture(scope="session")
def ref():
    import oracle_bind
    if not oracle_bind.ref_available():
        pytest.skip("oracle/_ref/libscref.so not built (needs /root/reference; `make -C oracle ref`)")
    return oracle_bind.ref_lib()


@pytest.fixture(scope="session")
def gpu_lib():
    import scgpu
    return scgpu.load_library()
