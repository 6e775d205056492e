import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
GOLDEN = Path(__file__).resolve().parent / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session", autouse=True)
def _build_lib():
    # the C-ABI library must exist for every test that imports the product (built in-tree, no GPU needed)
    import shutil
    from sensorium_b200.build import LIB, _nvcc, build
    if shutil.which(_nvcc()) is None:
        # no nvcc on this machine: host-only tests still run; tests that need the library fail loudly on load
        if not LIB.exists():
            import warnings
            warnings.warn("nvcc not found and libdwn_b200.so is not built: tests that call the C ABI will fail")
        return
    build()
