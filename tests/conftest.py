import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box: pytest -m gpu)")


def _cuda_devices() -> int:
    """CUDA devices the product library sees (0 when the library is missing or no GPU is present)."""
    try:
        from boondock_airband_b200 import engine
        return max(0, int(engine.load_library().ba_cuda_visible_devices()))
    except Exception:
        return 0


def pytest_collection_modifyitems(config, items):
    """`pytest tests` on a box without a GPU skips the gpu-marked tests instead of failing at the first one; an explicit
    `-m gpu` still runs (and fails loudly) there: the CUDA path has no CPU stand-in."""
    if "gpu" in (config.getoption("-m") or ""):
        return
    if _cuda_devices() > 0:
        return
    skip = pytest.mark.skip(reason="no CUDA device visible (the engine has no CPU path); run with -m gpu on a B200")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def oracle_built():
    """Compile the CPU oracle (test infrastructure) once per session."""
    from oracle import ba_oracle
    ba_oracle.build()
    return ba_oracle
