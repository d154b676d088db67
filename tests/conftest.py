import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _gpu_count():
    try:
        from waafle_b200 import engine
        return engine.load_library().wfl_device_count()
    except Exception:
        return 0


def pytest_collection_modifyitems(config, items):
    """`gpu` tests are skipped on a host without a CUDA device -- unless they were asked for (`-m gpu`) or
    WFL_REQUIRE_GPU=1 is set, in which case they fail loudly (there is no CPU fallback to hide behind)."""
    asked = "gpu" in (config.getoption("-m") or "") and "not gpu" not in (config.getoption("-m") or "")
    if asked or os.environ.get("WFL_REQUIRE_GPU") == "1" or _gpu_count() > 0:
        return
    skip = pytest.mark.skip(reason="no CUDA device on this host (the engine has no CPU fallback)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def engine():
    """One engine on cuda:0 for the whole GPU session (fails loudly without the .so / a GPU)."""
    from waafle_b200.engine import Engine
    e = Engine(0)
    yield e
    e.close()
