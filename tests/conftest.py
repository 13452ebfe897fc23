import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


@pytest.fixture(scope="session")
def repo_root():
    return ROOT


@pytest.fixture(scope="session")
def demo1():
    from flux_b200 import SceneData
    return SceneData.from_yaml(os.path.join(ROOT, "scenes", "demo1.yml"))


@pytest.fixture(scope="session")
def demo2():
    from flux_b200 import SceneData
    return SceneData.from_yaml(os.path.join(ROOT, "scenes", "demo2.yml"))


@pytest.fixture(scope="session")
def gpu_ctx():
    """One flux_ctx on cuda:0 for the whole GPU test session (fails loudly without the .so)."""
    from flux_b200.worker import GpuContext
    ctx = GpuContext(0)
    yield ctx
    ctx.close()
