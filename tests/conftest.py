import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (oracle/, C++ f64 restatement of the reference) — the checker."""
    from oracle import pyoracle
    pyoracle.lib()
    return pyoracle


@pytest.fixture(scope="session")
def rtw():
    """The product package; building librtw_cuda.so needs nvcc but no GPU."""
    import ray_tracing_weekend_b200 as R
    R.load()
    return R


SEED = 20261018


@pytest.fixture(scope="session")
def simple_scene(rtw, oracle):
    """scenes::simple built by the product host code and by the oracle, plus matching cameras."""
    world, lights, cb = rtw.scenes.simple(SEED)
    desc = oracle.scene_simple(SEED)
    return dict(world=world, lights=lights, cb=cb, desc=desc, oscene=oracle.Scene(desc))
