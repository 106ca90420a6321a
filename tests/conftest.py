import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle_lib():
    from oracle import oracle
    oracle.build()
    return oracle.lib()


@pytest.fixture(params=["fused", "unfused"])
def sweep_mode(request, monkeypatch):
    """Both Sweby drivers of the library: z + fused x/y pass (default) and the three separate sweeps (MOM5ADV_FUSE=0).
    The switch is read by mom5adv_init, i.e. per handle; spawned workers inherit the environment."""
    monkeypatch.setenv("MOM5ADV_FUSE", "1" if request.param == "fused" else "0")
    return request.param
