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


@pytest.fixture(params=["fused", "unfused", "fused_ldgsts"])
def sweep_mode(request, monkeypatch):
    """The Sweby drivers of the library: z + fused x/y pass with TMA staging (default; blocks with an odd ni fall back to LDGSTS
    staging by themselves), the three separate sweeps (MOM5ADV_FUSE=0) and the fused driver with per-thread LDGSTS staging forced
    (MOM5ADV_TMA=0).  The switches are read by mom5adv_init, i.e. per handle; spawned workers inherit the environment."""
    monkeypatch.setenv("MOM5ADV_FUSE", "0" if request.param == "unfused" else "1")
    monkeypatch.setenv("MOM5ADV_TMA", "0" if request.param == "fused_ldgsts" else "3")
    return request.param
