"""Shared fixtures.  `-m gpu` tests need a B200; everything else runs on CPU.

Only files under tests/ (plus __graft_entry__.smoke and bench.py's cpu_baseline legs) may import `oracle`.
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (ROOT, os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200, sm_100a)")
    config.addinivalue_line("markers", "slow: takes more than ~30 s")


def has_gpu() -> bool:
    try:
        from dicomautomaton_b200 import capi
        return capi.lib().dcma_b200_device_count() > 0
    except Exception:
        return False


@pytest.fixture(scope="session")
def oracle_mod():
    import oracle
    oracle.build_port()
    return oracle
