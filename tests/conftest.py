import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle_lib
    oracle_lib.build()
    oracle_lib.lib()
    return oracle_lib


@pytest.fixture(scope="session")
def pmc_factory():
    """Factory of PMC contexts on cuda:0; fails loudly if the CUDA library is absent."""
    import torch
    from cosmopmc_b200.pmc import PMC
    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    made = []

    def make():
        p = PMC(0)
        made.append(p)
        return p
    yield make
    for p in made:
        p.close()
