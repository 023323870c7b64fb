import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def emu_lib():
    """CPU logic-emulation build of the kernel sources (tests only, never the product path)."""
    from scamlgp_b200._capi import ScamlLib
    from tests.emu_build import build_emu

    return ScamlLib(build_emu())


@pytest.fixture(scope="session")
def engine():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from scamlgp_b200.engine import Engine

    return Engine(torch.device("cuda:0"))
