import os
import sys

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200)")


@pytest.fixture(scope="session")
def oracle():
    """The plain-C restatement (oracle/ipx_oracle.c); built on demand with gcc."""
    from oracle import pyoracle
    pyoracle.lib()
    return pyoracle


@pytest.fixture(scope="session")
def reflib():
    """The compiled reference (oracle/_ref/libipx_ref.so) when it is present."""
    from oracle import ipxlib
    if not os.path.exists(ipxlib.REF_LIB):
        pytest.skip("oracle/_ref/libipx_ref.so not built (needs /root/reference)")
    return ipxlib.IpxLibrary(ipxlib.REF_LIB)


@pytest.fixture(scope="session")
def gpulib():
    """IPX with the GPU drop-ins (ipx_b200/_build/libipx_gpu.so)."""
    from oracle import ipxlib
    if not os.path.exists(ipxlib.GPU_LIB):
        pytest.fail("ipx_b200/_build/libipx_gpu.so missing: run __graft_entry__.build()")
    return ipxlib.IpxLibrary(ipxlib.GPU_LIB)


def rel_err(a, b):
    """Norm-wise relative error ||a-b||inf / ||b||inf."""
    a, b = np.asarray(a), np.asarray(b)
    denom = np.abs(b).max() if b.size else 0.0
    num = np.abs(a - b).max() if a.size else 0.0
    return num / denom if denom > 0 else num
