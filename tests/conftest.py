"""Test configuration. `-m "not gpu"` tests run on a CPU-only box (oracle against golden vectors, host
logic, C-ABI surface); `-m gpu` tests are the parity tests proper: they call libzombie_b200.so through its
C ABI on cuda:0 and compare with the oracle (oracle/_ref = the unmodified reference, compiled by
oracle/Makefile) and with the committed golden fixtures under tests/golden/.

Nothing here reads /root/reference at run time."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def ref():
    """ctypes binding of the compiled reference (test oracle)."""
    from oracle import ref as _ref
    if not _ref.available():
        pytest.fail("oracle/_ref/libzombie_ref.so is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                    "in a container that has /root/reference")
    return _ref


@pytest.fixture(scope="session")
def ctx():
    """zb_context on cuda:0 — fails loudly (never skips) when the CUDA library or device is missing."""
    from zombie_b200 import native
    return native.Context(0)


def rel_err(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return np.abs(a - b)/np.maximum(1e-30, np.maximum(np.abs(a), np.abs(b)))
