"""Shared fixtures.  `-m gpu` tests call the product through its C ABI on a real B200; everything
else (oracle vs golden vectors / reference CPU code, host logic, ABI surface) runs on CPU."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

import __graft_entry__ as entry  # noqa: E402


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu)")


@pytest.fixture(scope="session")
def pkg():
    return entry.load_package()


@pytest.fixture(scope="session")
def oracle():
    from oracle.bindings import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def ref():
    """The reference's own code (oracle/_ref/libbsmr_ref.so); skipped when it was not built."""
    from oracle.bindings import Ref, REF_SO
    if not os.path.exists(REF_SO):
        pytest.skip("oracle/_ref/libbsmr_ref.so not built (needs /root/reference at build time)")
    return Ref()


@pytest.fixture(scope="session")
def ctx(pkg):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return pkg.Context(0)


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")
