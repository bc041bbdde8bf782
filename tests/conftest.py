import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    """ctypes handle on the C restatement (built on demand; test infrastructure only)."""
    import oracle_lib
    return oracle_lib.load()


@pytest.fixture(scope="session")
def ref_driver():
    """Path of oracle/_ref/ref_driver (the unmodified reference compiled with the shim), or skip."""
    path = os.path.join(ROOT, "oracle", "_ref", "ref_driver")
    if not os.path.exists(path):
        if os.path.isdir("/root/reference/src"):
            subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "ref"], check=True)
        else:
            pytest.skip("oracle/_ref/ref_driver not built and /root/reference absent")
    return path
