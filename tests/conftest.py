import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "ref: needs oracle/_ref (the compiled unmodified reference)")


@pytest.fixture(scope="session", autouse=True)
def _build_checkers():
    """Build the CPU checkers (oracle/liboracle.so, and oracle/_ref when /root/reference exists)."""
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "all"], check=True,
                   stdout=subprocess.DEVNULL)
