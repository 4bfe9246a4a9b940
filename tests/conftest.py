import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TESTS = os.path.join(ROOT, "tests")
for _p in (ROOT, TESTS):
    if _p not in sys.path:
        sys.path.insert(0, _p)


# The GPU witness parser (ppd_parse.cu) only takes witnesses above a size threshold in production (small ones are
# latency-bound, the host builder is faster); the tests run every witness through it.
os.environ.setdefault("PPD_GPU_PARSE_MIN_BYTES", "0")
os.environ.setdefault("PPD_GPU_DUMP_MIN_TOUCHED", "0")  # likewise for the GPU IR dump


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (oracle/libppd_oracle.so), built on demand.  Checker only."""
    import ppd_oracle_lib as oracle_lib

    return oracle_lib.load()


@pytest.fixture(scope="session")
def goldens():
    import json

    with open(os.path.join(ROOT, "tests", "golden", "reference_goldens.json")) as f:
        return json.load(f)
