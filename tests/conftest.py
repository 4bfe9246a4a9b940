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


def _cuda_device_present():
    """True when the CUDA driver reports a device (libcuda only: nothing of this repo or torch is loaded for the check)."""
    import ctypes

    try:
        cuda = ctypes.CDLL("libcuda.so.1")
        n = ctypes.c_int(0)
        return cuda.cuInit(0) == 0 and cuda.cuDeviceGetCount(ctypes.byref(n)) == 0 and n.value > 0
    except OSError:
        return False


def pytest_collection_modifyitems(config, items):
    # plain `pytest tests` on a box without a GPU: the gpu-marked tests are skipped, not errors in the CPU results
    if any(it.get_closest_marker("gpu") for it in items) and not _cuda_device_present():
        skip = pytest.mark.skip(reason="no CUDA device (run with -m gpu on the B200 box)")
        for it in items:
            if it.get_closest_marker("gpu"):
                it.add_marker(skip)


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
