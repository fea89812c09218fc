import json
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
    config.addinivalue_line("markers", "slow: long-running exhaustive check")


@pytest.fixture(scope="session")
def golden_tables():
    return np.load(os.path.join(GOLDEN, "eval_tables.npz"))


@pytest.fixture(scope="session")
def golden_cases():
    with open(os.path.join(GOLDEN, "eval_cases.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def golden_enum():
    with open(os.path.join(GOLDEN, "enum_golden.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def golden_mc():
    with open(os.path.join(GOLDEN, "mc_seeded.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def golden_ranges():
    with open(os.path.join(GOLDEN, "mc_ranges_seeded.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def golden_preflop():
    with open(os.path.join(GOLDEN, "preflop_order.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def cuda_device():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda", 0)


@pytest.fixture(scope="session", autouse=True)
def checked_build_reports_no_failure():
    """When the suite runs against the checked build of the library (NPK_LIBRARY=.../libnpk_checked.so, see
    tools/checked_build.sh: the stand-in for compute-sanitizer), no bounds / deck-restore check may have failed."""
    yield
    if not os.environ.get("NPK_LIBRARY"):
        return
    import ctypes
    import torch
    if not torch.cuda.is_available():
        return
    from neuron_poker_b200 import _lib
    L = _lib.ensure_init(0)
    checked, code = ctypes.c_int(0), ctypes.c_uint32(0)
    _lib.check(L.npk_checked_status(ctypes.byref(checked), ctypes.byref(code)))
    print("\nchecked build: %d, first failed check: %d" % (checked.value, code.value))
    assert code.value == 0, "NPK_CHECK code %d failed (csrc/npk_device.cuh)" % code.value
