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
