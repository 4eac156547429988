import gzip
import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


def load_golden(name):
    path = os.path.join(GOLDEN, name)
    if name.endswith(".gz"):
        with gzip.open(path, "rb") as f:
            return json.loads(f.read())
    with open(path, "rb") as f:
        return json.loads(f.read())


@pytest.fixture(scope="session")
def traces():
    return load_golden("replay_traces.json.gz")


@pytest.fixture(scope="session")
def kat():
    return load_golden("kat.json")


@pytest.fixture(scope="session")
def pawn_cases():
    return load_golden("pawn_cases.json.gz")


def n_mcts_golden():
    """number of recorded mcts.MCTS runs (for parametrize at collection time)"""
    return len(load_golden("mcts_golden.json"))


@pytest.fixture(scope="session")
def mcts_golden():
    return load_golden("mcts_golden.json")
