import gzip
import json
import os
import sys

import pytest

# every context the tests open checks the guard zones around its device buffers when it is closed
os.environ.setdefault("SB200_GUARD", "1")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _load(name):
    with gzip.open(os.path.join(ROOT, "tests", "golden", name), "rt") as f:
        return json.load(f)


@pytest.fixture(scope="session")
def golden_fixtures():
    """Outputs of the unmodified reference on its own 14 .lp files (+ generated Klee-Minty n=10)."""
    return _load("fixtures.json.gz")


@pytest.fixture(scope="session")
def golden_synth():
    """Outputs of the unmodified reference on the synthetic table G(m,n,n_eq,seed)."""
    return _load("synth.json.gz")


@pytest.fixture(scope="session")
def golden_parser_cases():
    """Outcomes of the unmodified reference on hand-written .lp texts (make_parser_cases.py)."""
    return _load("parser_cases.json.gz")


@pytest.fixture(scope="session")
def golden_cli():
    """What the reference's stock command line printed / wrote for every golden .lp text."""
    return _load("cli_outputs.json.gz")


@pytest.fixture(scope="session")
def golden_large():
    """Outputs of the unmodified reference at the sizes BASELINE.json quotes (make_golden_large.py)."""
    return _load("large.json.gz")


@pytest.fixture(scope="session")
def golden_bounded():
    """The unmodified reference on mid-size dense LPs with finite upper bounds (make_bounded_golden.py)."""
    return _load("bounded.json.gz")
