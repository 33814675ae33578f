import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(ROOT, "tests", "golden", "reference_golden.json")) as fh:
        return json.load(fh)


@pytest.fixture(scope="session")
def biopython_tutorial():
    """Published known-answer vectors of the Bio.motifs example (tests/golden/biopython_tutorial.json)."""
    with open(os.path.join(ROOT, "tests", "golden", "biopython_tutorial.json")) as fh:
        return json.load(fh)


@pytest.fixture(scope="session")
def ctx():
    """One libtfbs_cuda context on cuda:0.  No skip-on-missing: a GPU test without the CUDA
    library or device must fail loudly."""
    from discover_tfbs_b200 import _lib

    c = _lib.Context(0)
    yield c
    c.close()
