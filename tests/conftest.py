import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    # NaN / inf patterns are compared on purpose (the reference produces them at the edges): inf - inf is expected
    config.addinivalue_line("filterwarnings", "ignore:invalid value encountered:RuntimeWarning")


def pytest_sessionstart(session):
    """Built artefacts are not in git: on a fresh checkout compile the C-ABI library (nvcc cross-compiles sm_100a
    without a GPU) and the oracle before the first test needs them."""
    import __graft_entry__
    __graft_entry__.ensure_built()


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_cuda = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        has_cuda = False
    if has_cuda:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.build()
    return O


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")
