import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session", autouse=True)
def _built_oracle():
    # building the checker is not using it: oracle/c -> oracle/_build/liboracle.so
    from oracle import c_ref

    c_ref.build()
    yield


@pytest.fixture(scope="session", autouse=True)
def _cpu_oracle_nms_rule():
    """The NMS fixtures come from the CPU oracle, where torchvision switches from the coordinate-offset trick to the
    per-class loop at numel > 4000 (100 000 on CUDA, the product's default): compare like with like."""
    import importlib

    lnms = importlib.import_module("cddmsl_b200.layers.nms")
    old = lnms.COORD_TRICK_NUMEL_LIMIT
    lnms.COORD_TRICK_NUMEL_LIMIT = 4000
    yield
    lnms.COORD_TRICK_NUMEL_LIMIT = old
