import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _have_gpu():
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _have_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture
def concrete_normalised_10_rows_data():
    """The reference's only data fixture (tests/conftest.py:11-41 there): 10 x 7 rows + y."""
    rows = ("1.988 -1.536 -1.155 -0.274 1.73 -0.3 -0.629; 1.988 -1.536 -1.155 -0.274 1.996 -0.3 -0.629; "
            "-0.061 0.321 0.855 -0.589 -0.188 -1.064 1.261; -0.061 0.321 0.855 -0.589 -0.188 -1.064 2.003; "
            "-1.383 0.189 -0.241 -0.589 0.636 1.093 1.964; -0.718 -0.051 0.855 -0.589 -0.188 -0.356 -0.145; "
            "0.408 -0.298 0.855 -0.589 -0.188 -1.064 2.003; 0.408 -0.298 0.855 -0.589 -0.188 -1.064 -0.629; "
            "-0.718 -0.051 0.855 -0.589 -0.188 -0.356 -0.629; 1.346 -1.536 0.855 -0.589 -0.188 -1.064 -0.629")
    X = np.array([[float(v) for v in r.split()] for r in rows.split(";")])
    y = np.array([[3.346], [1.825], [0.008], [0.073], [0.346], [0.576], [0.296], [-0.313], [0.477], [-0.075]])
    return X, y
