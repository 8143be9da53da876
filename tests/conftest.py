import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def golden_cases():
    return json.load(open(os.path.join(GOLDEN, "golden_cases.json")))


@pytest.fixture(scope="session")
def matrix_100():
    """The reference's src/tests/matrix.txt as float64 (see tests/golden/make_golden.py)."""
    return np.load(os.path.join(GOLDEN, "matrix_100.npy"))


def case_inputs(name):
    """Rebuilds the seeded input matrices of a golden case (same recipe as make_golden.py)."""
    from oracle import oracle as orc

    if name.startswith("matrix_txt"):
        return np.load(os.path.join(GOLDEN, "matrix_100.npy")), None
    if name.startswith("readme_std"):
        return orc.generate_diagonal_dominant(50, 1e-4, seed=0), None
    if name.startswith("readme_gev"):
        return (orc.generate_diagonal_dominant(50, 1e-4, seed=0),
                orc.generate_diagonal_dominant(50, 1e-4, 1.0, seed=1))
    if name.startswith("test_dense_numpy_std"):
        return orc.generate_diagonal_dominant(50, 1e-3, seed=2), None
    if name.startswith("test_dense_numpy_gen"):
        return (orc.generate_diagonal_dominant(50, 1e-3, seed=2),
                orc.generate_diagonal_dominant(50, 1e-3, 1.0, seed=3))
    if name.startswith("main_f90"):
        return (orc.generate_diagonal_dominant(100, 1e-3, seed=4),
                orc.generate_diagonal_dominant(100, 1e-3, 1.0, seed=5))
    if name == "collapse_n1000_DPR":
        return orc.generate_diagonal_dominant(1000, 1e-2, seed=0), None
    if name == "collapse_n1000_gev_DPR":
        return (orc.generate_diagonal_dominant(1000, 1e-2, seed=0),
                orc.generate_diagonal_dominant(1000, 1e-2, 1.0, seed=1))
    if name == "collapse_n2000_DPR":
        return orc.generate_diagonal_dominant(2000, 5e-2, seed=0), None
    if name == "notconverged_DPR":
        return orc.generate_diagonal_dominant(400, 1e-3, seed=7), None
    raise KeyError(name)
