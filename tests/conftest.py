import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
GOLDEN_CASES = ["sphere_n2000_k50", "torus_n600_k20", "flat3torus_R6_n900_k24", "sheet_R20_n500_k16"]
GOLDEN_NB = {"sphere_n2000_k50": 10, "torus_n600_k20": 10, "flat3torus_R6_n900_k24": 10, "sheet_R20_n500_k16": 14}


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz")))


@pytest.fixture(params=GOLDEN_CASES)
def golden(request):
    g = load_golden(request.param)
    g["name"] = request.param
    g["nb"] = GOLDEN_NB[request.param]
    return g


def subspace_angle_max(A, B):
    """Largest principal angle between span(A) and span(B) (columns), robust for tiny angles."""
    import scipy.linalg
    return float(np.max(scipy.linalg.subspace_angles(A, B))) if A.shape[1] else 0.0


def eigen_clusters(evals, rtol=1e-6, atol=1e-9):
    """Group ascending eigenvalues into clusters of (near-)degenerate values -> list of slices."""
    out, s = [], 0
    for i in range(1, len(evals) + 1):
        if i == len(evals) or abs(evals[i] - evals[i - 1]) > max(atol, rtol * abs(evals[i])):
            out.append(slice(s, i))
            s = i
    return out
