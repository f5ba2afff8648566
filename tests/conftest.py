import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    import numpy as np

    z = np.load(os.path.join(GOLDEN_DIR, name), allow_pickle=False)
    d = {k: z[k] for k in z.files}
    d["events"] = d["events"].astype(np.float64)
    d["covariates"] = {k: d[k] for k in ("C", "W", "N", "adjacency", "weekday", "area")}
    return d


GOLDEN_CASES = ["ref_M11_T32_s0.npz", "ref_M11_T32_s1.npz", "ref_M23_T17_s2.npz", "ref_M382_T84_s0.npz"]


@pytest.fixture(params=GOLDEN_CASES)
def golden(request):
    return load_golden(request.param)
