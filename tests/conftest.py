import os
import sys

# the GPU box's container may run under a CPU quota smaller than its core count: BLAS pools sized by the core count then
# thrash (a dense oracle solve took minutes there).  Small pools, set before numpy loads its BLAS.
for _v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
    os.environ.setdefault(_v, "4")

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


@pytest.fixture(scope="session")
def golden_match():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "match_golden.npz"))


@pytest.fixture(scope="session")
def ctx():
    import monocularsfm_b200 as m
    c = m.Context(0)
    yield c
    c.close()
