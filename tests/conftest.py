import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, 'tests', 'golden')
DATA = os.path.join(ROOT, 'trackingbench_slam_b200', 'data')   # the package's input images (the reference's data/*.png)

# (n_levels, scale, target, initTh, minTh): A = the reference's own call sites
# (test_matcher.cpp:25-39, test_vo.cpp:194-200), B = BASELINE.json's "8 levels, scale 1.2".
PARAM_SETS = {
    'A1000': (5, 0.8, 1000, 80.0, 30.0),
    'A2000': (5, 0.8, 2000, 80.0, 30.0),
    'B1000': (8, 1.0 / 1.2, 1000, 20.0, 7.0),
    'B2000': (8, 1.0 / 1.2, 2000, 20.0, 7.0),
}


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')


@pytest.fixture(scope='session')
def images():
    import cv2
    out = {}
    for n in ('left', 'right', 'disparity'):
        out[n] = cv2.imread(os.path.join(DATA, n + '.png'), cv2.IMREAD_GRAYSCALE)
        assert out[n] is not None
    return out


@pytest.fixture(scope='session')
def oracle():
    import oracle as O
    O.build()
    return O


@pytest.fixture(scope='session')
def tb():
    import trackingbench_slam_b200 as tb
    from trackingbench_slam_b200 import build as B
    if not os.path.exists(B.LIB):   # a fresh checkout: build the CUDA library once (nvcc cross-compiles without a GPU)
        B.build()
    return tb


@pytest.fixture(scope='session')
def ctx(tb):
    c = tb.Context(0)
    yield c
    c.close()


def kp_bytes_equal(a, b):
    return len(a) == len(b) and np.ascontiguousarray(a).tobytes() == np.ascontiguousarray(b).tobytes()
