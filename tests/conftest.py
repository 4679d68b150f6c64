import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (B200); run with -m gpu on the GPU box')


@pytest.fixture(scope='session')
def ref_vectors():
    return np.load(os.path.join(GOLDEN, 'ref_vectors.npz'))


@pytest.fixture(scope='session')
def ranker_vectors():
    return np.load(os.path.join(GOLDEN, 'ref_rankers.npz'))


@pytest.fixture(scope='session')
def oracle_vectors():
    return np.load(os.path.join(GOLDEN, 'oracle_vectors.npz'))


@pytest.fixture(scope='session')
def eng():
    from es_pytorch_b200.engine import get_engine
    return get_engine(0)
