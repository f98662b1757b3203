import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line(
        'markers', 'gpu: needs a CUDA device (run on the B200 box)')


def load_golden(name):
    with np.load(os.path.join(GOLDEN, name + '.npz'), allow_pickle=False) as f:
        return {k: f[k] for k in f.files}


def cos_similarity(a, b):
    """Phase-invariant |<a,b>| / (|a||b|) over the last axis
    (tests/test_extraction/test_beamformer.py:18-22 of the reference)."""
    num = np.abs(np.sum(np.conj(a) * b, axis=-1))
    den = np.linalg.norm(a, axis=-1) * np.linalg.norm(b, axis=-1)
    return num / den


@pytest.fixture(scope='session')
def golden():
    return load_golden
