import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


@pytest.fixture(scope='session')
def golden():
    def load(name):
        return np.load(os.path.join(GOLDEN, name + '.npz'), allow_pickle=False)
    return load


def golden_params(g, prefix='init_'):
    p = {k: g[prefix + k] for k in ('E_user', 'E_item', 'W', 'b')}
    p['Feat'] = g['feat']
    p['expo'] = g['expo']
    return p


def rel_err(a, b):
    """max |a-b| / max |b|  — the norm-wise relative error of SURVEY.md Appendix D."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    denom = np.abs(b).max()
    return float(np.abs(a - b).max() / (denom if denom > 0 else 1.0))
