import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: test needs a CUDA GPU (B200); run with `pytest -m gpu`')


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:  # noqa
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason='no CUDA device')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope='session')
def golden():
    """Reference golden loss curves (tests/test_baseline.json of the reference; numbers only, read at test time)."""
    import json
    path = '/root/reference/tests/test_baseline.json'
    local = os.path.join(ROOT, 'tests', 'golden_losses.json')
    if os.path.exists(local):
        with open(local) as f:
            return json.load(f)
    if os.path.exists(path):
        with open(path) as f:
            data = json.load(f)
        return [{'top': d['top'], 'dtype': d['dtype'], 'num_local_experts': d['num_local_experts'],
                 'losses': [float(v) for v in d['losses'][:12]]} for d in data]
    pytest.skip('golden losses unavailable')
