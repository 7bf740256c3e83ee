import os
import sys

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), '..'))
for p in (ROOT, os.path.join(ROOT, 'tests')):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')
    config.addinivalue_line('markers', 'needs_reference: needs /root/reference (build container only)')


def pytest_collection_modifyitems(config, items):
    import torch
    has_gpu = torch.cuda.is_available()
    has_ref = os.path.isdir(os.environ.get('GATOR_REFERENCE', '/root/reference'))
    for it in items:
        if 'gpu' in it.keywords and not has_gpu:
            it.add_marker(pytest.mark.skip(reason='no CUDA device'))
        if 'needs_reference' in it.keywords and not has_ref:
            it.add_marker(pytest.mark.skip(reason='reference tree not present'))
