import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason='no CUDA device')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)


@pytest.fixture
def cpu_kernels(monkeypatch):
    """Replace every libb200gan entry point by its CPU stand-in (oracle/kernels_ref.py) so the
    host-side autograd algebra can be checked without a GPU.  Tests only -- the product has no
    such switch."""
    from gan_control_b200 import kernels
    from oracle import kernels_ref
    for name in kernels_ref.STAND_INS:
        monkeypatch.setattr(kernels, name, getattr(kernels_ref, name))
    return kernels_ref
