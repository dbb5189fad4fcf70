import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


if os.environ.get('G2_EMU') == '1':      # GPU tests on the CPU emulation of the kernel sources (tests/cuda_emu/emu_mode.py)
    sys.path.insert(0, os.path.join(ROOT, 'tests', 'cuda_emu'))
    import emu_mode
    emu_mode.enable()


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason='no CUDA device')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope='session')
def built_lib():
    """Build (if stale) and load libgenesis_b200.so."""
    from genesis_b200 import build, _lib
    build.build()
    return _lib.lib()
