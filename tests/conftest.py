import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
	sys.path.insert(0, ROOT)
os.environ.setdefault('CSLVR_B200_QUIET', '1')


def pytest_configure(config):
	config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope='session')
def lib_built():
	"""build (or reuse) the in-tree shared library; CPU-only: compiles, never computes."""
	from cslvr_b200 import _build
	return _build.build()
