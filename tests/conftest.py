import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
  sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
  config.addinivalue_line('markers', 'gpu: needs a CUDA (sm_100) device')


def pytest_collection_modifyitems(config, items):
  # `-m gpu` on a box without a GPU: fail loudly rather than silently skip.
  pass


@pytest.fixture(scope='session', autouse=True)
def _built_library():
  """The ctypes binding needs libbnf_sm100.so; build it if it is not there."""
  import __graft_entry__ as g
  if not os.path.exists(g.LIB):
    g.build()
