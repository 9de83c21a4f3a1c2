import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
  sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))
if HERE not in sys.path:
  sys.path.insert(0, HERE)          # sibling helper modules (py_env_mocks)


def pytest_configure(config):
  config.addinivalue_line('markers', 'gpu: test needs a CUDA device (run on the B200 box)')


@pytest.fixture(scope='session')
def cuda():
  import torch
  if not torch.cuda.is_available():
    pytest.skip('needs a CUDA device')
  return torch.device('cuda:0')
