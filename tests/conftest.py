import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
  sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
  config.addinivalue_line('markers', 'gpu: test needs a CUDA device (run on the B200 box)')
  # A fresh checkout has no built library (build artefacts are git-ignored): build it once, in-tree, with nvcc.
  lib = os.path.join(ROOT, 'deepvariant_b200', 'csrc', 'libdvb.so')
  if not os.path.exists(lib):
    import __graft_entry__
    __graft_entry__.build()


def _have_gpu() -> bool:
  try:
    import torch
    return torch.cuda.is_available()
  except Exception:  # pylint: disable=broad-except
    return False


def pytest_collection_modifyitems(config, items):
  # `-m gpu` on a box without a GPU must fail loudly rather than silently pass; but when the
  # whole suite is collected without -m on a CPU box, skip the gpu tests.
  if _have_gpu():
    return
  if 'gpu' in (config.getoption('-m') or ''):
    return
  skip = pytest.mark.skip(reason='no CUDA device')
  for item in items:
    if 'gpu' in item.keywords:
      item.add_marker(skip)


@pytest.fixture(params=['oracle', pytest.param('gpu', marks=pytest.mark.gpu)])
def backend(request):
  """Encoder factory: options -> object with the PileupImageEncoderNative surface.
  'oracle' = CPU restatement (checks the oracle against the reference's KATs);
  'gpu'    = the CUDA path through the C ABI (checks the product against the same KATs)."""
  if request.param == 'oracle':
    import oracle_lib
    return oracle_lib.OraclePileupImageEncoder
  from deepvariant_b200 import pileup_image
  return pileup_image.PileupImageEncoderNative
