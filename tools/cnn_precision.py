"""Error of both classifier precisions against the torch fp32 / fp64 oracles (development aid)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import cnn_oracle  # noqa: E402
from deepvariant_b200 import call_variants as cv, modeling  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
shape = (100, 221, 7)
for seed in (21, 22):
  w = modeling.random_weights(7, seed)
  g = torch.Generator().manual_seed(seed)
  imgs = torch.randint(0, 255, (n,) + shape, dtype=torch.uint8, generator=g)
  w32 = cnn_oracle.ReferenceModel(w).forward(imgs).numpy()
  w64 = cnn_oracle.ReferenceModel(w, dtype=torch.float64).forward(imgs).numpy()
  for prec in (0, 1):
    net = cv.GpuCnn(w, shape, device=0, max_batch=n, precision=prec)
    got = net.forward_host(imgs.numpy())
    print(f'seed {seed} precision {prec}: |ours-fp32| {np.abs(got - w32).max():.3e}  |ours-fp64| {np.abs(got - w64).max():.3e}  '
          f'|fp32-fp64| {np.abs(w32 - w64).max():.3e}', flush=True)
    net.close()
