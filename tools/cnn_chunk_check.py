"""Classifier chunk size: time and compare (bitwise) the forward of the same images at max_batch 4096 and 8192.  Development aid (GPU)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deepvariant_b200 import call_variants as cv, modeling  # noqa: E402

shape = (100, 221, 7)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
w = modeling.random_weights(7, 0)
x = torch.randint(0, 255, (B,) + shape, dtype=torch.uint8, device='cuda:0')
s = torch.cuda.current_stream()
out = {}
ref = None
for chunk in (4096, 8192):
  net = cv.GpuCnn(w, shape, device=0, max_batch=chunk, precision=0)
  p = torch.empty((B, 3), dtype=torch.float32, device='cuda:0')
  for _ in range(2):
    net.forward_device(x, p, stream=s)
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record(s)
  for _ in range(4):
    net.forward_device(x, p, stream=s)
  e1.record(s)
  torch.cuda.synchronize()
  ms = e0.elapsed_time(e1) / 4
  if ref is None:
    ref = p.clone()
  out[chunk] = {'ms_per_forward': ms, 'images_per_s': B / ms * 1e3, 'max_abs_diff_vs_4096': float((p - ref).abs().max()), 'finite': bool(torch.isfinite(p).all())}
  net.close()
  del net
  torch.cuda.empty_cache()
print(json.dumps(out))
