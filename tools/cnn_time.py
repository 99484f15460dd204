"""Device-side timing of the classifier alone (development aid)."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deepvariant_b200 import call_variants as cv, modeling  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--batch', type=int, default=2048)
ap.add_argument('--chunk', type=int, default=2048)
ap.add_argument('--steps', type=int, default=5)
ap.add_argument('--warmup', type=int, default=2)
ap.add_argument('--pacbio', action='store_true')
ap.add_argument('--precision', type=int, default=0)
a = ap.parse_args()
shape = (100, 147, 10) if a.pacbio else (100, 221, 7)
net = cv.GpuCnn(modeling.random_weights(shape[2], 0), shape, device=0, max_batch=a.chunk, precision=a.precision)
x = torch.randint(0, 255, (a.batch,) + shape, dtype=torch.uint8, device='cuda:0')
p = torch.empty((a.batch, 3), dtype=torch.float32, device='cuda:0')
s = torch.cuda.Stream(device='cuda:0') if os.environ.get('DVB_CNN_SIDE_STREAM') or os.environ.get('DVB_CNN_GRAPH') else torch.cuda.current_stream()   # graphs cannot be captured on the legacy default stream
torch.cuda.synchronize()
for _ in range(a.warmup):
  net.forward_device(x, p, stream=s)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(s)
for _ in range(a.steps):
  net.forward_device(x, p, stream=s)
e1.record(s)
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.steps
print(json.dumps({'precision': a.precision, 'batch': a.batch, 'chunk': a.chunk, 'ms_per_forward': ms, 'images_per_s': a.batch / ms * 1e3,
                  'tflops': a.batch * net.flops_per_image / ms / 1e9}))
