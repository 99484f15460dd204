"""Development probe: per-layer relative error of the CUDA CNN vs the torch fp32 oracle."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import cnn_oracle  # noqa: E402
from deepvariant_b200 import call_variants as cv, modeling  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 3
shape = (100, 221, 7)
w = modeling.random_weights(7, 0)
net = cv.GpuCnn(w, shape, device=0, max_batch=n)
g = torch.Generator().manual_seed(0)
imgs = torch.randint(0, 255, (n,) + shape, dtype=torch.uint8, generator=g)
probs = torch.empty((n, 3), dtype=torch.float32, device='cuda:0')
net.forward_device(imgs.to('cuda:0'), probs)
torch.cuda.synchronize()
want_p, tensors, pooled = cnn_oracle.ReferenceModel(w).forward(imgs, return_tensors=True)
ops, _ = modeling.inception_v3_graph(7)
seen = []
for o in ops:
  if o.dst in seen:
    continue
  seen.append(o.dst)
for name in seen:
  got = net.debug_tensor(name, n)
  ref = tensors[name].permute(0, 2, 3, 1).numpy()
  scale = max(float(np.abs(ref).max()), 1e-6)
  d = np.abs(got - ref)
  print(f'{name:12s} shape {tuple(ref.shape)} rel_err {d.max() / scale:.3e} mean_abs_err {d.mean():.3e} ref_absmean {np.abs(ref).mean():.3f}', flush=True)
print('probs gpu', probs.cpu().numpy().tolist())
print('probs ref', want_p.numpy().tolist())
print('max |dp|', float((probs.cpu() - want_p).abs().max()))
# quick timing
B = int(os.environ.get('PROBE_BATCH', '512'))
net2 = cv.GpuCnn(w, shape, device=0, max_batch=B)
x = torch.randint(0, 255, (B,) + shape, dtype=torch.uint8, device='cuda:0')
p = torch.empty((B, 3), dtype=torch.float32, device='cuda:0')
for _ in range(2):
  net2.forward_device(x, p)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
  net2.forward_device(x, p)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(f'batch {B}: {ms:.2f} ms/forward, {B / ms * 1e3:.0f} img/s, {B * net2.flops_per_image / ms / 1e9:.1f} TFLOP/s')
