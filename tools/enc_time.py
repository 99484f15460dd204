"""Quick device-side timing of the encoder kernel (development aid; bench.py is the contract)."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deepvariant_b200 import pileup_image as pi, synthetic  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--batch', type=int, default=16384)
ap.add_argument('--steps', type=int, default=10)
ap.add_argument('--warmup', type=int, default=3)
ap.add_argument('--pacbio', action='store_true')
a = ap.parse_args()

o = pi.default_options()
if a.pacbio:
  o.channels = pi.PILEUP_DEFAULT_CHANNELS + ['haplotype', 'supplementary_alignment',
                                             'diff_channels_alternate_allele_1', 'diff_channels_alternate_allele_2']
  o.width = 147
  o.sort_by_haplotypes = True
else:
  o.channels = list(pi.PILEUP_CHANNELS_WITH_INSERT_SIZE)
params = pi.to_params(o)
enc = pi.GpuEncoder(params, 0)
tb = synthetic.make_batch(a.batch, 'cuda:0', width=o.width, hp=a.pacbio)
out = torch.empty((a.batch,) + enc.shape, dtype=torch.uint8, device='cuda:0')
s = torch.cuda.current_stream()
for _ in range(a.warmup):
  enc.encode_device(tb, out, stream=s)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(s)
for _ in range(a.steps):
  enc.encode_device(tb, out, stream=s)
e1.record(s)
torch.cuda.synchronize()
enc.check()
ms = e0.elapsed_time(e1) / a.steps
ab = tb.algorithmic_bytes(enc.image_bytes, o.width)
print(json.dumps({'batch': a.batch, 'ms_per_launch': ms, 'images_per_s': a.batch / ms * 1e3,
                  'algorithmic_GBps': ab / ms / 1e6, 'out_GBps': a.batch * enc.image_bytes / ms / 1e6,
                  'bytes_per_image': ab / a.batch, 'grid_cap': None}))
