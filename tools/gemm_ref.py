"""cuBLAS fp16 GEMM time on the implicit-GEMM shapes of the classifier's layers (development aid: a library ceiling to read the
hand-written kernels against; the im2col gather is NOT included - A is a dense [M, K] matrix here)."""
import json
import sys

import torch

SHAPES = [  # name, pixels per image, N, K
    ('conv2 32->32 3x3 @47x108', 47 * 108, 32, 288), ('conv3 32->64 3x3 @47x108', 47 * 108, 64, 288),
    ('conv5 80->192 3x3 @21x51', 21 * 51, 192, 720), ('mixed0 5x5 48->64 @10x25', 250, 64, 1200),
    ('mixed0 3x3 96->96 @10x25', 250, 96, 864), ('mixed0 1x1 192->208 @10x25', 250, 208, 192),
    ('mixed5 1x1 768->704 @4x12', 48, 704, 768), ('mixed5 1x7 160->160 @4x12', 48, 160, 1120),
    ('mixed7 1x7 192->192 @4x12', 48, 192, 1344), ('mixed9 1x1 1280->1344 @1x5', 5, 1344, 1280), ('mixed9 3x3 448->384 @1x5', 5, 384, 4032),
]
n_img = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
out = []
for name, px, n, k in SHAPES:
  m = n_img * px
  a = torch.randn(m, k, device='cuda', dtype=torch.float16)
  b = torch.randn(k, n, device='cuda', dtype=torch.float16)
  for _ in range(3):
    torch.matmul(a, b)
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(5):
    torch.matmul(a, b)
  e1.record()
  torch.cuda.synchronize()
  us = e0.elapsed_time(e1) / 5 * 1e3
  out.append({'layer': name, 'M': m, 'N': n, 'K': k, 'us': round(us, 1), 'tflops': round(2.0 * m * n * k / us / 1e6, 1)})
  del a, b
print(json.dumps(out, indent=1))
