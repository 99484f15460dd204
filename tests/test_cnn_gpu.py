"""CUDA classifier (tcgen05 implicit-GEMM convs) vs the torch fp32 oracle.  `-m gpu`.

Tolerances: operands are fp16 (the configuration BASELINE.json names: "Inception-v3 fp16 inference"),
accumulation fp32.  Per-layer activations are compared with a relative-to-scale tolerance that grows
with depth; final probabilities within 5e-3 of the fp32 oracle (measured: see DESIGN.md)."""
import numpy as np
import pytest
import torch

import cnn_oracle
from deepvariant_b200 import call_variants as cv
from deepvariant_b200 import modeling

pytestmark = pytest.mark.gpu


def _images(n, shape, seed=0):
  """Pileup-like uint8 images: real encoder output when the geometry is WGS, else random."""
  g = torch.Generator().manual_seed(seed)
  return torch.randint(0, 255, (n,) + tuple(shape), dtype=torch.uint8, generator=g)


def _check_layers(shape, n, names, seed=0):
  w = modeling.random_weights(shape[2], seed)
  net = cv.GpuCnn(w, shape, device=0, max_batch=n)
  imgs = _images(n, shape, seed)
  probs = torch.empty((n, 3), dtype=torch.float32, device='cuda:0')
  net.forward_device(imgs.to('cuda:0'), probs)
  torch.cuda.synchronize()
  want_p, tensors, pooled = cnn_oracle.ReferenceModel(w).forward(imgs, return_tensors=True)
  report = []
  worst = 0.0
  for name in names:
    got = net.debug_tensor(name, n)
    ref = tensors[name].permute(0, 2, 3, 1).numpy()
    assert got.shape == ref.shape, (name, got.shape, ref.shape)
    scale = max(float(np.abs(ref).max()), 1e-6)
    err = float(np.abs(got - ref).max()) / scale
    report.append((name, err))
    worst = max(worst, err)
  return report, worst, probs.cpu(), want_p, net, pooled


STEM = ['s1', 's2', 's3', 'p1', 's4', 's5', 'p2']


def test_stem_layers_match_oracle():
  report, worst, _, _, _, _ = _check_layers((100, 221, 7), 3, STEM)
  print(report)
  for name, err in report:
    assert err < 6e-3, report


def test_stem_patches_are_exact():
  """Tensor 'input' = preprocess + im2col of conv1: patch k = (r*3+s)*C + c holds (x[2oh+r, 2ow+s, c] - 128)/128 exactly."""
  shape = (100, 221, 7)
  net = cv.GpuCnn(modeling.random_weights(7, 0), shape, device=0, max_batch=2)
  imgs = _images(2, shape, 9)
  probs = torch.empty((2, 3), dtype=torch.float32, device='cuda:0')
  net.forward_device(imgs.to('cuda:0'), probs)
  torch.cuda.synchronize()
  got = net.debug_tensor('input', 2)
  assert got.shape == (2, 49, 110, 64)
  x = (imgs.numpy().astype(np.float32) - 128.0) / 128.0
  for r in range(3):
    for s in range(3):
      want = x[:, r:r + 2 * 49:2, s:s + 2 * 110:2, :]
      np.testing.assert_array_equal(got[..., (r * 3 + s) * 7:(r * 3 + s) * 7 + 7], want)
  assert not got[..., 63:].any()


def test_all_block_outputs_match_oracle():
  names = STEM + [f'mixed{i}' for i in range(11)]
  report, worst, got_p, want_p, net, pooled = _check_layers((100, 221, 7), 4, names, seed=1)
  print(report)
  for name, err in report:
    assert err < 2e-2, report
  assert torch.allclose(got_p.sum(1), torch.ones(4), atol=1e-6)
  assert float((got_p - want_p).abs().max()) < 5e-3
  got_pooled = net.debug_tensor('pooled', 4).reshape(4, 2048)
  assert float(np.abs(got_pooled - pooled.numpy()).max()) < 5e-2


def test_branch_tensors_match_oracle():
  names = ['mixed0_b5a', 'mixed0_d2', 'mixed0_ap', 'mixed3_d2', 'mixed4_s2', 'mixed4_d4', 'mixed4_ap', 'mixed8_b3', 'mixed9_t1',
           'mixed9_d2', 'mixed10_ap']
  report, worst, _, _, _, _ = _check_layers((100, 221, 7), 2, names, seed=2)
  print(report)
  for name, err in report:
    assert err < 2e-2, report


def test_pacbio_geometry():
  report, worst, got_p, want_p, _, _ = _check_layers((100, 147, 10), 3, ['s1', 'p2', 'mixed3', 'mixed8', 'mixed10'], seed=3)
  print(report)
  assert worst < 2e-2 and float((got_p - want_p).abs().max()) < 5e-3


def test_batch_chunking_and_host_entry_point():
  shape = (100, 221, 7)
  w = modeling.random_weights(7, 4)
  net = cv.GpuCnn(w, shape, device=0, max_batch=5)       # 13 images -> chunks of 5, 5, 3
  imgs = _images(13, shape, 4)
  got = net.forward_host(imgs.numpy())
  want = cnn_oracle.ReferenceModel(w).forward(imgs).numpy()
  assert np.abs(got - want).max() < 5e-3
  probs = torch.empty((13, 3), dtype=torch.float32, device='cuda:0')
  net.forward_device(imgs.to('cuda:0'), probs)
  torch.cuda.synchronize()
  np.testing.assert_array_equal(probs.cpu().numpy(), got)   # deterministic, chunk-invariant
  for row in got:
    cv.round_gls(row.astype(np.float64).tolist(), 10)        # sums to 1 within 1e-6


def test_encoder_output_feeds_cnn_in_place():
  from deepvariant_b200 import pileup_image as pi, synthetic
  o = pi.default_options()
  o.channels = list(pi.PILEUP_CHANNELS_WITH_INSERT_SIZE)
  enc = pi.GpuEncoder(pi.to_params(o), 0)
  tb = synthetic.make_batch(24, 'cuda:0')
  images = torch.empty((24,) + enc.shape, dtype=torch.uint8, device='cuda:0')
  enc.encode_device(tb, images)
  w = modeling.random_weights(7, 5)
  net = cv.GpuCnn(w, enc.shape, device=0, max_batch=24)
  probs = torch.empty((24, 3), dtype=torch.float32, device='cuda:0')
  net.forward_device(images, probs)
  torch.cuda.synchronize()
  want = cnn_oracle.ReferenceModel(w).forward(images.cpu())
  assert float((probs.cpu() - want).abs().max()) < 5e-3
