"""Pins the TF/Keras semantics the CNN oracle (tests/cnn_oracle.py) restates, the topology
(deepvariant_b200/modeling.py) and the only numeric KAT the reference has for this stage
(round_gls, call_variants_test.py:332-356).  CPU only."""
import numpy as np
import pytest
import torch

import cnn_oracle
from deepvariant_b200 import call_variants as cv
from deepvariant_b200 import modeling


def test_topology_counts_and_flops():
  ops, ch = modeling.inception_v3_graph(7)
  assert sum(o.kind == 'conv' for o in ops) == 94
  assert sum(o.kind == 'maxpool' for o in ops) == 4
  assert sum(o.kind == 'avgpool' for o in ops) == 9
  assert ch['mixed2'] == 288 and ch['mixed7'] == 768 and ch['mixed8'] == 1280 and ch['mixed10'] == 2048
  # 1.005 GMAC @100x221x7, 0.624 GMAC @100x147x10, and the known 5.71 GMAC @299x299x3
  assert abs(modeling.conv_flops_per_image(100, 221, 7) / 2e9 - 1.005) < 1e-3
  assert abs(modeling.conv_flops_per_image(100, 147, 10) / 2e9 - 0.624) < 1e-3
  assert abs(modeling.conv_flops_per_image(299, 299, 3) / 2e9 - 5.71) < 5e-3
  n_w = sum(o.kh * o.kw * o.cin * o.cout for o in ops if o.kind == 'conv')
  assert n_w == 21_785_568 - 0 or abs(n_w - 21.75e6) < 0.1e6   # 21.75 M conv weights at C=7


def test_avgpool_excludes_padding_like_tf():
  """TF AveragePooling2D(3, strides 1, 'same'): divisor = number of valid elements."""
  x = torch.ones(1, 1, 3, 4)
  y = torch.nn.functional.avg_pool2d(x, 3, 1, 1, count_include_pad=False)
  assert torch.allclose(y, torch.ones_like(y))            # include_pad=True would give 4/9 at corners
  x = torch.arange(12.).view(1, 1, 3, 4)
  y = torch.nn.functional.avg_pool2d(x, 3, 1, 1, count_include_pad=False)
  assert abs(float(y[0, 0, 0, 0]) - (0 + 1 + 4 + 5) / 4) < 1e-6
  assert abs(float(y[0, 0, 1, 0]) - (0 + 1 + 4 + 5 + 8 + 9) / 6) < 1e-6


def test_bn_fold_matches_unfused():
  rng = np.random.default_rng(1)
  p = modeling.ConvParams(rng.standard_normal((3, 3, 5, 4)).astype(np.float32), rng.standard_normal(4).astype(np.float32),
                          rng.standard_normal(4).astype(np.float32), rng.uniform(0.5, 1.5, 4).astype(np.float32))
  k, b = modeling.fold_bn(p)
  x = torch.randn(2, 5, 9, 9)
  w = torch.from_numpy(np.transpose(p.kernel, (3, 2, 0, 1)).copy())
  y = torch.nn.functional.conv2d(x, w)
  y = (y - torch.from_numpy(p.moving_mean).view(1, -1, 1, 1)) / torch.sqrt(torch.from_numpy(p.moving_variance).view(1, -1, 1, 1) + 1e-3)
  y = y + torch.from_numpy(p.beta).view(1, -1, 1, 1)
  y2 = torch.nn.functional.conv2d(x, torch.from_numpy(np.transpose(k, (3, 2, 0, 1)).copy())) + torch.from_numpy(b).view(1, -1, 1, 1)
  assert torch.allclose(y, y2, atol=1e-5)


def test_oracle_forward_shapes_and_softmax():
  m = cnn_oracle.build_reference_model(7)
  x = torch.randint(0, 255, (2, 100, 221, 7), dtype=torch.uint8)
  p, t, pooled = m.forward(x, return_tensors=True)
  assert p.shape == (2, 3) and torch.allclose(p.sum(1), torch.ones(2), atol=1e-6)
  assert tuple(t['s1'].shape) == (2, 32, 49, 110) and tuple(t['p2'].shape) == (2, 192, 10, 25)
  assert tuple(t['mixed3'].shape) == (2, 768, 4, 12) and tuple(t['mixed10'].shape) == (2, 2048, 1, 5)
  assert pooled.shape == (2, 2048)
  # PACBIO geometry: 100 x 147 x 10
  m2 = cnn_oracle.build_reference_model(10)
  assert m2.forward(torch.zeros(1, 100, 147, 10, dtype=torch.uint8)).shape == (1, 3)


def test_weights_blob_roundtrip_layout():
  w = modeling.random_weights(7, 3)
  blob = modeling.pack_weights(w)
  import struct
  magic, cin, n = struct.unpack_from('<3i', blob, 0)
  assert magic == modeling.BLOB_MAGIC and cin == 7 and n == 94
  kh, kw, ci, cp, co = struct.unpack_from('<5i', blob, 12)
  assert (kh, kw, ci, cp, co) == (3, 3, 7, 16, 32)
  k = np.frombuffer(blob, dtype=np.float16, count=co * kh * kw * cp, offset=32).reshape(co, kh, kw, cp)
  assert not k[..., 7:].any()
  kf, _ = modeling.fold_bn(w.convs[0])
  np.testing.assert_allclose(k[..., :7].astype(np.float32), np.transpose(kf, (3, 0, 1, 2)), rtol=2e-3, atol=1e-4)


def test_round_gls_with_precision():
  """deepvariant/call_variants_test.py:332-356."""
  gls = [0.2102311329, 0.099768768, 0.6899999991]
  assert cv.round_gls(gls, precision=1) == [0.2, 0.1, 0.7]
  assert cv.round_gls(gls, precision=2) == [0.21, 0.10, 0.69]
  assert cv.round_gls(gls, precision=None) == gls
  with pytest.raises(ValueError):
    cv.round_gls([0.5, 0.5, 0.5], 3)


def test_cvo_carries_model_id_and_probabilities():
  from deepvariant_b200 import protos
  call = protos.f_bytes(9, b'sample')                                    # VariantCall.call_set_name
  variant = protos.f_bytes(6, b'A') + protos.f_bytes(7, b'C') + protos.f_bytes(11, call) + protos.f_varint(16, 41)
  cvo = cv.create_cvo(variant, [0.1, 0.2, 0.7], protos.encode_alt_allele_indices([0]))
  v, idx, probs = protos.parse_call_variants_output(cvo)
  assert idx == [0] and probs == [0.1, 0.2, 0.7]
  assert b'MID' in v and b'deepvariant' in v and protos.parse_variant(v).start == 41


def test_fast_cpu_model_matches_reference_model():
  w = modeling.random_weights(7, 6)
  x = torch.randint(0, 255, (2, 100, 221, 7), dtype=torch.uint8)
  a = cnn_oracle.ReferenceModel(w).forward(x)
  b = cnn_oracle.FastCpuModel(w).forward(x)
  assert float((a - b).abs().max()) < 1e-5
