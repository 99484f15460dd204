"""CNN ORACLE (TEST INFRASTRUCTURE): plain PyTorch fp32 restatement of the reference classifier.

  preprocess   deepvariant/dv_utils.py:356-380        (float32(x) - 128) / 128
  backbone     deepvariant/keras_modeling.py:268-274  tf.keras.applications.InceptionV3(
                                                      include_top=False, pooling='avg')
               (tf_keras==2.16.0 inception_v3.py — third-party, not in /root/reference; its
               published topology is restated in deepvariant_b200/modeling.py)
  head         deepvariant/keras_modeling.py:46-67    Dropout (inference no-op) + Dense(3, softmax) fp32

PARITY UNPINNED BY THE REFERENCE: the reference's tests use random weights and assert only shapes
and counts (call_variants_test.py:91-200, keras_modeling_test.py:63-100) and no Inception weights
ship in the repository, so this oracle cannot be checked against reference outputs here.  It is
instead written independently of the CUDA engine (unfolded BatchNorm, NCHW, torch's own conv
kernels) and the TF semantics that matter are pinned by unit tests in tests/test_cnn_oracle.py
('same' padding for stride-1 convs, avg-pool divisor excluding padding, BN eps 1e-3 / no gamma).
"""
from __future__ import annotations

from typing import Dict, List

import numpy as np
import torch
import torch.nn.functional as F

from deepvariant_b200 import modeling


class ReferenceModel:

  def __init__(self, weights: modeling.ModelWeights, device='cpu', dtype=torch.float32):
    self.ops, self.channels = modeling.inception_v3_graph(weights.in_channels)
    self.device = torch.device(device)
    self.dtype = dtype
    self.convs = []
    for p in weights.convs:
      self.convs.append(dict(
          w=torch.from_numpy(np.ascontiguousarray(np.transpose(p.kernel, (3, 2, 0, 1)))).to(self.device, dtype),
          beta=torch.from_numpy(p.beta).to(self.device, dtype),
          mean=torch.from_numpy(p.moving_mean).to(self.device, dtype),
          var=torch.from_numpy(p.moving_variance).to(self.device, dtype)))
    self.dense_w = torch.from_numpy(weights.dense_kernel).to(self.device, dtype)
    self.dense_b = torch.from_numpy(weights.dense_bias).to(self.device, dtype)

  def eval(self):
    return self

  @torch.no_grad()
  def forward(self, images_u8: torch.Tensor, return_tensors: bool = False):
    """images_u8: uint8 [N, H, W, C] -> probabilities float [N, 3]."""
    x = images_u8.to(self.device).to(self.dtype)
    x = (x - 128.0) / 128.0                      # dv_utils.preprocess_images
    t: Dict[str, torch.Tensor] = {'input': x.permute(0, 3, 1, 2).contiguous()}
    parts: Dict[str, Dict[int, torch.Tensor]] = {}
    ci = 0
    for o in self.ops:
      src = self._get(t, parts, o.src)
      if o.kind == 'conv':
        c = self.convs[ci]
        ci += 1
        pad = ((o.kh - 1) // 2, (o.kw - 1) // 2) if o.same else (0, 0)
        y = F.conv2d(src, c['w'], None, stride=o.stride, padding=pad)
        y = (y - c['mean'].view(1, -1, 1, 1)) / torch.sqrt(c['var'].view(1, -1, 1, 1) + modeling.BN_EPS)
        y = F.relu(y + c['beta'].view(1, -1, 1, 1))
      elif o.kind == 'maxpool':
        y = F.max_pool2d(src, 3, 2, 0)
      else:  # TF AveragePooling2D(padding='same') excludes padding from the divisor
        y = F.avg_pool2d(src, 3, 1, 1, count_include_pad=False)
      parts.setdefault(o.dst, {})[o.dst_channel_offset] = y
    feat = self._get(t, parts, 'mixed10')
    pooled = feat.mean(dim=(2, 3))               # GlobalAveragePooling2D
    logits = pooled @ self.dense_w + self.dense_b
    probs = torch.softmax(logits.float(), dim=1)
    if return_tensors:
      for name in list(parts):
        self._get(t, parts, name)
      return probs, t, pooled
    return probs

  def _get(self, t, parts, name):
    if name not in t:
      pieces = parts[name]
      t[name] = torch.cat([pieces[k] for k in sorted(pieces)], dim=1)
      assert t[name].shape[1] == self.channels[name], name
    return t[name]


def build_reference_model(in_channels: int, seed: int = 0, device='cpu') -> ReferenceModel:
  return ReferenceModel(modeling.random_weights(in_channels, seed), device)


def predict(model: ReferenceModel, images_u8: torch.Tensor) -> torch.Tensor:
  return model.forward(images_u8)


class FastCpuModel:
  """Same arithmetic as ReferenceModel with BN folded and channels_last tensors so that oneDNN's NHWC
  kernels are used — the fairest stand-in available here for the reference's TF-CPU (oneDNN)
  call_variants when bench.py times the CPU baseline.  Checked against ReferenceModel in
  tests/test_cnn_oracle.py."""

  def __init__(self, weights: modeling.ModelWeights):
    self.ops, _ = modeling.inception_v3_graph(weights.in_channels)
    self.convs = []
    for p in weights.convs:
      k, b = modeling.fold_bn(p)
      wt = torch.from_numpy(np.ascontiguousarray(np.transpose(k, (3, 2, 0, 1)))).contiguous(memory_format=torch.channels_last)
      self.convs.append((wt, torch.from_numpy(b)))
    self.dense_w = torch.from_numpy(weights.dense_kernel)
    self.dense_b = torch.from_numpy(weights.dense_bias)

  def forward(self, images_u8: torch.Tensor) -> torch.Tensor:
    with torch.inference_mode():
      x = ((images_u8.float() - 128.0) / 128.0).permute(0, 3, 1, 2).contiguous(memory_format=torch.channels_last)
      t = {'input': x}
      parts = {}

      def get(name):
        if name not in t:
          t[name] = torch.cat([parts[name][k] for k in sorted(parts[name])], 1)
        return t[name]

      ci = 0
      for o in self.ops:
        s = get(o.src)
        if o.kind == 'conv':
          wt, b = self.convs[ci]
          ci += 1
          pad = ((o.kh - 1) // 2, (o.kw - 1) // 2) if o.same else (0, 0)
          y = F.relu(F.conv2d(s, wt, b, stride=o.stride, padding=pad))
        elif o.kind == 'maxpool':
          y = F.max_pool2d(s, 3, 2, 0)
        else:
          y = F.avg_pool2d(s, 3, 1, 1, count_include_pad=False)
        parts.setdefault(o.dst, {})[o.dst_channel_offset] = y
      f = get('mixed10').mean((2, 3))
      return torch.softmax(f @ self.dense_w + self.dense_b, 1)
