"""Inception-v3 genotype classifier: topology description, weight container, BN folding and the
weights blob consumed by dvb_cnn_create (include/dvb.h).

Follows deepvariant/keras_modeling.py:246-336 (`inceptionv3`):
    backbone = tf.keras.applications.InceptionV3(include_top=False, weights=None,
                                                 input_shape=(H, W, C), pooling='avg')
    head     = Dropout(0.2) -> Dense(3, activation='softmax', dtype=float32)   (:46-67)
The backbone's arithmetic lives in the third-party dependency tf_keras==2.16.0
(tf_keras/src/applications/inception_v3.py, not vendored in the reference); its published
topology is restated in `conv_specs()` / `inception_v3_graph()` below:
    conv2d_bn = Conv2D(use_bias=False) -> BatchNormalization(axis=3, scale=False, eps=1e-3) -> ReLU
94 convolutions, 4 max-pools (3x3 s2 valid), 9 avg-pools (3x3 s1 'same', padding excluded from
the divisor), 15 concats, global average pool.

The graph is a flat op list over named tensors; both the torch fp32 oracle (tests/cnn_oracle.py)
and the CUDA engine (csrc/dvb_cnn.cu, which rebuilds the same list in C++) walk it in this order,
and the weights blob stores the convolutions in exactly this order.
"""
from __future__ import annotations

import dataclasses
import struct
from typing import Dict, List, Optional, Tuple

import numpy as np

BN_EPS = 1e-3   # tf_keras BatchNormalization default in inception_v3.conv2d_bn
NUM_CLASSES = 3  # deepvariant/dv_constants.py:77
BLOB_MAGIC = 0x31424E4E  # 'NNB1'


@dataclasses.dataclass
class Op:
  kind: str                 # 'conv' | 'maxpool' | 'avgpool'
  src: str
  dst: str
  dst_channel_offset: int = 0
  cin: int = 0
  cout: int = 0
  kh: int = 1
  kw: int = 1
  stride: int = 1
  same: bool = True         # 'same' (stride 1 only) or 'valid'
  name: str = ''


def inception_v3_graph(in_channels: int) -> Tuple[List[Op], Dict[str, int]]:
  """Returns (ops, tensor_channels).  Tensor 'input' has `in_channels` channels."""
  ops: List[Op] = []
  ch: Dict[str, int] = {'input': in_channels}
  counter = [0]

  def conv(src, dst, cout, kh, kw, stride=1, same=True, off=0, dst_total=None):
    counter[0] += 1
    if dst not in ch:
      ch[dst] = dst_total if dst_total is not None else cout
    ops.append(Op('conv', src, dst, off, ch[src], cout, kh, kw, stride, same, f'conv{counter[0]}'))

  def pool(kind, src, dst, off=0, dst_total=None):
    if dst not in ch:
      ch[dst] = dst_total if dst_total is not None else ch[src]
    ops.append(Op(kind, src, dst, off, ch[src], ch[src], 3, 3, 2 if kind == 'maxpool' else 1,
                  kind == 'avgpool', kind))

  # stem
  conv('input', 's1', 32, 3, 3, stride=2, same=False)
  conv('s1', 's2', 32, 3, 3, same=False)
  conv('s2', 's3', 64, 3, 3)
  pool('maxpool', 's3', 'p1')
  conv('p1', 's4', 80, 1, 1, same=False)
  conv('s4', 's5', 192, 3, 3, same=False)
  pool('maxpool', 's5', 'p2')
  x = 'p2'
  # mixed 0, 1, 2
  for i, pool_ch in enumerate((32, 64, 64)):
    m = f'mixed{i}'
    total = 64 + 64 + 96 + pool_ch
    conv(x, m, 64, 1, 1, off=0, dst_total=total)
    conv(x, f'{m}_b5a', 48, 1, 1)
    conv(f'{m}_b5a', m, 64, 5, 5, off=64)
    conv(x, f'{m}_d1', 64, 1, 1)
    conv(f'{m}_d1', f'{m}_d2', 96, 3, 3)
    conv(f'{m}_d2', m, 96, 3, 3, off=128)
    pool('avgpool', x, f'{m}_ap')
    conv(f'{m}_ap', m, pool_ch, 1, 1, off=224)
    x = m
  # mixed 3
  total = 384 + 96 + ch[x]
  conv(x, 'mixed3', 384, 3, 3, stride=2, same=False, off=0, dst_total=total)
  conv(x, 'mixed3_d1', 64, 1, 1)
  conv('mixed3_d1', 'mixed3_d2', 96, 3, 3)
  conv('mixed3_d2', 'mixed3', 96, 3, 3, stride=2, same=False, off=384)
  pool('maxpool', x, 'mixed3', off=480)
  x = 'mixed3'
  # mixed 4..7
  for i, c7 in zip((4, 5, 6, 7), (128, 160, 160, 192)):
    m = f'mixed{i}'
    conv(x, m, 192, 1, 1, off=0, dst_total=768)
    conv(x, f'{m}_s1', c7, 1, 1)
    conv(f'{m}_s1', f'{m}_s2', c7, 1, 7)
    conv(f'{m}_s2', m, 192, 7, 1, off=192)
    conv(x, f'{m}_d1', c7, 1, 1)
    conv(f'{m}_d1', f'{m}_d2', c7, 7, 1)
    conv(f'{m}_d2', f'{m}_d3', c7, 1, 7)
    conv(f'{m}_d3', f'{m}_d4', c7, 7, 1)
    conv(f'{m}_d4', m, 192, 1, 7, off=384)
    pool('avgpool', x, f'{m}_ap')
    conv(f'{m}_ap', m, 192, 1, 1, off=576)
    x = m
  # mixed 8
  total = 320 + 192 + ch[x]
  conv(x, 'mixed8_a1', 192, 1, 1)
  conv('mixed8_a1', 'mixed8', 320, 3, 3, stride=2, same=False, off=0, dst_total=total)
  conv(x, 'mixed8_b1', 192, 1, 1)
  conv('mixed8_b1', 'mixed8_b2', 192, 1, 7)
  conv('mixed8_b2', 'mixed8_b3', 192, 7, 1)
  conv('mixed8_b3', 'mixed8', 192, 3, 3, stride=2, same=False, off=320)
  pool('maxpool', x, 'mixed8', off=512)
  x = 'mixed8'
  # mixed 9, 10
  for i in (9, 10):
    m = f'mixed{i}'
    conv(x, m, 320, 1, 1, off=0, dst_total=2048)
    conv(x, f'{m}_t1', 384, 1, 1)
    conv(f'{m}_t1', m, 384, 1, 3, off=320)
    conv(f'{m}_t1', m, 384, 3, 1, off=704)
    conv(x, f'{m}_d1', 448, 1, 1)
    conv(f'{m}_d1', f'{m}_d2', 384, 3, 3)
    conv(f'{m}_d2', m, 384, 1, 3, off=1088)
    conv(f'{m}_d2', m, 384, 3, 1, off=1472)
    pool('avgpool', x, f'{m}_ap')
    conv(f'{m}_ap', m, 192, 1, 1, off=1856)
    x = m
  assert sum(1 for o in ops if o.kind == 'conv') == 94
  assert ch['mixed10'] == 2048
  return ops, ch


def out_hw(op: Op, h: int, w: int) -> Tuple[int, int]:
  if op.same:
    return h, w
  return (h - op.kh) // op.stride + 1, (w - op.kw) // op.stride + 1


def conv_flops_per_image(height: int, width: int, in_channels: int) -> float:
  ops, _ = inception_v3_graph(in_channels)
  hw = {'input': (height, width)}
  macs = 0
  for o in ops:
    h, w = hw[o.src]
    oh, ow = out_hw(o, h, w)
    hw[o.dst] = (oh, ow)
    if o.kind == 'conv':
      macs += oh * ow * o.cout * o.kh * o.kw * o.cin
  return 2.0 * macs


@dataclasses.dataclass
class ConvParams:
  kernel: np.ndarray        # [kh, kw, cin, cout] float32 (Keras layout)
  beta: np.ndarray          # [cout]
  moving_mean: np.ndarray   # [cout]
  moving_variance: np.ndarray  # [cout]


@dataclasses.dataclass
class ModelWeights:
  in_channels: int
  convs: List[ConvParams]
  dense_kernel: np.ndarray  # [2048, 3]
  dense_bias: np.ndarray    # [3]


def random_weights(in_channels: int, seed: int = 0) -> ModelWeights:
  """Seeded stand-in weights (no Inception weights ship with the reference): He-normal convs,
  BN mean ~ N(0, 0.1), var ~ U(0.5, 1.5), beta ~ N(0, 0.1)."""
  rng = np.random.default_rng(seed)
  ops, _ = inception_v3_graph(in_channels)
  convs = []
  for o in ops:
    if o.kind != 'conv':
      continue
    fan_in = o.kh * o.kw * o.cin
    convs.append(ConvParams(
        kernel=(rng.standard_normal((o.kh, o.kw, o.cin, o.cout)) * np.sqrt(2.0 / fan_in)).astype(np.float32),
        beta=(rng.standard_normal(o.cout) * 0.1).astype(np.float32),
        moving_mean=(rng.standard_normal(o.cout) * 0.1).astype(np.float32),
        moving_variance=rng.uniform(0.5, 1.5, o.cout).astype(np.float32)))
  dk = (rng.standard_normal((2048, NUM_CLASSES)) * np.sqrt(1.0 / 2048)).astype(np.float32)
  db = (rng.standard_normal(NUM_CLASSES) * 0.1).astype(np.float32)
  return ModelWeights(in_channels, convs, dk, db)


def fold_bn(p: ConvParams) -> Tuple[np.ndarray, np.ndarray]:
  """Conv (no bias) + BN(scale=False, eps=1e-3) -> (kernel', bias') in float32."""
  inv = (1.0 / np.sqrt(p.moving_variance.astype(np.float64) + BN_EPS))
  k = (p.kernel.astype(np.float64) * inv.reshape(1, 1, 1, -1)).astype(np.float32)
  b = (p.beta.astype(np.float64) - p.moving_mean.astype(np.float64) * inv).astype(np.float32)
  return k, b


def pad_cin(cin: int) -> int:
  """Channel count as stored on the device: a multiple of 8 (TMA strides are 16-byte multiples);
  the 7-channel input is padded to 16."""
  return 16 if cin < 16 else (cin + 7) // 8 * 8


BLOB_MAGIC_SPLIT = 0x32424E4E  # 'NNB2': main + residual fp16 planes per kernel
SPLIT_SCALE = 2048.0           # residual plane = fp16((k - fp16(k)) * 2^11)  (csrc/dvb_cnn.cu "precision = 1")


def pack_weights(w: ModelWeights, precision: int = 0) -> bytes:
  """Weights blob for dvb_cnn_create: header {magic, in_channels, n_conv}, then per conv (network
  order) {kh, kw, cin, cin_pad, cout} int32 + fp16 kernel [cout][kh][kw][cin_pad] (BN folded,
  zero padded) [+ the fp16 residual plane of the same shape when precision == 1] + fp32 bias[cout];
  then fp32 dense kernel [2048][3] and bias [3]."""
  ops, _ = inception_v3_graph(w.in_channels)
  conv_ops = [o for o in ops if o.kind == 'conv']
  assert len(conv_ops) == len(w.convs)
  out = bytearray(struct.pack('<3i', BLOB_MAGIC_SPLIT if precision == 1 else BLOB_MAGIC, w.in_channels, len(conv_ops)))
  for o, p in zip(conv_ops, w.convs):
    assert p.kernel.shape == (o.kh, o.kw, o.cin, o.cout), (o.name, p.kernel.shape)
    k, b = fold_bn(p)
    cp = pad_cin(o.cin)
    k32 = np.zeros((o.cout, o.kh, o.kw, cp), dtype=np.float32)
    k32[..., :o.cin] = np.transpose(k, (3, 0, 1, 2))
    kk = k32.astype(np.float16)
    out += struct.pack('<5i', o.kh, o.kw, o.cin, cp, o.cout)
    out += kk.tobytes()
    if precision == 1:
      out += ((k32 - kk.astype(np.float32)) * np.float32(SPLIT_SCALE)).astype(np.float16).tobytes()
    out += b.astype(np.float32).tobytes()
  out += w.dense_kernel.astype(np.float32).tobytes()
  out += w.dense_bias.astype(np.float32).tobytes()
  return bytes(out)


def save_npz(path: str, w: ModelWeights) -> None:
  d = {'in_channels': np.array(w.in_channels), 'dense_kernel': w.dense_kernel, 'dense_bias': w.dense_bias}
  for i, p in enumerate(w.convs):
    d[f'c{i}_kernel'], d[f'c{i}_beta'] = p.kernel, p.beta
    d[f'c{i}_mean'], d[f'c{i}_var'] = p.moving_mean, p.moving_variance
  np.savez(path, **d)


def load_npz(path: str) -> ModelWeights:
  d = np.load(path)
  n = sum(1 for k in d.files if k.endswith('_kernel') and k.startswith('c'))
  convs = [ConvParams(d[f'c{i}_kernel'], d[f'c{i}_beta'], d[f'c{i}_mean'], d[f'c{i}_var']) for i in range(n)]
  return ModelWeights(int(d['in_channels']), convs, d['dense_kernel'], d['dense_bias'])
