"""Reads the weights of the reference's released models - a TensorFlow SavedModel directory or a `model.ckpt` / `ckpt-N` prefix -
without TensorFlow, into modeling.ModelWeights (what `call_variants --checkpoint` hands to the classifier).

The reference loads them with `tf.saved_model.load(checkpoint_path)` / `model.load_weights(checkpoint_path)`
(deepvariant/call_variants.py:679, 759-762; deepvariant/keras_modeling.py:222-228, 322-336).  Both are TensorFlow *tensor bundles*
(`<prefix>.index` + `<prefix>.data-0000i-of-0000N`; a SavedModel keeps its bundle under `variables/variables`); the format lives
in the third-party dependency tensorflow==2.16.1 (tensorflow/core/util/tensor_bundle/, tensorflow/core/lib/io/table*, not vendored in
the reference) and is restated here from its published layout:

  index file   an immutable sorted table (the LevelDB table format): data blocks of prefix-compressed entries
               [varint shared][varint non_shared][varint value_len][key suffix][value] + restart array, each block followed by a
               1-byte compression tag (0 none, 1 snappy) and a masked CRC-32C; an index block of (separator key -> BlockHandle);
               a 48-byte footer (metaindex handle, index handle, padding, magic 0xdb4775248b80fb57).
  entries      key "" -> BundleHeaderProto{num_shards=1, endianness=2, version=3}; tensor name -> BundleEntryProto{dtype=1, shape=2,
               shard_id=3, offset=4, size=5, crc32c=6 (fixed32, masked)}
  data shards  raw little-endian tensor bytes at [offset, offset + size); DT_STRING: varint64 length per element, a fixed32
               masked CRC-32C of those lengths, then the bytes
  names        key "_CHECKPOINTABLE_OBJECT_GRAPH" holds a TrackableObjectGraph proto: nodes{children{node_id=1, local_name=2},
               attributes{name=1, full_name=2, checkpoint_key=3}}; `full_name` is the Keras variable name ("conv2d_12/kernel",
               "batch_normalization_12/moving_mean", "classification/bias"), `checkpoint_key` the bundle key.

Layer matching: tf_keras names layers by creation order (conv2d, conv2d_1, ... - a process-wide counter, so the first index is
arbitrary); inception_v3.conv2d_bn creates Conv2D then BatchNormalization, in the order modeling.inception_v3_graph lists the 94
convolutions.  Every assignment is shape-checked against that topology.

No released checkpoint ships with the reference, so this reader is pinned only against bundles written by `write_bundle` below
(same layout, prefix compression and multi-block tables exercised; tests/test_tf_checkpoint.py) - "parity unpinned" against
TensorFlow's own writer until one of the released models is available on the box."""
from __future__ import annotations

import os
import re
import struct
from typing import Dict, Iterator, List, Optional, Sequence, Tuple

import numpy as np

from deepvariant_b200 import modeling, protos, tfrecord

TABLE_MAGIC = 0xdb4775248b80fb57
OBJECT_GRAPH_KEY = '_CHECKPOINTABLE_OBJECT_GRAPH'
_DTYPES = {1: np.float32, 2: np.float64, 3: np.int32, 4: np.uint8, 5: np.int16, 6: np.int8, 9: np.int64, 10: np.bool_, 14: None, 19: np.float16}
# 14 = DT_BFLOAT16 (widened on read), 7 = DT_STRING (handled separately)


def _varint(buf: bytes, pos: int) -> Tuple[int, int]:
  shift = val = 0
  while True:
    b = buf[pos]
    pos += 1
    val |= (b & 0x7f) << shift
    if not b & 0x80:
      return val, pos
    shift += 7


def snappy_uncompress(buf: bytes) -> bytes:
  """Snappy block format (tables may be written with kSnappyCompression; the bundle writer itself uses none)."""
  n, pos = _varint(buf, 0)
  out = bytearray()
  while pos < len(buf):
    tag = buf[pos]
    pos += 1
    kind = tag & 3
    if kind == 0:
      ln = tag >> 2
      if ln >= 60:
        nb = ln - 59
        ln = int.from_bytes(buf[pos:pos + nb], 'little')
        pos += nb
      ln += 1
      out += buf[pos:pos + ln]
      pos += ln
      continue
    if kind == 1:
      ln = ((tag >> 2) & 7) + 4
      off = ((tag >> 5) << 8) | buf[pos]
      pos += 1
    elif kind == 2:
      ln = (tag >> 2) + 1
      off = int.from_bytes(buf[pos:pos + 2], 'little')
      pos += 2
    else:
      ln = (tag >> 2) + 1
      off = int.from_bytes(buf[pos:pos + 4], 'little')
      pos += 4
    if off == 0 or off > len(out):
      raise ValueError('corrupt snappy block')
    for _ in range(ln):                 # overlapping copies are legal
      out.append(out[-off])
  if len(out) != n:
    raise ValueError('corrupt snappy block: length mismatch')
  return bytes(out)


def _read_block(data: bytes, offset: int, size: int, verify_crc: bool) -> bytes:
  block, tag = data[offset:offset + size], data[offset + size]
  if verify_crc:
    want = struct.unpack_from('<I', data, offset + size + 1)[0]
    if tfrecord.masked_crc32c(data[offset:offset + size + 1]) != want:
      raise ValueError('table block checksum mismatch')
  if tag == 0:
    return block
  if tag == 1:
    return snappy_uncompress(block)
  raise ValueError(f'unknown table block compression {tag}')


def _block_entries(block: bytes) -> Iterator[Tuple[bytes, bytes]]:
  n_restarts = struct.unpack_from('<I', block, len(block) - 4)[0]
  end = len(block) - 4 - 4 * n_restarts
  pos, key = 0, b''
  while pos < end:
    shared, pos = _varint(block, pos)
    non_shared, pos = _varint(block, pos)
    vlen, pos = _varint(block, pos)
    key = key[:shared] + block[pos:pos + non_shared]
    pos += non_shared
    yield key, block[pos:pos + vlen]
    pos += vlen


def read_table(path: str, verify_crc: bool = True) -> Dict[bytes, bytes]:
  data = open(path, 'rb').read()
  if len(data) < 48 or struct.unpack_from('<Q', data, len(data) - 8)[0] != TABLE_MAGIC:
    raise ValueError(f'{path} is not a TensorFlow table (bad magic)')
  footer = data[-48:]
  _, p = _varint(footer, 0)        # metaindex handle
  _, p = _varint(footer, p)
  idx_off, p = _varint(footer, p)
  idx_size, p = _varint(footer, p)
  out: Dict[bytes, bytes] = {}
  for _, handle in _block_entries(_read_block(data, idx_off, idx_size, verify_crc)):
    off, q = _varint(handle, 0)
    size, _ = _varint(handle, q)
    for k, v in _block_entries(_read_block(data, off, size, verify_crc)):
      out[k] = v
  return out


def _parse_entry(buf: bytes) -> dict:
  e = {'dtype': 0, 'shape': [], 'shard_id': 0, 'offset': 0, 'size': 0, 'crc32c': None}
  for fn, wt, val, _ in protos.iter_fields(buf):
    if fn == 1:
      e['dtype'] = int(val)
    elif fn == 2:
      for f2, w2, v2, _ in protos.iter_fields(bytes(val)):
        if f2 == 2:                       # TensorShapeProto.dim
          size = 0
          for f3, w3, v3, _ in protos.iter_fields(bytes(v2)):
            if f3 == 1:
              size = protos._to_signed64(v3)   # pylint: disable=protected-access
          e['shape'].append(size)
    elif fn == 3:
      e['shard_id'] = int(val)
    elif fn == 4:
      e['offset'] = int(val)
    elif fn == 5:
      e['size'] = int(val)
    elif fn == 6:
      e['crc32c'] = int(val) if isinstance(val, int) else struct.unpack('<I', bytes(val))[0]
  return e


class Bundle:
  """A tensor bundle opened for reading: `keys()`, `tensor(name)`, `object_graph_names()`."""

  def __init__(self, prefix: str, verify_crc: bool = True):
    self.prefix = prefix
    table = read_table(prefix + '.index', verify_crc)
    self.num_shards = 1
    for fn, wt, val, _ in protos.iter_fields(table.get(b'', b'')):
      if fn == 1:
        self.num_shards = int(val)
      elif fn == 2 and int(val) != 0:
        raise ValueError('big-endian tensor bundles are not supported')
    self.entries = {k.decode(): _parse_entry(v) for k, v in table.items() if k != b''}
    self._shards: Dict[int, bytes] = {}
    self.verify_crc = verify_crc

  def keys(self) -> List[str]:
    return sorted(self.entries)

  def _bytes(self, e: dict) -> bytes:
    sid = e['shard_id']
    if sid not in self._shards:
      self._shards[sid] = open(f'{self.prefix}.data-{sid:05d}-of-{self.num_shards:05d}', 'rb').read()
    raw = self._shards[sid][e['offset']:e['offset'] + e['size']]
    if len(raw) != e['size']:
      raise ValueError('tensor bundle data shard is truncated')
    return raw

  def tensor(self, name: str) -> np.ndarray:
    e = self.entries[name]
    raw = self._bytes(e)
    if e['dtype'] == 7:
      raise TypeError(f'{name} is a string tensor: use string_scalar()')
    if self.verify_crc and e['crc32c'] is not None and tfrecord.masked_crc32c(raw) != e['crc32c']:
      raise ValueError(f'checksum mismatch for tensor {name}')
    if e['dtype'] == 14:                      # bfloat16 -> float32
      u = np.frombuffer(raw, dtype=np.uint16).astype(np.uint32) << 16
      return u.view(np.float32).reshape(e['shape'])
    dt = _DTYPES.get(e['dtype'])
    if dt is None:
      raise TypeError(f'unsupported dtype {e["dtype"]} for tensor {name}')
    return np.frombuffer(raw, dtype=dt).reshape(e['shape']).copy()

  def string_scalar(self, name: str) -> bytes:
    raw = self._bytes(self.entries[name])
    n, pos = _varint(raw, 0)
    return raw[pos + 4:pos + 4 + n]          # [varint length][fixed32 crc of the lengths][bytes]

  def object_graph_names(self) -> Dict[str, str]:
    """Keras variable name ("conv2d_3/kernel") -> bundle key, from the TrackableObjectGraph; {} when the bundle has none."""
    if OBJECT_GRAPH_KEY not in self.entries:
      return {}
    out: Dict[str, str] = {}
    for fn, wt, node, _ in protos.iter_fields(self.string_scalar(OBJECT_GRAPH_KEY)):
      if fn != 1:
        continue
      for f2, w2, attr, _ in protos.iter_fields(bytes(node)):
        if f2 != 2:
          continue
        full_name = key = ''
        for f3, w3, v3, _ in protos.iter_fields(bytes(attr)):
          if f3 == 2:
            full_name = bytes(v3).decode()
          elif f3 == 3:
            key = bytes(v3).decode()
        if full_name and key in self.entries and full_name not in out:
          out[full_name] = key
    return out


def resolve_prefix(path: str) -> str:
  """SavedModel directory -> <dir>/variables/variables; directory with one *.index -> that prefix; else `path` is the prefix."""
  if os.path.isdir(path):
    if os.path.exists(os.path.join(path, 'variables', 'variables.index')):
      return os.path.join(path, 'variables', 'variables')
    found = sorted(f for f in os.listdir(path) if f.endswith('.index'))
    if len(found) == 1:
      return os.path.join(path, found[0][:-len('.index')])
    raise FileNotFoundError(f'{path}: no variables/variables.index and no single *.index file')
  if path.endswith('.index'):
    path = path[:-len('.index')]
  if not os.path.exists(path + '.index'):
    raise FileNotFoundError(path + '.index')
  return path


def is_tf_checkpoint(path: str) -> bool:
  try:
    resolve_prefix(path)
    return True
  except (FileNotFoundError, NotADirectoryError):
    return False


def _layer_index(layer: str) -> int:
  m = re.search(r'_(\d+)$', layer)
  return int(m.group(1)) if m else 0


def load_inception_weights(path: str, in_channels: Optional[int] = None, verify_crc: bool = True) -> modeling.ModelWeights:
  """The 94 conv + BN layers and the classification head of deepvariant/keras_modeling.py:inceptionv3 from a tensor bundle."""
  b = Bundle(resolve_prefix(path), verify_crc)
  names = b.object_graph_names()
  if not names:      # name-keyed bundles (tf.compat.v1 Saver): the keys are the variable names themselves
    names = {k: k for k in b.keys()}
  by_layer: Dict[str, Dict[str, str]] = {}
  for full_name, key in names.items():
    parts = full_name.split('/')
    if len(parts) >= 2 and not any(s in full_name for s in ('optimizer', 'OPTIMIZER_SLOT', '.OPTIMIZER')):
      by_layer.setdefault(parts[-2], {})[parts[-1].split(':')[0]] = key
  convs = sorted((l for l in by_layer if re.fullmatch(r'conv2d(_\d+)?', l) and 'kernel' in by_layer[l]), key=_layer_index)
  bns = sorted((l for l in by_layer if re.fullmatch(r'batch_normalization(_\d+)?', l) and 'moving_mean' in by_layer[l]), key=_layer_index)
  if len(convs) != 94 or len(bns) != 94:
    raise ValueError(f'{path}: expected 94 conv2d and 94 batch_normalization layers, found {len(convs)} and {len(bns)}')
  first_kernel = b.tensor(by_layer[convs[0]]['kernel'])
  channels = int(first_kernel.shape[2])
  if in_channels is not None and channels != in_channels:
    raise ValueError(f'model has {channels} input channels, examples have {in_channels}')
  ops = [o for o in modeling.inception_v3_graph(channels)[0] if o.kind == 'conv']
  params = []
  for o, c, n in zip(ops, convs, bns):
    kernel = b.tensor(by_layer[c]['kernel']).astype(np.float32)
    if tuple(kernel.shape) != (o.kh, o.kw, o.cin, o.cout):
      raise ValueError(f'{c}/kernel has shape {kernel.shape}, the topology expects {(o.kh, o.kw, o.cin, o.cout)} for {o.name}')
    bn = by_layer[n]
    if 'gamma' in bn:
      raise ValueError(f'{n} has a gamma: inception_v3.conv2d_bn uses scale=False')
    beta, mean, var = (b.tensor(bn[k]).astype(np.float32) for k in ('beta', 'moving_mean', 'moving_variance'))
    if not beta.shape == mean.shape == var.shape == (o.cout,):
      raise ValueError(f'{n} has shapes {beta.shape}, {mean.shape}, {var.shape}; expected ({o.cout},)')
    params.append(modeling.ConvParams(kernel, beta, mean, var))
  dense = [l for l, v in by_layer.items() if 'kernel' in v and 'bias' in v and b.entries[v['kernel']]['shape'] == [2048, modeling.NUM_CLASSES]]
  if len(dense) != 1:
    raise ValueError(f'{path}: expected one Dense(2048 -> {modeling.NUM_CLASSES}) head, found {dense}')
  head = by_layer[dense[0]]
  return modeling.ModelWeights(channels, params, b.tensor(head['kernel']).astype(np.float32), b.tensor(head['bias']).astype(np.float32))


# ---- writer (exports our weights as a TensorFlow checkpoint; also what the reader's tests are built on) ---------------------------------
def _put_varint(out: bytearray, v: int) -> None:
  while v >= 0x80:
    out.append((v & 0x7f) | 0x80)
    v >>= 7
  out.append(v)


def _build_block(entries: Sequence[Tuple[bytes, bytes]], restart_interval: int = 16) -> bytes:
  out, restarts, prev = bytearray(), [], b''
  for i, (k, v) in enumerate(entries):
    shared = 0
    if i % restart_interval == 0:
      restarts.append(len(out))
    else:
      while shared < min(len(prev), len(k)) and prev[shared] == k[shared]:
        shared += 1
    _put_varint(out, shared)
    _put_varint(out, len(k) - shared)
    _put_varint(out, len(v))
    out += k[shared:] + v
    prev = k
  for r in restarts or [0]:
    out += struct.pack('<I', r)
  out += struct.pack('<I', len(restarts) or 1)
  return bytes(out)


def write_table(path: str, items: Dict[bytes, bytes], block_size: int = 4096) -> None:
  out = bytearray()
  index: List[Tuple[bytes, bytes]] = []

  def emit(block: bytes) -> bytes:
    off = len(out)
    out.extend(block + b'\0')
    out.extend(struct.pack('<I', tfrecord.masked_crc32c(block + b'\0')))
    h = bytearray()
    _put_varint(h, off)
    _put_varint(h, len(block))
    return bytes(h)

  pending: List[Tuple[bytes, bytes]] = []
  size = 0
  for k in sorted(items):
    pending.append((k, items[k]))
    size += len(k) + len(items[k])
    if size >= block_size:
      index.append((pending[-1][0], emit(_build_block(pending))))
      pending, size = [], 0
  if pending:
    index.append((pending[-1][0], emit(_build_block(pending))))
  meta = emit(_build_block([]))
  idx = emit(_build_block(index, restart_interval=1))
  footer = bytearray(meta + idx)
  footer += b'\0' * (40 - len(footer))
  footer += struct.pack('<Q', TABLE_MAGIC)
  out += footer
  with open(path, 'wb') as f:
    f.write(out)


def write_bundle(prefix: str, tensors: Dict[str, np.ndarray], keras_names: Optional[Dict[str, str]] = None) -> None:
  """tensors: bundle key -> array; keras_names: Keras variable name -> bundle key (written as the TrackableObjectGraph)."""
  dtype_code = {np.dtype(v): k for k, v in _DTYPES.items() if v is not None}
  data = bytearray()
  items: Dict[bytes, bytes] = {b'': protos.f_varint(1, 1) + protos.f_bytes(3, protos.f_varint(1, 1))}   # num_shards=1, version{producer=1}

  def entry(dtype: int, shape: Sequence[int], raw: bytes) -> bytes:
    e = protos.f_varint(1, dtype)
    e += protos.f_bytes(2, b''.join(protos.f_bytes(2, protos.f_varint(1, int(s)) if s else b'') for s in shape))
    if len(data):
      e += protos.f_varint(4, len(data))
    e += protos.f_varint(5, len(raw))
    e += bytes([6 << 3 | 5]) + struct.pack('<I', tfrecord.masked_crc32c(raw))
    data.extend(raw)
    return e

  if keras_names:
    nodes = [b'']                        # node 0 = root
    children = b''
    for i, (full_name, key) in enumerate(sorted(keras_names.items())):
      attr = protos.f_bytes(1, b'VARIABLE_VALUE') + protos.f_bytes(2, full_name.encode()) + protos.f_bytes(3, key.encode())
      nodes.append(protos.f_bytes(2, attr))
      children += protos.f_bytes(1, protos.f_varint(1, i + 1) + protos.f_bytes(2, f'v{i}'.encode()))
    nodes[0] = children
    graph = b''.join(protos.f_bytes(1, n) for n in nodes)
    lengths = bytearray()
    _put_varint(lengths, len(graph))
    raw = bytes(lengths) + struct.pack('<I', tfrecord.masked_crc32c(bytes(lengths))) + graph
    e = protos.f_varint(1, 7) + protos.f_bytes(2, b'')
    if len(data):
      e += protos.f_varint(4, len(data))
    e += protos.f_varint(5, len(raw))
    data.extend(raw)
    items[OBJECT_GRAPH_KEY.encode()] = e
  for key in sorted(tensors):
    a = np.ascontiguousarray(tensors[key])
    items[key.encode()] = entry(dtype_code[a.dtype], a.shape, a.tobytes())
  with open(prefix + '.data-00000-of-00001', 'wb') as f:
    f.write(data)
  write_table(prefix + '.index', items)


def save_inception_checkpoint(prefix: str, w: modeling.ModelWeights, first_layer_index: int = 0) -> None:
  """Writes ModelWeights as a Keras object-graph checkpoint with the layer names tf_keras would have given them."""
  tensors: Dict[str, np.ndarray] = {}
  names: Dict[str, str] = {}

  def add(full_name: str, key: str, a: np.ndarray) -> None:
    tensors[key] = np.asarray(a, dtype=np.float32)
    names[full_name] = key

  def layer(base: str, i: int) -> str:
    i += first_layer_index
    return base if i == 0 else f'{base}_{i}'

  for i, p in enumerate(w.convs):
    add(f'{layer("conv2d", i)}/kernel', f'layer_with_weights-{2 * i}/kernel/.ATTRIBUTES/VARIABLE_VALUE', p.kernel)
    for attr, a in (('beta', p.beta), ('moving_mean', p.moving_mean), ('moving_variance', p.moving_variance)):
      add(f'{layer("batch_normalization", i)}/{attr}', f'layer_with_weights-{2 * i + 1}/{attr}/.ATTRIBUTES/VARIABLE_VALUE', a)
  add('classification/kernel', 'layer_with_weights-188/kernel/.ATTRIBUTES/VARIABLE_VALUE', w.dense_kernel)
  add('classification/bias', 'layer_with_weights-188/bias/.ATTRIBUTES/VARIABLE_VALUE', w.dense_bias)
  write_bundle(prefix, tensors, names)
