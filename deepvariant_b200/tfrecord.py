"""TFRecord files as the reference reads and writes them between stages — gzip( TFRecord( proto ) ),
sharded `name-00000-of-0000N` — without TensorFlow.

  framing      tensorflow RecordWriter/RecordReader (used by third_party/nucleus/io/tfrecord_writer.cc,
               example_writer.cc:99-115): u64 length, u32 masked crc32c(length), data, u32 masked crc32c(data)
  sharding     third_party/nucleus/io/sharded_file_utils.py: 'x@N.gz' -> x-0000i-of-0000N.gz
  compression  GZIP when the path ends in .gz (third_party/nucleus/io/tfrecord.py:88-93)
The CRC-32C is computed by libdvb.so (dvb_masked_crc32c, slicing-by-8).
"""
from __future__ import annotations

import ctypes as C
import gzip
import os
import re
import struct
from typing import Iterator, List, Optional

from deepvariant_b200 import _lib

_SHARD_SPEC = re.compile(r'^(.*)@(\d+)(\..*)?$')


def masked_crc32c(data: bytes) -> int:
  return int(_lib.lib().dvb_masked_crc32c(data, len(data)))


def is_sharded_spec(path: str) -> bool:
  return _SHARD_SPEC.match(path) is not None


def shard_paths(spec: str) -> List[str]:
  """'a/b@3.gz' -> ['a/b-00000-of-00003.gz', ...]; an unsharded path is returned as is."""
  m = _SHARD_SPEC.match(spec)
  if not m:
    return [spec]
  base, n, suffix = m.group(1), int(m.group(2)), m.group(3) or ''
  return [f'{base}-{i:05d}-of-{n:05d}{suffix}' for i in range(n)]


def shard_path(spec: str, task: int) -> str:
  paths = shard_paths(spec)
  if len(paths) == 1:
    return paths[0]
  return paths[task]


def resolve_input_paths(spec: str) -> List[str]:
  """Accepts a sharded spec, a single file, or 'name.gz' whose shards exist as name-*-of-*.gz / name.gz-*-of-*."""
  if is_sharded_spec(spec):
    return shard_paths(spec)
  if os.path.exists(spec):
    return [spec]
  import glob
  found = sorted(glob.glob(spec + '-?????-of-?????')) or sorted(glob.glob(re.sub(r'(\.[^.]+(\.gz)?)$', r'-?????-of-?????\1', spec)))
  if not found:
    raise FileNotFoundError(spec)
  # one consistent -of-N set: shards left behind by an earlier run with another shard count would otherwise be read twice
  totals = {}
  for f in found:
    m = re.search(r'-(\d{5})-of-(\d{5})', f)
    totals.setdefault(int(m.group(2)), []).append(int(m.group(1)))
  if len(totals) != 1:
    raise ValueError(f'{spec}: shards of several runs are mixed ({sorted(totals)} shard counts): delete the stale ones')
  (n, have), = totals.items()
  if sorted(have) != list(range(n)):
    raise ValueError(f'{spec}: expected shards 0..{n - 1} of {n}, found {sorted(have)}')
  return found


def _open(path: str, mode: str):
  return gzip.open(path, mode) if path.endswith('.gz') else open(path, mode)


class Writer:
  """tfrecord.Writer: write(serialized_proto_bytes)."""

  def __init__(self, path: str, compresslevel: int = 6):
    self.path = path
    self._f = gzip.open(path, 'wb', compresslevel=compresslevel) if path.endswith('.gz') else open(path, 'wb')

  def write(self, record: bytes) -> None:
    header = struct.pack('<Q', len(record))
    self._f.write(header)
    self._f.write(struct.pack('<I', masked_crc32c(header)))
    self._f.write(record)
    self._f.write(struct.pack('<I', masked_crc32c(record)))

  def close(self):
    if self._f:
      self._f.close()
      self._f = None

  def __enter__(self):
    return self

  def __exit__(self, *a):
    self.close()


def _open_for_reading(path: str):
  """gzip by content, not by name: the reference's sharded files are called name.tfrecord.gz-0000i-of-0000N (nucleus reads them
  with compression_type GZIP because '.gz' occurs in the name, third_party/nucleus/io/tfrecord.py:88-93)."""
  with open(path, 'rb') as f:
    magic = f.read(2)
  return gzip.open(path, 'rb') if magic == b'\x1f\x8b' else open(path, 'rb')


def read_records(path: str, check_crc: bool = False) -> Iterator[bytes]:
  with _open_for_reading(path) as f:
    while True:
      header = f.read(12)
      if not header:
        return
      if len(header) < 12:
        raise IOError(f'{path}: truncated TFRecord header')
      (length,), (hcrc,) = struct.unpack('<Q', header[:8]), struct.unpack('<I', header[8:])
      if check_crc and masked_crc32c(header[:8]) != hcrc:
        raise IOError(f'{path}: corrupted record length')
      data = f.read(length)
      footer = f.read(4)
      if len(data) < length or len(footer) < 4:
        raise IOError(f'{path}: truncated TFRecord')
      if check_crc and masked_crc32c(data) != struct.unpack('<I', footer)[0]:
        raise IOError(f'{path}: corrupted record data')
      yield data
