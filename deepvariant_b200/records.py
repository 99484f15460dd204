"""Record I/O of the call_variants stage on the C++ side of libdvb.so (csrc/dvb_records.cu).

  NativeExamplesReader   call_variants.get_dataset (deepvariant/call_variants.py:449-538): sharded gzip TFRecords of
                         tf.Example -> batches of raw uint8 images in a caller-owned (pinned) buffer + the two proto
                         fields each CallVariantsOutput carries over, in tf.data's deterministic interleave order.
  NativeCvoWriter        round_gls + _create_cvo_proto + write_variant_call (call_variants.py:248-399): one output shard.
  interleave_order       the same order restated in Python for the tests.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple

import numpy as np

from deepvariant_b200 import _lib

DEFAULT_CYCLE_LENGTH = 32   # _DEFAULT_INPUT_READ_THREADS, deepvariant/call_variants.py:83


class BatchMeta:
  """variant/encoded and alt_allele_indices/encoded of one batch as two byte arenas with CSR offsets (copies: safe to keep
  after the reader moves on)."""

  def __init__(self, n: int, variant_blob: np.ndarray, variant_begin: np.ndarray, alt_blob: np.ndarray, alt_begin: np.ndarray):
    self.n = n
    self.variant_blob, self.variant_begin, self.alt_blob, self.alt_begin = variant_blob, variant_begin, alt_blob, alt_begin

  def as_ctypes(self) -> _lib.DvbExampleBatchMeta:
    m = _lib.DvbExampleBatchMeta()
    m.variant_blob, m.variant_begin = self.variant_blob.ctypes.data, self.variant_begin.ctypes.data
    m.alt_blob, m.alt_begin = self.alt_blob.ctypes.data, self.alt_begin.ctypes.data
    return m

  def variants(self) -> List[bytes]:
    b = self.variant_blob.tobytes()
    return [b[int(self.variant_begin[i]):int(self.variant_begin[i + 1])] for i in range(self.n)]

  def alt_allele_indices(self) -> List[bytes]:
    b = self.alt_blob.tobytes()
    return [b[int(self.alt_begin[i]):int(self.alt_begin[i + 1])] for i in range(self.n)]

  @classmethod
  def from_lists(cls, variants: Sequence[bytes], alts: Sequence[bytes]) -> 'BatchMeta':
    def pack(items):
      begin = np.zeros(len(items) + 1, dtype=np.int64)
      np.cumsum([len(x) for x in items], out=begin[1:])
      return np.frombuffer(b''.join(items) or b'\0', dtype=np.uint8).copy(), begin
    vb, vbeg = pack(list(variants))
    ab, abeg = pack(list(alts))
    return cls(len(variants), vb, vbeg, ab, abeg)


class NativeExamplesReader:

  def __init__(self, paths: Sequence[str], threads: int = 0, cycle_length: int = DEFAULT_CYCLE_LENGTH, verify_crc: bool = True):
    self._lib = _lib.lib()
    arr = (C.c_char_p * max(1, len(paths)))(*[p.encode() for p in paths])
    h = C.c_void_p()
    _lib.check(self._lib.dvb_examples_reader_open(arr, len(paths), threads, cycle_length, int(verify_crc), C.byref(h)))
    self._h = h

  def shape(self) -> Tuple[List[int], int]:
    """(image/shape of the first record, byte size of its image/encoded); ([0, 0, 0], 0) when there are no records."""
    s = (C.c_int64 * 3)()
    nb = C.c_int64()
    _lib.check(self._lib.dvb_examples_reader_shape(self._h, s, C.byref(nb)))
    return [int(x) for x in s], int(nb.value)

  def next_into(self, images: np.ndarray, max_n: Optional[int] = None) -> Optional[BatchMeta]:
    """Fills images[:n] (uint8 [B, image_bytes], C-contiguous, e.g. the numpy view of a pinned tensor); None at the end."""
    if images.dtype != np.uint8 or images.ndim != 2 or not images.flags['C_CONTIGUOUS']:
      raise ValueError('images must be a C-contiguous uint8 [B, image_bytes] array')
    cap = images.shape[0] if max_n is None else min(max_n, images.shape[0])
    n = C.c_int32()
    m = _lib.DvbExampleBatchMeta()
    _lib.check(self._lib.dvb_examples_reader_next(self._h, cap, images.ctypes.data, images.shape[1], C.byref(n), C.byref(m)))
    k = n.value
    if k == 0:
      return None

    def arr(ptr, count, dtype):
      if not count:
        return np.zeros(1, dtype=dtype)
      return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), shape=(count * np.dtype(dtype).itemsize,)).view(dtype).copy()
    vbeg, abeg = arr(m.variant_begin, k + 1, np.int64), arr(m.alt_begin, k + 1, np.int64)
    return BatchMeta(k, arr(m.variant_blob, int(vbeg[-1]), np.uint8), vbeg, arr(m.alt_blob, int(abeg[-1]), np.uint8), abeg)

  def close(self) -> None:
    if getattr(self, '_h', None) is not None:
      self._lib.dvb_examples_reader_close(self._h)
      self._h = None

  def __del__(self):
    self.close()

  def __enter__(self):
    return self

  def __exit__(self, *a):
    self.close()


class NativeCvoWriter:

  def __init__(self, path: str, gl_precision: Optional[int] = 10):
    self._lib = _lib.lib()
    self.path = path
    h = C.c_void_p()
    _lib.check(self._lib.dvb_cvo_writer_open(path.encode(), -1 if gl_precision is None else int(gl_precision), C.byref(h)))
    self._h = h
    self.n_written = 0

  def write_batch(self, meta: BatchMeta, probs: np.ndarray) -> None:
    probs = np.ascontiguousarray(probs, dtype=np.float32)
    if probs.shape != (meta.n, 3):
      raise ValueError(f'probs has shape {probs.shape}, expected {(meta.n, 3)}')
    m = meta.as_ctypes()
    _lib.check(self._lib.dvb_cvo_writer_write_batch(self._h, meta.n, C.byref(m), probs.ctypes.data))

  def close(self) -> int:
    if self._h is not None:
      h, self._h = self._h, None
      n = C.c_int64()
      _lib.check(self._lib.dvb_cvo_writer_close(h, C.byref(n)))
      self.n_written = int(n.value)
    return self.n_written

  def __del__(self):
    try:
      self.close()
    except Exception:  # pylint: disable=broad-except
      pass


def interleave_order(shard_sizes: Sequence[int], cycle_length: int = DEFAULT_CYCLE_LENGTH) -> List[Tuple[int, int]]:
  """(shard, record) pairs in the order tf.data's deterministic interleave (block_length 1) emits them: cycle_length slots,
  one record per slot per turn; an exhausted slot is cleared and refilled with the next shard when the cycle comes round
  to it again (tensorflow/core/kernels/data/interleave_dataset_op.cc GetNextInternal)."""
  n = len(shard_sizes)
  slots: List[Optional[List[int]]] = [None] * cycle_length   # [shard, next record]
  out: List[Tuple[int, int]] = []
  ci = next_file = num_open = 0
  while next_file < n or num_open > 0:
    s = slots[ci]
    if s is not None:
      if s[1] < shard_sizes[s[0]]:
        out.append((s[0], s[1]))
        s[1] += 1
      else:
        slots[ci] = None
        num_open -= 1
      ci = (ci + 1) % cycle_length
    elif next_file < n:
      slots[ci] = [next_file, 0]
      next_file += 1
      num_open += 1
    else:
      ci = (ci + 1) % cycle_length
  return out
