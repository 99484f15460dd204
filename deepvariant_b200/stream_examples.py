"""The reference's shared-memory example stream (make_examples --stream_examples -> call_variants --stream_examples, orchestrated by
fast_pipeline), both ends, over csrc/dvb_stream.cu:

  StreamOrchestrator   fast_pipeline.cc:125-165      creates / removes the per-shard buffer and its three named mutexes
  StreamProducer       stream_examples.cc:60-176     StartStreaming / StreamExample / EndStreaming / SignalShardFinished
  StreamConsumer       stream_examples_kernel.cc:166-240   Next(): the records of whichever shard has a buffer ready

Same names in /dev/shm as the reference's binaries use, so either end can be the reference's own process.  In this repository the
in-process fused flow (deepvariant_b200/fused.py) is the fast path; the stream is the drop-in boundary for mixed deployments."""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Tuple

import numpy as np

from deepvariant_b200 import _lib

ORCHESTRATOR, PRODUCER, CONSUMER = 0, 1, 2


class _Handle:

  def __init__(self, prefix: str, shard: int, role: int, buffer_size: int = 0):
    self._lib = _lib.lib()
    h = C.c_void_p()
    _lib.check(self._lib.dvb_stream_open(prefix.encode(), shard, role, buffer_size, C.byref(h)))
    self._h, self.prefix, self.shard = h, prefix, shard

  @property
  def buffer_size(self) -> int:
    return int(self._lib.dvb_stream_buffer_size(self._h))

  def close(self):
    if getattr(self, '_h', None):
      self._lib.dvb_stream_close(self._h)
      self._h = None

  def __del__(self):
    try:
      self.close()
    except Exception:  # pylint: disable=broad-except
      pass


class StreamOrchestrator:
  """Owns the shared objects of `num_shards` shards for the lifetime of a run."""

  def __init__(self, prefix: str, num_shards: int, buffer_size: int):
    self.prefix, self.num_shards = prefix, num_shards
    for shard in range(num_shards):
      _lib.lib().dvb_stream_remove(prefix.encode(), shard)          # a stale run's objects would carry stale mutex states
    self._handles = [_Handle(prefix, shard, ORCHESTRATOR, buffer_size) for shard in range(num_shards)]

  def remove(self):
    for h in self._handles:
      h.close()
    for shard in range(self.num_shards):
      _lib.lib().dvb_stream_remove(self.prefix.encode(), shard)


class StreamProducer(_Handle):

  def __init__(self, prefix: str, shard: int):
    super().__init__(prefix, shard, PRODUCER)

  def start_streaming(self):
    _lib.check(self._lib.dvb_stream_start(self._h))

  def stream_example(self, alt_indices_encoded: bytes, variant_encoded: bytes, image: np.ndarray):
    image = np.ascontiguousarray(image, dtype=np.uint8)
    _lib.check(self._lib.dvb_stream_put(self._h, alt_indices_encoded, len(alt_indices_encoded), variant_encoded, len(variant_encoded),
                                        C.c_void_p(image.ctypes.data), image.size))

  def end_streaming(self, data_written: bool):
    _lib.check(self._lib.dvb_stream_end(self._h, int(bool(data_written))))

  def signal_shard_finished(self):
    _lib.check(self._lib.dvb_stream_shard_finished(self._h))


class StreamConsumer:
  """All shards of a run from the call_variants side."""

  def __init__(self, prefix: str, num_shards: int, image_shape: Tuple[int, int, int], images_out: Optional[np.ndarray] = None,
               attach_timeout_s: float = 600.0):
    self._lib = _lib.lib()
    self._handles = [_Handle(prefix, shard, CONSUMER) for shard in range(num_shards)]
    for h in self._handles:       # every producer must hold its mutexes before the first poll (dvb_stream_wait_attached)
      _lib.check(self._lib.dvb_stream_wait_attached(h._h, int(attach_timeout_s * 1000)))  # pylint: disable=protected-access
    self._arr = (C.c_void_p * num_shards)(*[h._h for h in self._handles])  # pylint: disable=protected-access
    self.image_shape = tuple(int(x) for x in image_shape)
    self.image_bytes = int(np.prod(self.image_shape))
    cap = max(1, self._handles[0].buffer_size // max(self.image_bytes, 1))
    # one buffer's worth of images; pass a pinned array to hand the images to the GPU without another copy
    self.images = images_out if images_out is not None else np.empty((cap,) + self.image_shape, dtype=np.uint8)
    self._index = 0

  def next(self):
    """(images uint8[n, H, W, C], variants [bytes], alt_allele_indices [bytes]) of one drained buffer; None when every shard finished."""
    n, shard, done = C.c_int32(0), C.c_int32(-1), C.c_int32(0)
    meta = _lib.DvbExampleBatchMeta()
    while True:
      _lib.check(self._lib.dvb_stream_next(self._arr, len(self._handles), self._index, C.c_void_p(self.images.ctypes.data), self.images.nbytes,
                                           self.image_bytes, C.byref(n), C.byref(shard), C.byref(meta), C.byref(done)))
      self._index += 1
      if done.value:
        return None
      if n.value:
        break
    k = n.value
    vb = np.ctypeslib.as_array(C.cast(meta.variant_begin, C.POINTER(C.c_int64)), shape=(k + 1,))
    ab = np.ctypeslib.as_array(C.cast(meta.alt_begin, C.POINTER(C.c_int64)), shape=(k + 1,))
    vblob = C.string_at(meta.variant_blob, int(vb[k])) if vb[k] else b''
    ablob = C.string_at(meta.alt_blob, int(ab[k])) if ab[k] else b''
    variants = [vblob[vb[i]:vb[i + 1]] for i in range(k)]
    alts = [ablob[ab[i]:ab[i + 1]] for i in range(k)]
    return self.images[:k], variants, alts

  def close(self):
    for h in self._handles:
      h.close()
