"""deepvariant/realigner/python/ssw (pybind of ssw.h) mirrored over dvb_ssw_align (csrc/dvb_ssw.cu)."""
from __future__ import annotations

import ctypes as C
import dataclasses

from deepvariant_b200 import _lib


@dataclasses.dataclass
class Alignment:
  sw_score: int = 0
  ref_begin: int = 0
  ref_end: int = 0
  query_begin: int = 0
  query_end: int = 0
  mismatches: int = 0
  cigar_string: str = ''


class Aligner:
  """ssw.Aligner: defaults of StripedSmithWaterman::Aligner (match 2, mismatch 2, gap open 3, gap extend 1)."""

  def __init__(self, match_score: int = 2, mismatch_penalty: int = 2, gap_opening_penalty: int = 3, gap_extending_penalty: int = 1):
    self.params = (match_score, mismatch_penalty, gap_opening_penalty, gap_extending_penalty)
    self._ref = b''
    self._lib = _lib.lib()

  def set_reference_sequence(self, ref: str) -> int:
    self._ref = ref.encode() if isinstance(ref, str) else bytes(ref)
    return len(self._ref)

  def align(self, query: str) -> Alignment:
    q = query.encode() if isinstance(query, str) else bytes(query)
    out = _lib.DvbSswAlignment()
    cap = 16 * (len(q) + len(self._ref)) + 64
    buf = C.create_string_buffer(cap)
    _lib.check(self._lib.dvb_ssw_align(self._ref, len(self._ref), q, len(q), *self.params, C.byref(out), buf, cap))
    return Alignment(out.sw_score, out.ref_begin, out.ref_end, out.query_begin, out.query_end, out.mismatches, buf.value.decode())


def align_batch(pairs, match_score: int = 2, mismatch_penalty: int = 2, gap_opening_penalty: int = 3, gap_extending_penalty: int = 1,
                device: int = 0):
  """[(reference, query), ...] -> [Alignment, ...] in one launch (dvb_ssw_align_batch: the Smith-Waterman scans on the GPU, a warp per
  pair; the banded traceback on the host).  Field for field what Aligner.align returns for each pair.  Raises without a CUDA device."""
  import numpy as np
  n = len(pairs)
  if n == 0:
    return []
  refs = [r.encode() if isinstance(r, str) else bytes(r) for r, _ in pairs]
  qs = [q.encode() if isinstance(q, str) else bytes(q) for _, q in pairs]
  ref_arr = (C.c_char_p * n)(*refs)
  q_arr = (C.c_char_p * n)(*qs)
  rl = np.array([len(r) for r in refs], dtype=np.int64)
  ql = np.array([len(q) for q in qs], dtype=np.int64)
  out = (_lib.DvbSswAlignment * n)()
  stride = int(16 * (rl + ql).max() + 64)
  buf = C.create_string_buffer(stride * n)
  _lib.check(_lib.lib().dvb_ssw_align_batch(ref_arr, rl.ctypes.data_as(C.c_void_p), q_arr, ql.ctypes.data_as(C.c_void_p), n, match_score,
                                            mismatch_penalty, gap_opening_penalty, gap_extending_penalty, device, out, buf, stride))
  raw = buf.raw
  res = []
  for i in range(n):
    o = out[i]
    cg = raw[i * stride:i * stride + o.cigar_len].decode() if o.cigar_len < stride else ''
    res.append(Alignment(o.sw_score, o.ref_begin, o.ref_end, o.query_begin, o.query_end, o.mismatches, cg))
  return res
