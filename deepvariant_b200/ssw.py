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
