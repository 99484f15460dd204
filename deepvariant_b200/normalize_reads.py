"""`make_examples --normalize_reads`: left-normalisation of the indels of every read of a region before allele counting; the
normalised alignment then also feeds the pileups, as the reference rewrites read.alignment in place
(make_examples_core.py:2900-2953).

Restates AlleleCounter::NormalizeCigar (deepvariant/allelecounter.cc:777-845) with its helpers MergeOperations :558-586,
AdvanceReadReferencePointers :590-619, HandleHeadingIndel :624-640, ShiftOperation :647-684, FindAndMergeOperations /
SwipeAndMerge :689-730, CanDelBeShifted / CanInsBeShifted :732-769, and NormalizeAndAdd :847-871 on (bam_op, length) lists.
AlleleCounter::Add with a substitute cigar and a read shift (:873-978) is the plain Add of a read that carries that cigar at the
shifted position - base elements outside the counting interval are dropped either way and RefBases validity is the contig's, not
the interval's (:375-384) - so normalisation is a transformation of the reads and the allele counter itself is unchanged.

The reads interval (the reference bases NormalizeCigar may look at) is [min(region start, first read start), max(region end,
last read end capped at n_bases - 1)) (make_examples_core.py:2906-2921, allelecounter.cc:337-356).

Pinned by allelecounter_test.cc's NormalizeCigar* cases (:1224-1577) in tests/test_normalize_reads.py and by the reference's
golden.calling_examples.with_flags (tools/check_realigner_golden.py --with_flags)."""
from __future__ import annotations

import copy
from typing import Callable, List, Optional, Sequence, Tuple

M, I, D, N, S, H, P, EQ, X = range(9)
_MATCH = (M, EQ, X)
MAX_ITERATIONS = 100000000       # allelecounter.cc:793

Cigar = List[List[int]]


def _merge_operations(a: List[int], b: List[int]) -> bool:
  if a[0] == b[0] or (a[0] in _MATCH and b[0] in _MATCH):
    a[1] += b[1]
    b[1] = 0
    return True
  if a[0] in (I, D) and b[0] in (I, D):
    lo = min(a[1], b[1])
    rest = max(a[1], b[1]) - lo
    if a[1] > b[1]:
      b[0] = a[0]
    a[0], a[1] = M, lo
    b[1] = rest
    return True
  return False


def _swipe_and_merge(cigar: Cigar) -> bool:
  modified = False
  merged = True
  while merged:
    kept = [u for u in cigar if u[1] != 0]
    if len(kept) < len(cigar):
      modified = True
    cigar[:] = kept
    merged = False
    for a, b in zip(cigar, cigar[1:]):
      if _merge_operations(a, b):
        merged = modified = True
        break
  return modified


def _handle_heading_indel(i: int, cigar: Cigar) -> int:
  assert i == 0 or (i == 1 and cigar and cigar[0][0] == S)
  if cigar[i][0] == D:
    shift = cigar[i][1]
    del cigar[i]
    return shift
  if cigar[i][0] == I:
    cigar[i][0] = M
    return -cigar[i][1]
  return 0


def _shift_operation(shift: int, i: int, cigar: Cigar) -> int:
  if i == 0 or (i == 1 and cigar[0][0] == S):
    return _handle_heading_indel(i, cigar)
  prev = cigar[i - 1]
  if prev[0] == S:
    raise ValueError('soft clip in the middle of a cigar')
  if prev[0] not in _MATCH:
    return 0
  if shift > prev[1]:
    raise ValueError('indel shifted past the preceding match')
  prev[1] -= shift
  if i + 1 == len(cigar):
    cigar.append([M, shift])
  else:
    cigar[i + 1][1] += shift
  return 0


def normalize_cigar(read_seq: bytes, interval_offset: int, cigar: Sequence[Tuple[int, int]], ref_bases: bytes,
                    max_iterations: int = MAX_ITERATIONS) -> Tuple[bool, List[Tuple[int, int]], int]:
  """-> (is_modified, normalised cigar, read_shift).  `ref_bases` are the bases of the reads interval, `interval_offset` the read's
  alignment start relative to it.  The cigar is returned with the heading indel handled whether or not is_modified (that is what
  the counter adds); callers rewrite the read only when is_modified, as make_examples_core.py:2942-2950 does."""
  out: Cigar = [[int(op), int(ln)] for op, ln in cigar]
  if not out:
    return False, [], 0
  modified = False
  read_shift = 0
  n_ref, n_seq = len(ref_bases), len(read_seq)
  iteration = 0
  while iteration < max_iterations:
    iteration += 1
    read_offset = 0
    cur = interval_offset + read_shift
    prev_len = out[0][1]
    shifted = False
    for i, (op, op_len) in enumerate(out):
      if op == I or op == D:
        shift = 0
        while prev_len > 0:
          if op == D:
            ok = read_offset > 0 and cur + op_len - 1 < n_ref and read_seq[read_offset - 1] == ref_bases[cur + op_len - 1]
          else:
            ok = cur > 0 and read_offset + op_len - 1 < n_seq and read_seq[read_offset + op_len - 1] == ref_bases[cur - 1]
          if not ok:
            break
          cur -= 1
          prev_len -= 1
          read_offset -= 1
          shift += 1
        if shift > 0:
          read_shift += _shift_operation(shift, i, out)
          modified = shifted = True
          break
      prev_len = op_len
      if op in _MATCH:
        read_offset += op_len
        cur += op_len
      elif op in (S, I):
        read_offset += op_len
      elif op in (D, P, N):
        cur += op_len
    merged = _swipe_and_merge(out)
    modified = modified or merged
    if not shifted and not merged:
      break
  read_shift += _handle_heading_indel(0, out)
  return modified, [(op, ln) for op, ln in out], read_shift


def reads_interval(reads, region_start: int, region_end: int, n_bases: int) -> Tuple[int, int]:
  """make_examples_core.py:2906-2921 + allelecounter.cc:343-346."""
  lo, hi = region_start, region_end
  for r in reads:
    lo = min(lo, r.position)
    hi = max(hi, min(n_bases - 1, r.end()))
  return lo, hi


def normalize_region_reads(reads, fetch: Callable[[int, int], bytes], region_start: int, region_end: int, n_bases: int,
                           min_mapping_quality: int):
  """-> (pileup_reads, count_reads): `reads` are all reads of the region (after realignment); those that overlap
  [region_start, region_end) and pass the mapping-quality test are normalised (NormalizeAndAdd :847-871).  pileup_reads carry the
  rewritten alignment where NormalizeCigar reported a modification; count_reads is what the allele counter adds - it differs from
  pileup_reads (else it is None) only for a read whose sole change is the heading indel, which the reference counts with the
  rewritten cigar but leaves untouched in memory.  Reads are copied, not modified."""
  lo, hi = reads_interval(reads, region_start, region_end, n_bases)
  ref_bases = fetch(lo, hi) if hi > lo else b''
  pileup, count = [], []
  differs = False
  for r in reads:
    overlaps = r.position < region_end and r.end() > region_start
    if not overlaps or r.mapping_quality < min_mapping_quality:
      pileup.append(r)
      count.append(r)
      continue
    modified, cigar, shift = normalize_cigar(r.aligned_sequence, r.position - lo, r.cigar, ref_bases)
    if modified:
      nr = copy.copy(r)
      nr.cigar = cigar
      nr.position = r.position + shift
      pileup.append(nr)
      count.append(nr)
    elif shift != 0 or list(map(tuple, r.cigar)) != cigar:
      nr = copy.copy(r)
      nr.cigar = cigar
      nr.position = r.position + shift
      pileup.append(r)
      count.append(nr)
      differs = True
    else:
      pileup.append(r)
      count.append(r)
  return pileup, (count if differs else None)
