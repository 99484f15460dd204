"""FastPassAligner: aligns reads to candidate haplotypes (exact k-mer pass, then Smith-Waterman) and re-expresses the best alignment
against the reference - the aligner behind the alt-aligned pileups (make_examples_native.cc:553-626, alt_aligned_pileup_lib.cc:278-313)
and the local realigner (SURVEY.md 8(f) "next" row #3).

Restates deepvariant/realigner/fast_pass_aligner.cc: AlignReads :123-143, CalculateSswAlignmentScoreThreshold :111-121,
BuildIndex / AddReadToIndex :566-585, FastAlignReadsToHaplotype(s) :165-254, FastAlignStrings :256-279, AlignHaplotypesToReference
:322-357, SetPositionsMap :586-637, SswAlignReadsToHaplotypes :359-397, GetBestReadAlignment :645-667, RealignReadsToReference
:447-524, IsAlignmentNormalized :399-445, CalculateReadToRefAlignment / MergeCigarOp / LeftTrimHaplotypeToRefAlignment /
MergeOneBaseOperations :669-886.  Smith-Waterman = deepvariant_b200.ssw (libssw's tie-breaking, csrc/dvb_ssw.cu).
CIGAR operations are BAM codes here: 0 M, 1 I, 2 D, 4 S.  Pinned by the known answers of fast_pass_aligner_test.cc
(tests/test_fast_pass_aligner.py)."""
from __future__ import annotations

import copy
import dataclasses
import re
from typing import Dict, List, Optional, Sequence, Tuple

from deepvariant_b200 import ssw
from deepvariant_b200.protos import Read

M, I, D, S = 0, 1, 2, 4
K_NOT_ALIGNED = 65535            # ReadAlignment::kNotAligned (uint16 max)
_CIGAR_RE = re.compile(r'(\d+)([XIDS=])')
_OP_OF = {'=': M, 'X': M, 'S': S, 'D': D, 'I': I}


@dataclasses.dataclass
class ReadAlignment:
  position: int = K_NOT_ALIGNED
  cigar: str = ''
  score: int = 0

  def reset(self) -> None:
    self.position, self.cigar, self.score = K_NOT_ALIGNED, '', 0


@dataclasses.dataclass
class HaplotypeReadsAlignment:
  haplotype_index: int
  haplotype_score: int
  read_alignment_scores: List[ReadAlignment]
  cigar: str = ''
  cigar_ops: List[Tuple[int, int]] = dataclasses.field(default_factory=list)     # (op, length)
  ref_pos: int = 0
  hap_to_ref_positions_map: List[int] = dataclasses.field(default_factory=list)
  is_reference: bool = False


def cigar_string_to_ops(cigar: str) -> List[Tuple[int, int]]:
  return [(_OP_OF[op], int(n)) for n, op in _CIGAR_RE.findall(cigar)]


def set_positions_map(haplotype_size: int, cigar: str) -> List[int]:
  """SetPositionsMap: shift from a haplotype position to the reference position it aligns to."""
  out = [0] * haplotype_size
  shift = pos = 0
  for n, op in _CIGAR_RE.findall(cigar):
    n = int(n)
    if op in '=X':
      for _ in range(n):
        out[pos] = shift
        pos += 1
    elif op == 'S':
      shift -= n
      for _ in range(n):
        out[pos] = shift
        pos += 1
    elif op == 'D':
      shift += n
    elif op == 'I':
      for _ in range(n):
        out[pos] = shift
        shift -= 1
        pos += 1
  return out


def _aligned_length(cigar: Sequence[List[int]]) -> int:
  return sum(n for op, n in cigar if op != D)


def merge_cigar_op(op: int, length: int, read_len: int, cigar: List[List[int]]) -> None:
  """MergeCigarOp (:680-737)."""
  last_op = cigar[-1][0] if cigar else None
  before = _aligned_length(cigar)
  new_len = min(length, read_len - before) if op != D else length
  if new_len <= 0 or before == read_len:
    return
  if (op == I and last_op == D) or (op == D and last_op == I):
    one_before_last = cigar[-2] if len(cigar) > 1 else cigar[-1]
    if one_before_last[0] != M:
      cigar.insert(len(cigar) - 1, [M, 1])
    else:
      one_before_last[1] += 1
    if cigar[-1][1] == 1:
      cigar.pop()
    else:
      cigar[-1][1] -= 1
  elif op == last_op:
    cigar[-1][1] += new_len
  else:
    cigar.append([op, new_len])


def _left_trim(hap_to_ref: Sequence[Tuple[int, int]], read_to_hap_pos: int) -> List[List[int]]:
  ops = [list(o) for o in hap_to_ref]
  cur = 0
  while cur != read_to_hap_pos:
    if not ops:
      raise ValueError('haplotype-to-reference alignment is shorter than the read position')
    op, n = ops.pop(0)
    if op in (M, 5, S, I):
      if n + cur > read_to_hap_pos:
        ops.insert(0, [op, n - (read_to_hap_pos - cur)])
      cur = min(n + cur, read_to_hap_pos)
  if ops and ops[0][0] == D:
    ops.pop(0)
  return ops


def _merge_one_base(read_op: int, hap_op: int, read_len: int, out: List[List[int]]) -> None:
  assert (read_op, hap_op) not in ((D, I), (I, D))
  for op in (S, D, I, M):
    if read_op == op or hap_op == op:
      merge_cigar_op(op, 1, read_len, out)
      break


def _merge_bases(read_op: int, hap_op: int, count: int, read_len: int, out: List[List[int]]) -> None:
  """`count` calls of _merge_one_base(read_op, hap_op) (the reference merges base by base, fast_pass_aligner.cc:760-800): only a
  call that meets the I-after-D / D-after-I rewrite of MergeCigarOp depends on being a single base; once the last operation is
  anything else, the remaining calls just extend it (clamped to the read length, deletions unclamped) - one call does that."""
  assert (read_op, hap_op) not in ((D, I), (I, D))
  op = next(o for o in (S, D, I, M) if read_op == o or hap_op == o)
  while count > 0:
    last_op = out[-1][0] if out else None
    if (op == I and last_op == D) or (op == D and last_op == I):
      merge_cigar_op(op, 1, read_len, out)
      count -= 1
      continue
    merge_cigar_op(op, count, read_len, out)
    return


class FastPassAligner:
  ssw_batch_min = 4      # fewer (haplotype, read) pairs than this are aligned on the host even when ssw_device is set

  def __init__(self):
    self.kmer_size = 32
    self.read_size = 100
    self.max_num_of_mismatches = 2
    self.similarity_threshold = 0.85
    self.match_score, self.mismatch_penalty, self.gap_opening_penalty, self.gap_extending_penalty = 4, 6, 8, 1
    self.force_alignment = False
    self.normalize_reads = False
    self.reference = ''
    self.reads: List[str] = []
    self.haplotypes: List[str] = []
    self.region_position_in_chr = 0
    self.ref_prefix_len = self.ref_suffix_len = 0
    self.kmer_index: Dict[str, List[Tuple[int, int]]] = {}
    self.read_to_haplotype_alignments: List[HaplotypeReadsAlignment] = []
    self.ssw_alignment_score_threshold = 0
    self._ssw: Optional[ssw.Aligner] = None
    self.ssw_device: Optional[int] = None      # CUDA device for batched Smith-Waterman (None: host dvb_ssw_align per pair)

  # -- configuration (set_options :81-109: zero / unset fields keep the defaults) ----------------------------------------------
  def set_options(self, kmer_size=0, read_size=0, max_num_of_mismatches=0, realignment_similarity_threshold=0.0, match=0, mismatch=0,
                  gap_open=0, gap_extend=0, force_alignment=False) -> None:
    if kmer_size > 0:
      self.kmer_size = kmer_size
    if read_size > 0:
      self.read_size = read_size
    if max_num_of_mismatches > 0:
      self.max_num_of_mismatches = max_num_of_mismatches
    if realignment_similarity_threshold > 0.0:
      self.similarity_threshold = realignment_similarity_threshold
    if match > 0:
      self.match_score = match
    if mismatch > 0:
      self.mismatch_penalty = mismatch
    if gap_open > 0:
      self.gap_opening_penalty = gap_open
    if gap_extend > 0:
      self.gap_extending_penalty = gap_extend
    self.force_alignment = force_alignment
    if not 3 <= self.kmer_size <= 32:
      raise ValueError('kmer_size must be in [3, 32]')

  def calculate_ssw_alignment_score_threshold(self) -> None:
    t = (self.match_score * self.read_size * self.similarity_threshold -
         self.mismatch_penalty * self.read_size * (1 - self.similarity_threshold))
    t = int(t)                          # double -> uint16_t member
    self.ssw_alignment_score_threshold = 1 if t < 0 else t

  # -- k-mer index ---------------------------------------------------------------------------------------------------------------
  def build_index(self) -> None:
    self.kmer_index = {}
    k = self.kmer_size
    for read_id, read in enumerate(self.reads):
      if len(read) <= k:
        continue
      for i in range(len(read) - k + 1):
        self.kmer_index.setdefault(read[i:i + k], []).append((read_id, i))

  # -- fast pass -----------------------------------------------------------------------------------------------------------------
  def fast_align_strings(self, s1: str, s2: str, max_mismatches: int) -> Tuple[int, int]:
    """-> (score, number of mismatches); score 0 as soon as max_mismatches is reached."""
    matches = mismatches = 0
    for c1, c2 in zip(s1, s2):
      if c1 != c2 and c1 != 'N' and c2 != 'N':
        mismatches += 1
        if mismatches == max_mismatches:
          return 0, mismatches
      else:
        matches += 1
    return matches * self.match_score - mismatches * self.mismatch_penalty, mismatches

  def fast_align_reads_to_haplotype(self, haplotype: str, scores: List[ReadAlignment]) -> int:
    haplotype_score = 0
    is_ref = haplotype == self.reference
    coverage = [0] * len(haplotype)
    k = self.kmer_size
    for i in range(len(haplotype) - k + 1):
      hits = self.kmer_index.get(haplotype[i:i + k])
      if hits is None:
        continue                      # ... which also skips the zero-coverage test below (:181-184)
      for read_id, read_pos in hits:
        target_start = max(0, i - read_pos)
        read = self.reads[read_id]
        if target_start + len(read) > len(haplotype):
          continue
        ra = scores[read_id]
        if ra.position != K_NOT_ALIGNED and ra.position == target_start:
          continue
        new_score, n_mismatches = self.fast_align_strings(haplotype[target_start:target_start + len(read)], read, self.max_num_of_mismatches + 1)
        if n_mismatches <= self.max_num_of_mismatches:
          old = ra.score
          for p in range(target_start, target_start + len(read)):
            coverage[p] += 1
          if old < new_score:
            ra.score = new_score
            haplotype_score += new_score - old
            ra.position = target_start
            ra.cigar = f'{len(read)}='
      if coverage[i] == 0 and self.ref_prefix_len <= i < len(haplotype) - self.ref_suffix_len and not is_ref:
        return 0
    return haplotype_score

  def fast_align_reads_to_haplotypes(self) -> None:
    """All haplotypes in one native call (dvb_fast_pass_scores); fast_align_reads_to_haplotype below is the same pass in Python, kept for
    the known-answer tests and as the cross-check of tests/test_fast_pass_aligner.py."""
    import ctypes as C
    import numpy as np
    from deepvariant_b200 import _lib
    n_h, n_r = len(self.haplotypes), len(self.reads)
    if n_h and n_r:
      haps = [h.encode() for h in self.haplotypes]
      reads = [r.encode() for r in self.reads]
      hap_arr, read_arr = (C.c_char_p * n_h)(*haps), (C.c_char_p * n_r)(*reads)
      hap_lens = np.array([len(h) for h in haps], dtype=np.int64)
      read_lens = np.array([len(r) for r in reads], dtype=np.int64)
      hap_score = np.zeros(n_h, dtype=np.int32)
      position = np.zeros(n_h * n_r, dtype=np.int32)
      score = np.zeros(n_h * n_r, dtype=np.int32)
      ref = self.reference.encode()
      _lib.check(_lib.lib().dvb_fast_pass_scores(ref, len(ref), hap_arr, hap_lens.ctypes.data_as(C.c_void_p), n_h, read_arr,
                                                 read_lens.ctypes.data_as(C.c_void_p), n_r, self.kmer_size, self.max_num_of_mismatches,
                                                 self.match_score, self.mismatch_penalty, self.ref_prefix_len, self.ref_suffix_len,
                                                 hap_score.ctypes.data_as(C.c_void_p), position.ctypes.data_as(C.c_void_p),
                                                 score.ctypes.data_as(C.c_void_p)))
      position, score = position.reshape(n_h, n_r), score.reshape(n_h, n_r)
      for i in range(n_h):
        scores = [ReadAlignment(int(position[i, r]), f'{len(self.reads[r])}=' if score[i, r] > 0 or position[i, r] != K_NOT_ALIGNED else '',
                                int(score[i, r])) for r in range(n_r)]
        self.read_to_haplotype_alignments.append(HaplotypeReadsAlignment(i, int(hap_score[i]), scores))
      return
    self.fast_align_reads_to_haplotypes_py()

  def fast_align_reads_to_haplotypes_py(self) -> None:
    if not self.kmer_index:
      self.build_index()
    for i, haplotype in enumerate(self.haplotypes):
      scores = [ReadAlignment() for _ in self.reads]
      hap_score = self.fast_align_reads_to_haplotype(haplotype, scores)
      if hap_score == 0:
        for s in scores:
          s.reset()
      self.read_to_haplotype_alignments.append(HaplotypeReadsAlignment(i, hap_score, scores))

  # -- Smith-Waterman ------------------------------------------------------------------------------------------------------------
  def _ssw_align(self, reference: str, query: str) -> ssw.Alignment:
    if self._ssw is None:
      self._ssw = ssw.Aligner(self.match_score, self.mismatch_penalty, self.gap_opening_penalty, self.gap_extending_penalty)
    self._ssw.set_reference_sequence(reference)
    return self._ssw.align(query)

  def align_haplotypes_to_reference(self) -> None:
    if not self.read_to_haplotype_alignments:
      self.read_to_haplotype_alignments = [HaplotypeReadsAlignment(i, -1, [ReadAlignment() for _ in self.reads])
                                           for i in range(len(self.haplotypes))]
    # with `ssw_device` set the haplotype-to-reference alignments of the call share one dvb_ssw_align_batch launch
    todo = [ha for ha in self.read_to_haplotype_alignments if self.haplotypes[ha.haplotype_index] != self.reference]
    batched = {}
    if self.ssw_device is not None and len(todo) >= 2:
      als = ssw.align_batch([(self.reference, self.haplotypes[ha.haplotype_index]) for ha in todo], self.match_score, self.mismatch_penalty,
                            self.gap_opening_penalty, self.gap_extending_penalty, device=self.ssw_device)
      batched = {id(ha): al for ha, al in zip(todo, als)}
    for ha in self.read_to_haplotype_alignments:
      hap = self.haplotypes[ha.haplotype_index]
      if hap == self.reference:
        ha.is_reference, ha.cigar, ha.ref_pos = True, f'{len(hap)}=', 0
        ha.cigar_ops = cigar_string_to_ops(ha.cigar)
      else:
        al = batched[id(ha)] if batched else self._ssw_align(self.reference, hap)
        if al.sw_score > 0:
          ha.is_reference = al.cigar_string == f'{len(hap)}='
          ha.cigar, ha.ref_pos = al.cigar_string, al.ref_begin
          ha.cigar_ops = cigar_string_to_ops(ha.cigar)

  def calculate_position_maps(self) -> None:
    for ha in self.read_to_haplotype_alignments:
      ha.hap_to_ref_positions_map = set_positions_map(len(self.haplotypes[ha.haplotype_index]), ha.cigar)

  def ssw_align_reads_to_haplotypes(self, score_threshold: int) -> None:
    """SswAlignReadsToHaplotypes (fast_pass_aligner.cc:281-322).  With `ssw_device` set, all (haplotype, read) pairs of the call go
    through ONE dvb_ssw_align_batch launch (the Smith-Waterman scans on the GPU, csrc/dvb_ssw_gpu.cu); same results either way."""
    jobs = []
    for i, read in enumerate(self.reads):
      if any(ha.read_alignment_scores[i].score > 0 for ha in self.read_to_haplotype_alignments):
        continue
      for ha in self.read_to_haplotype_alignments:
        forced = self.force_alignment and ha.is_reference
        if ha.haplotype_score == 0 and not forced:
          continue
        jobs.append((i, ha, forced))
    if not jobs:
      return
    if self.ssw_device is not None and len(jobs) >= self.ssw_batch_min:
      als = ssw.align_batch([(self.haplotypes[ha.haplotype_index], self.reads[i]) for i, ha, _ in jobs], self.match_score, self.mismatch_penalty,
                            self.gap_opening_penalty, self.gap_extending_penalty, device=self.ssw_device)
    else:
      als = [self._ssw_align(self.haplotypes[ha.haplotype_index], self.reads[i]) for i, ha, _ in jobs]
    for (i, ha, forced), al in zip(jobs, als):
      if al.sw_score > 0 and (al.sw_score >= score_threshold or forced):
        ra = ha.read_alignment_scores[i]
        ra.score, ra.cigar, ra.position = al.sw_score, al.cigar_string, al.ref_begin

  def get_best_read_alignment(self, read_id: int) -> Optional[int]:
    best_score, best = 0, None
    for hap_index in range(len(self.haplotypes)):
      ha = self.read_to_haplotype_alignments[hap_index]
      s = ha.read_alignment_scores[read_id].score
      if s > best_score or (best_score > 0 and s == best_score and not ha.is_reference):
        best_score, best = s, hap_index
    return best

  # -- read -> reference ------------------------------------------------------------------------------------------------------------
  def calculate_read_to_ref_alignment(self, read_index: int, read_to_hap: ReadAlignment, hap_to_ref_ops: Sequence[Tuple[int, int]]) -> List[List[int]]:
    read_len = len(self.reads[read_index])
    out: List[List[int]] = []
    r2h = [list(o) for o in cigar_string_to_ops(read_to_hap.cigar)]
    h2r = _left_trim(hap_to_ref_ops, read_to_hap.position)
    if not h2r:
      raise ValueError('empty haplotype-to-reference alignment')
    if r2h and r2h[0][0] == S:
      merge_cigar_op(S, r2h[0][1], read_len, out)
      r2h.pop(0)
    cur_r = [None, 0]
    cur_h = [None, 0]
    while (r2h or h2r) and _aligned_length(out) < read_len:
      if r2h and not h2r and cur_h[1] == 0:
        merge_cigar_op(r2h[0][0], r2h[0][1], read_len, out)
        r2h.pop(0)
        continue
      if not r2h and cur_r[1] == 0 and h2r:
        break
      if cur_r[1] == 0:
        cur_r = r2h.pop(0)
      if cur_h[1] == 0:
        cur_h = h2r.pop(0)
      while cur_r[1] > 0 and cur_h[1] > 0:
        if (cur_r[0], cur_h[0]) in ((D, I), (I, D)):
          cur_h[1] -= 1
          cur_r[1] -= 1
          if cur_h[0] == D:
            h2r.insert(0, [M, 1])
            r2h.insert(0, [M, 1])
          continue
        # the pair of operations stays the same until one of them runs out: merge the whole run
        if cur_r[0] == I:
          _merge_bases(cur_r[0], cur_h[0], cur_r[1], read_len, out)
          cur_r[1] = 0
        elif cur_h[0] == D:
          _merge_bases(cur_r[0], cur_h[0], cur_h[1], read_len, out)
          cur_h[1] = 0
        else:
          k = min(cur_r[1], cur_h[1])
          _merge_bases(cur_r[0], cur_h[0], k, read_len, out)
          cur_h[1] -= k
          cur_r[1] -= k
    if cur_r[1] > 0 and cur_r[0] == S:
      _merge_bases(cur_r[0], cur_h[0], cur_r[1], read_len, out)
      cur_r[1] = 0
    if r2h or cur_r[1] > 0:
      out = []
    return out

  def is_alignment_normalized(self, cigar: Sequence[Sequence[int]], ref_offset: int, read_sequence: str) -> bool:
    if ref_offset < 0:
      return True
    ref_off, read_off = ref_offset, 0
    for op, n in cigar:
      if op == S:
        read_off += n
        continue
      if op != M:
        if op == D:
          if ref_off + n > len(self.reference):
            return False
          seq = self.reference[ref_off:ref_off + n]
        else:
          seq = read_sequence[read_off:read_off + n]
        if (ref_off > 0 and op == I and seq[-1] == self.reference[ref_off - 1]) or \
           (read_off > 0 and op == D and seq[-1] == read_sequence[read_off - 1]):
          return False
      if op != I:
        ref_off += n
      if op != D:
        read_off += n
    return True

  def realign_reads_to_reference(self, reads: Sequence[Read]) -> List[Read]:
    out: List[Read] = []
    for read_index, read in enumerate(reads):
      best = self.get_best_read_alignment(read_index)
      if best is None:
        out.append(Read() if self.force_alignment else copy.copy(read))
        continue
      ha = self.read_to_haplotype_alignments[best]
      ra = ha.read_alignment_scores[read_index]
      hap_to_ref = ha.hap_to_ref_positions_map[ra.position]
      offset = ha.ref_pos + ra.position + hap_to_ref
      cigar = self.calculate_read_to_ref_alignment(read_index, ra, ha.cigar_ops)
      if not self.normalize_reads and not self.is_alignment_normalized(cigar, offset, self.reads[read_index]):
        cigar = []
      realigned = copy.copy(read)
      if cigar:
        realigned.position = self.region_position_in_chr + offset
        realigned.cigar = [(op, n) for op, n in cigar]
      out.append(realigned)
    return out

  # -- the entry point ----------------------------------------------------------------------------------------------------------------
  def align_reads(self, reads: Sequence[Read]) -> List[Read]:
    self.reads += [r.aligned_sequence.decode().upper() for r in reads]
    self.calculate_ssw_alignment_score_threshold()
    self.fast_align_reads_to_haplotypes()      # the native pass indexes the reads' k-mers itself; build_index() serves the Python cross-check
    self.align_haplotypes_to_reference()
    self.calculate_position_maps()
    self.ssw_align_reads_to_haplotypes(self.ssw_alignment_score_threshold)
    self.read_to_haplotype_alignments.sort(key=lambda h: h.haplotype_score)      # std::sort with operator< on the score
    return self.realign_reads_to_reference(reads)
