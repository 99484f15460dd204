"""Local realigner of make_examples (--realign_reads, the reference's default for WGS / WES): select windows with evidence of variation,
assemble candidate haplotypes per window with a de Bruijn graph, realign each window's reads to the best haplotype and re-express the
alignment against the reference.

Restates, for one sample with the default flags (deepvariant/realigner/realigner.py:56-243):
  select_windows            deepvariant/realigner/window_selector.py:45-238 + window_selector.cc:40-140 (VARIANT_READS model: a position
                            is a candidate when 2..300 reads carry a non-reference allele over it; ALLELE_COUNT_LINEAR also restated)
  DeBruijnGraph             deepvariant/realigner/debruijn_graph.cc:123-487 (smallest k in 10..101 without a cycle, edge pruning at
                            weight 2, source-to-sink paths, haplotypes sorted)
  Realigner.realign_reads   realigner.py:706-860 (call_debruijn_graph, assign_reads_to_assembled_regions, call_fast_pass_aligner with
                            a 20-bp margin) and make_examples_core.RegionProcessor.realign_reads (:2479-2518, reads longer than 500 bp
                            are left alone)
The read aligner is deepvariant_b200.fast_pass_aligner (FastPassAligner + Smith-Waterman with libssw's tie-breaking).
Pinned by the reference's golden.calling_candidates / golden.calling_examples, which were made WITH the realigner
(tools/check_realigner_golden.py)."""
from __future__ import annotations

import collections
import dataclasses
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from deepvariant_b200 import candidates as cand, fast_pass_aligner
from deepvariant_b200.protos import Read

REF_ALIGN_MARGIN = 20                 # realigner.py:243
MIN_ALLELE_SUPPORT = 2                # realigner.py:269
SUBSTITUTION, INSERTION, DELETION, SOFT_CLIP = 2, 3, 4, 5


@dataclasses.dataclass
class WindowSelectorOptions:
  min_mapq: int = 20
  min_base_quality: int = 20
  min_windows_distance: int = 80
  max_window_size: int = 1000
  region_expansion_in_bp: int = 20
  model_type: str = 'VARIANT_READS'
  min_num_supporting_reads: int = 2
  max_num_supporting_reads: int = 300
  # ALLELE_COUNT_LINEAR defaults (realigner.py:247-258)
  bias: float = -0.683379
  coeff_soft_clip: float = 2.997
  coeff_substitution: float = -0.086644
  coeff_insertion: float = 2.493585
  coeff_deletion: float = 1.795914
  coeff_reference: float = -0.059787
  decision_boundary: float = 3
  keep_legacy_behavior: bool = False
  realign_all: bool = False
  enable_strict_insertion_filter: bool = False


@dataclasses.dataclass
class DeBruijnGraphOptions:
  min_k: int = 10
  max_k: int = 101
  step_k: int = 1
  min_mapq: int = 14
  min_base_quality: int = 15
  min_edge_weight: int = 2
  max_num_paths: int = 256


@dataclasses.dataclass
class RealignerOptions:
  ws: WindowSelectorOptions = dataclasses.field(default_factory=WindowSelectorOptions)
  dbg: DeBruijnGraphOptions = dataclasses.field(default_factory=DeBruijnGraphOptions)
  aln: Dict[str, float] = dataclasses.field(default_factory=lambda: dict(
      match=4, mismatch=6, gap_open=8, gap_extend=2, kmer_size=32, max_num_of_mismatches=2, realignment_similarity_threshold=0.16934))
  max_read_length_to_realign: int = 500
  normalize_reads: bool = False


# ---- window selector -------------------------------------------------------------------------------------------------------------------
def _update(counts: np.ndarray, by, start: int, end: int) -> None:
  start, end = max(start, 0), min(end, len(counts))
  if start < end:
    counts[start:end] += by


def candidate_positions_from_counts(sites: Sequence[dict], start: int, o: WindowSelectorOptions) -> List[int]:
  """`sites` = candidates.debug_allele_counts of the expanded region (AlleleCounter.Counts())."""
  n = len(sites)
  if o.model_type == 'VARIANT_READS':
    counts = np.zeros(n, dtype=np.int64)
    for i, s in enumerate(sites):
      summed: Dict[Tuple[str, int], int] = collections.OrderedDict()
      total = s['ref']
      for bases, typ, low, *_ in s['alleles']:
        if low:
          continue
        summed[(bases, typ)] = summed.get((bases, typ), 0) + 1
        if typ != 1:
          total += 1
      for (bases, typ), count in summed.items():
        if typ == 1 or count < MIN_ALLELE_SUPPORT:
          continue
        if o.enable_strict_insertion_filter and typ == INSERTION and len(bases) <= 2 and np.float32(count) / np.float32(total) < 0.08:
          continue
        if typ == SUBSTITUTION:
          _update(counts, count, i, i + 1)
        elif typ in (SOFT_CLIP, INSERTION):
          _update(counts, count, i + 1 - (len(bases) - 1), i + len(bases))
        elif typ == DELETION:
          _update(counts, count, i + 1, i + len(bases))
    return [start + i for i in range(n) if o.min_num_supporting_reads <= counts[i] <= o.max_num_supporting_reads]
  if o.model_type == 'ALLELE_COUNT_LINEAR':
    scores = np.full(n, np.float32(o.bias), dtype=np.float32)
    coeff = {SUBSTITUTION: o.coeff_substitution, SOFT_CLIP: o.coeff_soft_clip, INSERTION: o.coeff_insertion, DELETION: o.coeff_deletion,
             1: o.coeff_reference}
    for i, s in enumerate(sites):
      _update(scores, np.float32(s['ref'] * np.float32(o.coeff_reference)), i, i + 1)
      for bases, typ, *_ in s['alleles']:
        by = np.float32(np.float32(coeff[typ]))
        if typ in (SUBSTITUTION, 1):
          _update(scores, by, i, i + 1)
        elif typ in (SOFT_CLIP, INSERTION):
          _update(scores, by, i + 1 - (len(bases) - 1), i + len(bases))
        else:
          _update(scores, by, i + 1, i + len(bases))
    return [start + i for i in range(n) if scores[i] > o.decision_boundary]
  raise ValueError(f'Unknown enum option "{o.model_type}" for WindowSelectorModel.model_type')


def candidates_to_windows(positions: Sequence[int], o: WindowSelectorOptions) -> List[Tuple[int, int]]:
  windows: List[Tuple[int, int]] = []
  start = end = None
  for pos in sorted(positions):
    if start is None:
      start = end = pos
    elif pos > end + 2 * o.min_windows_distance:
      windows.append((start - o.min_windows_distance, end + o.min_windows_distance))
      start = end = pos
    else:
      end = pos
  if start is not None:
    windows.append((start - o.min_windows_distance, end + o.min_windows_distance))
  return sorted(windows)


def select_windows(table, ref_reader, contig: str, rows: np.ndarray, region: Tuple[int, int], o: WindowSelectorOptions) -> List[Tuple[int, int]]:
  if not len(rows):
    return []
  if o.realign_all:
    return [region]
  n_bases = ref_reader.n_bases(contig)
  start, end = max(region[0] - o.region_expansion_in_bp, 0), min(region[1] + o.region_expansion_in_bp, n_bases)
  copts = cand.CandidateOptions(min_mapping_quality=o.min_mapq, min_base_quality=o.min_base_quality,
                                keep_legacy_allele_counter_behavior=o.keep_legacy_behavior)
  sites = cand.debug_allele_counts(table, ref_reader, contig, start, end, rows, copts)
  return candidates_to_windows(candidate_positions_from_counts(sites, start, o), o)


_NOT_ACGT = np.ones(256, dtype=bool)
_NOT_ACGT[[65, 67, 71, 84]] = False


# ---- de Bruijn graph ----------------------------------------------------------------------------------------------------------------------
class DeBruijnGraph:

  def __init__(self, ref: str, reads: Sequence[Read], o: DeBruijnGraphOptions, k: int):
    self.o, self.k = o, k
    self.out: Dict[str, Dict[str, List]] = collections.OrderedDict()      # kmer -> {next kmer: [weight, is_ref]}
    self._add(ref, 0, len(ref) - k, True)
    self.source, self.sink = ref[:k], ref[len(ref) - k:]
    for r in reads:
      if r.mapping_quality >= o.min_mapq:
        self._add_read(r)

  def _vertex(self, kmer: str) -> None:
    if kmer not in self.out:
      self.out[kmer] = collections.OrderedDict()

  def _add(self, bases: str, start: int, end: int, is_ref: bool) -> None:
    k = self.k
    if end > 0:
      prev = bases[start:start + k]
      self._vertex(prev)
      for i in range(start + 1, end + 1):
        cur = bases[i:i + k]
        self._vertex(cur)
        e = self.out[prev].setdefault(cur, [0, False])
        e[0] += 1
        e[1] = e[1] or is_ref
        prev = cur

  def _add_read(self, read: Read) -> None:
    bases = read.aligned_sequence.decode().upper()
    n, k = len(bases), self.k
    stop = n - k
    # positions that break a run of usable bases (non-ACGT or below the base-quality floor), found once per read
    seq = np.frombuffer(bases.encode(), dtype=np.uint8)
    quals = np.frombuffer(bytes(read.aligned_quality), dtype=np.uint8)
    breaks = np.flatnonzero((_NOT_ACGT[seq]) | (quals[:n] < self.o.min_base_quality)).tolist() if len(quals) >= n else None
    if breaks is None:
      raise ValueError('read with fewer qualities than bases')
    breaks.append(n)
    i, b = 0, 0
    while i < stop:
      while breaks[b] < i:
        b += 1
      bad = breaks[b]
      self._add(bases, i, bad - k, False)
      i = bad + 1

  def has_cycle(self) -> bool:
    color: Dict[str, int] = {}
    for root in self.out:
      if root in color:
        continue
      stack = [(root, iter(self.out[root]))]
      color[root] = 1
      while stack:
        v, it = stack[-1]
        for w in it:
          c = color.get(w, 0)
          if c == 1:
            return True
          if c == 0:
            color[w] = 1
            stack.append((w, iter(self.out[w])))
            break
        else:
          color[v] = 2
          stack.pop()
    return False

  def _reachable(self, root: str, adj: Dict[str, Sequence[str]]) -> set:
    seen = {root}
    stack = [root]
    while stack:
      v = stack.pop()
      for w in adj.get(v, ()):
        if w not in seen:
          seen.add(w)
          stack.append(w)
    return seen

  def prune(self) -> None:
    for v in self.out:
      self.out[v] = collections.OrderedDict((w, e) for w, e in self.out[v].items() if e[1] or e[0] >= self.o.min_edge_weight)
    rev: Dict[str, List[str]] = {}
    for v, ws in self.out.items():
      for w in ws:
        rev.setdefault(w, []).append(v)
    keep = self._reachable(self.source, self.out) & self._reachable(self.sink, rev)
    self.out = collections.OrderedDict((v, collections.OrderedDict((w, e) for w, e in ws.items() if w in keep))
                                       for v, ws in self.out.items() if v in keep)

  def candidate_haplotypes(self) -> List[str]:
    terminated: List[List[str]] = []
    queue = collections.deque([[self.source]])
    while queue:
      if len(terminated) + len(queue) > self.o.max_num_paths:
        return []
      path = queue.popleft()
      for w in self.out.get(path[-1], ()):
        ext = path + [w]
        if w == self.sink or not self.out.get(w):
          terminated.append(ext)
        else:
          queue.append(ext)
    return sorted(''.join(v[0] for v in p) + p[-1][1:] for p in terminated)


def build_graph(ref: str, reads: Sequence[Read], o: DeBruijnGraphOptions) -> Optional[DeBruijnGraph]:
  """DeBruijnGraph::Build (:224-248): the smallest k for which neither the reference nor the graph has a cycle."""
  max_k = min(o.max_k, len(ref) - 1)
  min_k = -1
  for k in range(o.min_k, max_k + 1, o.step_k):
    if len({ref[i:i + k] for i in range(len(ref) - k + 1)}) == len(ref) - k + 1:
      min_k = k
      break
  if min_k < 0:
    return None
  for k in range(min_k, max_k + 1, o.step_k):
    g = DeBruijnGraph(ref, reads, o, k)
    if g.has_cycle():
      continue
    g.prune()
    return g
  return None


def candidate_haplotypes_native(ref: str, reads: Sequence[Read], o: DeBruijnGraphOptions) -> Optional[List[str]]:
  """build_graph(...).candidate_haplotypes() in native code (csrc/dvb_dbg.cu, dvb_dbg_candidate_haplotypes): None when no k gives
  an acyclic graph, else the sorted haplotypes (possibly none).  The Python DeBruijnGraph above is its cross-check in the tests."""
  import ctypes as C
  from deepvariant_b200 import _lib
  lib = _lib.lib()
  seqs = [bytes(r.aligned_sequence) for r in reads]
  begin = np.zeros(len(reads) + 1, dtype=np.int64)
  if reads:
    np.cumsum([len(x) for x in seqs], out=begin[1:])
  bases = b''.join(seqs)
  quals = np.frombuffer(b''.join(bytes(r.aligned_quality) for r in reads), dtype=np.uint8) if reads else np.zeros(1, np.uint8)
  if len(quals) < int(begin[-1]):
    raise ValueError('read with fewer qualities than bases')
  mapq = np.fromiter((r.mapping_quality for r in reads), dtype=np.int32, count=len(reads)) if reads else np.zeros(1, np.int32)
  ref_b = ref.encode()
  cap = 64 * (len(ref_b) + 64)
  while True:
    out = C.create_string_buffer(cap)
    need = lib.dvb_dbg_candidate_haplotypes(ref_b, len(ref_b), bases, quals.ctypes.data, begin.ctypes.data, mapq.ctypes.data, len(reads), o.min_k, o.max_k,
                                            o.step_k, o.min_mapq, o.min_base_quality, o.min_edge_weight, o.max_num_paths, out, cap, None)
    if need < 0:
      _lib.check(int(-need))
    if need == 0:
      return None
    if need <= cap:
      return out.value.decode().split('\n')[:-1] if need > 1 else []
    cap = int(need)


# ---- the realigner ----------------------------------------------------------------------------------------------------------------------------
def _overlap(a0: int, a1: int, b0: int, b1: int) -> int:
  return max(0, min(a1, b1) - max(a0, b0))


class Realigner:

  def __init__(self, ref_reader, options: Optional[RealignerOptions] = None):
    self.ref_reader = ref_reader
    self.o = options or RealignerOptions()
    self.native_graph = True      # de Bruijn graph in native code (csrc/dvb_dbg.cu); False = the Python restatement (the tests' cross-check)
    self.ssw_device: Optional[int] = None   # CUDA device for the Smith-Waterman alignments of FastPassAligner (batched launches); None = host

  def call_debruijn_graph(self, contig: str, windows: Sequence[Tuple[int, int]], reads: Sequence[Read]) -> List[Tuple[Tuple[int, int], List[str]]]:
    out = []
    for w0, w1 in windows:
      if w1 - w0 > self.o.ws.max_window_size or not self.ref_reader.is_valid_interval(contig, w0, w1):
        continue
      ref = self.ref_reader.query(contig, w0, w1)
      window_reads = [r for r in reads if w1 > r.position and w0 < r.end()]
      if self.native_graph:
        haplotypes = candidate_haplotypes_native(ref, window_reads, self.o.dbg)
        haplotypes = [ref] if haplotypes is None else haplotypes
      else:
        g = build_graph(ref, window_reads, self.o.dbg)
        haplotypes = [ref] if g is None else g.candidate_haplotypes()
      if haplotypes and haplotypes != [ref]:
        out.append(((w0, w1), haplotypes))
    return out

  def call_fast_pass_aligner(self, contig: str, region: Tuple[int, int], haplotypes: Sequence[str], reads: List[Read]) -> List[Read]:
    if not reads:
      return []
    span0, span1 = min(r.position for r in reads), max(r.end() for r in reads)
    ref_start = max(0, min(span0, region[0]) - REF_ALIGN_MARGIN)
    ref_end = min(self.ref_reader.n_bases(contig), max(span1, region[1]) + REF_ALIGN_MARGIN)
    prefix = self.ref_reader.query(contig, ref_start, region[0])
    ref = self.ref_reader.query(contig, region[0], region[1])
    if ref_end <= region[1]:
      return reads
    suffix = self.ref_reader.query(contig, region[1], ref_end)
    a = fast_pass_aligner.FastPassAligner()
    a.ssw_device = self.ssw_device
    a.normalize_reads = self.o.normalize_reads
    c = self.o.aln
    a.set_options(kmer_size=c['kmer_size'], read_size=len(reads[0].aligned_sequence), max_num_of_mismatches=c['max_num_of_mismatches'],
                  realignment_similarity_threshold=c['realignment_similarity_threshold'], match=c['match'], mismatch=c['mismatch'],
                  gap_open=c['gap_open'], gap_extend=c['gap_extend'], force_alignment=False)
    a.reference = prefix + ref + suffix
    a.region_position_in_chr = ref_start
    a.ref_prefix_len, a.ref_suffix_len = len(prefix), len(suffix)
    a.haplotypes = [prefix + h + suffix for h in haplotypes]
    return a.align_reads(reads)

  def realign_reads(self, table, contig: str, rows: np.ndarray, region: Tuple[int, int]) -> List[Read]:
    """RegionProcessor.realign_reads + Realigner.realign_reads over the table rows of a region -> the region's reads, realigned
    where a window's assembly asked for it, in the reference's output order (long reads, unassigned reads, then window by window)."""
    reads = [table.read(int(r)) for r in rows]
    limit = self.o.max_read_length_to_realign
    long_reads = [r for r in reads if limit and len(r.aligned_sequence) > limit]
    keep = [i for i, r in enumerate(reads) if not (limit and len(r.aligned_sequence) > limit)]
    short_reads = [reads[i] for i in keep]
    if not short_reads:
      return long_reads
    windows = select_windows(table, self.ref_reader, contig, np.asarray(rows)[keep], region, self.o.ws)
    assembled = self.call_debruijn_graph(contig, windows, short_reads)
    per_window: List[List[Read]] = [[] for _ in assembled]
    out: List[Read] = []
    for r in short_reads:
      overlaps = [_overlap(r.position, r.end(), w[0][0], w[0][1]) for w in assembled]
      best = max(range(len(assembled)), key=lambda i: overlaps[i]) if assembled else None
      if best is None or overlaps[best] == 0:
        out.append(r)
      else:
        per_window[best].append(r)
    for (window, haplotypes), window_reads in zip(assembled, per_window):
      out.extend(self.call_fast_pass_aligner(contig, window, haplotypes, window_reads))
    return long_reads + out
