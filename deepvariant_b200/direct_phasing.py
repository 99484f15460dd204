"""Read phasing for `make_examples --phase_reads` (PACBIO / ONT): assigns every read of a region to haplotype 1, 2 or 0 from
the candidates' per-allele read support, so that the haplotype channel and --sort_by_haplotypes see the HP values the
reference computes itself (the HP tag of the input file is discarded: make_examples_core.py:2998-3004, 3100-3111).

Restates deepvariant/direct_phasing.cc (DirectPhasing::PhaseReads :78-185, Build :833-889, CandidateFilter :788-816, AddCandidate
:736-786, UpdateReadToAllelesMap :650-666, UpdateStartingScore :561-596, CalculateScore / FindSupportingReads :480-559, MaxScore
:246-302, AssignPhasesToVertices :304-398, AssignPhasesToReads :429-463) on integer vertex ids instead of a boost graph: a dynamic
programme over the phasable positions whose states are ordered pairs of allele vertices (phase-1 allele, phase-2 allele), scored by
the number of reads that keep supporting the same phase.  The reference's containers that iterate in pointer / hash order
(btree_set<Edge>, flat_hash_map) only feed order-insensitive steps: the edges of a position are re-keyed by (source bases, target
bases), which are unique because the alleles of one position are distinct.  Not restated: methylation-aware phasing, GraphViz output.
Pinned by the haplotype channel and the row order of the reference's golden.pacbio_examples (tools/check_pacbio_golden.py)."""
from __future__ import annotations

import dataclasses
from typing import Dict, List, Optional, Sequence, Set, Tuple

K_MIN_REF_ALLELE_DEPTH = 3        # direct_phasing.cc:68
K_REF = 'REF'                     # :69
UNCALLED = 'UNCALLED_ALLELE'
PHASE_MAX_CANDIDATES = 5000       # make_examples_options.py:684-692
MIN_ALLELES_TO_PHASE = 1          # :676-683


@dataclasses.dataclass
class _Support:
  read_index: int
  is_first_allele: bool = False


@dataclasses.dataclass
class _Vertex:
  position: int
  bases: str
  read_support: List[_Support]
  phase: int = 0


@dataclasses.dataclass
class _Score:
  score: int
  src: Tuple[Optional[int], Optional[int]]
  read_support: Tuple[Set[int], Set[int]]


def _candidate_filter(c: dict, indel_end: List[int]) -> bool:
  called = [a for a in c['allele_support_ext'] if a != UNCALLED]
  if len(called) <= 1 and len(c['ref_support_ext']) < K_MIN_REF_ALLELE_DEPTH:
    return False
  for allele in called:
    if c['end'] <= indel_end[0] or len(allele) != c['end'] - c['start']:
      if indel_end[0] < c['end']:
        indel_end[0] = c['end']
      return False
  return True


class DirectPhasing:
  """phase(candidates, read_keys) -> [phase per read].  `candidates` are candidates.canonical_call dicts in coordinate order (they
  carry allele_support_ext / ref_support_ext, i.e. made with --track_ref_reads), `read_keys` the "fragment_name/read_number" keys of
  the region's reads in their order."""

  def __init__(self, min_alleles_to_phase: int = MIN_ALLELES_TO_PHASE):
    self.min_alleles_to_phase = min_alleles_to_phase

  # -- graph ------------------------------------------------------------------------------------------------------------------
  def _build(self, candidates: Sequence[dict], read_keys: Sequence[str]) -> None:
    self.vertices: List[_Vertex] = []
    self.positions: List[int] = []
    self.by_position: Dict[int, List[int]] = {}
    self.read_to_alleles: Dict[int, List[int]] = {}
    self.in_edges: Dict[int, Set[int]] = {}
    self.read_to_index = {k: i for i, k in enumerate(read_keys)}       # later duplicates win, as read_to_index_[key] = index does
    indel_end = [0]
    prev_start = None
    for c in candidates:
      if prev_start is not None and not prev_start < c['start']:
        raise ValueError('candidates must be sorted by position, one per position')
      prev_start = c['start']
      if _candidate_filter(c, indel_end):
        self._add_candidate(c)
        self.positions.append(c['start'])
    index_of = {p: i for i, p in enumerate(self.positions)}
    for alleles in self.read_to_alleles.values():
      for prev, cur in zip(alleles, alleles[1:]):
        i = index_of[self.vertices[cur].position]
        if i >= 1 and (i == 1 or self.positions[i - 1] == self.vertices[prev].position):
          self.in_edges.setdefault(cur, set()).add(prev)

  def _add_vertex(self, position: int, bases: str, read_infos: Sequence[dict]) -> None:
    support = [_Support(self.read_to_index[r['read_name']]) for r in read_infos
               if r['read_name'] in self.read_to_index and not r['is_low_quality']]
    v = len(self.vertices)
    self.vertices.append(_Vertex(position, bases, support))
    self.by_position.setdefault(position, []).append(v)
    for s in support:
      s.is_first_allele = s.read_index not in self.read_to_alleles
      self.read_to_alleles.setdefault(s.read_index, []).append(v)

  def _add_candidate(self, c: dict) -> None:
    if len(c['alts']) == 1 and c['alts'][0] == '.':
      return                                                   # reference site: methylation-aware phasing only
    if len(c['ref_support_ext']) >= K_MIN_REF_ALLELE_DEPTH:
      self._add_vertex(c['start'], K_REF, c['ref_support_ext'])
    for allele in sorted(a for a in c['allele_support_ext'] if a != UNCALLED):
      self._add_vertex(c['start'], allele, c['allele_support_ext'][allele])

  # -- scores -----------------------------------------------------------------------------------------------------------------
  def _update_starting_score(self, verts: Sequence[int]) -> None:
    for v1 in verts:
      for v2 in verts:
        self.scores.pop((v1, v2), None)
    for i, v1 in enumerate(verts):
      s1 = {s.read_index for s in self.vertices[v1].read_support}
      for v2 in verts[i:]:
        s2 = {s.read_index for s in self.vertices[v2].read_support}
        self.scores[(v1, v2)] = _Score(len(s1) if s1 == s2 else len(s1) + len(s2), (None, None), (s1, s2))

  def _calculate_score(self, e1: Tuple[int, int], e2: Tuple[int, int]) -> _Score:
    prev = self.scores.get((e1[0], e2[0]))
    if prev is None:
      return _Score(0, (None, None), (set(), set()))
    continuing, first = [], []
    for phase, to in enumerate((e1[1], e2[1])):
      support = self.vertices[to].read_support
      continuing.append({s.read_index for s in support if s.read_index in prev.read_support[phase]})
      first.append({s.read_index for s in support if s.is_first_allele})
    score = prev.score + len(continuing[0] | continuing[1]) + len(first[0] | first[1]) // 2
    if len(continuing[0]) < 2 and len(continuing[1]) < 2:
      score = prev.score
    return _Score(score, (e1[0], e2[0]), (continuing[0] | first[0], continuing[1] | first[1]))

  def _bases_greater(self, a: Tuple[Optional[int], Optional[int]], b: Tuple[Optional[int], Optional[int]]) -> bool:
    """CompareVertexPairByBases (:227-244)."""
    if a[0] is None or a[1] is None:
      return False
    if b[0] is None or b[1] is None:
      return True
    a0, b0 = self.vertices[a[0]].bases, self.vertices[b[0]].bases
    if a0 != b0:
      return a0 > b0
    return self.vertices[a[1]].bases > self.vertices[b[1]].bases

  def _max_score(self, i: int) -> Optional[Tuple[int, int]]:
    verts = self.by_position[self.positions[i]]
    best, best_score = None, 0
    for v1 in verts:
      for v2 in verts:
        s = self.scores.get((v1, v2))
        if s is None:
          continue
        if s.score > best_score:
          best, best_score = (v1, v2), s.score
        elif s.score == best_score and (best is None or self._bases_greater((v1, v2), best)):
          best, best_score = (v1, v2), s.score
    all_equal = all(self.scores[(v1, v2)].score == best_score for v1 in verts for v2 in verts if (v1, v2) in self.scores)
    return None if all_equal else best

  def _assign_phases_to_vertices(self) -> None:
    if not self.scores:
      return
    i = len(self.positions) - 1
    prev_score: Optional[Tuple[int, int]] = None
    while i >= 0:
      cur = None
      while i >= 0:
        cur = self._max_score(i)
        if cur is None:
          i -= 1
        else:
          break
      if prev_score is None:
        prev_score = cur
      n_in_block = 0
      while cur is not None:
        n_in_block += 1
        p1, p2 = cur
        self.vertices[p1].phase, self.vertices[p2].phase = (1, 2) if p1 != p2 else (0, 0)
        if p1 != p2:
          self.vertices[p1].phase, self.vertices[p2].phase = 1, 2
        if cur != prev_score and n_in_block > 1 and self.scores[cur].score == self.scores[prev_score].score:
          self.vertices[p1].phase = self.vertices[p2].phase = 0
          i -= 1
          break
        nxt = self.scores[cur].src
        if nxt not in self.scores:
          if n_in_block == 1:
            self.vertices[p1].phase = self.vertices[p2].phase = 0
          i -= 1
          prev_score = cur
          break
        if nxt == cur:
          i -= 1
          break
        prev_score, cur = cur, nxt
        i -= 1

  # -- the entry point ----------------------------------------------------------------------------------------------------------
  def phase(self, candidates: Sequence[dict], read_keys: Sequence[str]) -> List[int]:
    self._build(candidates, read_keys)
    self.scores: Dict[Tuple[int, int], _Score] = {}
    for i, pos in enumerate(self.positions):
      verts = self.by_position.get(pos, [])
      if i == 0 or not any(self.in_edges.get(v) for v in verts):
        self._update_starting_score(verts)
        continue
      incoming: Set[Tuple[int, int]] = set()
      for v in verts:
        if not self.in_edges.get(v):
          incoming.update((pv, v) for pv in self.by_position.get(self.positions[i - 1], []))
        else:
          incoming.update((s, v) for s in self.in_edges[v])
      keyed = {(self.vertices[s].bases, self.vertices[t].bases): (s, t) for s, t in sorted(incoming)}
      edges = [keyed[k] for k in sorted(keyed)]
      advancing = False
      for e1 in edges:
        for e2 in edges:
          prev = self.scores.get((e1[0], e2[0]))
          if prev is None:
            continue
          score = self._calculate_score(e1, e2)
          if prev.score < score.score:
            advancing = True
          existing = self.scores.get((e1[1], e2[1]))
          if existing is None or existing.score < score.score or (
              existing.score == score.score and self._bases_greater(score.src, existing.src)):
            self.scores[(e1[1], e2[1])] = score
      if i < len(self.positions) - 1 and (not advancing or self._all_scores_the_same(edges)):
        self._update_starting_score(verts)
    self._assign_phases_to_vertices()
    phases = []
    for key in read_keys:
      counts = [0, 0, 0]
      for v in self.read_to_alleles.get(self.read_to_index[key], []):
        counts[self.vertices[v].phase] += 1
      if counts[1] > counts[2] and counts[1] >= self.min_alleles_to_phase:
        phases.append(1)
      elif counts[2] > counts[1] and counts[2] >= self.min_alleles_to_phase:
        phases.append(2)
      else:
        phases.append(0)
    return phases

  def _all_scores_the_same(self, edges: Sequence[Tuple[int, int]]) -> bool:
    lo, hi = 2 ** 31 - 1, 0
    for e1 in edges:
      for e2 in edges:
        s = self.scores.get((e1[1], e2[1]))
        if s is not None:
          lo, hi = min(lo, s.score), max(hi, s.score)
    return not hi - lo > 1


def phase_reads(candidates: Sequence[dict], read_keys: Sequence[str], phase_max_candidates: int = PHASE_MAX_CANDIDATES,
                min_alleles_to_phase: int = MIN_ALLELES_TO_PHASE) -> List[int]:
  """make_examples_core.py:2992-3036: all zeros when there are more candidates than --phase_max_candidates."""
  if phase_max_candidates and len(candidates) > phase_max_candidates:
    return [0] * len(read_keys)
  return DirectPhasing(min_alleles_to_phase).phase(candidates, read_keys)
