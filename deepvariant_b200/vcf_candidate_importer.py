"""--variant_caller vcf_candidate_importer: the candidates are the records of a --proposed_variants VCF; the reads only supply the
evidence (AD / DP / VAF, the supporting read names the pileup encoder marks).  Mirrors

  VcfCandidateImporter.get_candidates / get_candidate_positions          deepvariant/vcf_candidate_importer.py:41-78
  VariantCaller::CallsFromVcf / CallPositionsFromVcf (record selection)  deepvariant/variant_calling.cc:393-476
  is_uncalled_genotype                                                   :383-391
  fetch_vcf_positions / filter_regions_by_vcf (regions without a record are skipped unless gVCF output is on)
                                                                         deepvariant/make_examples_core.py:891-975, 3443-3478

The per-record computation (ComputeVariant) is native: dvb_candidates_from_proposed in csrc/dvb_candidates.cu over the same allele
counter the very-sensitive caller uses.  The reference queries an indexed VCF per region; this reader holds the records of the
whole file by contig (a proposed-variants VCF is a call set, megabytes) and answers the same overlap query.
"""
from __future__ import annotations

import bisect
import ctypes as C
import gzip
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from deepvariant_b200 import _lib
from deepvariant_b200 import candidates as cand


class ProposedVariant:
  __slots__ = ('reference_name', 'start', 'end', 'reference_bases', 'alternate_bases', 'genotype')

  def __init__(self, reference_name: str, start: int, reference_bases: str, alternate_bases: List[str], genotype: Optional[List[int]]):
    self.reference_name, self.start, self.reference_bases, self.alternate_bases = reference_name, start, reference_bases, alternate_bases
    self.end = start + len(reference_bases)
    self.genotype = genotype          # of the FIRST sample column (variant.calls(0)); None when the file has no sample columns

  def is_uncalled_genotype(self) -> bool:
    """is_uncalled_genotype (variant_calling.cc:383-391): first call, at least two alleles, the first two missing."""
    g = self.genotype
    return g is not None and len(g) >= 2 and g[0] == -1 and g[1] == -1


def _parse_genotype(fmt: str, sample: str) -> List[int]:
  keys = fmt.split(':')
  if 'GT' not in keys:
    return []
  fields = sample.split(':')
  i = keys.index('GT')
  if i >= len(fields):
    return []
  return [-1 if a in ('.', '') else int(a) for a in fields[i].replace('|', '/').split('/')]


class ProposedVcfReader:
  """Records of a VCF by contig in file order + vcf_reader->Query(range): the records that overlap a range."""

  def __init__(self, path: str):
    self.path = path
    self.by_contig: Dict[str, List[ProposedVariant]] = {}
    with open(path, 'rb') as probe:
      zipped = probe.read(2) == b'\x1f\x8b'
    with (gzip.open if zipped else open)(path, 'rt') as f:
      for line in f:
        if line.startswith('#'):
          continue
        t = line.rstrip('\n').split('\t')
        if len(t) < 5:
          continue
        alts = [] if t[4] in ('.', '') else t[4].split(',')
        genotype = _parse_genotype(t[8], t[9]) if len(t) >= 10 else None
        self.by_contig.setdefault(t[0], []).append(ProposedVariant(t[0], int(t[1]) - 1, t[3], alts, genotype))
    self._starts = {c: [v.start for v in vs] for c, vs in self.by_contig.items()}
    self._sorted = {c: all(a <= b for a, b in zip(s, s[1:])) for c, s in self._starts.items()}
    self._longest = {c: max((v.end - v.start for v in vs), default=0) for c, vs in self.by_contig.items()}

  def has_contig(self, contig: str) -> bool:
    return contig in self.by_contig

  def query(self, contig: str, start: int, end: int) -> List[ProposedVariant]:
    vs = self.by_contig.get(contig)
    if not vs:
      return []
    if self._sorted[contig]:
      lo = bisect.bisect_left(self._starts[contig], start - self._longest[contig])
      hi = bisect.bisect_left(self._starts[contig], end)
      vs = vs[lo:hi]
    return [v for v in vs if v.start < end and v.end > start]

  def starting_in(self, contig: str, start: int, end: int, skip_uncalled_genotypes: bool = False) -> List[ProposedVariant]:
    """The records CallsFromVcf keeps for a range: overlapping it, starting at or after its start, called (training mode only)."""
    return [v for v in self.query(contig, start, end)
            if v.start >= start and not (skip_uncalled_genotypes and v.is_uncalled_genotype())]

  def positions(self, contig: str) -> List[int]:
    return self._starts.get(contig, [])


def region_has_proposed_variant(reader: ProposedVcfReader, contig: str, start: int, end: int) -> bool:
  """filter_regions_by_vcf for one region: a record STARTS inside it (variant_position is the start base)."""
  starts = reader.positions(contig)
  if reader._sorted.get(contig, True):    # pylint: disable=protected-access
    i = bisect.bisect_left(starts, start)
    return i < len(starts) and starts[i] < end
  return any(start <= s < end for s in starts)


def call_positions_from_vcf(reader: ProposedVcfReader, contig: str, start: int, end: int, skip_uncalled_genotypes: bool = False) -> List[int]:
  """get_candidate_positions: the first pass of --track_ref_reads."""
  return [v.start for v in reader.starting_in(contig, start, end, skip_uncalled_genotypes)]


def _padded(start: int, end: int, padding_pct: int, contig_len: int) -> Tuple[int, int]:
  if padding_pct <= 0:
    return start, end
  pad = int((end - start) * padding_pct / 100)
  return max(start - pad, 0), min(end + pad, contig_len)


def calls_from_vcf(table, ref_reader, contig: str, start: int, end: int, options: cand.CandidateOptions, reader: ProposedVcfReader,
                   rows: Optional[np.ndarray] = None, padding_pct: int = 0, skip_uncalled_genotypes: bool = False) -> cand.NativeCandidates:
  """get_candidates for [start, end) on `contig`: same arguments and result type as candidates.candidates_in_region."""
  lib = _lib.lib()
  if rows is None:
    rows = cand.region_reads(table, contig, start, end, options.max_reads_per_partition, options.random_seed)
  rows = np.ascontiguousarray(rows, dtype=np.int64)
  seq, ptr = cand._contig_buffer(ref_reader, contig)   # pylint: disable=protected-access
  end = min(end, len(seq))
  region = (start, end)
  start, end = _padded(start, end, padding_pct, len(seq))
  if not reader.has_contig(contig):
    import logging
    logging.warning('%s:%d-%d cannot be found in proposed VCF header. Skip this region.', contig, start + 1, end)
  proposed = reader.starting_in(contig, start, end, skip_uncalled_genotypes)
  positions = np.asarray([v.start for v in proposed], dtype=np.int32) if options.track_ref_reads else np.zeros(0, np.int32)
  alleles: List[bytes] = []
  first = [0]
  for v in proposed:
    alleles.append(v.reference_bases.encode())
    alleles.extend(a.encode() for a in v.alternate_bases)
    first.append(len(alleles))
  begin = np.zeros(len(alleles) + 1, dtype=np.int64)
  np.cumsum([len(a) for a in alleles], out=begin[1:])
  chars = b''.join(alleles)
  starts = np.asarray([v.start for v in proposed], dtype=np.int64)
  first_a = np.asarray(first, dtype=np.int32)
  co = options.to_c()
  h = C.c_void_p()
  _lib.check(lib.dvb_candidates_from_proposed(
      table.handle, contig.encode(), ptr, len(seq), start, end, rows.ctypes.data_as(C.c_void_p), len(rows), C.byref(co),
      positions.ctypes.data_as(C.c_void_p), len(positions), len(proposed), starts.ctypes.data_as(C.c_void_p),
      first_a.ctypes.data_as(C.c_void_p), begin.ctypes.data_as(C.c_void_p), chars, C.byref(h)))
  return cand.NativeCandidates(h, keep=region if padding_pct > 0 else None, interval=(start, end))
