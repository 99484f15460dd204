"""Population allele frequencies of candidate alleles — what fills DeepVariantCall.allele_frequency for the allele_frequency
pileup channel (make_examples --population_vcfs).  Restates deepvariant/allele_frequency.py:

  get_allele_frequency / get_ref_allele_frequency   :38-76
  get_ref_haplotype_and_offset                      :79-119
  update_haplotype                                  :122-172
  match_candidate_and_cohort_haplotypes             :175-246
  find_matching_allele_frequency                    :249-327
  add_allele_frequencies_to_candidates              :384-421

A candidate alt matches a cohort alt when splicing either into the reference span that covers the candidate and all overlapping
cohort records gives the same haplotype.  The cohort VCF is read as text (bgzip is gzip-compatible); its AF values are float32 as
htslib stores them (nucleus number_value = the float widened to double).
"""
from __future__ import annotations

import gzip
import struct
from typing import Dict, Iterable, List, Optional, Sequence

from deepvariant_b200.postprocess_variants import simplify_alleles
from deepvariant_b200.protos import DeepVariantCall


class CohortVariant:
  __slots__ = ('reference_name', 'start', 'end', 'reference_bases', 'alternate_bases', 'af')

  def __init__(self, reference_name: str, start: int, reference_bases: str, alternate_bases: List[str], af: Optional[List[float]]):
    self.reference_name, self.start, self.reference_bases, self.alternate_bases, self.af = reference_name, start, reference_bases, alternate_bases, af
    self.end = start + len(reference_bases)


def _f32(text: str) -> float:
  return struct.unpack('<f', struct.pack('<f', float(text)))[0]


class PopulationVcfReader:
  """The records of one population VCF, by contig, sorted by start (vcf.VcfReader.query: records overlapping a range)."""

  def __init__(self, path: str):
    self.by_contig: Dict[str, List[CohortVariant]] = {}
    opener = gzip.open if open(path, 'rb').read(2) == b'\x1f\x8b' else open
    with opener(path, 'rt') as f:
      for line in f:
        if line.startswith('#'):
          continue
        t = line.rstrip('\n').split('\t')
        if len(t) < 8:
          continue
        af = None
        for kv in t[7].split(';'):
          if kv.startswith('AF='):
            af = [_f32(x) for x in kv[3:].split(',')]
        alts = [] if t[4] == '.' else t[4].split(',')
        self.by_contig.setdefault(t[0], []).append(CohortVariant(t[0], int(t[1]) - 1, t[3], alts, af))

  def query(self, contig: str, start: int, end: int) -> List[CohortVariant]:
    if contig not in self.by_contig:
      return []   # the reference logs "population_vcf does not have contig" and goes on with no cohort variants
    return [v for v in self.by_contig[contig] if v.start < end and v.end > max(start, 0)]


def make_population_vcf_readers(paths: Sequence[str]) -> Dict[str, PopulationVcfReader]:
  """make_population_vcf_readers (:330-381): one VCF serves every contig (key '*'); several must hold one contig each."""
  if len(paths) == 1:
    return {'*': PopulationVcfReader(paths[0])}
  out: Dict[str, PopulationVcfReader] = {}
  for path in paths:
    reader = PopulationVcfReader(path)
    if not reader.by_contig:
      continue
    contig = next(iter(reader.by_contig))     # the contig of the file's first record
    if contig in out:
      raise ValueError('Variants on %s are included in multiple VCFs' % contig)
    out[contig] = reader
  return out


def get_allele_frequency(variant: CohortVariant, index: int) -> float:
  if variant.af:
    if index < len(variant.af):
      return variant.af[index]
    raise ValueError('Invalid index', index, 'for the info[AF] field', variant.af)
  raise ValueError('Variant does not have an AF field')


def get_ref_allele_frequency(variant: CohortVariant) -> float:
  s = 0
  for alt_idx, _ in enumerate(variant.alternate_bases):
    s += get_allele_frequency(variant, alt_idx)
  return 1 - s


def update_haplotype(variant, reference_haplotype: str, reference_offset: int) -> List[dict]:
  if variant.start < reference_offset:
    raise ValueError('The starting position of a variant is smaller than its corresponding reference offset', variant.start, reference_offset)
  offset_start = variant.start - reference_offset
  offset_suffix = variant.start + len(variant.reference_bases) - reference_offset
  return [{'haplotype': reference_haplotype[:offset_start] + alt + reference_haplotype[offset_suffix:], 'alt': alt, 'variant': variant}
          for alt in variant.alternate_bases]


def _simplified(variant):
  s = simplify_alleles(variant.reference_bases, *variant.alternate_bases)
  return variant.start, s[0]


def match_candidate_and_cohort_haplotypes(candidate_haps: Sequence[dict], cohort_haps_and_freqs: Sequence[dict]) -> Dict[str, float]:
  d: Dict[str, float] = {}
  for candidate_obj in candidate_haps:
    candidate_alt = candidate_obj['alt']
    candidate_variant = candidate_obj['variant']
    for cohort_obj in cohort_haps_and_freqs:
      if candidate_obj['haplotype'] == cohort_obj['haplotype']:
        cohort_variant = cohort_obj['variant']
        d[candidate_alt] = get_allele_frequency(cohort_variant, list(cohort_variant.alternate_bases).index(cohort_obj['alt']))
        if not d.get(candidate_variant.reference_bases):
          d[candidate_variant.reference_bases] = get_ref_allele_frequency(cohort_variant)
    if not d.get(candidate_alt):
      d[candidate_alt] = 0
  if sum(d.values()) == 0:
    candidate = candidate_haps[0]['variant']
    s_start, s_ref = _simplified(candidate)
    for cohort_obj in cohort_haps_and_freqs:
      c_start, c_ref = _simplified(cohort_obj['variant'])
      if s_start == c_start and s_ref == c_ref:
        d[s_ref] = get_ref_allele_frequency(cohort_obj['variant'])
    if not d.get(candidate.reference_bases):
      d[candidate.reference_bases] = 1
  return d


def find_matching_allele_frequency(variant, population_vcf_reader: PopulationVcfReader, ref_reader, padding_bases: int = 0) -> Dict[str, float]:
  cohort_variants = population_vcf_reader.query(variant.reference_name, variant.start - padding_bases, variant.end + padding_bases)
  start = min([variant.start] + [cv.start for cv in cohort_variants])
  end = max([variant.end] + [cv.end for cv in cohort_variants])
  if not cohort_variants or start < 0 or end > ref_reader.n_bases(variant.reference_name):
    # min() over no cohort variants / an invalid FASTA range: the reference's ValueError branch
    d = {variant.reference_bases: 1}
    for alt in variant.alternate_bases:
      d[alt] = 0
    return d
  reference_haplotype = ref_reader.query(variant.reference_name, start, end)
  candidate_haps = update_haplotype(variant, reference_haplotype, start)
  cohort_haps: List[dict] = []
  for cv in cohort_variants:
    cohort_haps.extend(update_haplotype(cv, reference_haplotype, start))
  return match_candidate_and_cohort_haplotypes(candidate_haps, cohort_haps)


def add_allele_frequencies_to_candidates(candidates: Iterable[DeepVariantCall], population_vcf_reader: Optional[PopulationVcfReader],
                                         ref_reader) -> List[DeepVariantCall]:
  out = []
  for candidate in candidates:
    if population_vcf_reader:
      d = find_matching_allele_frequency(candidate.variant, population_vcf_reader, ref_reader)
    else:
      d = {candidate.variant.reference_bases: 1}
      for alt in candidate.variant.alternate_bases:
        d[alt] = 0
    candidate.allele_frequency = {k: struct.unpack('<f', struct.pack('<f', float(v)))[0] for k, v in d.items()}   # map<string, float>
    out.append(candidate)
  return out
