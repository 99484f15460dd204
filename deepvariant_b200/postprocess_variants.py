"""CallVariantsOutput TFRecords -> VCF: the single-sample, single-site path of the reference's postprocess_variants stage
(SURVEY.md 8(f) "next" row #4), so that `run_deepvariant --output_vcf` closes without the reference.

Mirrors deepvariant/postprocess_variants.py with its default flags (qual_filter 1.0, multi_allelic_qual_filter 1.0,
cnn_homref_call_min_gq 20, multiallelic_mode 'product', group_variants, haplotype resolution on):

  sort_cvos                         postprocess_variants.cc:66-80 (stable sort by contig order, start, end; nucleus CompareVariants)
  group / _sort_grouped_variants    postprocess_variants.py:1467-1488, 1380-1382
  merge_predictions                 :1167-1308 (get_alt_alleles_to_remove :806-860, prune_alleles :944-970, product fusion :1234-1278,
                                    normalize_predictions :1057-1067, simplify_variant_alleles variant_utils.py:480-556)
  add_call_to_variant               :555-608 (most_likely_genotype :380-461, compute_quals :611-646, uncall_gt_if_no_ad :464-471,
                                    compute_filter_fields dv_vcf_constants.py:205-227, uncall_homref_gt_if_lowqual :474-495)
  maybe_resolve_conflicting_variants deepvariant/haplotypes.py:63-475
  write_vcf                         nucleus VcfWriter (third_party/nucleus/io/vcf_writer.cc:178-200 QUAL rounded to one decimal,
                                    vcf_conversion.cc:1176-1240 GL -> zero-shifted, truncated PL) with htslib's number formatting
Not restated: gVCF output, haploid contigs / PAR regions, the multiallelic keras model, small-model CVOs, phase-set stitching (PS), PON filters,
methylation FORMAT fields.  Pinned by the reference's golden.postprocess_single_site_input -> golden.postprocess_single_site_output.vcf
(tests/test_postprocess.py)."""
from __future__ import annotations

import copy
import dataclasses
import itertools
import math
from typing import Dict, Iterable, Iterator, List, Optional, Sequence, Tuple

import numpy as np

from deepvariant_b200 import candidates as cand, protos, tfrecord

VERSION = '1.10.0'
PASS, REF_FILTER, QUAL_FILTER, NO_CALL = 'PASS', 'RefCall', 'LowQual', 'NoCall'    # dv_vcf_constants.py:41-47
_MAX_CONFIDENCE = 1.0 - 1.25e-10            # genomics_math.py:100
_QUAL_PRECISION = 7
_FILTERED_ALT_PROB = -9.0
_MAX_OVERLAPPING_VARIANTS_TO_RESOLVE = 12   # haplotypes.py:63


@dataclasses.dataclass
class OutVariant:
  """The fields of nucleus Variant + its single VariantCall that this stage reads and writes."""
  reference_name: str
  start: int
  end: int
  reference_bases: str
  alternate_bases: List[str]
  info: Dict[str, list]                     # calls[0].info: AD, DP, VAF
  call_set_name: str = ''
  genotype: List[int] = dataclasses.field(default_factory=lambda: [-1, -1])
  genotype_likelihood: List[float] = dataclasses.field(default_factory=list)
  gq: Optional[int] = None
  quality: float = 0.0
  filter: List[str] = dataclasses.field(default_factory=list)
  is_phased: bool = False
  variant_info: Dict[str, list] = dataclasses.field(default_factory=dict)      # Variant.info (PS_CONTIG, ALT_PS from make_examples' phasing)


@dataclasses.dataclass
class Cvo:
  variant: OutVariant
  alt_allele_indices: List[int]
  genotype_probabilities: List[float]
  raw: bytes = b''


def parse_cvo(record: bytes) -> Cvo:
  variant_bytes, indices, probs = protos.parse_call_variants_output(record)
  c = cand.canonical_call(protos.f_bytes(1, variant_bytes))
  v = OutVariant(c['contig'], c['start'], c['end'], c['ref'], list(c['alts']), {k: list(vs) for k, vs in c['info'].items()},
                 call_set_name=c['call_set_name'], genotype=list(c['genotype']))
  for fn, wt, val, _ in protos.iter_fields(variant_bytes):          # Variant.info = 10: map<string, ListValue>
    if fn == 10:
      key, vals = '', []
      for f2, w2, v2, _ in protos.iter_fields(bytes(val)):
        if f2 == 1:
          key = bytes(v2).decode()
        elif f2 == 2:
          vals = [cand._value(bytes(v3)) for f3, w3, v3, _ in protos.iter_fields(bytes(v2)) if f3 == 1]   # pylint: disable=protected-access
      v.variant_info[key] = vals
  return Cvo(v, list(indices), [float(p) for p in probs], bytes(record))


# ---- genomics_math ------------------------------------------------------------------------------------------------------------
def ptrue_to_bounded_phred(ptrue: float) -> float:
  if not 0 <= ptrue <= 1:
    raise ValueError('ptrue must be between zero and one: {}'.format(ptrue))
  return -10 * math.log10(1.0 - min(ptrue, _MAX_CONFIDENCE))


def perror_to_bounded_log10_perror(perror: float) -> float:
  if not 0 <= perror <= 1:
    raise ValueError('perror must be between zero and one: {}'.format(perror))
  return math.log10(max(perror, 1.0 - _MAX_CONFIDENCE))


def log10sumexp(xs: Sequence[float]) -> float:
  m = max(xs)
  return m + math.log10(sum(pow(10.0, x - m) for x in xs))


def normalize_log10_probs(xs: Sequence[float]) -> np.ndarray:
  xs = np.array(xs)
  if np.max(xs) > 0.0:
    raise ValueError('log10_probs all must be <= 0', xs)
  return np.minimum(xs - log10sumexp(xs), 0.0)


# ---- variant_utils -------------------------------------------------------------------------------------------------------------
def genotype_order_in_likelihoods(num_alts: int) -> Iterator[Tuple[int, int]]:
  for j in range(num_alts + 1):
    for i in range(j + 1):
      yield i, j


def genotype_likelihood_index(allele_indices: Sequence[int]) -> int:
  g1, g2 = sorted(allele_indices)
  return g1 + (g2 * (g2 + 1) // 2)


def allele_indices_for_genotype_likelihood_index(gl_index: int) -> Tuple[int, int]:
  num_alts = 1
  while genotype_likelihood_index([num_alts, num_alts]) < gl_index:
    num_alts += 1
  return list(genotype_order_in_likelihoods(num_alts))[gl_index]


def simplify_alleles(*alleles: str) -> Tuple[str, ...]:
  shortest = min(len(a) for a in alleles)
  common = 0
  for i in range(1, shortest):
    if len({a[-i] for a in alleles}) != 1:
      break
    common = i
  return tuple(a[:-common] for a in alleles) if common else alleles


def simplify_variant_alleles(v: OutVariant) -> OutVariant:
  s = simplify_alleles(v.reference_bases, *v.alternate_bases)
  v.reference_bases, v.alternate_bases = s[0], list(s[1:])
  v.end = v.start + len(v.reference_bases)
  return v


# ---- merge_predictions ----------------------------------------------------------------------------------------------------------
def compute_quals(predictions: Sequence[float], prediction_index: int) -> Tuple[int, float]:
  gq = int(np.around(ptrue_to_bounded_phred(predictions[prediction_index])))
  qual = ptrue_to_bounded_phred(min(sum(predictions[1:]), 1.0))
  return gq, round(qual, _QUAL_PRECISION)


def expected_alt_allele_indices(num_alternate_bases: int) -> List[List[int]]:
  n = num_alternate_bases + 1
  lists = [sorted(set(x) - {0}) for x in itertools.combinations(range(n), 2)]
  return sorted([i - 1 for i in idx] for idx in lists)


def get_alt_alleles_to_remove(cvos: Sequence[Cvo], qual_filter: float) -> set:
  to_remove = set()
  if not qual_filter or not cvos:
    return to_remove
  max_qual, max_qual_allele = None, None
  canonical = cvos[0].variant
  for c in cvos:
    if len(c.alt_allele_indices) == 1:
      _, qual = compute_quals(c.genotype_probabilities, 0)
      allele = canonical.alternate_bases[c.alt_allele_indices[0]]
      if max_qual is None or max_qual < qual:
        max_qual, max_qual_allele = qual, allele
      if qual < qual_filter:
        to_remove.add(allele)
  if len(to_remove) == len(canonical.alternate_bases):
    to_remove -= {max_qual_allele}
  return to_remove


_ALT_ALLELE_INDEXED_FORMAT_FIELDS = (('AD', True), ('VAF', False), ('MF', True), ('MD', True), ('NAD', True), ('NAF', False))


def prune_alleles(v: OutVariant, to_remove: set) -> OutVariant:
  if not to_remove:
    return v
  new = copy.deepcopy(v)
  keep_alt = [a not in to_remove for a in v.alternate_bases]
  for field, ref_is_zero in _ALT_ALLELE_INDEXED_FORMAT_FIELDS:
    if field in new.info:
      vals = new.info[field]
      new.info[field] = [x for i, x in enumerate(vals) if (i == 0 or keep_alt[i - 1] if ref_is_zero else keep_alt[i])]
  new.alternate_bases = [a for a, k in zip(v.alternate_bases, keep_alt) if k]
  return new


def normalize_predictions(predictions: Sequence[float]) -> List[float]:
  if sum(predictions) == 0:
    predictions = [1.0] * len(predictions)
  denominator = sum(i if i != _FILTERED_ALT_PROB else 0.0 for i in predictions) or 1.0
  return [i / denominator if i != _FILTERED_ALT_PROB else 0.0 for i in predictions]


def is_valid_call_variants_outputs(cvos: Sequence[Cvo]) -> bool:
  if not cvos:
    return True
  if sorted(c.alt_allele_indices for c in cvos) != expected_alt_allele_indices(len(cvos[0].variant.alternate_bases)):
    return False
  f = cvos[0].variant
  key = (f.reference_name, f.start, f.end, f.reference_bases, f.alternate_bases)
  return all((c.variant.reference_name, c.variant.start, c.variant.end, c.variant.reference_bases, c.variant.alternate_bases) == key
             for c in cvos[1:])


def correct_nonautosome_probabilities(probabilities: List[float], n_alts: int) -> List[float]:
  """postprocess_variants.py:1070-1091: every heterozygous genotype of a haploid site gets probability zero, the rest renormalised."""
  probabilities = list(probabilities)
  index = 0
  for h1 in range(n_alts + 1):
    for h2 in range(h1 + 1):
      if h2 != h1:
        if len(probabilities) <= index:
          raise ValueError("Probabilties array doesn't match alt alleles.")
        probabilities[index] = 0
      index += 1
  total = sum(probabilities) or 1.0
  return [p / total for p in probabilities]


def read_bed(path: str) -> List[Tuple[str, int, int]]:
  out = []
  with open(path) as f:
    for line in f:
      parts = line.split()
      if len(parts) >= 3 and not line.startswith(('#', 'track', 'browser')):
        out.append((parts[0], int(parts[1]), int(parts[2])))
  return out


def _is_haploid_site(v: OutVariant, haploid_contigs: Sequence[str], par_regions: Sequence[Tuple[str, int, int]]) -> bool:
  """is_non_autosome and not is_in_regions(variant, par_regions) (postprocess_variants.py:1094-1112)."""
  if not haploid_contigs or v.reference_name not in haploid_contigs:
    return False
  return not any(c == v.reference_name and s < v.end and v.start < e for c, s, e in par_regions)


def merge_predictions(cvos: Sequence[Cvo], qual_filter: float = 1.0, multiallelic_mode: str = 'product', haploid_contigs: Sequence[str] = (),
                      par_regions: Sequence[Tuple[str, int, int]] = ()) -> Tuple[OutVariant, List[float]]:
  canonical, predictions = _merge_predictions(cvos, qual_filter, multiallelic_mode)
  if _is_haploid_site(canonical, haploid_contigs, par_regions):
    predictions = correct_nonautosome_probabilities(predictions, len(canonical.alternate_bases))
  return canonical, predictions


def _merge_predictions(cvos: Sequence[Cvo], qual_filter: float = 1.0, multiallelic_mode: str = 'product') -> Tuple[OutVariant, List[float]]:
  if not cvos:
    raise ValueError('Expected 1 or more call_variants_outputs.')
  if not is_valid_call_variants_outputs(cvos):
    raise ValueError('`call_variants_outputs` did not pass sanity check.')
  first = cvos[0]
  canonical = copy.deepcopy(first.variant)
  if len(cvos) == 1:
    return simplify_variant_alleles(canonical), list(first.genotype_probabilities)
  to_remove = get_alt_alleles_to_remove(cvos, qual_filter)
  original_alts = list(canonical.alternate_bases)
  canonical = prune_alleles(canonical, to_remove)
  alleles = [canonical.reference_bases] + list(canonical.alternate_bases)
  ordering = [(alleles[i], alleles[j]) for i, j in genotype_order_in_likelihoods(len(canonical.alternate_bases))]
  if multiallelic_mode == 'product':
    example_info = []
    for c in cvos:
      example_alts = frozenset(original_alts[i] for i in c.alt_allele_indices)
      if to_remove & example_alts:
        continue
      example_info.append((c.genotype_probabilities, example_alts))
    predictions = []
    for a1, a2 in ordering:
      probs = [p[int(a1 in alts) + int(a2 in alts)] for p, alts in example_info]
      predictions.append(_FILTERED_ALT_PROB if _FILTERED_ALT_PROB in probs else float(np.prod(probs)))
    normalized = normalize_predictions(predictions)
  elif multiallelic_mode == 'min':
    flat: Dict[Tuple[str, str], List[float]] = {}
    for c in cvos:
      set1 = frozenset([first.variant.reference_bases])
      set2 = frozenset(original_alts[i] for i in c.alt_allele_indices)
      if to_remove & set2:
        continue
      p11, p12, p22 = c.genotype_probabilities
      for s1, s2, p in ((set1, set1, p11), (set1, set2, p12), (set2, set2, p22)):
        for key in itertools.product(s1, s2):
          flat.setdefault(key, []).append(p)
    predictions = [min([x for x in flat.get(k, []) if x != _FILTERED_ALT_PROB] or [0]) for k in ordering]
    if sum(predictions) == 0:
      predictions = [1.0] * len(predictions)
    normalized = normalize_predictions(predictions)
  else:
    raise ValueError(f'unknown multiallelic_mode {multiallelic_mode}')
  return simplify_variant_alleles(canonical), normalized


# ---- add_call_to_variant -------------------------------------------------------------------------------------------------------------
def most_likely_genotype(predictions: Sequence[float], n_alleles: int = 2) -> Tuple[int, List[int]]:
  if n_alleles < 2:
    raise ValueError('n_alleles must be >= 2 but got', n_alleles)
  index_of_max = int(np.argmax(predictions))
  index = 0
  for h1 in range(0, n_alleles + 1):
    for h2 in range(0, h1 + 1):
      if index == index_of_max:
        return index, [h2, h1]
      index += 1
  raise ValueError('No corresponding GenotypeType for predictions', predictions)


def maybe_phase_genotype(v: OutVariant, genotype: List[int]) -> Tuple[bool, List[int]]:
  """postprocess_variants.py:498-552: Variant.info['ALT_PS'] holds the haplotype (0 = none, 1, 2) of REF and of every ALT."""
  if not (v.variant_info.get('PS_CONTIG') and v.variant_info.get('ALT_PS')):
    return False, genotype
  phase_info = [int(p) for p in v.variant_info['ALT_PS']]
  if max(genotype) >= len(phase_info):
    return False, genotype
  h1, h2 = phase_info[genotype[0]], phase_info[genotype[1]]
  is_phased = 0 not in (h1, h2) and h1 != h2
  if is_phased:
    genotype = [genotype[h1 - 1], genotype[h2 - 1]]
  return is_phased, genotype


def compute_filter_fields(v: OutVariant, min_quality: float) -> List[str]:
  gt = set(v.genotype)
  if gt == {-1}:
    return [NO_CALL]
  if gt == {0}:
    return [REF_FILTER]
  if v.quality < min_quality:
    return [QUAL_FILTER]
  return [PASS]


def add_call_to_variant(v: OutVariant, predictions: Sequence[float], qual_filter: float = 1.0, sample_name: Optional[str] = None,
                        cnn_homref_call_min_gq: float = 20.0) -> OutVariant:
  n_alleles = len(v.alternate_bases) + 1
  index, genotype = most_likely_genotype(predictions, n_alleles=n_alleles)
  v.gq, v.quality = compute_quals(predictions, index)
  v.call_set_name = sample_name
  v.is_phased, genotype = maybe_phase_genotype(v, genotype)
  v.genotype = genotype
  v.genotype_likelihood = [perror_to_bounded_log10_perror(p) for p in predictions]
  if sum(v.info.get('AD', [])) == 0:               # uncall_gt_if_no_ad
    v.genotype = [-1, -1]
    v.genotype_likelihood = [0, 0]
    v.gq = 0
  v.filter = compute_filter_fields(v, qual_filter)
  if v.filter == [REF_FILTER] and v.gq < cnn_homref_call_min_gq:      # uncall_homref_gt_if_lowqual
    v.genotype = [-1, -1]
    v.filter = [NO_CALL]
  return v


# ---- haplotypes.maybe_resolve_conflicting_variants ------------------------------------------------------------------------------------
def _nonref_genotype_count(v: OutVariant) -> int:
  return sum(g > 0 for g in v.genotype)


def _group_overlapping_variants(sorted_variants: Iterable[OutVariant]) -> Iterator[List[OutVariant]]:
  curr, prev_chrom, prev_max_end = [], None, -1
  for v in sorted_variants:
    if v.reference_name != prev_chrom or v.start >= prev_max_end:
      if curr:
        yield curr
      curr, prev_chrom, prev_max_end = [v], v.reference_name, v.end
    else:
      curr.append(v)
      prev_max_end = max(prev_max_end, v.end)
  if curr:
    yield curr


def _all_variants_compatible(variants: Sequence[OutVariant], counts: Sequence[int], ploidy: int = 2) -> bool:
  min_start = min(v.start for v in variants)
  span = np.zeros(max(v.end - min_start for v in variants), dtype=int)
  for cnt, v in zip(counts, variants):
    span[v.start - min_start:v.end - min_start] += cnt
  return bool(np.all(span <= ploidy))


def _allele_indices_with_num_alts(v: OutVariant, num_alts: int) -> List[Tuple[int, int]]:
  n = len(v.alternate_bases)
  if num_alts == 0:
    return [(0, 0)]
  if num_alts == 1:
    return [(0, i) for i in range(1, n + 1)]
  return [(i, j) for i in range(1, n + 1) for j in range(i, n + 1)]


def _resolve_overlapping_variants(variants: List[OutVariant], qual_filter: float) -> Iterator[OutVariant]:
  if len(variants) == 1:
    yield variants[0]
    return
  counts = [_nonref_genotype_count(v) for v in variants]
  if _all_variants_compatible(variants, counts) or len(variants) > _MAX_OVERLAPPING_VARIANTS_TO_RESOLVE:
    yield from variants
    return
  valid = [conf for conf in itertools.product([0, 1, 2], repeat=len(variants)) if _all_variants_compatible(variants, conf)]
  n_likelihoods = [genotype_likelihood_index((len(v.alternate_bases),) * 2) + 1 for v in variants]
  containers: List[List[List[float]]] = [[[] for _ in range(n)] for n in n_likelihoods]
  best_config, best_likelihood = None, None
  for conf in valid:
    for config in itertools.product(*[_allele_indices_with_num_alts(v, k) for v, k in zip(variants, conf)]):
      likelihood = 0
      for v, alleles in zip(variants, config):
        likelihood += v.genotype_likelihood[genotype_likelihood_index(alleles)]
      if best_likelihood is None or likelihood > best_likelihood:
        best_likelihood, best_config = likelihood, config
      for cont, alleles in zip(containers, config):
        cont[genotype_likelihood_index(alleles)].append(likelihood)
  scaled = []
  for cont in containers:
    if not all(bool(x) for x in cont):
      raise ValueError('All genotypes must have some probability mass: {}'.format(cont))
    scaled.append(normalize_log10_probs([log10sumexp(x) for x in cont]))
  marginal = tuple(allele_indices_for_genotype_likelihood_index(int(np.argmax(s))) for s in scaled)
  if marginal == best_config:
    for v, alleles, gls in zip(variants, best_config, scaled):
      new = copy.deepcopy(v)
      new.genotype = list(alleles)
      new.genotype_likelihood = [float(g) for g in gls]
      new.filter = compute_filter_fields(new, qual_filter)
      yield new
  else:
    yield from variants


def maybe_resolve_conflicting_variants(sorted_variants: Iterable[OutVariant], qual_filter: float = 1.0) -> Iterator[OutVariant]:
  for group in _group_overlapping_variants(sorted_variants):
    if len(group) == 1:
      yield group[0]
      continue
    reference_calls = [c for c in group if _nonref_genotype_count(c) == 0]
    variant_calls = [c for c in group if _nonref_genotype_count(c) > 0]
    resolved = []
    for sub in _group_overlapping_variants(variant_calls):
      resolved.extend(_resolve_overlapping_variants(sub, qual_filter))
    yield from sorted(reference_calls + resolved, key=lambda v: (v.reference_name, v.start, v.end))


# ---- the stage -----------------------------------------------------------------------------------------------------------------------
def sort_cvos(cvos: List[Cvo], contig_order: Sequence[str]) -> List[Cvo]:
  pos = {c: i for i, c in enumerate(contig_order)}
  return sorted(cvos, key=lambda c: (pos.get(c.variant.reference_name, len(pos)), c.variant.start, c.variant.end))   # stable


def call_variants_outputs_to_variants(cvos: Sequence[Cvo], sample_name: str, qual_filter: float = 1.0,
                                      multi_allelic_qual_filter: float = 1.0, multiallelic_mode: str = 'product',
                                      group_variants: bool = True, haploid_contigs: Sequence[str] = (),
                                      par_regions: Sequence[Tuple[str, int, int]] = ()) -> Iterator[OutVariant]:
  """group_call_variants_outputs (:1467-1488): --group_variants groups by the variant's range; without it itertools.groupby
  falls back to equality of consecutive records (the vcf_candidate_importer flow, where one range can hold several variants)."""
  key = (lambda c: (c.variant.reference_name, c.variant.start, c.variant.end)) if group_variants else (lambda c: c.raw)
  for _, group in itertools.groupby(cvos, key=key):
    outputs = sorted(group, key=lambda c: sorted(c.alt_allele_indices))
    canonical, predictions = merge_predictions(outputs, multi_allelic_qual_filter, multiallelic_mode, haploid_contigs, par_regions)
    yield add_call_to_variant(canonical, predictions, qual_filter=qual_filter, sample_name=sample_name)


def _fmt_float(x: float) -> str:
  """htslib's kputd as VCF text uses it: the value as a C float, six significant digits, no trailing zeros."""
  return '%g' % float(np.float32(x))


def vcf_header_lines(contigs: Sequence[Tuple[str, int]], sample_name: str) -> List[str]:
  """dv_vcf_constants.deepvariant_header (deepvariant/dv_vcf_constants.py:60-202) as the writer prints it."""
  lines = [
      '##fileformat=VCFv4.2',
      '##FILTER=<ID=PASS,Description="All filters passed">',
      '##FILTER=<ID=RefCall,Description="Genotyping model thinks this site is reference.">',
      '##FILTER=<ID=LowQual,Description="Confidence in this variant being real is below calling threshold.">',
      '##FILTER=<ID=NoCall,Description="Site has depth=0 resulting in no call.">',
      '##INFO=<ID=END,Number=1,Type=Integer,Description="End position (for use with symbolic alleles)">',
      '##FORMAT=<ID=GT,Number=1,Type=String,Description="Genotype">',
      '##FORMAT=<ID=GQ,Number=1,Type=Integer,Description="Conditional genotype quality">',
      '##FORMAT=<ID=DP,Number=1,Type=Integer,Description="Read depth">',
      '##FORMAT=<ID=MIN_DP,Number=1,Type=Integer,Description="Minimum DP observed within the GVCF block.">',
      '##FORMAT=<ID=AD,Number=R,Type=Integer,Description="Read depth for each allele">',
      '##FORMAT=<ID=VAF,Number=A,Type=Float,Description="Variant allele fractions.">',
      '##FORMAT=<ID=PL,Number=G,Type=Integer,Description="Phred-scaled genotype likelihoods rounded to the closest integer">',
      '##FORMAT=<ID=PS,Number=1,Type=Integer,Description="Phase set">',
      '##FORMAT=<ID=MF,Number=R,Type=Float,Description="Methylation fraction for each of the reference and alternate allele">',
      '##FORMAT=<ID=MD,Number=R,Type=Integer,Description="Methylation depth for each of the reference and alternate allele">',
      '##FORMAT=<ID=MT,Number=1,Type=String,Description="Methylation type: 0/0=Unmethylated, 0/1=Heterozygous, 1/1=Methylated">',
      '##FORMAT=<ID=MI,Number=1,Type=Float,Description="Allele-specific methylation score: p-value for Wilcoxon Rank-Sum test based on the '
      'observed difference in methylation between haplotypes.">',
      '##FORMAT=<ID=MED_DP,Number=1,Type=Integer,Description="Median DP observed within the GVCF block rounded to the nearest integer.">',
      f'##DeepVariant_version={VERSION}',
  ]
  lines += [f'##contig=<ID={name},length={n}>' for name, n in contigs]
  lines.append('#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t' + sample_name)
  return lines


def vcf_line(v: OutVariant) -> str:
  qual = math.floor(v.quality * 10 + 0.5) / 10                     # VcfWriter::Write with round_qual_values
  keys, vals = ['GT'], [('|' if v.is_phased else '/').join('.' if g < 0 else str(g) for g in v.genotype)]
  if v.gq is not None:
    keys.append('GQ')
    vals.append(str(int(v.gq)))
  if 'DP' in v.info:
    keys.append('DP')
    vals.append(','.join(str(int(x)) for x in v.info['DP']))
  if 'AD' in v.info:
    keys.append('AD')
    vals.append(','.join(str(int(x)) for x in v.info['AD']))
  if 'VAF' in v.info:
    keys.append('VAF')
    vals.append(','.join(_fmt_float(x) for x in v.info['VAF']))
  if v.genotype_likelihood:
    m = max(v.genotype_likelihood)
    keys.append('PL')
    vals.append(','.join(str(int(-10 * (x - m))) for x in v.genotype_likelihood))     # ZeroShiftLikelihoods, truncation to int
  return '\t'.join([v.reference_name, str(v.start + 1), '.', v.reference_bases, ','.join(v.alternate_bases) or '.', _fmt_float(qual),
                    ';'.join(v.filter) or '.', '.', ':'.join(keys), ':'.join(vals)])


def get_sample_name(cvos: Sequence[Cvo], flag: str = '') -> str:
  """get_sample_name (postprocess_variants.py:1633-1660): the flag, else the call_set_name of the first record, else 'default'."""
  if flag:
    return flag
  if cvos and cvos[0].variant.call_set_name:
    return cvos[0].variant.call_set_name
  return 'default'


# ---- --cpus: the CVO -> Variant conversion over worker processes (deepvariant/postprocess_variants.py --cpus, :160-172, 1998-2087) --------
def cvo_range_key(record: bytes) -> Tuple[str, int, int]:
  """(reference_name, start, end) of a serialized CallVariantsOutput without building the variant: CallVariantsOutput.variant = 1,
  Variant.reference_name = 14 / start = 16 / end = 13 - all the parent process needs to sort the records and cut them into chunks."""
  name, start, end = '', 0, 0
  for fn, wt, val, _ in protos.iter_fields(record):
    if fn == 1:
      for f2, w2, v2, _ in protos.iter_fields(bytes(val)):
        if f2 == 14:
          name = bytes(v2).decode()
        elif f2 == 16:
          start = v2 if v2 < (1 << 63) else v2 - (1 << 64)
        elif f2 == 13:
          end = v2 if v2 < (1 << 63) else v2 - (1 << 64)
      break
  return name, start, end


def independent_chunks(keys: Sequence[Tuple[str, int, int]], target: int) -> List[Tuple[int, int]]:
  """Cuts SORTED (contig, start, end) keys into [begin, end) chunks of about `target` records such that no variant range of one chunk
  overlaps a range of another - grouping by range, multi-allelic merging and the resolution of overlapping variants never look across
  such a cut, so the chunks can be converted independently and concatenated."""
  chunks, begin, reach, contig = [], 0, -1, None
  for i, (c, s, e) in enumerate(keys):
    if i - begin >= target and (c != contig or s >= reach):
      chunks.append((begin, i))
      begin = i
    if c != contig:
      contig, reach = c, e
    else:
      reach = max(reach, e)
  if begin < len(keys):
    chunks.append((begin, len(keys)))
  return chunks


def _convert_chunk(args):
  records, sample, qual_filter, multi_allelic_qual_filter, multiallelic_mode, group_variants, haploid, par, disable_haplotype_resolution = args
  cvos = [parse_cvo(r) for r in records]                      # already in sorted order
  variants = call_variants_outputs_to_variants(cvos, sample, qual_filter, multi_allelic_qual_filter, multiallelic_mode, group_variants, haploid, par)
  if not disable_haplotype_resolution:
    variants = maybe_resolve_conflicting_variants(variants, qual_filter)
  return list(variants)


class _VcfTextWriter:
  """Plain text, or - for *.gz - BGZF with a tabix index beside it, as the reference leaves it (deepvariant_b200/bgzf_tabix.py)."""

  def __init__(self, path: str, header: str):
    from deepvariant_b200 import bgzf_tabix
    self._bgzf = bgzf_tabix.BgzfVcfWriter(path) if path.endswith('.gz') else None
    self._f = None if self._bgzf else open(path, 'w')
    if self._bgzf:
      self._bgzf.write_header(header)
    else:
      self._f.write(header)

  def write(self, line: str, v: 'OutVariant') -> None:
    if self._bgzf:
      self._bgzf.write_record(line, v.reference_name, v.start, v.end)
    else:
      self._f.write(line)

  def __enter__(self):
    return self

  def __exit__(self, *exc):
    if self._bgzf:
      self._bgzf.close()
    else:
      self._f.close()
    return False


def postprocess_variants(infile: str, outfile: str, contigs: Sequence[Tuple[str, int]], sample_name: str = '', qual_filter: float = 1.0,
                         multi_allelic_qual_filter: float = 1.0, multiallelic_mode: str = 'product', only_keep_pass: bool = False,
                         disable_haplotype_resolution: bool = False, group_variants: bool = True, nonvariant_site_tfrecord_path: str = '',
                         gvcf_outfile: str = '', base_at=None, haploid_contigs: str = '', par_regions_bed: str = '', cpus: int = 0,
                         chunk_records: int = 5000) -> dict:
  """CVO TFRecord shards (`infile` may be a sharded spec or a glob) -> VCF text (`outfile`, gzip when it ends in .gz).  With
  nonvariant_site_tfrecord_path (the --gvcf output of make_examples, every shard) and gvcf_outfile also the gVCF: the variants merged
  with the reference blocks (deepvariant_b200/gvcf.py; `base_at(contig, position)` supplies the reference base where a block is split)."""
  import gzip
  if bool(nonvariant_site_tfrecord_path) != bool(gvcf_outfile):
    raise ValueError('gVCF creation requires both nonvariant_site_tfrecord_path and gvcf_outfile')          # postprocess_variants.py:2245-2251
  records = [r for p in tfrecord.resolve_input_paths(infile) for r in tfrecord.read_records(p)]
  if cpus > 1 and len(records) > chunk_records:
    return _postprocess_variants_parallel(records, outfile, contigs, sample_name, qual_filter, multi_allelic_qual_filter, multiallelic_mode, only_keep_pass,
                                          disable_haplotype_resolution, group_variants, nonvariant_site_tfrecord_path, gvcf_outfile, base_at, haploid_contigs,
                                          par_regions_bed, cpus, chunk_records)
  cvos = [parse_cvo(r) for r in records]
  blocks = []
  if gvcf_outfile:
    from deepvariant_b200 import gvcf
    blocks = [gvcf.parse_variant_record(r) for p in tfrecord.resolve_input_paths(nonvariant_site_tfrecord_path) for r in tfrecord.read_records(p)]
  sample = get_sample_name(cvos, sample_name)
  if not sample_name and not cvos and blocks and blocks[0].call_set_name:
    sample = blocks[0].call_set_name                # get_sample_name (postprocess_variants.py:1651-1676): the gVCF records name the sample
  cvos = sort_cvos(cvos, [c for c, _ in contigs])
  haploid = tuple(item for part in (haploid_contigs or '').split(',') for item in part.split())          # is_non_autosome (:1094-1101)
  par = read_bed(par_regions_bed) if par_regions_bed else ()
  variants = call_variants_outputs_to_variants(cvos, sample, qual_filter, multi_allelic_qual_filter, multiallelic_mode, group_variants, haploid, par)
  if not disable_haplotype_resolution:
    variants = maybe_resolve_conflicting_variants(variants, qual_filter)
  return _write_outputs(variants, len(cvos), blocks, outfile, gvcf_outfile, contigs, sample, only_keep_pass, base_at)


def _postprocess_variants_parallel(records, outfile, contigs, sample_name, qual_filter, multi_allelic_qual_filter, multiallelic_mode, only_keep_pass,
                                   disable_haplotype_resolution, group_variants, nonvariant_site_tfrecord_path, gvcf_outfile, base_at, haploid_contigs,
                                   par_regions_bed, cpus, chunk_records) -> dict:
  """The same result with the per-record work (proto parsing, merge_predictions, genotype / GQ / QUAL, haplotype resolution) spread over
  `cpus` worker processes: the parent only extracts the range keys, sorts, cuts independent chunks and writes."""
  import multiprocessing
  keys = [cvo_range_key(r) for r in records]
  pos = {c: i for i, (c, _) in enumerate(contigs)}
  order = sorted(range(len(records)), key=lambda i: (pos.get(keys[i][0], len(pos)), keys[i][1], keys[i][2]))      # stable, like sort_cvos
  records = [records[i] for i in order]
  keys = [keys[i] for i in order]
  blocks = []
  if gvcf_outfile:
    from deepvariant_b200 import gvcf
    blocks = [gvcf.parse_variant_record(r) for p in tfrecord.resolve_input_paths(nonvariant_site_tfrecord_path) for r in tfrecord.read_records(p)]
  sample = get_sample_name([parse_cvo(records[0])] if records else [], sample_name)
  haploid = tuple(item for part in (haploid_contigs or '').split(',') for item in part.split())
  par = read_bed(par_regions_bed) if par_regions_bed else ()
  jobs = [(records[b:e], sample, qual_filter, multi_allelic_qual_filter, multiallelic_mode, group_variants, haploid, par, disable_haplotype_resolution)
          for b, e in independent_chunks(keys, chunk_records)]
  # forkserver: the workers start from a clean single-threaded process (this one may already run BLAS / CUDA / decoder threads, and a
  # fork() of a multi-threaded process can inherit a held lock); the jobs are plain tuples of bytes
  with multiprocessing.get_context('forkserver').Pool(min(cpus, len(jobs))) as pool:
    variants = (v for chunk in pool.imap(_convert_chunk, jobs) for v in chunk)          # in chunk order = genome order
    return _write_outputs(variants, len(records), blocks, outfile, gvcf_outfile, contigs, sample, only_keep_pass, base_at)


def _write_outputs(variants, n_cvos, blocks, outfile, gvcf_outfile, contigs, sample, only_keep_pass, base_at) -> dict:
  n = 0
  if gvcf_outfile:
    from deepvariant_b200 import gvcf
    variants = list(variants)
  header = '\n'.join(vcf_header_lines(contigs, sample)) + '\n'
  with _VcfTextWriter(outfile, header) as w:
    for v in variants:
      if only_keep_pass and v.filter != [PASS]:
        continue
      w.write(vcf_line(v) + '\n', v)
      n += 1
  out = {'n_cvo_records': n_cvos, 'n_variants_written': n, 'sample_name': sample}
  if gvcf_outfile:
    order = {c: i for i, (c, _) in enumerate(contigs)}
    # ShardedVariantReader: the shards are each sorted, merged by (contig index, start)
    blocks.sort(key=lambda b: (order[b.reference_name], b.start, b.end))
    m = 0
    with _VcfTextWriter(gvcf_outfile, header) as w:
      for rec in gvcf.merge_variants_and_nonvariants(variants, blocks, [c for c, _ in contigs], base_at):
        w.write(gvcf.gvcf_line(rec) + '\n', rec)
        m += 1
    out['n_gvcf_records_written'] = m
  return out
