"""gVCF reference blocks: `make_examples --gvcf` (the non-variant site records) and the merge of those records with the called
variants that `postprocess_variants --nonvariant_site_tfrecord_path --gvcf_outfile` writes.

make_examples side restates VariantCaller.reference_confidence / _calc_reference_confidence / make_gvcfs
(deepvariant/variant_caller.py:154-254, 256-413) with _rescale_read_counts_if_necessary / _quantize_gq (:76-122) and the
nucleus helpers normalize_log10_probs / log10sumexp (third_party/nucleus/util/genomics_math.py:183-262) and Log10PTrueToPhred
(third_party/nucleus/util/math.cc:78-83), over AlleleCounter::SummaryCounts (deepvariant/allelecounter.cc:986-1007), which
dvb_candidates_summary_counts returns from the same counter the candidates came from (calls_and_gvcfs, variant_caller.py:415-468).
Options are make_examples' (make_examples_core.py:225-236): p_error 0.001, max_gq 50, ploidy 2, gq_resolution = --gvcf_gq_binsize
(5); VerySensitiveCaller caches the confidences up to a coverage of 100 and RESCALES deeper sites onto the table
(very_sensitive_caller.py:46, variant_caller.py:211-218) - restated, since it changes the numbers.

postprocess side restates nucleus MergeAndWriteVariantsAndNonVariants with CreateRecordFromTemplate / TransfromToGvcf /
ZeroScaleGl (third_party/nucleus/io/merge_variants.cc:52-101, 162-231).

Pinned by golden.postprocess_gvcf_input.tfrecord.gz (make_examples side, from BAM + FASTA) and golden.postprocess_gvcf_output*.g.vcf
(postprocess side): tools/check_gvcf_golden.py, tests/test_gvcf.py."""
from __future__ import annotations

import dataclasses
import itertools
import math
import statistics
import struct
from typing import Callable, Dict, Iterable, Iterator, List, Optional, Sequence, Tuple

import numpy as np

from deepvariant_b200 import candidates as cand
from deepvariant_b200 import protos
from deepvariant_b200.postprocess_variants import OutVariant

GVCF_ALT_ALLELE = '<*>'
GVCF_ALT_ALLELE_GL = -99.0                 # merge_variants.cc:48
CANONICAL_DNA_BASES = frozenset('ACGT')
EXTENDED_IUPAC_CODES = frozenset('ACGTRYSWKMBDHVN')
LOG_10 = math.log(10.0)
IMPOSSIBLE_PROBABILITY_LOG10 = 999.0


@dataclasses.dataclass
class GvcfOptions:
  sample_name: str = ''
  p_error: float = 0.001
  max_gq: int = 50
  gq_resolution: int = 5
  ploidy: int = 2
  haploid_contigs: Tuple[str, ...] = ()
  par_regions: Tuple[Tuple[str, int, int], ...] = ()
  include_med_dp: bool = False
  max_cache_coverage: int = 100            # very_sensitive_caller.py:46; 0 = no table, exact at every depth


def _log10sumexp(xs: Sequence[float]) -> float:
  m = max(xs)
  return m + math.log10(sum(pow(10.0, x - m) for x in xs))


def _normalize_log10_probs(xs: Sequence[float]) -> List[float]:
  if max(xs) > 0.0:
    raise ValueError('log10_probs all must be <= 0', xs)
  lse = _log10sumexp(xs)
  return [min(x - lse, 0.0) for x in xs]


def _log10_ptrue_to_phred(log10_ptrue: float, value_if_not_finite: float) -> float:
  ptrue = math.pow(10.0, log10_ptrue)
  if 1.0 - ptrue <= 0.0:
    return value_if_not_finite
  return -10.0 * math.log10(1.0 - ptrue)


def rescale_read_counts_if_necessary(n_ref: int, n_total: int, max_allowed: int) -> Tuple[int, int]:
  if n_total > max_allowed:
    ratio = n_ref / (1.0 * n_total)
    n_ref = int(math.ceil(ratio * max_allowed))
    n_total = max_allowed
  return n_ref, n_total


def quantize_gq(raw_gq: int, binsize: int) -> int:
  if raw_gq < 1:
    return 0
  return ((raw_gq - 1) // binsize) * binsize + 1


class ReferenceConfidence:
  """reference_confidence(n_ref, n_total, is_haploid) -> (raw GQ, [log10 p(0/0), p(0/<*>), p(<*>/<*>)])."""

  def __init__(self, options: GvcfOptions):
    self.o = options
    self.p_error = float(np.float32(options.p_error))      # VariantCallerOptions.p_error is a proto `float` (deepvariant.proto)
    self._cache: Dict[Tuple[int, int, bool], Tuple[int, List[float]]] = {}

  def calc(self, n_ref: int, n_total: int, is_haploid: bool = False) -> Tuple[int, List[float]]:
    if n_ref < 0:
      raise ValueError(f'n_ref={n_ref} must be >= 0')
    if n_total < n_ref:
      raise ValueError(f'n_total={n_total} must be >= n_ref={n_ref}')
    if self.o.ploidy != 2:
      raise ValueError(f'ploidy={self.o.ploidy} but we only support ploidy=2')
    if n_total == 0:
      probs = _normalize_log10_probs([-1.0, -IMPOSSIBLE_PROBABILITY_LOG10, -1.0] if is_haploid else [-1.0, -1.0, -1.0])
    else:
      n_alts = n_total - n_ref
      logp = math.log(self.p_error) / LOG_10
      log1p = math.log1p(-self.p_error) / LOG_10
      p_ref = n_ref * log1p + n_alts * logp
      p_het = -IMPOSSIBLE_PROBABILITY_LOG10 if is_haploid else -n_total * math.log(self.o.ploidy) / LOG_10
      p_hom_alt = n_ref * logp + n_alts * log1p
      probs = _normalize_log10_probs([p_ref, p_het, p_hom_alt])
    gq = _log10_ptrue_to_phred(probs[0], self.o.max_gq)
    return int(min(math.floor(gq), self.o.max_gq)), probs

  def __call__(self, n_ref: int, n_total: int, is_haploid: bool = False) -> Tuple[int, List[float]]:
    if self.o.max_cache_coverage > 0:
      n_ref, n_total = rescale_read_counts_if_necessary(n_ref, n_total, self.o.max_cache_coverage)
    key = (n_ref, n_total, is_haploid)
    hit = self._cache.get(key)
    if hit is None:
      hit = self._cache[key] = self.calc(n_ref, n_total, is_haploid)
    return hit


def make_gvcfs(contig: str, start: int, ref_bases: str, summary_counts: np.ndarray, options: GvcfOptions,
               confidence: Optional[ReferenceConfidence] = None) -> Iterator[OutVariant]:
  """summary_counts[i] = (ref_supporting_read_count, total_read_count) of position start + i whose reference base is ref_bases[i]."""
  confidence = confidence or ReferenceConfidence(options)
  haploid = contig in options.haploid_contigs

  def site(i: int):
    base = ref_bases[i]
    n_ref, n_total = int(summary_counts[i][0]), int(summary_counts[i][1])
    if base not in CANONICAL_DNA_BASES:
      if base not in EXTENDED_IUPAC_CODES:
        raise ValueError(f'Invalid reference base={base} found during gvcf calculation')
      return (None, True, None, None, n_total, i)
    is_haploid = haploid and not any(c == contig and s <= start + i < e for c, s, e in options.par_regions)
    raw_gq, probs = confidence(n_ref, n_total, is_haploid)
    return (quantize_gq(raw_gq, options.gq_resolution), max(probs) == probs[0], raw_gq, probs, n_total, i)

  def record(first: int, last: int, genotype, probs, gq: int, min_dp: int, med_dp: int) -> OutVariant:
    info = {'MIN_DP': [min_dp]}
    if options.include_med_dp:
      info['MED_DP'] = [med_dp]
    return OutVariant(contig, start + first, start + last + 1, ref_bases[first], [GVCF_ALT_ALLELE], info, call_set_name=options.sample_name,
                      genotype=list(genotype), genotype_likelihood=list(probs), gq=gq)

  for (quantized, valid), group in itertools.groupby((site(i) for i in range(len(summary_counts))), key=lambda t: (t[0], t[1])):
    if quantized is None:
      continue
    group = list(group)
    if valid:
      k = min(range(len(group)), key=lambda j: group[j][2])            # the first record with the smallest raw GQ
      depths = [g[4] for g in group]
      yield record(group[0][5], group[-1][5], (0, 0), group[k][3], group[k][2], min(depths), int(statistics.median(depths)))
    else:
      for g in group:
        yield record(g[5], g[5], (-1, -1), g[3], g[2], g[4], g[4])


# ---- Variant proto <-> OutVariant (variants.proto: Variant 6, 7, 11, 13, 14, 16; VariantCall 2, 6, 7, 9) ----------------------------
def _info_entry(key: str, ints: Sequence[int]) -> bytes:
  values = b''.join(protos.f_bytes(1, protos.f_varint(7, int(v))) for v in ints)          # ListValue.values -> Value.int_value
  return protos.f_bytes(2, protos.f_bytes(1, key.encode()) + protos.f_bytes(2, values))


def serialize_gvcf_record(v: OutVariant) -> bytes:
  call = b''.join(_info_entry(k, v.info[k]) for k in sorted(v.info))
  if v.gq is not None:
    call = _info_entry('GQ', [v.gq]) + call
  call += protos.f_bytes(6, b''.join(struct.pack('<d', float(x)) for x in v.genotype_likelihood))
  call += protos.f_bytes(7, protos.packed_varints([g & 0xFFFFFFFFFFFFFFFF for g in v.genotype]))
  call += protos.f_bytes(9, v.call_set_name.encode())
  out = protos.f_bytes(6, v.reference_bases.encode()) + b''.join(protos.f_bytes(7, a.encode()) for a in v.alternate_bases)
  out += protos.f_bytes(11, call) + protos.f_varint(13, v.end) + protos.f_bytes(14, v.reference_name.encode()) + protos.f_varint(16, v.start)
  return out


def parse_variant_record(record: bytes) -> OutVariant:
  """A serialized nucleus Variant with one call (a gVCF block of make_examples, or any Variant) -> OutVariant."""
  c = cand.canonical_call(protos.f_bytes(1, record))
  likelihoods: List[float] = []
  quality, filters = 0.0, []
  for fn, wt, val, _ in protos.iter_fields(record):
    if fn == 11:
      for f2, w2, v2, _ in protos.iter_fields(bytes(val)):
        if f2 == 6:
          raw = bytes(v2) if w2 == 2 else struct.pack('<Q', v2)
          likelihoods += [struct.unpack('<d', raw[i:i + 8])[0] for i in range(0, len(raw), 8)]
    elif fn == 8:
      quality = struct.unpack('<d', struct.pack('<Q', val))[0] if isinstance(val, int) else struct.unpack('<d', bytes(val))[0]
    elif fn == 9:
      filters.append(bytes(val).decode())
  info = {k: list(vs) for k, vs in c['info'].items()}
  gq = info.pop('GQ', [None])[0]
  return OutVariant(c['contig'], c['start'], c['end'], c['ref'], list(c['alts']), info, call_set_name=c['call_set_name'],
                    genotype=list(c['genotype']), genotype_likelihood=likelihoods, gq=gq, quality=quality, filter=filters)


# ---- postprocess side: merge variants and non-variant blocks -------------------------------------------------------------------------
def _from_template(t: OutVariant, start: int, end: int, base_at: Callable[[str, int], str]) -> OutVariant:
  v = dataclasses.replace(t, start=start, end=end, info={k: list(x) for k, x in t.info.items()},
                          genotype_likelihood=list(t.genotype_likelihood), genotype=list(t.genotype))
  if start != t.start:
    v.reference_bases = base_at(t.reference_name, start)
  return v


def transform_to_gvcf(v: OutVariant) -> OutVariant:
  """ZeroScaleGl + TransfromToGvcf: the variant as the gVCF prints it - likelihoods shifted to max 0, `<*>` appended with
  likelihood -99 for every genotype that contains it, AD 0 and VAF 0 for it."""
  out = dataclasses.replace(v, alternate_bases=list(v.alternate_bases), info={k: list(x) for k, x in v.info.items()},
                            genotype_likelihood=list(v.genotype_likelihood))
  if out.genotype_likelihood:
    m = max(out.genotype_likelihood)
    out.genotype_likelihood = [x - m for x in out.genotype_likelihood]
  if GVCF_ALT_ALLELE not in out.alternate_bases:
    out.alternate_bases.append(GVCF_ALT_ALLELE)
    out.genotype_likelihood += [GVCF_ALT_ALLELE_GL] * (len(out.alternate_bases) + 1)
    if 'AD' in out.info:
      out.info['AD'].append(0)
    if 'VAF' in out.info:
      out.info['VAF'].append(0.0)
  return out


def merge_variants_and_nonvariants(variants: Iterable[OutVariant], nonvariants: Iterable[OutVariant], contig_order: Sequence[str],
                                   base_at: Callable[[str, int], str]) -> Iterator[OutVariant]:
  """Yields the records of the gVCF in order (both inputs sorted by contig order, then position)."""
  index = {c: i for i, c in enumerate(contig_order)}
  inf = len(index) + 1
  vi, ni = iter(variants), iter(nonvariants)
  v, n = next(vi, None), next(ni, None)
  while v is not None or n is not None:
    vc = index[v.reference_name] if v is not None else inf
    nc = index[n.reference_name] if n is not None else inf
    if vc < nc or (vc == nc and v.end <= n.start):
      yield transform_to_gvcf(v)
      v = next(vi, None)
    elif nc < vc or (nc == vc and n.end <= v.start):
      yield n
      n = next(ni, None)
    else:
      if n.start < v.start:
        yield _from_template(n, n.start, v.start, base_at)
      if n.end > v.end:
        n = _from_template(n, v.end, n.end, base_at)
      else:
        n = next(ni, None)


def gvcf_line(v: OutVariant) -> str:
  """One record as nucleus VcfWriter prints it (END for a single symbolic ALT, vcf_conversion.cc:1072-1084)."""
  from deepvariant_b200.postprocess_variants import _fmt_float   # pylint: disable=g-import-not-at-top
  qual = math.floor(v.quality * 10 + 0.5) / 10
  info = f'END={v.end}' if len(v.alternate_bases) == 1 and v.alternate_bases[0].startswith('<') else '.'
  keys, vals = ['GT'], [('|' if v.is_phased else '/').join('.' if g < 0 else str(g) for g in v.genotype)]
  if v.gq is not None:
    keys.append('GQ')
    vals.append(str(int(v.gq)))
  for k in ('DP', 'MIN_DP', 'MED_DP', 'AD'):
    if k in v.info:
      keys.append(k)
      vals.append(','.join(str(int(x)) for x in v.info[k]))
  if 'VAF' in v.info:
    keys.append('VAF')
    vals.append(','.join(_fmt_float(x) for x in v.info['VAF']))
  if v.genotype_likelihood:
    m = max(v.genotype_likelihood)
    keys.append('PL')
    vals.append(','.join(str(int(-10 * (x - m))) for x in v.genotype_likelihood))
  return '\t'.join([v.reference_name, str(v.start + 1), '.', v.reference_bases, ','.join(v.alternate_bases) or '.', _fmt_float(qual),
                    ';'.join(v.filter) or '.', info, ':'.join(keys), ':'.join(vals)])
