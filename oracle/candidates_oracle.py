"""CPU restatement of the reference's candidate generation - TEST INFRASTRUCTURE ONLY (tests/ and tools/ import it; the
product path is csrc/dvb_candidates.cu and never loads this file).

Keeps the reference's own structure - one AlleleCount per position holding a dict of read key -> allele, walked read by
read - in plain Python, for small cases:
  AlleleCounter.add            deepvariant/allelecounter.cc:880-978 (Add), :402-473 (MakeIndelReadAllele), :475-546 (AddReadAlleles),
                               :195-243 (CanBasesBeUsed, GetAvgBaseQuality), :360-399 (RefBases, GetPrevBase)
  sum_allele_counts / total    deepvariant/allelecounter.cc:78-116, 157-169
  call_variant                 deepvariant/variant_calling_multisample.cc:1117-1233 (CallVariant), :175-196, :203-260, :95-128,
                               :560-606, :608-640, :1235-1290, :1330-1360 for one sample with the make_examples defaults
Pinned by: the KATs of deepvariant/allelecounter_test.cc and variant_calling_test.cc transcribed in tests/test_candidates.py,
and (through the product, which must agree with it on random inputs) by the reference's golden files
(tools/check_candidates_golden.py)."""
import struct

REFERENCE, SUBSTITUTION, INSERTION, DELETION, SOFT_CLIP = 1, 2, 3, 4, 5
_CANON = frozenset(b'ACGT')


def _f32(x):
  return struct.unpack('<f', struct.pack('<f', x))[0]


class Options:
  def __init__(self, **kw):
    self.min_mapping_quality = 5
    self.min_base_quality = 10
    self.keep_legacy_behavior = False
    self.track_ref_reads = False
    self.min_count_snps = 2
    self.min_count_indels = 2
    self.min_fraction_snps = 0.12
    self.min_fraction_indels = 0.06
    self.min_fraction_multiplier = 1.0
    self.vsc_min_indel_fraction_for_small_indels = 0.0
    self.vsc_min_indel_fraction_for_large_indels = 0.0
    self.vsc_small_indel_threshold = 0
    self.small_model_vaf_context_window_size = 0
    self.sample_name = ''
    for k, v in kw.items():
      if not hasattr(self, k):
        raise AttributeError(k)
      setattr(self, k, v)


class AlleleCount:
  def __init__(self, ref_base):
    self.ref_base = ref_base
    self.ref_supporting_read_count = 0
    self.read_alleles = {}      # read key -> dict(bases, type, low_quality, mapq, avg_bq, reverse); insertion-ordered


class AlleleCounter:
  def __init__(self, contig_bases: bytes, start: int, end: int, options: Options, candidate_positions=()):
    self.contig, self.start, self.end, self.o = contig_bases, start, end, options
    self.counts = [AlleleCount(contig_bases[p:p + 1].decode()) for p in range(start, end)]
    self.candidate_positions = sorted(p - start for p in candidate_positions)

  def _ref_bases(self, rel_start, n):
    a = self.start + rel_start
    if a < 0 or a + n > len(self.contig):
      return b''
    return self.contig[a:a + n]

  def _usable(self, read, offset, n):
    """-> (usable, is_low_quality)"""
    q, s = read.aligned_quality, read.aligned_sequence
    total = 0
    for i in range(n):
      total += q[offset + i]
      if q[offset + i] < self.o.min_base_quality and self.o.keep_legacy_behavior:
        return False, False
      if s[offset + i] not in _CANON:
        return False, False
    return True, (not self.o.keep_legacy_behavior) and total < self.o.min_base_quality * n

  def _indel(self, read, interval_offset, read_offset, op, n):
    prev = self._ref_bases(interval_offset - 1, 1) if read_offset == 0 else read.aligned_sequence[read_offset - 1:read_offset]
    if not prev or prev[0] not in _CANON:
      return None
    low = False
    if op != 2:
      ok, low = self._usable(read, read_offset, n)
      if not ok:
        return None
    if op == 2:
      bases = self._ref_bases(interval_offset, n)
      if not bases or any(b not in _CANON for b in bases):
        return None
      typ, avg = DELETION, read.aligned_quality[max(0, read_offset - 1)]
    else:
      bases = read.aligned_sequence[read_offset:read_offset + n]
      typ = INSERTION if op == 1 else SOFT_CLIP
      avg = sum(read.aligned_quality[read_offset:read_offset + n]) // max(1, n)
    return dict(position=interval_offset - 1, bases=(prev + bases).decode(), type=typ, low_quality=low, avg_bq=avg)

  def add(self, read):
    if read.mapping_quality < self.o.min_mapping_quality or not read.aligned_sequence:
      return
    to_add = []
    read_offset, off = 0, read.position - self.start
    n_sites = self.end - self.start
    for op, n in read.cigar:
      if op in (0, 7, 8):
        for i in range(n):
          if 0 <= off + i < n_sites and read_offset + i < len(read.aligned_sequence):
            ok, low = self._usable(read, read_offset + i, 1)
            if ok:
              b = read.aligned_sequence[read_offset + i:read_offset + i + 1]
              to_add.append(dict(position=off + i, bases=b.decode(), low_quality=low, avg_bq=read.aligned_quality[read_offset + i],
                                 type=REFERENCE if b == self.contig[self.start + off + i:self.start + off + i + 1] else SUBSTITUTION))
        read_offset += n
        off += n
      elif op in (1, 4):
        to_add.append(self._indel(read, off, read_offset, op, n) if read_offset + n <= len(read.aligned_sequence) else None)
        read_offset += n
      elif op == 2:
        to_add.append(self._indel(read, off, read_offset, op, n))
        off += n
      elif op in (3, 6):
        off += n
    key = f'{read.fragment_name}/{read.read_number}'
    for i, a in enumerate(to_add):
      if a is None or not 0 <= a['position'] < n_sites:
        continue
      if i + 1 < len(to_add) and to_add[i + 1] is not None and to_add[i + 1]['position'] == a['position']:
        continue
      ac = self.counts[a['position']]
      if a['type'] == REFERENCE and not a['low_quality']:
        ac.ref_supporting_read_count += 1
      if a['type'] != REFERENCE or (self.o.track_ref_reads and a['position'] in self.candidate_positions):
        ac.read_alleles[key] = dict(bases=a['bases'], type=a['type'], low_quality=a['low_quality'], mapq=read.mapping_quality,
                                    avg_bq=a['avg_bq'], reverse=read.reverse_strand)


def sum_allele_counts(ac: AlleleCount):
  sums = {}
  for a in ac.read_alleles.values():
    if not a['low_quality']:
      sums[(a['bases'], a['type'])] = sums.get((a['bases'], a['type']), 0) + 1
  return [dict(bases=k[0], type=k[1], count=v) for k, v in sorted(sums.items())]


def total_allele_counts(ac: AlleleCount):
  return ac.ref_supporting_read_count + sum(1 for a in ac.read_alleles.values() if not a['low_quality'] and a['type'] != REFERENCE)


def _min_fraction(o: Options, a):
  if a['type'] == SUBSTITUTION:
    return _f32(o.min_fraction_snps)
  if o.vsc_small_indel_threshold > 0 and o.vsc_min_indel_fraction_for_small_indels > 0 and o.vsc_min_indel_fraction_for_large_indels > 0:
    return _f32(o.vsc_min_indel_fraction_for_small_indels if len(a['bases']) <= o.vsc_small_indel_threshold + 1
                else o.vsc_min_indel_fraction_for_large_indels)
  return _f32(o.min_fraction_indels)


def _reason(o: Options, a, total, trio):
  if a['type'] == REFERENCE:
    return 'ref'
  if a['count'] < (o.min_count_snps if a['type'] == SUBSTITUTION else o.min_count_indels):
    return 'low_support'
  if a['type'] == SOFT_CLIP:
    return 'other'
  if a['count'] / total < _min_fraction(o, a) * (_f32(o.min_fraction_multiplier) if trio else 1.0):
    return 'low_ratio'
  return None


def select_alt_alleles(o: Options, ac: AlleleCount):
  total = total_allele_counts(ac)
  out = []
  for a in sum_allele_counts(ac):
    r = _reason(o, a, total, False)
    if r is None or (r in ('low_ratio', 'low_support') and _reason(o, a, total, True) is None):
      out.append(a)
  return out


def call_variant(counter: AlleleCounter, index: int, reference_name: str):
  """-> the canonical dict of deepvariant_b200.candidates.canonical_call, or None."""
  o, ac = counter.o, counter.counts[index]
  if ac.ref_base not in 'ACGT' or not ac.ref_base:
    return None
  alts = select_alt_alleles(o, ac)
  if not alts:
    return None
  dels = [a for a in alts if a['type'] == DELETION]
  ref_bases = ac.ref_base
  if dels:
    longest = max(len(a['bases']) for a in dels)
    ref_bases += next(a['bases'] for a in dels if len(a['bases']) == longest)[1:]
  amap = {}
  for a in alts:
    b = a['bases']
    if a['type'] == SUBSTITUTION:
      alt = b if len(b) > 1 and len(ref_bases) > 1 else b + ref_bases[1:]
    elif a['type'] == INSERTION:
      alt = b + ref_bases[1:]
    elif a['type'] == DELETION:
      alt = b[0] + ref_bases[len(b):]
    else:
      continue
    amap[(a['type'], b)] = (alt, a['count'])
  alt_list = sorted(v[0] for v in amap.values())
  assert len(set(alt_list)) == len(alt_list)
  dp = total_allele_counts(ac)
  count_of = {v[0]: v[1] for v in amap.values()}
  start = counter.start + index
  out = {'ref': ref_bases, 'alts': alt_list, 'start': start, 'end': start + len(ref_bases), 'contig': reference_name,
         'info': {'AD': [ac.ref_supporting_read_count] + [count_of[a] for a in alt_list], 'DP': [dp], 'VAF': [count_of[a] / dp for a in alt_list]},
         'call_set_name': o.sample_name, 'genotype': [-1, -1], 'allele_support': {}, 'allele_support_ext': {}, 'ref_support': [],
         'ref_support_ext': [], 'af_at_position': {}}

  def info(key, a):
    return {'read_name': key, 'is_low_quality': int(a['low_quality']), 'mapping_quality': a['mapq'], 'average_base_quality': a['avg_bq'],
            'is_reverse_strand': int(a['reverse']), 'sample_name': o.sample_name}
  for key, a in ac.read_alleles.items():
    if a['type'] != REFERENCE:
      supported = amap.get((a['type'], a['bases']), ('UNCALLED_ALLELE',))[0]
      out['allele_support'].setdefault(supported, []).append(key)
      out['allele_support_ext'].setdefault(supported, []).append(info(key, a))
    else:
      out['ref_support'].append(key)
      out['ref_support_ext'].append(info(key, a))
  for k in out['allele_support']:
    out['allele_support'][k].sort()
    out['allele_support_ext'][k].sort(key=lambda d: d['read_name'])
  out['ref_support'].sort()
  out['ref_support_ext'].sort(key=lambda d: d['read_name'])
  w = o.small_model_vaf_context_window_size
  if w > 0:
    half = w // 2
    for j in range(index - min(index, half), index + min(len(counter.counts) - index, half + 1)):
      c = counter.counts[j]
      depth = c.ref_supporting_read_count + len(c.read_alleles)
      out['af_at_position'][str(counter.start + j)] = (100 * len(c.read_alleles)) // depth if depth > 0 else 0
  return out


def candidates(contig_bases: bytes, reference_name: str, start: int, end: int, reads, options: Options):
  """candidates_in_region for one sample (make_examples_core.py:2832-2960): optional first pass for the tracked positions."""
  positions = ()
  if options.track_ref_reads:
    first = AlleleCounter(contig_bases, start, end, options)
    for r in reads:
      first.add(r)
    positions = [start + i for i, ac in enumerate(first.counts) if ac.ref_base in 'ACGT' and ac.ref_base and select_alt_alleles(options, ac)]
  counter = AlleleCounter(contig_bases, start, end, options, positions)
  for r in reads:
    counter.add(r)
  calls = [call_variant(counter, i, reference_name) for i in range(end - start)]
  return [c for c in calls if c is not None], counter
