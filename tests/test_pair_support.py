"""The (candidate, read) support walk of the encoder's pre-pass (DvbBatch.allele_begin).
CPU: dvb_allele::ElementAt (CIGAR-only walk to one position) == the last commit of the full AlleleCounter walk at that position
(host instantiation of the same header), on random reads with indels, clips, skips, N bases, low qualities and contig ends;
allele keys derived from (ref, alt) invert BuildAlleleMap.  GPU: support classes derived on the device == the host packer's
read-name search over DeepVariantCall.allele_support, and the images are bit-identical."""
import ctypes as C

import numpy as np
import pytest

from deepvariant_b200 import _lib

OPS = {'M': 0, 'I': 1, 'D': 2, 'N': 3, 'S': 4, 'H': 5, 'P': 6, '=': 7, 'X': 8}


def _both(seq, qual, cigar, pos, contig, start, end, target, min_bq=10, legacy=0):
  l = _lib.lib()
  seq_a = np.frombuffer(seq, np.uint8).copy()
  qual_a = np.asarray(qual, np.uint8)
  cig = np.array([(ln << 4) | OPS[op] for op, ln in cigar], np.uint32)
  ctg = np.frombuffer(contig, np.uint8).copy()
  w = np.zeros(6, np.int32)
  a = np.zeros(6, np.int32)
  _lib.check(l.dvb_debug_read_allele_at(seq_a.ctypes.data, qual_a.ctypes.data, len(seq_a), cig.ctypes.data, len(cig), pos, ctg.ctypes.data,
                                        len(ctg), start, end, target, min_bq, legacy, w.ctypes.data, a.ctypes.data))
  return w.tolist(), a.tolist()


def test_known_cases_of_the_target_walk():
  contig = b'ACGTACGTACGTACGTACGTACGTACGTACGT'
  # a substitution at 10 (reference G), then the same base superseded by an insertion anchored on it
  w, a = _both(b'ACGTACTTAC', [30] * 10, [('M', 10)], 4, contig, 0, 32, 10)
  assert w == a == [1, 2, 0, 0, 6, 0]
  w, a = _both(b'ACGTACTGGTAC', [30] * 12, [('M', 7), ('I', 2), ('M', 3)], 4, contig, 0, 32, 10)
  assert w == a == [1, 3, 0, ord('T'), 7, 2]
  # a deletion anchored at 10; a leading soft clip anchored left of the alignment start
  w, a = _both(b'ACGTACGAC', [30] * 9, [('M', 7), ('D', 2), ('M', 2)], 4, contig, 0, 32, 10)
  assert w == a == [1, 4, 0, ord('G'), 7, 2]
  w, a = _both(b'TTACGT', [30] * 6, [('S', 2), ('M', 4)], 11, contig, 0, 32, 10)
  assert w == a == [1, 5, 0, ord('G'), 0, 2]
  # an insertion right after the deletion: anchored on the last deleted base, not on the deletion's anchor
  w, a = _both(b'ACGTACGTTAC', [30] * 11, [('M', 7), ('D', 2), ('I', 2), ('M', 2)], 4, contig, 0, 32, 10)
  assert w == a == [1, 4, 0, ord('G'), 7, 2]
  w, a = _both(b'ACGTACGTTAC', [30] * 11, [('M', 7), ('D', 2), ('I', 2), ('M', 2)], 4, contig, 0, 32, 12)
  assert w == a == [1, 3, 0, ord('G'), 7, 2]
  # an unusable insertion (N inside) leaves the base under it in place
  w, a = _both(b'ACGTACTGNTAC', [30] * 12, [('M', 7), ('I', 2), ('M', 3)], 4, contig, 0, 32, 10)
  assert w == a == [1, 2, 0, 0, 6, 0]
  # nothing there: the read ends before the position / a skip covers it
  w, a = _both(b'ACGT', [30] * 4, [('M', 4)], 4, contig, 0, 32, 10)
  assert w == a == [0] * 6
  w, a = _both(b'ACGTAC', [30] * 6, [('M', 3), ('N', 8), ('M', 3)], 4, contig, 0, 32, 10)
  assert w == a == [0] * 6


@pytest.mark.parametrize('legacy', [0, 1])
def test_target_walk_equals_full_walk_on_random_reads(legacy):
  rng = np.random.default_rng(5 + legacy)
  n_checked = n_found = 0
  kinds = set()
  for trial in range(1500):
    n = int(rng.integers(60, 400))
    contig = rng.choice(np.frombuffer(b'ACGT', np.uint8), n)
    if trial % 3 == 0:
      contig[rng.integers(0, n, 3)] = ord('N')
    contig = contig.tobytes()
    cigar, seq_len, ref_len = [], 0, 0
    for k in range(int(rng.integers(1, 9))):
      op = 'MMMMIDDSN=XP'[int(rng.integers(0, 12))]
      ln = int(rng.integers(1, 12 if op in 'M=X' else 5))
      if cigar and cigar[-1][0] == op:
        continue
      cigar.append((op, ln))
      seq_len += ln if op in 'MIS=X' else 0
      ref_len += ln if op in 'MDN=XP' else 0
    if trial % 11 == 0:
      seq_len = max(1, seq_len - 2)                     # a record whose CIGAR consumes more bases than it has
    pos = int(rng.integers(0, max(1, n - 5)))
    seq = rng.choice(np.frombuffer(b'ACGTACGTACGTN', np.uint8), seq_len)
    ctg = np.frombuffer(contig, np.uint8)
    so, ro = 0, pos
    for op, ln in cigar:                                # aligned bases mostly agree with the reference: REF and SUBSTITUTION both occur
      if op in 'M=X':
        for i in range(ln):
          if so + i < seq_len and ro + i < n and rng.random() < 0.8:
            seq[so + i] = ctg[ro + i]
      so += ln if op in 'MIS=X' else 0
      ro += ln if op in 'MDN=XP' else 0
    qual = rng.choice(np.array([2, 9, 10, 11, 30, 30, 30, 40], np.uint8), seq_len)
    start = int(rng.integers(0, max(1, min(pos + 1, n - 1))))
    end = int(rng.integers(start + 1, n + 1))
    for target in range(start, end):
      if target < pos - 2 or target > pos + ref_len + 2:
        continue
      w, a = _both(seq.tobytes(), qual, cigar, pos, contig, start, end, target, 10, legacy)
      assert w == a, (trial, cigar, pos, start, end, target, w, a)
      n_checked += 1
      n_found += w[0]
      kinds.add(w[1])
  assert n_checked > 15000 and n_found > 5000 and kinds >= {0, 1, 2, 3, 4, 5}, (n_checked, n_found, kinds)


# ---- allele keys and the whole derivation --------------------------------------------------------------------------------------
def test_read_allele_keys_invert_the_allele_map():
  from deepvariant_b200 import packing
  k = packing.read_allele_key
  assert k('A', 'C') == (2, b'C')
  assert k('ATT', 'CTT') == (2, b'C')                 # a SNP beside a 2-bp deletion: the alt carries the reference tail
  assert k('A', 'AGT') == (3, b'AGT')
  assert k('ATT', 'AGTT') == (3, b'AG')               # insertion of G, reference allele widened by the deletion
  assert k('ATT', 'A') == (4, b'ATT') and k('ATT', 'AT') == (4, b'AT')
  assert k('ATT', 'C') == (4, b'CTT')                 # a deletion anchored on a read base that differs from the reference
  assert k('A', 'A') is None and k('AT', 'CA') is None and k('ATT', 'AGTA') is None and k('ATT', 'AG') is None and k('', 'A') is None


def _quickstart():
  import os
  from deepvariant_b200 import bam, fasta
  g = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
  table = bam.NativeBamTable(os.path.join(g, 'quickstart.chr20_10mb.bam'), bam.ReadRequirements(min_mapping_quality=5))
  return table, fasta.IndexedFastaReader(os.path.join(g, 'quickstart.chr20_10mb.fa.gz'))


def _generator(ref, sort_by_support=False):
  from deepvariant_b200 import make_examples_native as men, pileup_image as pi
  pic = pi.default_options()
  pic.channels = list(pi.PILEUP_CHANNELS_WITH_INSERT_SIZE)
  pic.sort_by_alt_allele_support = sort_by_support
  return men.ExamplesGenerator(men.MakeExamplesOptions(pic_options=pic), test_mode=True, ref_reader=ref), pi.to_params(pic)


def _emulate_prepass(packed, ref, contig):
  """The derivation of dvb_pair_prepass_kernel in Python over the host instantiation of ElementAt: per (image, read) pair the
  read allele at variant_start, matched against the image's keys; repeated read keys resolved to the last entry holder."""
  a = packed.arrays
  mq, bq, flags = packed.support
  legacy, track, repeated = bool(flags & 1), bool(flags & 2), bool(flags & 4)
  ctg = np.frombuffer(ref._contig(contig), np.uint8)
  l = _lib.lib()
  support = np.zeros(packed.n_pairs, np.uint8)
  group = np.zeros(packed.n_pairs, np.uint8)
  for i in range(packed.n_images):
    v = int(a['variant_start'][i])
    p0, p1 = int(a['pair_begin'][i]), int(a['pair_begin'][i + 1])
    a0, a1 = int(a['allele_begin'][i]), int(a['allele_begin'][i + 1])
    gdef = int(a['image_group_default'][i]) if 'image_group_default' in a else 0
    raw = []
    for p in range(p0, p1):
      r = int(a['pair_read'][p])
      entry = (False, 0, gdef)
      if a['read_mapq'][r] >= mq:
        s0, s1 = int(a['read_seq_begin'][r]), int(a['read_seq_begin'][r + 1])
        c0, c1 = int(a['read_cigar_begin'][r]), int(a['read_cigar_begin'][r + 1])
        seq, qual, cig = a['bases'][s0:s1].copy(), a['quals'][s0:s1].copy(), a['cigar'][c0:c1].copy()
        w = np.zeros(6, np.int32)
        e = np.zeros(6, np.int32)
        lo, hi = max(0, v - 1000), min(len(ctg), v + 1000)
        _lib.check(l.dvb_debug_read_allele_at(seq.ctypes.data, qual.ctypes.data, len(seq), cig.ctypes.data, len(cig), int(a['read_pos'][r]),
                                              ctg.ctypes.data, len(ctg), lo, hi, v, bq, int(legacy), w.ctypes.data, e.ctypes.data))
        assert w.tolist() == e.tolist()
        # the device knows the reference only as (base at v, canonical run after v): the run the planner fetched must agree
        if e[0] and e[1] == 4:
          assert e[5] <= int(a['image_ref_run'][i])
        if e[0]:
          t = int(e[1])
          if t == 1:
            entry = (track, 0, gdef)
          elif t == 5:
            entry = (True, 0, gdef)
          else:
            key = bytes(seq[e[4]:e[4] + 1]) if t == 2 else bytes([e[3]]) + (bytes(seq[e[4]:e[4] + e[5]]) if t == 3 else
                                                                         bytes(ctg[v + 1:v + 1 + e[5]]))
            entry = (True, 0, gdef)
            for k in range(a0, a1):
              kb = bytes(a['allele_bases'][int(a['allele_bases_begin'][k]):int(a['allele_bases_begin'][k + 1])])
              if int(a['allele_type'][k]) == t and kb == key:
                entry = (True, int(a['allele_class'][k]), int(a['allele_group'][k]) if 'allele_group' in a else 0)
                break
      raw.append(entry)
    for j, p in enumerate(range(p0, p1)):
      q = j
      if repeated:
        rank = a['read_name_rank'][a['pair_read'][p]]
        q = -1
        for k in range(p1 - p0 - 1, -1, -1):
          if raw[k][0] and a['read_name_rank'][a['pair_read'][p0 + k]] == rank:
            q = k
            break
      if q >= 0 and raw[q][0]:
        support[p], group[p] = raw[q][1], raw[q][2]
      else:
        support[p], group[p] = 0, gdef
  return support, group


def _regions_with_candidates(table, ref, opts, step=1000, lo=10_000_000, hi=10_010_000):
  from deepvariant_b200 import candidates as cand
  for p0 in range(lo, hi, step):
    found = cand.candidates_in_region(table, ref, 'chr20', p0, p0 + step, opts)
    if found.records:
      yield ('chr20', p0, p0 + step), found.calls()


@pytest.mark.parametrize('track_ref,legacy,sort_by_support', [(False, False, False), (True, False, True), (False, True, False)])
def test_derivation_from_allele_keys_equals_the_read_name_search(track_ref, legacy, sort_by_support):
  """BASELINE config 1's reads (NA12878 chr20:10,000,000-10,010,000): for every candidate of the very-sensitive caller and
  every read of its pileup, the class derived from (alt-allele keys, the read's own CIGAR walk) equals what the host packer finds
  by searching DeepVariantCall.allele_support for the read's name - soft clips, indels, multi-allelic sites, low-quality bases."""
  from deepvariant_b200 import candidates as cand
  table, ref = _quickstart()
  opts = cand.CandidateOptions(min_mapping_quality=5, min_base_quality=10, keep_legacy_allele_counter_behavior=legacy, track_ref_reads=track_ref)
  gen, _ = _generator(ref, sort_by_support)
  n_pairs = n_images = n_multi = 0
  classes = set()
  for region, calls in _regions_with_candidates(table, ref, opts):
    gen.support_options = None
    plans_n, by_name = gen.pack_region_native(calls, table, region)
    gen.support_options = (5, 10, legacy, track_ref)
    plans_k, by_key = gen.pack_region_native(calls, table, region)
    assert gen.last_region_derived_support and by_key.support is not None and by_name.support is None
    assert [(p.variant.start, p.alt_combination) for p in plans_k] == [(p.variant.start, p.alt_combination) for p in plans_n]
    assert not by_key.arrays['pair_support'][:by_key.n_pairs].any()          # the names were not searched
    for name in ('pair_begin', 'pair_read', 'read_pos', 'variant_start'):
      np.testing.assert_array_equal(by_key.arrays[name], by_name.arrays[name])
    support, group = _emulate_prepass(by_key, ref, 'chr20')
    np.testing.assert_array_equal(support, by_name.arrays['pair_support'][:by_name.n_pairs])
    if sort_by_support:
      np.testing.assert_array_equal(group, by_name.arrays['pair_allele_group'][:by_name.n_pairs])
    n_pairs += by_key.n_pairs
    n_images += by_key.n_images
    n_multi += sum(1 for c in calls if len(c.variant.alternate_bases) > 1)
    classes |= set(support.tolist())
  assert n_images >= 20 and n_pairs > 1000 and classes == {0, 1, 2} and n_multi >= 1


def test_concat_packed_carries_the_allele_keys():
  from deepvariant_b200 import candidates as cand, fused
  table, ref = _quickstart()
  opts = cand.CandidateOptions(min_mapping_quality=5, min_base_quality=10)
  gen, _ = _generator(ref)
  gen.support_options = (5, 10, False, False)
  parts = [gen.pack_region_native(calls, table, region)[1] for region, calls in _regions_with_candidates(table, ref, opts, step=2500)]
  assert len(parts) >= 3
  cat = fused.concat_packed(parts)
  want_s, want_g = zip(*(_emulate_prepass(p, ref, 'chr20') for p in parts))
  got_s, _ = _emulate_prepass(cat, ref, 'chr20')
  np.testing.assert_array_equal(got_s, np.concatenate(want_s))
  b = cat.as_ctypes()
  assert b.n_alleles == sum(int(p.arrays['allele_begin'][p.n_images]) for p in parts) and b.allele_begin and b.image_ref_run
  plain = gen.pack_region_native([], table, ('chr20', 10_000_000, 10_000_100))[1]
  gen.support_options = None
  named = [gen.pack_region_native(calls, table, region)[1] for region, calls in _regions_with_candidates(table, ref, opts, step=5000)]
  with pytest.raises(ValueError):
    fused.concat_packed([parts[0], named[0]])
  del plain


@pytest.mark.gpu
@pytest.mark.parametrize('track_ref,sort_by_support', [(False, False), (True, True)])
def test_device_derives_the_support_classes_of_the_read_name_search(track_ref, sort_by_support):
  """The encoder with allele keys (pair_support derived in the pre-pass kernel) against the same batch with the host packer's
  name-searched classes: identical classes / groups per pair and bit-identical images; also through the chunked upload of
  dvb_encode_classify_host (several phases) by way of a small classifier batch."""
  import ctypes as C
  from deepvariant_b200 import candidates as cand, fused, pileup_image as pi
  table, ref = _quickstart()
  opts = cand.CandidateOptions(min_mapping_quality=5, min_base_quality=10, track_ref_reads=track_ref)
  gen, params = _generator(ref, sort_by_support)
  enc = pi.GpuEncoder(params, device=0)
  by_name, by_key = [], []
  for region, calls in _regions_with_candidates(table, ref, opts):
    gen.support_options = None
    by_name.append(gen.pack_region_native(calls, table, region)[1])
    gen.support_options = (5, 10, False, track_ref)
    by_key.append(gen.pack_region_native(calls, table, region)[1])
  names, keys = fused.concat_packed(by_name), fused.concat_packed(by_key)
  want = enc.encode_host(names)
  got = enc.encode_host(keys)
  sup, grp = enc.last_pair_support(keys.n_pairs)
  np.testing.assert_array_equal(sup, names.arrays['pair_support'][:names.n_pairs])
  if sort_by_support:
    np.testing.assert_array_equal(grp, names.arrays['pair_allele_group'][:names.n_pairs])
  np.testing.assert_array_equal(got, want)
  assert set(sup.tolist()) == {0, 1, 2} and keys.n_images >= 20
  # a single region, and the REPEATED_KEYS resolution forced on (no key repeats here: the resolution must be the identity)
  one = by_key[0]
  one.support = one.support[:2] + (one.support[2] | _lib.SUPPORT_REPEATED_KEYS,)
  np.testing.assert_array_equal(enc.encode_host(one), enc.encode_host(by_name[0]))


def _repeated_keys_case(tmp_path):
  """40x of 60-bp single-end reads over a 3-kb genome whose names come from a pool of 150: many alignments share a key, and the
  two haplotypes (SNPs, a 2-bp insertion, a 3-bp deletion planted on one) mix under one key."""
  import test_bam_native as tb
  from deepvariant_b200 import bam
  rng = np.random.default_rng(21)
  n = 3000
  genome = ''.join(rng.choice(list('ACGT'), n))
  snps, ins, dele = (700, 1300, 1301, 2100), 1700, 2500
  recs = []
  for i in range(2000):
    pos = int(rng.integers(100, n - 200))
    hap = int(rng.integers(0, 2))
    seq, cigar, p, run = [], [], pos, 0
    while len(seq) < 60:
      if p in snps and hap:
        seq.append('ACGT'[('ACGT'.index(genome[p]) + 1 + (p == 1301)) % 4]); run += 1; p += 1
      elif p == ins and hap and run > 0:
        seq += [genome[p], 'G', 'T']; cigar += [(0, run + 1), (1, 2)]; run = 0; p += 1
      elif p == dele and hap and run > 0:
        seq.append(genome[p]); cigar += [(0, run + 1), (2, 3)]; run = 0; p += 4
      else:
        seq.append(genome[p]); run += 1; p += 1
    if run:
      cigar.append((0, run))
    if cigar[-1][0] != 0:
      continue
    merged = []
    for op, k in cigar:
      if merged and merged[-1][0] == op:
        merged[-1] = (op, merged[-1][1] + k)
      else:
        merged.append((op, k))
    recs.append((pos, tb._record(0, pos, f'q{int(rng.integers(0, 150))}', int(rng.choice([3, 20, 60])), 0x10 if i % 3 == 0 else 0, merged, ''.join(seq),
                                 rng.choice([8, 30, 30, 40], len(seq)).tolist())))
  recs.sort(key=lambda t: t[0])
  path = str(tmp_path / 'repeated.bam')
  open(path, 'wb').write(tb._bam([r for _, r in recs], refs=(('chr20', n),)))

  class Ref:
    def n_bases(self, contig): return n
    def is_valid_interval(self, contig, s, e): return 0 <= s <= e <= n
    def query(self, contig, s, e): return genome[s:e]
    def _contig(self, contig): return genome.encode()
    contig_order = ['chr20']
  return bam.NativeBamTable(path, bam.ReadRequirements(min_mapping_quality=5)), Ref()


@pytest.mark.parametrize('track_ref', [False, True])
def test_repeated_read_keys_resolve_like_the_read_name_map(tmp_path, track_ref):
  """AlleleCount.read_alleles is keyed by read name: when several alignments of a region carry one key, the last one that holds
  an entry at the site decides for all of them (and, with track_ref_reads, a reference-matching one can take the entry back)."""
  from deepvariant_b200 import candidates as cand
  table, ref = _repeated_keys_case(tmp_path)
  opts = cand.CandidateOptions(min_mapping_quality=5, min_base_quality=10, track_ref_reads=track_ref)
  gen, _ = _generator(ref)
  region = ('chr20', 0, 3000)
  calls = cand.candidates_in_region(table, ref, 'chr20', 0, 3000, opts).calls()
  assert len(calls) >= 6
  gen.support_options = None
  _, by_name = gen.pack_region_native(calls, table, region)
  gen.support_options = (5, 10, False, track_ref)
  _, by_key = gen.pack_region_native(calls, table, region)
  assert by_key.support[2] & _lib.SUPPORT_REPEATED_KEYS
  support, _ = _emulate_prepass(by_key, ref, 'chr20')
  np.testing.assert_array_equal(support, by_name.arrays['pair_support'][:by_name.n_pairs])
  # the resolution matters here: without it some pairs come out differently
  by_key.support = by_key.support[:2] + (by_key.support[2] & ~_lib.SUPPORT_REPEATED_KEYS,)
  naive, _ = _emulate_prepass(by_key, ref, 'chr20')
  assert (naive != support).any()


@pytest.mark.gpu
@pytest.mark.parametrize('track_ref', [False, True])
def test_device_resolves_repeated_read_keys(tmp_path, track_ref):
  from deepvariant_b200 import candidates as cand, pileup_image as pi
  table, ref = _repeated_keys_case(tmp_path)
  opts = cand.CandidateOptions(min_mapping_quality=5, min_base_quality=10, track_ref_reads=track_ref)
  gen, params = _generator(ref)
  enc = pi.GpuEncoder(params, device=0)
  calls = cand.candidates_in_region(table, ref, 'chr20', 0, 3000, opts).calls()
  gen.support_options = None
  _, by_name = gen.pack_region_native(calls, table, ('chr20', 0, 3000))
  gen.support_options = (5, 10, False, track_ref)
  _, by_key = gen.pack_region_native(calls, table, ('chr20', 0, 3000))
  want = enc.encode_host(by_name)
  got = enc.encode_host(by_key)
  sup, _ = enc.last_pair_support(by_key.n_pairs)
  np.testing.assert_array_equal(sup, by_name.arrays['pair_support'][:by_name.n_pairs])
  np.testing.assert_array_equal(got, want)
