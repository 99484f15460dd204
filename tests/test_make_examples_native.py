"""Host driver KATs transcribed from deepvariant/make_examples_native_test.cc and
deepvariant/alt_aligned_pileup_lib_test.cc, plus the region driver end to end."""
import dataclasses
import json
import os

import numpy as np
import pytest

import oracle_lib
from deepvariant_b200 import make_examples_native as men
from deepvariant_b200 import packing, protos, tfrecord
from deepvariant_b200 import pileup_image as pi
from deepvariant_b200.protos import DeepVariantCall, Read, Variant, parse_cigar_string


def _variant(ref, alts, start=10):
  return Variant(reference_name='chr1', start=start, end=start + len(ref), reference_bases=ref, alternate_bases=list(alts))


@pytest.mark.parametrize('mode,ref,alts,indices,expected', [
    ('ADD_HET_ALT_IMAGES', 'A', ['T'], [], [['T']]),
    ('ADD_HET_ALT_IMAGES', 'AT', ['A'], [], [['A']]),
    ('ADD_HET_ALT_IMAGES', 'A', ['ATT'], [], [['ATT']]),
    ('ADD_HET_ALT_IMAGES', 'AT', ['A', 'ATT'], [], [['A'], ['ATT'], ['A', 'ATT']]),
    ('NO_HET_ALT_IMAGES', 'AT', ['A', 'ATT'], [], [['A'], ['ATT']]),
    ('ADD_HET_ALT_IMAGES', 'AT', ['A', 'ATT', 'ATTG'], [[0], [0, 1]], [['A'], ['A', 'ATT']]),
    ('NO_HET_ALT_IMAGES', 'AT', ['A', 'ATT', 'ATTG'], [[0], [1], [0, 1]], [['A'], ['ATT']]),
    ('ADD_HET_ALT_IMAGES', 'AT', ['A', 'ATT', 'ATTG'], [[0, 2], [1, 2]], [['A', 'ATTG'], ['ATT', 'ATTG']]),
])
def test_alt_allele_combinations(mode, ref, alts, indices, expected):
  """make_examples_native_test.cc:489-544 (unordered)."""
  call = DeepVariantCall(variant=_variant(ref, alts), make_examples_alt_allele_indices=indices)
  got = men.alt_allele_combinations(call, mode)
  assert sorted(map(tuple, got)) == sorted(map(tuple, expected))


def test_alt_allele_combinations_unspecified_mode_is_fatal():
  with pytest.raises(ValueError):
    men.alt_allele_combinations(DeepVariantCall(variant=_variant('A', ['T'])), 'UNSPECIFIED')


class _Ref:
  """In-memory reference like nucleus InMemoryFastaReader."""

  def __init__(self, contigs):
    self.c = contigs

  def n_bases(self, name):
    return len(self.c[name])

  def is_valid_interval(self, name, s, e):
    return name in self.c and 0 <= s <= e <= len(self.c[name])

  def query(self, name, s, e):
    return self.c[name][s:e]


@pytest.mark.parametrize('start,expected', [(10, 'AGTGGGGGGGGGATGGGGGTG'), (5, 'NNNNNAGTGGGGGGGGGATGG'), (16, 'GGGGGGATGGGGGTGNNNNNN')])
def test_get_reference_bases_for_pileup(start, expected):
  """make_examples_native_test.cc:792-838."""
  pic = dataclasses.replace(pi.default_options(), width=21, channels=list(pi.PILEUP_DEFAULT_CHANNELS))
  gen = men.ExamplesGenerator(men.MakeExamplesOptions(pic_options=pic), test_mode=True,
                              ref_reader=_Ref({'chr1': 'AGTGGGGGGGGGATGGGGGTG'}))
  assert gen.get_reference_bases_for_pileup(_variant('A', ['T'], start)) == expected


@pytest.mark.parametrize('ref_start,ref_len,cigar,expected,read_start,read_len', [
    (10, 20, '20M5I10M', '10M5I10M', 10, 25), (10, 20, '20M5D10M', '10M5D5M', 10, 15), (22, 10, '20M5I20M', '10M', 27, 10),
    (22, 10, '20M5D20M', '3D7M', 20, 7), (50, 20, '20M5I10M', '', 35, 0), (10, 40, '20M5I10M', '10M5I10M', 10, 25)])
def test_trim_cigar(ref_start, ref_len, cigar, expected, read_start, read_len):
  """alt_aligned_pileup_lib_test.cc:146-162."""
  got, rs, rl = men.trim_cigar(parse_cigar_string(cigar), ref_start, ref_len)
  assert got == parse_cigar_string(expected) and (rs, rl) == (read_start, read_len)


SEQ22 = 'ACGTACGTAAAAAAGTGTGATC'


@pytest.mark.parametrize('read_start,trim_start,trim_len,cigar,exp_start,exp_bases,exp_cigar,exp_quals', [
    (10, 15, 5, '22M', 15, 'CGTAA', '5M', [6, 7, 8, 9, 10]),
    (10, 15, 5, '2M3I17M', 15, 'AAAAA', '5M', [9, 10, 11, 12, 13]),
    (10, 15, 5, '2M3D20M', 15, 'GTACG', '5M', [3, 4, 5, 6, 7]),
    (10, 8, 5, '22M', 10, 'ACG', '3M', [1, 2, 3]),
    (10, 10, 22, '22M', 10, SEQ22, '22M', list(range(1, 23)))])
def test_trim_read(read_start, trim_start, trim_len, cigar, exp_start, exp_bases, exp_cigar, exp_quals):
  """alt_aligned_pileup_lib_test.cc:196-252."""
  r = Read(fragment_name='r', reference_name='chr1', position=read_start, cigar=parse_cigar_string(cigar),
           aligned_sequence=SEQ22.encode(), aligned_quality=bytes(range(1, 23)))
  t = men.trim_read(r, trim_start, trim_start + trim_len)
  assert (t.position, t.aligned_sequence.decode(), t.cigar, list(t.aligned_quality)) == (
      exp_start, exp_bases, parse_cigar_string(exp_cigar), exp_quals)
  assert r.position == read_start and len(r.aligned_sequence) == 22   # input untouched


def test_trim_reads_min_overlap_and_original_positions():
  reads = [Read(fragment_name=f'r{i}', reference_name='chr1', position=p, cigar=[(0, 30)], aligned_sequence=b'A' * 30,
                aligned_quality=bytes([30] * 30)) for i, p in enumerate((0, 60, 95))]
  out, orig = men.trim_reads(reads, 20, 100, 15)
  assert [r.fragment_name for r in out] == ['r1']     # r0 keeps 10 bp (<15), r2 keeps 5 bp
  assert orig == [60] and out[0].position == 60


def test_encoded_variant_type_and_alt_indices():
  assert men.encoded_variant_type(_variant('A', ['T'])) == 1
  assert men.encoded_variant_type(_variant('A', ['T', 'G'])) == 1
  assert men.encoded_variant_type(_variant('AT', ['A'])) == 2
  assert men.encoded_variant_type(_variant('A', ['ATT', 'C'])) == 2
  assert men.encoded_variant_type(_variant('A', [])) == 0
  enc, idx = men.encode_alt_alleles(_variant('AT', ['A', 'ATT', 'ATTG']), ['A', 'ATTG'])
  assert idx == [0, 2] and protos.parse_alt_allele_indices(enc) == [0, 2]


def _region_fixture():
  rng = np.random.default_rng(5)
  ref_seq = ''.join(rng.choice(list('ACGT'), 3000))
  ref = _Ref({'chr1': ref_seq})
  reads = []
  for i in range(120):
    pos = int(rng.integers(900, 1500))
    seq = ref_seq[pos:pos + 100]
    reads.append(Read(fragment_name=f'frag{i // 2}', read_number=i % 2, reference_name='chr1', position=pos,
                      reverse_strand=bool(i % 3 == 0), mapping_quality=int(rng.integers(3, 61)), cigar=[(0, 100)],
                      aligned_sequence=seq.encode(), aligned_quality=bytes(rng.integers(5, 41, 100).tolist()),
                      fragment_length=int(rng.integers(-600, 600))))
  names = [r.key() for r in reads]
  cands = [
      DeepVariantCall(variant=_variant(ref_seq[1200], ['A' if ref_seq[1200] != 'A' else 'C'], 1200),
                      allele_support={('A' if ref_seq[1200] != 'A' else 'C'): names[:20]}),
      DeepVariantCall(variant=_variant(ref_seq[1300:1302], [ref_seq[1300], ref_seq[1300:1302] + 'TT'], 1300),
                      allele_support={ref_seq[1300]: names[20:30], ref_seq[1300:1302] + 'TT': names[30:44]}),
      DeepVariantCall(variant=_variant(ref_seq[40], ['T' if ref_seq[40] != 'T' else 'G'], 40)),   # N-padded window, no reads
  ]
  pic = dataclasses.replace(pi.default_options(), channels=list(pi.PILEUP_CHANNELS_WITH_INSERT_SIZE), num_channels=7)
  return ref, reads, cands, pic


def test_plan_and_finish_region_with_oracle_pixels(tmp_path):
  """Host logic of CreateAndWriteExamplesForCandidate: 1 + 3 + 1 examples, features as the reference writes them."""
  ref, reads, cands, pic = _region_fixture()
  gen = men.ExamplesGenerator(men.MakeExamplesOptions(pic_options=pic), test_mode=True, ref_reader=ref)
  stats = {}
  plans = gen.plan_region(cands, reads, stats)
  assert [p.alt_combination for p in plans[1:4]] == [[cands[1].variant.alternate_bases[0]], [cands[1].variant.alternate_bases[1]],
                                                     cands[1].variant.alternate_bases]
  assert len(plans) == 5
  # read query: [start-5, end+5) overlap (make_examples_native.cc:643-648)
  for p in plans:
    v = p.variant
    want = [r for r in reads if r.position < v.end + 5 and r.end() > v.start - 5]
    assert [id(r) for r in p.spec.reads] == [id(r) for r in want]
  assert plans[4].spec.ref_bases.startswith('N' * 70) and not plans[4].spec.reads
  # support classes: het-alt image counts both alts as "this image", single-alt images see the other as class 2
  sup1 = dict(zip([r.key() for r in plans[1].spec.reads], plans[1].spec.support))
  sup3 = dict(zip([r.key() for r in plans[3].spec.reads], plans[3].spec.support))
  other = set(cands[1].allele_support[cands[1].variant.alternate_bases[1]])
  assert all(sup1[k] == 2 for k in sup1 if k in other) and all(sup3[k] == 1 for k in sup3 if k in other)
  params = pi.to_params(pic)
  images = oracle_lib.encode_batch(params, packing.pack_images([p.spec for p in plans], params))
  recs = gen.finish_region(plans, images, stats)
  assert stats == {'n_examples': 5, 'n_snps': 2, 'n_indels': 3}
  ex = protos.parse_tf_example(recs[2])
  assert ex['locus'][1][0] == b'chr1:1301-1302' and ex['variant_type'][1] == [2] and ex['image/shape'][1] == [100, 221, 7]
  assert protos.parse_alt_allele_indices(ex['alt_allele_indices/encoded'][1][0]) == [1]
  assert protos.parse_variant(ex['variant/encoded'][1][0]).alternate_bases == cands[1].variant.alternate_bases
  np.testing.assert_array_equal(np.frombuffer(ex['image/encoded'][1][0], np.uint8).reshape(100, 221, 7), images[2])


def test_trim_reads_for_pileup_plans_trimmed_reads_with_original_positions():
  ref, reads, cands, pic = _region_fixture()
  pic = dataclasses.replace(pic, width=147)
  gen = men.ExamplesGenerator(men.MakeExamplesOptions(pic_options=pic, trim_reads_for_pileup=True), test_mode=True, ref_reader=ref)
  plans = gen.plan_region(cands[:1], reads, {})
  spec = plans[0].spec
  lo, hi = 1200 - 73, 1201 + 73
  assert all(lo <= r.position and r.end() <= hi for r in spec.reads)
  assert any(sp != r.position for sp, r in zip(spec.sort_positions, spec.reads))   # sorted by pre-trim position


@pytest.mark.gpu
def test_write_examples_in_region_on_gpu_matches_oracle(tmp_path):
  ref, reads, cands, pic = _region_fixture()
  path = str(tmp_path / 'examples.tfrecord.gz')
  gen = men.ExamplesGenerator(men.MakeExamplesOptions(pic_options=pic), {'main_sample': path}, ref_reader=ref)
  stats, shape = gen.write_examples_in_region(cands, [reads], [0], 'main_sample', [0.0])
  gen.signal_shard_finished()
  assert stats == {'n_examples': 5, 'n_snps': 2, 'n_indels': 3} and shape == [100, 221, 7]
  recs = list(tfrecord.read_records(path, check_crc=True))
  assert len(recs) == 5
  info = json.load(open(path + '.example_info.json'))
  assert info == {'version': '1.10.0', 'shape': [100, 221, 7], 'channels': [1, 2, 3, 4, 5, 6, 19]}
  gen2 = men.ExamplesGenerator(men.MakeExamplesOptions(pic_options=pic), test_mode=True, ref_reader=ref)
  plans = gen2.plan_region(cands, reads, {})
  params = pi.to_params(pic)
  want = oracle_lib.encode_batch(params, packing.pack_images([p.spec for p in plans], params))
  for i, rec in enumerate(recs):
    img = np.frombuffer(protos.parse_tf_example(rec)['image/encoded'][1][0], np.uint8).reshape(100, 221, 7)
    np.testing.assert_array_equal(img, want[i])
  with pytest.raises(KeyError):
    gen.write_examples_in_region(cands, [reads], [0], 'no_such_role', [0.0])
