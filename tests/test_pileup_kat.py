"""Known-answer tests transcribed (as data) from the reference's own tests of this path:

  deepvariant/pileup_image_test.py:138-660          (encode_reference / encode_read KATs)
  deepvariant/pileup_image_native_test.cc:277-413   (BuildPileupForOneSample, 4 cases)
  deepvariant/pileup_image_native_test.cc:872-969   (sort order)
  deepvariant/pileup_channel_lib_test.cc:566-633    (insert size, supplementary)
  deepvariant/pileup_channel_lib_test.cc:851-951    (ref-row values)

Every test runs twice: backend 'oracle' pins the CPU oracle (runs here, `-m "not gpu"`),
backend 'gpu' pins the CUDA path through the C ABI (`-m gpu`, on the B200).
"""
import dataclasses
import itertools

import numpy as np
import pytest

from deepvariant_b200 import pileup_image as pi
from deepvariant_b200.protos import DeepVariantCall, Read, Variant, parse_cigar_string

_counter = itertools.count()


def make_read(bases, start, quals=None, cigar=None, mapq=50, chrom='chr1', name=None,
              fragment_length=None):
  """third_party/nucleus/testing/test_utils.py:288-316."""
  if quals is not None and len(bases) != len(list(quals)):
    raise ValueError('Incompatable bases and quals')
  return Read(
      fragment_name=name if name else 'read_' + str(next(_counter)),
      proper_placement=True, read_number=1, number_reads=2,
      aligned_sequence=bases.encode(),
      aligned_quality=bytes(list(quals)) if quals is not None else b'',
      fragment_length=fragment_length or 0,
      reference_name=chrom, position=start, mapping_quality=mapq,
      cigar=parse_cigar_string(cigar) if cigar else [])


def cc_make_read(chrom, start, bases, cigar_elements, name='', hp_tag=-1):
  """third_party/nucleus/testing/test_utils.cc:128-152 + deepvariant/testing_utils.cc:75-88:
  every base quality 30, mapq 90, read_number 0."""
  r = Read(fragment_name=name, aligned_sequence=bases.encode(), number_reads=2, proper_placement=True,
           aligned_quality=bytes([30] * len(bases)), mapping_quality=90, reference_name=chrom,
           position=start, cigar=parse_cigar_string(''.join(cigar_elements)))
  if hp_tag >= 0:
    r.hp_values = [hp_tag]
  return r


def _make_dv_call(ref_bases='A', alt_bases='C'):
  return DeepVariantCall(
      variant=Variant(reference_name='chr1', start=10, end=11, reference_bases=ref_bases,
                      alternate_bases=[alt_bases]),
      allele_support={'C': ['read1/1', 'read2/1']})


def _options(channels=None, read_requirements=None, **kwargs):
  o = pi.default_options(read_requirements)
  o.channels = list(channels if channels is not None else pi.PILEUP_DEFAULT_CHANNELS)
  o.num_channels = len(o.channels)
  return dataclasses.replace(o, **kwargs)


def cc_default_options(width, height, ref_band_height, channels):
  """deepvariant/testing_utils.cc:139-165 MakeDefaultPileupImageOptions (no read requirements)."""
  o = pi.default_options(pi.ReadRequirements(0, 0))
  return dataclasses.replace(o, width=width, height=height, reference_band_height=ref_band_height,
                             channels=list(channels), num_channels=len(channels),
                             multi_allelic_mode='UNSPECIFIED', alt_aligned_pileup='', types_to_alt_align='')


def dstack(*rows):
  return np.dstack([np.array(r) for r in rows]).astype(np.uint8)


FULL_EXPECTED = dstack((250, 30, 30, 180, 100), (63, 69, 76, 82, 88), (211,) * 5, (70,) * 5,
                       (254,) * 5, (50, 50, 254, 50, 50))


def test_reference_encoding(backend):
  got = backend(_options()).encode_reference('ACGTN')
  np.testing.assert_array_equal(
      got, dstack((250, 30, 180, 100, 0), (254,) * 5, (254,) * 5, (70,) * 5, (152,) * 5, (50,) * 5))


def test_encode_read_matches(backend):
  dv_call = _make_dv_call()
  read = make_read('ACCGT', start=10, cigar='5M', quals=range(10, 15), name='read1')
  got = backend(_options()).encode_read(dv_call, 'ACAGT', read, 10, dv_call.variant.alternate_bases)
  np.testing.assert_array_equal(got, FULL_EXPECTED)


@pytest.mark.parametrize('hp_value,hp_color,polish', [
    (None, 0, None), (0, 0, None), (1, 127, None), (2, 254, None),
    (None, 0, 2), (0, 0, 2), (1, 254, 2), (2, 127, 2)])
def test_encode_read_matches_with_hp_channel(backend, hp_value, hp_color, polish):
  dv_call = _make_dv_call()
  read = make_read('ACCGT', start=10, cigar='5M', quals=range(10, 15), name='read1')
  if hp_value is not None:
    read.hp_values = [hp_value]
  kw = {} if polish is None else {'hp_tag_for_assembly_polishing': polish}
  enc = backend(_options(pi.PILEUP_DEFAULT_CHANNELS + ['haplotype'], **kw))
  got = enc.encode_read(dv_call, 'ACAGT', read, 10, dv_call.variant.alternate_bases)
  np.testing.assert_array_equal(got, np.dstack([FULL_EXPECTED, np.full((1, 5, 1), hp_color, np.uint8)]))


@pytest.mark.parametrize('bases_start,bases_end',
                         [(s, e) for s in range(0, 5) for e in range(6, 12)])
def test_encode_read_spans2(backend, bases_start, bases_end):
  bases = 'AAAACCGTCCC'
  quals = [9, 9, 9, 10, 11, 12, 13, 14, 8, 8, 8]
  ref_start, ref_size = 10, 5
  read_bases = bases[bases_start:bases_end]
  read_quals = quals[bases_start:bases_end]
  read_start = 7 + bases_start
  expected = np.zeros((1, ref_size, 6), dtype=np.uint8)
  for i in range(read_start, read_start + len(read_bases)):
    if ref_start <= i < ref_start + ref_size:
      expected[0, i - ref_start] = FULL_EXPECTED[0, i - ref_start]
  read = make_read(read_bases, start=read_start, cigar=f'{len(read_bases)}M', quals=read_quals, name='read1')
  dv_call = _make_dv_call()
  got = backend(_options()).encode_read(dv_call, 'ACAGT', read, ref_start, dv_call.variant.alternate_bases)
  np.testing.assert_array_equal(got, expected)


def test_encode_read_deletion(backend):
  read = make_read('AAG', start=2, cigar='2M2D1M', quals=range(10, 13), name='read1')
  dv_call = _make_dv_call()
  got = backend(_options()).encode_read(dv_call, 'AACAG', read, 2, dv_call.variant.alternate_bases)
  np.testing.assert_array_equal(
      got, dstack((250, 0, 0, 0, 180), (63, 69, 0, 0, 76), (211, 211, 0, 0, 211), (70, 70, 0, 0, 70),
                  (254, 254, 0, 0, 254), (50, 254, 0, 0, 50)))


def test_encode_read_insertion(backend):
  read = make_read('AAACAG', start=2, cigar='2M1I3M', quals=range(10, 16), name='read1')
  dv_call = _make_dv_call()
  got = backend(_options()).encode_read(dv_call, 'AACAG', read, 2, dv_call.variant.alternate_bases)
  np.testing.assert_array_equal(
      got, dstack((250, 0, 30, 250, 180), (63, 76, 82, 88, 95), (211,) * 5, (70,) * 5, (254,) * 5,
                  (50, 254, 50, 50, 50)))


def _low_qual_call():
  return DeepVariantCall(variant=Variant(reference_name='chr1', start=2, end=3, reference_bases='A',
                                         alternate_bases=['C']))


@pytest.mark.parametrize('min_bq,min_mq', [(0, 0), (1, 3), (4, 4), (2, 0), (3, 1)])
def test_ignores_reads_with_low_quality_bases(backend, min_bq, min_mq):
  enc = backend(_options(read_requirements=pi.ReadRequirements(min_bq, min_mq)))
  for base_qual in range(min_bq + 5):
    read = make_read('AAA', start=1, cigar='3M', quals=[min_bq, base_qual, min_bq], mapq=min_mq)
    actual = enc.encode_read(_low_qual_call(), 'AACAG', read, 1, ['C'])
    assert (actual is None) == (base_qual < min_bq)


@pytest.mark.parametrize('min_bq,min_mq', [(0, 0), (1, 3), (4, 4)])
def test_keeps_reads_with_low_quality_bases(backend, min_bq, min_mq):
  enc = backend(_options(read_requirements=pi.ReadRequirements(min_bq, min_mq)))
  for base_qual in range(1, min_bq + 5):
    read = make_read('AAA', start=1, cigar='3M', quals=[base_qual - 1, min_bq, base_qual + 1], mapq=min_mq)
    assert enc.encode_read(_low_qual_call(), 'AACAG', read, 1, ['C']) is not None


@pytest.mark.parametrize('min_bq,min_mq', [(0, 0), (1, 3), (4, 4), (0, 2)])
def test_ignores_reads_with_low_mapping_quality(backend, min_bq, min_mq):
  enc = backend(_options(read_requirements=pi.ReadRequirements(min_bq, min_mq)))
  for mapping_qual in range(min_mq + 5):
    read = make_read('AAA', start=1, cigar='3M', quals=[min_bq] * 3, mapq=mapping_qual)
    actual = enc.encode_read(_low_qual_call(), 'AACAG', read, 1, ['C'])
    assert (actual is None) == (mapping_qual < min_mq)


@pytest.mark.parametrize('read_name,read_number,alt_allele,read_base,supports_alt', [
    ('read1', 1, 'C', 'C', True), ('read1', 2, 'C', 'C', False), ('read2', 1, 'C', 'G', False),
    ('read2', 2, 'C', 'G', False), ('read3', 1, 'C', 'C', False), ('read3', 2, 'C', 'C', True),
    ('read1', 1, 'G', 'C', False), ('read1', 2, 'G', 'C', False), ('read2', 1, 'G', 'G', True),
    ('read2', 2, 'G', 'G', True), ('read3', 1, 'G', 'C', False), ('read3', 2, 'G', 'C', False)])
def test_read_support_is_respected(backend, read_name, read_number, alt_allele, read_base, supports_alt):
  dv_call = DeepVariantCall(
      variant=Variant(reference_name='chr1', start=10, end=11, reference_bases='A', alternate_bases=['C', 'G']),
      allele_support={'C': ['read1/1', 'read3/2'], 'G': ['read2/1', 'read2/2']})
  read = make_read(read_base, start=10, cigar='1M', quals=[50], name=read_name)
  read.read_number = read_number
  actual = backend(_options()).encode_read(dv_call, 'TAT', read, 9, [alt_allele])
  expected = [{'C': 30, 'G': 180}[read_base], 254, 211, 70, [152, 254][supports_alt], 254]
  assert list(actual[0, 1]) == expected


@pytest.mark.parametrize('read_name,read_number,alt_allele,read_base,other_color,expected_color', [
    ('read1', 1, 'C', 'C', True, int(254.0 * 1.0)), ('read1', 2, 'C', 'C', True, int(254.0 * 0.6)),
    ('read2', 1, 'C', 'G', True, int(254.0 * 0.3)), ('read1', 1, 'C', 'C', False, int(254.0 * 1.0)),
    ('read1', 2, 'C', 'C', False, int(254.0 * 0.6)), ('read2', 1, 'C', 'G', False, int(254.0 * 0.6))])
def test_read_support_multiallelic(backend, read_name, read_number, alt_allele, read_base, other_color,
                                   expected_color):
  dv_call = DeepVariantCall(
      variant=Variant(reference_name='chr1', start=10, end=11, reference_bases='A', alternate_bases=['C', 'G']),
      allele_support={'C': ['read1/1'], 'G': ['read2/1', 'read2/2']})
  read = make_read(read_base, start=10, cigar='1M', quals=[50], name=read_name)
  read.read_number = read_number
  enc = backend(_options(other_allele_supporting_read_alpha=0.3 if other_color else 0.6))
  actual = enc.encode_read(dv_call, 'TAT', read, 9, [alt_allele])
  assert actual[0, 1, 4] == expected_color


# ---- PileupCustomChannels (pileup_image_test.py:663-712) for the channels this build has ----

SEQ50 = 'TTTTATGACAAAAAAGATGCGACGGTTCCGTAACCCATAAGAAAGAACGT'


def _custom(backend, channel_set, cigar='20M5D20M5S', fragment_length=10):
  dv_call = _make_dv_call()
  read = make_read(SEQ50, start=500, cigar=cigar, quals=range(1, 51), name='read1',
                   fragment_length=fragment_length)
  return backend(_options(channel_set)).encode_read(dv_call, SEQ50, read, 500, ['C'])


def test_custom_blank(backend):
  assert set(np.unique(_custom(backend, ['blank'])[:, :, 0])) == {0}


def test_custom_insert_size(backend):
  assert set(np.unique(_custom(backend, ['insert_size'], fragment_length=22)[:, :, 0])) == {0, 5}


@pytest.mark.parametrize('fraglen,expected', [(22, 5), (-22, 5), (1001, 254), (0, 0)])
def test_insert_size_values(backend, fraglen, expected):
  """pileup_channel_lib_test.cc:566-607."""
  dv_call = _make_dv_call()
  seq = 'GATTGGGCCCCAAAAA'[:15]
  read = make_read(seq, start=1, cigar='15M', quals=[30] * 15, fragment_length=fraglen, name='r')
  got = backend(_options(['insert_size'])).encode_read(dv_call, 'A' * 15, read, 1, ['C'])
  assert set(np.unique(got)) == {expected}


@pytest.mark.parametrize('supp,expected', [(True, 254), (False, 0)])
def test_supplementary_alignment(backend, supp, expected):
  """pileup_channel_lib_test.cc:609-633 (unsupporting alpha 0.0, supporting 1.0)."""
  dv_call = _make_dv_call()
  read = make_read('A', start=1, cigar='1M', quals=[30], name='r')
  read.supplementary_alignment = supp
  o = _options(['supplementary_alignment'], allele_unsupporting_read_alpha=0.0,
               allele_supporting_read_alpha=1.0, read_requirements=pi.ReadRequirements(0, 0))
  got = backend(o).encode_read(dv_call, 'A', read, 1, ['C'])
  assert got[0, 0, 0] == expected


def test_ref_rows_custom_options(backend):
  """pileup_channel_lib_test.cc:851-951, the channels this build has."""
  o = dataclasses.replace(
      pi.PileupImageOptions(), width=221, height=100, reference_base_quality=20, base_quality_cap=20,
      allele_unsupporting_read_alpha=1.0, positive_strand_color=20, base_color_offset_a_and_g=1,
      base_color_offset_t_and_c=1, base_color_stride=1, reference_matching_read_alpha=1.0,
      channels=['read_base', 'base_quality', 'mapping_quality', 'strand', 'read_supports_variant',
                'base_differs_from_ref', 'blank', 'insert_size'])
  got = backend(o).encode_reference('GGGCGCTTTTA')
  assert got[0, 10, 0] == 4 and got[0, 8, 0] == 2 and got[0, 0, 0] == 3 and got[0, 3, 0] == 1
  assert got[0, 0, 1] == 254 and got[0, 1, 2] == 254 and got[0, 1, 3] == 20
  assert got[0, 1, 4] == 254 and got[0, 0, 5] == 254 and got[0, 0, 6] == 0 and got[0, 0, 7] == 254


def test_ref_row_wgs_pacbio_channels(backend):
  """SURVEY §8(a) a8 ref-band values; verified on golden.calling_examples / golden.pacbio_examples."""
  o = _options(pi.PILEUP_CHANNELS_WITH_INSERT_SIZE + ['haplotype', 'supplementary_alignment'])
  got = backend(o).encode_reference('ACGTN')
  assert list(got[0, 0]) == [250, 254, 254, 70, 152, 50, 254, 0, 0]


# ---- BuildPileupForOneSample (pileup_image_native_test.cc:277-413) ----

REF11 = 'ACGTACTCCCA'
REF_ROW = [[250, 30, 180, 100, 250, 30, 100, 30, 30, 30, 250], [254] * 11, [254] * 11]
Z = [[0] * 11] * 3


def _ins_row(first):
  return [[first, 30, 180, 100, 0, 30, 100, 30, 30, 30, 0], [190] * 10 + [0], [254] * 10 + [0]]


MATCH_ROW = [[250, 30, 180, 100, 180, 30, 100, 30, 30, 30, 250], [190] * 11, [254] * 11]

BUILD_CASES = {
    'simple_case': dict(
        alts=['G'], reads=[('ACGTGCTCCCA', ['11M'], 'read_2', -1), ('ACGTGCTCCCA', ['11M'], 'read_3', -1)],
        expected=[REF_ROW, MATCH_ROW, MATCH_ROW, Z]),
    'no_reads': dict(alts=['G'], reads=[], expected=[REF_ROW, Z, Z, Z]),
    'numer_of_reads_greater_than_max_reads': dict(
        alts=['AGG'],
        reads=[('ACGTAGGCTCCCA', ['5M', '2I', '5M'], 'read_2', -1), ('ACGTAGGCTCCCA', ['5M', '2I', '5M'], 'read_3', -1),
               ('ACGTAGGGCTCCCA', ['5M', '3I', '5M'], 'read_4', -1), ('ACGTAGGGCTCCCA', ['5M', '3I', '5M'], 'read_5', -1)],
        expected=[REF_ROW, _ins_row(250), _ins_row(250), _ins_row(250)]),
    'image_creation_with_haplotype_sorting': dict(
        alts=['AGG', 'AGGG'],
        reads=[('ACGTAGGCTCCCA', ['5M', '2I', '5M'], 'read_2', 2), ('TCGTAGGCTCCCA', ['5M', '2I', '5M'], 'read_3', 0),
               ('CCGTAGGGCTCCCA', ['5M', '3I', '5M'], 'read_4', 1)],
        expected=[REF_ROW, _ins_row(250), _ins_row(100), _ins_row(30)]),
}


@pytest.mark.parametrize('name', sorted(BUILD_CASES))
def test_build_pileup_for_one_sample(backend, name):
  case = BUILD_CASES[name]
  o = cc_default_options(11, 4, 1, ['read_base', 'base_quality', 'mapping_quality'])
  dv_call = DeepVariantCall(variant=Variant(reference_name='chr1', start=5, end=6, reference_bases='A',
                                            alternate_bases=case['alts']))
  reads = [cc_make_read('chr1', 0, b, c, n, hp) for b, c, n, hp in case['reads']]
  got = backend(o).build_pileup_for_one_sample(dv_call, REF11, reads, 0, ['G'])
  assert got.shape == (4, 11, 3)
  # UnorderedElementsAreArray over rows (pileup_image_native_test.cc:268-269)
  got_rows = sorted(tuple(got[r].T.reshape(-1).tolist()) for r in range(4))
  exp_rows = sorted(tuple(np.array(e, dtype=np.uint8).reshape(-1).tolist()) for e in case['expected'])
  assert got_rows == exp_rows
  # the reference band is always first and blank rows always last
  np.testing.assert_array_equal(got[0].T, np.array(REF_ROW, dtype=np.uint8))


def test_sorts_by_haplotype_then_allele_support(backend):
  """pileup_image_native_test.cc:872-969: exact row ORDER."""
  o = cc_default_options(9, 8, 1, ['read_base'])
  o = dataclasses.replace(o, sort_by_haplotypes=True, sort_by_alt_allele_support=True, random_seed=12345)
  call = DeepVariantCall(
      variant=Variant(reference_name='chr1', start=4, end=5, reference_bases='A', alternate_bases=['G', 'T']),
      allele_support={'G': ['read_z_supports_g/0', 'read_g2/0'], 'T': ['read_a_supports_t/0']})
  z = cc_make_read('chr1', 0, 'CCCCGCCCC', ['9M'], 'read_z_supports_g', 1)
  a = cc_make_read('chr1', 0, 'CCCCTCCCC', ['9M'], 'read_a_supports_t', 1)
  ref1 = cc_make_read('chr1', 1, 'CCCCACCCC', ['9M'], 'read_ref1', 1)
  g2 = cc_make_read('chr1', 0, 'GGGGGCCCC', ['9M'], 'read_g2', 2)
  ref2 = cc_make_read('chr1', 0, 'AAAAACCCC', ['9M'], 'read_ref2', 2)
  ref3 = cc_make_read('chr1', 0, 'TTTTTCCCC', ['9M'], 'read_ref3', 2)
  expected_order = [z, a, ref1, g2, ref2, ref3]
  enc = backend(o)
  ref_bases = 'CCCCACCCC'
  got = enc.build_pileup_for_one_sample(call, ref_bases, [ref3, ref2, g2, ref1, a, z], 0, ['G', 'T'])
  assert got.shape == (8, 9, 1)
  np.testing.assert_array_equal(got[0:1], enc.encode_reference(ref_bases))
  for i, r in enumerate(expected_order):
    np.testing.assert_array_equal(got[1 + i:2 + i], enc.encode_read(call, ref_bases, r, 0, ['G', 'T']),
                                  err_msg=f'row {i + 1}')
  assert not got[7].any()


def test_downsample_prefix_kat():
  """std::shuffle(iota(n), mt19937_64(2101079370)) is implementation-defined: both standard libraries are restated.
  libstdc++ prefixes: SURVEY Appendix A (g++ 13.3).  libc++ (shuffle_stdlib 0, the default): which reads the reference's
  golden.allele_frequency_examples drops at n = 96 and n = 103 (tools/check_downsample_golden.py) = the tail of the permutation."""
  import oracle_lib
  exp = {96: [32, 69, 31, 60, 53, 68, 49, 39, 76, 54, 18, 82],
         100: [32, 69, 31, 60, 53, 68, 49, 39, 76, 54, 18, 82],
         150: [32, 69, 31, 60, 53, 68, 107, 39, 76, 54, 113, 122],
         300: [181, 69, 31, 249, 259, 68, 107, 216, 76, 54, 113, 122]}
  for n, prefix in exp.items():
    t = oracle_lib.shuffle_table(n, 2101079370, 95, shuffle_stdlib=1)
    assert t[:12].tolist() == prefix
    assert sorted(t.tolist()) == list(range(n))
  for flavour in (0, 1):
    assert oracle_lib.shuffle_table(95, 2101079370, 95, flavour).tolist() == list(range(95))
  t96 = oracle_lib.shuffle_table(96, 2101079370, 95)
  assert t96[95] == 91 and sorted(t96.tolist()) == list(range(96))            # the golden image at chr20:61645 lacks query read 91
  t103 = oracle_lib.shuffle_table(103, 2101079370, 95)
  assert sorted(t103[95:].tolist()) == [15, 24, 46, 50, 60, 67, 75, 102]      # chr20:61350 lacks exactly these eight
  assert sorted(t103.tolist()) == list(range(103))
  # the product's host table (csrc/dvb_encoder.cu) is an independent restatement of both
  import ctypes
  from deepvariant_b200 import _lib
  for flavour in (0, 1):
    for n in (96, 103, 150, 300, 1000):
      got = np.zeros(n, dtype=np.int32)
      assert _lib.lib().dvb_shuffle_table(n, 2101079370, flavour, got.ctypes.data_as(ctypes.c_void_p)) == 0
      np.testing.assert_array_equal(got, oracle_lib.shuffle_table(n, 2101079370, 95, flavour))


# ---- "Opt Channels": whole-read statistics (pileup_channel_lib_test.cc:419-504, 696-832, 851-951) ------------------------

OPT = ['read_mapping_percent', 'avg_base_quality', 'identity', 'gap_compressed_identity', 'gc_content']


def _scaled(value, cap):
  return int(np.float32(254.0) * (np.float32(min(value, cap)) / np.float32(cap)))


@pytest.mark.parametrize('channel,bases,cigar,quals,value,cap', [
    ('read_mapping_percent', 'AAAAATTTTT', '5M5D', [30] * 10, 50, 100),          # ReadMappingPercentTest.BasicCase
    ('avg_base_quality', 'AAAAATTTTT', '10M', list(range(1, 11)), 5, 93),         # AvgBaseQualityTest.BasicCase
    ('identity', 'AAAAATTTTT', '5M1I4M', [30] * 10, 90, 100),                     # IdentityTest.BasicCase
    ('identity', 'AAAAATTTTT', '5=1X4=', [30] * 10, 90, 100),                     # IdentityTest.PacBioStyleCigar
    ('gap_compressed_identity', 'AAAAATTTTT', '3M4I3M', [30] * 10, 85, 100),      # GapCompressedIdentityTest.InsertionCase
    ('gap_compressed_identity', 'AAAAATTTTT', '3M4D3M', [30] * 10, 85, 100),      # .DeletionCase
    ('gap_compressed_identity', 'AAAAATTTTT', '3=2X2I3=', [30] * 10, 66, 100),    # .PacBioStyleCigar
    ('gc_content', 'GGGGGCCCCC', '10M', [30] * 10, 100, 100),                     # GcContestTest.AllGc
    ('gc_content', 'GGGGGTTTTT', '10M', [30] * 10, 50, 100),                      # GcContestTest.HalfGc
])
def test_opt_channel_read_statistics(backend, channel, bases, cigar, quals, value, cap):
  """The statistic the reference's unit tests assert, seen through the channel's pixel value ScaleColor(value, cap)."""
  read = make_read(bases, start=2, cigar=cigar, quals=quals, name='r')
  got = backend(_options([channel], read_requirements=pi.ReadRequirements(0, 0))).encode_read(_make_dv_call(), 'A' * 21, read, 0, ['C'])
  drawn = got[0, :, 0][got[0, :, 0] > 0] if value else got[0, :, 0]
  assert len(drawn) > 0 and set(np.unique(drawn)) == {_scaled(value, cap)}, (np.unique(got), _scaled(value, cap))


def test_opt_channels_get_channel_data_kat(backend):
  """GetChannelDataTest (pileup_channel_lib_test.cc:696-832): read GGGCGCTTTTAT / 11M, every quality 33 ->
  mapping percent 231, average base quality 90, identity 231, gap-compressed identity 254, GC content 127;
  GetRefChannelDataTest (:851-951): reference row 254, 254, 254, 254 and the window's own GC content (127)."""
  read = make_read('GGGCGCTTTTAT', start=1, cigar='11M', quals=[33] * 12, name='r')
  enc = backend(_options(OPT, read_requirements=pi.ReadRequirements(0, 0)))
  got = enc.encode_read(_make_dv_call(), 'GGGCGCTTTTAT', read, 1, ['C'])
  assert got.shape == (1, 12, 5)
  np.testing.assert_array_equal(got[0, 3], [231, 90, 231, 254, 127])
  np.testing.assert_array_equal(got[0, :11], np.tile([231, 90, 231, 254, 127], (11, 1)))
  assert not got[0, 11].any()                                    # 11M covers 11 of the 12 columns
  ref = enc.encode_reference('GGGCGCTTTTAT')
  np.testing.assert_array_equal(ref[0], np.tile([254, 254, 254, 254, 127], (12, 1)))
  assert enc.encode_reference('ATATATATATAT')[0, 0, 4] == 0 and enc.encode_reference('GCGCGCGCGCGC')[0, 5, 4] == 254


# ---- per-base homopolymer channels (pileup_channel_lib_test.cc:506-564, 696-832, 851-951) --------------------------------

@pytest.mark.parametrize('bases,expected', [
    ('GGGATAATA', [1, 1, 1, 0, 0, 0, 0, 0, 0]),     # IsHomoPolymerTest.IsHomopolymerBeginning
    ('ATTGGGTTA', [0, 0, 0, 1, 1, 1, 0, 0, 0]),     # .IsHomopolymerMiddle
    ('ATAATAGGG', [0, 0, 0, 0, 0, 0, 1, 1, 1]),     # .IsHomopolymerEnd
    ('AAAAAAAAA', [1] * 9),                         # .IsHomopolymerAll
])
def test_is_homopolymer(backend, bases, expected):
  read = make_read(bases, start=3, cigar=f'{len(bases)}M', quals=[30] * len(bases), name='r')
  got = backend(_options(['is_homopolymer'], read_requirements=pi.ReadRequirements(0, 0))).encode_read(_make_dv_call(), 'C' * 15, read, 0, ['C'])
  assert got[0, 3:3 + len(bases), 0].tolist() == [254 * e for e in expected]
  assert not got[0, :3].any() and not got[0, 3 + len(bases):].any()


def test_homopolymer_weighted(backend):
  """HomoPolymerWeightedTest.BasicCase (run lengths 1,1,2,2,3,3,3,4,4,4,4,5,5,5,5,5) and .WeightedHomoPolymerMax (10 G + 20 A)."""
  enc = backend(_options(['homopolymer_weighted'], read_requirements=pi.ReadRequirements(0, 0)))
  seq = 'GATTGGGCCCCAAAAA'
  got = enc.encode_read(_make_dv_call(), 'C' * 20, make_read(seq, start=2, cigar='16M', quals=[30] * 16, name='r'), 0, ['C'])
  runs = [1, 1, 2, 2, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 5]
  assert got[0, 2:18, 0].tolist() == [_scaled(r, 30) for r in runs]
  seq = 'G' * 10 + 'A' * 20 + 'C' * 35          # a 35-long run saturates at 30 -> 254
  got = enc.encode_read(_make_dv_call(), 'T' * 70, make_read(seq, start=1, cigar='65M', quals=[30] * 65, name='r'), 0, ['C'])
  assert got[0, 1:66, 0].tolist() == [_scaled(10, 30)] * 10 + [_scaled(20, 30)] * 20 + [254] * 35


def test_homopolymer_channels_get_channel_data_kat(backend):
  """GetChannelDataTest / GetRefChannelDataTest with read = reference window = GGGCGCTTTTAT (11M):
  is_homopolymer[1] = 254, [4] = 0; homopolymer_weighted[1] = 25 (run 3), [9] = 33 (run 4); same values on the reference row."""
  chans = ['is_homopolymer', 'homopolymer_weighted']
  enc = backend(_options(chans, read_requirements=pi.ReadRequirements(0, 0)))
  read = make_read('GGGCGCTTTTAT', start=1, cigar='11M', quals=[33] * 12, name='r')
  got = enc.encode_read(_make_dv_call(), 'GGGCGCTTTTAT', read, 1, ['C'])
  assert got[0, 1, 0] == 254 and got[0, 4, 0] == 0 and got[0, 1, 1] == 25 and got[0, 9, 1] == 33
  ref = enc.encode_reference('GGGCGCTTTTAT')
  assert ref[0, 1, 0] == 254 and ref[0, 4, 0] == 0 and ref[0, 1, 1] == 25 and ref[0, 9, 1] == 33
  # an insertion anchor takes the value at the first inserted base's index, a deletion anchor at the base before it
  ins = make_read('ACGTTTTGCA', start=2, cigar='3M4I3M', quals=[30] * 10, name='r')
  got = enc.encode_read(_make_dv_call(), 'A' * 12, ins, 0, ['C'])
  assert got[0, 4, 0] == 254 and got[0, 4, 1] == _scaled(4, 30)       # anchor column 2+3-1 = 4 <- read index 3 ('T' of TTTT)


# ---- channels_enum_to_blank: GetChannelDataTest with its three parameter sets (pileup_channel_lib_test.cc:696-849) ---------------

GET_CHANNEL_DATA_CHANNELS = ['read_base', 'base_quality', 'mapping_quality', 'strand', 'read_supports_variant', 'base_differs_from_ref',
                             'read_mapping_percent', 'avg_base_quality', 'identity', 'gap_compressed_identity', 'gc_content', 'is_homopolymer',
                             'homopolymer_weighted', 'blank', 'insert_size', 'supplementary_alignment']


@pytest.mark.parametrize('to_blank', [(), (1,), (1, 3)])      # {}, {CH_READ_BASE}, {CH_READ_BASE, CH_MAPPING_QUALITY}
def test_get_channel_data_with_channels_enum_to_blank(backend, to_blank):
  """All 16 channels of the reference's test in one read row, its options (mapping_quality_cap 1, positive_strand_color 20,
  allele_unsupporting_read_alpha 1, base colour offsets / stride 1, base_quality_cap 20, matching alpha 1, mismatching alpha 0), read =
  window = GGGCGCTTTTAT / 11M at position 1, qualities 33, fragment length 1000.  A blanked channel's read pixels are 0; the rest and
  the reference row are what they are without blanking."""
  o = pi.PileupImageOptions(reference_band_height=5, mapping_quality_cap=1, positive_strand_color=20, allele_unsupporting_read_alpha=1.0,
                            base_color_offset_a_and_g=1, base_color_offset_t_and_c=1, base_color_stride=1, base_quality_cap=20,
                            reference_matching_read_alpha=1.0, reference_mismatching_read_alpha=0.0, width=13, height=100,
                            channels=list(GET_CHANNEL_DATA_CHANNELS), num_channels=16, channels_enum_to_blank=tuple(to_blank))
  read = Read(fragment_name='r', read_number=0, aligned_sequence=b'GGGCGCTTTTAT', aligned_quality=bytes([33] * 12), position=1, mapping_quality=90,
              cigar=parse_cigar_string('11M'), fragment_length=1000)
  enc = backend(o)
  got = enc.encode_read(DeepVariantCall(variant=Variant(start=-100)), 'NGGGCGCTTTTAT', read, 0, [])[0]   # columns = reference positions (image_start_pos 0)
  ch = {n: i for i, n in enumerate(GET_CHANNEL_DATA_CHANNELS)}
  if 1 not in to_blank:
    assert [got[c, ch['read_base']] for c in (11, 9, 1, 4)] == [4, 2, 3, 1]
  else:
    assert not got[:, ch['read_base']].any()
  assert got[1, ch['base_quality']] == 254
  assert got[1, ch['mapping_quality']] == (0 if 3 in to_blank else 254)
  assert got[1, ch['strand']] == 20 and got[1, ch['read_supports_variant']] == 254 and got[1, ch['base_differs_from_ref']] == 254
  assert got[3, ch['read_mapping_percent']] == 231 and got[3, ch['avg_base_quality']] == 90 and got[9, ch['identity']] == 231
  assert got[9, ch['gap_compressed_identity']] == 254 and got[3, ch['gc_content']] == 127
  assert got[1, ch['is_homopolymer']] == 254 and got[4, ch['is_homopolymer']] == 0
  assert got[1, ch['homopolymer_weighted']] == 25 and got[9, ch['homopolymer_weighted']] == 33
  assert got[1, ch['blank']] == 0 and got[1, ch['insert_size']] == 254
  # the reference band ignores the blank set (CalculateRefRows draws every channel): same row with and without it
  ref_row = enc.encode_reference('NGGGCGCTTTTAT')
  plain = backend(dataclasses.replace(o, channels_enum_to_blank=())).encode_reference('NGGGCGCTTTTAT')
  np.testing.assert_array_equal(ref_row, plain)
  assert ref_row[0, 1, ch['read_base']] == 3 and ref_row[0, 1, ch['strand']] == 20
