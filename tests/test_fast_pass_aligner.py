"""Smith-Waterman with libssw's tie-breaking (csrc/dvb_ssw.cu) and the FastPassAligner (deepvariant_b200/fast_pass_aligner.py) against the
known answers of the reference's tests, transcribed as data: deepvariant/realigner/ssw_test.cc:47-58, python/ssw_misc_test.py:44-84,
python/ssw_wrap_test.py:37-72, fast_pass_aligner_test.cc:171-202, 311-362, 431-480, 482-756.  CPU-only (host code)."""
import pytest

from deepvariant_b200 import fast_pass_aligner as fpa, ssw
from deepvariant_b200.fast_pass_aligner import D, I, M, S

REF = 'ATCAAGGGAAAAAGTGCCCAGGGCCAAATATGTTTTGGGTTTTGCAGGACAAAGTATGGTTGAAACTGAGCTGAAGATATG'     # FastPassAlignerTest::SetUp


@pytest.mark.parametrize('params,ref,query,expected', [
    ((4, 2, 4, 2), 'tttt', 'ttAtt', dict(cigar_string='2=1I2=')),
    ((4, 2, 4, 2), 'TTTTGGGGGGGGGGGGG', 'TTATTGGGGGGGGGGGGG', dict(cigar_string='2=1I15=')),
    ((2, 2, 3, 1), 'CAGCCTTTCTGACCCGGAAATCAAAATAGGCACAACAAA', 'CTGAGCCGGTAAATC',
     dict(sw_score=21, ref_begin=8, ref_end=21, query_begin=0, query_end=14, mismatches=2, cigar_string='4=1X4=1I5=')),
    ((2, 2, 3, 1), 'CTGAGCCGGTAAATC', 'CAGCCTTTCTGACCCGGAAATCAAAATAGGCACAACAAA',
     dict(sw_score=21, query_begin=8, query_end=21, ref_begin=0, ref_end=14, mismatches=2, cigar_string='8S4=1X4=1D5=17S')),
    ((4, 6, 8, 1), 'TTTGCCGAAGTTAAACCC', 'GCCGAAGTTA', dict(cigar_string='10=', ref_begin=3)),
])
def test_ssw_known_answers(params, ref, query, expected):
  a = ssw.Aligner(*params)
  assert a.set_reference_sequence(ref) == len(ref)
  al = a.align(query)
  for k, v in expected.items():
    assert getattr(al, k) == v, (k, al)


HAP1 = 'AAGTGCCCAGGGCCAAATGTTTTGGGTTTTGCAGGACAAAGTATGGTT'           # reference with 1 del
HAP2 = 'AAGTGCCCAGGGCCAAATATGCACAGGGTTTTGCAGGACAAAGTATGGTT'         # reference with 1 sub
READS = ['CAGGGCCAAATGTTT', 'GCCATATATGCACAGGGTTATG', 'TTGGGTTGCAGGACA', 'ACAGGGTTTTTTGCAGGACAA', 'TGTTGGGTTCAGCAGTTTT']


def test_ssw_align_reads_to_haplotypes():
  """SswAlignReadsToHaplotypes_Test (:431-480): positions, cigars (soft clips, indel placement) and scores of 5 reads x 2 haplotypes."""
  a = fpa.FastPassAligner()
  a.reference = REF
  a.reads = list(READS)
  a.set_options(kmer_size=3)
  a.haplotypes = [HAP1, HAP2]
  a.align_haplotypes_to_reference()
  a.ssw_align_reads_to_haplotypes(40)
  got = [[(r.position, r.cigar, r.score) for r in ha.read_alignment_scores] for ha in a.read_to_haplotype_alignments]
  none = (fpa.K_NOT_ALIGNED, '', 0)
  assert got[0] == [(7, '15=', 60), none, (21, '5=2D10=', 51), (23, '3S3=2I13=', 55), none]
  assert got[1] == [(7, '11=4S', 44), (11, '4=1X14=1X2=', 68), (25, '2S3=2D10=', 43), (22, '6=2I13=', 67), none]


def test_align_haplotypes_to_reference():
  """AlignHaplotypesToReference_Test (:319-362)."""
  a = fpa.FastPassAligner()
  a.reference = 'AGAAGGTCCCTTTGCCGAAGTTAAACCCTTTCGCGC'
  a.haplotypes = ['GTCCCTTTGCCGAAGTTAAACCCTTT', 'GTCCCTTTGCCGAGTTAAACCCTTT', 'GTCCCTATGCCGAAGTTAAACCCTTT']
  a.align_haplotypes_to_reference()
  got = [(h.cigar, h.cigar_ops, h.ref_pos, h.is_reference) for h in a.read_to_haplotype_alignments]
  assert got == [('26=', [(M, 26)], 5, True), ('12=1D13=', [(M, 12), (D, 1), (M, 13)], 5, False),
                 ('6=1X19=', [(M, 6), (M, 1), (M, 19)], 5, False)]


def test_fast_align_reads_to_haplotype():
  """FastAlignReadsToHaplotypeTest (:171-202)."""
  a = fpa.FastPassAligner()
  a.reference = REF
  a.reads = ['AAACCC', 'CTCTCT', 'TGAGCTGAAG']
  a.set_options(kmer_size=3)
  a.ref_prefix_len = 100       # the reference's test leaves ref_prefix_len_ / ref_suffix_len_ unset: no zero-coverage check
  a.build_index()
  assert a.kmer_index['CTC'] == [(1, 0), (1, 2)] and a.kmer_index['AAA'] == [(0, 0)]
  scores = [fpa.ReadAlignment() for _ in a.reads]
  assert a.fast_align_reads_to_haplotype('TGAGCTGAAGTTAAACCC', scores) == 16 * 4
  assert [(s.position, s.cigar, s.score) for s in scores] == [(12, '6=', 24), (fpa.K_NOT_ALIGNED, '', 0), (0, '10=', 40)]


READ_TO_REF_REFERENCE = 'CTCTGTAATCGGATCATGTTTTGGGTTTTGCAGGACAAAGTATGGTTGAAACTGAGCTGAAGATATG'
READ_TO_REF_CASES = [      # (name, haplotype, read, read-to-haplotype cigar, expected read-to-reference cigar)      :552-756
    ('ins_snp_merge', 'CGGATCATGTTTTTTGGGTTTTCAGGACAAAGTATGGTTGAAACTG', 'GATCATGATTTTTGGGTTTTCAG', '7=1X15=', [(M, 7), (I, 2), (M, 11), (D, 1), (M, 3)]),
    ('ins_ins_merge', 'CGGATCATGTTTTTTGGGTTTTCAGGACAAAGTATGGTTGAAACTG', 'GATCATGTTTTTTTGGGTTTTCAG', '7=1I16=', [(M, 7), (I, 3), (M, 11), (D, 1), (M, 3)]),
    ('del_del_merge', 'CGGATCATGTTTGGGTTTTCAGGACAAAGTATGGTTGAAACTG', 'GATCATGTTGGGTTTTCAGGACAAA', '7=1D18=', [(M, 7), (D, 2), (M, 9), (D, 1), (M, 9)]),
    ('del_ins_merge', 'CGGATCATGTTTGGGTTTTCAGGACAAAGTATGGTTGAAACTG', 'GATCATGTTTTTGGGTTTTCAGGACAAA', '7=2I19=', [(M, 7), (I, 1), (M, 11), (D, 1), (M, 9)]),
    ('del_ins_merge2', 'CGGATCATGTGGGTTTTCAGGACAAAGTATGGTTGAAACTG', 'GATCATGTTTGGGTTTTCAGGACAAA', '7=2I17=', [(M, 7), (D, 1), (M, 10), (D, 1), (M, 9)]),
    ('ins_del_merge', 'CGGATCATGTTTTTTGGGTTTTCAGGACAAAGTATGGTTGAAACTG', 'GATCATGTTTTTGGGTTTTCAGGACAAA', '7=1D21=', [(M, 7), (I, 1), (M, 11), (D, 1), (M, 9)]),
    ('2ins_3del_merge', 'CGGATCATGTTTTTTGGGTTTTCAGGACAAAGTATGGTTGAAACTG', 'GATCATGTTTGGGTTTTCAGGACAAA', '7=3D19=', [(M, 7), (D, 1), (M, 10), (D, 1), (M, 9)]),
    ('1ins_1del_back_to_back', 'CGGATCATGTTTTGGGTTTTCAGGACAAAGTATGGTTGAAACTG', 'GATCATGTTTTGGGTTTTCCAGGACAAA', '18=1I9=', [(M, 28)]),
    ('1ins_1del_consecutive', 'CGGATCATGTTTTGGGTTTTTTGCAGGACAAAGTATGGTTGAAACTG', 'GATCATGTTTTGGGTTTTGCAGGACAAA', '16=2D12=', [(M, 28)]),
    ('1del_1ins_consecutive2', 'CGGATCATGTTTTGGGTTTTGCGCAGGACAAAGTATGGTTGAAACTG', 'GATCATGTTTTGGGTTGCGCAGGACAAA', '16=2D12=', [(M, 28)]),
    ('two_dels_different_positions', 'CGGATCATGTTTGGGTTTTGCAGGACAAAGTATGGTTGAAACTG', 'GATCATGTTTGGTTTT', '10=1D6=', [(M, 7), (D, 1), (M, 3), (D, 1), (M, 6)]),
]


@pytest.mark.parametrize('name,hap,read,cigar,expected', READ_TO_REF_CASES, ids=[c[0] for c in READ_TO_REF_CASES])
def test_calculate_read_to_ref_alignment(name, hap, read, cigar, expected):
  a = fpa.FastPassAligner()
  a.reference = READ_TO_REF_REFERENCE
  a.haplotypes = [hap]
  a.align_haplotypes_to_reference()
  a.reads = [read]
  got = a.calculate_read_to_ref_alignment(0, fpa.ReadAlignment(2, cigar, 100), a.read_to_haplotype_alignments[0].cigar_ops)
  assert [tuple(x) for x in got] == expected


def test_calculate_read_to_ref_alignment_simple_and_soft_clipped_haplotype():
  a = fpa.FastPassAligner()            # CalculateReadToRefAlignment_MatchMismatch_Test (:482-509)
  a.reference = REF
  a.haplotypes = ['TGTTTAGGGTTTTGCAGGACAAAGTATGGTTGAAACTG']
  a.align_haplotypes_to_reference()
  a.reads = ['TGTTTAGGGTTTTGCAGGA']
  assert a.calculate_read_to_ref_alignment(0, fpa.ReadAlignment(7, '19=', 100), a.read_to_haplotype_alignments[0].cigar_ops) == [[M, 19]]
  a = fpa.FastPassAligner()            # ..._HaplotypeSoftClipped_Test (:511-550)
  a.reference = 'nnnnnnnnnnnTGTTTTGGGTTTTGCAGGACAAAGTATGGTTGAAACTGAGCTGAAGATATG'
  a.haplotypes = ['GATCATGTTTAGGGTTTTGCAGGACAAAGTATGGTTGAAACTG']
  a.align_haplotypes_to_reference()
  a.reads = ['GATCATGTTTAGGGTTTT']
  assert a.calculate_read_to_ref_alignment(0, fpa.ReadAlignment(0, '19=', 100), a.read_to_haplotype_alignments[0].cigar_ops) == [[S, 5], [M, 13]]


def test_positions_map_and_merge_cigar_op():
  assert fpa.set_positions_map(10, '10=') == [0] * 10                                        # :364-429
  assert fpa.set_positions_map(10, '3=2I5=') == [0, 0, 0, 0, -1, -2, -2, -2, -2, -2]
  assert fpa.set_positions_map(8, '3=2D5=') == [0, 0, 0, 2, 2, 2, 2, 2]
  c = []
  fpa.merge_cigar_op(M, 5, 10, c)                                                            # :963-1041
  assert c == [[M, 5]]
  fpa.merge_cigar_op(D, 2, 10, c)
  fpa.merge_cigar_op(M, 3, 10, c)
  fpa.merge_cigar_op(M, 1, 10, c)
  assert c == [[M, 5], [D, 2], [M, 4]]
  fpa.merge_cigar_op(M, 5, 10, c)                                                            # clipped to the read length
  assert c == [[M, 5], [D, 2], [M, 5]]
  fpa.merge_cigar_op(M, 1, 10, c)
  assert c == [[M, 5], [D, 2], [M, 5]]


def test_score_threshold_and_normalisation():
  a = fpa.FastPassAligner()
  a.set_options(read_size=100, realignment_similarity_threshold=0.8, match=4, mismatch=6)      # :1043-1058: 4*100*0.8 - 6*100*0.2
  a.calculate_ssw_alignment_score_threshold()
  assert a.ssw_alignment_score_threshold == 200
  a.reference = 'AAAAACCCCCGGGGGTTTTT'
  assert a.is_alignment_normalized([(M, 10), (M, 10)], 0, 'AAAAACCCCCGGGGGTTTTT')
  assert not a.is_alignment_normalized([(M, 7), (D, 2), (M, 11)], 0, 'AAAAACCCGGGGGTTTTT')      # deleting CC after C: can shift left
  assert a.is_alignment_normalized([(M, 5), (D, 2), (M, 13)], 0, 'AAAAACCCGGGGGTTTTT')
  assert not a.is_alignment_normalized([(M, 7), (I, 1), (M, 13)], 0, 'AAAAACCCCCCGGGGGTTTTT')
  assert a.is_alignment_normalized([(M, 5), (I, 1), (M, 15)], 0, 'AAAAACCCCCCGGGGGTTTTT')


@pytest.mark.parametrize('seed', range(4))
def test_native_fast_pass_equals_the_python_pass(seed):
  """dvb_fast_pass_scores (all haplotypes in one call) == fast_align_reads_to_haplotype per haplotype, incl. the zero-coverage rule."""
  import random
  rng = random.Random(seed)
  ref = ''.join(rng.choice('ACGT') for _ in range(300))
  haps = [ref]
  for _ in range(4):
    h = list(ref)
    for _ in range(rng.randrange(1, 4)):
      p = rng.randrange(60, 240)
      kind = rng.random()
      if kind < 0.4:
        h[p] = rng.choice('ACGT')
      elif kind < 0.7:
        h[p] = h[p] + rng.choice('ACGT') * rng.randrange(1, 4)
      else:
        h[p] = ''
    haps.append(''.join(h))
  reads = []
  for _ in range(80):
    src = rng.choice(haps)
    p = rng.randrange(0, len(src) - 60)
    r = list(src[p:p + rng.randrange(40, 60)])
    for _ in range(rng.choice([0, 0, 1, 2, 4])):
      r[rng.randrange(len(r))] = rng.choice('ACGTN')
    reads.append(''.join(r))
  def make():
    a = fpa.FastPassAligner()
    a.reference, a.haplotypes, a.reads = ref, haps, list(reads)
    a.set_options(kmer_size=rng.choice([8, 12, 32]))
    a.ref_prefix_len, a.ref_suffix_len = 20, 20
    a.build_index()
    return a
  state = rng.getstate()
  a = make()
  rng.setstate(state)
  b = make()
  a.fast_align_reads_to_haplotypes()
  b.fast_align_reads_to_haplotypes_py()
  assert [(h.haplotype_index, h.haplotype_score, [(r.position, r.cigar, r.score) for r in h.read_alignment_scores]) for h in a.read_to_haplotype_alignments] == \
         [(h.haplotype_index, h.haplotype_score, [(r.position, r.cigar, r.score) for r in h.read_alignment_scores]) for h in b.read_to_haplotype_alignments]
  assert any(h.haplotype_score > 0 for h in a.read_to_haplotype_alignments)


def test_run_merge_equals_base_by_base_merge():
  """_merge_bases (one call per run of equal operation pairs) against `count` calls of _merge_one_base, the reference's
  base-by-base form (fast_pass_aligner.cc:760-800), from every kind of preceding CIGAR - incl. the I-after-D / D-after-I rewrite
  of MergeCigarOp and the clamp at the read length."""
  import copy
  import numpy as np
  from deepvariant_b200 import fast_pass_aligner as fpa
  rng = np.random.default_rng(3)
  ops = [fpa.M, fpa.I, fpa.D, fpa.S]
  n = 0
  for _ in range(4000):
    read_len = int(rng.integers(1, 40))
    start = [[ops[int(rng.integers(0, 4))], int(rng.integers(1, 6))] for _ in range(int(rng.integers(0, 4)))]
    start = [o for i, o in enumerate(start) if i == 0 or o[0] != start[i - 1][0]]
    read_op, hap_op = ops[int(rng.integers(0, 4))], ops[int(rng.integers(0, 4))]
    if (read_op, hap_op) in ((fpa.D, fpa.I), (fpa.I, fpa.D)):
      continue
    count = int(rng.integers(1, 30))
    a, b = copy.deepcopy(start), copy.deepcopy(start)
    for _ in range(count):
      fpa._merge_one_base(read_op, hap_op, read_len, a)
    fpa._merge_bases(read_op, hap_op, count, read_len, b)
    assert a == b, (start, read_op, hap_op, count, read_len, a, b)
    n += 1
  assert n > 3000
