"""--normalize_reads: known answers of deepvariant/allelecounter_test.cc NormalizeCigar* (:1224-1577) for
deepvariant_b200/normalize_reads.py, and the region-level rewrite."""
import pytest

from deepvariant_b200 import normalize_reads as nr
from deepvariant_b200.protos import Read, parse_cigar_string

REF151 = (b'GTCAAAGGGTGTTGCATCTGCTTAAACTCACACATCTCGAAGGTTGCTGTGAAGGTAAACAG'
          b'AAAGCAACGTAAGGCACGGATGTTGATTCGTGTGTCGTGTGTGTGTGTGTGTGTGTGTGTGT'
          b'GCGAAATTTGTACAGCAGTACCTGCAT')
REF19 = b'AGTGGGGGGGGGATGGGGG'
REF26 = b'ATAGACAGATAGATAGATCGATAGAT'
REF34 = b'ATGTTCCTTCCTTCCTTCCTTCCTTCCTTCCACT'

# (name, interval bases, interval_offset, read bases, cigar, expected cigar, expected shift)
CASES = [
    ('Del', REF151, 82, b'TGTTGATTCGTGTGTCGTGTGTGTGTGTGTGCGAAATTTGTACAGCAGTACCTGCAT', '31M12D26M', '16M12D41M', 0),
    ('Ins', REF151, 82, b'TGTTGATTCGTGTGTGTCGTGTGTGTGTGTGTGTGTGTGTGTGTGCGAAATTTGTACAGCAGTACCTGCAT', '13M2I56M', '9M2I60M', 0),
    ('InsDel', REF19, 0, b'AGTGGGGGGGGGGATGGGG', '7M1I10M1D1M', '3M1I11M1D4M', 0),
    ('InsertAtTheEnd', REF19, 0, b'AGTGGGGGGGGGGG', '12M2I', '3M2I9M', 0),
    ('TwoDelsMerged', REF26[:22], 5, b'CAGATAGA', '4M9D1M3D3M', '2M12D6M', 0),
    ('DelInsMerged', REF34, 4, b'TCCTTCCTTCCTCCTTCCTTCCTTCCTTCCTTCCA', '11M1D4M8I12M', '4M7I24M', 0),
    ('InsShiftedToEdge', REF34, 8, b'TCCTTCCTTCCTTCCTTCCTTCCTTCCACT', '4M4I22M', '30M', -4),
    ('InsShiftedAllTheWayToSoftClip', REF34, 8, b'GGGTCCTTCCTTCCTTCCTTCCTTCCTTCCACT', '3S4M4I22M', '3S30M', -4),
    ('DelInsMergedNoShift', REF34, 4, b'TCCTTCCTTCCTCCTTCCTTCCTTCCTTCCTTCCA', '11M1D8I16M', '4M7I24M', 0),
]


@pytest.mark.parametrize('name,ref,offset,seq,cigar,want,want_shift', CASES, ids=[c[0] for c in CASES])
def test_normalize_cigar_known_answers(name, ref, offset, seq, cigar, want, want_shift):
  modified, got, shift = nr.normalize_cigar(seq, offset, parse_cigar_string(cigar), ref)
  assert modified
  assert got == parse_cigar_string(want)
  assert shift == want_shift


def test_normalized_read_still_spells_the_same_haplotype():
  # the rewritten alignment must produce the very same sequence when applied to the reference
  def spell(ref, pos, cigar, seq):
    out, r, q = b'', pos, 0
    for op, ln in cigar:
      if op in (0, 7, 8):
        out += seq[q:q + ln]; r += ln; q += ln
      elif op == 1:
        out += seq[q:q + ln]; q += ln
      elif op == 2:
        r += ln
      elif op == 4:
        q += ln
    return out, r
  for name, ref, offset, seq, cigar, want, want_shift in CASES:
    _, got, shift = nr.normalize_cigar(seq, offset, parse_cigar_string(cigar), ref)
    assert sum(ln for op, ln in got if op in (0, 1, 4, 7, 8)) == len(seq), name


def test_already_normalised_and_empty():
  assert nr.normalize_cigar(b'ACGT', 3, [], REF34) == (False, [], 0)
  seq = REF34[4:24]
  assert nr.normalize_cigar(seq, 4, [(0, 20)], REF34) == (False, [(0, 20)], 0)
  # adjacent matches merge and report a modification (SwipeAndMerge on a non-normalised record)
  assert nr.normalize_cigar(seq, 4, [(0, 5), (7, 15)], REF34) == (True, [(0, 20)], 0)
  # zero-length operations are removed
  assert nr.normalize_cigar(seq, 4, [(0, 5), (1, 0), (0, 15)], REF34) == (True, [(0, 20)], 0)


def test_heading_indel_alone_is_not_a_modification():
  # HandleHeadingIndel after the loop does not set is_modified (allelecounter.cc:843-844): the counter adds the rewritten
  # alignment, the read in memory keeps its own
  ref = b'ACGTACGTACGTACGTACGT'
  modified, cigar, shift = nr.normalize_cigar(b'GGACGT', 6, [(1, 2), (0, 4)], ref)
  assert (modified, cigar, shift) == (False, [(0, 2), (0, 4)], -2)
  modified, cigar, shift = nr.normalize_cigar(b'ACGT', 4, [(2, 4), (0, 4)], ref)
  assert (modified, cigar, shift) == (False, [(0, 4)], 4)


def _read(name, pos, seq, cigar, mapq=60):
  return Read(fragment_name=name, read_number=1, reference_name='chr1', position=pos, mapping_quality=mapq,
              cigar=parse_cigar_string(cigar), aligned_sequence=seq, aligned_quality=bytes([30] * len(seq)))


def test_normalize_region_reads():
  contig = REF151
  fetch = lambda a, b: contig[a:b]
  reads = [
      _read('del', 82, CASES[0][3], '31M12D26M'),
      _read('lowmq', 82, CASES[0][3], '31M12D26M', mapq=0),
      _read('plain', 10, contig[10:40], '30M'),
      _read('outside', 0, contig[0:20], '20M'),
      _read('heading', 60, b'GG' + contig[60:80], '2I20M'),
  ]
  pileup, count = nr.normalize_region_reads(reads, fetch, 30, 151, len(contig), 5)
  assert [r.fragment_name for r in pileup] == [r.fragment_name for r in reads]
  assert pileup[0].cigar == parse_cigar_string('16M12D41M') and pileup[0].position == 82
  assert reads[0].cigar == parse_cigar_string('31M12D26M')                 # inputs are not modified
  assert pileup[1] is reads[1] and pileup[2] is reads[2] and pileup[3] is reads[3]
  assert pileup[4] is reads[4]
  assert count is not None and count[4].cigar == [(0, 2), (0, 20)] and count[4].position == 58
  assert count[0] is pileup[0]
  pileup, count = nr.normalize_region_reads(reads[:4], fetch, 30, 151, len(contig), 5)
  assert count is None


def test_reads_interval():
  reads = [_read('a', 5, b'A' * 10, '10M'), _read('b', 95, b'A' * 10, '10M')]
  assert nr.reads_interval(reads, 20, 50, 100) == (5, 99)       # read end capped at n_bases - 1
  assert nr.reads_interval(reads, 20, 50, 1000) == (5, 105)
  assert nr.reads_interval([], 20, 50, 1000) == (20, 50)


def test_with_flags_golden_report_is_current():
  """tools/check_realigner_golden.py --with_flags: the reference's golden.calling_examples.with_flags (made with --min_mapping_quality 1
  --keep_legacy_allele_counter_behavior --normalize_reads) reproduced from BAM + FASTA, every tf.Example feature."""
  import json, os
  r = json.load(open(os.path.join(os.path.dirname(__file__), 'golden', 'with_flags_golden_report.json')))
  assert r['golden_examples'] == r['images_identical'] == r['tf_examples_equal_feature_by_feature'] == 84
  assert r['example_order_equal'] and r['reads_rewritten_by_normalization'] > 0
