"""deepvariant_b200/allele_frequency.py against the known answers of the reference's deepvariant/allele_frequency_test.py
(:47-172 update_haplotype, 5 cases; :175-213 reference span; :216-415 find_matching_allele_frequency, 17 cases over the reference's
own 1-KB population VCF) and, end to end, against golden.allele_frequency_examples (tests/golden/allele_frequency_golden_report.json,
tools/check_allele_frequency_golden.py: 78 of 78 images on all 8 channels)."""
import json
import os

import pytest

from deepvariant_b200 import allele_frequency as af

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


class V:
  def __init__(self, start, ref, alts, contig='chr20'):
    self.reference_name, self.start, self.end, self.reference_bases, self.alternate_bases = contig, start, start + len(ref), ref, list(alts)


class SpanReader:
  """Answers the FASTA queries the KATs make from the committed spans (tools/make_allele_frequency_fixtures.py)."""

  def __init__(self):
    self.spans = json.load(open(os.path.join(GOLDEN, 'allele_frequency_ref_spans.json')))

  def n_bases(self, contig):
    return self.spans[f'{contig}:n']

  def query(self, contig, a, b):
    return self.spans[f'{contig}:{a}-{b}']


@pytest.mark.parametrize('start,ref,alt,hap,offset,want', [
    (60168, 'C', 'T', 'GCACCT', 60165, 'GCATCT'),
    (60284, 'ATTCCAG', 'AT', 'TTTCCATTCCAGTCCAT', 60279, 'TTTCCATTCCAT'),
    (60279, 'TTTCCA', 'TTTCCATTCCA', 'TTTCCATTCCAGTCCAT', 60279, 'TTTCCATTCCATTCCAGTCCAT'),
    (60284, 'ATTCCAG', 'AT', 'TTTCCATTCCAG', 60279, 'TTTCCAT'),
    (60279, 'TTTCCA', 'TTTCCATTCCA', 'TTTCCATTCCAG', 60279, 'TTTCCATTCCATTCCAG'),
])
def test_update_haplotype(start, ref, alt, hap, offset, want):
  v = V(start, ref, [alt])
  got = af.update_haplotype(v, hap, offset)
  assert got == [{'haplotype': want, 'alt': alt, 'variant': v}]
  with pytest.raises(ValueError):
    af.update_haplotype(v, hap, start + 1)


# (start, reference_bases, alternate_bases, expected) - allele_frequency_test.py:216-415
FIND_KATS = [
    (60168, 'C', ['T'], dict(C=0.9998, T=0.0002)),                                        # matched_snp_1
    (60285, 'TTCCAG', ['T'], dict(T=0.001198, TTCCAG=0.998802)),                          # matched_del_1
    (60284, 'ATTCCAG', ['A'], dict(A=0, ATTCCAG=1)),                                      # unmatched_del_1
    (60284, 'ATTCCAG', ['AT'], dict(AT=0.001198, ATTCCAG=0.998802)),                      # matched_del_2: diff representation
    (60150, 'C', ['T'], dict(C=1, T=0)),                                                  # unmatched_snp_1
    (60168, 'C', ['T', 'A'], dict(C=0.9998, T=0.0002, A=0)),                              # mixed_snp_1
    (60168, 'C', ['A'], dict(C=0.9998, A=0)),                                             # unmatched_snp_2: non-1 ref allele
    (60279, 'TTTCCA', ['T', 'TTTCCATTCCA'], dict(TTTCCA=0.999401, T=0.000399, TTTCCATTCCA=0.0002)),   # matched_mult_1
    (60279, 'TTTCCA', ['T', 'TATCCATTCCA'], dict(TTTCCA=0.999401, T=0.000399, TATCCATTCCA=0)),        # unmatched_mult_1
    (60295, 'TTCCAT', ['T'], dict(T=0.000399, TTCCAT=0.923922)),                          # matched_del_3: diff representation
    (60279, 'TTTCCA', ['T'], dict(TTTCCA=0.999401, T=0.000399)),                          # matched_del_4: multi-allelic cohort
    (9074790, 'CT', ['C', 'CTTT'], dict(C=0.167732, CTTT=0.215256, CT=0.442092)),         # matched_mult_2: left align
    (9074790, 'C', ['CTTT'], dict(CTTT=0.145367, C=0.442092)),                            # matched_ins_1: left align
    (9074790, 'CTT', ['CTTA'], dict(CTTA=0, CTT=0.442092)),                               # unmatched_ins_1: left align
    (61065, 'T', ['C'], dict(C=0.079872, T=0.919729)),                                    # matched_mnps_1
    (62022, 'G', ['C', 'T'], dict(G=0.996206, C=0.003594, T=0)),                          # matched_snp_2
]
# the reference gives the two left-align cases explicit ends (9074794, 9074793) longer than their reference_bases
ENDS = {(9074790, 'CT'): 9074794, (9074790, 'C'): 9074794, (9074790, 'CTT'): 9074793}


def _variant(start, ref, alts):
  v = V(start, ref, alts)
  v.end = ENDS.get((start, ref), v.end)
  return v


@pytest.mark.parametrize('start,ref,alts,want', FIND_KATS)
def test_find_matching_allele_frequency(start, ref, alts, want):
  pop = af.PopulationVcfReader(os.path.join(GOLDEN, 'allele_frequencies_vcf.vcf'))
  got = af.find_matching_allele_frequency(_variant(start, ref, alts), pop, SpanReader())
  assert set(got) == set(want)
  for k in want:
    assert got[k] == pytest.approx(want[k], abs=5e-7), (k, got)   # assertAlmostEqual: 7 places


def test_reference_span_covers_candidate_and_cohort_records():
  """test_get_ref_haplotype_and_offset: the span of a candidate and two overlapping cohort records."""
  assert SpanReader().query('chr20', 60279, 60291) == 'TTTCCATTCCAG'


def test_candidates_without_a_reader_get_zero_frequencies():
  from deepvariant_b200.protos import DeepVariantCall, Variant
  c = DeepVariantCall(variant=Variant(reference_name='chr20', start=5, end=6, reference_bases='A', alternate_bases=['C', 'G']))
  out = af.add_allele_frequencies_to_candidates([c], None, None)
  assert out[0].allele_frequency == {'A': 1.0, 'C': 0.0, 'G': 0.0}
  pop = af.PopulationVcfReader(os.path.join(GOLDEN, 'allele_frequencies_vcf.vcf'))
  assert pop.query('chrUn', 0, 10) == []


def test_golden_report_says_every_example_is_reproduced():
  r = json.load(open(os.path.join(GOLDEN, 'allele_frequency_golden_report.json')))
  assert r['golden_examples'] == 78 and r['images_identical_all_8_channels'] == 78 and r['same_examples_in_same_order']
  assert r['of_them_allele_frequency_channel_equal'] == r['golden_images_with_nonzero_allele_frequency_pixels'] == 8
  assert r['downsampled_examples_identical'] == r['downsampled_examples'] == 51
