"""Fixture for tests/test_allele_frequency.py: the reference's tiny population VCF (deepvariant/testdata/input/allele_frequencies_vcf.vcf.gz,
1 KB, test data) and every reference span the 17 known answers of allele_frequency_test.py:216-415 query from
input/grch38.chr20_and_21_10M.fa.gz.  Run in the build container (needs /root/reference)."""
import gzip
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from deepvariant_b200 import allele_frequency as af, fasta  # noqa: E402
import test_allele_frequency as t  # noqa: E402

T = '/root/reference/deepvariant/testdata/input/'


class Recorder:
  def __init__(self, ref):
    self.ref, self.spans = ref, {}

  def n_bases(self, contig):
    self.spans[f'{contig}:n'] = self.ref.n_bases(contig)
    return self.spans[f'{contig}:n']

  def query(self, contig, a, b):
    self.spans[f'{contig}:{a}-{b}'] = self.ref.query(contig, a, b)
    return self.spans[f'{contig}:{a}-{b}']


def main():
  text = gzip.open(T + 'allele_frequencies_vcf.vcf.gz', 'rt').read()
  out = os.path.join(ROOT, 'tests', 'golden', 'allele_frequencies_vcf.vcf')
  open(out, 'w').write(text)
  rec = Recorder(fasta.IndexedFastaReader(T + 'grch38.chr20_and_21_10M.fa.gz'))
  pop = af.PopulationVcfReader(out)
  for start, ref, alts, _ in t.FIND_KATS:
    af.find_matching_allele_frequency(t._variant(start, ref, alts), pop, rec)
  rec.query('chr20', 60279, 60291)   # test_get_ref_haplotype_and_offset
  json.dump(rec.spans, open(os.path.join(ROOT, 'tests', 'golden', 'allele_frequency_ref_spans.json'), 'w'), indent=0)
  print(len(rec.spans), 'spans')


if __name__ == '__main__':
  main()
