"""Builds the small inputs `tools/cli_throughput.py` runs the product CLI on when /root/reference is not there (the GPU box):
the reads of BASELINE config 1 (quick-start NA12878 chr20:10,000,000-10,010,000, widened by 1 kb) re-written through the
product's own BAM writer, and a FASTA of chr20 that is N outside [9,990,000, 10,020,000).  Inputs, not golden outputs: nothing is pinned
on them.  Run here (needs /root/reference):  python tools/make_quickstart_fixture.py"""
import gzip
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from deepvariant_b200 import bam, fasta  # noqa: E402

T = '/root/reference/deepvariant/testdata/input/'
OUT = os.path.join(ROOT, 'tests', 'golden')


def main():
  table = bam.NativeBamTable(T + 'NA12878_S1.chr20.10_10p1mb.bam', bam.ReadRequirements(min_mapping_quality=0, keep_duplicates=True, keep_failed_vendor_quality_checks=True,
                                                                                          keep_secondary_alignments=True, keep_supplementary_alignments=True,
                                                                                          keep_improperly_placed=True),
                             regions=[('chr20', 9_999_000, 10_011_000)])
  reads = [table.read(i) for i in range(table.n_reads)]
  ref = fasta.IndexedFastaReader(T + 'ucsc.hg19.chr20.unittest.fasta.gz')
  n = ref.n_bases('chr20')
  bam.write_bam(os.path.join(OUT, 'quickstart.chr20_10mb.bam'), reads, [('chr20', n)], sample_name='NA12878')
  lo, hi = 9_990_000, 10_020_000
  window = ref.query('chr20', lo, hi)
  path = os.path.join(OUT, 'quickstart.chr20_10mb.fa.gz')
  with gzip.open(path, 'wb', compresslevel=9) as f:
    f.write(b'>chr20\n')
    seq_iter = ('N' * lo, window, 'N' * (n - hi))
    buf = ''
    for part in seq_iter:
      buf += part
      k = len(buf) // 50 * 50
      f.write(('\n'.join(buf[i:i + 50] for i in range(0, k, 50)) + '\n').encode() if k else b'')
      buf = buf[k:]
    if buf:
      f.write((buf + '\n').encode())
  open(path + '.fai', 'w').write(f'chr20\t{n}\t7\t50\t51\n')
  print(len(reads), 'reads;', os.path.getsize(os.path.join(OUT, 'quickstart.chr20_10mb.bam')), 'bytes of BAM;', os.path.getsize(path), 'bytes of FASTA')


if __name__ == '__main__':
  main()
