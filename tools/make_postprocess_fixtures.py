"""Copies the reference's postprocess_variants golden pairs (CVO TFRecord in, VCF out) into tests/golden/ so that
tests/test_postprocess.py runs where /root/reference is absent.  Run in the build container."""
import gzip
import os
import shutil

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
T = '/root/reference/deepvariant/testdata/'
OUT = os.path.join(ROOT, 'tests', 'golden')
PAIRS = [
    ('golden.postprocess_single_site_input-00000-of-00001.tfrecord.gz', 'golden.postprocess_single_site_output.vcf'),
    ('golden.postprocess_single_site_input-00000-of-00001.tfrecord.gz', 'golden.postprocess_single_site_output.pass_only.vcf'),
    ('golden.vcf_candidate_importer_postprocess_single_site_input-00000-of-00001.tfrecord.gz',
     'golden.vcf_candidate_importer_postprocess_single_site_output.vcf'),
    ('golden.postprocess_pacbio_input-00000-of-00001.tfrecord.gz', 'golden.postprocess_single_site_output_pacbio.vcf.gz'),
]
for src, vcf in PAIRS:
  shutil.copy(T + src, os.path.join(OUT, src))
  dst = os.path.join(OUT, vcf[:-3] if vcf.endswith('.gz') else vcf)
  with (gzip.open if vcf.endswith('.gz') else open)(T + vcf, 'rt') as f, open(dst, 'w') as g:
    g.write(f.read())
  print(src, '->', os.path.basename(dst))

# gVCF: make_examples --gvcf records (non-variant blocks) and the merged g.vcf postprocess_variants writes from them
for src in ['golden.postprocess_gvcf_input.tfrecord.gz', 'golden.postprocess_pacbio_gvcf_input.tfrecord.gz']:
  shutil.copy(T + src, os.path.join(OUT, src))
for vcf in ['golden.postprocess_gvcf_output.g.vcf', 'golden.postprocess_gvcf_output_pacbio.g.vcf', 'golden.haploid_chr20.postprocess_gvcf_output.g.vcf',
            'golden.haploid_chr20.postprocess_single_site_output.vcf']:      # the last two: --haploid_contigs chr20
  with open(T + vcf, 'rb') as f, gzip.open(os.path.join(OUT, vcf + '.gz'), 'wb') as g:
    g.write(f.read())
  print(vcf, '-> .gz')
