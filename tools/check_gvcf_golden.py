"""Pins deepvariant_b200/gvcf.py against the reference's gVCF goldens.

make_examples side (`--gvcf`): the non-variant blocks of every 1-kb region of chr20:10,000,000-10,010,000, from BAM + FASTA through
region reads -> realigner -> allele counter (dvb_candidates_summary_counts) -> make_gvcfs, against
golden.postprocess_gvcf_input.tfrecord.gz (scripts/create_golden.sh:169-180), record by record and field by field (likelihoods as
doubles, bit for bit).  The PACBIO golden (golden.postprocess_pacbio_gvcf_input.tfrecord.gz: --norealign_reads --phase_reads
--track_ref_reads, padded regions) the same way.
postprocess side: golden CVO + golden blocks -> golden.postprocess_gvcf_output.g.vcf / ..._pacbio.g.vcf byte for byte.
Writes tests/golden/gvcf_golden_report.json."""
import json
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from deepvariant_b200 import bam, candidates as cand, fasta, gvcf, realigner, tfrecord  # noqa: E402
from deepvariant_b200 import postprocess_variants as pp  # noqa: E402

T = '/root/reference/deepvariant/testdata/'


def fields(v):
  return (v.reference_name, v.start, v.end, v.reference_bases, tuple(v.alternate_bases), tuple(v.genotype), tuple(v.genotype_likelihood), v.gq,
          tuple(sorted((k, tuple(x)) for k, x in v.info.items())), v.call_set_name)


def compare(ours, golden):
  g = [fields(gvcf.parse_variant_record(r)) for r in golden]
  o = [fields(v) for v in ours]
  round_trip = [fields(gvcf.parse_variant_record(gvcf.serialize_gvcf_record(v))) for v in ours] == o
  same = sum(1 for a, b in zip(o, g) if a == b)
  first = next(({'ours': str(a)[:300], 'golden': str(b)[:300]} for a, b in zip(o, g) if a != b), None)
  return {'golden_records': len(g), 'ours_records': len(o), 'identical_in_order': same, 'all_equal': o == g, 'serialize_parse_round_trip': round_trip,
          'first_difference': first}


def wgs_blocks(include_med_dp=False):
  bam_path = T + 'input/NA12878_S1.chr20.10_10p1mb.bam'
  ref = fasta.IndexedFastaReader(T + 'input/ucsc.hg19.chr20.unittest.fasta.gz')
  table = bam.NativeBamTable(bam_path, bam.ReadRequirements(min_mapping_quality=5))
  copts = cand.CandidateOptions(sample_name=cand.sample_name_from_bam(bam_path))
  rl = realigner.Realigner(ref)
  refs = [(c, ref.n_bases(c)) for c in ref.contig_order]
  go = gvcf.GvcfOptions(sample_name=copts.sample_name, include_med_dp=include_med_dp)
  conf = gvcf.ReferenceConfidence(go)
  out = []
  for contig, s, e in cand.regions_to_process(refs, 1000, ('chr20', 9999999, 10010000)):
    rows = cand.region_reads(table, contig, s, e)
    reads = rl.realign_reads(table, contig, rows, (s, e))
    t2 = bam.scratch_table(reads, refs, bam.ReadRequirements(min_mapping_quality=5))
    found = cand.candidates_in_region(t2, ref, contig, s, e, copts, rows=t2.query_indices(contig, s, e))
    out += list(gvcf.make_gvcfs(contig, s, ref.query(contig, s, e), found.summary_counts, go, conf))
    t2.close()
  return out


def pacbio_blocks():
  bam_path = T + 'input/test_pacbio.chr20_100kbp_at_9mb.bam'
  ref = fasta.IndexedFastaReader(T + 'input/grch38.chr20_and_21_10M.fa.gz')
  table = bam.NativeBamTable(bam_path, bam.ReadRequirements(min_mapping_quality=1), parse_aux=True)
  copts = cand.CandidateOptions(sample_name=cand.sample_name_from_bam(bam_path), min_mapping_quality=1, track_ref_reads=True,
                                vsc_min_fraction_indels=0.12, partition_size=25000, max_reads_per_partition=600)
  go = gvcf.GvcfOptions(sample_name=copts.sample_name)
  conf = gvcf.ReferenceConfidence(go)
  out = []
  for contig, s, e in cand.regions_to_process([(c, ref.n_bases(c)) for c in ref.contig_order], 25000, ('chr20', 8999999, 9100000)):
    rows = cand.region_reads(table, contig, s, e, copts.max_reads_per_partition, copts.random_seed)
    found = cand.candidates_in_region(table, ref, contig, s, e, copts, rows=rows, padding_pct=20)
    lo = s - found.interval[0]                       # summary_counts(left_padding, right_padding): the unpadded region only
    out += list(gvcf.make_gvcfs(contig, s, ref.query(contig, s, e), found.summary_counts[lo:lo + (e - s)], go, conf))
  return out


def main():
  report = {}
  golden = list(tfrecord.read_records(T + 'golden.postprocess_gvcf_input.tfrecord.gz'))
  report['make_examples_wgs'] = compare(wgs_blocks(), golden)
  report['make_examples_pacbio'] = compare(pacbio_blocks(), list(tfrecord.read_records(T + 'golden.postprocess_pacbio_gvcf_input.tfrecord.gz')))
  ref = fasta.IndexedFastaReader(T + 'input/ucsc.hg19.chr20.unittest.fasta.gz')
  contigs = [(c, ref.n_bases(c)) for c in ref.contig_order]
  post = {}
  with tempfile.TemporaryDirectory() as tmp:
    for name, cvo, nv, gold in [
        ('wgs', 'golden.postprocess_single_site_input-00000-of-00001.tfrecord.gz', 'golden.postprocess_gvcf_input.tfrecord.gz', 'golden.postprocess_gvcf_output.g.vcf'),
        ('pacbio', 'golden.postprocess_pacbio_input-00000-of-00001.tfrecord.gz', 'golden.postprocess_pacbio_gvcf_input.tfrecord.gz',
         'golden.postprocess_gvcf_output_pacbio.g.vcf')]:
      r = pp.postprocess_variants(T + cvo, os.path.join(tmp, 'o.vcf'), contigs, nonvariant_site_tfrecord_path=T + nv,
                                  gvcf_outfile=os.path.join(tmp, 'o.g.vcf'), base_at=lambda c, p: ref.query(c, p, p + 1))
      a, b = open(os.path.join(tmp, 'o.g.vcf')).read().splitlines(), open(T + gold).read().splitlines()
      post[name] = {'lines': len(b), 'byte_identical': a == b, 'records': r['n_gvcf_records_written']}
    # --include_med_dp end to end: blocks from BAM + FASTA with MED_DP -> golden.postprocess_gvcf_output.med_dp.g.vcf (its input
    # TFRecord is not kept in testdata, scripts/create_golden.sh:437-470)
    med = os.path.join(tmp, 'med.tfrecord.gz')
    with tfrecord.Writer(med) as w:
      for b in wgs_blocks(include_med_dp=True):
        w.write(gvcf.serialize_gvcf_record(b))
    pp.postprocess_variants(T + 'golden.postprocess_single_site_input-00000-of-00001.tfrecord.gz', os.path.join(tmp, 'o.vcf'), contigs,
                            nonvariant_site_tfrecord_path=med, gvcf_outfile=os.path.join(tmp, 'm.g.vcf'), base_at=lambda c, p: ref.query(c, p, p + 1))
    a, b = open(os.path.join(tmp, 'm.g.vcf')).read().splitlines(), open(T + 'golden.postprocess_gvcf_output.med_dp.g.vcf').read().splitlines()
    post['wgs_med_dp_from_bam'] = {'lines': len(b), 'byte_identical': a == b}
  report['postprocess'] = post
  with open(os.path.join(ROOT, 'tests/golden/gvcf_golden_report.json'), 'w') as f:
    json.dump(report, f, indent=1)
  print(json.dumps(report, indent=1))


if __name__ == '__main__':
  main()
