"""Pins --variant_caller vcf_candidate_importer against the reference's golden.vcf_candidate_importer_calling_examples.tfrecord.gz
(make_examples_test.py:654-694: --mode calling --regions chr20:59,777,000-60,000,000 --norealign_reads --proposed_variants
input/vcf_candidate_importer.indels.chr20.vcf.gz, channels with insert_size): proposed VCF records -> native ComputeVariant
(dvb_candidates_from_proposed) -> pileups of 100 x 221 x 7 through the planner + CPU oracle; compares the variant protos (alleles,
AD / DP / VAF, genotype, sample) and the images.

Run in the build container (needs /root/reference).  Writes tests/golden/vcf_candidate_importer_golden_report.json and the fixture
tests/golden/vcf_candidate_importer_golden_subset.npz for the GPU suite."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import oracle_lib  # noqa: E402
from deepvariant_b200 import bam, candidates as cand, fasta, packing, protos, tfrecord  # noqa: E402
from deepvariant_b200 import make_examples_native as men  # noqa: E402
from deepvariant_b200 import pileup_image as pi  # noqa: E402
from deepvariant_b200 import vcf_candidate_importer as vci  # noqa: E402

T = '/root/reference/deepvariant/testdata/'


def variant_fields(raw):
  d = cand.canonical_call(protos.f_bytes(1, raw))
  return {k: d[k] for k in ('contig', 'start', 'end', 'ref', 'alts', 'genotype', 'call_set_name', 'info')}


def main():
  golden, order, golden_variant = {}, [], {}
  for r in tfrecord.read_records(T + 'golden.vcf_candidate_importer_calling_examples.tfrecord.gz'):
    e = protos.parse_tf_example(r)
    v = protos.parse_variant(e['variant/encoded'][1][0])
    idx = tuple(protos.parse_alt_allele_indices(e['alt_allele_indices/encoded'][1][0]))
    golden[(v.start, idx)] = np.frombuffer(e['image/encoded'][1][0], dtype=np.uint8).reshape(e['image/shape'][1])
    golden_variant[(v.start, idx)] = e['variant/encoded'][1][0]
    order.append((v.start, idx))
  bam_path = T + 'input/NA12878_S1.chr20.10_10p1mb.bam'
  ref = fasta.IndexedFastaReader(T + 'input/ucsc.hg19.chr20.unittest.fasta.gz')
  proposed = vci.ProposedVcfReader(T + 'input/vcf_candidate_importer.indels.chr20.vcf.gz')
  table = bam.NativeBamTable(bam_path, bam.ReadRequirements(min_mapping_quality=5))
  copts = cand.CandidateOptions(sample_name=cand.sample_name_from_bam(bam_path), min_mapping_quality=5)
  pic = pi.default_options(pi.ReadRequirements(min_base_quality=10, min_mapping_quality=5))
  pic.channels = list(pi.PILEUP_CHANNELS_WITH_INSERT_SIZE)
  pic.num_channels = 7
  gen = men.ExamplesGenerator(men.MakeExamplesOptions(pic_options=pic), test_mode=True, ref_reader=ref)
  params = pi.to_params(pic)
  refs = [(c, ref.n_bases(c)) for c in ref.contig_order]
  images, ours_order, variants, specs, n_regions = {}, [], {}, {}, 0
  for contig, s, e in cand.regions_to_process(refs, 1000, ('chr20', 59776999, 60000000)):
    if not vci.region_has_proposed_variant(proposed, contig, s, e):
      continue
    n_regions += 1
    rows = cand.region_reads(table, contig, s, e)
    found = vci.calls_from_vcf(table, ref, contig, s, e, copts, proposed, rows=rows)
    calls = found.calls()
    plans = gen.plan_region(calls, [table.read(int(i)) for i in rows], {})
    for p in plans:
      batch = packing.pack_images([p.spec], params)
      key = (p.variant.start, tuple(p.variant.alternate_bases.index(a) for a in p.alt_combination))
      images[key] = oracle_lib.encode_batch(params, batch)[0]
      variants[key] = p.variant.serialize()
      specs[key] = p.spec
      ours_order.append(key)
  full = [k for k in golden if k in images and np.array_equal(images[k], golden[k])]
  var_eq = [k for k in golden if k in variants and variant_fields(variants[k]) == variant_fields(golden_variant[k])]
  var_bytes = [k for k in golden if k in variants and variants[k] == golden_variant[k]]
  report = {'source': 'deepvariant/testdata/golden.vcf_candidate_importer_calling_examples.tfrecord.gz (v1.10.0), 100 x 221 x 7',
            'proposed_records': sum(len(v) for v in proposed.by_contig.values()), 'regions_with_a_proposed_record': n_regions,
            'golden_examples': len(golden), 'examples_planned': len(images), 'same_examples_in_same_order': ours_order == order,
            'images_identical': len(full), 'variants_identical_fields': len(var_eq), 'variants_identical_bytes': len(var_bytes),
            'not_identical': [dict(start=k[0], alt_indices=list(k[1]), planned=k in images,
                                   channels_differing=[c for c in range(7) if k in images and not np.array_equal(images[k][..., c], golden[k][..., c])],
                                   ours=variant_fields(variants[k]) if k in variants else None, golden=variant_fields(golden_variant[k]))
                              for k in golden if k not in full or k not in var_eq][:10]}
  with open(os.path.join(ROOT, 'tests/golden/vcf_candidate_importer_golden_report.json'), 'w') as f:
    json.dump(report, f, indent=1)
  print(json.dumps(report, indent=1))
  keep = [k for k in dict.fromkeys(order) if k in full]
  sub = packing.pack_images([specs[k] for k in keep], params)
  np.savez_compressed(os.path.join(ROOT, 'tests/golden/vcf_candidate_importer_golden_subset.npz'), n_images=sub.n_images, n_reads=sub.n_reads,
                      n_pairs=sub.n_pairs, ref_stride=sub.ref_stride, golden_images=np.stack([golden[k] for k in keep]),
                      keys=np.array([[k[0]] + list(k[1]) + [-1] * (2 - len(k[1])) for k in keep]), **{'arr_' + k: v for k, v in sub.arrays.items()})


def main_training():
  """The training-mode golden (make_examples_test.py:654-694, mode='training'): the proposed records are the TRUTH VCF (make_examples_core.py
  :1878-1881), uncalled genotypes skipped, realigner on (the test leaves the default), no --regions: 223 examples over 100 kb WITH read
  evidence.  Labels / truth genotypes are not built here (training is out of scope): images, alleles, AD / DP / VAF are compared."""
  from deepvariant_b200 import realigner
  golden, order, golden_variant = {}, [], {}
  for r in tfrecord.read_records(T + 'golden.vcf_candidate_importer.training_examples.tfrecord.gz'):
    e = protos.parse_tf_example(r)
    v = protos.parse_variant(e['variant/encoded'][1][0])
    idx = tuple(protos.parse_alt_allele_indices(e['alt_allele_indices/encoded'][1][0]))
    golden[(v.start, idx)] = np.frombuffer(e['image/encoded'][1][0], dtype=np.uint8).reshape(e['image/shape'][1])
    golden_variant[(v.start, idx)] = e['variant/encoded'][1][0]
    order.append((v.start, idx))
  bam_path = T + 'input/NA12878_S1.chr20.10_10p1mb.bam'
  ref = fasta.IndexedFastaReader(T + 'input/ucsc.hg19.chr20.unittest.fasta.gz')
  proposed = vci.ProposedVcfReader(T + 'input/test_nist.b37_chr20_100kbp_at_10mb.vcf.gz')
  reqs = bam.ReadRequirements(min_mapping_quality=5)
  table = bam.NativeBamTable(bam_path, reqs)
  copts = cand.CandidateOptions(sample_name=cand.sample_name_from_bam(bam_path), min_mapping_quality=5)
  pic = pi.default_options(pi.ReadRequirements(min_base_quality=10, min_mapping_quality=5))
  pic.channels = list(pi.PILEUP_CHANNELS_WITH_INSERT_SIZE)
  pic.num_channels = 7
  gen = men.ExamplesGenerator(men.MakeExamplesOptions(pic_options=pic), test_mode=True, ref_reader=ref)
  params = pi.to_params(pic)
  rl = realigner.Realigner(ref, realigner.RealignerOptions())
  refs = [(c, ref.n_bases(c)) for c in ref.contig_order]
  images, ours_order, variants, specs, n_regions = {}, [], {}, {}, 0
  for contig, s, e in cand.regions_to_process(refs, 1000, None):
    if not vci.region_has_proposed_variant(proposed, contig, s, e):
      continue
    n_regions += 1
    rows = cand.region_reads(table, contig, s, e)
    if len(rows):
      reads = rl.realign_reads(table, contig, rows, (s, e))
      t2 = bam.scratch_table(reads, refs, reqs)
    else:
      t2 = table
    rows2 = t2.query_indices(contig, s, e)
    found = vci.calls_from_vcf(t2, ref, contig, s, e, copts, proposed, rows=rows2, skip_uncalled_genotypes=True)
    for p in gen.plan_region(found.calls(), [t2.read(int(i)) for i in rows2], {}):
      key = (p.variant.start, tuple(p.variant.alternate_bases.index(a) for a in p.alt_combination))
      images[key] = oracle_lib.encode_batch(params, packing.pack_images([p.spec], params))[0]
      variants[key] = p.variant.serialize()
      specs[key] = p.spec
      ours_order.append(key)
    if t2 is not table:
      t2.close()

  def fields(raw):
    d = variant_fields(raw)
    d.pop('genotype', None)      # the labeler writes the truth genotype into the training example's variant
    return d
  full = [k for k in golden if k in images and np.array_equal(images[k], golden[k])]
  var_eq = [k for k in golden if k in variants and fields(variants[k]) == fields(golden_variant[k])]
  with_evidence = [k for k in golden if golden[k][5:].any()]
  report = {'source': 'deepvariant/testdata/golden.vcf_candidate_importer.training_examples.tfrecord.gz (v1.10.0), 100 x 221 x 7; proposed = the truth VCF',
            'proposed_records': sum(len(v) for v in proposed.by_contig.values()), 'regions_with_a_proposed_record': n_regions,
            'golden_examples': len(golden), 'examples_planned': len(images), 'same_examples_in_same_order': ours_order == order,
            'golden_images_with_read_rows': len(with_evidence), 'images_identical': len(full), 'variants_identical_fields_but_genotype': len(var_eq),
            'not_identical': [dict(start=k[0], alt_indices=list(k[1]), planned=k in images,
                                   channels_differing=[c for c in range(7) if k in images and not np.array_equal(images[k][..., c], golden[k][..., c])],
                                   ours=fields(variants[k]) if k in variants else None, golden=fields(golden_variant[k]))
                              for k in golden if k not in full or k not in var_eq][:10],
            'extra_examples': [list(map(str, k)) for k in images if k not in golden][:10]}
  path = os.path.join(ROOT, 'tests/golden/vcf_candidate_importer_golden_report.json')
  whole = json.load(open(path))
  whole['training_golden'] = report
  json.dump(whole, open(path, 'w'), indent=1)
  print(json.dumps(report, indent=1))
  # fixture: examples with evidence (multi-allelic and extended-reference ones first)
  multi = [k for k in full if k in var_eq and len(variant_fields(golden_variant[k])['alts']) > 1]
  keep = list(dict.fromkeys(multi[:6] + [k for k in full if k in var_eq and k in with_evidence][:6]))
  sub = packing.pack_images([specs[k] for k in keep], params)
  np.savez_compressed(os.path.join(ROOT, 'tests/golden/vcf_candidate_importer_training_subset.npz'), n_images=sub.n_images, n_reads=sub.n_reads,
                      n_pairs=sub.n_pairs, ref_stride=sub.ref_stride, golden_images=np.stack([golden[k] for k in keep]),
                      keys=np.array([[k[0]] + list(k[1]) + [-1] * (2 - len(k[1])) for k in keep]), **{'arr_' + k: v for k, v in sub.arrays.items()})


def main_cli():
  """--cli: the same golden through the make_examples stage CLI with the test's flags; records compared in order with the golden's
  (image bytes, shape, alt_allele_indices, locus, variant fields).  The encoder handle is the CPU oracle here (no GPU in the build
  container); everything else is the product flow."""
  import tempfile
  import test_candidates as tc
  from deepvariant_b200 import cli
  men.ExamplesGenerator._gpu = lambda self: tc.OracleEncoder(pi.to_params(self.options.pic_options, height=self.pileup_image_height))  # pylint: disable=protected-access
  with tempfile.TemporaryDirectory() as tmp:
    out = os.path.join(tmp, 'examples.tfrecord.gz')
    cli.make_examples(['--mode', 'calling', '--ref', T + 'input/ucsc.hg19.chr20.unittest.fasta.gz', '--reads', T + 'input/NA12878_S1.chr20.10_10p1mb.bam',
                       '--regions', 'chr20:59,777,000-60,000,000', '--norealign_reads', '--variant_caller', 'vcf_candidate_importer',
                       '--proposed_variants', T + 'input/vcf_candidate_importer.indels.chr20.vcf.gz', '--examples', out,
                       '--channel_list', 'read_base,base_quality,mapping_quality,strand,read_supports_variant,base_differs_from_ref,insert_size'])
    g = [protos.parse_tf_example(r) for r in tfrecord.read_records(T + 'golden.vcf_candidate_importer_calling_examples.tfrecord.gz')]
    o = [protos.parse_tf_example(r) for r in tfrecord.read_records(out)]
    keys = ('image/encoded', 'image/shape', 'alt_allele_indices/encoded', 'locus')
    equal = sum(1 for a, b in zip(g, o) if all(a[k] == b[k] for k in keys) and
                variant_fields(a['variant/encoded'][1][0]) == variant_fields(b['variant/encoded'][1][0]))
    info_equal = json.load(open(out + '.example_info.json')) == json.load(open(T + 'golden.vcf_candidate_importer_calling_examples.tfrecord.gz.example_info.json'))
  path = os.path.join(ROOT, 'tests/golden/vcf_candidate_importer_golden_report.json')
  report = json.load(open(path))
  report['stage_cli'] = {'golden_records': len(g), 'records_written': len(o), 'records_equal_in_order': equal, 'example_info_json_equal': info_equal}
  json.dump(report, open(path, 'w'), indent=1)
  print(report['stage_cli'])


if __name__ == '__main__':
  main()
  main_training()
  if '--cli' in sys.argv[1:]:
    main_cli()
