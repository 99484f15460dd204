"""Pins the allele_frequency channel (and, with it, the whole 8-channel image incl. read support and down-sampling) against the
reference's golden.allele_frequency_examples.tfrecord.gz (scripts/create_golden.sh:421-432: make_examples --mode calling --regions
chr20:61001-62000 --population_vcfs cohort-chr20_and_chr21_100k.vcf.gz --channel_list '...,insert_size,allele_frequency' over
input/grch38_1k_subset_chr20_and_chr21.bam, realigner on): region reads -> realigner -> candidates -> population allele frequencies
(deepvariant_b200/allele_frequency.py) -> pileups of 100 x 221 x 8 through the planner + channel planes + CPU oracle.

Run in the build container (needs /root/reference).  Writes tests/golden/allele_frequency_golden_report.json and a small fixture
tests/golden/allele_frequency_golden_subset.npz (packed batches + golden images of a few examples) for the GPU suite."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import oracle_lib  # noqa: E402
from deepvariant_b200 import allele_frequency as af  # noqa: E402
from deepvariant_b200 import bam, candidates as cand, fasta, packing, protos, realigner, tfrecord  # noqa: E402
from deepvariant_b200 import make_examples_native as men  # noqa: E402
from deepvariant_b200 import pileup_image as pi  # noqa: E402

T = '/root/reference/deepvariant/testdata/'


def main():
  golden, order = {}, []
  for r in tfrecord.read_records(T + 'golden.allele_frequency_examples.tfrecord.gz'):
    e = protos.parse_tf_example(r)
    v = protos.parse_variant(e['variant/encoded'][1][0])
    idx = tuple(protos.parse_alt_allele_indices(e['alt_allele_indices/encoded'][1][0]))
    golden[(v.start, idx)] = np.frombuffer(e['image/encoded'][1][0], dtype=np.uint8).reshape(e['image/shape'][1])
    order.append((v.start, idx))
  bam_path = T + 'input/grch38_1k_subset_chr20_and_chr21.bam'
  ref = fasta.IndexedFastaReader(T + 'input/grch38.chr20_and_21_10M.fa.gz')
  pop = af.PopulationVcfReader(T + 'input/cohort-chr20_and_chr21_100k.vcf.gz')
  table = bam.NativeBamTable(bam_path, bam.ReadRequirements(min_mapping_quality=5))
  copts = cand.CandidateOptions(sample_name=cand.sample_name_from_bam(bam_path), small_model_vaf_context_window_size=51, min_mapping_quality=5)
  pic = pi.default_options(pi.ReadRequirements(min_base_quality=10, min_mapping_quality=5))
  pic.channels = list(pi.PILEUP_CHANNELS_WITH_INSERT_SIZE) + ['allele_frequency']
  pic.num_channels = 8
  gen = men.ExamplesGenerator(men.MakeExamplesOptions(pic_options=pic), test_mode=True, ref_reader=ref)
  params = pi.to_params(pic)
  rl = realigner.Realigner(ref, realigner.RealignerOptions())
  refs = [(c, ref.n_bases(c)) for c in ref.contig_order]
  images, ours_order, packed_of, n_with_af = {}, [], {}, 0
  for contig, s, e in cand.regions_to_process(refs, 1000, ('chr20', 61000, 62000)):
    rows = cand.region_reads(table, contig, s, e)
    reads = rl.realign_reads(table, contig, rows, (s, e))
    t2 = bam.scratch_table(reads, refs, bam.ReadRequirements(min_mapping_quality=5))
    rows2 = t2.query_indices(contig, s, e)
    found = cand.candidates_in_region(t2, ref, contig, s, e, copts, rows=rows2)
    calls = af.add_allele_frequencies_to_candidates(found.calls(), pop, ref)
    n_with_af += sum(1 for c in calls if any(c.allele_frequency.get(a, 0) > 0 for a in c.variant.alternate_bases))
    plans = gen.plan_region(calls, [t2.read(int(i)) for i in rows2], {})
    for p in plans:
      batch = packing.pack_images([p.spec], params)
      key = (p.variant.start, tuple(p.variant.alternate_bases.index(a) for a in p.alt_combination))
      images[key] = oracle_lib.encode_batch(params, batch)[0]
      packed_of[key] = p.spec
      ours_order.append(key)
    t2.close()
  full = [k for k in golden if k in images and np.array_equal(images[k], golden[k])]
  ch_eq = [sum(1 for k in golden if k in images and np.array_equal(images[k][..., c], golden[k][..., c])) for c in range(8)]
  af_nonzero_golden = [k for k in golden if golden[k][5:, :, 7].any()]
  af_nonzero_equal = [k for k in af_nonzero_golden if k in images and np.array_equal(images[k][..., 7], golden[k][..., 7])]
  capped = [k for k in golden if k in packed_of and len(packed_of[k].reads) > 95]
  report = {'source': 'deepvariant/testdata/golden.allele_frequency_examples.tfrecord.gz (v1.10.0), 100 x 221 x 8',
            'golden_examples': len(golden), 'examples_planned': len(images), 'same_examples_in_same_order': ours_order == order,
            'images_identical_all_8_channels': len(full), 'images_identical_per_channel': ch_eq,
            'golden_images_with_nonzero_allele_frequency_pixels': len(af_nonzero_golden), 'of_them_allele_frequency_channel_equal': len(af_nonzero_equal),
            'allele_frequency_pixel_values_in_golden': sorted(int(v) for v in set(np.concatenate([golden[k][..., 7].ravel() for k in golden]).tolist())),
            'candidates_with_a_population_frequency': n_with_af, 'downsampled_examples': len(capped),
            'downsampled_examples_identical': sum(1 for k in capped if k in full),
            'not_identical': [dict(start=k[0], alt_indices=list(k), channels_differing=[c for c in range(8) if not np.array_equal(images[k][..., c], golden[k][..., c])])
                              for k in golden if k in images and k not in full][:20]}
  with open(os.path.join(ROOT, 'tests/golden/allele_frequency_golden_report.json'), 'w') as f:
    json.dump(report, f, indent=1)
  print(json.dumps(report, indent=1))
  # fixture for the GPU suite: a few identical examples, preferring ones with allele-frequency pixels and down-sampled ones
  keep = [k for k in full if k in af_nonzero_golden][:6] + [k for k in full if k in capped and k not in af_nonzero_golden][:3]
  if keep:
    sub = packing.pack_images([packed_of[k] for k in keep], params)
    np.savez_compressed(os.path.join(ROOT, 'tests/golden/allele_frequency_golden_subset.npz'), n_images=sub.n_images, n_reads=sub.n_reads, n_pairs=sub.n_pairs,
                        ref_stride=sub.ref_stride, golden_images=np.stack([golden[k] for k in keep]), keys=np.array([[k[0]] + list(k[1]) + [-1] * (2 - len(k[1])) for k in keep]),
                        **{'arr_' + k: v for k, v in sub.arrays.items()})


def main_cli():
  """--cli: the same golden through the make_examples stage CLI (flags of scripts/create_golden.sh:424-432), records compared in order
  with the golden's (image bytes, shape, alt_allele_indices, locus) and the example_info.json beside them.  The encoder handle is the
  CPU oracle here (no GPU in the build container); everything else is the product flow."""
  import tempfile
  import test_candidates as tc
  from deepvariant_b200 import cli
  men.ExamplesGenerator._gpu = lambda self: tc.OracleEncoder(pi.to_params(self.options.pic_options, height=self.pileup_image_height))  # pylint: disable=protected-access
  with tempfile.TemporaryDirectory() as tmp:
    out = os.path.join(tmp, 'examples.tfrecord.gz')
    cli.make_examples(['--mode', 'calling', '--ref', T + 'input/grch38.chr20_and_21_10M.fa.gz', '--regions', 'chr20:61001-62000',
                       '--population_vcfs', T + 'input/cohort-chr20_and_chr21_100k.vcf.gz', '--reads', T + 'input/grch38_1k_subset_chr20_and_chr21.bam',
                       '--examples', out, '--channel_list',
                       'read_base,base_quality,mapping_quality,strand,read_supports_variant,base_differs_from_ref,insert_size,allele_frequency'])
    g = [protos.parse_tf_example(r) for r in tfrecord.read_records(T + 'golden.allele_frequency_examples.tfrecord.gz')]
    o = [protos.parse_tf_example(r) for r in tfrecord.read_records(out)]
    keys = ('image/encoded', 'image/shape', 'alt_allele_indices/encoded', 'locus')
    equal = sum(1 for a, b in zip(g, o) if all(a[k] == b[k] for k in keys))
    info_equal = json.load(open(out + '.example_info.json')) == json.load(open(T + 'golden.allele_frequency_examples.tfrecord.gz.example_info.json'))
  path = os.path.join(ROOT, 'tests/golden/allele_frequency_golden_report.json')
  report = json.load(open(path))
  report['stage_cli'] = {'golden_records': len(g), 'records_written': len(o), 'records_equal_in_order': equal, 'example_info_json_equal': info_equal}
  json.dump(report, open(path, 'w'), indent=1)
  print(report['stage_cli'])


if __name__ == '__main__':
  main()
  if '--cli' in sys.argv[1:]:
    main_cli()
