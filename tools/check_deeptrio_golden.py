"""Pins the multi-sample image (FillPileupArrayBySample, deepvariant/pileup_image_native.h:313-336; CreateAndWriteExamplesForCandidate
with three samples, make_examples_native.cc:632-736) against the reference's DeepTrio golden: deeptrio/testdata/
golden_child.calling_examples.tfrecord.gz (scripts/create_golden_deeptrio.sh:142-160: --reads HG001 --reads_parent1 NA12891
--reads_parent2 NA12892, region 20:10,000,000-10,010,000, BASE_CHANNELS + insert_size, realigner on) - images of 140 x 221 x 7 =
parent1 | child | parent2 blocks of 40 | 60 | 40 rows (--pileup_image_height_parent 40 --pileup_image_height_child 60, the heights of
the released DeepTrio WGS models, so every block of a 30x sample is down-sampled).  The candidates are the golden's own (golden_child.calling_candidates.tfrecord.gz: the
trio caller's cross-sample allele filters are out of scope); per partition every sample's reads are realigned on their own
(realign_reads_per_sample_multisample, make_examples_core.py:2520-2538) and the three pileups come from the same DeepVariantCall.

Run in the build container (needs /root/reference).  Writes tests/golden/deeptrio_golden_report.json and a small fixture
(tests/golden/deeptrio_golden_subset.npz) for the oracle / CUDA tests."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import oracle_lib  # noqa: E402
from deepvariant_b200 import bam, candidates as cand, fasta, packing, protos, realigner, tfrecord  # noqa: E402
from deepvariant_b200 import make_examples_native as men  # noqa: E402
from deepvariant_b200 import pileup_image as pi  # noqa: E402

T = '/root/reference/deeptrio/testdata/'


class OracleEncoder:
  def __init__(self, params):
    self.params = params
    self.shape = (params.height, params.width, params.num_channels + params.num_alt_channels)

  def encode_host(self, batch):
    return oracle_lib.encode_batch(self.params, batch)


def main():
  golden, order = {}, []
  for r in tfrecord.read_records(T + 'golden_child.calling_examples.tfrecord.gz'):
    e = protos.parse_tf_example(r)
    v = protos.parse_variant(e['variant/encoded'][1][0])
    idx = tuple(protos.parse_alt_allele_indices(e['alt_allele_indices/encoded'][1][0]))
    golden[(v.start, idx)] = np.frombuffer(e['image/encoded'][1][0], dtype=np.uint8).reshape(e['image/shape'][1])
    order.append((v.start, idx))
  calls = [protos.parse_deepvariant_call(r) for r in tfrecord.read_records(T + 'golden_child.calling_candidates.tfrecord.gz')]
  ref = fasta.IndexedFastaReader(T + 'input/hs37d5.chr20.fa.gz')
  # samples_in_order = [parent1, child, parent2] (deeptrio/make_examples.py:316-319); the child's order is [0, 1, 2]
  bams = ['NA12891.chr20.10_10p1mb_sorted.bam', 'HG001.chr20.10_10p1mb_sorted.bam', 'NA12892.chr20.10_10p1mb_sorted.bam']
  tables = [bam.NativeBamTable(T + 'input/' + b, bam.ReadRequirements(min_mapping_quality=5)) for b in bams]
  pic = pi.default_options(pi.ReadRequirements(min_base_quality=10, min_mapping_quality=5))
  pic.channels = list(pi.PILEUP_CHANNELS_WITH_INSERT_SIZE)
  pic.num_channels = 7
  samples = [men.SampleOptions(role='parent1', name='parent1', pileup_height=40, order=[0, 1, 2]),
             men.SampleOptions(role='child', name='child', pileup_height=60, order=[0, 1, 2]),
             men.SampleOptions(role='parent2', name='parent2', pileup_height=40, order=[2, 1, 0])]
  gen = men.ExamplesGenerator(men.MakeExamplesOptions(pic_options=pic, sample_options=samples), test_mode=True, ref_reader=ref)
  gen._make_encoder = OracleEncoder   # pylint: disable=protected-access  (no GPU in the build container; the GPU suite runs the fixture through the CUDA encoder)
  assert gen.image_shape() == [140, 221, 7]
  bounds = [(0, 40), (40, 100), (100, 140)]
  rl = realigner.Realigner(ref, realigner.RealignerOptions())
  refs = [(c, ref.n_bases(c)) for c in ref.contig_order]
  images, ours_order, spec_of = {}, [], {}
  for contig, s, e in cand.regions_to_process(refs, 1000, ('20', 9999999, 10010000)):
    region_calls = [c for c in calls if c.variant.reference_name == contig and s <= c.variant.start < e]
    if not region_calls:
      continue
    reads_per_sample = []
    for t in tables:
      rows = cand.region_reads(t, contig, s, e)
      reads_per_sample.append(rl.realign_reads(t, contig, rows, (s, e)))
    plans, specs_of = gen.plan_region_by_sample(region_calls, reads_per_sample, samples[1].order, {})
    imgs = gen.encode_plans_by_sample(specs_of)
    for p, per_sample, img in zip(plans, specs_of, imgs):
      key = (p.variant.start, tuple(p.variant.alternate_bases.index(a) for a in p.alt_combination))
      images[key] = img
      spec_of[key] = per_sample
      ours_order.append(key)
  full = [k for k in golden if k in images and np.array_equal(images[k], golden[k])]
  blocks = [sum(1 for k in golden if k in images and np.array_equal(images[k][bounds[b][0]:bounds[b][1]], golden[k][bounds[b][0]:bounds[b][1]])) for b in range(3)]
  report = {'source': 'deeptrio/testdata/golden_child.calling_examples.tfrecord.gz (v1.10.0), 140 x 221 x 7 = parent1 (40) | child (60) | parent2 (40)',
            'golden_examples': len(golden), 'examples_planned': len(images), 'same_examples_in_same_order': ours_order == order,
            'images_identical': len(full), 'blocks_identical_parent1_child_parent2': blocks,
            'not_identical': [dict(start=k[0], alt_indices=list(k[1]), blocks_differing=[b for b in range(3) if not np.array_equal(images[k][bounds[b][0]:bounds[b][1]], golden[k][bounds[b][0]:bounds[b][1]])])
                              for k in golden if k in images and k not in full][:20]}
  with open(os.path.join(ROOT, 'tests/golden/deeptrio_golden_report.json'), 'w') as f:
    json.dump(report, f, indent=1)
  print(json.dumps(report, indent=1))
  keep = full[:2] + [k for k in full if len(k[1]) == 2][:1]
  if keep:
    arrays = {}
    for b in range(3):
      sub = packing.pack_images([spec_of[k][b][1] for k in keep], pi.to_params(pic, height=bounds[b][1] - bounds[b][0]))
      arrays.update({f's{b}_n': np.array([sub.n_images, sub.n_reads, sub.n_pairs, sub.ref_stride])})
      arrays.update({f's{b}_arr_' + k: v for k, v in sub.arrays.items()})
    np.savez_compressed(os.path.join(ROOT, 'tests/golden/deeptrio_golden_subset.npz'), golden_images=np.stack([golden[k] for k in keep]), **arrays)


if __name__ == '__main__':
  main()
