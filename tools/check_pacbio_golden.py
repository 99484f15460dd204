"""Cross-checks the TRIMMED-read path (TrimReads / TrimCigar, alt_aligned_pileup_lib.cc:91-276; long reads clipped to the
window) against the reference's PacBio golden examples.

deepvariant/testdata/golden.pacbio_examples.tfrecord.gz (make_examples_test.py:790-831): 100x147x10, channels
[read_base, base_quality, mapping_quality, strand, read_supports_variant, base_differs_from_ref, haplotype,
base_methylation, alt-aligned x2] from input/test_pacbio.chr20_100kbp_at_9mb.bam with --trim_reads_for_pileup,
--phase_reads, --sort_by_haplotypes, realigner off.  Phasing (upstream of this path) decides the haplotype channel and the
ROW ORDER, no candidates file ships (read support unknown) and base_methylation / alt-aligned channels are not built here,
so the comparison is: per example, the MULTISET of read rows on the five channels that depend only on the trimmed read
(read_base, base_quality, mapping_quality, strand, base_differs_from_ref), plus the reference band.

Run in the build container (needs /root/reference).  Writes tests/golden/pacbio_golden_report.json and a 12-example
fixture tests/golden/pacbio_golden_subset.npz (packed trimmed reads + the golden images' seven channels)."""
import collections
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import oracle_lib  # noqa: E402
from deepvariant_b200 import bam, fasta, packing, protos, tfrecord  # noqa: E402
from deepvariant_b200 import make_examples_native as men  # noqa: E402
from deepvariant_b200 import pileup_image as pi  # noqa: E402

REF = '/root/reference/deepvariant/testdata/'
OUT = os.path.join(ROOT, 'tests', 'golden')
CH = [0, 1, 2, 3, 5]


def main():
  examples = [protos.parse_tf_example(r) for r in tfrecord.read_records(REF + 'golden.pacbio_examples.tfrecord.gz')]
  reader = bam.BamReader(REF + 'input/test_pacbio.chr20_100kbp_at_9mb.bam', bam.ReadRequirements(min_mapping_quality=1), parse_aux=True)
  ref = fasta.IndexedFastaReader(REF + 'input/grch38.chr20_and_21_10M.fa.gz')
  pic = pi.default_options(pi.ReadRequirements(min_base_quality=10, min_mapping_quality=1))
  pic.channels = pi.PILEUP_DEFAULT_CHANNELS + ['haplotype']
  pic.num_channels = 7
  pic.width = 147
  gen = men.ExamplesGenerator(men.MakeExamplesOptions(pic_options=pic, trim_reads_for_pileup=True), test_mode=True, ref_reader=ref)
  params = pi.to_params(pic)
  region_start, part = 8_999_999, 25000
  specs, golden, meta = [], [], []
  cache = {}
  for ex in examples:
    v = protos.parse_variant(ex['variant/encoded'][1][0])
    idx = protos.parse_alt_allele_indices(ex['alt_allele_indices/encoded'][1][0])
    comb = [v.alternate_bases[i] for i in idx]
    p0 = region_start + (v.start - region_start) // part * part
    if p0 not in cache:
      cache[p0] = reader.query(v.reference_name, p0, min(p0 + part, 9_100_000))
    cand = protos.DeepVariantCall(variant=v, allele_support={})
    plans = [p for p in gen.plan_region([cand], cache[p0], {}) if p.alt_combination == comb]
    if len(plans) != 1:
      continue
    specs.append(plans[0].spec)
    golden.append(np.frombuffer(ex['image/encoded'][1][0], dtype=np.uint8).reshape(ex['image/shape'][1])[..., :7])
    meta.append(dict(start=v.start, ref=v.reference_bases, alts=v.alternate_bases, vtype=plans[0].variant_type, n_reads=len(plans[0].spec.reads)))
  print(len(examples), 'golden examples;', len(specs), 'planned')
  ours = oracle_lib.encode_batch(params, packing.pack_images(specs, params))
  golden = np.stack(golden)
  tot_rows = hit_rows = full = band = same_count = 0
  for i, m in enumerate(meta):
    g_rows = [golden[i, r][:, CH].tobytes() for r in range(5, 100) if golden[i, r].any()]
    o_rows = [ours[i, r][:, CH].tobytes() for r in range(5, 100) if ours[i, r].any()]
    inter = sum((collections.Counter(g_rows) & collections.Counter(o_rows)).values())
    m.update(rows_golden=len(g_rows), rows_ours=len(o_rows), rows_matched=int(inter),
             ref_band_equal=bool(np.array_equal(ours[i, :5][..., CH], golden[i, :5][..., CH])))
    tot_rows += len(g_rows); hit_rows += inter
    full += int(inter == len(g_rows) == len(o_rows)); band += int(m['ref_band_equal']); same_count += int(len(g_rows) == len(o_rows))
  print(f'reference band equal in {band} of {len(meta)}; same number of read rows in {same_count}; '
        f'every read row matched (as a multiset, 5 channels) in {full}; rows: {hit_rows} of {tot_rows} ({100.0 * hit_rows / tot_rows:.1f} %)')
  # fixture: the deepest examples of each variant type (SNP / indel) — multiset compare needs no phasing
  order = sorted(range(len(meta)), key=lambda i: (-meta[i]['n_reads'], i))
  keep = sorted([i for i in order if meta[i]['vtype'] == 1][:6] + [i for i in order if meta[i]['vtype'] == 2][:6])
  sub = packing.pack_images([specs[i] for i in keep], params)
  np.savez_compressed(os.path.join(OUT, 'pacbio_golden_subset.npz'), n_images=sub.n_images, n_reads=sub.n_reads, n_pairs=sub.n_pairs,
                      ref_stride=sub.ref_stride, golden_images=golden[keep], example_index=np.array(keep), channels=np.array(CH),
                      **{'arr_' + k: v for k, v in sub.arrays.items()})
  json.dump({'source': 'deepvariant/testdata/golden.pacbio_examples.tfrecord.gz (v1.10.0), channels ' + str(CH) + ' as row multisets',
             'n_examples': len(meta), 'ref_band_equal': band, 'same_row_count': same_count, 'all_rows_matched': full,
             'golden_read_rows': tot_rows, 'golden_read_rows_matched': hit_rows, 'examples': meta},
            open(os.path.join(OUT, 'pacbio_golden_report.json'), 'w'), indent=1)


if __name__ == '__main__':
  main()
