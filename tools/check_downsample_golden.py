"""Pins the DOWN-SAMPLING path (DownsampleReadIndices: libstdc++ std::shuffle with a fresh mt19937_64, then the first 95
reads that encode, then the position sort) against the reference's own output.

The only golden file whose pileups hit the 95-row cap is deepvariant/testdata/golden.allele_frequency_examples.tfrecord.gz
(scripts/create_golden.sh:421-432: region chr20:61001-62000 of input/grch38_1k_subset_chr20_and_chr21.bam, 100x221x8 =
the 7 WGS channels + allele_frequency).  No candidates file ships for it, so read support (channel 4) cannot be
reconstructed; the other six per-read channels (read_base, base_quality, mapping_quality, strand, base_differs_from_ref,
insert_size) and the ROW ORDER can: they depend only on which reads the shuffle keeps.  The golden was made with the
realigner on, so only examples whose reads the realigner left untouched can match in full.

Run in the build container (needs /root/reference).  Writes tests/golden/downsample_golden_subset.npz (+ report)."""
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
CH = [0, 1, 2, 3, 5, 6]   # every per-read channel except read_supports_variant


def main():
  examples = [protos.parse_tf_example(r) for r in tfrecord.read_records(REF + 'golden.allele_frequency_examples.tfrecord.gz')]
  reader = bam.BamReader(REF + 'input/grch38_1k_subset_chr20_and_chr21.bam', bam.ReadRequirements(min_mapping_quality=5))
  ref = fasta.IndexedFastaReader(REF + 'input/grch38.chr20_and_21_10M.fa.gz')
  pic = pi.default_options(pi.ReadRequirements(min_base_quality=10, min_mapping_quality=5))
  pic.channels = list(pi.PILEUP_CHANNELS_WITH_INSERT_SIZE)
  pic.num_channels = 7
  gen = men.ExamplesGenerator(men.MakeExamplesOptions(pic_options=pic), test_mode=True, ref_reader=ref)
  params = pi.to_params(pic)
  region = ('chr20', 61000, 62000)
  reads = reader.query(*region)
  print(len(examples), 'golden examples;', len(reads), 'reads in', region)
  specs, golden, meta = [], [], []
  for ex in examples:
    v = protos.parse_variant(ex['variant/encoded'][1][0])
    idx = protos.parse_alt_allele_indices(ex['alt_allele_indices/encoded'][1][0])
    comb = [v.alternate_bases[i] for i in idx]
    cand = protos.DeepVariantCall(variant=v, allele_support={})
    plans = [p for p in gen.plan_region([cand], reads, {}) if p.alt_combination == comb]
    assert len(plans) == 1, (v.start, comb)
    specs.append(plans[0].spec)
    shape = ex['image/shape'][1]
    golden.append(np.frombuffer(ex['image/encoded'][1][0], dtype=np.uint8).reshape(shape)[..., :7])
    meta.append(dict(start=v.start, ref=v.reference_bases, alts=v.alternate_bases, comb=comb, n_reads=len(plans[0].spec.reads)))
  ours = oracle_lib.encode_batch(params, packing.pack_images(specs, params))
  golden = np.stack(golden)
  full, report = [], []
  for i, m in enumerate(meta):
    rows_g = int(golden[i, 5:].reshape(95, -1).any(1).sum())
    rows_o = int(ours[i, 5:].reshape(95, -1).any(1).sum())
    eq_rows = [bool(np.array_equal(ours[i, r][:, CH], golden[i, r][:, CH])) for r in range(100)]
    m.update(rows_golden=rows_g, rows_ours=rows_o, capped=bool(m['n_reads'] > 95), rows_equal_in_place=int(sum(eq_rows[5:5 + rows_g])),
             exact_six_channels=bool(all(eq_rows)))
    report.append(m)
    if all(eq_rows):
      full.append(i)
  capped = [r for r in report if r['capped']]
  capped_exact = [i for i in full if report[i]['capped']]
  print(f'{len(capped)} examples have more than 95 overlapping reads (down-sampled); '
        f'{len(capped_exact)} of them reproduce ALL 100 rows, in order, on the six channels; '
        f'{len(full)} of {len(report)} examples overall')
  tot = sum(r['rows_golden'] for r in capped)
  hit = sum(r['rows_equal_in_place'] for r in capped)
  print(f'down-sampled examples: {hit} of {tot} golden read rows equal IN PLACE ({100.0 * hit / max(tot, 1):.1f} %)')
  keep = capped_exact[:12]
  if keep:
    sub = packing.pack_images([specs[i] for i in keep], params)
    np.savez_compressed(os.path.join(OUT, 'downsample_golden_subset.npz'), n_images=sub.n_images, n_reads=sub.n_reads, n_pairs=sub.n_pairs,
                        ref_stride=sub.ref_stride, golden_images=golden[keep], example_index=np.array(keep), channels=np.array(CH),
                        **{'arr_' + k: v for k, v in sub.arrays.items()})
  json.dump({'source': 'deepvariant/testdata/golden.allele_frequency_examples.tfrecord.gz (v1.10.0), channels 0-6',
             'compared_channels': CH, 'n_examples': len(report), 'n_downsampled': len(capped), 'n_downsampled_exact': len(capped_exact),
             'downsampled_rows': tot, 'downsampled_rows_equal_in_place': hit, 'examples': report},
            open(os.path.join(OUT, 'downsample_golden_report.json'), 'w'), indent=1)


if __name__ == '__main__':
  main()
