"""Derives tests/golden/ fixtures from the REFERENCE's own golden files (run in the build container,
where /root/reference exists; the fixtures travel, the reference does not).

Inputs (google/deepvariant v1.10.0, deepvariant/testdata):
  golden.calling_candidates.tfrecord.gz   78 DeepVariantCall protos (incl. allele_support read names)
  golden.calling_examples.tfrecord.gz     84 tf.Examples, 100x221x7, produced by the reference's make_examples
  input/NA12878_S1.chr20.10_10p1mb.bam    reads;  input/ucsc.hg19.chr20.unittest.fasta.gz  reference

For every candidate the reads of its 1-kb partition are fetched from the BAM with the reference's read
filter, the per-candidate image specs are planned by deepvariant_b200.make_examples_native (the product's
host logic) and encoded by the CPU ORACLE; the result is compared with the reference's golden image.
The golden images were made from *realigned* reads (realigner on, scripts/create_golden.sh:169-180), which
this repository does not rebuild, so only candidates whose overlapping reads the realigner left untouched
can match; those that match EXACTLY are saved as fixtures:
  tests/golden/wgs_golden_subset.npz   packed DvbBatch arrays + the reference's golden image bytes
  tests/golden/wgs_golden_report.json  per-example match report
so that tests can pin the oracle (CPU) and the CUDA encoder (GPU) against the reference's own output.
"""
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


def _fields(buf, nested=()):
  out = []
  for fn, wt, val, _ in protos.iter_fields(buf):
    v = val if isinstance(val, int) else bytes(val)
    if fn in nested and not isinstance(v, int):
      v = tuple(sorted(map(repr, _fields(v, (2,)))))   # map entries / sub-messages: order-insensitive
    out.append((fn, wt, v))
  return sorted(map(repr, out))


def variants_equal(a: bytes, b: bytes) -> bool:
  """Parsed-proto equality as in assertDeepVariantExamplesEqual (make_examples_test.py:1047-1066): map-field
  byte order inside the serialized Variant is not significant."""
  return _fields(a, (11,)) == _fields(b, (11,))


def main():
  cands = [protos.parse_deepvariant_call(r) for r in tfrecord.read_records(REF + 'golden.calling_candidates.tfrecord.gz')]
  examples = [protos.parse_tf_example(r) for r in tfrecord.read_records(REF + 'golden.calling_examples.tfrecord.gz')]
  print(len(cands), 'candidates', len(examples), 'golden examples')
  reader = bam.BamReader(REF + 'input/NA12878_S1.chr20.10_10p1mb.bam', bam.ReadRequirements(min_mapping_quality=5))
  ref = fasta.IndexedFastaReader(REF + 'input/ucsc.hg19.chr20.unittest.fasta.gz')
  pic = pi.default_options(pi.ReadRequirements(min_base_quality=10, min_mapping_quality=5))
  pic.channels = list(pi.PILEUP_CHANNELS_WITH_INSERT_SIZE)
  pic.num_channels = 7
  opts = men.MakeExamplesOptions(pic_options=pic)
  gen = men.ExamplesGenerator(opts, test_mode=True, ref_reader=ref)
  params = pi.to_params(pic)
  # region chr20:10,000,000-10,010,000 (1-based inclusive) -> [9999999, 10010000), 1000-bp partitions
  region_start, part = 9_999_999, 1000
  plans_all, golden_imgs, meta = [], [], []
  gi = 0
  for cand in cands:
    v = cand.variant
    p0 = region_start + (v.start - region_start) // part * part
    reads = reader.query(v.reference_name, p0, min(p0 + part, 10_010_000))
    stats = {}
    plans = gen.plan_region([cand], reads, stats)
    for pl in plans:
      ex = examples[gi]
      gi += 1
      assert protos.parse_variant(ex['variant/encoded'][1][0]).start == v.start, 'example order mismatch'
      enc_idx, want_idx = men.encode_alt_alleles(v, pl.alt_combination)
      assert ex['alt_allele_indices/encoded'][1][0] == enc_idx, (ex['alt_allele_indices/encoded'], enc_idx)
      assert ex['locus'][1][0].decode() == f'{v.reference_name}:{v.start + 1}-{v.end}'
      assert ex['variant_type'][1][0] == pl.variant_type
      assert variants_equal(ex['variant/encoded'][1][0], v.serialize())
      golden_imgs.append(np.frombuffer(ex['image/encoded'][1][0], dtype=np.uint8).reshape(100, 221, 7))
      plans_all.append(pl)
      meta.append(dict(start=v.start, ref=v.reference_bases, alts=v.alternate_bases, comb=pl.alt_combination,
                       n_reads=len(pl.spec.reads), vtype=pl.variant_type))
  assert gi == len(examples)
  batch = packing.pack_images([p.spec for p in plans_all], params)
  ours = oracle_lib.encode_batch(params, batch)
  golden = np.stack(golden_imgs)
  exact, report = [], []
  for i, m in enumerate(meta):
    diff = ours[i] != golden[i]
    rows_ours = int(ours[i, 5:].reshape(95, -1).any(1).sum())
    rows_gold = int(golden[i, 5:].reshape(95, -1).any(1).sum())
    ref_rows_equal = bool(np.array_equal(ours[i, :5], golden[i, :5]))
    import collections
    ours_rows = collections.Counter(ours[i, r].tobytes() for r in range(5, 5 + rows_ours))
    gold_rows = collections.Counter(golden[i, r].tobytes() for r in range(5, 5 + rows_gold))
    m['golden_rows_reproduced'] = int(sum((ours_rows & gold_rows).values()))
    m.update(exact=bool(not diff.any()), differing_bytes=int(diff.sum()), rows_ours=rows_ours, rows_golden=rows_gold,
             ref_band_equal=ref_rows_equal, differing_rows=int(diff.reshape(100, -1).any(1).sum()))
    report.append(m)
    if not diff.any():
      exact.append(i)
  print(f'{len(exact)} of {len(meta)} golden examples reproduced bit-exactly; reference band equal in '
        f'{sum(r["ref_band_equal"] for r in report)}; same row count in {sum(r["rows_ours"] == r["rows_golden"] for r in report)}')
  tot_gold = sum(r['rows_golden'] for r in report)
  tot_hit = sum(r['golden_rows_reproduced'] for r in report)
  print(f'read rows: {tot_hit} of {tot_gold} golden read rows reproduced byte-for-byte ({100.0 * tot_hit / tot_gold:.1f} %)')
  os.makedirs(OUT, exist_ok=True)
  sub = packing.pack_images([plans_all[i].spec for i in exact], params)
  np.savez_compressed(os.path.join(OUT, 'wgs_golden_subset.npz'), n_images=sub.n_images, n_reads=sub.n_reads,
                      n_pairs=sub.n_pairs, ref_stride=sub.ref_stride, golden_images=golden[exact],
                      example_index=np.array(exact), **{'arr_' + k: v for k, v in sub.arrays.items()})
  json.dump({'source': 'deepvariant/testdata/golden.calling_examples.tfrecord.gz (v1.10.0)', 'n_examples': len(meta),
             'n_exact': len(exact), 'golden_read_rows': tot_gold, 'golden_read_rows_reproduced': tot_hit,
             'note': 'differences are reads the reference realigner rewrote (e.g. 100M1S -> 101M): upstream of this path',
             'examples': report}, open(os.path.join(OUT, 'wgs_golden_report.json'), 'w'), indent=1)
  # a tiny slice of the golden tf.Example file itself, for the wire-format tests
  with tfrecord.Writer(os.path.join(OUT, 'golden.calling_examples.first3.tfrecord.gz')) as w:
    for i, r in enumerate(tfrecord.read_records(REF + 'golden.calling_examples.tfrecord.gz')):
      if i < 3:
        w.write(r)
  with tfrecord.Writer(os.path.join(OUT, 'golden.calling_candidates.first8.tfrecord.gz')) as w:
    for i, r in enumerate(tfrecord.read_records(REF + 'golden.calling_candidates.tfrecord.gz')):
      if i < 8:
        w.write(r)


if __name__ == '__main__':
  main()
