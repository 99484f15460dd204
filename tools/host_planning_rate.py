"""Host planning / packing rate of the make_examples tail on the reference's golden candidates (CPU only, one core):
Read-object planner vs numpy table path vs C++ region packer.  Needs /root/reference testdata (build container only)."""
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests'))
from deepvariant_b200 import bam, fasta, packing, protos, tfrecord  # noqa: E402
from test_bam_native import REF_INPUT, _wgs_generator  # noqa: E402


def main():
  td = os.path.dirname(REF_INPUT.rstrip('/')) + '/'
  cands = [protos.parse_deepvariant_call(r) for r in tfrecord.read_records(td + 'golden.calling_candidates.tfrecord.gz')]
  path = os.path.join(REF_INPUT, 'NA12878_S1.chr20.10_10p1mb.bam')
  req = bam.ReadRequirements(min_mapping_quality=5)
  t0 = time.perf_counter()
  table = bam.NativeBamTable(path, req)
  t_open = time.perf_counter() - t0
  reader = bam.BamReader(path, req)
  gen, params = _wgs_generator(fasta.IndexedFastaReader(os.path.join(REF_INPUT, 'ucsc.hg19.chr20.unittest.fasta.gz')))
  region_start, part = 9_999_999, 1000
  by_part = {}
  for c in cands:
    by_part.setdefault(region_start + (c.variant.start - region_start) // part * part, []).append(c)
  print(f'native BAM decode {t_open * 1e3:.1f} ms, {table.n_reads} reads')
  reps = 20
  for name in ('reads', 'numpy', 'native'):
    best = 1e9
    for _ in range(3):
      t = time.perf_counter()
      n = 0
      for _ in range(reps if name != 'reads' else 2):
        for p0, cs in sorted(by_part.items()):
          region = (cs[0].variant.reference_name, p0, min(p0 + part, 10_010_000))
          if name == 'reads':
            plans = gen.plan_region(cs, reader.query(*region), {})
            b = packing.pack_images([p.spec for p in plans], params)
          elif name == 'numpy':
            _, specs = gen.plan_region_from_table(cs, table, {}, region)
            b = packing.pack_images_from_table(specs, table, params)
          else:
            _, b = gen.pack_region_native(cs, table, region)
          n += b.n_images
      best = min(best, (time.perf_counter() - t) / n)
    print(f'{name:7s} {1 / best:9.0f} images/s/core')


if __name__ == '__main__':
  main()
