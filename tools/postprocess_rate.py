"""Throughput of postprocess_variants (CVO -> VCF) on one host: the reference's golden PACBIO CallVariantsOutputs replicated along chr20
(every copy shifted by 100 kb, so the copies never overlap), serial against --cpus workers; the outputs must be identical.
Writes profiles/r02d_postprocess_rate.json."""
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from deepvariant_b200 import postprocess_variants as pp, protos, tfrecord  # noqa: E402


def shifted(record: bytes, delta: int) -> bytes:
  out = bytearray()
  for fn, wt, val, raw in protos.iter_fields(record):
    if fn != 1:
      out += raw
      continue
    v = bytearray()
    for f2, w2, v2, raw2 in protos.iter_fields(bytes(val)):
      v += protos.f_varint(f2, v2 + delta) if f2 in (13, 16) else raw2
    out += protos.f_bytes(1, bytes(v))
  return bytes(out)


def main():
  copies = int(sys.argv[1]) if len(sys.argv) > 1 else 300
  cpus = int(sys.argv[2]) if len(sys.argv) > 2 else min(os.cpu_count() or 1, 8)
  base = list(tfrecord.read_records(os.path.join(ROOT, 'tests/golden/golden.postprocess_pacbio_input-00000-of-00001.tfrecord.gz')))
  with tempfile.TemporaryDirectory() as tmp:
    path = os.path.join(tmp, 'cvo.tfrecord.gz')
    w = tfrecord.Writer(path)
    for k in range(copies):
      for r in base:
        w.write(shifted(r, 100000 * k))
    w.close()
    n = copies * len(base)
    contigs = [('chr20', 64444167 + 100000 * copies)]
    t0 = time.time()
    a = pp.postprocess_variants(path, os.path.join(tmp, 'serial.vcf'), contigs)
    t1 = time.time()
    b = pp.postprocess_variants(path, os.path.join(tmp, 'parallel.vcf'), contigs, cpus=cpus, chunk_records=max(2000, n // (8 * cpus)))
    t2 = time.time()
    same = open(os.path.join(tmp, 'serial.vcf')).read() == open(os.path.join(tmp, 'parallel.vcf')).read()
  out = {'cvo_records': n, 'variants_written': a['n_variants_written'], 'serial_s': round(t1 - t0, 2), 'serial_records_per_s': round(n / (t1 - t0)),
         'cpus': cpus, 'parallel_s': round(t2 - t1, 2), 'parallel_records_per_s': round(n / (t2 - t1)), 'outputs_identical': same and a == b,
         'host_cores': os.cpu_count()}
  print(json.dumps(out))
  json.dump(out, open(os.path.join(ROOT, 'profiles', 'r02d_postprocess_rate.json'), 'w'), indent=1)


if __name__ == '__main__':
  main()
