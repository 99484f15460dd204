"""Pins candidate generation (csrc/dvb_candidates.cu) against the reference's golden candidates.

Runs where /root/reference exists.  Re-derives golden.calling_candidates.tfrecord.gz (78 DeepVariantCalls; make_examples
--regions chr20:10,000,000-10,010,000 on NA12878_S1.chr20.10_10p1mb.bam, scripts/create_golden.sh:165-176) from the raw BAM
with our allele counter + caller and compares, per candidate: reference / alternate bases, AD / DP / VAF, allele_support as sets
of read keys, allele_support_ext (mapping quality, average base quality, strand, low-quality flag per read) and
allele_frequency_at_position.  The golden was made with the realigner on, so candidates whose reads the realigner rewrote are
expected to differ; the report lists them.  Writes tests/golden/candidates_golden_report.json and a fixture of the candidates
that match (tests/golden/candidates_golden_subset.json) used by tests/test_candidates.py where the reference is absent.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from deepvariant_b200 import bam, candidates as cand, fasta, protos, tfrecord  # noqa: E402

TESTDATA = '/root/reference/deepvariant/testdata'


def _value(buf):
  for fn, wt, val, _ in protos.iter_fields(buf):
    if fn == 7:
      return protos._to_signed32(val)   # pylint: disable=protected-access
    if fn == 2:
      import struct
      return struct.unpack('<d', struct.pack('<Q', val))[0] if isinstance(val, int) else struct.unpack('<d', bytes(val))[0]
    if fn == 3:
      return bytes(val).decode()
  return None


def _read_support(buf):
  d = {'read_name': '', 'is_low_quality': 0, 'mapping_quality': 0, 'average_base_quality': 0, 'is_reverse_strand': 0, 'sample_name': ''}
  names = {1: 'read_name', 2: 'is_low_quality', 3: 'mapping_quality', 4: 'average_base_quality', 5: 'is_reverse_strand', 7: 'sample_name'}
  for fn, wt, val, _ in protos.iter_fields(buf):
    if fn in names:
      d[names[fn]] = bytes(val).decode() if wt == 2 else int(val)
  return d


def canonical(record: bytes) -> dict:
  """Semantic content of a DeepVariantCall, independent of map / field order."""
  out = {'ref': '', 'alts': [], 'start': 0, 'end': 0, 'contig': '', 'info': {}, 'call_set_name': '', 'genotype': [],
         'allele_support': {}, 'allele_support_ext': {}, 'ref_support': [], 'ref_support_ext': [], 'af_at_position': {}}
  for fn, wt, val, _ in protos.iter_fields(record):
    val = bytes(val) if wt == 2 else val
    if fn == 1:
      for f2, w2, v2, _ in protos.iter_fields(val):
        v2 = bytes(v2) if w2 == 2 else v2
        if f2 == 6:
          out['ref'] = v2.decode()
        elif f2 == 7:
          out['alts'].append(v2.decode())
        elif f2 == 13:
          out['end'] = int(v2)
        elif f2 == 14:
          out['contig'] = v2.decode()
        elif f2 == 16:
          out['start'] = int(v2)
        elif f2 == 11:
          for f3, w3, v3, _ in protos.iter_fields(v2):
            v3 = bytes(v3) if w3 == 2 else v3
            if f3 == 2:
              key, vals = '', []
              for f4, w4, v4, _ in protos.iter_fields(v3):
                if f4 == 1:
                  key = bytes(v4).decode()
                elif f4 == 2:
                  vals = [_value(bytes(v5)) for f5, w5, v5, _ in protos.iter_fields(bytes(v4)) if f5 == 1]
              out['info'][key] = vals
            elif f3 == 7:
              out['genotype'] = [protos._to_signed32(x) for x in protos.unpack_varints(v3)] if w3 == 2 else out['genotype'] + [protos._to_signed32(v3)]   # pylint: disable=protected-access
            elif f3 == 9:
              out['call_set_name'] = v3.decode()
    elif fn == 2:
      key, names = '', []
      for f2, w2, v2, _ in protos.iter_fields(val):
        if f2 == 1:
          key = bytes(v2).decode()
        elif f2 == 2:
          names = [bytes(v3).decode() for f3, w3, v3, _ in protos.iter_fields(bytes(v2)) if f3 == 1]
      out['allele_support'][key] = sorted(names)
    elif fn == 4:
      out['ref_support'].append(val.decode())
    elif fn == 5:
      key, infos = '', []
      for f2, w2, v2, _ in protos.iter_fields(val):
        if f2 == 1:
          key = bytes(v2).decode()
        elif f2 == 2:
          infos = [_read_support(bytes(v3)) for f3, w3, v3, _ in protos.iter_fields(bytes(v2)) if f3 == 1]
      out['allele_support_ext'][key] = sorted(infos, key=lambda d: d['read_name'])
    elif fn == 6:
      out['ref_support_ext'] = sorted((_read_support(bytes(v3)) for f3, w3, v3, _ in protos.iter_fields(val) if f3 == 1),
                                      key=lambda d: d['read_name'])
    elif fn == 7:
      k = v = 0
      for f2, w2, v2, _ in protos.iter_fields(val):
        if f2 == 1:
          k = int(v2)
        elif f2 == 2:
          v = int(v2)
      out['af_at_position'][str(k)] = v
  out['ref_support'].sort()
  return out


def diff_fields(a: dict, b: dict):
  return [k for k in a if a[k] != b[k]]


def main():
  golden = [canonical(r) for r in tfrecord.read_records(os.path.join(TESTDATA, 'golden.calling_candidates.tfrecord.gz'))]
  bam_path = os.path.join(TESTDATA, 'input/NA12878_S1.chr20.10_10p1mb.bam')
  ref = fasta.IndexedFastaReader(os.path.join(TESTDATA, 'input/ucsc.hg19.chr20.unittest.fasta.gz'))
  table = bam.NativeBamTable(bam_path, bam.ReadRequirements(min_mapping_quality=5))
  opts = cand.CandidateOptions(sample_name=cand.sample_name_from_bam(bam_path), small_model_vaf_context_window_size=51)
  ours = []
  region = ('chr20', 9999999, 10010000)       # --regions chr20:10,000,000-10,010,000
  for contig, s, e in cand.regions_to_process([(c, ref.n_bases(c)) for c in ref.contig_order], opts.partition_size, region):
    ours += [canonical(r) for r in cand.candidates_in_region(table, ref, contig, s, e, opts).records]
  g_by = {(c['start'], c['ref'], tuple(c['alts'])): c for c in golden}
  o_by = {(c['start'], c['ref'], tuple(c['alts'])): c for c in ours}
  exact, partial, fixture = [], [], []
  for key, g in g_by.items():
    o = o_by.get(key)
    if o is None:
      continue
    d = diff_fields(g, o)
    if not d:
      exact.append(key[0])
      fixture.append(g)
    else:
      partial.append({'start': key[0], 'differs_in': d})
  report = {
      'golden_candidates': len(golden), 'ours_candidates': len(ours), 'sample_name': opts.sample_name,
      'same_site_and_alleles': len(exact) + len(partial), 'identical_in_every_field': len(exact),
      'same_alleles_different_counts_or_support': partial,
      'golden_only': sorted(k[0] for k in g_by if k not in o_by), 'ours_only': sorted(k[0] for k in o_by if k not in g_by),
      'note': 'golden made with --realign_reads (default); ours = no realigner',
  }
  os.makedirs(os.path.join(ROOT, 'tests/golden'), exist_ok=True)
  with open(os.path.join(ROOT, 'tests/golden/candidates_golden_report.json'), 'w') as f:
    json.dump(report, f, indent=1)
  print(json.dumps({k: v for k, v in report.items() if not isinstance(v, list) or len(v) < 30}, indent=1))
  return fixture, table, ref, opts


if __name__ == '__main__':
  main()
