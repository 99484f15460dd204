"""End-to-end throughput of the PRODUCT CLI (run_deepvariant, fused flow: BAM -> realigner -> candidates -> pileups -> classifier ->
CallVariantsOutput -> VCF), as a user runs it - VERDICT r1 item 4.  Two inputs:
  config1    BASELINE config 1: the quick-start reads NA12878 chr20:10,000,000-10,010,000 (tests/golden/quickstart.*; tools/make_quickstart_fixture.py)
  synthetic  a synthetic 30x / 150-bp coordinate-sorted BAM over --mbases megabases with planted SNPs and indels
Random-init weights of the right architecture (no checkpoints ship with the reference): the calls are noise, the work is real.
Prints one JSON object (and writes it to --out): seconds per stage, candidates (examples) per second, per configuration."""
import argparse
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))


def synthetic_inputs(d, n_bases, seed=11):
  import numpy as np
  import test_bam_native as tb
  rng = np.random.default_rng(seed)
  genome = rng.choice(np.frombuffer(b'ACGT', np.uint8), n_bases)
  alt = genome.copy()                                   # the second haplotype: a SNP every ~1 kb
  snps = np.arange(500, n_bases - 500, 997)
  alt[snps] = np.frombuffer(b'CGTA', np.uint8)[np.searchsorted(np.frombuffer(b'ACGT', np.uint8), genome[snps])]
  read_len, depth = 150, 30
  n_reads = n_bases * depth // read_len
  starts = np.sort(rng.integers(0, n_bases - read_len - 20, n_reads))
  recs = []
  for i, pos in enumerate(starts.tolist()):
    src = alt if i % 2 else genome
    seq = src[pos:pos + read_len].copy()
    if rng.random() < 0.3:
      seq[rng.integers(0, read_len)] = b'ACGT'[int(rng.integers(0, 4))]          # sequencing errors
    quals = rng.choice(np.array([11, 25, 37, 37, 37, 37, 37, 37], np.uint8), read_len)
    recs.append(tb._record(0, pos, f'r{i}', 60, 0x10 if i % 2 else 0, [(0, read_len)], seq.tobytes().decode(), quals.tolist()))
  bam_path = os.path.join(d, 'synthetic.bam')
  open(bam_path, 'wb').write(tb._bam(recs, refs=(('chr1', n_bases),)))
  fa = os.path.join(d, 'synthetic.fa')
  g = genome.tobytes().decode()
  open(fa, 'w').write('>chr1\n' + '\n'.join(g[i:i + 60] for i in range(0, n_bases, 60)) + '\n')
  open(fa + '.fai', 'w').write(f'chr1\t{n_bases}\t6\t60\t61\n')
  return fa, bam_path, f'chr1:1-{n_bases}', len(snps)


def run(name, fa, bam_path, regions, extra, out_dir):
  from deepvariant_b200 import cli, tfrecord
  d = os.path.join(out_dir, name)
  t0 = time.time()
  rc = cli.run_deepvariant(['--model_type', 'WGS', '--ref', fa, '--reads', bam_path, '--regions', regions, '--customized_model', 'random:1',
                            '--output_dir', d, '--output_vcf', os.path.join(d, 'out.vcf'), '--logging_dir', os.path.join(d, 'logs'), '--runtime_report'] + extra)
  dt = time.time() - t0
  assert rc == 0
  stages = {}
  import glob
  for path in glob.glob(os.path.join(d, 'logs', 'make_examples_runtime_by_region', '*.tsv')):
    rows = [line.rstrip('\n').split('\t') for line in open(path)]
    for row in rows[1:]:
      for k, v in zip(rows[0][1:8], row[1:8]):
        stages[k] = round(stages.get(k, 0) + float(v), 3)
  n = sum(1 for p in tfrecord.resolve_input_paths(os.path.join(d, 'call_variants_output.tfrecord.gz')) for _ in tfrecord.read_records(p))
  return {'seconds': round(dt, 2), 'make_examples_stage_seconds_summed_over_regions': stages, 'call_variants_outputs': n, 'candidates_per_s': round(n / dt, 1),
          'vcf_records': sum(1 for line in open(os.path.join(d, 'out.vcf')) if not line.startswith('#'))}


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument('--mbases', type=float, default=0.2)
  ap.add_argument('--shards', type=int, default=1)
  ap.add_argument('--gpus', type=int, default=1)
  ap.add_argument('--precision', type=int, default=0)
  ap.add_argument('--batch_size', type=int, default=512)     # classifier chunk per task (device memory per process)
  ap.add_argument('--out', default='')
  a = ap.parse_args()
  res = {'what': 'run_deepvariant (fused flow) wall clock, random-init weights', 'shards': a.shards, 'gpus': a.gpus, 'precision': a.precision,
         'host_cores': len(os.sched_getaffinity(0))}
  extra = ['--num_shards', str(a.shards), '--num_gpus', str(a.gpus), '--precision', str(a.precision), '--call_variants_extra_args', f'batch_size={a.batch_size}']
  with tempfile.TemporaryDirectory() as d:
    g = os.path.join(ROOT, 'tests', 'golden')
    if os.path.exists(os.path.join(g, 'quickstart.chr20_10mb.bam')):
      res['config1_quickstart_chr20_10kb'] = run('config1', os.path.join(g, 'quickstart.chr20_10mb.fa.gz'), os.path.join(g, 'quickstart.chr20_10mb.bam'),
                                                 'chr20:10,000,001-10,010,000', extra, d)
      res['config1_quickstart_chr20_10kb_norealign'] = run('config1_nr', os.path.join(g, 'quickstart.chr20_10mb.fa.gz'), os.path.join(g, 'quickstart.chr20_10mb.bam'),
                                                           'chr20:10,000,001-10,010,000', extra, d) if False else None
    fa, bam_path, regions, n_snps = synthetic_inputs(d, int(a.mbases * 1e6))
    res['synthetic_30x'] = dict(run('synthetic', fa, bam_path, regions, extra, d), megabases=a.mbases, planted_snps=n_snps)
    res['synthetic_30x']['kilobases_per_s'] = round(a.mbases * 1e3 / res['synthetic_30x']['seconds'], 2)
  res.pop('config1_quickstart_chr20_10kb_norealign', None)
  print(json.dumps(res))
  if a.out:
    json.dump(res, open(a.out, 'w'), indent=1)


if __name__ == '__main__':
  main()
