"""CUDA allele counting (csrc/dvb_candidates.cu: dvb_allele_count_kernel / dvb_allele_flag_kernel; SURVEY 8(f) next row #2, device
half) against its host instantiation (the same dvb_allele::WalkRead, sink and flag function compiled for the CPU, which
tests/test_candidates.py pins against the allele counter, the Python oracle and the reference's golden candidates), bit for bit.
Named to sort last: these kernels were written after the round's GPU budget was spent and are first run by the driver."""
import json
import os
import random

import numpy as np
import pytest

import test_candidates as tc
from deepvariant_b200 import bam, candidates as cand

pytestmark = pytest.mark.gpu


def _compare(counter, table, ref, contig, start, end, rows, o):
  want_counts, want_flags = cand.debug_dense_counts_host(table, ref, contig, start, end, rows, o, windowed=True)
  got_counts, got_flags = counter.count_region(ref, contig, start, end, rows, o)
  np.testing.assert_array_equal(got_counts, want_counts)
  np.testing.assert_array_equal(got_flags, want_flags)
  return got_flags


def test_make_examples_cli_generates_candidates_on_gpu(tmp_path, monkeypatch):
  from deepvariant_b200 import make_examples_native as men, pileup_image as pi
  fa, bam_path, genome, sites = tc._planted_case(tmp_path)
  examples, cands = tc._run_cli(tmp_path, fa, bam_path, 'gpu')
  tc._check_planted(examples, cands, genome, sites)
  monkeypatch.setattr(men.ExamplesGenerator, '_gpu', lambda self: tc.OracleEncoder(pi.to_params(self.options.pic_options, height=self.pileup_image_height)))
  want_examples, want_cands = tc._run_cli(tmp_path, fa, bam_path, 'oracle')
  assert cands == want_cands and examples == want_examples          # CUDA encoder == CPU oracle, record for record


def test_run_deepvariant_from_bam_to_vcf(tmp_path):
  """run_deepvariant --ref --reads --output_vcf on the planted genome: candidates (host counter + caller), pileups (CUDA encoder),
  genotype likelihoods (tcgen05 classifier, random-init weights), CallVariantsOutput shards, VCF."""
  from deepvariant_b200 import cli
  fa, bam_path, genome, sites = tc._planted_case(tmp_path)
  out_dir, vcf, gvcf_path = str(tmp_path / 'work'), str(tmp_path / 'out.vcf'), str(tmp_path / 'out.g.vcf')
  assert cli.run_deepvariant(['--model_type', 'WGS', '--ref', fa, '--reads', bam_path, '--output_dir', out_dir, '--output_vcf', vcf,
                              '--output_gvcf', gvcf_path, '--regions', 'chr20:1001-5000', '--customized_model', 'random:3']) == 0
  # the gVCF tiles the calling region: reference blocks and the four variants, every position in exactly one record
  nxt = 1001
  for r in (line.split('\t') for line in open(gvcf_path) if not line.startswith('#')):
    assert int(r[1]) == nxt and r[4].endswith('<*>')
    nxt = int(r[7][4:]) + 1 if r[7].startswith('END=') else int(r[1]) + len(r[3])
  assert nxt == 5001
  lines = [line.split('\t') for line in open(vcf).read().split('\n') if line and not line.startswith('##')]
  assert lines[0][-1] == 'planted'
  records = lines[1:]
  assert [int(r[1]) - 1 for r in records] == sorted(sites.values())
  for r in records:
    assert r[8] == 'GT:GQ:DP:AD:VAF:PL' and r[6] in ('PASS', 'RefCall', 'NoCall', 'LowQual')
    fields = dict(zip(r[8].split(':'), r[9].split(':')))
    assert len(fields['PL'].split(',')) == 3 and 0 in [int(x) for x in fields['PL'].split(',')]
    assert int(fields['DP']) >= sum(int(x) for x in fields['AD'].split(',')) > 0


@pytest.mark.parametrize('seed', range(6))
def test_device_counts_and_flags_equal_the_host_instantiation_random(tmp_path, seed):
  rng = random.Random(3000 + seed)
  contig, reads = tc._random_case(rng, rng.choice([5, 300, 1200]))
  ref = tc.FakeRef([('chr1', contig)])
  table = tc._table(tmp_path, reads, [('chr1', contig)])
  counter = cand.GpuAlleleCounter(table)
  o = cand.CandidateOptions(min_mapping_quality=rng.choice([0, 5]), vsc_min_count_snps=rng.choice([1, 2]),
                            keep_legacy_allele_counter_behavior=seed == 3, sample_name='s')
  for _ in range(3):     # several regions through one handle (scratch buffers are reused and must be re-zeroed)
    start, end = rng.randrange(0, 60), rng.randrange(120, len(contig) + 1)
    rows = np.arange(len(reads)) if rng.random() < 0.5 else np.array(sorted(rng.sample(range(len(reads)), max(1, len(reads) // 2))))
    _compare(counter, table, ref, 'chr1', start, end, rows, o)
  _compare(counter, table, ref, 'chr1', 10, 20, np.zeros(0, np.int64), o)      # no reads: all zero
  assert counter.launch_count >= 7
  counter.close()


def test_gpu_candidates_equal_host_candidates_on_the_golden_fixture():
  """The reference's golden candidates through the device pass: counts + flags on the GPU, exact calls on the flagged sites."""
  fx = json.load(open(os.path.join(tc.GOLDEN, 'candidates_golden_subset.json')))
  contig = b'N' * fx['slice_start'] + fx['slice'].encode()
  contig += b'N' * (fx['n_bases'] - len(contig))
  ref = tc.FakeRef([(fx['contig'], contig)])
  table = bam.NativeBamTable(os.path.join(tc.GOLDEN, 'candidates_golden_subset.bam'), bam.ReadRequirements(min_mapping_quality=5))
  counter = cand.GpuAlleleCounter(table)
  o = cand.CandidateOptions(sample_name=fx['sample_name'], small_model_vaf_context_window_size=51)
  for part in fx['partitions']:
    rows = cand.region_reads(table, fx['contig'], part['start'], part['end'])
    _compare(counter, table, ref, fx['contig'], part['start'], part['end'], rows, o)
    got = cand.candidates_in_region_gpu(counter, ref, fx['contig'], part['start'], part['end'], o)
    want = cand.candidates_in_region(table, ref, fx['contig'], part['start'], part['end'], o)
    assert got.records == want.records and len(got.records) == len(part['expected'])
  counter.close()


def test_one_launch_over_a_long_interval(tmp_path):
  """A 200-kb interval with 30x of reads in one launch pair (the shape the bench times)."""
  rng = np.random.default_rng(9)
  n = 200_000
  genome = rng.choice(np.frombuffer(b'ACGT', np.uint8), n).tobytes()
  recs = []
  for i in range(60_000):
    pos = int(rng.integers(0, n - 120))
    seq = bytearray(genome[pos:pos + 100])
    for j in rng.integers(0, 100, 1):
      seq[j] = b'ACGT'[int(rng.integers(0, 4))]
    recs.append((pos, tc.tb._record(0, pos, f'q{i}', 60, 0, [(0, 100)], seq.decode(), rng.integers(5, 41, 100).tolist())))
  recs.sort(key=lambda t: t[0])
  path = str(tmp_path / 'big.bam')
  open(path, 'wb').write(tc.tb._bam([r for _, r in recs], refs=(('chr1', n),)))
  table = bam.NativeBamTable(path, bam.ReadRequirements(min_mapping_quality=5))
  ref = tc.FakeRef([('chr1', genome)])
  counter = cand.GpuAlleleCounter(table)
  flags = _compare(counter, table, ref, 'chr1', 0, n, np.arange(table.n_reads), cand.CandidateOptions())
  assert 0 < int((flags != 0).sum()) < n // 20
  counter.close()
