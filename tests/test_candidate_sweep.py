"""make_examples --mode candidate_sweep and --candidate_positions: known answers transcribed from
deepvariant/make_examples_core_test.py (:482-545 test_partition_by_candidates, :546-663 test_merge_ranges_from_files_sequential),
the reference's golden.candidate_positions files, and the two-pass flow through the stage CLI."""
import json
import os

import numpy as np
import pytest

from deepvariant_b200 import candidates as cand

GOLDEN = os.path.join(os.path.dirname(__file__), 'golden')
R, P = cand.END_OF_REGION, cand.END_OF_PARTITION


def _lit(s):          # '1:1-10' -> ('1', 0, 10)
  name, span = s.split(':')
  a, b = span.split('-')
  return (name, int(a) - 1, int(b))


@pytest.mark.parametrize('regions,positions,max_size,want', [
    (['1:1-10'], [2, 4, 5, R], 2, ['1:1-5', '1:6-10']),
    (['1:1-10', '1:15-20'], [2, 4, 5, R, 16, 19, R], 2, ['1:1-5', '1:6-10', '1:15-20']),
    (['1:1-10', '2:1-20'], [2, 4, 5, R, 2, 3, 12, R], 2, ['1:1-5', '1:6-10', '2:1-4', '2:5-20']),
    (['1:1-1000200'], [3, 5, 7, R], 2, ['1:1-6', '1:7-1000006', '1:1000007-1000200']),
], ids=['one_interval', 'two_intervals', 'two_intervals_different_contigs', 'candidate_far_apart'])
def test_partition_by_candidates(regions, positions, max_size, want):
  assert sorted(cand.partition_by_candidates([_lit(r) for r in regions], positions, max_size)) == sorted(_lit(w) for w in want)


def test_partition_by_candidates_errors():
  with pytest.raises(ValueError, match='max_size'):
    cand.partition_by_candidates([('1', 0, 10)], [R], 0)
  with pytest.raises(ValueError, match='Terminating item'):
    cand.partition_by_candidates([('1', 0, 10)], [2, 4], 2)


@pytest.mark.parametrize('arrays,want', [
    ([[1, 2, 3, P, 7, 8, 9, P, R], [4, 5, 6, P]], [1, 2, 3, 4, 5, 6, 7, 8, 9, R]),
    ([[1, 3, 7, P], [9, 11, P, R]], [1, 3, 7, 9, 11, R]),
    ([[1, 2, 3, 4, 7, P, R]], [1, 2, 3, 4, 7, R]),
    ([[1, 2, 3, 4, 7, P, R], []], [1, 2, 3, 4, 7, R]),
], ids=['simple', 'one_partition_in_each_shard', 'one_shard', 'empty_shard'])
def test_merge_ranges_from_files_sequential(arrays, want):
  assert cand.merge_ranges_from_files_sequential([np.array(a, dtype=np.int32) for a in arrays]) == want


@pytest.mark.parametrize('arrays', [[[1, 7, 3, P], [4, 5, P, R]], [[1, 3, 7, P], [4, 5, P, R]]], ids=['unordered_input', 'unordered_input_2'])
def test_merge_rejects_unordered_input(arrays):
  with pytest.raises(AssertionError):
    cand.merge_ranges_from_files_sequential([np.array(a, dtype=np.int32) for a in arrays])


def test_golden_shards_merge_to_the_unsharded_golden():
  whole = np.fromfile(os.path.join(GOLDEN, 'golden.candidate_positions'), dtype=np.int32)
  merged = cand.load_candidate_positions(os.path.join(GOLDEN, 'golden.candidate_positions@3'))
  assert merged == [int(x) for x in whole if x != P]
  assert merged == cand.load_candidate_positions(os.path.join(GOLDEN, 'golden.candidate_positions'))
  assert len(merged) == 83 and merged[-1] == R
  # 82 candidates, at most 200 per partition: the calling region stays one partition
  assert cand.regions_to_process([('chr20', 63025520)], 1000, ('chr20', 9999999, 10010000), candidates=merged) == [('chr20', 9999999, 10010000)]


def test_candidate_sweep_report_is_current():
  r = json.load(open(os.path.join(GOLDEN, 'candidate_sweep_report.json')))
  assert r['unsharded_byte_identical'] and r['shards_byte_identical'] == [True, True, True]


def test_two_pass_flow_through_the_cli(tmp_path, monkeypatch):
  """candidate_sweep on 2 shards, then calling with --candidate_positions: the partitions are cut by candidate count (here: the
  whole region in one partition) and the same sites come out as with fixed 1-kb partitions (the records themselves may differ:
  a partition's reads are what its in-memory reader holds, and the 4-kb partition is reservoir-sampled to 1500 reads)."""
  import test_candidates as tc
  from deepvariant_b200 import cli, make_examples_native as men, pileup_image as pi, tfrecord
  monkeypatch.setattr(men.ExamplesGenerator, '_gpu', lambda self: tc.OracleEncoder(pi.to_params(self.options.pic_options, height=self.pileup_image_height)))
  fa, bam_path, genome, sites = tc._planted_case(tmp_path)
  common = ['--ref', fa, '--reads', bam_path, '--channel_list', 'BASE_CHANNELS,insert_size', '--regions', 'chr20:1001-5000', '--norealign_reads']
  pos = str(tmp_path / 'positions@2')
  for task in (0, 1):
    assert cli.make_examples(['--mode', 'candidate_sweep', '--examples', str(tmp_path / 'sweep.tfrecord@2.gz'), '--candidate_positions', pos, '--task', str(task)] + common) == 0
  merged = cand.load_candidate_positions(pos)
  assert merged == sorted(sites.values()) + [R]
  outs = []
  for tag, extra in (('fixed', []), ('swept', ['--candidate_positions', pos])):
    ex, cs = str(tmp_path / f'{tag}.tfrecord.gz'), str(tmp_path / f'{tag}.candidates.tfrecord.gz')
    assert cli.make_examples(['--mode', 'calling', '--examples', ex, '--candidates', cs] + extra + common) == 0
    outs.append((len(list(tfrecord.read_records(ex))), [cand.canonical_call(r)['start'] for r in tfrecord.read_records(cs)]))
  assert outs[0] == outs[1] == (4, sorted(sites.values()))


def test_calling_intervals_and_several_regions(tmp_path, monkeypatch):
  contigs = [('chr1', 1000), ('chr2', 500)]
  assert cand.calling_intervals(contigs) == [('chr1', 0, 1000), ('chr2', 0, 500)]
  assert cand.calling_intervals(contigs, ('chr2', 100, 9999)) == [('chr2', 100, 500)]
  # any order, overlaps merged, empty and unknown-contig regions dropped, output in contig order
  assert cand.calling_intervals(contigs, [('chr2', 100, 200), ('chr1', 50, 80), ('chr1', 70, 120), ('chr1', 300, 300), ('chrX', 0, 5)]) == \
      [('chr1', 50, 120), ('chr2', 100, 200)]
  assert cand.regions_to_process(contigs, 100, [('chr2', 100, 250), ('chr1', 950, 2000)]) == [('chr1', 950, 1000), ('chr2', 100, 200), ('chr2', 200, 250)]
  # the stage CLI: two literals and a BED file; the planted variants inside them come out, the others do not
  import test_candidates as tc
  from deepvariant_b200 import cli, make_examples_native as men, pileup_image as pi, tfrecord
  monkeypatch.setattr(men.ExamplesGenerator, '_gpu', lambda self: tc.OracleEncoder(pi.to_params(self.options.pic_options, height=self.pileup_image_height)))
  fa, bam_path, genome, sites = tc._planted_case(tmp_path)
  bed = tmp_path / 'r.bed'
  bed.write_text(f'chr20\t{sites["dele"] - 100}\t{sites["dele"] + 100}\n')
  assert cli.parse_regions(f'chr20:1,401-1,600 chr20 {bed}') == [('chr20', 1400, 1600), ('chr20', 0, 1 << 40), ('chr20', sites['dele'] - 100, sites['dele'] + 100)]
  cs = str(tmp_path / 'c.tfrecord.gz')
  assert cli.make_examples(['--mode', 'calling', '--ref', fa, '--reads', bam_path, '--examples', str(tmp_path / 'e.tfrecord.gz'), '--candidates', cs,
                            '--channel_list', 'BASE_CHANNELS,insert_size', '--norealign_reads',
                            '--regions', f'chr20:{sites["snp_het"] - 50}-{sites["snp_het"] + 50} {bed}']) == 0
  assert [cand.canonical_call(r)['start'] for r in tfrecord.read_records(cs)] == [sites['snp_het'], sites['dele']]
