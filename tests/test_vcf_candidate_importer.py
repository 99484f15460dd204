"""--variant_caller vcf_candidate_importer (deepvariant_b200/vcf_candidate_importer.py + dvb_candidates_from_proposed): the known answers
of the reference's own tests of that path (deepvariant/variant_calling_test.cc:895-1237, built here from reads through the allele
counter instead of hand-made AlleleCounts), and the two goldens of make_examples_test.py:654-694 as committed fixtures
(tools/check_vcf_candidate_importer_golden.py: 22 / 22 records through the stage CLI, 223 / 223 images WITH read evidence)."""
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import oracle_lib  # noqa: E402
import test_candidates as tc  # noqa: E402
from deepvariant_b200 import _lib, candidates as cand, packing, ssw  # noqa: E402
from deepvariant_b200 import pileup_image as pi  # noqa: E402
from deepvariant_b200 import vcf_candidate_importer as vci  # noqa: E402

GOLDEN = os.path.join(ROOT, 'tests', 'golden')

# the VCF of variant_calling_test.cc:932-945 (testdata/input/test_calls_from_vcf.vcf.gz), as that test prints it
TEST_CALLS_VCF = ('##fileformat=VCFv4.2\n##contig=<ID=contigInHeaderWithCandidates,length=10>\n##contig=<ID=contigInHeaderNoCandidates,length=10>\n'
                  '#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\tdefault\n'
                  'contigInHeaderWithCandidates\t3\t.\tT\tG\t60\tPASS\t.\tGT\t./.\n'
                  'contigNotInHeaderWithCandidates\t1\t.\tA\tG\t60\tPASS\t.\tGT\t./.\n')


def _vcf(tmp_path, text, name='proposed.vcf'):
  path = str(tmp_path / name)
  with open(path, 'w') as f:
    f.write(text)
  return vci.ProposedVcfReader(path)


def _records(contig, rows):
  head = '##fileformat=VCFv4.2\n#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\ts\n'
  return head + ''.join(f'{contig}\t{pos + 1}\t.\t{ref}\t{",".join(alts) or "."}\t60\tPASS\t.\tGT\t{gt}\n' for pos, ref, alts, gt in rows)


def _calls(tmp_path, contig_bases, reads, proposed_rows, start, end, **options):
  kw = dict(vsc_min_count_snps=0, vsc_min_count_indels=0, vsc_min_fraction_snps=0.0, vsc_min_fraction_indels=0.0, sample_name='sample')
  kw.update(options)
  ref = tc.FakeRef([('chr1', contig_bases)])
  table = tc._table(tmp_path, reads, [('chr1', contig_bases)])
  reader = _vcf(tmp_path, _records('chr1', proposed_rows))
  found = vci.calls_from_vcf(table, ref, 'chr1', start, end, cand.CandidateOptions(**kw), reader, rows=np.arange(table.n_reads))
  return [cand.canonical_call(r) for r in found.records]


def test_reader_query_and_uncalled_genotypes(tmp_path):
  """TestCallsFromVcfQueryingVcf / TestCallPositionsFromVcfQueryingVcf / TestTrainUncalledGenotypes (:895-970, 1045-1072)."""
  reader = _vcf(tmp_path, TEST_CALLS_VCF)
  assert [v.start for v in reader.starting_in('contigInHeaderWithCandidates', 0, 5)] == [2]
  assert reader.starting_in('contigInHeaderNoCandidates', 0, 5) == [] and reader.starting_in('contigNotInVcf', 0, 5) == []
  assert vci.call_positions_from_vcf(reader, 'contigInHeaderWithCandidates', 0, 5) == [2]
  assert reader.starting_in('contigInHeaderWithCandidates', 0, 5, skip_uncalled_genotypes=True) == []      # ./. is skipped in training mode
  assert vci.region_has_proposed_variant(reader, 'contigInHeaderWithCandidates', 0, 5)
  assert not vci.region_has_proposed_variant(reader, 'contigInHeaderWithCandidates', 3, 10)
  # a record that overlaps the range but starts before it belongs to the previous range (CallsFromVcf, :404-408)
  reader = _vcf(tmp_path, _records('chr1', [(8, 'ACGT', ['A'], '0/1'), (12, 'C', ['T'], '1|1'), (13, 'G', [], '.')]), 'b.vcf')
  assert [v.start for v in reader.query('chr1', 10, 20)] == [8, 12, 13]
  assert [v.start for v in reader.starting_in('chr1', 10, 20)] == [12, 13]
  assert [v.genotype for v in reader.query('chr1', 0, 20)] == [[0, 1], [1, 1], [-1]]
  assert not reader.query('chr1', 0, 20)[2].is_uncalled_genotype()          # one allele only: not the ./. pattern


def test_calls_from_vcf_details(tmp_path):
  """TestCallsFromVcfDetails (:971-1043): 5 T, 3 A, 2 G reads under a proposed T -> G."""
  contig = b'GGTGGGGGGG'
  reads = [tc._read(f'r{i}', 2, b, '1M') for i, b in enumerate('AAAGGTTTTT')]
  got = _calls(tmp_path, contig, reads, [(2, 'T', ['G'], './.')], 0, 5)
  assert len(got) == 1
  g = got[0]
  assert (g['ref'], g['alts'], g['start'], g['end'], g['genotype'], g['call_set_name']) == ('T', ['G'], 2, 3, [-1, -1], 'sample')
  assert g['info'] == {'AD': [5, 2], 'DP': [10], 'VAF': [0.2]}
  assert sorted(g['allele_support']) == ['A', 'G'] and len(g['allele_support']['G']) == 2 and len(g['allele_support']['A']) == 3
  assert g['ref_support'] == [] and g['af_at_position'] == {}                     # no reference reads without --track_ref_reads, no VAF context
  ext = g['allele_support_ext']['G'][0]
  assert ext['mapping_quality'] == 0 and ext['sample_name'] == ''               # read name + is_low_quality only (AddSupportingReads :697-701)


def test_calls_from_variants_in_region(tmp_path):
  """TestCallsFromVariantsInRegion (:1074-1108): two proposed SNPs, positions without reads in between."""
  contig = b'G' * 10 + b'AGGGT' + b'G' * 10
  reads = [tc._read(f'read_{i}', 11, 'C', '1M') for i in range(10)]
  reads += [tc._read(f'r5_{i}', 14, 'C', '1M') for i in range(9)] + [tc._read(f'r5ref_{i}', 14, 'T', '1M') for i in range(2)]
  got = _calls(tmp_path, contig, reads, [(11, 'G', ['C'], '0/1'), (14, 'T', ['C'], '0/1')], 10, 15)
  assert [(g['start'], g['ref'], g['alts'], g['info']['AD'], g['info']['DP']) for g in got] == [(11, 'G', ['C'], [0, 10], [10]), (14, 'T', ['C'], [2, 9], [11])]
  assert got[0]['allele_support']['C'] == sorted(f'read_{i}/0' for i in range(10))
  assert len(got[1]['allele_support']['C']) == 9


@pytest.mark.parametrize('alleles,proposed_alts,ad', [
    ([('C', 10), ('G', 10)], ['C', 'G'], [0, 10, 10]),
    ([('C', 1000), ('G', 10)], ['C'], [0, 1000]),          # G falls under the 0.1 fraction: not in the map, its reads are UNCALLED_ALLELE
    ([('C', 10), ('G', 1000)], ['G'], [0, 1000]),
    ([('A', 1000), ('C', 500), ('G', 500)], ['C', 'G'], [1000, 500, 500])])
def test_compute_variant_multi_allelic(tmp_path, alleles, proposed_alts, ad):
  """TestComputeVariantMultiAllelic (:1110-1162)."""
  contig = b'GGGGGGGGGGATGC' + b'G' * 10
  reads = [tc._read(f'r{b}{i}', 10, b, '1M') for b, n in alleles for i in range(n)]
  got = _calls(tmp_path, contig, reads, [(10, 'A', proposed_alts, '0/1')], 10, 11, vsc_min_count_snps=10, vsc_min_count_indels=10,
               vsc_min_fraction_snps=0.1, vsc_min_fraction_indels=0.1)
  assert len(got) == 1 and got[0]['alts'] == proposed_alts and got[0]['info']['AD'] == ad and got[0]['info']['DP'] == [sum(n for _, n in alleles)]
  dropped = [b for b, n in alleles if b != 'A' and b not in proposed_alts]
  assert ('UNCALLED_ALLELE' in got[0]['allele_support']) == bool(dropped)


def test_compute_variant_different_refs(tmp_path):
  """TestComputeVariantDifferentRefs (:1165-1195): the reads' longest deletion extends the proposed CA -> C to CAA -> CA."""
  contig = b'GGGGGGGGGGCAAT' + b'G' * 10
  reads = [tc._read(f'ref{i}', 10, 'CA', '2M') for i in range(9)]
  reads += [tc._read(f'd1_{i}', 10, 'CA', '1M1D1M') for i in range(6)] + [tc._read(f'd2_{i}', 10, 'CT', '1M2D1M') for i in range(3)]
  got = _calls(tmp_path, contig, reads, [(10, 'CA', ['C'], '0/1')], 10, 11, vsc_min_count_snps=2, vsc_min_count_indels=2)
  assert len(got) == 1
  g = got[0]
  assert (g['ref'], g['alts'], g['start'], g['end']) == ('CAA', ['CA'], 10, 13)
  assert g['info']['AD'] == [9, 6] and g['info']['DP'] == [18]
  assert len(g['allele_support']['CA']) == 6 and len(g['allele_support']['C']) == 3      # the CAA -> C reads keep their own allele


def test_compute_variant_different_refs_2(tmp_path):
  """TestComputeVariantDifferentRefs2 (:1197-1237): TACAC -> T is counted under the proposed TACACACACAC -> TACACAC."""
  contig = b'GGGGGGGGGGTACACACACACG' + b'G' * 10
  reads = [tc._read(f'ref{i}', 10, 'TA', '2M') for i in range(8)] + [tc._read(f'del{i}', 10, 'TA', '1M4D1M') for i in range(4)]
  got = _calls(tmp_path, contig, reads, [(10, 'TACACACACAC', ['TACACAC', 'T'], '0/1')], 10, 11, vsc_min_count_snps=2, vsc_min_count_indels=2)
  assert len(got) == 1
  g = got[0]
  assert (g['ref'], g['alts'], g['end']) == ('TACACACACAC', ['TACACAC', 'T'], 21)
  assert g['info']['AD'] == [8, 4, 0] and g['info']['DP'] == [12] and g['info']['VAF'] == [4 / 12, 0.0]
  assert list(g['allele_support']) == ['TACACAC'] and len(g['allele_support']['TACACAC']) == 4


def test_no_evidence_non_canonical_and_errors(tmp_path):
  contig = b'GGGGGGGGGGANGC' + b'G' * 10
  reads = [tc._read('far', 20, 'GG', '2M')]
  got = _calls(tmp_path, contig, reads, [(10, 'A', ['C', 'AT'], '0/1'), (11, 'N', ['C'], '0/1'), (12, 'G', [], '0/0')], 10, 15)
  assert [(g['start'], g['alts']) for g in got] == [(10, ['C', 'AT']), (12, [])]             # the N site is dropped (ComputeVariant :519-523)
  assert got[0]['info'] == {'AD': [0, 0, 0], 'DP': [0], 'VAF': [0.0, 0.0]} and got[0]['allele_support'] == {}
  assert got[1]['info']['AD'] == [0] and got[1]['info']['VAF'] == []
  # --track_ref_reads: the reference-supporting reads of the proposed positions are listed
  reads = [tc._read(f'ref{i}', 10, 'AN', '2M') for i in range(3)] + [tc._read('alt', 10, 'CN', '2M')]
  g = _calls(tmp_path, contig, reads, [(10, 'A', ['C'], '0/1')], 10, 11, track_ref_reads=True)[0]
  assert g['info']['AD'] == [3, 1] and g['ref_support'] == ['ref0/0', 'ref1/0', 'ref2/0'] and g['allele_support']['C'] == ['alt/0']
  # a proposed reference allele the genome contradicts: the reference QCHECK-fails, here an error
  with pytest.raises(_lib.DvbError, match='incorrect ref bases'):
    _calls(tmp_path, contig, reads, [(10, 'T', ['C'], '0/1')], 10, 11)
  table = tc._table(tmp_path, reads, [('chr1', contig)])
  with pytest.raises(_lib.DvbError):
    co = cand.CandidateOptions().to_c()
    _lib.check(_lib.lib().dvb_candidates_from_proposed(table.handle, b'chr1', None, 0, 0, 1, None, 0, co, None, 0, 0, None, None, None, None, None))


def test_libssw_penalises_n_like_a_mismatch():
  """libssw 1.2.5's ssw_cpp scores N against anything as a mismatch (BuildSwScoreMatrix).  The read below (59 N-ridden bases, then 42
  clean ones; HSQ1004:134:C0D8DACXX:1:2307:21360:100640 of the reference's NA12878 test BAM) is the one read of the importer's
  training golden whose row tells the two conventions apart: 59S42M in the golden, 4S97M with N scored 0."""
  ref = 'ATTTAAGGGTTAGTGTGCATTTTAATTGACCATTCAATTTCAAATAAATGTGGAGGAAATTACAACCCTCTTAACAGAGAAACGACAATTTAAGGAGAATGAGACCTTGATTGAAACAATTAGGAACTTAGAAACCAGACTCTGATGTGG'
  read = 'NNNNA' + 'N' * 54 + 'GAATGAGACCTTGATTGAAACCANNAGGAACTTAGAAACCAG'
  aligner = ssw.Aligner(match_score=4, mismatch_penalty=6, gap_opening_penalty=8, gap_extending_penalty=2)     # the realigner's scores
  aligner.set_reference_sequence(ref)
  a = aligner.align(read)
  assert a.cigar_string.startswith('59S') and a.ref_begin == ref.index('GAATGAGACC')
  # N inside an otherwise matching stretch costs a mismatch, it does not end the alignment
  assert a.sw_score == 4 * 39 - 6 * 3 and a.cigar_string == '59S21=1X1=2X17='       # C/A, then the two N over TT


def _fixture(name):
  d = np.load(os.path.join(GOLDEN, name))
  arrays = {k[4:]: d[k] for k in d.files if k.startswith('arr_')}
  pb = packing.PackedBatch(int(d['n_images']), int(d['n_reads']), int(d['n_pairs']), int(d['ref_stride']), arrays)
  o = pi.default_options(pi.ReadRequirements(min_base_quality=10, min_mapping_quality=5))
  o.channels = list(pi.PILEUP_CHANNELS_WITH_INSERT_SIZE)
  return pb, d['golden_images'], o


@pytest.mark.parametrize('name', ['vcf_candidate_importer_golden_subset.npz', 'vcf_candidate_importer_training_subset.npz'])
def test_oracle_reproduces_importer_golden_images(name):
  pb, golden, o = _fixture(name)
  assert golden.shape[1:] == (100, 221, 7) and golden.shape[0] >= 8
  if 'training' in name:
    assert all(g[5:].any() for g in golden)          # every image of this fixture has read rows
  np.testing.assert_array_equal(oracle_lib.encode_batch(pi.to_params(o), pb), golden)


@pytest.mark.gpu
@pytest.mark.parametrize('name', ['vcf_candidate_importer_golden_subset.npz', 'vcf_candidate_importer_training_subset.npz'])
def test_cuda_encoder_reproduces_importer_golden_images(name):
  pb, golden, o = _fixture(name)
  enc = pi.GpuEncoder(pi.to_params(o), 0)
  np.testing.assert_array_equal(enc.encode_host(pb), golden)


def test_golden_report_is_current():
  r = json.load(open(os.path.join(GOLDEN, 'vcf_candidate_importer_golden_report.json')))
  assert r['golden_examples'] == r['images_identical'] == r['variants_identical_fields'] == 20 and r['same_examples_in_same_order']
  assert r['stage_cli'] == {'golden_records': 22, 'records_written': 22, 'records_equal_in_order': 22, 'example_info_json_equal': True}
  t = r['training_golden']
  assert t['golden_examples'] == t['images_identical'] == t['variants_identical_fields_but_genotype'] == t['golden_images_with_read_rows'] == 223
  assert t['same_examples_in_same_order'] and t['extra_examples'] == []


def _proposed_for_planted(tmp_path, genome, sites):
  """The four planted variants as a proposed VCF, plus a SNP nobody carries, a record in a partition without reads and a ./. record."""
  g = genome
  rows = [(sites['snp_het'], g[sites['snp_het']], ['ACGT'[('ACGT'.index(g[sites['snp_het']]) + 1) % 4]], '0/1'),
          (1800, g[1800], ['ACGT'[('ACGT'.index(g[1800]) + 2) % 4]], '0/1'),                        # no read shows it: AD [n, 0]
          (sites['snp_hom'], g[sites['snp_hom']], ['ACGT'[('ACGT'.index(g[sites['snp_hom']]) + 1) % 4]], '1/1'),
          (sites['ins'], g[sites['ins']], [g[sites['ins']] + 'GT'], '0/1'),
          (sites['dele'], g[sites['dele']:sites['dele'] + 4], [g[sites['dele']]], './.'),
          (5500, g[5500], ['ACGT'[('ACGT'.index(g[5500]) + 1) % 4]], '0/1')]                        # outside --regions
  path = str(tmp_path / 'proposed.vcf')
  with open(path, 'w') as f:
    f.write(_records('chr20', rows))
  return path, rows


def test_make_examples_cli_with_proposed_variants_cpu_plumbing(tmp_path, monkeypatch):
  """make_examples --variant_caller vcf_candidate_importer --proposed_variants: one candidate per proposed record of the regions, in
  file order, evidence from the reads (with and without the realigner), pileups through the same planner.  The encoder is the CPU
  oracle here; test_cuda_encoder_reproduces_importer_golden_images covers the kernel on the reference's own golden."""
  from deepvariant_b200 import make_examples_native as men, protos
  monkeypatch.setattr(men.ExamplesGenerator, '_gpu', lambda self: tc.OracleEncoder(pi.to_params(self.options.pic_options, height=self.pileup_image_height)))
  fa, bam_path, genome, sites = tc._planted_case(tmp_path)
  vcf, rows = _proposed_for_planted(tmp_path, genome, sites)
  in_regions = [r for r in rows if r[0] < 5000]
  for tag, realign in (('plain', False), ('realigned', True)):
    examples, cands = tc._run_cli(tmp_path, fa, bam_path, 'vci_' + tag, realign=realign,
                                  extra=('--variant_caller', 'vcf_candidate_importer', '--proposed_variants', vcf))
    calls = [cand.canonical_call(r) for r in cands]
    assert [(c['start'], c['ref'], c['alts']) for c in calls] == [(p, ref, alts) for p, ref, alts, _ in in_regions]
    by = {c['start']: c for c in calls}
    assert by[1800]['info']['AD'][1] == 0 and by[1800]['info']['AD'][0] > 10 and by[1800]['info']['VAF'] == [0.0]
    assert by[sites['snp_hom']]['info']['AD'][0] == 0 and by[sites['snp_hom']]['info']['AD'][1] > 10
    assert 0.3 < by[sites['snp_het']]['info']['VAF'][0] < 0.7
    assert by[sites['ins']]['info']['AD'][1] > 5 and by[sites['dele']]['info']['AD'][1] > 5       # ./. is only skipped in training mode
    assert all(c['call_set_name'] == 'planted' and c['genotype'] == [-1, -1] and c['af_at_position'] == {} for c in calls)
    ex = [protos.parse_tf_example(r) for r in examples]
    assert [protos.parse_variant(e['variant/encoded'][1][0]).start for e in ex] == [r[0] for r in in_regions]
    # the proposed SNP without support: no read row carries the supports-variant mark (channel 4 is 152 = "does not support" under reads)
    img = np.frombuffer(ex[1]['image/encoded'][1][0], np.uint8).reshape(100, 221, 7)
    assert img[5:, :, 0].any() and not (img[5:, :, 4] == 254).any()
    img = np.frombuffer(ex[2]['image/encoded'][1][0], np.uint8).reshape(100, 221, 7)
    assert (img[5:, :, 4] == 254).any()
  # flag errors of make_examples_options.py:1440-1490
  from deepvariant_b200 import cli
  base = ['--mode', 'calling', '--ref', fa, '--reads', bam_path, '--examples', str(tmp_path / 'x.tfrecord.gz'), '--regions', 'chr20:1001-2000']
  with pytest.raises(SystemExit, match='--proposed_variants is required'):
    cli.make_examples(base + ['--variant_caller', 'vcf_candidate_importer'])
  with pytest.raises(SystemExit, match='needs --variant_caller'):
    cli.make_examples(base + ['--proposed_variants', vcf])


def test_run_deepvariant_extra_args_helpers(capsys):
  """scripts/run_deepvariant.py:330-386: --make_examples_extra_args / --postprocess_variants_extra_args."""
  from deepvariant_b200 import cli
  d = cli.extra_args_to_dict('variant_caller=vcf_candidate_importer,proposed_variants=X.vcf.gz,regions=chr1:1-2,chr2,--realign_reads=false,sort_by_haplotypes=TRUE')
  assert d == {'variant_caller': 'vcf_candidate_importer', 'proposed_variants': 'X.vcf.gz', 'regions': 'chr1:1-2,chr2', 'realign_reads': False,
               'sort_by_haplotypes': True}
  assert cli.extra_args_to_dict('') == {} and cli.extra_args_to_dict(None) == {}
  with pytest.raises(ValueError):
    cli.extra_args_to_dict('just_a_flag')
  args = ['--mode', 'calling', '--partition_size', '25000', '--norealign_reads', '--phase_reads']
  out = cli.apply_extra_args(args, cli.extra_args_to_dict('partition_size=1000,realign_reads=true,phase_reads=false,variant_caller=vcf_candidate_importer'),
                             cli._value_flags(args))
  assert out == ['--mode', 'calling', '--partition_size', '1000', '--realign_reads', '--variant_caller', 'vcf_candidate_importer']
  err = capsys.readouterr().err
  assert 'Warning: --partition_size is previously set to 25000, now to 1000.' in err and '--phase_reads is previously set to True, now to False' in err
  assert args == ['--mode', 'calling', '--partition_size', '25000', '--norealign_reads', '--phase_reads']       # the input list is not touched
  assert cli.apply_extra_args(['--a', '1'], {'realign_reads': False, 'group_variants': False, 'other': False}, {'a'}) == ['--a', '1', '--nogroup_variants', '--norealign_reads']
