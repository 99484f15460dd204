"""gVCF: reference-confidence blocks (make_examples --gvcf) and their merge with the called variants (postprocess_variants
--gvcf_outfile).  Known answers transcribed from deepvariant/variant_caller_test.py (:135-256 test_ref_calc / test_rescale_read_counts,
:274-310 test_gvcf_basic*, :355-403 test_make_gvcfs, :405-493 test_quantize_gvcfs); the reference's golden pairs copied by
tools/make_postprocess_fixtures.py; the from-BAM pin is tools/check_gvcf_golden.py (tests/golden/gvcf_golden_report.json)."""
import gzip
import json
import os

import numpy as np
import pytest

from deepvariant_b200 import gvcf, tfrecord
from deepvariant_b200 import postprocess_variants as pp

GOLDEN = os.path.join(os.path.dirname(__file__), 'golden')

REF_CALC = [
    (0, 0, 0.01, 100, [-0.477121, -0.477121, -0.477121], 1, False),
    (10, 0, 0.01, 100, [-0.000469, -2.967121, -19.956821], 29, False),
    (10, 1, 0.01, 100, [-0.044109, -1.015126, -16.009190], 10, False),
    (10, 2, 0.01, 100, [-1.063830, -0.039211, -13.037641], 0, False),
    (10, 5, 0.01, 100, [-7.011524, -0.000000, -7.011524], 0, False),
    (10, 10, 0.01, 100, [-19.956821, -2.967121, -0.000469], 0, False),
    (20, 0, 0.01, 100, [-0.000001, -5.933304, -39.912704], 59, False),
    (20, 1, 0.01, 100, [-0.000050, -3.937719, -35.921484], 39, False),
    (20, 2, 0.01, 100, [-0.004935, -1.946968, -31.935098], 19, False),
    (20, 3, 0.01, 100, [-0.328657, -0.275056, -28.267550], 2, False),
    (20, 17, 0.01, 100, [-28.267550, -0.275056, -0.328657], 0, False),
    (10, 0, 0.1, 100, [-0.001215, -2.553940, -9.543640], 25, False),
    (10, 1, 0.1, 100, [-0.010811, -1.609294, -7.644752], 16, False),
    (10, 0, 0.001, 100, [-0.000428, -3.006383, -29.996083], 30, False),
    (10, 1, 0.001, 100, [-0.297847, -0.304236, -24.294371], 3, False),
    (10, 0, 1e-04, 100, [-0.000424, -3.010290, -39.999990], 30, False),
    (10, 1, 1e-04, 100, [-1.032394, -0.042303, -33.032046], 0, False),
    (30, 0, 0.01, 100, [-0.000000, -8.899956, -59.869056], 88, False),
    (40, 0, 0.01, 100, [-0.000000, -11.866608, -79.825408], 100, False),
    (100, 0, 0.01, 100, [0.000000, -29.666519, -199.563519], 100, False),
    (10, 8, 0.01, 100, [-11.97381, -9.949651e02, -0.0000000000004609646], 0, True),
    (10, 1, 0.01, 100, [0.0, -996.960717, -15.965082], 100, True),
    (10, 5, 0.01, 100, [-0.30103, -989.2792, -0.3010300], 3, True),
]


def _confidence(p_error, max_gq, gq_resolution=1):
  return gvcf.GvcfOptions(sample_name='s', p_error=p_error, max_gq=max_gq, gq_resolution=gq_resolution, max_cache_coverage=0)


@pytest.mark.parametrize('total_n,alt_n,p_error,max_gq,likelihoods,want_gq,haploid', REF_CALC)
def test_ref_calc(total_n, alt_n, p_error, max_gq, likelihoods, want_gq, haploid):
  gq, got = gvcf.ReferenceConfidence(_confidence(p_error, max_gq))(total_n - alt_n, total_n, haploid)
  np.testing.assert_allclose(likelihoods, got, atol=1e-6, rtol=1e-6)
  assert gq == want_gq


@pytest.mark.parametrize('n_ref,n_total,max_allowed,want', [
    (0, 10, 100, (0, 10)), (5, 10, 100, (5, 10)), (10, 10, 100, (10, 10)), (10, 100, 100, (10, 100)), (100, 100, 100, (100, 100)),
    (0, 200, 100, (0, 100)), (0, 10000, 100, (0, 100)), (1, 200, 100, (1, 100)), (1, 100000, 100, (1, 100)), (2, 200, 100, (1, 100)),
    (3, 200, 100, (2, 100)), (4, 200, 100, (2, 100)), (10, 200, 100, (5, 100)), (50, 200, 100, (25, 100)), (100, 200, 100, (50, 100)),
    (200, 200, 100, (100, 100)), (99, 100, 100, (99, 100))])
def test_rescale_read_counts(n_ref, n_total, max_allowed, want):
  assert gvcf.rescale_read_counts_if_necessary(n_ref, n_total, max_allowed) == want


def test_cache_rescales_deep_sites():
  # VerySensitiveCaller's table (max coverage 100): a site with 400 reads is answered from the rescaled cell
  exact = gvcf.ReferenceConfidence(_confidence(0.001, 50))
  cached = gvcf.ReferenceConfidence(gvcf.GvcfOptions(p_error=0.001, max_gq=50, max_cache_coverage=100))
  assert cached(396, 400) == exact(99, 100)
  assert cached(90, 100) == exact(90, 100)


def _blocks(counts, start=1, **kw):
  o = _confidence(0.01, 100, kw.pop('gq_resolution', 1))
  o.include_med_dp = kw.pop('include_med_dp', False)
  summary = np.array([(n_ref, n_ref + n_alt) for n_alt, n_ref, _ in counts], dtype=np.int32).reshape(-1, 2)
  return list(gvcf.make_gvcfs('chr1', start, ''.join(b for _, _, b in counts), summary, o))


@pytest.mark.parametrize('base', 'ACGT')
@pytest.mark.parametrize('include_med_dp', [True, False])
def test_gvcf_basic(base, include_med_dp):
  (b,) = _blocks([(0, 0, base)], start=100, include_med_dp=include_med_dp)
  assert (b.reference_name, b.start, b.end, b.reference_bases, b.alternate_bases, b.gq, b.call_set_name) == ('chr1', 100, 101, base, ['<*>'], 1, 's')
  np.testing.assert_allclose(b.genotype_likelihood, [-0.47712125472] * 3)
  assert b.info == ({'MIN_DP': [0], 'MED_DP': [0]} if include_med_dp else {'MIN_DP': [0]})


@pytest.mark.parametrize('base', 'NRWB')
def test_gvcf_basic_skips_iupac_ref_base(base):
  assert _blocks([(0, 0, base)], start=100) == []


@pytest.mark.parametrize('base', 'X>!')
def test_gvcf_basic_raises_with_bad_ref_base(base):
  with pytest.raises(ValueError, match='Invalid reference base='):
    _blocks([(0, 0, base)], start=100)


@pytest.mark.parametrize('counts,want', [
    ([(0, 0, 'A')], [(1, 2, 'A', 1, 0)]),
    ([(0, 0, 'A'), (0, 0, 'C')], [(1, 3, 'A', 1, 0)]),
    ([(0, 0, 'C'), (0, 0, 'A')], [(1, 3, 'C', 1, 0)]),
    ([(0, 0, 'A'), (0, 0, 'C'), (0, 0, 'T')], [(1, 4, 'A', 1, 0)]),
    ([(0, 0, 'A'), (0, 100, 'C')], [(1, 2, 'A', 1, 0), (2, 3, 'C', 100, 100)]),
    ([(0, 100, 'A'), (0, 0, 'C')], [(1, 2, 'A', 100, 100), (2, 3, 'C', 1, 0)]),
    ([(0, 0, 'A'), (0, 20, 'C'), (0, 100, 'T')], [(1, 2, 'A', 1, 0), (2, 3, 'C', 59, 20), (3, 4, 'T', 100, 100)]),
])
def test_make_gvcfs(counts, want):
  assert [(b.start, b.end, b.reference_bases, b.gq, b.info['MIN_DP'][0]) for b in _blocks(counts)] == want


QUANTIZE_COUNTS = [(0, 18, 'A'), (0, 19, 'C'), (35, 0, 'A'), (10, 10, 'T'), (4, 12, 'A'), (1, 30, 'A'), (1, 34, 'C'), (0, 20, 'T'), (0, 19, 'G')]
_SINGLES = [(1, 2, 'A', 53, 18, 18), (2, 3, 'C', 56, 19, 19), (3, 4, 'A', 0, 35, 35), (4, 5, 'T', 0, 20, 20), (5, 6, 'A', 0, 16, 16), (6, 7, 'A', 72, 31, 31),
            (7, 8, 'C', 83, 35, 35), (8, 9, 'T', 59, 20, 20), (9, 10, 'G', 56, 19, 19)]


@pytest.mark.parametrize('gq_resolution,want', [
    (1, _SINGLES),
    (3, _SINGLES),
    (4, [(1, 3, 'A', 53, 18, 18)] + _SINGLES[2:]),
    (10, [(1, 3, 'A', 53, 18, 18)] + _SINGLES[2:7] + [(8, 10, 'T', 56, 19, 19)]),
    (45, [(1, 3, 'A', 53, 18, 18)] + _SINGLES[2:5] + [(6, 10, 'A', 56, 25, 19)]),
])
def test_quantize_gvcfs(gq_resolution, want):
  got = _blocks(QUANTIZE_COUNTS, gq_resolution=gq_resolution, include_med_dp=True)
  assert [(b.start, b.end, b.reference_bases, b.gq, b.info['MED_DP'][0], b.info['MIN_DP'][0]) for b in got] == want
  # het / hom-alt sites are never merged and are not called 0/0
  assert [b.genotype for b in got if b.gq == 0] == [[-1, -1]] * 3


def test_record_round_trip():
  for b in _blocks(QUANTIZE_COUNTS, include_med_dp=True):
    assert gvcf.parse_variant_record(gvcf.serialize_gvcf_record(b)) == b


def test_transform_to_gvcf():
  v = pp.OutVariant('chr1', 10, 11, 'A', ['C', 'G'], {'AD': [3, 4, 5], 'DP': [12], 'VAF': [0.33, 0.41]}, genotype=[1, 2],
                    genotype_likelihood=[-3.0, -2.0, -4.0, -1.0, -0.5, -6.0], gq=20, quality=30.0, filter=['PASS'])
  g = gvcf.transform_to_gvcf(v)
  assert g.alternate_bases == ['C', 'G', '<*>'] and g.info['AD'] == [3, 4, 5, 0] and g.info['VAF'] == [0.33, 0.41, 0.0]
  assert g.genotype_likelihood == [-2.5, -1.5, -3.5, -0.5, 0.0, -5.5, -99.0, -99.0, -99.0, -99.0]
  assert v.alternate_bases == ['C', 'G'] and len(v.genotype_likelihood) == 6           # the VCF record is untouched
  assert gvcf.transform_to_gvcf(g).alternate_bases == ['C', 'G', '<*>']               # idempotent
  line = gvcf.gvcf_line(g).split('\t')
  assert line[4] == 'C,G,<*>' and line[7] == '.' and line[8] == 'GT:GQ:DP:AD:VAF:PL'
  assert line[9] == '1/2:20:12:3,4,5,0:0.33,0.41,0:25,15,35,5,0,55,990,990,990,990'


def test_merge_splits_blocks_around_variants():
  base_at = lambda c, p: 'ACGT'[p % 4]
  block = lambda s, e, c='chr1': pp.OutVariant(c, s, e, base_at(c, s), ['<*>'], {'MIN_DP': [7]}, genotype=[0, 0], genotype_likelihood=[0.0, -1.0, -2.0], gq=5)
  var = lambda s, e, c='chr1': pp.OutVariant(c, s, e, 'A' * (e - s), ['T'], {}, genotype=[0, 1], genotype_likelihood=[-1.0, 0.0, -2.0], gq=9, quality=3.0)
  merged = list(gvcf.merge_variants_and_nonvariants([var(5, 6), var(12, 15), var(3, 4, 'chr2')], [block(0, 10), block(10, 13), block(13, 14), block(14, 20), block(0, 8, 'chr2')],
                                                    ['chr1', 'chr2'], base_at))
  got = [(v.reference_name, v.start, v.end, v.reference_bases if v.alternate_bases == ['<*>'] else 'VAR') for v in merged]
  assert got == [('chr1', 0, 5, 'A'), ('chr1', 5, 6, 'VAR'), ('chr1', 6, 10, 'G'), ('chr1', 10, 12, 'G'), ('chr1', 12, 15, 'VAR'), ('chr1', 15, 20, 'T'),
                 ('chr2', 0, 3, 'A'), ('chr2', 3, 4, 'VAR'), ('chr2', 4, 8, 'A')]
  assert [v.end for v in merged if v.alternate_bases == ['<*>']] == [5, 10, 12, 20, 3, 8]
  # only variants / only blocks
  assert len(list(gvcf.merge_variants_and_nonvariants([var(5, 6)], [], ['chr1'], base_at))) == 1
  assert len(list(gvcf.merge_variants_and_nonvariants([], [block(0, 10)], ['chr1'], base_at))) == 1


@pytest.mark.parametrize('cvo,blocks,golden', [
    ('golden.postprocess_single_site_input-00000-of-00001.tfrecord.gz', 'golden.postprocess_gvcf_input.tfrecord.gz', 'golden.postprocess_gvcf_output.g.vcf.gz'),
    ('golden.postprocess_pacbio_input-00000-of-00001.tfrecord.gz', 'golden.postprocess_pacbio_gvcf_input.tfrecord.gz', 'golden.postprocess_gvcf_output_pacbio.g.vcf.gz'),
])
def test_postprocess_gvcf_golden_byte_for_byte(tmp_path, cvo, blocks, golden):
  want = gzip.open(os.path.join(GOLDEN, golden), 'rt').read().splitlines()
  contigs = [(l.split('ID=')[1].split(',')[0], int(l.split('length=')[1].rstrip('>'))) for l in want if l.startswith('##contig')]

  def base_at(contig, pos):
    raise AssertionError('the golden blocks need no split inside a block whose base is unknown here')

  # blocks are split where a variant cuts them: the new first base comes from the reference; take it from the golden itself
  bases = {(l.split('\t')[0], int(l.split('\t')[1]) - 1): l.split('\t')[3][0] for l in want if not l.startswith('#')}
  r = pp.postprocess_variants(os.path.join(GOLDEN, cvo), str(tmp_path / 'o.vcf'), contigs, nonvariant_site_tfrecord_path=os.path.join(GOLDEN, blocks),
                              gvcf_outfile=str(tmp_path / 'o.g.vcf'), base_at=lambda c, p: bases[(c, p)])
  got = open(tmp_path / 'o.g.vcf').read().splitlines()
  assert got == want
  assert r['n_gvcf_records_written'] == sum(1 for l in want if not l.startswith('#'))


def test_gvcf_needs_both_paths(tmp_path):
  with pytest.raises(ValueError, match='both'):
    pp.postprocess_variants(os.path.join(GOLDEN, 'golden.postprocess_single_site_input-00000-of-00001.tfrecord.gz'), str(tmp_path / 'o.vcf'), [('chr20', 1)],
                            gvcf_outfile=str(tmp_path / 'o.g.vcf'))


def test_gvcf_golden_report_is_current():
  r = json.load(open(os.path.join(GOLDEN, 'gvcf_golden_report.json')))
  assert r['make_examples_wgs']['all_equal'] and r['make_examples_wgs']['golden_records'] == 235
  assert r['make_examples_pacbio']['all_equal'] and r['make_examples_pacbio']['golden_records'] == 1496
  assert all(v['byte_identical'] for v in r['postprocess'].values()) and 'wgs_med_dp_from_bam' in r['postprocess']


def test_cli_bam_to_gvcf_cpu_plumbing(tmp_path, monkeypatch):
  """make_examples --gvcf -> (stand-in call_variants) -> postprocess_variants --gvcf_outfile through the stage CLIs: the g.vcf tiles
  the calling region exactly - every position in one record, variants and blocks interleaved in order.  The encoder is the CPU
  oracle here (as in tests/test_candidates.py); tests/test_zz_allele_count_gpu.py runs run_deepvariant --output_gvcf on the GPU."""
  import test_candidates as tc
  from deepvariant_b200 import cli, make_examples_native as men, pileup_image as pi, protos
  monkeypatch.setattr(men.ExamplesGenerator, '_gpu', lambda self: tc.OracleEncoder(pi.to_params(self.options.pic_options, height=self.pileup_image_height)))
  fa, bam_path, genome, sites = tc._planted_case(tmp_path)
  ex = str(tmp_path / 'ex.tfrecord@2.gz')
  blocks = str(tmp_path / 'gvcf.tfrecord@2.gz')
  for task in (0, 1):
    assert cli.make_examples(['--mode', 'calling', '--ref', fa, '--reads', bam_path, '--examples', ex, '--gvcf', blocks, '--task', str(task),
                              '--channel_list', 'BASE_CHANNELS,insert_size', '--regions', 'chr20:1-6000', '--norealign_reads',
                              '--runtime_by_region', str(tmp_path / 'runtime.tsv@2')]) == 0
  # --runtime_by_region: the reference's columns, one line per region of the task, counts that add up
  rows = []
  for task in (0, 1):
    lines = open(tmp_path / f'runtime.tsv-0000{task}-of-00002').read().splitlines()
    assert lines[0].split('\t') == list(cli.RUNTIME_BY_REGION_COLUMNS) and len(lines) == 1 + 3
    rows += [dict(zip(cli.RUNTIME_BY_REGION_COLUMNS, l.split('\t'))) for l in lines[1:]]
  assert sorted(r['region'] for r in rows) == sorted(f'chr20:{s + 1}-{s + 1000}' for s in range(0, 6000, 1000))
  assert sum(int(r['num candidates']) for r in rows) == 4 == sum(int(r['num examples']) for r in rows)
  assert all(float(r['find candidates']) >= 0 for r in rows if int(r['num reads'])) and rows[0]['small model total'] == 'NA'
  cvo_path = str(tmp_path / 'cvo.tfrecord.gz')
  with tfrecord.Writer(cvo_path) as w:
    for p in tfrecord.resolve_input_paths(ex):
      for r in tfrecord.read_records(p):
        e = protos.parse_tf_example(r)
        w.write(protos.encode_call_variants_output(e['variant/encoded'][1][0], protos.parse_alt_allele_indices(e['alt_allele_indices/encoded'][1][0]),
                                                   [0.01, 0.9, 0.09]))
  out, gout = str(tmp_path / 'o.vcf'), str(tmp_path / 'o.g.vcf')
  assert cli.postprocess_variants(['--ref', fa, '--infile', cvo_path, '--outfile', out, '--nonvariant_site_tfrecord_path', blocks, '--gvcf_outfile', gout]) == 0
  recs = [l.split('\t') for l in open(gout) if not l.startswith('#')]
  variants = [l.split('\t') for l in open(out) if not l.startswith('#')]
  assert [int(v[1]) - 1 for v in variants] == sorted(sites.values())
  nxt = 1
  for r in recs:
    assert int(r[1]) == nxt, r[:5]
    nxt = int(r[7][4:]) + 1 if r[7].startswith('END=') else int(r[1]) + len(r[3])
    assert r[4].endswith('<*>') and r[3][0] == genome[int(r[1]) - 1]
  assert nxt == 6001
  assert sum(1 for r in recs if r[4] != '<*>') == 4
  # the regions without reads (first and last kilobase) are blocks of GQ 1, depth 0 (no early exit with --gvcf)
  assert recs[0][9].split(':')[:3] == ['0/0', '1', '0'] and recs[-1][9].split(':')[:3] == ['0/0', '1', '0']
  assert all(l.split('\t')[9].strip() for l in open(gout) if l.startswith('#CHROM')) and '\tplanted\n' in open(gout).read()


# ---- --haploid_contigs / --par_regions_bed (postprocess_variants.py:1070-1112) ------------------------------------------------------------
@pytest.mark.parametrize('probabilities,n_alts,want', [
    ([0.98, 0.02, 0], 1, [1.0, 0, 0]),
    ([0.2, 0.5, 0.3], 1, [0.4, 0, 0.6]),
    ([0.0, 1.0, 0.0], 1, [0, 0, 0]),
    ([0.02, 0.03, 0.45, 0.07, 0.3, 0.13], 2, [0.033, 0, 0.75, 0, 0, 0.216]),
])
def test_correct_nonautosome_probabilities(probabilities, n_alts, want):
  # postprocess_variants_test.py:2140-2184
  np.testing.assert_allclose(pp.correct_nonautosome_probabilities(probabilities, n_alts), want, atol=1e-3)


def test_haploid_goldens_byte_for_byte(tmp_path):
  """golden.haploid_chr20.*: the WGS golden CVOs + blocks with --haploid_contigs chr20 (scripts/create_golden.sh:293-305)."""
  want_g = gzip.open(os.path.join(GOLDEN, 'golden.haploid_chr20.postprocess_gvcf_output.g.vcf.gz'), 'rt').read().splitlines()
  want_v = gzip.open(os.path.join(GOLDEN, 'golden.haploid_chr20.postprocess_single_site_output.vcf.gz'), 'rt').read().splitlines()
  contigs = [(l.split('ID=')[1].split(',')[0], int(l.split('length=')[1].rstrip('>'))) for l in want_g if l.startswith('##contig')]
  bases = {(l.split('\t')[0], int(l.split('\t')[1]) - 1): l.split('\t')[3][0] for l in want_g if not l.startswith('#')}
  kw = dict(nonvariant_site_tfrecord_path=os.path.join(GOLDEN, 'golden.postprocess_gvcf_input.tfrecord.gz'), gvcf_outfile=str(tmp_path / 'o.g.vcf'),
            base_at=lambda c, p: bases[(c, p)])
  cvo = os.path.join(GOLDEN, 'golden.postprocess_single_site_input-00000-of-00001.tfrecord.gz')
  pp.postprocess_variants(cvo, str(tmp_path / 'o.vcf'), contigs, haploid_contigs='chr20', **kw)
  assert open(tmp_path / 'o.vcf').read().splitlines() == want_v
  assert open(tmp_path / 'o.g.vcf').read().splitlines() == want_g
  # no heterozygous genotype survives on a haploid contig ...
  gts = [l.split('\t')[9].split(':')[0] for l in want_v if not l.startswith('#')]
  assert not any(g in ('0/1', '1/2', '0/2') for g in gts)
  # ... unless the site lies in a pseudo-autosomal region: with all of chr20 declared PAR the diploid output comes back
  bed = tmp_path / 'par.bed'
  bed.write_text('chr20\t0\t63025520\n')
  pp.postprocess_variants(cvo, str(tmp_path / 'p.vcf'), contigs, haploid_contigs='chr20', par_regions_bed=str(bed), **kw)
  diploid = open(os.path.join(GOLDEN, 'golden.postprocess_single_site_output.vcf')).read().splitlines()
  assert open(tmp_path / 'p.vcf').read().splitlines() == diploid
