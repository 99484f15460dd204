"""postprocess_variants (deepvariant_b200/postprocess_variants.py): CallVariantsOutput -> VCF.

Pinned by the reference's own golden pairs, byte for byte (header included): golden.postprocess_single_site_input ->
golden.postprocess_single_site_output.vcf (78 records incl. multi-allelic merging, allele pruning, RefCall / NoCall), its
--only_keep_pass variant, the vcf_candidate_importer pair (--nogroup_variants) and the PACBIO pair (341 records, phased
genotypes from ALT_PS); plus the unit KATs of deepvariant/postprocess_variants_test.py and haplotypes_test.py transcribed
as data (values recomputed with the reference's formulas).  CPU-only (host code)."""
import os

import numpy as np
import pytest

from deepvariant_b200 import postprocess_variants as pp

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
CHR20 = [('chr20', 63025520)]


@pytest.mark.parametrize('infile,vcf,kwargs', [
    ('golden.postprocess_single_site_input-00000-of-00001.tfrecord.gz', 'golden.postprocess_single_site_output.vcf', {}),
    ('golden.postprocess_single_site_input-00000-of-00001.tfrecord.gz', 'golden.postprocess_single_site_output.pass_only.vcf',
     {'only_keep_pass': True}),
    ('golden.vcf_candidate_importer_postprocess_single_site_input-00000-of-00001.tfrecord.gz',
     'golden.vcf_candidate_importer_postprocess_single_site_output.vcf', {'group_variants': False}),
    ('golden.postprocess_pacbio_input-00000-of-00001.tfrecord.gz', 'golden.postprocess_single_site_output_pacbio.vcf', {}),
])
def test_reference_golden_vcfs_reproduced_byte_for_byte(tmp_path, infile, vcf, kwargs):
  out = str(tmp_path / 'out.vcf')
  r = pp.postprocess_variants(os.path.join(GOLDEN, infile), out, CHR20, **kwargs)
  got, want = open(out).read(), open(os.path.join(GOLDEN, vcf)).read()
  assert got == want
  assert r['n_variants_written'] == sum(1 for line in want.split('\n') if line and not line.startswith('#')) > 20


def test_group_variants_rejects_two_variants_on_one_range():
  """postprocess_variants_test.py:444-453: with --group_variants the importer's CVOs fail merge_predictions' sanity check."""
  with pytest.raises(ValueError, match='sanity check'):
    pp.postprocess_variants(os.path.join(GOLDEN, 'golden.vcf_candidate_importer_postprocess_single_site_input-00000-of-00001.tfrecord.gz'),
                            os.devnull, CHR20)


def test_gzip_output_and_cli(tmp_path):
  import gzip
  from deepvariant_b200 import cli
  fa = tmp_path / 'ref.fa'
  fa.write_text('>chr20\nACGT\n')
  (tmp_path / 'ref.fa.fai').write_text('chr20\t63025520\t7\t60\t61\n')
  out = str(tmp_path / 'o.vcf.gz')
  assert cli.postprocess_variants(['--ref', str(fa), '--infile', os.path.join(GOLDEN, 'golden.postprocess_single_site_input-00000-of-00001.tfrecord.gz'),
                                   '--outfile', out]) == 0
  assert gzip.open(out, 'rt').read() == open(os.path.join(GOLDEN, 'golden.postprocess_single_site_output.vcf')).read()


def _variant(ref='A', alts=('C',), start=10, ad=(10, 10), **kw):
  return pp.OutVariant('chr1', start, start + len(ref), ref, list(alts), {'AD': list(ad), 'DP': [sum(ad)]}, **kw)


@pytest.mark.parametrize('probs,expected', [                     # postprocess_variants_test.py test_most_likely_genotype
    ([0.9, 0.05, 0.05], (0, [0, 0])), ([0.05, 0.9, 0.05], (1, [0, 1])), ([0.05, 0.05, 0.9], (2, [1, 1])),
    ([0.7, 0.1, 0.1, 0.05, 0.04, 0.01], (0, [0, 0])), ([0.1, 0.1, 0.1, 0.05, 0.64, 0.01], (4, [1, 2])), ([0, 0, 0, 0, 0, 1], (5, [2, 2])),
])
def test_most_likely_genotype(probs, expected):
  n_alleles = 2 if len(probs) == 3 else 3
  assert pp.most_likely_genotype(probs, n_alleles=n_alleles) == expected


@pytest.mark.parametrize('probs,index,gq,qual', [               # test_compute_quals (postprocess_variants_test.py)
    ([0.01, 0.98, 0.01], 1, 17, 20), ([0.01, 0.01, 0.98], 2, 17, 20), ([0.001, 0.99, 0.009], 1, 20, 30), ([0.9, 0.05, 0.05], 0, 10, 0.4575749),
    ([1.0, 0.0, 0.0], 0, 99, 0.0), ([0.0, 1.0, 0.0], 1, 99, 99.0308995),      # capped by _MAX_CONFIDENCE = 1 - 1.25e-10
])
def test_compute_quals(probs, index, gq, qual):
  got_gq, got_qual = pp.compute_quals(probs, index)
  assert got_gq == gq and abs(got_qual - qual) < 1e-6 and got_qual == round(got_qual, 7)


def test_add_call_to_variant_filters():
  v = pp.add_call_to_variant(_variant(), [0.001, 0.999 - 1e-3, 1e-3], sample_name='s')
  assert v.genotype == [0, 1] and v.filter == ['PASS'] and v.call_set_name == 's'
  v = pp.add_call_to_variant(_variant(), [0.999, 0.0005, 0.0005])
  assert v.genotype == [0, 0] and v.filter == ['RefCall'] and v.gq == 30
  v = pp.add_call_to_variant(_variant(), [0.9, 0.05, 0.05])                       # hom-ref with GQ 10 < 20 -> ./. NoCall
  assert v.genotype == [-1, -1] and v.filter == ['NoCall'] and v.gq == 10
  v = pp.add_call_to_variant(_variant(), [0.85, 0.1, 0.05], qual_filter=1.0)
  assert v.filter == ['NoCall']
  v = pp.add_call_to_variant(_variant(), [0.45, 0.5, 0.05], qual_filter=10.0)     # variant call with QUAL 2.6 < 10
  assert v.filter == ['LowQual'] and v.genotype == [0, 1]
  v = pp.add_call_to_variant(_variant(ad=(0, 0)), [0.01, 0.98, 0.01])             # uncall_gt_if_no_ad
  assert v.genotype == [-1, -1] and v.gq == 0 and v.genotype_likelihood == [0, 0] and v.filter == ['NoCall']


def test_simplify_alleles_and_prune():
  assert pp.simplify_alleles('CAA', 'CA') == ('CA', 'C')                            # variant_utils_test.py simplify cases
  assert pp.simplify_alleles('ATT', 'TT') == ('AT', 'T')
  assert pp.simplify_alleles('AT', 'A', 'ATT') == ('AT', 'A', 'ATT')
  assert pp.simplify_alleles('CACA', 'CA', 'CCA') == ('CAC', 'C', 'CC')
  v = pp.OutVariant('1', 5, 6, 'A', ['C', 'G', 'T'], {'AD': [1, 2, 3, 4], 'VAF': [0.2, 0.3, 0.4], 'DP': [10]})
  p = pp.prune_alleles(v, {'G'})
  assert p.alternate_bases == ['C', 'T'] and p.info == {'AD': [1, 2, 4], 'VAF': [0.2, 0.4], 'DP': [10]}
  assert v.alternate_bases == ['C', 'G', 'T']                                       # the input is not modified


def _cvo(v, idx, probs):
  return pp.Cvo(v, idx, probs)


def test_merge_predictions_multiallelic():
  base = _variant(alts=('C', 'G'), ad=(10, 10, 10))
  cvos = [_cvo(base, [0], [0.01, 0.98, 0.01]), _cvo(base, [1], [0.01, 0.98, 0.01]), _cvo(base, [0, 1], [0.01, 0.01, 0.98])]
  v, p = pp.merge_predictions(cvos)
  assert v.alternate_bases == ['C', 'G'] and len(p) == 6 and abs(sum(p) - 1) < 1e-12
  assert int(np.argmax(p)) == 4                                                      # 1/2
  # an alt whose 1 - p(ref) falls below the filter is pruned and its examples are ignored
  cvos = [_cvo(base, [0], [0.01, 0.98, 0.01]), _cvo(base, [1], [0.999, 0.0005, 0.0005]), _cvo(base, [0, 1], [0.01, 0.98, 0.01])]
  v, p = pp.merge_predictions(cvos)
  assert v.alternate_bases == ['C'] and v.info['AD'] == [10, 10] and len(p) == 3 and int(np.argmax(p)) == 1
  with pytest.raises(ValueError):
    pp.merge_predictions([_cvo(base, [0], [0.01, 0.98, 0.01]), _cvo(base, [0, 1], [0.01, 0.01, 0.98])])   # missing [1]
  _, p_min = pp.merge_predictions([_cvo(base, [0], [0.1, 0.8, 0.1]), _cvo(base, [1], [0.2, 0.7, 0.1]), _cvo(base, [0, 1], [0.05, 0.05, 0.9])],
                                  multiallelic_mode='min')
  assert len(p_min) == 6 and abs(sum(p_min) - 1) < 1e-12


def test_haplotype_resolution():
  """haplotypes_test.py: two overlapping hom-alt calls cannot both be true; compatible calls pass through unchanged."""
  def called(start, end, gt, gls):
    v = pp.OutVariant('1', start, end, 'A' * (end - start), ['T'], {'AD': [5, 5]}, genotype=list(gt), genotype_likelihood=list(gls), quality=30.0)
    v.filter = ['PASS']
    return v
  a, b = called(10, 14, (0, 1), [-2, -0.01, -2]), called(12, 13, (0, 1), [-2, -0.01, -2])
  assert [(v.start, v.genotype) for v in pp.maybe_resolve_conflicting_variants([a, b])] == [(10, [0, 1]), (12, [0, 1])]
  a, b = called(10, 14, (1, 1), [-3, -1, -0.05]), called(12, 13, (1, 1), [-3, -0.3, -0.35])
  got = list(pp.maybe_resolve_conflicting_variants([a, b]))
  # joint log10 likelihoods: het + het = -1.3 beats hom-alt + hom-ref = -3.05 and het + hom-ref = -4; 1/1 + anything else is invalid
  assert got[0].genotype == [0, 1] and got[1].genotype == [0, 1] and got[0].filter == ['PASS']
  assert abs(sum(10 ** g for g in got[1].genotype_likelihood) - 1) < 1e-6
  c = called(30, 31, (1, 1), [-3, -1, -0.05])
  assert list(pp.maybe_resolve_conflicting_variants([c]))[0] is c


def test_vcf_number_formatting():
  v = pp.OutVariant('chr1', 9, 10, 'A', ['C'], {'DP': [3], 'AD': [1, 2], 'VAF': [2 / 3]}, call_set_name='s', genotype=[0, 1],
                    genotype_likelihood=[-3.76, -0.0004, -4.2], gq=36, quality=37.649, filter=['PASS'])
  assert pp.vcf_line(v) == 'chr1\t10\t.\tA\tC\t37.6\tPASS\t.\tGT:GQ:DP:AD:VAF:PL\t0/1:36:3:1,2:0.666667:37,0,41'
  v.quality, v.is_phased, v.genotype = 0.05, True, [1, 0]
  assert pp.vcf_line(v).split('\t')[5] == '0.1' and pp.vcf_line(v).split('\t')[9].startswith('1|0:')
  v.quality = 0.04
  assert pp.vcf_line(v).split('\t')[5] == '0'


# ---- --cpus: the conversion over worker processes gives the same bytes --------------------------------------------------------------

def test_independent_chunks_never_cut_through_overlapping_ranges():
  keys = [('a', 0, 5), ('a', 3, 4), ('a', 4, 9), ('a', 9, 10), ('a', 9, 10), ('a', 12, 13), ('b', 1, 2), ('b', 1, 8), ('b', 8, 9)]
  chunks = pp.independent_chunks(keys, 1)
  assert chunks == [(0, 3), (3, 5), (5, 6), (6, 8), (8, 9)]      # [0,5) [3,4) [4,9) overlap; the two equal ranges stay together; contig change cuts
  assert pp.independent_chunks(keys, 100) == [(0, 9)] and pp.independent_chunks([], 3) == []
  rng = np.random.RandomState(1)
  starts = np.sort(rng.randint(0, 3000, 500))
  keys = [('c', int(s), int(s + rng.randint(1, 30))) for s in starts]
  chunks = pp.independent_chunks(keys, 7)
  assert chunks[0][0] == 0 and chunks[-1][1] == 500 and all(a[1] == b[0] for a, b in zip(chunks, chunks[1:]))
  for b, e in chunks[1:]:
    assert keys[b][1] >= max(k[2] for k in keys[:b])               # nothing before the cut reaches past it


def test_cvo_range_key_equals_the_parsed_variant_range():
  for name in ('golden.postprocess_single_site_input-00000-of-00001.tfrecord.gz', 'golden.postprocess_pacbio_input-00000-of-00001.tfrecord.gz'):
    from deepvariant_b200 import tfrecord
    for r in tfrecord.read_records(os.path.join(GOLDEN, name)):
      v = pp.parse_cvo(r).variant
      assert pp.cvo_range_key(r) == (v.reference_name, v.start, v.end)


@pytest.mark.parametrize('name,vcf', [('golden.postprocess_single_site_input-00000-of-00001.tfrecord.gz', 'golden.postprocess_single_site_output.vcf'),
                                      ('golden.postprocess_pacbio_input-00000-of-00001.tfrecord.gz', None)])
def test_worker_processes_give_the_serial_output_byte_for_byte(tmp_path, name, vcf):
  contigs = CHR20
  serial, parallel = str(tmp_path / 'serial.vcf'), str(tmp_path / 'parallel.vcf')
  a = pp.postprocess_variants(os.path.join(GOLDEN, name), serial, contigs)
  b = pp.postprocess_variants(os.path.join(GOLDEN, name), parallel, contigs, cpus=3, chunk_records=9)     # many small chunks
  assert a == b and a['n_variants_written'] > 50
  assert open(serial).read() == open(parallel).read()
  if vcf:
    assert open(parallel).read() == open(os.path.join(GOLDEN, vcf)).read()                               # ... which is the reference's golden VCF
