"""Pins against the REFERENCE'S OWN golden outputs (fixtures derived by tools/make_golden_fixtures.py from
deepvariant/testdata/golden.calling_{candidates,examples}.tfrecord.gz + NA12878_S1.chr20.10_10p1mb.bam):
the 7 golden pileup images whose reads the (out-of-scope) realigner did not rewrite must be reproduced
byte-for-byte by the oracle (CPU) and by the CUDA encoder (GPU) from the packed candidate + read inputs."""
import json
import os

import numpy as np
import pytest

import oracle_lib
from deepvariant_b200 import packing, protos, tfrecord
from deepvariant_b200 import pileup_image as pi

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def _fixture():
  d = np.load(os.path.join(GOLDEN, 'wgs_golden_subset.npz'))
  arrays = {k[4:]: d[k] for k in d.files if k.startswith('arr_')}
  pb = packing.PackedBatch(int(d['n_images']), int(d['n_reads']), int(d['n_pairs']), int(d['ref_stride']), arrays)
  o = pi.default_options(pi.ReadRequirements(min_base_quality=10, min_mapping_quality=5))
  o.channels = list(pi.PILEUP_CHANNELS_WITH_INSERT_SIZE)
  return pb, d['golden_images'], pi.to_params(o)


def test_oracle_reproduces_reference_golden_images():
  pb, golden, params = _fixture()
  assert golden.shape == (7, 100, 221, 7)
  got = oracle_lib.encode_batch(params, pb)
  np.testing.assert_array_equal(got, golden)


@pytest.mark.gpu
def test_cuda_encoder_reproduces_reference_golden_images():
  pb, golden, params = _fixture()
  enc = pi.GpuEncoder(params, 0)
  got = enc.encode_host(pb)
  np.testing.assert_array_equal(got, golden)
  assert enc.last_rows_kept[:7].tolist() == [int(g[5:].reshape(95, -1).any(1).sum()) for g in golden]


def test_golden_report_numbers():
  """Regression pin of the per-example match report: all 84 reference bands, 7 whole images and >= 80 % of
  the 4309 golden read rows are reproduced; the rest are reads the reference realigner rewrote."""
  r = json.load(open(os.path.join(GOLDEN, 'wgs_golden_report.json')))
  assert r['n_examples'] == 84 and r['n_exact'] == 7
  assert all(e['ref_band_equal'] for e in r['examples'])
  assert r['golden_read_rows'] == 4309 and r['golden_read_rows_reproduced'] >= 3467


def test_reference_tf_examples_parse_and_reencode():
  """The 7-feature tf.Example of make_examples (make_examples_native.cc:388-474) through our wire codec."""
  recs = list(tfrecord.read_records(os.path.join(GOLDEN, 'golden.calling_examples.first3.tfrecord.gz'), check_crc=True))
  assert len(recs) == 3
  for rec in recs:
    ex = protos.parse_tf_example(rec)
    assert sorted(ex) == ['alt_allele_indices/encoded', 'image/encoded', 'image/shape', 'locus', 'sequencing_type',
                          'variant/encoded', 'variant_type']
    assert ex['image/shape'][1] == [100, 221, 7] and len(ex['image/encoded'][1][0]) == 154700
    v = protos.parse_variant(ex['variant/encoded'][1][0])
    assert ex['locus'][1][0].decode() == f'{v.reference_name}:{v.start + 1}-{v.end}'
    assert protos.parse_alt_allele_indices(ex['alt_allele_indices/encoded'][1][0]) == [0]
    again = protos.parse_tf_example(protos.encode_tf_example(ex))
    assert again == ex
  img = np.frombuffer(protos.parse_tf_example(recs[0])['image/encoded'][1][0], np.uint8).reshape(100, 221, 7)
  assert (img[:5, :, 1:] == np.array([254, 254, 70, 152, 50, 254], np.uint8)).all()


def test_reference_candidates_parse():
  cands = [protos.parse_deepvariant_call(r) for r in
           tfrecord.read_records(os.path.join(GOLDEN, 'golden.calling_candidates.first8.tfrecord.gz'), check_crc=True)]
  assert len(cands) == 8
  c = cands[0]
  assert (c.variant.reference_name, c.variant.start, c.variant.end, c.variant.reference_bases, c.variant.alternate_bases) == (
      'chr20', 10000116, 10000117, 'C', ['T'])
  assert len(c.allele_support['T']) == 30 and all('/' in n for n in c.allele_support['T'])


def test_tfrecord_roundtrip_and_sharding(tmp_path):
  spec = str(tmp_path / 'ex.tfrecord@3.gz')
  paths = tfrecord.shard_paths(spec)
  assert [os.path.basename(p) for p in paths] == ['ex.tfrecord-00000-of-00003.gz', 'ex.tfrecord-00001-of-00003.gz',
                                                   'ex.tfrecord-00002-of-00003.gz']
  payload = [os.urandom(n) for n in (0, 1, 17, 70000)]
  with tfrecord.Writer(paths[1]) as w:
    for p in payload:
      w.write(p)
  assert list(tfrecord.read_records(paths[1], check_crc=True)) == payload
  assert tfrecord.masked_crc32c(b'') == ((0 >> 15 | 0 << 17) + 0xa282ead8) & 0xFFFFFFFF
  # CRC-32C check value (RFC 3720 appendix B.4)
  from deepvariant_b200 import _lib
  assert _lib.lib().dvb_crc32c(b'123456789', 9) == 0xE3069283


# ---- the down-sampling path (more than 95 reads overlap the candidate) pinned by the reference's golden file -------------

def _downsample_fixture():
  """tools/check_downsample_golden.py: 12 of the 51 down-sampled examples of the reference's
  golden.allele_frequency_examples.tfrecord.gz (channels 0-6 of its 100x221x8 images) with the reads of its BAM.  No
  candidates file ships for that golden, so read support is unknown: channel 4 is not compared."""
  d = np.load(os.path.join(GOLDEN, 'downsample_golden_subset.npz'))
  arrays = {k[4:]: d[k] for k in d.files if k.startswith('arr_')}
  pb = packing.PackedBatch(int(d['n_images']), int(d['n_reads']), int(d['n_pairs']), int(d['ref_stride']), arrays)
  o = pi.default_options(pi.ReadRequirements(min_base_quality=10, min_mapping_quality=5))
  o.channels = list(pi.PILEUP_CHANNELS_WITH_INSERT_SIZE)
  return pb, d['golden_images'], d['channels'].tolist(), o


def test_oracle_reproduces_downsampled_golden_rows_in_order():
  pb, golden, ch, o = _downsample_fixture()
  assert golden.shape == (12, 100, 221, 7) and ch == [0, 1, 2, 3, 5, 6]
  n_reads = np.diff(pb.arrays['pair_begin'])
  assert (n_reads > 95).all() and n_reads.max() >= 190          # every image is down-sampled; some from ~200 reads
  got = oracle_lib.encode_batch(pi.to_params(o), pb)
  np.testing.assert_array_equal(got[..., ch], golden[..., ch])  # all 100 rows, in the reference's order
  # ... and NOT with libstdc++'s std::shuffle (what a gcc build would do): the golden pins the standard library
  p1 = pi.to_params(o)
  p1.shuffle_stdlib = 1
  other = oracle_lib.encode_batch(p1, pb)
  assert not any(np.array_equal(other[i][..., ch], golden[i][..., ch]) for i in range(12))


@pytest.mark.gpu
def test_cuda_encoder_reproduces_downsampled_golden_rows_in_order():
  pb, golden, ch, o = _downsample_fixture()
  enc = pi.GpuEncoder(pi.to_params(o), 0)
  got = enc.encode_host(pb)
  np.testing.assert_array_equal(got[..., ch], golden[..., ch])
  assert enc.last_rows_kept[:12].tolist() == [95] * 12
  p1 = pi.to_params(o)
  p1.shuffle_stdlib = 1
  enc1 = pi.GpuEncoder(p1, 0)
  np.testing.assert_array_equal(enc1.encode_host(pb), oracle_lib.encode_batch(p1, pb))   # the libstdc++ flavour stays bit-exact vs the oracle


def test_downsample_report_numbers():
  r = json.load(open(os.path.join(GOLDEN, 'downsample_golden_report.json')))
  assert r['n_examples'] == 78 and r['n_downsampled'] == 51 and r['n_downsampled_exact'] == 51
  assert r['downsampled_rows'] == r['downsampled_rows_equal_in_place'] == 4845
  assert all(e['exact_six_channels'] for e in r['examples'])


# ---- trimmed long reads (PACBIO path: TrimReads / TrimCigar + the window clip) against the reference's PacBio golden ---------

def _pacbio_fixture():
  """tools/check_pacbio_golden.py: 12 of the 401 examples of the reference's golden.pacbio_examples.tfrecord.gz (first seven of
  its ten channels) with the TRIMMED reads of its BAM as planned by make_examples_native.trim_reads.  Phasing (upstream) decides
  the haplotype channel and the row order and no candidates file ships, so rows are compared as multisets on channels 0,1,2,3,5."""
  d = np.load(os.path.join(GOLDEN, 'pacbio_golden_subset.npz'))
  arrays = {k[4:]: d[k] for k in d.files if k.startswith('arr_')}
  pb = packing.PackedBatch(int(d['n_images']), int(d['n_reads']), int(d['n_pairs']), int(d['ref_stride']), arrays)
  o = pi.default_options(pi.ReadRequirements(min_base_quality=10, min_mapping_quality=1))
  o.channels = pi.PILEUP_DEFAULT_CHANNELS + ['haplotype']
  o.width = 147
  return pb, d['golden_images'], d['channels'].tolist(), pi.to_params(o)


def _assert_row_multisets_equal(got, golden, ch):
  import collections
  for i in range(golden.shape[0]):
    np.testing.assert_array_equal(got[i, :5][..., ch], golden[i, :5][..., ch])          # reference band
    g = collections.Counter(golden[i, r][:, ch].tobytes() for r in range(5, 100) if golden[i, r].any())
    o = collections.Counter(got[i, r][:, ch].tobytes() for r in range(5, 100) if got[i, r].any())
    assert g == o, f'example {i}: {sum((g & o).values())} of {sum(g.values())} golden rows reproduced'


def test_oracle_reproduces_pacbio_golden_rows():
  pb, golden, ch, params = _pacbio_fixture()
  assert golden.shape == (12, 100, 147, 7) and ch == [0, 1, 2, 3, 5]
  _assert_row_multisets_equal(oracle_lib.encode_batch(params, pb), golden, ch)


@pytest.mark.gpu
def test_cuda_encoder_reproduces_pacbio_golden_rows():
  pb, golden, ch, params = _pacbio_fixture()
  got = pi.GpuEncoder(params, 0).encode_host(pb)
  _assert_row_multisets_equal(got, golden, ch)
  np.testing.assert_array_equal(got, oracle_lib.encode_batch(params, pb))


def test_pacbio_report_numbers():
  r = json.load(open(os.path.join(GOLDEN, 'pacbio_golden_report.json')))
  assert r['n_examples'] == 401 and r['ref_band_equal'] == 401 and r['same_row_count'] == 401 and r['all_rows_matched'] == 401
  assert r['golden_read_rows'] == r['golden_read_rows_matched'] == 13689


# ---- CallVariantsOutput written by the reference's call_variants (a17: _create_cvo_proto + set_model_id) -----------------------

def _proto_fields(buf, nested=()):
  out = []
  for fn, wt, val, _ in protos.iter_fields(buf):
    v = val if isinstance(val, int) else bytes(val)
    if fn in nested and not isinstance(v, int):
      v = tuple(sorted(map(repr, _proto_fields(v, (2,)))))   # sub-messages with map fields: entry order is not significant
    out.append((fn, wt, v))
  return sorted(map(repr, out))


def test_create_cvo_equals_reference_call_variants_output():
  """10 (example, CallVariantsOutput) pairs of the reference's goldens (all 84 checked by hand in the build container):
  given the example's variant / alt_allele_indices and the reference's probabilities, create_cvo() emits the same proto —
  same length, same fields, calls[0].info['MID'] = 'deepvariant' — up to the order of map entries (protobuf maps serialise in
  hash order; the reference compares parsed protos)."""
  from deepvariant_b200 import call_variants as cv
  pairs = json.load(open(os.path.join(GOLDEN, 'cvo_golden_pairs.json')))['pairs']
  assert len(pairs) == 10
  for p in pairs:
    golden = bytes.fromhex(p['cvo'])
    v, idx, probs = protos.parse_call_variants_output(golden)
    assert len(probs) == 3 and abs(sum(probs) - 1) < 1e-6
    mine = cv.create_cvo(bytes.fromhex(p['variant']), probs, bytes.fromhex(p['alt_allele_indices']))
    v2, idx2, probs2 = protos.parse_call_variants_output(mine)
    assert len(mine) == len(golden) and idx2 == idx and probs2 == probs
    assert _proto_fields(v2, (11,)) == _proto_fields(v, (11,))
    assert protos.encode_call_variants_output(v, idx, probs) == golden     # wire round trip of the golden record itself


def _allele_frequency_fixture():
  """tools/check_allele_frequency_golden.py: examples of the reference's golden.allele_frequency_examples.tfrecord.gz (100 x 221 x 8:
  the 7 WGS channels + allele_frequency), all 8 channels, with the batches the product flow packed for them from BAM + FASTA +
  population VCF (realigned reads, candidates, read support, the allele_frequency pair plane)."""
  d = np.load(os.path.join(GOLDEN, 'allele_frequency_golden_subset.npz'))
  arrays = {k[4:]: d[k] for k in d.files if k.startswith('arr_')}
  pb = packing.PackedBatch(int(d['n_images']), int(d['n_reads']), int(d['n_pairs']), int(d['ref_stride']), arrays)
  o = pi.default_options(pi.ReadRequirements(min_base_quality=10, min_mapping_quality=5))
  o.channels = list(pi.PILEUP_CHANNELS_WITH_INSERT_SIZE) + ['allele_frequency']
  return pb, d['golden_images'], o


def test_oracle_reproduces_allele_frequency_golden_images():
  pb, golden, o = _allele_frequency_fixture()
  assert golden.shape[1:] == (100, 221, 8) and golden.shape[0] >= 6
  assert sum(1 for g in golden if g[5:, :, 7].any()) >= 6               # images whose reads carry a population frequency
  assert 'pair_channel_0' in pb.arrays
  np.testing.assert_array_equal(oracle_lib.encode_batch(pi.to_params(o), pb), golden)


@pytest.mark.gpu
def test_cuda_encoder_reproduces_allele_frequency_golden_images():
  pb, golden, o = _allele_frequency_fixture()
  enc = pi.GpuEncoder(pi.to_params(o), 0)
  np.testing.assert_array_equal(enc.encode_host(pb), golden)


def _deeptrio_fixture():
  """tools/check_deeptrio_golden.py: examples of the reference's DeepTrio golden (deeptrio/testdata/golden_child.calling_examples,
  140 x 221 x 7 = parent1 (40 rows) | child (60) | parent2 (40), every block down-sampled) with the three per-sample batches the
  multi-sample planner packed for them (each sample's realigned reads against the same DeepVariantCall)."""
  d = np.load(os.path.join(GOLDEN, 'deeptrio_golden_subset.npz'))
  o = pi.default_options(pi.ReadRequirements(min_base_quality=10, min_mapping_quality=5))
  o.channels = list(pi.PILEUP_CHANNELS_WITH_INSERT_SIZE)
  blocks = []
  for b, h in enumerate((40, 60, 40)):
    n = d[f's{b}_n']
    arrays = {k[len(f's{b}_arr_'):]: d[k] for k in d.files if k.startswith(f's{b}_arr_')}
    blocks.append((packing.PackedBatch(int(n[0]), int(n[1]), int(n[2]), int(n[3]), arrays), pi.to_params(o, height=h)))
  return blocks, d['golden_images']


def test_oracle_reproduces_deeptrio_golden_images():
  blocks, golden = _deeptrio_fixture()
  assert golden.shape[1:] == (140, 221, 7)
  got = np.concatenate([oracle_lib.encode_batch(params, pb) for pb, params in blocks], axis=1)   # FillPileupArrayBySample: blocks stacked
  np.testing.assert_array_equal(got, golden)
  assert any((np.diff(pb.arrays['pair_begin'])[:pb.n_images] > params.height - 5).any() for pb, params in blocks)   # down-sampled blocks


@pytest.mark.gpu
def test_cuda_encoder_reproduces_deeptrio_golden_images():
  blocks, golden = _deeptrio_fixture()
  got = np.concatenate([pi.GpuEncoder(params, 0).encode_host(pb) for pb, params in blocks], axis=1)
  np.testing.assert_array_equal(got, golden)


def test_multi_sample_golden_report_says_every_example_is_reproduced():
  import json
  r = json.load(open(os.path.join(GOLDEN, 'deeptrio_golden_report.json')))
  assert r['golden_examples'] == r['images_identical'] == 88 and r['same_examples_in_same_order']
  assert r['blocks_identical_parent1_child_parent2'] == [88, 88, 88]
