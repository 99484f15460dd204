"""The plane-backed channels (include/dvb.h "channel planes"): allele_frequency, read_supports_variant_fuzzy,
allele_sample_probability, base_methylation, base_6ma, the Ultima homopolymer-quality channels, and mean_coverage.

  * known answers transcribed (as data) from the reference's own tests:
      deepvariant/channels/read_supports_variant_fuzzy_channel_test.cc:98-160   (6 tests, 11 expectations)
      deepvariant/pileup_channel_lib_test.cc:230-368                            (ReadSupportsAltFuzzy, 7 expectations)
      deepvariant/pileup_channel_lib_test.cc:635-684                            (AlleleSampleProbabilityChannelTest: 207, 146, 146)
  * the native value functions (csrc/dvb_channels.cu) against an independent float32 restatement written here;
  * pixel placement: the oracle and (`-m gpu`) the CUDA encoder against a third, minimal model of the CIGAR walk in this file,
    and against each other on seeded random batches with every plane channel at once.
"""
import dataclasses
import math

import numpy as np
import pytest

from deepvariant_b200 import channels, packing
from deepvariant_b200 import pileup_image as pi
from deepvariant_b200.protos import DeepVariantCall, Read, Variant, parse_cigar_string

F32 = np.float32


def _call(ref, alts, support, alt_ps=None, ref_support=(), start=10):
  return DeepVariantCall(variant=Variant(reference_name='chr1', start=start, end=start + len(ref), reference_bases=ref, alternate_bases=list(alts),
                                         alt_ps=list(alt_ps) if alt_ps else None),
                         allele_support={k: list(v) for k, v in support.items()}, ref_support=list(ref_support))


def _named(name, number=0, hp=None):
  return Read(fragment_name=name, read_number=number, hp_values=[hp] if hp is not None else None)


# ---- read_supports_variant_fuzzy_channel_test.cc ---------------------------------------------------------------------------------------

def test_fuzzy_exact_match():
  c = _call('A', ['AC', 'ACC'], {'AC': ['read1/0'], 'ACC': ['read2/0']}, alt_ps=[0, 1, 1])
  r1, r2 = _named('read1', 0, 1), _named('read2', 0, 1)
  assert channels.fuzzy_read_supports_alt(c, r1, ['AC']) == 1
  assert channels.fuzzy_read_supports_alt(c, r2, ['AC']) == 10
  assert channels.fuzzy_read_supports_alt(c, r1, ['ACC']) == 10
  assert channels.fuzzy_read_supports_alt(c, r2, ['ACC']) == 1


@pytest.mark.parametrize('alts,support,alt_ps,hp,want', [
    (['AC', 'ACC'], {'ACC': ['read1/0']}, [0, 1, 1], 1, 10),    # FuzzyMatch1bp
    (['AC', 'ACCC'], {'ACCC': ['read1/0']}, [0, 1, 1], 1, 9),   # FuzzyMatch2bp
    (['AC', 'ACC'], {'ACC': ['read1/0']}, [0, 1, 1], 2, 0),     # FuzzyMatchPhaseMismatch
    (['AC', 'ACC'], {'ACC': ['read1/0']}, [0, 0, 0], 1, 10),    # FuzzyMatchPhaseZero
])
def test_fuzzy_match_cases(alts, support, alt_ps, hp, want):
  c = _call('A', alts, support, alt_ps=alt_ps)
  assert channels.fuzzy_read_supports_alt(c, _named('read1', 0, hp), ['AC']) == want


def test_fuzzy_ref_support():
  c = _call('A', ['ATGC', 'AA'], {}, ref_support=['read1/0'])
  r = _named('read1', 0)
  assert channels.fuzzy_read_supports_alt(c, r, ['ATGC']) == 0
  assert channels.fuzzy_read_supports_alt(c, r, ['AA']) == 10


# ---- pileup_channel_lib_test.cc ReadSupportsAltFuzzy ------------------------------------------------------------------------------------

def test_fuzzy_lib_kats():
  assert channels.fuzzy_read_supports_alt(DeepVariantCall(), _named('', 0), []) == 0                                  # AlleleUnsupporting
  c = _call('', ['GGGCGCATT'], {'GGGCGCATT': ['FRAG1/1']})
  assert channels.fuzzy_read_supports_alt(c, _named('FRAG1', 1), ['GGGCGCATT']) == 1                                 # AlleleSupporting
  c = _call('', ['GGGCGCATT'], {'GGGCGCATT': ['FRAG2/2']})
  assert channels.fuzzy_read_supports_alt(c, _named('FRAG2', 2), []) == 0                                            # OtherAlleleSupporting
  # OtherAlleleFuzzySupporting: ALT_PS = [1, 1, 1, 2] -> phases of the four alts 1, 1, 2, 0 (values[0] is the reference's)
  c = _call('', ['GGGCGCATT', 'GGGCGCAT', 'GGGCGCATTT', 'GGGCGCA'],
            {'GGGCGCATT': ['Read1/1'], 'GGGCGCAT': ['Read2/1'], 'GGGCGCATTT': ['Read3/1'], 'GGGCGCA': ['Read4/1']}, alt_ps=[1, 1, 1, 2])
  got = [channels.fuzzy_read_supports_alt(c, _named(f'Read{i}', 1, hp), ['GGGCGCATT']) for i, hp in ((1, 1), (2, 1), (3, 1), (4, 2))]
  assert got == [1, 10, 10, 0]


def test_fuzzy_colors():
  o = pi.default_options()
  assert [channels.fuzzy_supports_alt_color(c, o) for c in (0, 1, 2, 10, 9, 8)] == [152, 254, 152, 228, 203, 177]
  with pytest.raises(ValueError):
    channels.fuzzy_supports_alt_color(3, o)


# ---- AlleleSampleProbabilityChannelTest -------------------------------------------------------------------------------------------------

@pytest.mark.parametrize('name,want', [('read1', 207), ('read3', 146), ('read4', 146)])
def test_allele_sample_probability_kats(name, want):
  c = _call('', [], {'A': ['read1/0', 'read2/0']}, ref_support=['read3/0'])
  assert channels.allele_sample_probability_color(c, f'{name}/0') == want


def test_allele_sample_probability_no_reads_is_zero():
  assert channels.allele_sample_probability_color(DeepVariantCall(), 'x/0') == 0   # max_val == 0


# ---- allele_frequency -----------------------------------------------------------------------------------------------------------------------

def _af_color_model(af, mn):
  """allele_frequency_channel.cc:76-87 in float32."""
  af, mn = F32(af), F32(mn)
  if af <= mn:
    return 0
  la, lm = F32(math.log10(float(af))), F32(math.log10(float(mn)))
  return int(F32(F32(lm - la) / lm) * F32(254)) & 0xFF


@pytest.mark.parametrize('af', [0.0, 1e-5, 1.1e-5, 1e-4, 0.001, 0.0123, 0.1, 0.25, 0.5, 0.9, 1.0])
def test_allele_frequency_color(af):
  assert channels.allele_frequency_color(af, 1e-5) == _af_color_model(af, 1e-5)


def test_allele_frequency_color_endpoints():
  assert channels.allele_frequency_color(1.0, 1e-5) == 254 and channels.allele_frequency_color(1e-5, 1e-5) == 0
  assert channels.allele_frequency_color(0.001, 1e-5) == 101   # (−5 − (−3)) / −5 · 254 = 101.6


def test_read_allele_frequency_follows_the_image_alts():
  c = _call('A', ['C', 'G'], {'C': ['r1/1'], 'G': ['r2/1', 'r1/1']})
  c.allele_frequency = {'C': 0.25, 'G': 0.5}
  assert channels.read_allele_frequency(c, 'r1/1', ['C']) == 0.25
  assert channels.read_allele_frequency(c, 'r1/1', ['G']) == 0.5     # C is skipped: not an alt of this image
  assert channels.read_allele_frequency(c, 'r2/1', ['C']) == 0.0
  assert channels.read_allele_frequency(c, 'r3/1', ['C', 'G']) == 0.0
  del c.allele_frequency['G']
  assert channels.read_allele_frequency(c, 'r2/1', ['G']) == 0.0     # supported alt without a frequency entry


# ---- per-base value functions against an independent float32 restatement ---------------------------------------------------------------

def _bq_color(q):
  return int(F32(F32(254) * F32(q)) / F32(93)) & 0xFF


def _hmer_model(seq, qual, tp, is_deletion):
  out, i, n = [0] * len(seq), 0, len(seq)
  while i < n:
    j = i
    while j < n and seq[j] == seq[i]:
      j += 1
    p = F32(0)
    for k in range(i, j):
      if tp is not None and tp[k] != 0 and (tp[k] < 0) == is_deletion:
        p = F32(p + F32(math.pow(10, qual[k] / -10.0)))
    q = 93 if p == 0 else min(int(F32(-10) * F32(math.log10(float(p)))), 93)
    for k in range(i, j):
      out[k] = _bq_color(q)
    i = j
  return out


def _ultima_read(seq, qual, tp=None, t0=None, mods=None):
  return Read(fragment_name='u', aligned_sequence=seq.encode(), aligned_quality=bytes(qual), tp_values=tp, t0_value=t0, base_modifications=mods,
              cigar=parse_cigar_string(f'{len(seq)}M'), mapping_quality=60)


def test_hmer_quality_known_values():
  # TAAAAAG: the A run carries one +1 error of Q30 and one -1 error of Q20
  seq, qual = 'TAAAAAG', [40, 30, 35, 35, 20, 35, 40]
  tp = [0, 1, 0, 0, -1, 0, 0]
  r = _ultima_read(seq, qual, tp)
  ins, dele = channels.hmer_quality_plane(r, False), channels.hmer_quality_plane(r, True)
  assert list(ins) == [254] + [_bq_color(30)] * 5 + [254] == [254, 81, 81, 81, 81, 81, 254]
  assert list(dele) == [254] + [_bq_color(20)] * 5 + [254] == [254, 54, 54, 54, 54, 54, 254]
  # no tp tag: every homopolymer keeps kMaxQScore
  assert list(channels.hmer_quality_plane(_ultima_read(seq, qual), False)) == [254] * 7


def test_hmer_quality_matches_float32_model_on_random_reads():
  rng = np.random.RandomState(5)
  for _ in range(200):
    n = int(rng.randint(1, 120))
    seq = ''.join(rng.choice(list('ACGT'), p=[0.4, 0.2, 0.2, 0.2]) for _ in range(n))
    qual = [int(q) for q in rng.randint(0, 60, n)]
    tp = [int(t) for t in rng.choice([0, 0, 0, 1, -1, 2, -2], n)]
    r = _ultima_read(seq, qual, tp)
    for is_del in (False, True):
      assert list(channels.hmer_quality_plane(r, is_del)) == _hmer_model(seq, qual, tp, is_del), (seq, qual, tp, is_del)
  # a tag shorter than the read is zero beyond its end (GetTPValues)
  r = _ultima_read('AAAA', [10, 10, 10, 10], [1])
  assert list(channels.hmer_quality_plane(r, False)) == _hmer_model('AAAA', [10] * 4, [1, 0, 0, 0], False)


def test_t0_and_base_modification_planes():
  r = _ultima_read('AAAATTTT', [30] * 8, t0=b'5555IIII')
  assert list(channels.t0_plane(r)) == [_bq_color(20)] * 4 + [_bq_color(40)] * 4
  assert list(channels.t0_plane(_ultima_read('ACG', [30] * 3, t0=b'5'))) == [_bq_color(20), 0, 0]     # shorter tag
  assert list(channels.t0_plane(_ultima_read('ACG', [30] * 3))) == [0, 0, 0]                          # no tag
  vals = bytes([0, 1, 127, 128, 254, 255])
  r = _ultima_read('ACGTAC', [30] * 6, mods={'5mC': vals})
  want = [int(F32(254) * (F32(v) / F32(255))) for v in vals]
  assert list(channels.base_modification_plane(r, '5mC')) == want == [0, 0, 126, 127, 253, 254]
  assert list(channels.base_modification_plane(r, '6mA')) == [0] * 6                                  # that modification is absent
  with pytest.raises(ValueError):
    channels.base_modification_plane(_ultima_read('ACG', [30] * 3, mods={'5mC': b'\x01'}), '5mC')


# ---- pixel placement: oracle and CUDA encoder against a minimal model of the walk -------------------------------------------------------

ALL_PLANE_CHANNELS = ['read_base', 'allele_frequency', 'read_supports_variant_fuzzy', 'allele_sample_probability', 'base_methylation', 'base_6ma',
                      'homopolymer_insertion_quality', 'homopolymer_deletion_quality', 'inter_homopolymer_insertion_quality', 'mean_coverage']


def _opts(chs, **kw):
  o = pi.default_options(pi.ReadRequirements(0, 0))
  return dataclasses.replace(o, channels=list(chs), num_channels=len(chs), **kw)


def _place(read, plane, image_start, width, anchor=True):
  """One channel's row: CalculateBaseLevelData (pileup_channel_lib.cc:171-261) with data[col] = plane[read_index]."""
  row = [0] * width
  ref_i, read_i = read.position, 0

  def put(r, i):
    col = r - image_start
    if 0 <= col < width:
      row[col] = int(plane[i])
  for op, ln in read.cigar:
    if op in (0, 7, 8):
      for _ in range(ln):
        put(ref_i, read_i)
        ref_i += 1
        read_i += 1
    elif op in (1, 4):
      if op == 1 and ref_i > 0 and anchor:
        put(ref_i - 1, read_i)
      read_i += ln
    elif op in (2, 3):
      if op == 2 and read_i > 0 and anchor:
        put(ref_i - 1, read_i - 1)
      ref_i += ln
  return row


def test_plane_channels_are_placed_like_fill_read_base(backend):
  width, start = 21, 100
  ref = 'ACGTACGTACGTACGTACGTA'
  rng = np.random.RandomState(11)
  reads = []
  for k, (pos, cigar) in enumerate([(95, '4S10M2I6M'), (104, '5M3D7M1I3M'), (90, '30M'), (110, '2I8M')]):
    n = sum(ln for op, ln in parse_cigar_string(cigar) if op in (0, 1, 4, 7, 8))
    seq = ''.join(rng.choice(list('ACGT')) for _ in range(n))
    reads.append(Read(fragment_name=f'r{k}', read_number=1, position=pos, cigar=parse_cigar_string(cigar), aligned_sequence=seq.encode(),
                      aligned_quality=bytes(int(q) for q in rng.randint(20, 40, n)), mapping_quality=60, hp_values=[2 if k == 2 else 1],
                      tp_values=[int(t) for t in rng.choice([0, 1, -1], n)], t0_value=bytes(int(c) for c in rng.randint(33, 33 + 60, n)),
                      base_modifications={'5mC': bytes(int(v) for v in rng.randint(0, 256, n)), '6mA': bytes(int(v) for v in rng.randint(0, 256, n))}))
  c = _call('C', ['CA', 'CAA', 'G'], {'CA': ['r0/1'], 'CAA': ['r1/1'], 'G': ['r2/1']}, alt_ps=[0, 1, 1, 0], ref_support=['r3/1'], start=110)
  c.allele_frequency = {'CA': 0.02, 'CAA': 0.3, 'G': 0.9}
  o = _opts(ALL_PLANE_CHANNELS, width=width, height=12, reference_band_height=2, mean_coverage=6.9)
  img = backend(o).build_pileup_for_one_sample(c, ref, reads, start, ['CA'])
  assert img.shape == (12, width, len(ALL_PLANE_CHANNELS))
  # rows: 2 reference rows, then the four reads sorted by position (90, 95, 104, 110)
  order = [2, 0, 1, 3]
  pair = channels.pair_planes(c, reads, ['CA'], o)
  assert pair[0] == [channels.allele_frequency_color(0.02, 1e-5), 0, 0, 0]
  # exact; CAA one base from CA on a compatible phase; G one base from CA but read 2 sits on the other haplotype (the first loop never reports
  # class 2, so it ends as 0); the reference allele C is one base from CA too (reference-support loop)
  assert pair[1] == [254, 228, 152, 228]
  for row, k in enumerate(order):
    r = reads[k]
    drawn = np.array(_place(r, [1] * len(r.aligned_sequence), start, width)) > 0
    got = img[2 + row]
    for ch, slot in ((1, 0), (2, 1), (3, 2)):
      assert list(got[:, ch]) == [pair[slot][k] if d else 0 for d in drawn], (row, ch)
    for ch, slot in ((4, 0), (5, 1), (6, 2), (7, 3), (8, 4)):
      assert list(got[:, ch]) == _place(r, channels.base_plane(r, slot), start, width), (row, ch)
  # reference band of the plane channels: 0 except the fuzzy channel's SupportsAltColor(0); mean_coverage paints 255 / 200
  assert (img[:2, :, [1, 3, 4, 5, 6, 7, 8]] == 0).all() and (img[:2, :, 2] == 152).all()
  cov = img[:, :, 9]
  assert (cov[:2] == 255).all() and (cov[2:8] == 200).all() and (cov[8:] == 0).all()   # int(6.9) = 6 rows below the band, blank rows included
  assert (img[6:, :, :9] == 0).all()


@pytest.mark.parametrize('mean_coverage,painted', [(0.0, 3), (1.0, 4), (250.0, 8), (-9.0, 0)])
def test_mean_coverage_overlay(backend, mean_coverage, painted):
  """pileup_image_native.cc:422-444: rows [0, min(int(mean_coverage) + band, height))."""
  o = _opts(['read_base', 'mean_coverage', 'base_quality'], width=11, height=8, reference_band_height=3, mean_coverage=mean_coverage)
  read = Read(fragment_name='a', read_number=1, position=3, cigar=parse_cigar_string('6M'), aligned_sequence=b'ACGTAC', aligned_quality=bytes([30] * 6),
              mapping_quality=60)
  c = _call('A', ['C'], {'C': []}, start=5)
  img = backend(o).build_pileup_for_one_sample(c, 'ACGTACGTACG', [read], 0, ['C'])
  cov = img[:, :, 1]
  want = np.zeros((8, 11), dtype=np.uint8)
  want[:min(painted, 3)] = 255
  want[3:painted] = 200
  assert np.array_equal(cov, want)
  assert (img[3, 3:9, 0] > 0).all() and (img[3, 3:9, 2] > 0).all()   # the read row's own channels are untouched


def test_missing_plane_is_an_error():
  import oracle_lib
  o = _opts(['read_base', 'base_methylation'], width=11, height=4, reference_band_height=1)
  params = pi.to_params(o)
  read = Read(fragment_name='a', read_number=1, position=3, cigar=parse_cigar_string('6M'), aligned_sequence=b'ACGTAC', aligned_quality=bytes([30] * 6))
  spec = packing.ImageSpec(ref_bases='ACGTACGTACG', image_start_pos=0, variant_start=5, reads=[read], support=[0], allele_group=[0])
  batch = packing.pack_images([spec], params)
  del batch.arrays['base_channel_0']
  with pytest.raises(oracle_lib.OracleError):
    oracle_lib.encode_batch(params, batch)


@pytest.mark.gpu
def test_missing_plane_is_an_error_on_the_device():
  from deepvariant_b200 import _lib
  o = _opts(['read_base', 'allele_frequency'], width=11, height=4, reference_band_height=1)
  params = pi.to_params(o)
  read = Read(fragment_name='a', read_number=1, position=3, cigar=parse_cigar_string('6M'), aligned_sequence=b'ACGTAC', aligned_quality=bytes([30] * 6))
  spec = packing.ImageSpec(ref_bases='ACGTACGTACG', image_start_pos=0, variant_start=5, reads=[read], support=[0], allele_group=[0], pair_channels={0: [7]})
  batch = packing.pack_images([spec], params)
  enc = pi.GpuEncoder(params, device=0)
  assert enc.encode_host(batch)[0, 1, 3, 1] == 7
  del batch.arrays['pair_channel_0']
  with pytest.raises(_lib.DvbError):
    enc.encode_host(batch)


@pytest.mark.gpu
@pytest.mark.parametrize('chs,width', [(ALL_PLANE_CHANNELS, 221), (pi.PILEUP_DEFAULT_CHANNELS + ['base_methylation'], 147),
                                       (pi.PILEUP_CHANNELS_WITH_INSERT_SIZE[:6] + ['allele_frequency'], 221),
                                       (pi.PILEUP_DEFAULT_CHANNELS + ['is_homopolymer', 'base_6ma', 'mean_coverage', 'allele_sample_probability'], 99)])
def test_random_planes_match_oracle_on_the_device(chs, width):
  """Seeded random batches (down-sampled images included) with random planes: CUDA == oracle, through the device entry point and the
  host entry point (staging of the plane segments, chunked or not)."""
  import torch
  import oracle_lib
  from deepvariant_b200 import synthetic
  tb = synthetic.make_batch(150, 'cuda:0', width=width, chunk=21, deep_fraction=0.05)
  g = torch.Generator(device='cuda:0')
  g.manual_seed(3)
  for k in range(3):
    tb.tensors[f'pair_channel_{k}'] = torch.randint(0, 256, (max(tb.n_pairs, 1),), generator=g, device='cuda:0').to(torch.uint8)
  for k in range(5):
    tb.tensors[f'base_channel_{k}'] = torch.randint(0, 256, (max(tb.n_bases, 1),), generator=g, device='cuda:0').to(torch.uint8)
  o = _opts(chs, width=width, mean_coverage=31.5, read_requirements=pi.ReadRequirements(10, 10))
  params = pi.to_params(o)
  enc = pi.GpuEncoder(params, device=0)
  out = torch.full((tb.n_images,) + enc.shape, 0xAB, dtype=torch.uint8, device='cuda:0')
  enc.encode_device(tb, out)
  enc.check()
  packed = tb.to_packed()
  want = oracle_lib.encode_batch(params, packed)
  assert np.array_equal(out.cpu().numpy(), want)
  assert np.array_equal(enc.encode_host(packed), want)
