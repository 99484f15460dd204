"""Parity tests proper: the CUDA encoder (through the C ABI) vs the CPU oracle on the same seeded
inputs — bit-exact — plus size-independent properties at full batch sizes.  All `-m gpu`."""
import dataclasses

import numpy as np
import pytest
import torch

import oracle_lib
from deepvariant_b200 import pileup_image as pi
from deepvariant_b200 import synthetic

pytestmark = pytest.mark.gpu

DEV = 'cuda:0'


def _options(channels, **kw):
  o = pi.default_options()
  o.channels = list(channels)
  o.num_channels = len(channels)
  return dataclasses.replace(o, **kw)


WGS = pi.PILEUP_CHANNELS_WITH_INSERT_SIZE
PACBIO = pi.PILEUP_DEFAULT_CHANNELS + ['haplotype', 'supplementary_alignment']


def _encode_both(options, tb, height=None):
  params = pi.to_params(options, height=height)
  enc = pi.GpuEncoder(params, device=0)
  out = torch.empty((tb.n_images,) + enc.shape, dtype=torch.uint8, device=DEV)
  out.fill_(0xAB)  # poison: every byte must be written
  rows = torch.full((tb.n_images,), -1, dtype=torch.int32, device=DEV)
  enc.encode_device(tb, out, rows)
  enc.check()
  want = oracle_lib.encode_batch(params, tb.to_packed())
  return out.cpu().numpy(), want, rows.cpu().numpy(), enc


def _assert_same(got, want):
  if not np.array_equal(got, want):
    bad = np.argwhere(got != want)
    i = bad[0]
    raise AssertionError(f'{len(bad)} differing bytes; first at image {i[0]} row {i[1]} col {i[2]} ch {i[3]}: '
                         f'gpu {got[tuple(i)]} oracle {want[tuple(i)]}; images affected '
                         f'{np.unique(bad[:, 0])[:10]}, rows {np.unique(bad[:, 1])[:10]}')


@pytest.mark.parametrize('seed_chunk', [0, 1])
def test_wgs_synthetic_matches_oracle(seed_chunk):
  tb = synthetic.make_batch(300, DEV, chunk=seed_chunk, deep_fraction=0.05)
  got, want, rows, _ = _encode_both(_options(WGS), tb)
  _assert_same(got, want)
  assert rows.min() >= 0 and rows.max() == 95  # deep images hit the 95-row cap


def test_pacbio_channels_hp_sort_matches_oracle():
  tb = synthetic.make_batch(200, DEV, width=147, chunk=3, hp=True, deep_fraction=0.05)
  o = _options(PACBIO, width=147, sort_by_haplotypes=True,
               read_requirements=pi.ReadRequirements(min_base_quality=10, min_mapping_quality=1))
  got, want, _, _ = _encode_both(o, tb)
  _assert_same(got, want)


def test_alt_channel_padding_matches_oracle():
  """PACBIO production layout: 8 computed + 2 alt-aligned (zero) channels = 10."""
  tb = synthetic.make_batch(64, DEV, width=147, chunk=4, hp=True)
  o = _options(PACBIO + ['diff_channels_alternate_allele_1', 'diff_channels_alternate_allele_2'], width=147,
               sort_by_haplotypes=True, alt_aligned_pileup='diff_channels')
  got, want, _, enc = _encode_both(o, tb)
  assert enc.shape == (100, 147, 10)
  _assert_same(got, want)
  assert not got[..., 8:].any()


def test_polish_tag_and_group_sort_matches_oracle():
  tb = synthetic.make_batch(100, DEV, chunk=5, hp=True)
  tb.tensors['pair_allele_group'] = torch.randint(0, 3, (tb.n_pairs,), device=DEV).to(torch.uint8)
  o = _options(WGS + ['haplotype'], sort_by_haplotypes=True, hp_tag_for_assembly_polishing=2,
               sort_by_alt_allele_support=True)
  got, want, _, _ = _encode_both(o, tb)
  _assert_same(got, want)


@pytest.mark.parametrize('width,height,band', [(11, 4, 1), (9, 8, 1), (33, 20, 3), (299, 140, 5)])
def test_odd_shapes_match_oracle(width, height, band):
  tb = synthetic.make_batch(50, DEV, width=width, chunk=width, mean_depth=6.0 if height < 30 else 40.0)
  o = _options(pi.PILEUP_DEFAULT_CHANNELS, width=width, height=height, reference_band_height=band)
  got, want, _, _ = _encode_both(o, tb)
  _assert_same(got, want)


def test_single_channel_and_sixteen_channels():
  tb = synthetic.make_batch(40, DEV, chunk=9, hp=True)
  got, want, _, _ = _encode_both(_options(['base_differs_from_ref']), tb)
  _assert_same(got, want)
  many = (WGS + ['haplotype', 'supplementary_alignment', 'blank'] + pi.PILEUP_DEFAULT_CHANNELS)[:16]
  got, want, _, _ = _encode_both(_options(many), tb)
  _assert_same(got, want)


def test_low_quality_thresholds_reject_rows():
  tb = synthetic.make_batch(120, DEV, chunk=11)
  o = _options(WGS, read_requirements=pi.ReadRequirements(min_base_quality=26, min_mapping_quality=30))
  got, want, rows, _ = _encode_both(o, tb)
  _assert_same(got, want)
  depth = (tb.tensors['pair_begin'][1:] - tb.tensors['pair_begin'][:-1]).cpu().numpy()
  assert (rows < np.minimum(depth, 95)).any()  # some reads were rejected


def test_empty_batch_and_images_without_reads():
  o = _options(WGS)
  params = pi.to_params(o)
  enc = pi.GpuEncoder(params, device=0)
  tb = synthetic.make_batch(8, DEV, mean_depth=0.0, deep_fraction=0.0)
  assert tb.n_pairs == 0
  out = torch.full((8,) + enc.shape, 7, dtype=torch.uint8, device=DEV)
  enc.encode_device(tb, out)
  enc.check()
  want = oracle_lib.encode_batch(params, tb.to_packed())
  _assert_same(out.cpu().numpy(), want)
  assert not out[:, 5:].any() and out[:, :5].any()
  # n_images == 0 is a no-op
  tb0 = synthetic.TorchBatch(tb.tensors, 0, 0, 0, tb.ref_stride)
  enc.encode_device(tb0, out)
  enc.check()


def test_host_entry_point_matches_device_entry_point():
  tb = synthetic.make_batch(97, DEV, chunk=13, deep_fraction=0.03)
  o = _options(WGS)
  params = pi.to_params(o)
  enc = pi.GpuEncoder(params, device=0)
  out = torch.empty((tb.n_images,) + enc.shape, dtype=torch.uint8, device=DEV)
  enc.encode_device(tb, out)
  enc.check()
  host = enc.encode_host(tb.to_packed())
  np.testing.assert_array_equal(host, out.cpu().numpy())
  assert enc.last_rows_kept[:97].max() == 95


def test_bad_cigar_is_reported_not_fatal():
  from deepvariant_b200 import _lib
  tb = synthetic.make_batch(16, DEV, chunk=17)
  tb.tensors['cigar'][3] = (5 << 4) | 9  # op 9 does not exist
  o = _options(WGS)
  enc = pi.GpuEncoder(pi.to_params(o), device=0)
  out = torch.empty((16,) + enc.shape, dtype=torch.uint8, device=DEV)
  enc.encode_device(tb, out)
  with pytest.raises(_lib.DvbError) as e:
    enc.check()
  assert e.value.status == 3
  enc.check()  # error word is cleared
  with pytest.raises(_lib.DvbError):
    enc.encode_host(tb.to_packed())  # host path validates before launching


def test_unaligned_output_pointer():
  """Row staging keys on the 16-byte phase of the destination: offset the output by 1..15 bytes."""
  tb = synthetic.make_batch(20, DEV, chunk=19)
  o = _options(WGS)
  params = pi.to_params(o)
  enc = pi.GpuEncoder(params, device=0)
  want = oracle_lib.encode_batch(params, tb.to_packed())
  nbytes = want.size
  for off in (1, 4, 7, 15):
    raw = torch.full((nbytes + 32,), 0xEE, dtype=torch.uint8, device=DEV)
    view = raw[off:off + nbytes].view((20,) + enc.shape)
    enc.encode_device(tb, view)
    enc.check()
    _assert_same(view.cpu().numpy(), want)
    assert int(raw[off - 1]) == 0xEE and int(raw[off + nbytes]) == 0xEE  # no overrun


def test_full_size_properties():
  """At bench batch size the oracle is too slow; check size-independent properties instead:
  determinism, 5 identical reference rows, blank tail, rows_kept == number of non-blank read rows,
  read rows sorted by position, and a sampled subset bit-exact against the oracle."""
  n = 8192
  tb = synthetic.make_batch(n, DEV, chunk=23)
  o = _options(WGS)
  params = pi.to_params(o)
  enc = pi.GpuEncoder(params, device=0)
  out = torch.empty((n,) + enc.shape, dtype=torch.uint8, device=DEV)
  rows = torch.zeros(n, dtype=torch.int32, device=DEV)
  enc.encode_device(tb, out, rows)
  enc.check()
  out2 = torch.empty_like(out)
  enc.encode_device(tb, out2)
  enc.check()
  assert torch.equal(out, out2)
  assert torch.equal(out[:, 0:1].expand(-1, 5, -1, -1), out[:, :5])
  nonblank = out[:, 5:].reshape(n, 95, -1).any(2)
  # a kept read may lie fully outside the window only if it does not overlap it: never here
  assert torch.equal(nonblank.sum(1).to(torch.int32), rows)
  assert torch.equal(nonblank, torch.arange(95, device=DEV).view(1, 95) < rows.view(n, 1))
  sample = torch.arange(0, n, 97, device=DEV)
  # exact check on a strided sample, re-packed as its own small batch
  idx = sample.cpu().numpy()
  packed = tb.to_packed()
  from subbatch_util import take_images
  small = take_images(packed, idx)
  want = oracle_lib.encode_batch(params, small)
  _assert_same(out[sample].cpu().numpy(), want)


def test_pair_prepass_and_single_kernel_paths_are_identical(monkeypatch):
  """The encoder runs as pre-pass (one record per (image, read) pair) + image kernel; DVB_ENC_PREPASS=0 keeps everything in
  the image kernel.  Both must give the oracle's bytes — WGS and PACBIO layouts, down-sampled images included."""
  import torch
  from deepvariant_b200 import synthetic
  for pacbio in (False, True):
    o = pi.default_options()
    if pacbio:
      o.channels = pi.PILEUP_DEFAULT_CHANNELS + ['haplotype', 'supplementary_alignment', 'diff_channels_alternate_allele_1',
                                                 'diff_channels_alternate_allele_2']
      o.width = 147
      o.sort_by_haplotypes = True
    else:
      o.channels = list(pi.PILEUP_CHANNELS_WITH_INSERT_SIZE)
    params = pi.to_params(o)
    tb = synthetic.make_batch(300, 'cuda:0', width=o.width, hp=pacbio)
    want = oracle_lib.encode_batch(params, tb.to_packed())
    for flag in ('1', '0'):
      monkeypatch.setenv('DVB_ENC_PREPASS', flag)
      enc = pi.GpuEncoder(params, 0)
      out = torch.empty((tb.n_images,) + enc.shape, dtype=torch.uint8, device='cuda:0')
      rows = torch.zeros(tb.n_images, dtype=torch.int32, device='cuda:0')
      enc.encode_device(tb, out, rows)
      enc.check()
      np.testing.assert_array_equal(out.cpu().numpy(), want, err_msg=f'pacbio={pacbio} prepass={flag}')
      assert enc.launch_count == (2 if flag == '1' else 1)
      enc.close()


def test_opt_channels_random_batch_matches_oracle():
  """12-channel layout: the six base channels + the five whole-read "Opt Channels" + insert_size, with indels, soft clips,
  down-sampled images and multi-allelic support classes from the synthetic generator."""
  chans = pi.PILEUP_DEFAULT_CHANNELS + ['read_mapping_percent', 'avg_base_quality', 'identity', 'gap_compressed_identity', 'gc_content',
                                        'insert_size']
  tb = synthetic.make_batch(200, DEV)
  got, want, rows, _ = _encode_both(_options(chans), tb)
  assert got.shape[-1] == 12
  np.testing.assert_array_equal(got, want)
  assert (got[:, :5, :, 6:10] == 254).all()          # reference band of the four fixed-value statistics
  assert len(np.unique(got[:, 0, 0, 10])) > 3         # GC content of the window differs from image to image


def test_homopolymer_channels_random_batch_matches_oracle():
  """The per-base homopolymer channels (separate kernel instantiation) next to the base channels and the whole-read statistics,
  on the synthetic batch (indels -> anchor pixels, soft clips, down-sampling) and on a PACBIO-width layout with haplotypes."""
  chans = pi.PILEUP_DEFAULT_CHANNELS + ['is_homopolymer', 'homopolymer_weighted', 'gc_content', 'insert_size']
  tb = synthetic.make_batch(200, DEV)
  got, want, _, _ = _encode_both(_options(chans), tb)
  np.testing.assert_array_equal(got, want)
  assert set(np.unique(got[..., 6])) == {0, 254} and len(np.unique(got[..., 7])) > 4
  chans = pi.PILEUP_DEFAULT_CHANNELS + ['haplotype', 'homopolymer_weighted']
  tb = synthetic.make_batch(120, DEV, width=147, hp=True)
  got, want, _, _ = _encode_both(_options(chans, width=147, sort_by_haplotypes=True), tb)
  np.testing.assert_array_equal(got, want)
