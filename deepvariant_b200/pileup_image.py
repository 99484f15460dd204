"""Host-side mirror of the reference's pileup-image interface.

  default_options()            deepvariant/pileup_image.py:36-74
  PileupImageOptions           deepvariant/protos/deepvariant.proto (PileupImageOptions)
  PileupImageEncoderNative     deepvariant/python/pileup_image_native_pybind.cc:79-112
      .encode_reference(ref_bases)                       -> uint8[1, W, C]
      .encode_read(dv_call, ref_bases, read, start, alts) -> uint8[1, W, C] | None
      .build_pileup_for_one_sample(dv_call, ref_bases, reads, start, alts) -> uint8[H, W, C]
      .all_channels_enum(alt_aligned_pileup)             -> list[int]

Same names, argument meaning and None-on-rejected-read behaviour as the pybind class;
the work itself is ONE batched CUDA launch through libdvb.so (no CPU path).
"""
from __future__ import annotations

import dataclasses
from typing import Dict, List, Optional, Sequence

import numpy as np

from deepvariant_b200 import _lib
from deepvariant_b200 import packing
from deepvariant_b200.protos import DeepVariantCall, Read

# deepvariant/dv_constants.py:39-73
PILEUP_DEFAULT_WIDTH = 221
PILEUP_DEFAULT_HEIGHT = 100
PILEUP_DEFAULT_CHANNELS = [
    'read_base', 'base_quality', 'mapping_quality', 'strand', 'read_supports_variant',
    'base_differs_from_ref',
]
PILEUP_CHANNELS_WITH_INSERT_SIZE = PILEUP_DEFAULT_CHANNELS + ['insert_size']

# Channel-name <-> DeepVariantChannelEnum (pileup_channel_lib.cc:449-520,
# deepvariant.proto:1288-1343).  Names that map to CH_UNSPECIFIED in
# Channels::ChannelStrToEnum (the alt-aligned pseudo channels) carry their enum for
# example_info.json but are not computed per read.
CHANNEL_ENUM: Dict[str, int] = {
    'read_base': 1, 'base_quality': 2, 'mapping_quality': 3, 'strand': 4,
    'read_supports_variant': 5, 'base_differs_from_ref': 6, 'haplotype': 7,
    'allele_frequency': 8, 'diff_channels_alternate_allele_1': 9,
    'diff_channels_alternate_allele_2': 10, 'read_mapping_percent': 11,
    'avg_base_quality': 12, 'identity': 13, 'gap_compressed_identity': 14, 'gc_content': 15,
    'is_homopolymer': 16, 'homopolymer_weighted': 17, 'blank': 18, 'insert_size': 19,
    'base_channels_alternate_allele_1': 20, 'base_channels_alternate_allele_2': 21,
    'mean_coverage': 22, 'base_methylation': 23, 'base_6ma': 24,
    'read_supports_variant_fuzzy': 25, 'supplementary_alignment': 26, 'allele_sample_probability': 27,
    'homopolymer_insertion_quality': 28, 'homopolymer_deletion_quality': 29, 'inter_homopolymer_insertion_quality': 30,
}
ALT_ALIGNED_PSEUDO_CHANNELS = (
    'diff_channels_alternate_allele_1', 'diff_channels_alternate_allele_2',
    'base_channels_alternate_allele_1', 'base_channels_alternate_allele_2')
SUPPORTED_ENUMS = (1, 2, 3, 4, 5, 6, 7, 8, 11, 12, 13, 14, 15, 16, 17, 18, 19, 22, 23, 24, 25, 26, 27, 28, 29, 30)


@dataclasses.dataclass
class ReadRequirements:
  """third_party/nucleus/protos/reads.proto:421-466 (fields the path reads)."""
  min_base_quality: int = 0
  min_mapping_quality: int = 0


@dataclasses.dataclass
class PileupImageOptions:
  reference_band_height: int = 0
  base_color_offset_a_and_g: int = 0
  base_color_offset_t_and_c: int = 0
  base_color_stride: int = 0
  allele_supporting_read_alpha: float = 0.0
  allele_unsupporting_read_alpha: float = 0.0
  other_allele_supporting_read_alpha: float = 0.0
  reference_matching_read_alpha: float = 0.0
  reference_mismatching_read_alpha: float = 0.0
  indel_anchoring_base_char: str = ''
  reference_alpha: float = 0.0
  reference_base_quality: int = 0
  positive_strand_color: int = 0
  negative_strand_color: int = 0
  base_quality_cap: int = 0
  mapping_quality_cap: int = 0
  height: int = 0
  width: int = 0
  num_channels: int = 0
  channels: List[str] = dataclasses.field(default_factory=list)
  read_overlap_buffer_bp: int = 0
  read_requirements: ReadRequirements = dataclasses.field(default_factory=ReadRequirements)
  multi_allelic_mode: str = 'UNSPECIFIED'
  random_seed: int = 0
  sequencing_type: int = 0
  alt_aligned_pileup: str = ''
  types_to_alt_align: str = ''
  sort_by_haplotypes: bool = False
  hp_tag_for_assembly_polishing: int = 0
  sort_by_alt_allele_support: bool = False
  min_non_zero_allele_frequency: float = 0.0   # PileupImageOptions field 33 (allele_frequency channel)
  mean_coverage: float = 0.0                   # SampleOptions.mean_coverage of the one sample (mean_coverage channel)
  channels_enum_to_blank: Sequence[int] = ()   # SampleOptions.channels_enum_to_blank of the sample: channel enums whose read pixels stay 0


def default_options(read_requirements: Optional[ReadRequirements] = None) -> PileupImageOptions:
  """deepvariant/pileup_image.py:36-74."""
  if not read_requirements:
    read_requirements = ReadRequirements(min_base_quality=10, min_mapping_quality=10)
  return PileupImageOptions(
      reference_band_height=5,
      base_color_offset_a_and_g=40,
      base_color_offset_t_and_c=30,
      base_color_stride=70,
      allele_supporting_read_alpha=1.0,
      allele_unsupporting_read_alpha=0.6,
      other_allele_supporting_read_alpha=0.6,
      reference_matching_read_alpha=0.2,
      reference_mismatching_read_alpha=1.0,
      indel_anchoring_base_char='*',
      reference_alpha=0.4,
      reference_base_quality=60,
      positive_strand_color=70,
      negative_strand_color=240,
      base_quality_cap=40,
      mapping_quality_cap=60,
      height=PILEUP_DEFAULT_HEIGHT,
      width=PILEUP_DEFAULT_WIDTH,
      read_overlap_buffer_bp=5,
      read_requirements=read_requirements,
      multi_allelic_mode='ADD_HET_ALT_IMAGES',
      random_seed=2101079370,
      sequencing_type=0,
      alt_aligned_pileup='none',
      types_to_alt_align='indels',
      min_non_zero_allele_frequency=0.00001,
  )


def all_channels_enum(options: PileupImageOptions, alt_aligned_representation: str = '') -> List[int]:
  """PileupImageEncoderNative::AllChannelsEnum (pileup_image_native.cc:125-151)."""
  out = []
  for name in options.channels:
    if name in ALT_ALIGNED_PSEUDO_CHANNELS:
      continue  # ChannelStrToEnum -> CH_UNSPECIFIED, dropped
    if name not in CHANNEL_ENUM:
      raise ValueError(f'Channel "{name}" should have a corresponding enum')  # pileup_channel_lib.cc:518
    out.append(CHANNEL_ENUM[name])
  if alt_aligned_representation == 'diff_channels':
    out += [9, 10]
  elif alt_aligned_representation == 'base_channels':
    out += [20, 21]
  return out


def to_params(options: PileupImageOptions, height: Optional[int] = None) -> _lib.DvbPileupParams:
  """PileupImageOptions -> the C-ABI parameter block (include/dvb.h DvbPileupParams)."""
  enums = all_channels_enum(options, '')
  if len(enums) > _lib.DVB_MAX_CHANNELS:
    raise ValueError('too many channels')
  p = _lib.DvbPileupParams()
  p.width = options.width
  p.height = height if height is not None else options.height
  p.reference_band_height = options.reference_band_height
  p.num_channels = len(enums)
  for i, e in enumerate(enums):
    p.channels[i] = e
  # options.channels lists the alt-aligned pseudo channels last (make_examples_options.py:1095-1103)
  p.num_alt_channels = sum(1 for c in options.channels if c in ALT_ALIGNED_PSEUDO_CHANNELS)
  p.base_color_offset_a_and_g = options.base_color_offset_a_and_g
  p.base_color_offset_t_and_c = options.base_color_offset_t_and_c
  p.base_color_stride = options.base_color_stride
  p.allele_supporting_read_alpha = options.allele_supporting_read_alpha
  p.allele_unsupporting_read_alpha = options.allele_unsupporting_read_alpha
  p.other_allele_supporting_read_alpha = options.other_allele_supporting_read_alpha
  p.reference_matching_read_alpha = options.reference_matching_read_alpha
  p.reference_mismatching_read_alpha = options.reference_mismatching_read_alpha
  p.indel_anchoring_base_char = ord(options.indel_anchoring_base_char[0]) if options.indel_anchoring_base_char else 0
  p.reference_base_quality = options.reference_base_quality
  p.positive_strand_color = options.positive_strand_color
  p.negative_strand_color = options.negative_strand_color
  p.base_quality_cap = options.base_quality_cap
  p.mapping_quality_cap = options.mapping_quality_cap
  p.min_base_quality = options.read_requirements.min_base_quality
  p.min_mapping_quality = options.read_requirements.min_mapping_quality
  p.sort_by_haplotypes = int(options.sort_by_haplotypes)
  p.hp_tag_for_assembly_polishing = options.hp_tag_for_assembly_polishing
  p.sort_by_alt_allele_support = int(options.sort_by_alt_allele_support)
  p.random_seed = options.random_seed & 0xFFFFFFFF
  p.max_reads_per_image = 0
  p.mean_coverage = options.mean_coverage
  blank = set(getattr(options, 'channels_enum_to_blank', ()) or ())
  p.blank_channel_mask = sum(1 << i for i, e in enumerate(enums) if e in blank)
  return p


class GpuEncoder:
  """Owns one DvbEncoder handle (one per device); thin RAII wrapper over the C ABI."""

  def __init__(self, params: _lib.DvbPileupParams, device: int = 0):
    import ctypes as C
    self._lib = _lib.lib()
    self.params = params
    self.device = device
    h = C.c_void_p()
    _lib.check(self._lib.dvb_encoder_create(C.byref(params), device, C.byref(h)))
    self._h = h
    self.image_bytes = int(self._lib.dvb_image_bytes(C.byref(params)))
    self.shape = (params.height, params.width, params.num_channels + params.num_alt_channels)

  def encode_host(self, batch: packing.PackedBatch) -> np.ndarray:
    """Host arrays in -> uint8[n_images, H, W, C] host array out (dvb_encode_batch_host)."""
    import ctypes as C
    out = np.empty((batch.n_images,) + self.shape, dtype=np.uint8)
    self.last_rows_kept = np.zeros(max(batch.n_images, 1), dtype=np.int32)
    cb = batch.as_ctypes()
    _lib.check(self._lib.dvb_encode_batch_host(self._h, C.byref(cb), out.ctypes.data_as(C.c_void_p),
                                               self.last_rows_kept.ctypes.data_as(C.c_void_p)))
    return out

  def last_pair_support(self, n_pairs: int):
    """(pair_support, pair_allele_group) the device derived for the last batch that carried allele keys
    (packing.attach_alleles; dvb_encoder_last_pair_support)."""
    import ctypes as C
    support = np.zeros(max(n_pairs, 1), dtype=np.uint8)
    group = np.zeros(max(n_pairs, 1), dtype=np.uint8)
    _lib.check(self._lib.dvb_encoder_last_pair_support(self._h, n_pairs, C.c_void_p(support.ctypes.data), C.c_void_p(group.ctypes.data)))
    return support[:n_pairs], group[:n_pairs]

  def encode_classify_host(self, batch, cnn, probs: Optional[np.ndarray] = None) -> np.ndarray:
    """Host batch in -> float32[n_images, 3] genotype probabilities out (dvb_encode_classify_host):
    the images stay in HBM between the encoder and the classifier.  `batch` is anything with
    `.n_images` and `.as_ctypes()` over HOST arrays (packing.PackedBatch, a CPU synthetic.TorchBatch)."""
    import ctypes as C
    n = int(batch.n_images)
    if probs is None:
      probs = np.empty((n, 3), dtype=np.float32)
    self.last_rows_kept = np.zeros(max(n, 1), dtype=np.int32)
    cb = batch.as_ctypes()
    _lib.check(self._lib.dvb_encode_classify_host(self._h, cnn._h, C.byref(cb), C.c_void_p(probs.ctypes.data),  # pylint: disable=protected-access
                                                  self.last_rows_kept.ctypes.data_as(C.c_void_p)))
    return probs

  def encode_device(self, dev_batch: 'packing.DeviceBatch', out, rows_kept=None, stream=None) -> None:
    """Device pointers in/out; asynchronous on `stream` (a torch.cuda.Stream or None)."""
    import ctypes as C
    cb = dev_batch.as_ctypes()
    sp = C.c_void_p(stream.cuda_stream) if stream is not None else C.c_void_p(0)
    rk = C.c_void_p(rows_kept.data_ptr()) if rows_kept is not None else C.c_void_p(0)
    _lib.check(self._lib.dvb_encode_batch_device(self._h, C.byref(cb), C.c_void_p(out.data_ptr()), rk, sp))

  def check(self, stream=None) -> None:
    import ctypes as C
    sp = C.c_void_p(stream.cuda_stream) if stream is not None else C.c_void_p(0)
    _lib.check(self._lib.dvb_encoder_check(self._h, sp))

  @property
  def launch_count(self) -> int:
    return int(self._lib.dvb_encoder_launch_count(self._h))

  def close(self):
    if getattr(self, '_h', None):
      self._lib.dvb_encoder_destroy(self._h)
      self._h = None

  def __del__(self):
    try:
      self.close()
    except Exception:  # pylint: disable=broad-except
      pass


class PileupImageEncoderNative:
  """Drop-in for deepvariant.python.pileup_image_native.PileupImageEncoderNative."""

  def __init__(self, options: PileupImageOptions, device: int = 0):
    if not (options.width % 2 == 1 and options.width >= 3):
      raise ValueError(f'Width must be odd; found {options.width}')  # pileup_image_native.cc:114
    self.options = options
    self._device = device
    self._enc_full: Optional[GpuEncoder] = None   # H rows, reference band as configured
    self._enc_row: Optional[GpuEncoder] = None    # 1 reference row + 1 read row (encode_read/_reference)

  # -- handles are created lazily so that constructing the object needs no GPU ------------
  def _full(self) -> GpuEncoder:
    if self._enc_full is None:
      self._enc_full = GpuEncoder(to_params(self.options), self._device)
    return self._enc_full

  def _row(self) -> GpuEncoder:
    if self._enc_row is None:
      o = dataclasses.replace(self.options, reference_band_height=1)
      p = to_params(o, height=2)
      p.num_alt_channels = 0
      self._enc_row = GpuEncoder(p, self._device)
    return self._enc_row

  def all_channels_enum(self, alt_aligned_representation: str = '') -> List[int]:
    return all_channels_enum(self.options, alt_aligned_representation)

  def encode_reference(self, ref_bases: str) -> np.ndarray:
    """EncodeReference (pileup_image_native.cc:512-527) -> uint8[1, len(ref_bases), C]."""
    enc = self._row_for_width(len(ref_bases))
    batch = packing.pack_images(
        [packing.ImageSpec(ref_bases=ref_bases, image_start_pos=0, variant_start=-1, reads=[],
                           support=[], allele_group=[])], enc.params)
    return enc.encode_host(batch)[0, 0:1]

  def encode_read(self, dv_call: DeepVariantCall, ref_bases: str, read: Read, image_start_pos: int,
                  alt_alleles: Sequence[str]) -> Optional[np.ndarray]:
    """EncodeRead (pileup_image_native.cc:477-510): uint8[1, W, C] or None when the read is
    rejected (low mapq / low base quality at the call site)."""
    enc = self._row_for_width(len(ref_bases))
    spec = packing.image_spec_for(dv_call, ref_bases, [read], image_start_pos, list(alt_alleles),
                                  self.options)
    out = enc.encode_host(packing.pack_images([spec], enc.params))
    if enc.last_rows_kept[0] == 0:  # EncodeRead returned nullptr
      return None
    return out[0, 1:2]

  def build_pileup_for_one_sample(self, dv_call: DeepVariantCall, ref_bases: str,
                                  reads: Sequence[Read], image_start_pos: int,
                                  alt_alleles: Sequence[str]) -> np.ndarray:
    """BuildPileupForOneSample (pileup_image_native.cc:296-447) -> uint8[H, W, C]."""
    if len(ref_bases) != self.options.width:
      raise ValueError('ref_bases.size() != options.width')  # CHECK_EQ, pileup_image_native.cc:308
    enc = self._full()
    spec = packing.image_spec_for(dv_call, ref_bases, list(reads), image_start_pos,
                                  list(alt_alleles), self.options)
    return enc.encode_host(packing.pack_images([spec], enc.params))[0]

  def _row_for_width(self, width: int) -> GpuEncoder:
    if width == self.options.width:
      return self._row()
    o = dataclasses.replace(self.options, reference_band_height=1, width=width)
    p = to_params(o, height=2)
    p.num_alt_channels = 0
    key = ('w', width)
    cache = self.__dict__.setdefault('_row_cache', {})
    if key not in cache:
      cache[key] = GpuEncoder(p, self._device)
    return cache[key]
