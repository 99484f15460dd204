"""Values of the plane-backed pileup channels (include/dvb.h "channel planes").

Nine of the reference's channels are functions of data the pileup loop does not hold: DeepVariantCall maps keyed by read-name
strings, or per-base aux tags of the alignment records.  The caller that owns that data computes the channel's pixel VALUE once per
(image, read) pair or once per base and hands it to the encoder as a plane (DvbBatch.pair_channel / base_channel); the CUDA kernel
places the bytes where FillReadBase would.  This module is that caller-side step:

  per (image, read) pair -- string logic here, colour arithmetic in libdvb (csrc/dvb_channels.cu):
    allele_frequency            ReadAlleleFrequency + AlleleFrequencyColor   channels/allele_frequency_channel.cc:76-118
    read_supports_variant_fuzzy ReadSupportsAlt + SupportsAltColor            channels/read_supports_variant_fuzzy_channel.cc:68-310
    allele_sample_probability   FillReadBase + ScaleColor                     channels/allele_sample_probability_channel.cc:48-101
  per base -- one native pass per read (csrc/dvb_channels.cu):
    base_methylation / base_6ma                      channels/base_methylation_channel.cc:54-99, base_6ma_channel.cc:54-99
    homopolymer_insertion_quality / _deletion_       channels/homopolymer_indel_quality_channel.cc:68-183 (tp tag)
    inter_homopolymer_insertion_quality              channels/inter_homopolymer_insertion_quality_channel.cc:75-127 (t0 tag)
  mean_coverage is painted by the kernel itself (DvbPileupParams.mean_coverage).
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence

import numpy as np

from deepvariant_b200 import _lib
from deepvariant_b200.protos import DeepVariantCall, Read

# read_supports_variant_fuzzy_channel.cc:62-67
FUZZY_ONE_BASE, FUZZY_TWO_BASES, FUZZY_THREE_BASES = 10, 9, 8
K_READ_SUPPORT_ALT_WITHIN = {FUZZY_ONE_BASE: 0.90, FUZZY_TWO_BASES: 0.80, FUZZY_THREE_BASES: 0.70}
K_MAX_PIXEL_VALUE_AS_FLOAT = np.float32(254.0)


def _hp_value(read: Read) -> int:
  return read.hp_values[0] if read.hp_values else 0


# ---- allele_frequency -----------------------------------------------------------------------------------------------------------

def read_allele_frequency(dv_call: DeepVariantCall, read_key: str, alt_alleles: Sequence[str]) -> float:
  """ReadAlleleFrequency (allele_frequency_channel.cc:90-118): over ALL alts of the variant in order, the frequency of the first
  alt of THIS image whose support lists the read; 0 when there is none (or the alt has no entry in allele_frequency)."""
  for alt_allele in dv_call.variant.alternate_bases:
    names = dv_call.allele_support.get(alt_allele)
    if names is None:
      continue
    if alt_allele in alt_alleles and read_key in names:
      return dv_call.allele_frequency.get(alt_allele, 0.0)
  return 0.0


def allele_frequency_color(allele_frequency: float, min_non_zero_allele_frequency: float) -> int:
  return int(_lib.lib().dvb_channel_allele_frequency_color(allele_frequency, min_non_zero_allele_frequency))


# ---- read_supports_variant_fuzzy -------------------------------------------------------------------------------------------------

def _allele_phases(values: Optional[Sequence[int]], num_alt_alleles: int) -> List[int]:
  """CalculateAlelePhases (:75-97): phase of alt i = info[key].values[i + 1] (values[0] belongs to the reference allele)."""
  out = [0] * num_alt_alleles
  if values is not None:
    for i in range(num_alt_alleles):
      out[i] = values[i + 1] if len(values) > i + 1 else 0
  return out


def _calculate_read_support(all_alt_alleles: Sequence[str], allele_support: Dict[str, List[str]], alt_allele: str, alt_alleles: Sequence[str],
                            key: str, read: Read, alt_allele_phases: Sequence[int]) -> int:
  """CalculateReadSupport (:118-190)."""
  for read_name in allele_support[alt_allele]:
    if read_name != key:
      continue
    if alt_allele in alt_alleles:
      return 1
    for image_alt in alt_alleles:
      global_index = 0
      for candidate_alt in all_alt_alleles:
        if candidate_alt == image_alt:
          break
        global_index += 1
      if global_index >= len(alt_allele_phases):
        raise ValueError('image alt allele is not an alt of the candidate')   # CHECK_LT
      hp_value = _hp_value(read)
      phase = alt_allele_phases[global_index]
      if phase == 0 or hp_value == 0 or phase == hp_value:
        d = abs(len(image_alt) - len(alt_allele))
        if d == 1:
          return FUZZY_ONE_BASE
        if d == 2:
          return FUZZY_TWO_BASES
    return 2
  return 0


def fuzzy_read_supports_alt(dv_call: DeepVariantCall, read: Read, alt_alleles: Sequence[str]) -> int:
  """ReadSupportsVariantFuzzyChannel::ReadSupportsAlt (:208-286): 1 exact, 10 / 9 an indel one / two bases away on a compatible
  phase, 2 another alt, 0 reference."""
  key = read.key()
  alts = dv_call.variant.alternate_bases
  phases = _allele_phases(dv_call.variant.alt_ps, len(alts))
  for alt_allele in alts:
    if alt_allele in dv_call.allele_support:
      s = _calculate_read_support(alts, dv_call.allele_support, alt_allele, alt_alleles, key, read, phases)
      if s in (1, FUZZY_ONE_BASE, FUZZY_TWO_BASES):
        return s
  for alt_allele in dv_call.variant.alternate_bases_rejected:
    if alt_allele in dv_call.rejected_allele_support:
      s = _calculate_read_support(alts, dv_call.rejected_allele_support, alt_allele, alt_alleles, key, read, phases)
      if s != 0:
        return s
  if dv_call.ref_support:
    ref = dv_call.variant.reference_bases
    s = _calculate_read_support(alts, {ref: list(dv_call.ref_support)}, ref, alt_alleles, key, read, phases)
    if s in (FUZZY_ONE_BASE, FUZZY_TWO_BASES):
      return s
  return 0


def fuzzy_supports_alt_color(read_supports_alt: int, options) -> int:
  """SupportsAltColor (:288-310): int(254 * alpha) in float32."""
  if read_supports_alt == 0:
    alpha = options.allele_unsupporting_read_alpha
  elif read_supports_alt == 1:
    alpha = options.allele_supporting_read_alpha
  elif read_supports_alt in K_READ_SUPPORT_ALT_WITHIN:
    alpha = K_READ_SUPPORT_ALT_WITHIN[read_supports_alt]
  elif read_supports_alt == 2:
    alpha = options.other_allele_supporting_read_alpha
  else:
    raise ValueError(f'read_supports_alt can only be 0/1/8/9/10/2, not {read_supports_alt}')
  return int(K_MAX_PIXEL_VALUE_AS_FLOAT * np.float32(alpha))


# ---- allele_sample_probability --------------------------------------------------------------------------------------------------

def allele_sample_probability_color(dv_call: DeepVariantCall, read_key: str) -> int:
  """AlleleSampleProbabilityChannel::FillReadBase (:48-79): sqrt-scaled share of the reads that support the read's allele; the
  reference walks the allele_support protobuf MAP and stops at the first allele listing the read, so with several alleles its total
  depends on the map's (unspecified, hash-seeded) iteration order; here the order is fixed: the variant's alts, then other keys sorted."""
  total_reads = 0
  supporting = None
  alts = [a for a in dv_call.variant.alternate_bases if a in dv_call.allele_support]
  for allele in alts + sorted(k for k in dv_call.allele_support if k not in alts):
    names = dv_call.allele_support[allele]
    total_reads += len(names)
    if read_key in names:
      supporting = len(names)
      break
  if supporting is None:
    supporting = len(dv_call.ref_support)
  total_reads += len(dv_call.ref_support)
  return int(_lib.lib().dvb_channel_allele_sample_probability_color(supporting, float(total_reads)))


# ---- per (image, read) planes of one image ----------------------------------------------------------------------------------------

def pair_planes(dv_call: DeepVariantCall, reads: Sequence[Read], alt_alleles: Sequence[str], options) -> Dict[int, List[int]]:
  """{plane slot: one value per read} for the per-pair plane channels options.channels lists."""
  out: Dict[int, List[int]] = {}
  names = set(options.channels)
  if 'allele_frequency' in names:
    m = float(getattr(options, 'min_non_zero_allele_frequency', 0.00001))
    out[0] = [allele_frequency_color(read_allele_frequency(dv_call, r.key(), alt_alleles), m) for r in reads]
  if 'read_supports_variant_fuzzy' in names:
    out[1] = [fuzzy_supports_alt_color(fuzzy_read_supports_alt(dv_call, r, alt_alleles), options) for r in reads]
  if 'allele_sample_probability' in names:
    out[2] = [allele_sample_probability_color(dv_call, r.key()) for r in reads]
  return out


# ---- per-base planes of one read -----------------------------------------------------------------------------------------------------

def _u8(n: int) -> np.ndarray:
  return np.zeros(max(n, 1), dtype=np.uint8)


def base_modification_plane(read: Read, key: str) -> np.ndarray:
  """ScaleColorVector(GetBaseModification(read), 255); a read without that modification leaves its pixels unwritten (0)."""
  n = len(read.aligned_sequence)
  v = (read.base_modifications or {}).get(key)
  out = _u8(n)
  if v:
    src = np.frombuffer(bytes(v), dtype=np.uint8)
    if src.size != n:
      raise ValueError(f'{key}: {src.size} modification bytes for {n} bases')   # vector.at(read_index) would throw
    _lib.check(_lib.lib().dvb_channel_base_modification_plane(C.c_void_p(src.ctypes.data), n, C.c_void_p(out.ctypes.data)))
  return out[:n]


def hmer_quality_plane(read: Read, is_deletion: bool) -> np.ndarray:
  """HomoPolymerInDelQuality(read, is_deletion) from the read's tp tag (GetTPValues: int8 per base, zero beyond the tag)."""
  n = len(read.aligned_sequence)
  seq = np.frombuffer(bytes(read.aligned_sequence), dtype=np.uint8)
  qual = np.frombuffer(bytes(read.aligned_quality), dtype=np.uint8)
  out = _u8(n)
  tp_ptr = C.c_void_p(0)
  tp = None
  if read.tp_values:
    tp = np.zeros(n, dtype=np.int8)
    k = min(n, len(read.tp_values))
    tp[:k] = np.array(read.tp_values[:k], dtype=np.int64).astype(np.int8)
    tp_ptr = C.c_void_p(tp.ctypes.data)
  if n:
    _lib.check(_lib.lib().dvb_channel_hmer_quality_plane(C.c_void_p(seq.ctypes.data), C.c_void_p(qual.ctypes.data), n, tp_ptr, int(is_deletion),
                                                         C.c_void_p(out.ctypes.data)))
  return out[:n]


def t0_plane(read: Read) -> np.ndarray:
  """GetT0QualityValues(read) from the t0 tag."""
  n = len(read.aligned_sequence)
  out = _u8(n)
  t0 = bytes(read.t0_value or b'')
  if n:
    _lib.check(_lib.lib().dvb_channel_t0_plane(n, t0, len(t0), C.c_void_p(out.ctypes.data)))
  return out[:n]


BASE_PLANE_OF_CHANNEL = {23: 0, 24: 1, 28: 2, 29: 3, 30: 4}


def base_plane(read: Read, slot: int) -> np.ndarray:
  if slot == 0:
    return base_modification_plane(read, '5mC')
  if slot == 1:
    return base_modification_plane(read, '6mA')
  if slot == 2:
    return hmer_quality_plane(read, False)
  if slot == 3:
    return hmer_quality_plane(read, True)
  return t0_plane(read)
