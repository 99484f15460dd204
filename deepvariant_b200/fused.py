"""Fused make_examples + call_variants: region -> packed reads -> dvb_encode_classify_host -> CallVariantsOutput shards.

The reference offers the same short-cut as `fast_pipeline` (deepvariant/fast_pipeline.cc: make_examples streams its examples to
call_variants through shared memory, stream_examples.cc:94-156) - the tf.Example files between the two stages are skipped, the
CallVariantsOutput files postprocess_variants reads are the same.  Here the pileup images never leave HBM: the regions of one
task are packed on the host (dvb_pack_region_from_bam), concatenated until a classifier-sized batch is full, and handed to
dvb_encode_classify_host (pileup encoder -> Inception-v3 -> genotype probabilities) in one call; the CallVariantsOutput records
(variant + alt_allele_indices + rounded likelihoods) are written by the same native writer call_variants uses.
Trimmed / alt-aligned pileups (PACBIO, indel candidates of --alt_aligned_pileup) are composed on the host from several
encoder images, so their finished images go through dvb_cnn_forward_host instead.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np

from deepvariant_b200 import _lib, packing


def concat_packed(batches: Sequence[packing.PackedBatch]) -> packing.PackedBatch:
  """One DvbBatch holding the images of all `batches` (read tables stacked, indices shifted)."""
  batches = [b for b in batches if b is not None and b.n_images]
  if len(batches) == 1:
    return batches[0]
  if not batches:
    raise ValueError('nothing to concatenate')
  ref_stride = batches[0].ref_stride
  out = {name: [] for name, _ in _lib.BATCH_ARRAYS}
  reads = pairs = bases = cigar = 0
  for b in batches:
    if b.ref_stride != ref_stride:
      raise ValueError('batches of different image widths')
    a = b.arrays
    nb, nc = int(a['read_seq_begin'][b.n_reads]), int(a['read_cigar_begin'][b.n_reads])
    out['ref_bases'].append(a['ref_bases'][:b.n_images * ref_stride])
    out['image_start_pos'].append(a['image_start_pos'][:b.n_images])
    out['variant_start'].append(a['variant_start'][:b.n_images])
    out['pair_begin'].append(a['pair_begin'][:b.n_images] + pairs)
    out['pair_read'].append(a['pair_read'][:b.n_pairs] + np.int32(reads))
    out['pair_support'].append(a['pair_support'][:b.n_pairs])
    out['pair_allele_group'].append(a['pair_allele_group'][:b.n_pairs])
    for k in ('read_pos', 'read_sort_pos', 'read_mapq', 'read_flags', 'read_fragment_length', 'read_hp', 'read_name_rank'):
      out[k].append(a[k][:b.n_reads])
    out['read_seq_begin'].append(a['read_seq_begin'][:b.n_reads] + bases)
    out['read_cigar_begin'].append(a['read_cigar_begin'][:b.n_reads] + cigar)
    out['bases'].append(a['bases'][:nb])
    out['quals'].append(a['quals'][:nb])
    out['cigar'].append(a['cigar'][:nc])
    reads += b.n_reads; pairs += b.n_pairs; bases += nb; cigar += nc
  out['pair_begin'].append(np.array([pairs], dtype=np.int64))
  out['read_seq_begin'].append(np.array([bases], dtype=np.int64))
  out['read_cigar_begin'].append(np.array([cigar], dtype=np.int64))
  arrays = {}
  for name, dtype in _lib.BATCH_ARRAYS:
    v = np.ascontiguousarray(np.concatenate(out[name]).astype(np.dtype(dtype), copy=False))
    arrays[name] = v if v.size else np.zeros(1, dtype=np.dtype(dtype))
  merged = packing.PackedBatch(sum(b.n_images for b in batches), reads, pairs, ref_stride, arrays)
  # allele keys (device-side pair support): all batches carry them, under the same allele-counter options, or none does
  with_keys = [b.support is not None for b in batches]
  if any(with_keys):
    if not all(with_keys) or len({b.support[:2] + (b.support[2] & ~_lib.SUPPORT_REPEATED_KEYS,) for b in batches}) != 1:
      raise ValueError('batches with and without allele keys (or of different allele-counter options) cannot be concatenated')
    groups = ['allele_group' in b.arrays for b in batches]
    if any(groups) != all(groups):
      raise ValueError('allele_group present in some batches only')
    alleles = abases = 0
    parts = {name: [] for name, _ in _lib.ALLELE_ARRAYS}
    for b in batches:
      a = b.arrays
      na = int(a['allele_begin'][b.n_images])
      nab = int(a['allele_bases_begin'][na])
      parts['allele_begin'].append(a['allele_begin'][:b.n_images] + alleles)
      parts['allele_bases_begin'].append(a['allele_bases_begin'][:na] + abases)
      for k in ('allele_type', 'allele_class') + (('allele_group',) if groups[0] else ()):
        parts[k].append(a[k][:na])
      parts['allele_bases'].append(a['allele_bases'][:nab])
      parts['image_ref_run'].append(a['image_ref_run'][:b.n_images])
      if groups[0]:
        parts['image_group_default'].append(a['image_group_default'][:b.n_images])
      alleles += na; abases += nab
    parts['allele_begin'].append(np.array([alleles], dtype=np.int64))
    parts['allele_bases_begin'].append(np.array([abases], dtype=np.int64))
    for name, dtype in _lib.ALLELE_ARRAYS:
      if parts[name]:
        v = np.ascontiguousarray(np.concatenate(parts[name]).astype(np.dtype(dtype), copy=False))
        arrays[name] = v if v.size else np.zeros(1, dtype=np.dtype(dtype))
    flags = 0
    for b in batches:
      flags |= b.support[2]
    merged.support = batches[0].support[:2] + (flags,)
  return merged


class FusedCaller:
  """Sink of an ExamplesGenerator in fused mode: collects the planned images of region after region, runs encoder + classifier
  on classifier-sized batches and writes CallVariantsOutput records in example order."""

  def __init__(self, encoder, cnn, cvo_path: str, batch_images: int = 2048, gl_precision: int = 10):
    from deepvariant_b200 import records
    self.encoder, self.cnn = encoder, cnn
    self.batch_images = int(batch_images)
    self.writer = records.NativeCvoWriter(cvo_path, gl_precision)
    self._packed: List[packing.PackedBatch] = []
    self._images: List[np.ndarray] = []
    self._order: List[tuple] = []            # ('p' | 'i', n, plans) in arrival order
    self._pending = 0
    self.n_examples = 0
    self.n_batches = 0
    self.probabilities: Optional[List[np.ndarray]] = None    # set to [] to keep them (tests)

  def add_packed(self, plans, packed: packing.PackedBatch) -> None:
    if not plans:
      return
    if packed.n_images != len(plans):
      raise ValueError('one packed image per plan expected')
    if self._packed and (self._packed[-1].support is None) != (packed.support is None):
      self.flush()       # batches with allele keys (device-side support) and batches with name-searched support do not share a launch
    self._packed.append(packed)
    self._order.append(('p', len(plans), list(plans)))
    self._pending += len(plans)
    if self._pending >= self.batch_images:
      self.flush()

  def add_images(self, plans, images: np.ndarray) -> None:
    if not plans:
      return
    self._images.append(np.ascontiguousarray(images, dtype=np.uint8))
    self._order.append(('i', len(plans), list(plans)))
    self._pending += len(plans)
    if self._pending >= self.batch_images:
      self.flush()

  def flush(self) -> None:
    if not self._pending:
      return
    from deepvariant_b200 import make_examples_native as men, records
    probs_p = self.encoder.encode_classify_host(concat_packed(self._packed), self.cnn) if self._packed else None
    probs_i = self.cnn.forward_host(np.concatenate(self._images)) if self._images else None
    kp = ki = 0
    probs, variants, alts = [], [], []
    for kind, n, plans in self._order:
      if kind == 'p':
        probs.append(probs_p[kp:kp + n]); kp += n
      else:
        probs.append(probs_i[ki:ki + n]); ki += n
      for plan in plans:
        variants.append(plan.variant.serialize())
        alts.append(men.encode_alt_alleles(plan.variant, plan.alt_combination)[0])
    probs = np.concatenate(probs)
    vbeg = np.zeros(len(variants) + 1, dtype=np.int64)
    vbeg[1:] = np.cumsum([len(v) for v in variants])
    abeg = np.zeros(len(alts) + 1, dtype=np.int64)
    abeg[1:] = np.cumsum([len(x) for x in alts])
    meta = records.BatchMeta(len(variants), np.frombuffer(b''.join(variants), dtype=np.uint8).copy(), vbeg,
                             np.frombuffer(b''.join(alts) or b'\0', dtype=np.uint8).copy(), abeg)
    self.writer.write_batch(meta, probs)
    if self.probabilities is not None:
      self.probabilities.append(probs.copy())
    self.n_examples += len(variants)
    self.n_batches += 1
    self._packed, self._images, self._order, self._pending = [], [], [], 0

  def close(self) -> int:
    self.flush()
    return self.writer.close()
