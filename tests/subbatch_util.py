"""Test helper: the sub-batch holding only the given images (read table kept whole)."""
import numpy as np

from deepvariant_b200 import packing


def take_images(pb: packing.PackedBatch, idx) -> packing.PackedBatch:
  idx = np.asarray(idx, dtype=np.int64)
  a = pb.arrays
  begin = a['pair_begin']
  lens = begin[idx + 1] - begin[idx]
  new_begin = np.zeros(len(idx) + 1, dtype=np.int64)
  new_begin[1:] = np.cumsum(lens)
  sel = np.concatenate([np.arange(begin[i], begin[i + 1]) for i in idx]) if len(idx) else np.zeros(0, np.int64)
  out = dict(a)
  out['ref_bases'] = np.ascontiguousarray(a['ref_bases'].reshape(pb.n_images, pb.ref_stride)[idx].reshape(-1))
  out['image_start_pos'] = np.ascontiguousarray(a['image_start_pos'][idx])
  out['variant_start'] = np.ascontiguousarray(a['variant_start'][idx])
  out['pair_begin'] = new_begin
  for k in ('pair_read', 'pair_support', 'pair_allele_group'):
    v = np.ascontiguousarray(a[k][sel])
    out[k] = v if v.size else np.zeros(1, dtype=a[k].dtype)
  return packing.PackedBatch(len(idx), pb.n_reads, int(new_begin[-1]), pb.ref_stride, out)
