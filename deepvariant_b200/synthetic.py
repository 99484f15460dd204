"""Synthetic 30x short-read workload for the pileup encoder.

Generates a DvbBatch (Structure-of-Arrays, include/dvb.h) directly as torch tensors on the
target device (torch's CUDA generator is counter-based Philox; seed 2101079370 + chunk):

  ref window   uniform over ACGT
  depth        n ~ Poisson(32) reads per candidate; 1% of candidates forced to n in [96, 300]
               (exercises down-sampling)
  reads        150 bp, alignment start uniform in [pos-154, pos+5]
  CIGAR        96% 150M, 2% one insertion (len U[1,10]), 2% one deletion (len U[1,10]) at a uniform
               offset; 3% additionally soft-clipped (U[1,30]) at the read start
  bases        reference with 0.5% substitutions inside the window, random outside / inserted /
               clipped; reads flagged supporting (P=0.5) carry the alt base at pos
  quals        {2: 0.01, 11: 0.04, 25: 0.10, 37: 0.85};  mapq {60: 0.90, U[0,59]: 0.10}
  strand       reverse P=0.5;  fragment_length ~ N(400, 100)
  support      1 for alt carriers; in 10% of candidates carriers are class 2 with P=0.3

Every candidate owns its reads (pair_read = arange), so the batch is also a valid input for the
CPU oracle after `.cpu()`.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict

import numpy as np
import torch

from deepvariant_b200 import _lib

READ_LEN = 150
SEED = 2101079370


class TorchBatch:
  """DvbBatch whose arrays are torch tensors (any device)."""

  def __init__(self, tensors: Dict[str, torch.Tensor], n_images: int, n_reads: int, n_pairs: int,
               ref_stride: int):
    self.tensors = tensors
    self.n_images, self.n_reads, self.n_pairs, self.ref_stride = n_images, n_reads, n_pairs, ref_stride
    self.n_bases = int(tensors['bases'].numel())
    self.n_cigar = int(tensors['cigar'].numel())

  def as_ctypes(self) -> _lib.DvbBatch:
    b = _lib.DvbBatch()
    b.n_images, b.n_reads, b.n_pairs = self.n_images, self.n_reads, self.n_pairs
    b.n_bases, b.n_cigar, b.ref_stride = self.n_bases, self.n_cigar, self.ref_stride
    for name, _ in _lib.BATCH_ARRAYS:
      setattr(b, name, C.c_void_p(self.tensors[name].data_ptr()))
    for member, k in _lib.PLANE_ARRAYS:   # optional channel planes (tensors['pair_channel_<k>'] / ['base_channel_<k>'])
      t = self.tensors.get(f'{member}_{k}')
      if t is not None:
        getattr(b, member)[k] = t.data_ptr()
    return b

  def to(self, device, non_blocking=False) -> 'TorchBatch':
    return TorchBatch({k: v.to(device, non_blocking=non_blocking) for k, v in self.tensors.items()},
                      self.n_images, self.n_reads, self.n_pairs, self.ref_stride)

  def pin(self) -> 'TorchBatch':
    return TorchBatch({k: v.pin_memory() for k, v in self.tensors.items()}, self.n_images, self.n_reads,
                      self.n_pairs, self.ref_stride)

  def input_bytes(self) -> int:
    return int(sum(v.numel() * v.element_size() for v in self.tensors.values()))

  def algorithmic_bytes(self, image_bytes: int, width: int) -> int:
    """B_enc = H*W*C + sum_reads(2*L + 4*n_cigar + 16) + W + 32 per image."""
    return int(self.n_images * (image_bytes + width + 32) + 2 * self.n_bases + 4 * self.n_cigar +
               16 * self.n_pairs)

  def to_packed(self):
    """Host numpy view for the oracle / the host C-ABI entry point."""
    from deepvariant_b200 import packing
    arrays = {}
    for name, dtype in _lib.BATCH_ARRAYS:
      a = self.tensors[name].cpu().numpy()
      if dtype == 'uint32':
        a = a.view(np.uint32)
      a = np.ascontiguousarray(a)
      arrays[name] = a if a.size else np.zeros(1, dtype=a.dtype)
    for member, k in _lib.PLANE_ARRAYS:
      t = self.tensors.get(f'{member}_{k}')
      if t is not None:
        a = np.ascontiguousarray(t.cpu().numpy())
        arrays[f'{member}_{k}'] = a if a.size else np.zeros(1, dtype=np.uint8)
    pb = packing.PackedBatch(self.n_images, self.n_reads, self.n_pairs, self.ref_stride, arrays)
    return pb


def _choice(gen, probs, values, n, device):
  """n draws from a small discrete distribution (inverse-CDF on uniform draws)."""
  cdf = torch.cumsum(torch.tensor(probs, dtype=torch.float32, device=device), 0)[:-1].contiguous()
  idx = torch.bucketize(torch.rand(n, generator=gen, device=device), cdf, right=True)
  return torch.tensor(values, device=device)[idx]


@torch.no_grad()
def make_batch(n_images: int, device='cpu', width: int = 221, seed: int = SEED, chunk: int = 0,
               mean_depth: float = 32.0, deep_fraction: float = 0.01, hp: bool = False) -> TorchBatch:
  dev = torch.device(device)
  gen = torch.Generator(device=dev)
  gen.manual_seed(seed + 7919 * chunk)
  N = n_images
  half = (width - 1) // 2
  ref_stride = (width + 15) // 16 * 16
  acgt = torch.tensor([65, 67, 71, 84], dtype=torch.uint8, device=dev)

  def rnd(shape):
    return torch.rand(shape, generator=gen, device=dev)

  def rint(lo, hi, shape):  # inclusive
    return torch.randint(lo, hi + 1, shape, generator=gen, device=dev)

  ref_idx = rint(0, 3, (N, ref_stride))
  ref = acgt[ref_idx]
  ref[:, width:] = 0
  pos = (10_000 + (torch.arange(N, device=dev) % 1_000_000) * 1_000).to(torch.int32)
  image_start = pos - half

  depth = torch.poisson(torch.full((N,), mean_depth, device=dev), generator=gen).to(torch.int64)
  deep = rnd((N,)) < deep_fraction
  depth = torch.where(deep, rint(96, 300, (N,)), depth)
  pair_begin = torch.zeros(N + 1, dtype=torch.int64, device=dev)
  pair_begin[1:] = torch.cumsum(depth, 0)
  R = int(pair_begin[-1].item())
  img_of = torch.repeat_interleave(torch.arange(N, device=dev), depth)

  L = READ_LEN
  # ---- CIGAR ----
  u = rnd((R,))
  kind = torch.zeros(R, dtype=torch.int64, device=dev)          # 0 plain, 1 ins, 2 del
  kind[u < 0.02] = 1
  kind[(u >= 0.02) & (u < 0.04)] = 2
  clip = torch.where(rnd((R,)) < 0.03, rint(1, 30, (R,)), torch.zeros(R, dtype=torch.int64, device=dev))
  ilen = rint(1, 10, (R,))
  aligned = L - clip                                             # bases after the soft clip
  # first match length a: ins needs a + ilen + b = aligned (b>=1); del needs a + b = aligned
  a_max = torch.where(kind == 1, aligned - ilen - 1, aligned - 1)
  a = 1 + (rnd((R,)) * a_max.to(torch.float32)).to(torch.int64).clamp(max=a_max - 1)
  b = torch.where(kind == 1, aligned - ilen - a, aligned - a)
  a = torch.where(kind == 0, aligned, a)
  ops = torch.zeros((R, 4), dtype=torch.int64, device=dev)
  valid = torch.zeros((R, 4), dtype=torch.bool, device=dev)
  ops[:, 0] = (clip << 4) | 4
  valid[:, 0] = clip > 0
  ops[:, 1] = (a << 4) | 0
  valid[:, 1] = True
  ops[:, 2] = (ilen << 4) | torch.where(kind == 1, 1, 2)
  valid[:, 2] = kind > 0
  ops[:, 3] = (b << 4) | 0
  valid[:, 3] = kind > 0
  n_cig = valid.sum(1)
  cigar = ops[valid].to(torch.int32)
  cig_begin = torch.zeros(R + 1, dtype=torch.int64, device=dev)
  cig_begin[1:] = torch.cumsum(n_cig, 0)

  # ---- positions ----
  start = pos[img_of].to(torch.int64) - 154 + rint(0, 159, (R,))   # alignment position of 1st M base
  # ---- bases ----
  j = torch.arange(L, device=dev).view(1, L)
  m = j - clip.view(R, 1)                                        # index in the aligned part
  av, iv = a.view(R, 1), ilen.view(R, 1)
  k1 = (kind == 1).view(R, 1)
  k2 = (kind == 2).view(R, 1)
  inserted = k1 & (m >= av) & (m < av + iv)
  refoff = torch.where(k1 & (m >= av + iv), m - iv, torch.where(k2 & (m >= av), m + iv, m))
  has_ref = (m >= 0) & ~inserted
  col = start.view(R, 1) + refoff - image_start[img_of].to(torch.int64).view(R, 1)
  in_win = has_ref & (col >= 0) & (col < width)
  colc = col.clamp(0, width - 1)
  ref_at = ref[img_of.view(R, 1).expand(R, L), colc]
  rand_base = acgt[rint(0, 3, (R, L))]
  sub = rnd((R, L)) < 0.005
  bases = torch.where(in_win & ~sub, ref_at, rand_base)
  # alt carriers
  carrier = rnd((R,)) < 0.5
  at_pos = in_win & (col == half)
  ref_center = ref[:, half]
  alt_base = acgt[(torch.bucketize(ref_center.to(torch.int64), acgt.to(torch.int64)) + 1 + rint(0, 2, (N,))) % 4]  # != ref base
  bases = torch.where(at_pos & carrier.view(R, 1), alt_base[img_of].view(R, 1).expand(R, L), bases)
  covers = at_pos.any(1)
  multi = rnd((N,)) < 0.10
  support = torch.where(carrier & covers, 1, 0)
  other = multi[img_of] & (rnd((R,)) < 0.3)
  support = torch.where((support == 1) & other, 2, support).to(torch.uint8)

  quals = _choice(gen, [0.01, 0.04, 0.10, 0.85], [2, 11, 25, 37], R * L, dev).to(torch.uint8)
  mapq = torch.where(rnd((R,)) < 0.9, torch.full((R,), 60, device=dev), rint(0, 59, (R,))).to(torch.int32)
  flags = (rnd((R,)) < 0.5).to(torch.uint8)
  hp_t = torch.zeros(R, dtype=torch.int32, device=dev)
  if hp:
    hv = _choice(gen, [0.2, 0.4, 0.4], [0, 1, 2], R, dev).to(torch.int32)
    hp_t = hv
    flags = flags | 4
    flags = torch.where(rnd((R,)) < 0.02, flags | 2, flags)
  fraglen = (400 + 100 * torch.randn(R, generator=gen, device=dev)).to(torch.int32)
  fraglen = torch.where(rnd((R,)) < 0.5, fraglen, -fraglen)
  name_rank = torch.randperm(R, generator=gen, device=dev).to(torch.int32)
  seq_begin = torch.arange(R + 1, device=dev, dtype=torch.int64) * L

  t = {
      'ref_bases': ref.reshape(-1).contiguous(),
      'image_start_pos': image_start.to(torch.int32),
      'variant_start': pos,
      'pair_begin': pair_begin,
      'pair_read': torch.arange(R, device=dev, dtype=torch.int32),
      'pair_support': support,
      'pair_allele_group': torch.zeros(max(R, 1), dtype=torch.uint8, device=dev)[:R] if R else torch.zeros(1, dtype=torch.uint8, device=dev),
      'read_pos': start.to(torch.int32),
      'read_sort_pos': start.to(torch.int32),
      'read_mapq': mapq,
      'read_flags': flags.to(torch.uint8),
      'read_fragment_length': fraglen,
      'read_hp': hp_t,
      'read_name_rank': name_rank,
      'read_seq_begin': seq_begin,
      'read_cigar_begin': cig_begin,
      'bases': bases.reshape(-1).contiguous(),
      'quals': quals,
      'cigar': cigar,
  }
  for k in list(t):
    if t[k].numel() == 0:
      t[k] = torch.zeros(1, dtype=t[k].dtype, device=dev)
  tb = TorchBatch(t, N, R, R, ref_stride)
  tb.n_bases = R * L
  tb.n_cigar = int(cig_begin[-1].item())
  return tb
