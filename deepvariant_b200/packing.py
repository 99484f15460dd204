"""Packs candidates + reads into the Structure-of-Arrays batch of include/dvb.h (DvbBatch).

This is host-side glue between the reference-shaped objects (DeepVariantCall, Read) and the
CUDA encoder.  The string work the reference does per read per image — the read-name search
in ReadSupportsVariantChannel::ReadSupportsAlt (channels/read_supports_variant_channel.cc:75-104),
the allele-group map of BuildPileupForOneSample (pileup_image_native.cc:345-360,384-392) and the
(fragment_name, read_number) tie-break of SortImageRows (pileup_image_native.cc:75-102) — is
resolved here ONCE into small integers so that the device works on integers only.
"""
from __future__ import annotations

import ctypes as C
import dataclasses
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from deepvariant_b200 import _lib
from deepvariant_b200.protos import DeepVariantCall, Read

READ_REVERSE_STRAND = 1
READ_SUPPLEMENTARY = 2
READ_HAS_HP = 4
READ_HP_MULTI = 8


@dataclasses.dataclass
class ImageSpec:
  """One BuildPileupForOneSample call."""
  ref_bases: str
  image_start_pos: int
  variant_start: int
  reads: List[Read]                 # in InMemoryReader::Query order
  support: List[int]                # per read: 0 / 1 / 2
  allele_group: List[int]           # per read (only used with sort_by_alt_allele_support)
  sort_positions: Optional[List[int]] = None  # alignment positions before trimming
  pair_channels: Optional[Dict[int, List[int]]] = None   # per-pair channel planes: slot -> one pixel value per read (channels.pair_planes)


PAIR_PLANE_CHANNELS = ('allele_frequency', 'read_supports_variant_fuzzy', 'allele_sample_probability')


def read_supports_alt(dv_call: DeepVariantCall, read_key: str, alt_alleles: Sequence[str]) -> int:
  """ReadSupportsAlt (channels/read_supports_variant_channel.cc:75-104): iterate ALL alternate
  bases in order, first name match wins: 1 if that alt is in alt_alleles, else 2; 0 if none."""
  for alt_allele in dv_call.variant.alternate_bases:
    names = dv_call.allele_support.get(alt_allele)
    if names is None:
      continue
    if read_key in names:
      return 1 if alt_allele in alt_alleles else 2
  return 0


def allele_groups(dv_call: DeepVariantCall, reads: Sequence[Read]) -> List[int]:
  """read_name_to_allele_group_map (pileup_image_native.cc:345-360): later alts overwrite
  earlier ones; reads in no alt's support list get num_alt_alleles (sorts last)."""
  n_alt = len(dv_call.variant.alternate_bases)
  m: Dict[str, int] = {}
  for i, alt in enumerate(dv_call.variant.alternate_bases):
    for name in dv_call.allele_support.get(alt, ()):
      m[name] = i
  return [m.get(r.key(), n_alt) for r in reads]


def image_spec_for(dv_call: DeepVariantCall, ref_bases: str, reads: List[Read], image_start_pos: int,
                   alt_alleles: List[str], options, sort_positions: Optional[List[int]] = None) -> ImageSpec:
  support_sets = {alt: set(names) for alt, names in dv_call.allele_support.items()}
  view = DeepVariantCall(variant=dv_call.variant, allele_support=support_sets)
  support = [read_supports_alt(view, r.key(), alt_alleles) for r in reads]
  groups = allele_groups(dv_call, reads) if options.sort_by_alt_allele_support else [0] * len(reads)
  pair_channels = None
  if any(c in PAIR_PLANE_CHANNELS for c in options.channels):
    from deepvariant_b200 import channels   # the value functions call into libdvb
    pair_channels = channels.pair_planes(dv_call, reads, alt_alleles, options)
  return ImageSpec(ref_bases=ref_bases, image_start_pos=image_start_pos,
                   variant_start=dv_call.variant.start, reads=reads, support=support,
                   allele_group=groups, sort_positions=sort_positions, pair_channels=pair_channels)


@dataclasses.dataclass
class PackedBatch:
  """Host numpy arrays laid out exactly as DvbBatch expects."""
  n_images: int
  n_reads: int
  n_pairs: int
  ref_stride: int
  arrays: Dict[str, np.ndarray]
  # (min_mapping_quality, min_base_quality, DVB_SUPPORT_* flags) of the allele counter when the batch carries allele keys
  # (arrays['allele_begin'] ...): pair_support / pair_allele_group are then derived on the device (attach_alleles below)
  support: Optional[Tuple[int, int, int]] = None

  def as_ctypes(self) -> _lib.DvbBatch:
    b = _lib.DvbBatch()
    b.n_images = self.n_images
    b.n_reads = self.n_reads
    b.n_pairs = self.n_pairs
    b.n_bases = int(self.arrays['read_seq_begin'][-1])
    b.n_cigar = int(self.arrays['read_cigar_begin'][-1])
    b.ref_stride = self.ref_stride
    for name, dtype in _lib.BATCH_ARRAYS:
      a = self.arrays[name]
      assert a.dtype == np.dtype(dtype) and a.flags['C_CONTIGUOUS'], name
      setattr(b, name, a.ctypes.data_as(C.c_void_p))
    if self.support is not None:
      for name, dtype in _lib.ALLELE_ARRAYS:
        a = self.arrays.get(name)
        if a is None:
          continue
        assert a.dtype == np.dtype(dtype) and a.flags['C_CONTIGUOUS'], name
        setattr(b, name, a.ctypes.data_as(C.c_void_p))
      b.n_alleles = int(self.arrays['allele_begin'][self.n_images])
      b.n_allele_bases = int(self.arrays['allele_bases_begin'][b.n_alleles])
      b.support_min_mapping_quality, b.support_min_base_quality, b.support_flags = self.support
    for member, k in _lib.PLANE_ARRAYS:
      a = self.arrays.get(f'{member}_{k}')
      if a is not None:
        want = self.n_pairs if member == 'pair_channel' else b.n_bases
        assert a.dtype == np.uint8 and a.flags['C_CONTIGUOUS'] and a.size >= want, (member, k, a.size, want)
        getattr(b, member)[k] = a.ctypes.data
    return b

  def input_bytes(self) -> int:
    return int(sum(a.nbytes for a in self.arrays.values()))


class DeviceBatch:
  """The same batch resident in HBM (torch tensors own the memory)."""

  def __init__(self, packed: PackedBatch, device, stream=None, pinned: Optional[Dict] = None):
    import torch
    self.n_images, self.n_reads, self.n_pairs = packed.n_images, packed.n_reads, packed.n_pairs
    self.ref_stride = packed.ref_stride
    self.tensors = {}
    for name, _ in _lib.BATCH_ARRAYS:
      a = packed.arrays[name]
      t = torch.from_numpy(a.view(np.int32) if a.dtype == np.uint32 else a)
      self.tensors[name] = t.to(device, non_blocking=True)
    for member, k in _lib.PLANE_ARRAYS:
      a = packed.arrays.get(f'{member}_{k}')
      if a is not None:
        self.tensors[f'{member}_{k}'] = torch.from_numpy(a).to(device, non_blocking=True)
    self.n_bases = int(packed.arrays['read_seq_begin'][-1])
    self.n_cigar = int(packed.arrays['read_cigar_begin'][-1])

  def as_ctypes(self) -> _lib.DvbBatch:
    b = _lib.DvbBatch()
    b.n_images, b.n_reads, b.n_pairs = self.n_images, self.n_reads, self.n_pairs
    b.n_bases, b.n_cigar, b.ref_stride = self.n_bases, self.n_cigar, self.ref_stride
    for name, _ in _lib.BATCH_ARRAYS:
      setattr(b, name, C.c_void_p(self.tensors[name].data_ptr()))
    for member, k in _lib.PLANE_ARRAYS:
      t = self.tensors.get(f'{member}_{k}')
      if t is not None:
        getattr(b, member)[k] = t.data_ptr()
    return b


def name_ranks(reads: Sequence[Read]) -> np.ndarray:
  """Dense rank of (fragment_name, read_number) — std::tuple<std::string,int> operator<
  (pileup_image_native.cc:98-101): byte-wise string compare, then int compare."""
  keys = [(r.fragment_name.encode(), r.read_number) for r in reads]
  order = sorted(set(keys))
  rank = {k: i for i, k in enumerate(order)}
  return np.array([rank[k] for k in keys], dtype=np.uint32)


def pack_images(specs: Sequence[ImageSpec], params: _lib.DvbPileupParams) -> PackedBatch:
  """Builds a PackedBatch; reads shared between images (same object) are stored once."""
  width = params.width
  ref_stride = (width + 15) // 16 * 16
  n_images = len(specs)
  ref = np.zeros((n_images, ref_stride), dtype=np.uint8)
  image_start = np.zeros(n_images, dtype=np.int32)
  variant_start = np.zeros(n_images, dtype=np.int32)
  pair_begin = np.zeros(n_images + 1, dtype=np.int64)

  read_index: Dict[int, int] = {}
  reads: List[Read] = []
  sort_pos: List[int] = []
  pair_read: List[int] = []
  pair_support: List[int] = []
  pair_group: List[int] = []
  enums = [params.channels[c] for c in range(params.num_channels)]
  pair_slots = sorted(k for e, (member, k) in _lib.PLANE_OF_CHANNEL.items() if member == 'pair_channel' and e in enums)
  base_slots = sorted(k for e, (member, k) in _lib.PLANE_OF_CHANNEL.items() if member == 'base_channel' and e in enums)
  pair_planes: Dict[int, List[int]] = {k: [] for k in pair_slots}
  for i, s in enumerate(specs):
    rb = s.ref_bases.encode() if isinstance(s.ref_bases, str) else bytes(s.ref_bases)
    if len(rb) != width:
      raise ValueError(f'ref_bases has {len(rb)} bases, expected width {width}')
    ref[i, :width] = np.frombuffer(rb, dtype=np.uint8)
    image_start[i] = s.image_start_pos
    variant_start[i] = s.variant_start
    for j, r in enumerate(s.reads):
      sp = s.sort_positions[j] if s.sort_positions else r.position
      key = (id(r), sp)
      k = read_index.get(key)
      if k is None:
        k = len(reads)
        read_index[key] = k
        reads.append(r)
        sort_pos.append(sp)
      pair_read.append(k)
      pair_support.append(s.support[j])
      pair_group.append(s.allele_group[j] if s.allele_group else 0)
    for k in pair_slots:
      if s.pair_channels is None or k not in s.pair_channels:
        raise ValueError(f'the channel list names a per-pair plane channel (slot {k}) but the ImageSpec carries no values for it')
      pair_planes[k].extend(s.pair_channels[k])
    pair_begin[i + 1] = len(pair_read)

  n_reads = len(reads)
  seq_begin = np.zeros(n_reads + 1, dtype=np.int64)
  cig_begin = np.zeros(n_reads + 1, dtype=np.int64)
  flags = np.zeros(n_reads, dtype=np.uint8)
  hp = np.zeros(n_reads, dtype=np.int32)
  for k, r in enumerate(reads):
    if len(r.aligned_quality) != len(r.aligned_sequence):
      raise ValueError('aligned_quality and aligned_sequence differ in length')
    seq_begin[k + 1] = seq_begin[k] + len(r.aligned_sequence)
    cig_begin[k + 1] = cig_begin[k] + len(r.cigar)
    f = 0
    if r.reverse_strand:
      f |= READ_REVERSE_STRAND
    if r.supplementary_alignment:
      f |= READ_SUPPLEMENTARY
    if r.hp_values:  # info contains "HP" with >= 1 value
      f |= READ_HAS_HP
      hp[k] = r.hp_values[0]
      if len(r.hp_values) > 1:
        f |= READ_HP_MULTI
    flags[k] = f
  bases = np.frombuffer(b''.join(bytes(r.aligned_sequence) for r in reads), dtype=np.uint8).copy() \
      if n_reads else np.zeros(0, dtype=np.uint8)
  quals = np.frombuffer(b''.join(bytes(r.aligned_quality) for r in reads), dtype=np.uint8).copy() \
      if n_reads else np.zeros(0, dtype=np.uint8)
  cigar = np.array([(ln << 4) | op for r in reads for op, ln in r.cigar], dtype=np.uint32)

  def _pad(a: np.ndarray) -> np.ndarray:
    # never hand the C side a NULL / zero-length allocation for an empty array
    return a if a.size else np.zeros(1, dtype=a.dtype)

  arrays = {
      'ref_bases': ref.reshape(-1),
      'image_start_pos': image_start,
      'variant_start': variant_start,
      'pair_begin': pair_begin,
      'pair_read': np.array(pair_read, dtype=np.int32),
      'pair_support': np.array(pair_support, dtype=np.uint8),
      'pair_allele_group': np.array(pair_group, dtype=np.uint8),
      'read_pos': np.array([r.position for r in reads], dtype=np.int32),
      'read_sort_pos': np.array(sort_pos, dtype=np.int32),
      'read_mapq': np.array([r.mapping_quality for r in reads], dtype=np.int32),
      'read_flags': flags,
      'read_fragment_length': np.array([r.fragment_length for r in reads], dtype=np.int32),
      'read_hp': hp,
      'read_name_rank': name_ranks(reads),
      'read_seq_begin': seq_begin,
      'read_cigar_begin': cig_begin,
      'bases': bases,
      'quals': quals,
      'cigar': cigar,
  }
  for k in pair_slots:
    arrays[f'pair_channel_{k}'] = np.array(pair_planes[k], dtype=np.uint8)
  if base_slots:
    from deepvariant_b200 import channels   # one native pass per read
    for k in base_slots:
      arrays[f'base_channel_{k}'] = np.concatenate([channels.base_plane(r, k) for r in reads]) if n_reads else np.zeros(0, dtype=np.uint8)
  arrays = {k: np.ascontiguousarray(_pad(v)) for k, v in arrays.items()}
  pb = PackedBatch(n_images=n_images, n_reads=n_reads, n_pairs=len(pair_read), ref_stride=ref_stride,
                   arrays=arrays)
  return pb


# ---------------------------------------------------------------------------------------------
# Table path: DvbBatch straight from the native BAM read table (no Read objects)
# ---------------------------------------------------------------------------------------------

@dataclasses.dataclass
class TableImageSpec:
  """One BuildPileupForOneSample call whose reads are rows of a bam.NativeBamTable."""
  ref_bases: str
  image_start_pos: int
  variant_start: int
  read_rows: np.ndarray            # int64 rows of the table, InMemoryReader::Query order (= file order)
  support: np.ndarray              # uint8 per read: 0 / 1 / 2
  allele_group: Optional[np.ndarray] = None


def table_support(dv_call: DeepVariantCall, key_to_local: Dict[str, List[int]], n_reads: int, alt_alleles: Sequence[str]) -> np.ndarray:
  """ReadSupportsAlt for every read of one image at once: walk the alts in order, first name match wins
  (channels/read_supports_variant_channel.cc:81-103).  key_to_local maps 'fragment_name/read_number' to the
  positions of the reads with that key in the image's read list."""
  out = np.zeros(n_reads, dtype=np.uint8)
  for alt_allele in dv_call.variant.alternate_bases:
    names = dv_call.allele_support.get(alt_allele)
    if not names:
      continue
    cls = 1 if alt_allele in alt_alleles else 2
    for name in names:
      for j in key_to_local.get(name, ()):   # a key can name several rows (same QNAME + read number aligned twice)
        if out[j] == 0:
          out[j] = cls
  return out


def table_allele_groups(dv_call: DeepVariantCall, key_to_local: Dict[str, List[int]], n_reads: int) -> np.ndarray:
  n_alt = len(dv_call.variant.alternate_bases)
  out = np.full(n_reads, n_alt, dtype=np.uint8)
  for i, alt in enumerate(dv_call.variant.alternate_bases):
    for name in dv_call.allele_support.get(alt, ()):
      for j in key_to_local.get(name, ()):
        out[j] = i
  return out


def _ragged_gather(values: np.ndarray, begin: np.ndarray, rows: np.ndarray):
  """Concatenation of values[begin[r]:begin[r+1]] for r in rows, and the new CSR offsets."""
  lens = (begin[rows + 1] - begin[rows]).astype(np.int64)
  new_begin = np.zeros(len(rows) + 1, dtype=np.int64)
  np.cumsum(lens, out=new_begin[1:])
  total = int(new_begin[-1])
  if total == 0:
    return np.zeros(0, dtype=values.dtype), new_begin
  idx = np.repeat(begin[rows] - new_begin[:-1], lens) + np.arange(total, dtype=np.int64)
  return values[idx], new_begin


def pack_images_from_table(specs: Sequence[TableImageSpec], table, params: _lib.DvbPileupParams) -> PackedBatch:
  """Same PackedBatch as pack_images() gives for the equivalent Read lists, gathered from the flat table arrays."""
  width = params.width
  ref_stride = (width + 15) // 16 * 16
  n_images = len(specs)
  ref = np.zeros((n_images, ref_stride), dtype=np.uint8)
  image_start = np.zeros(n_images, dtype=np.int32)
  variant_start = np.zeros(n_images, dtype=np.int32)
  pair_begin = np.zeros(n_images + 1, dtype=np.int64)
  for i, s in enumerate(specs):
    rb = s.ref_bases.encode() if isinstance(s.ref_bases, str) else bytes(s.ref_bases)
    if len(rb) != width:
      raise ValueError(f'ref_bases has {len(rb)} bases, expected width {width}')
    ref[i, :width] = np.frombuffer(rb, dtype=np.uint8)
    image_start[i] = s.image_start_pos
    variant_start[i] = s.variant_start
    pair_begin[i + 1] = pair_begin[i] + len(s.read_rows)
  all_rows = np.concatenate([s.read_rows for s in specs]).astype(np.int64) if n_images else np.zeros(0, dtype=np.int64)
  # reads are stored once, in order of first use (what pack_images does with its id() map)
  uniq, first = np.unique(all_rows, return_index=True)
  order = np.argsort(first, kind='stable')
  rows = uniq[order]
  remap = np.empty(len(uniq), dtype=np.int32)
  remap[order] = np.arange(len(uniq), dtype=np.int32)
  pair_read = remap[np.searchsorted(uniq, all_rows)] if len(all_rows) else np.zeros(0, dtype=np.int32)
  pair_support = np.concatenate([s.support for s in specs]).astype(np.uint8) if n_images else np.zeros(0, dtype=np.uint8)
  pair_group = np.concatenate([s.allele_group if s.allele_group is not None else np.zeros(len(s.read_rows), dtype=np.uint8)
                               for s in specs]).astype(np.uint8) if n_images else np.zeros(0, dtype=np.uint8)
  n_reads = len(rows)
  flag = table.flag[rows]
  hp_raw = table.hp[rows]
  has_hp = hp_raw != table.HP_ABSENT
  flags = (((flag & 0x10) != 0) * READ_REVERSE_STRAND + ((flag & 0x800) != 0) * READ_SUPPLEMENTARY + has_hp * READ_HAS_HP).astype(np.uint8)
  bases, seq_begin = _ragged_gather(table.bases, table.seq_begin, rows)
  quals, _ = _ragged_gather(table.quals, table.seq_begin, rows)
  cigar, cig_begin = _ragged_gather(table.cigar, table.cigar_begin, rows)
  # dense rank of (fragment_name bytes, read_number) among the reads of this batch
  keys = [(table.names[int(table.name_begin[r]):int(table.name_begin[r + 1])], int(table.read_number[r])) for r in rows]
  rank = {k: i for i, k in enumerate(sorted(set(keys)))}
  name_rank = np.array([rank[k] for k in keys], dtype=np.uint32)

  def _pad(a: np.ndarray) -> np.ndarray:
    return a if a.size else np.zeros(1, dtype=a.dtype)

  arrays = {
      'ref_bases': ref.reshape(-1), 'image_start_pos': image_start, 'variant_start': variant_start, 'pair_begin': pair_begin,
      'pair_read': pair_read.astype(np.int32), 'pair_support': pair_support, 'pair_allele_group': pair_group,
      'read_pos': table.pos[rows].astype(np.int32), 'read_sort_pos': table.pos[rows].astype(np.int32),
      'read_mapq': table.mapq[rows].astype(np.int32), 'read_flags': flags,
      'read_fragment_length': table.fragment_length[rows].astype(np.int32),
      'read_hp': np.where(has_hp, hp_raw, 0).astype(np.int32), 'read_name_rank': name_rank,
      'read_seq_begin': seq_begin, 'read_cigar_begin': cig_begin, 'bases': bases, 'quals': quals, 'cigar': cigar.astype(np.uint32),
  }
  arrays = {k: np.ascontiguousarray(_pad(v)) for k, v in arrays.items()}
  return PackedBatch(n_images=n_images, n_reads=n_reads, n_pairs=int(pair_begin[-1]), ref_stride=ref_stride, arrays=arrays)


def read_allele_key(ref: str, alt: str) -> Optional[Tuple[int, bytes]]:
  """The (AlleleType, bases) key AlleleCount.read_alleles holds for reads that support `alt` of a candidate whose reference
  allele is `ref` - the inverse of BuildAlleleMap (variant_calling_multisample.cc:560-606): a substitution's alt is its read
  base + ref[1:], an insertion's is anchor + inserted bases + ref[1:], a deletion's is anchor + ref[1 + deleted:].
  None when (ref, alt) is not of that form (not a candidate of the very-sensitive caller)."""
  if not ref or not alt:
    return None
  if len(alt) == len(ref):
    return (2, alt[:1].encode()) if alt[1:] == ref[1:] and alt[0] != ref[0] else None
  if len(alt) > len(ref):
    return (3, alt[:len(alt) - len(ref) + 1].encode()) if alt.endswith(ref[1:]) else None
  n = len(ref) - len(alt) + 1
  return (4, (alt[:1] + ref[1:n]).encode()) if alt[1:] == ref[n:] else None


def attach_alleles(packed: 'PackedBatch', images: Sequence['RegionImage'], min_mapping_quality: int, min_base_quality: int,
                   keep_legacy: bool, track_ref_reads: bool, with_groups: bool) -> 'PackedBatch':
  """Adds the allele keys of `images` (RegionImage.alleles / ref_run) to a packed batch: the encoder's pre-pass then derives
  pair_support / pair_allele_group on the device (DvbBatch.allele_begin, include/dvb.h) and the arrays of the same name are ignored."""
  n = len(images)
  assert n == packed.n_images
  counts = np.fromiter((len(im.alleles) for im in images), dtype=np.int64, count=n)
  begin = np.zeros(n + 1, dtype=np.int64)
  np.cumsum(counts, out=begin[1:])
  flat = [a for im in images for a in im.alleles]
  lens = np.fromiter((len(a[1]) for a in flat), dtype=np.int64, count=len(flat))
  bases_begin = np.zeros(len(flat) + 1, dtype=np.int64)
  np.cumsum(lens, out=bases_begin[1:])
  u8 = lambda k: np.ascontiguousarray(np.fromiter((a[k] for a in flat), dtype=np.uint8, count=len(flat))) if flat else np.zeros(1, np.uint8)
  a = packed.arrays
  a['allele_begin'] = begin
  a['allele_type'], a['allele_class'] = u8(0), u8(2)
  if with_groups:
    a['allele_group'] = u8(3)
    a['image_group_default'] = np.ascontiguousarray(np.fromiter((im.group_default for im in images), dtype=np.uint8, count=n)) if n else np.zeros(1, np.uint8)
  a['allele_bases_begin'] = bases_begin
  a['allele_bases'] = np.frombuffer(b''.join(x[1] for x in flat) or b'\0', dtype=np.uint8).copy()
  a['image_ref_run'] = np.ascontiguousarray(np.fromiter((im.ref_run for im in images), dtype=np.int32, count=n)) if n else np.zeros(1, np.int32)
  repeated = packed.n_reads > 0 and int(a['read_name_rank'][:packed.n_reads].max()) + 1 < packed.n_reads
  packed.support = (int(min_mapping_quality), int(min_base_quality),
                    (_lib.SUPPORT_KEEP_LEGACY if keep_legacy else 0) | (_lib.SUPPORT_TRACK_REF_READS if track_ref_reads else 0) |
                    (_lib.SUPPORT_REPEATED_KEYS if repeated else 0))
  return packed


@dataclasses.dataclass
class RegionImage:
  """One image of a region for pack_region_native: a candidate x alt combination, with allele_support flattened into
  (key, class, group) entries in alt order.  The key blob / lengths / groups are shared by the images of a candidate."""
  ref_id: int
  variant_start: int
  variant_end: int
  image_start_pos: int
  ref_bases: bytes
  keys_blob: bytes                 # the "fragment_name/read_number" keys of all alts, concatenated
  key_lens: np.ndarray             # int64 per entry
  support_class: np.ndarray        # uint8 per entry: 1 = alt of this image, 2 = other alt
  support_group: Optional[np.ndarray] = None   # uint8 per entry: alt index
  group_default: int = 0
  # device-side support (attach_alleles): (AlleleType, key bases, class, group) per alt, and the canonical reference run after variant_start
  alleles: Optional[List[Tuple[int, bytes, int, int]]] = None
  ref_run: int = 0


def pack_region_native(table, images: Sequence[RegionImage], region_ref_id: int, region_start: int, region_end: int,
                       read_overlap_buffer_bp: int, params: _lib.DvbPileupParams, with_groups: bool = False) -> PackedBatch:
  """pack_images_from_table() done by the C++ region packer (dvb_pack_region_from_bam, csrc/dvb_bam.cu): the read query,
  the read-name support search and the gathers run over the native table; Python only flattens the candidates."""
  lib = _lib.lib()
  width = params.width
  n = len(images)
  for im in images:
    if len(im.ref_bases) != width:
      raise ValueError(f'ref_bases has {len(im.ref_bases)} bases, expected width {width}')
  ref = np.frombuffer(b''.join(im.ref_bases for im in images) or bytes(width), dtype=np.uint8)
  meta = np.array([(im.ref_id, im.variant_start, im.variant_end, im.image_start_pos, len(im.key_lens), im.group_default)
                   for im in images], dtype=np.int64).reshape(n, 6)
  ref_id, vstart, vend, istart = (np.ascontiguousarray(meta[:, k], dtype=np.int32) if n else np.zeros(1, np.int32) for k in range(4))
  support_begin = np.zeros(n + 1, dtype=np.int64)
  np.cumsum(meta[:, 4], out=support_begin[1:])
  lens = np.concatenate([im.key_lens for im in images]) if n else np.zeros(0, dtype=np.int64)
  name_begin = np.zeros(len(lens) + 1, dtype=np.int64)
  np.cumsum(lens, out=name_begin[1:])
  names = b''.join(im.keys_blob for im in images) or b'\0'
  cat8 = lambda parts: np.ascontiguousarray(np.concatenate(parts + [np.zeros(1, np.uint8)]), dtype=np.uint8)
  classes = cat8([im.support_class for im in images])
  groups = cat8([im.support_group for im in images]) if with_groups else None
  gdef = np.ascontiguousarray(np.append(meta[:, 5], 0), dtype=np.uint8) if with_groups else None
  c = _lib.DvbRegionCandidates()
  c.n_images = n
  c.ref_id, c.variant_start, c.variant_end, c.image_start_pos = (a.ctypes.data for a in (ref_id, vstart, vend, istart))
  c.ref_bases, c.ref_stride = ref.ctypes.data, width
  c.support_begin, c.support_class = support_begin.ctypes.data, classes.ctypes.data
  c.support_group = groups.ctypes.data if with_groups else None
  c.support_name_begin = name_begin.ctypes.data
  c.support_names = C.cast(C.c_char_p(names), C.c_void_p)
  c.group_default = gdef.ctypes.data if with_groups else None
  h = C.c_void_p()
  _lib.check(lib.dvb_pack_region_from_bam(table.handle, C.byref(c), region_ref_id, region_start, region_end,
                                          read_overlap_buffer_bp, width, C.byref(h)))
  try:
    b = _lib.DvbBatch()
    _lib.check(lib.dvb_packed_region_batch(h, C.byref(b)))
    counts = {'ref_bases': b.n_images * b.ref_stride, 'image_start_pos': b.n_images, 'variant_start': b.n_images,
              'pair_begin': b.n_images + 1, 'pair_read': b.n_pairs, 'pair_support': b.n_pairs, 'pair_allele_group': b.n_pairs,
              'read_seq_begin': b.n_reads + 1, 'read_cigar_begin': b.n_reads + 1, 'bases': b.n_bases, 'quals': b.n_bases,
              'cigar': b.n_cigar}
    arrays = {}
    for name, dtype in _lib.BATCH_ARRAYS:
      cnt = int(counts.get(name, b.n_reads))
      ptr = getattr(b, name)
      if cnt and ptr:
        arrays[name] = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), shape=(cnt * np.dtype(dtype).itemsize,)).view(dtype).copy()
      else:
        arrays[name] = np.zeros(1, dtype=dtype)
    return PackedBatch(n_images=int(b.n_images), n_reads=int(b.n_reads), n_pairs=int(b.n_pairs), ref_stride=int(b.ref_stride),
                       arrays=arrays)
  finally:
    lib.dvb_packed_region_free(h)
