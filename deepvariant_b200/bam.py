"""Minimal BAM reader (BGZF + BAM records via zlib; no htslib) producing the Read fields the pileup
path uses, with the conversions of nucleus' SamReader (third_party/nucleus/io/sam_reader.cc:760-860):
  fragment_name = qname, fragment_length = isize, read_number = 0 if (FREAD1 or unpaired) else 1,
  number_reads = 2 if paired else 1, aligned_sequence upper-case from the 4-bit codes ("=ACMGRSVTWYHKDBN"),
  aligned_quality raw phred bytes, HP aux tag -> info['HP'] (with parse_sam_aux_fields).
and its ReadRequirements filter (sam_reader.cc:217-245, utils.cc:261-266).
This is SURVEY §8(f) "next" row #1 in its simplest form: whole-file decode, in-memory region query.
"""
from __future__ import annotations

import gzip
import os
import re
import struct
from typing import Dict, Iterable, List, Optional, Sequence

import numpy as np

from deepvariant_b200.protos import Read

_SEQ = '=ACMGRSVTWYHKDBN'
FPAIRED, FPROPER, FUNMAP, FMUNMAP, FREVERSE, FMREVERSE, FREAD1, FREAD2, FSECONDARY, FQCFAIL, FDUP, FSUPP = (
    0x1, 0x2, 0x4, 0x8, 0x10, 0x20, 0x40, 0x80, 0x100, 0x200, 0x400, 0x800)


class ReadRequirements:
  """reads.proto ReadRequirements defaults as make_examples sets them (make_examples_options.py:957-964)."""

  def __init__(self, min_mapping_quality=5, keep_duplicates=False, keep_failed_vendor_quality_checks=False,
               keep_secondary_alignments=False, keep_supplementary_alignments=False, keep_unaligned=False,
               keep_improperly_placed=False):
    self.__dict__.update(locals())


def _parse_aux_hp(aux: bytes) -> Optional[int]:
  i = 0
  sizes = {b'A': 1, b'c': 1, b'C': 1, b's': 2, b'S': 2, b'i': 4, b'I': 4, b'f': 4}
  fmts = {b'c': '<b', b'C': '<B', b's': '<h', b'S': '<H', b'i': '<i', b'I': '<I'}
  n = len(aux)
  while i + 3 <= n:
    tag, typ = aux[i:i + 2], aux[i + 2:i + 3]
    i += 3
    if typ in sizes:
      if tag == b'HP' and typ in fmts:
        return struct.unpack_from(fmts[typ], aux, i)[0]
      i += sizes[typ]
    elif typ in (b'Z', b'H'):
      j = aux.index(b'\0', i)
      i = j + 1
    elif typ == b'B':
      sub = aux[i:i + 1]
      cnt = struct.unpack_from('<i', aux, i + 1)[0]
      i += 5 + cnt * sizes[sub]
    else:
      break
  return None


_AUX_SIZES = {b'A': 1, b'c': 1, b'C': 1, b's': 2, b'S': 2, b'i': 4, b'I': 4, b'f': 4}
_AUX_FMTS = {b'c': 'b', b'C': 'B', b's': 'h', b'S': 'H', b'i': 'i', b'I': 'I', b'f': 'f'}


def parse_aux_tags(aux: bytes, wanted=(b'MM', b'ML', b'MN', b'tp', b't0', b'X5', b'X6')) -> dict:
  """The wanted tags of one record's aux bytes (BAM encoding): integers as int, Z as bytes, B arrays as lists."""
  out, i, n = {}, 0, len(aux)
  while i + 3 <= n:
    tag, typ = aux[i:i + 2], aux[i + 2:i + 3]
    i += 3
    if typ in _AUX_SIZES:
      if tag in wanted and typ in _AUX_FMTS:
        out[tag] = struct.unpack_from('<' + _AUX_FMTS[typ], aux, i)[0]
      i += _AUX_SIZES[typ]
    elif typ in (b'Z', b'H'):
      j = aux.index(b'\0', i)
      if tag in wanted:
        out[tag] = bytes(aux[i:j])
      i = j + 1
    elif typ == b'B':
      sub = aux[i:i + 1]
      cnt = struct.unpack_from('<i', aux, i + 1)[0]
      if tag in wanted:
        out[tag] = list(struct.unpack_from(f'<{cnt}{_AUX_FMTS[sub]}', aux, i + 5))
      i += 5 + cnt * _AUX_SIZES[sub]
    else:
      break
  return out


_COMPLEMENT = bytes.maketrans(b'ACGTNacgtn', b'TGCANtgcan')
_BASE_MODIFICATION = re.compile(r'([ACGTUN])([-+])([a-z]+|[0-9]+)([.?]?)')       # kBaseModificationRegexp, sam_reader.cc:518-519


def parse_base_modifications(aligned_sequence: bytes, reverse_strand: bool, mm: Optional[str], ml: Optional[Sequence[int]],
                             mn: Optional[int] = None) -> dict:
  """ParseBaseModifications (third_party/nucleus/io/sam_reader.cc:521-716): the MM / ML (/ MN) aux tags -> {'5mC' | '6mA': one
  probability byte per base of the aligned sequence}.  MM counts occurrences of the modified base along the ORIGINAL read (the
  reverse complement of a reverse-strand alignment), skipping mm_delta of them before each modified one; ML concatenates the
  probabilities of all modifications; a second entry for the same modification (6mA on A+ and T-) is merged by maximum; an MN that
  differs from the sequence length, or ML running out, voids everything (hard-clipped records)."""
  result: dict = {}
  if mm is None or ml is None:
    return result
  n = len(aligned_sequence)
  if (mn if mn is not None else n) != n:
    return result
  seq = bytes(aligned_sequence).translate(_COMPLEMENT)[::-1] if reverse_strand else bytes(aligned_sequence)
  ml_offset = 0
  body = mm[:-1] if mm.endswith(';') else mm
  for entry in body.split(';'):
    parts = entry.split(',')
    if len(parts) <= 1:
      continue
    m = _BASE_MODIFICATION.fullmatch(parts[0])
    if m is None:
      raise ValueError(f'bad MM entry {parts[0]!r}')      # the reference dereferences an empty capture here
    base, strand, modification = m.group(1), m.group(2), m.group(3)
    if base == 'C' and strand == '+' and modification == 'm':
      spec = '5mC'
    elif (base == 'A' and strand == '+' and modification == 'a') or (base == 'T' and strand == '-' and modification == 'a'):
      spec = '6mA'
    else:
      ml_offset += len(parts) - 1
      continue
    deltas = parts[1:]
    values = bytearray(n)
    base_count, k = 0, 0
    mm_delta = int(deltas[0])
    target = ord(base)
    for pos in range(n):
      if seq[pos] != target:
        continue
      if base_count != mm_delta:
        base_count += 1
        continue
      if ml_offset + k >= len(ml):
        return {}
      values[pos] = ml[k + ml_offset] & 0xFF
      base_count = 0
      k += 1
      if k >= len(deltas):
        ml_offset += k
        out = bytes(values[::-1]) if reverse_strand else bytes(values)
        if spec in result:
          # std::max over the chars of two std::strings (:697-701): a SIGNED comparison on the reference's platforms, so a probability
          # byte >= 128 loses against the other entry's 0 - kept as it is, the pixels must equal the reference's
          signed = lambda b: b - 256 if b >= 128 else b
          out = bytes(a if signed(a) >= signed(b) else b for a, b in zip(result[spec], out))
        result[spec] = out
        break
      mm_delta = int(deltas[k])
  return result


def apply_aux_tags(read: Read, aux: bytes) -> None:
  """Fills the per-base aux data of the optional channels (Read.base_modifications, tp_values, t0_value) from a record's aux bytes.
  X5 / X6 are this package's own scratch tags: the already-parsed 5mC / 6mA bytes of a read that went through a scratch BAM."""
  tags = parse_aux_tags(aux)
  mods = {}
  if b'MM' in tags and b'ML' in tags and isinstance(tags[b'ML'], list):
    mods = parse_base_modifications(read.aligned_sequence, read.reverse_strand, tags[b'MM'].decode(), tags[b'ML'], tags.get(b'MN'))
  for tag, key in ((b'X5', '5mC'), (b'X6', '6mA')):
    if isinstance(tags.get(tag), list):
      mods[key] = bytes(tags[tag])
  if mods:
    read.base_modifications = mods
  if isinstance(tags.get(b'tp'), list):
    read.tp_values = tags[b'tp']
  if isinstance(tags.get(b't0'), bytes):
    read.t0_value = tags[b't0']


class BamReader:

  def __init__(self, path: str, read_requirements: Optional[ReadRequirements] = None, parse_aux: bool = False):
    self.requirements = read_requirements or ReadRequirements()
    with gzip.open(path, 'rb') as f:   # BGZF is a series of gzip members
      data = f.read()
    if data[:4] != b'BAM\1':
      raise ValueError(f'{path} is not a BAM file')
    l_text = struct.unpack_from('<i', data, 4)[0]
    pos = 8 + l_text
    n_ref = struct.unpack_from('<i', data, pos)[0]
    pos += 4
    self.references: List[str] = []
    for _ in range(n_ref):
      l_name = struct.unpack_from('<i', data, pos)[0]
      self.references.append(data[pos + 4:pos + 4 + l_name - 1].decode())
      pos += 4 + l_name + 4
    self.reads: List[Read] = []
    while pos + 4 <= len(data):
      block_size = struct.unpack_from('<i', data, pos)[0]
      rec = data[pos + 4:pos + 4 + block_size]
      pos += 4 + block_size
      r = self._convert(rec, parse_aux)
      if r is not None:
        self.reads.append(r)
    self._by_contig: Dict[str, List[Read]] = {}
    for r in self.reads:
      self._by_contig.setdefault(r.reference_name, []).append(r)

  def _convert(self, rec: bytes, parse_aux: bool) -> Optional[Read]:
    ref_id, pos_, l_read_name, mapq, _bin, n_cigar, flag, l_seq, next_ref, next_pos, tlen = struct.unpack_from('<iiBBHHHiiii', rec, 0)
    req = self.requirements
    if ((flag & FDUP and not req.keep_duplicates) or (flag & FQCFAIL and not req.keep_failed_vendor_quality_checks) or
        (flag & FSECONDARY and not req.keep_secondary_alignments) or (flag & FSUPP and not req.keep_supplementary_alignments)):
      return None
    mapped = not flag & FUNMAP
    if not mapped and not req.keep_unaligned:
      return None
    paired = bool(flag & FPAIRED)
    off = 32
    name = rec[off:off + l_read_name - 1].decode()
    off += l_read_name
    cigar = []
    for k in range(n_cigar):
      v = struct.unpack_from('<I', rec, off + 4 * k)[0]
      cigar.append((v & 0xF, v >> 4))
    off += 4 * n_cigar
    seq_bytes = rec[off:off + (l_seq + 1) // 2]
    off += (l_seq + 1) // 2
    seq = ''.join(_SEQ[b >> 4] + _SEQ[b & 0xF] for b in seq_bytes)[:l_seq]
    qual = rec[off:off + l_seq]
    off += l_seq
    contig = self.references[ref_id] if ref_id >= 0 else ''
    mate_contig = self.references[next_ref] if (paired and not flag & FMUNMAP and next_ref >= 0) else ''
    number_reads = 2 if paired else 1
    proper = bool(flag & FPROPER)
    # IsReadProperlyPlaced (utils.cc:261-266)
    if not req.keep_improperly_placed and mapped:
      if not (number_reads < 2 or proper or not mate_contig or contig == mate_contig):
        return None
    if mapped and mapq < req.min_mapping_quality:
      return None
    r = Read(fragment_name=name, read_number=0 if (flag & FREAD1 or not paired) else 1, reference_name=contig,
             position=pos_, reverse_strand=bool(flag & FREVERSE), mapping_quality=mapq, cigar=cigar if mapped else [],
             aligned_sequence=seq.encode(), aligned_quality=bytes(qual), fragment_length=tlen,
             supplementary_alignment=bool(flag & FSUPP), secondary_alignment=bool(flag & FSECONDARY),
             duplicate_fragment=bool(flag & FDUP), failed_vendor_quality_checks=bool(flag & FQCFAIL),
             proper_placement=proper, number_reads=number_reads)
    if parse_aux:
      apply_aux_tags(r, rec[off:])
      hp = _parse_aux_hp(rec[off:])
      if hp is not None:
        r.hp_values = [hp]
    return r

  def query(self, contig: str, start: int, end: int) -> List[Read]:
    """Reads overlapping [start, end) in file order (ReadOverlapsRegion, utils.cc:172-188)."""
    return [r for r in self._by_contig.get(contig, ()) if end > r.position and start < r.end()]


# ---------------------------------------------------------------------------------------------
# Native decode (libdvb.so dvb_bam_*, csrc/dvb_bam.cu): BAM -> Structure-of-Arrays read table
# ---------------------------------------------------------------------------------------------

class NativeBamTable:
  """Flat numpy arrays of every read that passes the ReadRequirements filter, decoded by the C++ block-parallel
  BGZF/BAM decoder (include/dvb.h DvbReadTable).  `reads()` materialises the same `Read` objects `BamReader`
  produces (used by the planner and by the parity tests); `query()` answers region queries on the arrays."""

  def __init__(self, path: str, read_requirements: Optional[ReadRequirements] = None, parse_aux: bool = False,
               threads: int = 0, regions=None, ref_reader=None):
    """regions: optional [(contig, start, end), ...] (0-based, half-open) - only reads overlapping one of them are decoded
    (dvb_bam_open_regions: the .bai linear index is used when they lie on one contig and the index is beside the file)."""
    import ctypes as C
    import numpy as np
    from deepvariant_b200 import _lib
    lib = _lib.lib()
    req = read_requirements or ReadRequirements()
    creq = _lib.DvbReadRequirements(
        int(req.min_mapping_quality), int(req.keep_duplicates), int(req.keep_failed_vendor_quality_checks),
        int(req.keep_secondary_alignments), int(req.keep_supplementary_alignments), int(req.keep_unaligned),
        int(req.keep_improperly_placed))
    h = C.c_void_p()
    scratch = None
    if is_cram(path):
      # CRAM: the wanted containers become an uncompressed BAM in a scratch directory, which the decoder below reads (sam_reader.cc
      # hands CRAM to htslib with the FASTA; here that FASTA is `ref_reader`)
      import tempfile
      if ref_reader is None:
        raise ValueError(f'{path} is a CRAM file: the reference FASTA is needed to decode it (ref_reader=)')
      scratch = tempfile.TemporaryDirectory()
      cram_to_bam(path, os.path.join(scratch.name, 'reads.bam'), ref_reader, regions)
      path = os.path.join(scratch.name, 'reads.bam')
    if regions:
      names = (C.c_char_p * len(regions))(*[r[0].encode() for r in regions])
      starts = np.array([r[1] for r in regions], dtype=np.int64)
      ends = np.array([min(int(r[2]), (1 << 62)) for r in regions], dtype=np.int64)
      _lib.check(lib.dvb_bam_open_regions(path.encode(), C.byref(creq), 3 if parse_aux else 0, threads, names, starts.ctypes.data_as(C.c_void_p),
                                          ends.ctypes.data_as(C.c_void_p), len(regions), C.byref(h)))
    else:
      _lib.check(lib.dvb_bam_open(path.encode(), C.byref(creq), 3 if parse_aux else 0, threads, C.byref(h)))   # HP + the raw aux bytes
    if scratch is not None:
      scratch.cleanup()
    self._load(h, parse_aux)

  def _load(self, h, parse_aux: bool) -> None:
    """The arrays of an open DvbBam handle (owned by this object from here on)."""
    import ctypes as C
    import numpy as np
    from deepvariant_b200 import _lib
    lib = _lib.lib()
    try:
      t = _lib.DvbReadTable()
      _lib.check(lib.dvb_bam_table(h, C.byref(t)))
      n = t.n_reads

      def arr(ptr, count, dtype):
        if not count:
          return np.zeros(0, dtype=dtype)
        buf = (C.c_char * (count * np.dtype(dtype).itemsize)).from_address(ptr)
        return np.frombuffer(buf, dtype=dtype, count=count).copy()   # own the data: valid after close()

      self.n_reads = n
      self.n_records_seen = int(t.n_records_seen)
      self.references = [lib.dvb_bam_ref_name(h, i).decode() for i in range(t.n_refs)]
      self.reference_lengths = [int(lib.dvb_bam_ref_length(h, i)) for i in range(t.n_refs)]
      self.ref_id = arr(t.ref_id, n, np.int32)
      self.pos = arr(t.pos, n, np.int32)
      self.end = arr(t.end, n, np.int32)
      self.mapq = arr(t.mapq, n, np.uint8)
      self.flag = arr(t.flag, n, np.uint16)
      self.fragment_length = arr(t.fragment_length, n, np.int32)
      self.hp = arr(t.hp, n, np.int32)
      self.read_number = arr(t.read_number, n, np.uint8)
      self.number_reads = arr(t.number_reads, n, np.uint8)
      self.seq_begin = arr(t.seq_begin, n + 1, np.int64)
      self.cigar_begin = arr(t.cigar_begin, n + 1, np.int64)
      self.name_begin = arr(t.name_begin, n + 1, np.int64)
      self.bases = arr(t.bases, t.n_bases, np.uint8)
      self.quals = arr(t.quals, t.n_bases, np.uint8)
      self.cigar = arr(t.cigar, t.n_cigar, np.uint32)
      self.names = arr(t.names, t.n_name_bytes, np.uint8).tobytes()
      self.aux_begin = arr(t.aux_begin, n + 1, np.int64) if t.n_aux_bytes else None
      self.aux = arr(t.aux, t.n_aux_bytes, np.uint8).tobytes() if t.n_aux_bytes else b''
    except BaseException:
      lib.dvb_bam_close(h)
      raise
    # region queries: a coordinate-sorted table answers them with two binary searches (rows with pos < end form a prefix; the running
    # maximum of `end` is monotone, so rows that can reach `start` form a suffix of it) instead of a scan of every read per partition
    key = self.ref_id.astype(np.int64) * (1 << 32) + self.pos.astype(np.int64)
    mapped = self.ref_id >= 0
    self._sorted = bool(np.all(np.diff(key[mapped]) >= 0)) and (not mapped.any() or not (~mapped)[:int(np.nonzero(mapped)[0][-1]) + 1].any())
    self._key = key if self._sorted else None
    self._cummax_end = None
    if self._sorted and n:
      # per contig running maximum: encode (ref_id, end) so that the maximum restarts with every contig
      self._cummax_end = np.maximum.accumulate(self.ref_id.astype(np.int64) * (1 << 32) + self.end.astype(np.int64))
    self._handle = h            # kept open: the native region packer (packing.pack_region_native) reads the C++ table
    self._close = lib.dvb_bam_close
    self.parse_aux = parse_aux
    self._reads: Optional[List[Read]] = None

  @classmethod
  def derived(cls, source: 'NativeBamTable', rows, alignments=None) -> 'NativeBamTable':
    """The reads `rows` of `source` (in that order) as a table of their own; alignments[i] = (position, [(op, length), ...]) replaces
    the alignment of rows[i], None keeps it (dvb_bam_derive): realigned / normalised reads without a BAM file in between."""
    import ctypes as C
    import numpy as np
    from deepvariant_b200 import _lib
    lib = _lib.lib()
    rows = np.ascontiguousarray(rows, dtype=np.int64)
    h = C.c_void_p()
    if alignments is None:
      _lib.check(lib.dvb_bam_derive(source.handle, rows.ctypes.data_as(C.c_void_p), len(rows), None, None, None, C.byref(h)))
    else:
      if len(alignments) != len(rows):
        raise ValueError('one alignment (or None) per row')
      pos = np.zeros(len(rows), dtype=np.int32)
      begin = np.zeros(len(rows) + 1, dtype=np.int64)
      ops: list = []
      for i, a in enumerate(alignments):
        if a is not None:
          pos[i] = a[0]
          ops.extend((int(n) << 4) | int(op) for op, n in a[1])
        begin[i + 1] = len(ops)
      cig = np.asarray(ops, dtype=np.uint32)
      _lib.check(lib.dvb_bam_derive(source.handle, rows.ctypes.data_as(C.c_void_p), len(rows), pos.ctypes.data_as(C.c_void_p),
                                    begin.ctypes.data_as(C.c_void_p), cig.ctypes.data_as(C.c_void_p), C.byref(h)))
    self = cls.__new__(cls)
    self._load(h, source.parse_aux)
    return self

  def close(self) -> None:
    if getattr(self, '_handle', None) is not None:
      self._close(self._handle)
      self._handle = None

  def __del__(self):
    self.close()

  @property
  def handle(self):
    if self._handle is None:
      raise ValueError('NativeBamTable is closed')
    return self._handle

  HP_ABSENT = -(1 << 31)

  def read(self, i: int) -> Read:
    fl = int(self.flag[i])
    mapped = not fl & FUNMAP
    c0, c1 = int(self.cigar_begin[i]), int(self.cigar_begin[i + 1])
    s0, s1 = int(self.seq_begin[i]), int(self.seq_begin[i + 1])
    rid = int(self.ref_id[i])
    r = Read(fragment_name=self.names[int(self.name_begin[i]):int(self.name_begin[i + 1])].decode(),
             read_number=int(self.read_number[i]), reference_name=self.references[rid] if rid >= 0 else '',
             position=int(self.pos[i]), reverse_strand=bool(fl & FREVERSE), mapping_quality=int(self.mapq[i]),
             cigar=[(int(v) & 0xF, int(v) >> 4) for v in self.cigar[c0:c1]] if mapped else [],
             aligned_sequence=self.bases[s0:s1].tobytes(), aligned_quality=self.quals[s0:s1].tobytes(),
             fragment_length=int(self.fragment_length[i]), supplementary_alignment=bool(fl & FSUPP),
             secondary_alignment=bool(fl & FSECONDARY), duplicate_fragment=bool(fl & FDUP),
             failed_vendor_quality_checks=bool(fl & FQCFAIL), proper_placement=bool(fl & FPROPER),
             number_reads=int(self.number_reads[i]))
    if self.parse_aux and int(self.hp[i]) != self.HP_ABSENT:
      r.hp_values = [int(self.hp[i])]
    if self.parse_aux and self.aux_begin is not None and self.aux_begin[i + 1] > self.aux_begin[i]:
      apply_aux_tags(r, self.aux[int(self.aux_begin[i]):int(self.aux_begin[i + 1])])   # MM / ML / MN, tp, t0 (the optional channels' per-base data)
    r._table, r._row = self, i        # where the record lives natively (scratch_table derives from it; copy.copy keeps the tags)
    return r

  def reads(self) -> List[Read]:
    if self._reads is None:
      self._reads = [self.read(i) for i in range(self.n_reads)]
    return self._reads

  def query_indices(self, contig: str, start: int, end: int):
    """Indices (file order) of reads overlapping [start, end): ReadOverlapsRegion (utils.cc:172-188)."""
    import numpy as np
    if contig not in self.references:
      return np.zeros(0, dtype=np.int64)
    rid = self.references.index(contig)
    if self._sorted and self.n_reads:
      hi = int(np.searchsorted(self._key, rid * (1 << 32) + end, side='left'))            # first row with pos >= end (or a later contig)
      lo = int(np.searchsorted(self._cummax_end, rid * (1 << 32) + start, side='right'))  # first row whose running max(end) exceeds start
      if lo >= hi:
        return np.zeros(0, dtype=np.int64)
      sel = np.nonzero((self.end[lo:hi] > start) & (self.ref_id[lo:hi] == rid))[0]
      return sel + lo
    return np.nonzero((self.ref_id == rid) & (self.pos < end) & (self.end > start))[0]

  def query(self, contig: str, start: int, end: int) -> List[Read]:
    rs = self.reads()
    return [rs[i] for i in self.query_indices(contig, start, end)]


def is_cram(path: str) -> bool:
  try:
    with open(path, 'rb') as f:
      return f.read(4) == b'CRAM'
  except OSError:
    return False          # the native open reports the missing / unreadable file


def sam_header_text(path: str) -> str:
  """The SAM header of a BAM or CRAM file (CRAM: the block of the first container, raw or gzip as htslib / htsjdk write it)."""
  import gzip
  import zlib
  if not is_cram(path):
    with gzip.open(path, 'rb') as f:
      head = f.read(8)
      if head[:4] != b'BAM\1':
        raise ValueError(f'{path} is not a BAM file')
      return f.read(struct.unpack('<i', head[4:])[0]).decode(errors='replace')

  def itf8(b, o):
    v = b[o]
    if v < 0x80:
      return v, o + 1
    if v < 0xC0:
      return ((v & 0x3f) << 8) | b[o + 1], o + 2
    if v < 0xE0:
      return ((v & 0x1f) << 16) | (b[o + 1] << 8) | b[o + 2], o + 3
    if v < 0xF0:
      return ((v & 0x0f) << 24) | (b[o + 1] << 16) | (b[o + 2] << 8) | b[o + 3], o + 4
    return ((v & 0x0f) << 28) | (b[o + 1] << 20) | (b[o + 2] << 12) | (b[o + 3] << 4) | (b[o + 4] & 0x0f), o + 5

  with open(path, 'rb') as f:
    f.seek(26)
    length = struct.unpack('<i', f.read(4))[0]
    b = f.read(length + 256)
  o = 0
  for _ in range(4):                      # ref id, start, span, number of records
    _, o = itf8(b, o)
  for _ in range(2):                      # record counter, bases (LTF8)
    n = 0
    while n < 8 and b[o] & (0x80 >> n):
      n += 1
    o += 1 + n
  _, o = itf8(b, o)                       # number of blocks
  n_land, o = itf8(b, o)
  for _ in range(n_land):
    _, o = itf8(b, o)
  o += 4                                  # CRC32
  method = b[o]
  o += 2
  _, o = itf8(b, o)
  csz, o = itf8(b, o)
  _, o = itf8(b, o)
  data = b[o:o + csz]
  if method == 1:
    data = zlib.decompress(data, 31)
  elif method != 0:
    raise ValueError(f'{path}: header block compression method {method}')
  return data[4:4 + struct.unpack('<i', data[:4])[0]].decode(errors='replace')


def cram_to_bam(cram_path: str, bam_path: str, ref_reader, regions=None) -> int:
  """CRAM 3.0 -> uncompressed BAM (dvb_cram_to_bam, csrc/dvb_cram.cu); `ref_reader` supplies the contigs the file was compressed
  against (fasta.IndexedFastaReader: `contig_order`, `_contig(name)` = the whole sequence).  regions: [(contig, start, end)] - only the
  containers overlapping them are decoded, and only their contigs are handed over.  Returns the number of records written."""
  import ctypes as C
  import numpy as np
  from deepvariant_b200 import _lib
  wanted = None if not regions else {r[0] for r in regions}
  names = [c for c in (ref_reader.contig_order if ref_reader is not None else []) if wanted is None or c in wanted]
  seqs = [ref_reader._contig(c) for c in names]   # pylint: disable=protected-access
  name_arr = (C.c_char_p * max(1, len(names)))(*[n.encode() for n in names])
  seq_arr = (C.c_char_p * max(1, len(names)))(*seqs)
  lens = np.asarray([len(s) for s in seqs], dtype=np.int64)
  n = C.c_int64(0)
  if regions:
    rc = (C.c_char_p * len(regions))(*[r[0].encode() for r in regions])
    rs = np.asarray([r[1] for r in regions], dtype=np.int64)
    re_ = np.asarray([min(int(r[2]), 1 << 62) for r in regions], dtype=np.int64)
    _lib.check(_lib.lib().dvb_cram_to_bam(cram_path.encode(), bam_path.encode(), name_arr, seq_arr, lens.ctypes.data_as(C.c_void_p), len(names),
                                          rc, rs.ctypes.data_as(C.c_void_p), re_.ctypes.data_as(C.c_void_p), len(regions), C.byref(n)))
  else:
    _lib.check(_lib.lib().dvb_cram_to_bam(cram_path.encode(), bam_path.encode(), name_arr, seq_arr, lens.ctypes.data_as(C.c_void_p), len(names),
                                          None, None, None, 0, C.byref(n)))
  return int(n.value)


# ---- writer: reads that exist only in memory (realigned reads) become a BAM so that the native table / packer can take them ---------
_SEQ_CODE = {c: i for i, c in enumerate('=ACMGRSVTWYHKDBN')}


def _bgzf(payload: bytes, block: int = 0xff00, level: int = 1) -> bytes:
  """BGZF blocks of `block` payload bytes; level 0 = stored deflate blocks (5 bytes of framing per block, still under the 64 KiB limit)."""
  import zlib
  out = bytearray()
  for i in list(range(0, len(payload), block)) + [None]:
    ch = b'' if i is None else payload[i:i + block]
    co = zlib.compressobj(level, zlib.DEFLATED, -15)
    body = co.compress(ch) + co.flush()
    out += struct.pack('<4BI2BH2BHH', 0x1f, 0x8b, 8, 4, 0, 0, 0xff, 6, ord('B'), ord('C'), 2, len(body) + 25)
    out += body + struct.pack('<II', zlib.crc32(ch) & 0xffffffff, len(ch))
  return bytes(out)


_SEQ_CODE_LUT = np.full(256, 15, dtype=np.uint8)
for _ch, _code in _SEQ_CODE.items():
  _SEQ_CODE_LUT[ord(_ch)] = _code


def write_bam(path: str, reads, references, sample_name: str = '', level: int = 1) -> None:
  """references = [(name, length)].  Flags are rebuilt from the Read fields the reader fills (same decisions under
  ReadRequirements); mate fields are written so that IsReadProperlyPlaced still passes; HP is kept as an aux tag."""
  ref_index = {name: i for i, (name, _) in enumerate(references)}
  text = ('@HD\tVN:1.6\n' + (f'@RG\tID:rg\tSM:{sample_name}\n' if sample_name else '')).encode()
  out = bytearray(b'BAM\1' + struct.pack('<i', len(text)) + text + struct.pack('<i', len(references)))
  for name, length in references:
    out += struct.pack('<i', len(name) + 1) + name.encode() + b'\0' + struct.pack('<i', length)
  for r in reads:
    paired = r.number_reads == 2
    flag = ((FPAIRED | (FREAD1 if r.read_number == 0 else 0x80)) if paired else 0) | (FPROPER if r.proper_placement else 0) | \
        (FREVERSE if r.reverse_strand else 0) | (FSECONDARY if r.secondary_alignment else 0) | (FQCFAIL if r.failed_vendor_quality_checks else 0) | \
        (FDUP if r.duplicate_fragment else 0) | (FSUPP if r.supplementary_alignment else 0)
    rid = ref_index.get(r.reference_name, -1)
    seq = r.aligned_sequence
    codes = _SEQ_CODE_LUT[np.frombuffer(bytes(seq), dtype=np.uint8)]
    if len(codes) & 1:
      codes = np.append(codes, np.uint8(0))
    packed = ((codes[0::2] << 4) | codes[1::2]).astype(np.uint8).tobytes()
    aux = b''
    if r.hp_values:
      aux = b'HPi' + struct.pack('<i', int(r.hp_values[0]))
    for tag, key in ((b'X5', '5mC'), (b'X6', '6mA')):          # parsed base modifications travel as this package's own byte-array tags
      v = (r.base_modifications or {}).get(key)
      if v:
        aux += tag + b'BC' + struct.pack('<i', len(v)) + bytes(v)
    if r.tp_values:
      aux += b'tpBc' + struct.pack('<i', len(r.tp_values)) + struct.pack(f'<{len(r.tp_values)}b', *[max(-128, min(127, int(x))) for x in r.tp_values])
    if r.t0_value:
      aux += b't0Z' + bytes(r.t0_value) + b'\0'
    body = struct.pack('<iiBBHHHiiii', rid, r.position, len(r.fragment_name) + 1, r.mapping_quality, 0, len(r.cigar), flag, len(seq),
                       rid if paired else -1, r.position if paired else -1, r.fragment_length)
    body += r.fragment_name.encode() + b'\0' + b''.join(struct.pack('<I', (ln << 4) | op) for op, ln in r.cigar) + bytes(packed) + \
        bytes(r.aligned_quality) + aux
    out += struct.pack('<i', len(body)) + body
  with open(path, 'wb') as f:
    f.write(_bgzf(bytes(out), level=level))


def scratch_table(reads, references, read_requirements: Optional[ReadRequirements] = None, parse_aux: bool = False) -> 'NativeBamTable':
  """Read objects -> NativeBamTable: how realigned / normalised reads reach the native candidate generator and the region packer
  (in_memory_sam_reader.replace_reads, make_examples_core.py:2290-2300).  Reads that came out of one open NativeBamTable (read() tags
  them; the realigner and the normaliser hand on shallow copies with a new position / cigar) become a table derived natively from
  that one - only position and cigar are taken from the objects (dvb_bam_derive).  Anything else goes through a temporary BAM."""
  source = getattr(reads[0], '_table', None) if len(reads) else None
  if source is not None and getattr(source, '_handle', None) is not None and all(getattr(r, '_table', None) is source for r in reads):
    return NativeBamTable.derived(source, [r._row for r in reads], [(r.position, r.cigar) if r.cigar else None for r in reads])
  import tempfile
  with tempfile.TemporaryDirectory() as tmp:
    path = os.path.join(tmp, 'reads.bam')
    write_bam(path, reads, references, level=0)       # a scratch file read back at once: stored blocks, no deflate work
    return NativeBamTable(path, read_requirements, parse_aux=parse_aux)
