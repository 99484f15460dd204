"""Plain-Python mirrors of the reference protos the hot path touches, plus a minimal
protobuf wire codec (no generated code, no TensorFlow, no protoc at run time).

Only the fields the path reads or writes are modelled; field numbers are the wire
contract and are cited from the reference's .proto files:

  Read, LinearAlignment      third_party/nucleus/protos/reads.proto:40-237
  Position                   third_party/nucleus/protos/position.proto:38-47
  CigarUnit                  third_party/nucleus/protos/cigar.proto:34-93
  Value / ListValue          third_party/nucleus/protos/struct.proto:53-93
  Variant, VariantCall       third_party/nucleus/protos/variants.proto
  DeepVariantCall            deepvariant/protos/deepvariant.proto:262-317
  CallVariantsOutput         deepvariant/protos/deepvariant.proto:363-401
  tf.Example / Features      tensorflow/core/example/{example,feature}.proto (public TF wire format)

Unknown fields are preserved verbatim where a message is re-serialised (Variant), so a
`variant/encoded` feature round-trips byte-for-byte.
"""
from __future__ import annotations

import dataclasses
import struct
from typing import Dict, Iterable, Iterator, List, Optional, Tuple

# ----------------------------------------------------------------------------
# wire primitives
# ----------------------------------------------------------------------------

WT_VARINT, WT_I64, WT_LEN, WT_I32 = 0, 1, 2, 5


def _enc_varint(v: int) -> bytes:
  if v < 0:
    v += 1 << 64
  out = bytearray()
  while True:
    b = v & 0x7F
    v >>= 7
    if v:
      out.append(b | 0x80)
    else:
      out.append(b)
      return bytes(out)


def _dec_varint(buf: bytes, pos: int) -> Tuple[int, int]:
  result = 0
  shift = 0
  while True:
    b = buf[pos]
    pos += 1
    result |= (b & 0x7F) << shift
    if not b & 0x80:
      return result, pos
    shift += 7
    if shift > 70:
      raise ValueError('malformed varint')


def _to_signed64(v: int) -> int:
  return v - (1 << 64) if v >= (1 << 63) else v


def _to_signed32(v: int) -> int:
  v &= 0xFFFFFFFFFFFFFFFF
  v = _to_signed64(v)
  return v


def iter_fields(buf: bytes) -> Iterator[Tuple[int, int, object, bytes]]:
  """Yields (field_number, wire_type, value, raw_bytes_of_the_whole_field)."""
  pos = 0
  n = len(buf)
  while pos < n:
    start = pos
    key, pos = _dec_varint(buf, pos)
    fn, wt = key >> 3, key & 7
    if wt == WT_VARINT:
      val, pos = _dec_varint(buf, pos)
    elif wt == WT_I64:
      val = buf[pos:pos + 8]
      pos += 8
    elif wt == WT_LEN:
      ln, pos = _dec_varint(buf, pos)
      val = buf[pos:pos + ln]
      pos += ln
    elif wt == WT_I32:
      val = buf[pos:pos + 4]
      pos += 4
    else:
      raise ValueError(f'unsupported wire type {wt}')
    if pos > n:
      raise ValueError('truncated message')
    yield fn, wt, val, buf[start:pos]


def f_varint(fn: int, v: int) -> bytes:
  return _enc_varint((fn << 3) | WT_VARINT) + _enc_varint(v)


def f_bytes(fn: int, v: bytes) -> bytes:
  return _enc_varint((fn << 3) | WT_LEN) + _enc_varint(len(v)) + v


def f_double(fn: int, v: float) -> bytes:
  return _enc_varint((fn << 3) | WT_I64) + struct.pack('<d', v)


def f_float(fn: int, v: float) -> bytes:
  return _enc_varint((fn << 3) | WT_I32) + struct.pack('<f', v)


def packed_varints(vals: Iterable[int]) -> bytes:
  return b''.join(_enc_varint(v) for v in vals)


def unpack_varints(buf: bytes) -> List[int]:
  out = []
  pos = 0
  while pos < len(buf):
    v, pos = _dec_varint(buf, pos)
    out.append(v)
  return out


# ----------------------------------------------------------------------------
# nucleus messages
# ----------------------------------------------------------------------------

# CigarUnit.Operation (cigar.proto:38-82) -> BAM op code used in the packed batch.
CIGAR_ENUM_TO_BAM = {1: 0, 2: 1, 3: 2, 4: 3, 5: 4, 6: 5, 7: 6, 8: 7, 9: 8}
CIGAR_CHAR_TO_BAM = {c: i for i, c in enumerate('MIDNSHP=X')}
BAM_TO_CIGAR_CHAR = 'MIDNSHP=X'


@dataclasses.dataclass
class Read:
  """nucleus.genomics.v1.Read — the subset the pileup path uses."""
  fragment_name: str = ''
  read_number: int = 0
  reference_name: str = ''
  position: int = 0
  reverse_strand: bool = False
  mapping_quality: int = 0
  cigar: List[Tuple[int, int]] = dataclasses.field(default_factory=list)  # (bam_op, length)
  aligned_sequence: bytes = b''
  aligned_quality: bytes = b''
  fragment_length: int = 0
  supplementary_alignment: bool = False
  secondary_alignment: bool = False
  duplicate_fragment: bool = False
  failed_vendor_quality_checks: bool = False
  proper_placement: bool = False
  number_reads: int = 0
  hp_values: Optional[List[int]] = None  # info['HP'] int values (None = tag absent)
  # per-base aux data of the optional channels (deepvariant_b200/channels.py); None / empty = absent
  base_modifications: Optional[Dict[str, bytes]] = None   # Read.base_modifications: '5mC' / '6mA' -> one ML byte per base
  tp_values: Optional[List[int]] = None                   # info['tp'] int values (Ultima)
  t0_value: Optional[bytes] = None                        # info['t0'] string value (phred + 33 text)

  def key(self) -> str:
    # read_supports_variant_channel.cc:78-79
    return f'{self.fragment_name}/{self.read_number}'

  def end(self) -> int:
    """nucleus ReadEnd (third_party/nucleus/util/utils.cc:222-240)."""
    pos = self.position
    for op, ln in self.cigar:
      if op in (0, 7, 2, 3, 8):
        pos += ln
    return pos


def parse_cigar_string(cigar: str) -> List[Tuple[int, int]]:
  out = []
  num = ''
  for ch in cigar:
    if ch.isdigit():
      num += ch
    else:
      out.append((CIGAR_CHAR_TO_BAM[ch], int(num)))
      num = ''
  return out


def _parse_position(buf: bytes) -> Tuple[str, int, bool]:
  name, pos, rev = '', 0, False
  for fn, wt, val, _ in iter_fields(buf):
    if fn == 1:
      name = bytes(val).decode()
    elif fn == 2:
      pos = _to_signed64(val)
    elif fn == 3:
      rev = bool(val)
  return name, pos, rev


def _parse_list_value_ints(buf: bytes) -> List[int]:
  """ListValue{repeated Value values = 1}; Value.int_value = 7 (struct.proto:63,93).
  Non-int kinds read as 0, like Value::int_value() on another oneof case."""
  out = []
  for fn, wt, val, _ in iter_fields(buf):
    if fn == 1:
      iv = 0
      for fn2, wt2, val2, _ in iter_fields(bytes(val)):
        if fn2 == 7:
          iv = _to_signed32(val2)
      out.append(iv)
  return out


def parse_read(buf: bytes) -> Read:
  r = Read()
  for fn, wt, val, _ in iter_fields(buf):
    if fn == 4:
      r.fragment_name = bytes(val).decode()
    elif fn == 5:
      r.proper_placement = bool(val)
    elif fn == 6:
      r.duplicate_fragment = bool(val)
    elif fn == 7:
      r.fragment_length = _to_signed32(val)
    elif fn == 8:
      r.read_number = _to_signed32(val)
    elif fn == 9:
      r.number_reads = _to_signed32(val)
    elif fn == 10:
      r.failed_vendor_quality_checks = bool(val)
    elif fn == 11:  # LinearAlignment
      for fn2, wt2, val2, _ in iter_fields(bytes(val)):
        if fn2 == 1:
          r.reference_name, r.position, r.reverse_strand = _parse_position(bytes(val2))
        elif fn2 == 2:
          r.mapping_quality = _to_signed32(val2)
        elif fn2 == 3:
          op, ln = 0, 0
          for fn3, wt3, val3, _ in iter_fields(bytes(val2)):
            if fn3 == 1:
              op = val3
            elif fn3 == 2:
              ln = _to_signed64(val3)
          r.cigar.append((CIGAR_ENUM_TO_BAM.get(op, 15), ln))
    elif fn == 12:
      r.secondary_alignment = bool(val)
    elif fn == 13:
      r.supplementary_alignment = bool(val)
    elif fn == 14:
      r.aligned_sequence = bytes(val)
    elif fn == 15:
      r.aligned_quality = bytes(val)
    elif fn == 17:  # map<string, ListValue> info
      k, v = '', b''
      for fn2, wt2, val2, _ in iter_fields(bytes(val)):
        if fn2 == 1:
          k = bytes(val2).decode()
        elif fn2 == 2:
          v = bytes(val2)
      if k == 'HP':
        r.hp_values = _parse_list_value_ints(v)
  return r


def serialize_read(r: Read) -> bytes:
  out = bytearray()
  if r.fragment_name:
    out += f_bytes(4, r.fragment_name.encode())
  if r.proper_placement:
    out += f_varint(5, 1)
  if r.duplicate_fragment:
    out += f_varint(6, 1)
  if r.fragment_length:
    out += f_varint(7, r.fragment_length)
  if r.read_number:
    out += f_varint(8, r.read_number)
  if r.number_reads:
    out += f_varint(9, r.number_reads)
  posb = bytearray()
  if r.reference_name:
    posb += f_bytes(1, r.reference_name.encode())
  if r.position:
    posb += f_varint(2, r.position)
  if r.reverse_strand:
    posb += f_varint(3, 1)
  aln = bytearray(f_bytes(1, bytes(posb)))
  if r.mapping_quality:
    aln += f_varint(2, r.mapping_quality)
  for op, ln in r.cigar:
    aln += f_bytes(3, f_varint(1, op + 1) + f_varint(2, ln))
  out += f_bytes(11, bytes(aln))
  if r.secondary_alignment:
    out += f_varint(12, 1)
  if r.supplementary_alignment:
    out += f_varint(13, 1)
  if r.aligned_sequence:
    out += f_bytes(14, r.aligned_sequence)
  if r.aligned_quality:
    out += f_bytes(15, r.aligned_quality)
  if r.hp_values is not None:
    lv = b''.join(f_bytes(1, f_varint(7, v)) for v in r.hp_values)
    out += f_bytes(17, f_bytes(1, b'HP') + f_bytes(2, lv))
  return bytes(out)


@dataclasses.dataclass
class Variant:
  """nucleus.genomics.v1.Variant (variants.proto:52-73): reference_name=14, start=16,
  end=13, reference_bases=6, alternate_bases=7.  `raw` keeps the original serialisation so that
  re-emitting the variant is byte-exact (the CVO / tf.Example carry it opaquely)."""
  reference_name: str = ''
  start: int = 0
  end: int = 0
  reference_bases: str = ''
  alternate_bases: List[str] = dataclasses.field(default_factory=list)
  raw: Optional[bytes] = None
  # only read by the fuzzy read-support channel (channels.py); not parsed from / written to the wire here
  alternate_bases_rejected: List[str] = dataclasses.field(default_factory=list)   # variants.proto:75
  alt_ps: Optional[List[int]] = None       # info['ALT_PS'] int values
  alt_ps_ext: Optional[List[int]] = None   # info['ALT_PS_EXT'] int values

  def serialize(self) -> bytes:
    if self.raw is not None:
      return self.raw
    out = bytearray()  # canonical (field-number) order, like C++ SerializeToString
    if self.reference_bases:
      out += f_bytes(6, self.reference_bases.encode())
    for a in self.alternate_bases:
      out += f_bytes(7, a.encode())
    if self.end:
      out += f_varint(13, self.end)
    if self.reference_name:
      out += f_bytes(14, self.reference_name.encode())
    if self.start:
      out += f_varint(16, self.start)
    return bytes(out)


def parse_variant(buf: bytes) -> Variant:
  v = Variant(raw=bytes(buf))
  for fn, wt, val, _ in iter_fields(buf):
    if fn == 14:
      v.reference_name = bytes(val).decode()
    elif fn == 16:
      v.start = _to_signed64(val)
    elif fn == 13:
      v.end = _to_signed64(val)
    elif fn == 6:
      v.reference_bases = bytes(val).decode()
    elif fn == 7:
      v.alternate_bases.append(bytes(val).decode())
  return v


@dataclasses.dataclass
class DeepVariantCall:
  """learning.genomics.deepvariant.DeepVariantCall (deepvariant.proto:262-317):
  variant=1, allele_support=2 (map<string, SupportingReads{read_names=1}>),
  make_examples_alt_allele_indices=8 (repeated AltAlleleIndices{indices=1})."""
  variant: Variant = dataclasses.field(default_factory=Variant)
  allele_support: Dict[str, List[str]] = dataclasses.field(default_factory=dict)
  make_examples_alt_allele_indices: List[List[int]] = dataclasses.field(default_factory=list)
  # fields only the optional channels read (deepvariant.proto:280-287): rejected_allele_support=10, allele_frequency=3, ref_support=4
  rejected_allele_support: Dict[str, List[str]] = dataclasses.field(default_factory=dict)
  allele_frequency: Dict[str, float] = dataclasses.field(default_factory=dict)
  ref_support: List[str] = dataclasses.field(default_factory=list)


def _parse_indices(buf: bytes) -> List[int]:
  out = []
  for fn, wt, val, _ in iter_fields(buf):
    if fn == 1:
      if wt == WT_LEN:
        out.extend(_to_signed32(v) for v in unpack_varints(bytes(val)))
      else:
        out.append(_to_signed32(val))
  return out


def parse_deepvariant_call(buf: bytes) -> DeepVariantCall:
  c = DeepVariantCall()
  for fn, wt, val, _ in iter_fields(buf):
    if fn == 1:
      c.variant = parse_variant(bytes(val))
    elif fn == 2:
      k, names = '', []
      for fn2, wt2, val2, _ in iter_fields(bytes(val)):
        if fn2 == 1:
          k = bytes(val2).decode()
        elif fn2 == 2:
          names = [bytes(v3).decode() for f3, w3, v3, _ in iter_fields(bytes(val2)) if f3 == 1]
      c.allele_support[k] = names
    elif fn == 10:
      k, names = '', []
      for fn2, wt2, val2, _ in iter_fields(bytes(val)):
        if fn2 == 1:
          k = bytes(val2).decode()
        elif fn2 == 2:
          names = [bytes(v3).decode() for f3, w3, v3, _ in iter_fields(bytes(val2)) if f3 == 1]
      c.rejected_allele_support[k] = names
    elif fn == 3:
      k, f = '', 0.0
      for fn2, wt2, val2, _ in iter_fields(bytes(val)):
        if fn2 == 1:
          k = bytes(val2).decode()
        elif fn2 == 2:
          f = struct.unpack('<f', struct.pack('<I', val2))[0] if isinstance(val2, int) else struct.unpack('<f', bytes(val2))[0]
      c.allele_frequency[k] = f
    elif fn == 4:
      c.ref_support.append(bytes(val).decode())
    elif fn == 8:
      c.make_examples_alt_allele_indices.append(_parse_indices(bytes(val)))
  return c


def serialize_deepvariant_call(c: DeepVariantCall) -> bytes:
  """Inverse of parse_deepvariant_call for the fields modelled here (variant=1, allele_support=2 map entries
  {key=1, value=2 SupportingReads{read_names=1}}, make_examples_alt_allele_indices=8)."""
  out = bytearray(f_bytes(1, c.variant.serialize()))
  for alt, names in c.allele_support.items():
    out += f_bytes(2, f_bytes(1, alt.encode()) + f_bytes(2, b''.join(f_bytes(1, n.encode()) for n in names)))
  for idx in c.make_examples_alt_allele_indices:
    out += f_bytes(8, encode_alt_allele_indices(idx))
  return bytes(out)


def encode_alt_allele_indices(indices: Iterable[int]) -> bytes:
  """CallVariantsOutput.AltAlleleIndices{repeated int32 indices = 1} (proto3 -> packed)."""
  indices = list(indices)
  if not indices:
    return b''
  return f_bytes(1, packed_varints(indices))


def parse_alt_allele_indices(buf: bytes) -> List[int]:
  return _parse_indices(buf)


# ----------------------------------------------------------------------------
# tf.Example
# ----------------------------------------------------------------------------

def encode_tf_example(features: Dict[str, Tuple[str, list]]) -> bytes:
  """features: name -> ('bytes'|'int64'|'float', [values]).
  Example{Features features=1}; Features{map<string,Feature> feature=1};
  Feature{BytesList bytes_list=1 | FloatList float_list=2 | Int64List int64_list=3};
  BytesList{repeated bytes value=1}; Int64List{repeated int64 value=1 [packed]}."""
  body = bytearray()
  for name, (kind, values) in features.items():
    if kind == 'bytes':
      feat = f_bytes(1, b''.join(f_bytes(1, bytes(v)) for v in values))
    elif kind == 'int64':
      feat = f_bytes(3, f_bytes(1, packed_varints(values)))
    elif kind == 'float':
      feat = f_bytes(2, f_bytes(1, struct.pack(f'<{len(values)}f', *values)))
    else:
      raise ValueError(kind)
    entry = f_bytes(1, name.encode()) + f_bytes(2, feat)
    body += f_bytes(1, entry)
  return f_bytes(1, bytes(body))


def parse_tf_example(buf: bytes) -> Dict[str, Tuple[str, list]]:
  out: Dict[str, Tuple[str, list]] = {}
  for fn, wt, val, _ in iter_fields(buf):
    if fn != 1:
      continue
    for fn2, wt2, entry, _ in iter_fields(bytes(val)):
      if fn2 != 1:
        continue
      name, feat = '', b''
      for fn3, wt3, v3, _ in iter_fields(bytes(entry)):
        if fn3 == 1:
          name = bytes(v3).decode()
        elif fn3 == 2:
          feat = bytes(v3)
      for fn4, wt4, v4, _ in iter_fields(feat):
        if fn4 == 1:
          out[name] = ('bytes', [bytes(x) for f5, w5, x, _ in iter_fields(bytes(v4)) if f5 == 1])
        elif fn4 == 3:
          vals: List[int] = []
          for f5, w5, x, _ in iter_fields(bytes(v4)):
            if f5 == 1:
              if w5 == WT_LEN:
                vals.extend(_to_signed64(t) for t in unpack_varints(bytes(x)))
              else:
                vals.append(_to_signed64(x))
          out[name] = ('int64', vals)
        elif fn4 == 2:
          vals_f: List[float] = []
          for f5, w5, x, _ in iter_fields(bytes(v4)):
            if f5 == 1:
              if w5 == WT_LEN:
                vals_f.extend(struct.unpack(f'<{len(x)//4}f', bytes(x)))
              else:
                vals_f.append(struct.unpack('<f', bytes(x))[0])
          out[name] = ('float', vals_f)
  return out


# ----------------------------------------------------------------------------
# CallVariantsOutput
# ----------------------------------------------------------------------------

def encode_call_variants_output(variant_encoded: bytes, alt_allele_indices: Iterable[int],
                                genotype_probabilities: Iterable[float]) -> bytes:
  """CallVariantsOutput{variant=1, alt_allele_indices=2, genotype_probabilities=3 (packed double)}
  (deepvariant.proto:363-373)."""
  probs = list(genotype_probabilities)
  out = bytearray()
  out += f_bytes(1, variant_encoded)
  out += f_bytes(2, encode_alt_allele_indices(alt_allele_indices))
  out += f_bytes(3, struct.pack(f'<{len(probs)}d', *probs))
  return bytes(out)


def parse_call_variants_output(buf: bytes):
  variant, idx, probs = b'', [], []
  for fn, wt, val, _ in iter_fields(buf):
    if fn == 1:
      variant = bytes(val)
    elif fn == 2:
      idx = _parse_indices(bytes(val))
    elif fn == 3:
      if wt == WT_LEN:
        probs.extend(struct.unpack(f'<{len(val)//8}d', bytes(val)))
      else:
        probs.append(struct.unpack('<d', bytes(val))[0])
  return variant, idx, probs
