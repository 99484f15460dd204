"""BGZF-compressed VCF text with its tabix index: what `postprocess_variants --outfile x.vcf.gz` leaves on disk in the reference
(nucleus VcfWriter -> htslib bgzf, then build_index -> tbx_index_build, deepvariant/postprocess_variants.py:1583-1594, 2355-2368), so
that bcftools / tabix / hap.py can query the output.  htslib is not in this image; the formats are restated from the SAM/tabix
specifications: BGZF = gzip members of at most 64 KiB with the 'BC' extra field holding the member size and a 28-byte empty member
at the end; virtual offset = member start << 16 | offset inside the member; .tbi = "TBI\\1", the VCF column preset (format 2,
seq 1, beg 2, end 0, meta '#', skip 0), the names of the contigs that have records, and per contig the UCSC binning index
(min_shift 14, 5 levels: bin -> chunks of virtual offsets), the htslib pseudo-bin 37450 (file span + record count) and the 16-kb
linear index, itself BGZF-compressed.  As htslib's VCF writer does, the header is flushed into its own member(s) and no record
straddles two members.  Checked against the structure of the reference's golden .tbi files (same names, bins, intervals and
counts for the same VCF) and by reading the index back (tests/test_bgzf_tabix.py); byte equality with htslib's files is not a goal
(member boundaries depend on the deflate implementation)."""
from __future__ import annotations

import struct
import zlib
from typing import Dict, List, Optional, Tuple

BLOCK = 0xff00
META_BIN = 37450
EOF_MEMBER = bytes.fromhex('1f8b08040000000000ff0600424302001b0003000000000000000000')


def bgzf_member(data: bytes, level: int = 6) -> bytes:
  co = zlib.compressobj(level, zlib.DEFLATED, -15)
  body = co.compress(data) + co.flush()
  return struct.pack('<4BI2BH2BHH', 0x1f, 0x8b, 8, 4, 0, 0, 0xff, 6, ord('B'), ord('C'), 2, len(body) + 25) + body + \
      struct.pack('<II', zlib.crc32(data) & 0xffffffff, len(data))


def reg2bin(beg: int, end: int) -> int:
  """The UCSC binning scheme of SAM / tabix (0-based half-open interval)."""
  end -= 1
  if beg >> 14 == end >> 14:
    return ((1 << 15) - 1) // 7 + (beg >> 14)
  if beg >> 17 == end >> 17:
    return ((1 << 12) - 1) // 7 + (beg >> 17)
  if beg >> 20 == end >> 20:
    return ((1 << 9) - 1) // 7 + (beg >> 20)
  if beg >> 23 == end >> 23:
    return ((1 << 6) - 1) // 7 + (beg >> 23)
  if beg >> 26 == end >> 26:
    return ((1 << 3) - 1) // 7 + (beg >> 26)
  return 0


class BgzfVcfWriter:
  """write_header(text) once, write_record(line, contig, start, end) in file order (contigs contiguous, positions sorted), close()
  -> path (BGZF) and path + '.tbi'."""

  def __init__(self, path: str, write_index: bool = True):
    self._f = open(path, 'wb')
    self._path = path
    self._coffset = 0
    self._buf = bytearray()
    self._records: List[Tuple[str, int, int, int]] = []      # contig, beg, end, virtual offset of the line
    self._write_index = write_index

  def _flush(self) -> None:
    if self._buf:
      member = bgzf_member(bytes(self._buf))
      self._f.write(member)
      self._coffset += len(member)
      self._buf = bytearray()

  def write_header(self, text: str) -> None:
    data = text.encode()
    for i in range(0, len(data), BLOCK):
      self._buf += data[i:i + BLOCK]
      self._flush()

  def write_record(self, line: str, contig: str, start: int, end: int) -> None:
    data = line.encode()
    if len(self._buf) + len(data) > BLOCK:
      self._flush()
    self._records.append((contig, start, max(end, start + 1), (self._coffset << 16) | len(self._buf)))
    pos = 0
    while len(self._buf) + len(data) - pos > BLOCK:           # only a line longer than a member has to straddle
      take = BLOCK - len(self._buf)
      self._buf += data[pos:pos + take]
      pos += take
      self._flush()
    self._buf += data[pos:]

  def close(self) -> None:
    self._flush()
    end_offset = self._coffset << 16
    self._f.write(EOF_MEMBER)
    self._f.close()
    if self._write_index:
      with open(self._path + '.tbi', 'wb') as f:
        f.write(build_tbi(self._records, end_offset))


def build_tbi(records: List[Tuple[str, int, int, int]], end_offset: int) -> bytes:
  names: List[str] = []
  per: Dict[str, dict] = {}
  for i, (contig, beg, end, voff) in enumerate(records):
    nxt = records[i + 1][3] if i + 1 < len(records) else end_offset
    c = per.get(contig)
    if c is None:
      names.append(contig)
      c = per[contig] = {'bins': {}, 'linear': [], 'first': voff, 'last': nxt, 'n': 0, 'last_bin': None}
    b = reg2bin(beg, end)
    chunks = c['bins'].setdefault(b, [])
    if c['last_bin'] == b and chunks and chunks[-1][1] == voff:
      chunks[-1][1] = nxt                                     # a run of consecutive records in one bin is one chunk
    else:
      chunks.append([voff, nxt])
    c['last_bin'] = b
    c['last'] = nxt
    c['n'] += 1
    w0, w1 = beg >> 14, (end - 1) >> 14
    if len(c['linear']) <= w1:
      c['linear'] += [None] * (w1 + 1 - len(c['linear']))
    for w in range(w0, w1 + 1):
      if c['linear'][w] is None:
        c['linear'][w] = voff
  blob = b''.join(n.encode() + b'\0' for n in names)
  out = bytearray(b'TBI\1' + struct.pack('<8i', len(names), 2, 1, 2, 0, ord('#'), 0, len(blob)) + blob)
  for n in names:
    c = per[n]
    out += struct.pack('<i', len(c['bins']) + 1)
    out += struct.pack('<IiQQQQ', META_BIN, 2, c['first'], c['last'], c['n'], 0)
    for b in sorted(c['bins']):
      out += struct.pack('<Ii', b, len(c['bins'][b])) + b''.join(struct.pack('<QQ', x, y) for x, y in c['bins'][b])
    linear = c['linear']
    for w in range(len(linear) - 2, -1, -1):                  # unset windows take the next window's offset (hts_idx_finish)
      if linear[w] is None:
        linear[w] = linear[w + 1]
    out += struct.pack('<i', len(linear)) + b''.join(struct.pack('<Q', x) for x in linear)
  out += struct.pack('<Q', 0)                                 # n_no_coor
  data = bytes(out)
  return b''.join(bgzf_member(data[i:i + BLOCK]) for i in range(0, len(data), BLOCK)) + EOF_MEMBER


# ---- reading back (tests, and region queries on our own output) ---------------------------------------------------------------------------
def parse_tbi(data: bytes) -> dict:
  import gzip
  d = gzip.decompress(data)
  assert d[:4] == b'TBI\1'
  n_ref, fmt, col_seq, col_beg, col_end, meta, skip, l_nm = struct.unpack('<8i', d[4:36])
  names = [x.decode() for x in d[36:36 + l_nm].split(b'\0')[:-1]]
  p = 36 + l_nm
  refs = []
  for _ in range(n_ref):
    (n_bin,) = struct.unpack('<i', d[p:p + 4])
    p += 4
    bins = {}
    for _ in range(n_bin):
      b, n_chunk = struct.unpack('<Ii', d[p:p + 8])
      p += 8
      bins[b] = [struct.unpack('<QQ', d[p + 16 * i:p + 16 * i + 16]) for i in range(n_chunk)]
      p += 16 * n_chunk
    (n_intv,) = struct.unpack('<i', d[p:p + 4])
    p += 4
    linear = list(struct.unpack('<%dQ' % n_intv, d[p:p + 8 * n_intv]))
    p += 8 * n_intv
    refs.append({'bins': bins, 'linear': linear})
  return {'format': fmt, 'columns': (col_seq, col_beg, col_end), 'meta': chr(meta), 'skip': skip, 'names': names, 'refs': refs,
          'n_no_coor': struct.unpack('<Q', d[p:p + 8])[0] if len(d) >= p + 8 else None}


def read_from_virtual_offset(path: str, voff: int, end_voff: Optional[int] = None) -> bytes:
  """The uncompressed bytes of [voff, end_voff) (to the end of the data when end_voff is None)."""
  out = bytearray()
  with open(path, 'rb') as f:
    coffset, uoffset = voff >> 16, voff & 0xffff
    while True:
      f.seek(coffset)
      head = f.read(18)
      if len(head) < 18:
        break
      size = struct.unpack('<H', head[16:18])[0] + 1
      f.seek(coffset)
      member = f.read(size)
      data = zlib.decompress(member[18:-8], -15)
      if end_voff is not None and coffset == end_voff >> 16:
        out += data[uoffset:end_voff & 0xffff]
        break
      out += data[uoffset:]
      coffset += size
      uoffset = 0
      if end_voff is not None and coffset > end_voff >> 16:
        break
  return bytes(out)
