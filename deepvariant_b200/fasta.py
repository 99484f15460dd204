"""Indexed FASTA reader (plain or gzip/bgzip + .fai) with the semantics make_examples relies on:
nucleus IndexedFastaReader with keep_true_case=false (deepvariant/make_examples_native.cc:126-135):
bases are upper-cased; query(contig, start, end) is 0-based half-open and must lie inside the contig."""
from __future__ import annotations

import gzip
from typing import Dict, Tuple


class IndexedFastaReader:

  def __init__(self, fasta_path: str, fai_path: str = None):
    fai_path = fai_path or fasta_path + '.fai'
    self._index: Dict[str, Tuple[int, int, int, int]] = {}
    self.contig_order = []
    with open(fai_path) as f:
      for line in f:
        p = line.rstrip('\n').split('\t')
        if len(p) < 5:
          continue
        self._index[p[0]] = (int(p[1]), int(p[2]), int(p[3]), int(p[4]))  # length, offset, linebases, linewidth
        self.contig_order.append(p[0])
    opener = gzip.open if fasta_path.endswith('.gz') else open
    with opener(fasta_path, 'rb') as f:
      self._raw = f.read()          # offsets in .fai refer to the uncompressed stream
    self._cache: Dict[str, bytes] = {}

  def n_bases(self, contig: str) -> int:
    return self._index[contig][0]

  def has_contig(self, contig: str) -> bool:
    return contig in self._index

  def _contig(self, contig: str) -> bytes:
    if contig not in self._cache:
      length, offset, linebases, linewidth = self._index[contig]
      n_lines = (length + linebases - 1) // linebases
      raw = self._raw[offset:offset + n_lines * linewidth]
      seq = raw.replace(b'\n', b'').replace(b'\r', b'')[:length]
      self._cache[contig] = seq.upper()
    return self._cache[contig]

  def is_valid_interval(self, contig: str, start: int, end: int) -> bool:
    return contig in self._index and 0 <= start <= end <= self._index[contig][0]

  def query(self, contig: str, start: int, end: int) -> str:
    if not self.is_valid_interval(contig, start, end):
      raise ValueError(f'Invalid interval {contig}:{start}-{end}')
    return self._contig(contig)[start:end].decode()
