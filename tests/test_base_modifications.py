"""MM / ML / MN -> Read.base_modifications (deepvariant_b200/bam.py parse_base_modifications) against the reference's known answers
(third_party/nucleus/io/sam_reader_test.cc:543-700, Parse5mCAuxTagTest x7 + Parse6mATagTest), and the tags' way through both BAM
readers, the scratch BAM of the realigner path and into the base_methylation channel plane."""
import os
import struct

import numpy as np
import pytest

from deepvariant_b200 import bam
from deepvariant_b200.protos import Read, parse_cigar_string


@pytest.mark.parametrize('seq,reverse,mm,ml,want', [
    ('TCTCTCTCTCTCTCTCTCTC', False, 'C+m?,1,1,1,1,1', [1, 2, 3, 4, 5], [0, 0, 0, 1, 0, 0, 0, 2, 0, 0, 0, 3, 0, 0, 0, 4, 0, 0, 0, 5]),                     # BasicCase
    ('ACACACACACTCTCTCTCTC', False, 'A-a.,0,0,0,0,0;C+m?,1,1,1,1,1', [1] * 5 + [2] * 5, [0, 0, 0, 2] * 5),                                               # MultipleModifications
    ('ACACACACACTCTCTCTCTC', False, 'A-a.,0,0,0,0,0;C+m?,1,1,1,1,1;T-a.,0,0,0,0,0', [1] * 5 + [2] * 5 + [1] * 5, [0, 0, 0, 2] * 5),                       # ThreeModifications
    ('CACAACAAACAAAAC', False, 'A-a.,0,0,0,0,0;C+m?,0,3', [0, 0, 0, 0, 0, 1, 2], [1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 2]),                          # VariableMMDelta
    ('TTTTTGGGGG', True, 'C+m?,0,0,0,0,0', [1, 1, 1, 1, 1], [0, 0, 0, 0, 0, 1, 1, 1, 1, 1]),                                                             # ReverseStrandModifications
])
def test_parse_5mc_known_answers(seq, reverse, mm, ml, want):
  got = bam.parse_base_modifications(seq.encode(), reverse, mm, ml)
  assert list(got['5mC']) == want


def test_mn_tag_must_match_the_sequence_length():
  assert bam.parse_base_modifications(b'CCCCCTTTTT', False, 'C+m?,0,0,0,0,0', [1] * 5, mn=11) == {}          # MismatchMNTag
  assert bam.parse_base_modifications(b'CCCCCTTTTT', False, 'C+m?,0,0,0,0,0', [1] * 5, mn=10) != {}          # MatchMNTag
  assert bam.parse_base_modifications(b'CCCCC', False, 'C+m?,0,0,0,0,0', [1] * 3) == {}                      # ML runs out: everything is void
  assert bam.parse_base_modifications(b'CCCCC', False, None, [1]) == {} and bam.parse_base_modifications(b'CCCCC', False, 'C+m?,0', None) == {}


def test_parse_5mc_and_6ma_on_both_strands():
  got = bam.parse_base_modifications(b'ACCCAGGGTGGGTGGG', False, 'C+m?,0,0,0;A+a?,0,0;T-a?,0,0;', [7, 8, 9, 1, 2, 3, 4])     # Parse5mCand6mA
  assert list(got['5mC']) == [0, 7, 8, 9] + [0] * 12
  assert list(got['6mA']) == [1, 0, 0, 0, 2, 0, 0, 0, 3, 0, 0, 0, 4, 0, 0, 0]
  # the merge of the two 6mA entries is std::max over chars: bytes >= 128 are negative there and lose against the other entry's zero
  got = bam.parse_base_modifications(b'ACCCAGGGTGGGTGGG', False, 'A+a?,0,0;T-a?,0,0', [200, 100, 3, 250])
  assert list(got['6mA']) == [0, 0, 0, 0, 100, 0, 0, 0, 3, 0, 0, 0, 0, 0, 0, 0]


def _record(name, pos, seq, qual, cigar, flag, aux):
  ops = parse_cigar_string(cigar)
  codes = bam._SEQ_CODE_LUT[np.frombuffer(seq.encode(), dtype=np.uint8)]   # pylint: disable=protected-access
  if len(codes) & 1:
    codes = np.append(codes, np.uint8(0))
  packed = ((codes[0::2] << 4) | codes[1::2]).astype(np.uint8).tobytes()
  body = struct.pack('<iiBBHHHiiii', 0, pos, len(name) + 1, 60, 0, len(ops), flag, len(seq), -1, -1, 0)
  body += name.encode() + b'\0' + b''.join(struct.pack('<I', (ln << 4) | op) for op, ln in ops) + packed + bytes(qual) + aux
  return struct.pack('<i', len(body)) + body


def _write(path, records):
  text = b'@HD\tVN:1.6\tSO:coordinate\n'
  hdr = b'BAM\1' + struct.pack('<i', len(text)) + text + struct.pack('<i', 1) + struct.pack('<i', 6) + b'chr20\0' + struct.pack('<i', 100000)
  open(path, 'wb').write(bam._bgzf(hdr + b''.join(records)))   # pylint: disable=protected-access


def test_tags_reach_the_reads_through_both_readers_and_the_scratch_bam(tmp_path):
  seq = 'TCTCTCTCTCTCTCTCTCTC'
  mm = b'MMZC+m?,1,1,1,1,1;\0' + b'MLBC' + struct.pack('<i', 5) + bytes([10, 20, 128, 254, 255]) + b'MNi' + struct.pack('<i', 20)
  ultima = b'tpBc' + struct.pack('<i', 20) + struct.pack('<20b', *([0, 1, -1, 2] * 5)) + b't0Z' + bytes(range(40, 60)) + b'\0' + b'HPC' + bytes([2])
  path = str(tmp_path / 'aux.bam')
  _write(path, [_record('a', 100, seq, [30] * 20, '20M', 0, mm), _record('b', 120, seq, [30] * 20, '20M', 16, mm), _record('c', 140, seq, [30] * 20, '20M', 0, ultima),
                _record('d', 160, seq, [30] * 20, '20M', 0, b'')])
  want_fwd = [0, 0, 0, 10, 0, 0, 0, 20, 0, 0, 0, 128, 0, 0, 0, 254, 0, 0, 0, 255]
  for reader in (bam.BamReader(path, bam.ReadRequirements(min_mapping_quality=0), parse_aux=True).query('chr20', 0, 1000),
                 bam.NativeBamTable(path, bam.ReadRequirements(min_mapping_quality=0), parse_aux=True).reads()):
    a, b, c, d = reader
    assert list(a.base_modifications['5mC']) == want_fwd
    # reverse-strand alignment: MM walks the reverse complement (GAGAGA...: no C at all) -> the entry never completes
    assert b.base_modifications is None
    assert c.tp_values == [0, 1, -1, 2] * 5 and c.t0_value == bytes(range(40, 60)) and c.hp_values == [2] and c.base_modifications is None
    assert d.base_modifications is None and d.tp_values is None and d.t0_value is None
  # without parse_aux nothing is parsed
  assert bam.NativeBamTable(path, bam.ReadRequirements(min_mapping_quality=0)).reads()[0].base_modifications is None
  # the scratch BAM of the realigner / normaliser path carries the parsed data on
  reads = bam.NativeBamTable(path, bam.ReadRequirements(min_mapping_quality=0), parse_aux=True).reads()
  again = bam.scratch_table(reads, [('chr20', 100000)], bam.ReadRequirements(min_mapping_quality=0), parse_aux=True).reads()
  assert list(again[0].base_modifications['5mC']) == want_fwd and again[2].tp_values == reads[2].tp_values and again[2].t0_value == reads[2].t0_value
  # ... and the channel plane is ScaleColorVector(255) of the bytes
  from deepvariant_b200 import channels
  assert list(channels.base_plane(again[0], 0)) == [int(np.float32(254) * (np.float32(v) / np.float32(255))) for v in want_fwd]
