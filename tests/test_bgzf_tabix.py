"""*.vcf.gz outputs: BGZF members + tabix index (deepvariant_b200/bgzf_tabix.py), against the structure of the reference's golden .tbi
files and by using the index to answer region queries."""
import gzip
import os
import random
import struct

import pytest

from deepvariant_b200 import bgzf_tabix as bt
from deepvariant_b200 import postprocess_variants as pp

GOLDEN = os.path.join(os.path.dirname(__file__), 'golden')


def reg2bins(beg, end):
  end -= 1
  out = [0]
  for shift, offset in ((26, 1), (23, 9), (20, 73), (17, 585), (14, 4681)):
    out += list(range(offset + (beg >> shift), offset + (end >> shift) + 1))
  return out


def query(path, index, contig, beg, end):
  """A tabix reader in miniature: bins that overlap, chunks not before the linear-index offset, lines filtered by position."""
  if contig not in index['names']:
    return []
  ref = index['refs'][index['names'].index(contig)]
  lo = ref['linear'][min(beg >> 14, len(ref['linear']) - 1)] if ref['linear'] else 0
  chunks = sorted(c for b in reg2bins(beg, end) if b != bt.META_BIN for c in ref['bins'].get(b, []) if c[1] > lo)
  out = []
  for c0, c1 in chunks:
    for line in bt.read_from_virtual_offset(path, max(c0, lo) if c0 < lo < c1 else c0, c1).decode().splitlines():
      f = line.split('\t')
      start = int(f[1]) - 1
      stop = int(f[7][4:]) if f[7].startswith('END=') else start + len(f[3])
      if f[0] == contig and start < end and stop > beg:
        out.append(line)
  return out


def _contigs(lines):
  return [(l.split('ID=')[1].split(',')[0], int(l.split('length=')[1].rstrip('>'))) for l in lines if l.startswith('##contig')]


@pytest.mark.parametrize('cvo,golden_vcf,golden_tbi', [
    ('golden.postprocess_single_site_input-00000-of-00001.tfrecord.gz', 'golden.postprocess_single_site_output.vcf', 'golden.postprocess_single_site_output.vcf.gz.tbi'),
    ('golden.postprocess_pacbio_input-00000-of-00001.tfrecord.gz', 'golden.postprocess_single_site_output_pacbio.vcf', 'golden.postprocess_single_site_output_pacbio.vcf.gz.tbi'),
])
def test_vcf_gz_is_bgzf_with_the_golden_index_structure(tmp_path, cvo, golden_vcf, golden_tbi):
  want = open(os.path.join(GOLDEN, golden_vcf)).read()
  out = str(tmp_path / 'o.vcf.gz')
  pp.postprocess_variants(os.path.join(GOLDEN, cvo), out, _contigs(want.splitlines()))
  raw = open(out, 'rb').read()
  assert gzip.decompress(raw).decode() == want                                     # same text as the golden
  assert raw[:4] == b'\x1f\x8b\x08\x04' and raw[12:14] == b'BC' and raw.endswith(bt.EOF_MEMBER)
  ours, golden = bt.parse_tbi(open(out + '.tbi', 'rb').read()), bt.parse_tbi(open(os.path.join(GOLDEN, golden_tbi), 'rb').read())
  for key in ('format', 'columns', 'meta', 'skip', 'names', 'n_no_coor'):
    assert ours[key] == golden[key], key
  for o, g in zip(ours['refs'], golden['refs']):
    assert len(o['linear']) == len(g['linear'])
    assert o['bins'][bt.META_BIN][1] == g['bins'][bt.META_BIN][1]                  # (records, 0)
    # htslib folds a 16-kb bin into its parent when its chunks share one BGZF member; queries see the same records either way
    assert sum(len(c) for b, c in o['bins'].items() if b != bt.META_BIN) >= sum(len(c) for b, c in g['bins'].items() if b != bt.META_BIN)
  # the header sits in its own member: the first record starts at offset 0 of a member, as in the golden
  first = ours['refs'][0]['bins'][bt.META_BIN][0][0]
  assert first & 0xffff == 0 and golden['refs'][0]['bins'][bt.META_BIN][0][0] & 0xffff == 0
  # every chunk reads back as whole lines
  records = [l for l in want.splitlines() if not l.startswith('#')]
  span = ours['refs'][0]['bins'][bt.META_BIN][0]
  assert bt.read_from_virtual_offset(out, span[0], span[1]).decode().splitlines() == records
  # region queries through the index == filtering the text
  rng = random.Random(3)
  positions = [int(l.split('\t')[1]) - 1 for l in records]
  contig = records[0].split('\t')[0]
  for _ in range(40):
    a = rng.randrange(min(positions) - 2000, max(positions) + 2000)
    b = a + rng.choice([1, 50, 2000, 40000])
    brute = [l for l in records if int(l.split('\t')[1]) - 1 < b and int(l.split('\t')[1]) - 1 + len(l.split('\t')[3]) > a]
    assert query(out, ours, contig, a, b) == brute
  assert query(out, ours, 'chrNope', 0, 10) == []


def test_many_members_two_contigs_and_gvcf_ends(tmp_path):
  path = str(tmp_path / 'big.g.vcf.gz')
  w = bt.BgzfVcfWriter(path)
  w.write_header('##fileformat=VCFv4.2\n#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\ts\n')
  rng = random.Random(5)
  lines = []
  for contig, n in (('chr1', 3000), ('chr2', 1500)):
    pos = 0
    for _ in range(n):
      length = rng.choice([1, 1, 3, 700, 20000])
      line = f'{contig}\t{pos + 1}\t.\tA\t<*>\t0\t.\tEND={pos + length}\tGT:PL\t0/0:{rng.random()},{"9" * rng.randrange(1, 60)}\n'
      w.write_record(line, contig, pos, pos + length)
      lines.append((contig, pos, pos + length, line))
      pos += length + rng.choice([0, 0, 5, 30000])
  w.close()
  raw = open(path, 'rb').read()
  assert gzip.decompress(raw).decode().splitlines()[2:] == [l[3].rstrip('\n') for l in lines]
  # members: each at most 64 KiB, chained by the BC size field up to the EOF member
  p, members = 0, 0
  while p < len(raw):
    size = struct.unpack('<H', raw[p + 16:p + 18])[0] + 1
    p += size
    members += 1
  assert p == len(raw) and members > 5
  index = bt.parse_tbi(open(path + '.tbi', 'rb').read())
  assert index['names'] == ['chr1', 'chr2']
  assert [r['bins'][bt.META_BIN][1][0] for r in index['refs']] == [3000, 1500]
  for contig in ('chr1', 'chr2'):
    own = [l for l in lines if l[0] == contig]
    hi = own[-1][2]
    for _ in range(60):
      a = rng.randrange(0, hi)
      b = a + rng.choice([1, 100, 16384, 300000])
      brute = [l[3].rstrip('\n') for l in own if l[1] < b and l[2] > a]
      assert query(path, index, contig, a, b) == brute, (contig, a, b)


def test_reg2bin_known_values():
  assert bt.reg2bin(0, 1) == 4681 and bt.reg2bin(10000000, 10000001) == 5291 and bt.reg2bin(16383, 16385) == 585
  assert bt.reg2bin(0, 1 << 29) == 0 and bt.reg2bin((1 << 26) - 1, (1 << 26) + 1) == 0 and bt.reg2bin(1 << 26, (1 << 26) + 1) == 4681 + 4096
