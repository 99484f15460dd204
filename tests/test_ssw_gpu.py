"""Batched Smith-Waterman on the device (csrc/dvb_ssw_gpu.cu: a warp per alignment, anti-diagonal wavefront over strips of 32 query
rows) against the host implementation (csrc/dvb_ssw.cu, pinned by the reference's ssw / fast_pass_aligner known answers in
tests/test_fast_pass_aligner.py): every field and the CIGAR string, pair by pair.  `-m gpu`."""
import os
import random

import pytest

from deepvariant_b200 import ssw

pytestmark = pytest.mark.gpu

KATS = [   # deepvariant/realigner/ssw_test.cc:47-58, python/ssw_misc_test.py:44-84, python/ssw_wrap_test.py:37-72
    ((4, 2, 4, 2), 'tttt', 'ttAtt', dict(cigar_string='2=1I2=')),
    ((4, 2, 4, 2), 'TTTTGGGGGGGGGGGGG', 'TTATTGGGGGGGGGGGGG', dict(cigar_string='2=1I15=')),
    ((2, 2, 3, 1), 'CAGCCTTTCTGACCCGGAAATCAAAATAGGCACAACAAA', 'CTGAGCCGGTAAATC',
     dict(sw_score=21, ref_begin=8, ref_end=21, query_begin=0, query_end=14, mismatches=2, cigar_string='4=1X4=1I5=')),
    ((2, 2, 3, 1), 'CTGAGCCGGTAAATC', 'CAGCCTTTCTGACCCGGAAATCAAAATAGGCACAACAAA',
     dict(sw_score=21, query_begin=8, query_end=21, ref_begin=0, ref_end=14, mismatches=2, cigar_string='8S4=1X4=1D5=17S')),
    ((4, 6, 8, 1), 'TTTGCCGAAGTTAAACCC', 'GCCGAAGTTA', dict(cigar_string='10=', ref_begin=3)),
]


def test_reference_known_answers_through_the_batch_entry_point():
  for params, ref, query, expected in KATS:
    al = ssw.align_batch([(ref, query)], *params)[0]
    for k, v in expected.items():
      assert getattr(al, k) == v, (k, al)


def _mutate(rng, s, n_sub, n_indel):
  s = list(s)
  for _ in range(n_sub):
    s[rng.randrange(len(s))] = rng.choice('ACGT')
  for _ in range(n_indel):
    at = rng.randrange(1, len(s) - 1)
    if rng.random() < 0.5:
      del s[at:at + rng.randint(1, 6)]
    else:
      s[at:at] = [rng.choice('ACGT') for _ in range(rng.randint(1, 6))]
  return ''.join(s)


def test_batch_equals_host_alignments_on_read_to_haplotype_shaped_pairs():
  """Trimmed long reads against haplotype windows (the alt-aligned pileup's workload), short reads against realigner windows, queries
  longer than one 32-row strip and than the reference, N bases, a pair with nothing to align, an empty query, a reference longer than
  the kernel's shared-memory window (host scans take over), ties between equally good placements (repeats)."""
  rng = random.Random(7)
  pairs = []
  for k in range(160):
    ref_len = rng.choice([40, 147, 221, 300, 420])
    ref = ''.join(rng.choice('ACGT') for _ in range(ref_len))
    lo = rng.randrange(0, max(1, ref_len - 30))
    hi = min(ref_len, lo + rng.choice([20, 33, 64, 100, 150, 260]))
    q = _mutate(rng, ref[lo:hi], rng.randint(0, 6), rng.randint(0, 3))
    if k % 9 == 0:
      q = ''.join(rng.choice('ACGT') for _ in range(rng.randint(3, 12))) + q + ''.join(rng.choice('ACGT') for _ in range(rng.randint(3, 30)))   # soft clips
    if k % 13 == 0:
      q = q[:len(q) // 2] + 'N' + q[len(q) // 2 + 1:]
    pairs.append((ref, q))
  pairs.append(('ACGT' * 30, 'ACGTACGTACGT'))                     # a repeat: many equally good placements
  pairs.append(('A' * 50, 'C' * 20))                             # nothing aligns
  pairs.append(('ACGTACGT', ''))                                 # empty query
  pairs.append((''.join(rng.choice('ACGT') for _ in range(2500)), ''.join(rng.choice('ACGT') for _ in range(90))))   # beyond the shared-memory window
  long_ref = ''.join(rng.choice('ACGT') for _ in range(900))
  pairs.append((long_ref, _mutate(rng, long_ref[100:700], 20, 8)))                                                     # 19 strips
  for params in ((2, 2, 3, 1), (4, 6, 8, 2)):
    got = ssw.align_batch(pairs, *params)
    host = ssw.Aligner(*params)
    for (ref, q), g in zip(pairs, got):
      host.set_reference_sequence(ref)
      assert g == host.align(q), (ref[:30], q[:30], params)


def test_alt_aligned_pileups_through_the_gpu_flow_equal_the_host_flow(tmp_path, monkeypatch):
  """make_examples --alt_aligned_pileup diff_channels on the planted genome (a 2-bp insertion and a 3-bp deletion among the candidates):
  CUDA encoder + batched GPU Smith-Waterman for the read-to-haplotype alignments against the CPU oracle encoder + host Smith-Waterman
  (the flow tools/check_pacbio_end_to_end.py pins on all 131 indel examples of the reference's golden.pacbio_examples): the tf.Examples
  must be the same, record for record - images 100 x 221 x 9 with both alt-aligned diff channels."""
  import test_candidates as tc
  from deepvariant_b200 import fast_pass_aligner as fpa, make_examples_native as men, pileup_image as pi, protos
  fa, bam_path, genome, sites = tc._planted_case(tmp_path)
  calls = {'batch': 0}
  real_batch = ssw.align_batch

  def counting_batch(pairs, *a, **k):
    calls['batch'] += len(pairs)
    return real_batch(pairs, *a, **k)
  monkeypatch.setattr(ssw, 'align_batch', counting_batch)
  monkeypatch.setattr(fpa.ssw, 'align_batch', counting_batch)
  extra = ('--alt_aligned_pileup', 'diff_channels')
  gpu_examples, gpu_cands = tc._run_cli(tmp_path, fa, bam_path, 'alt_gpu', extra=extra)
  assert calls['batch'] > 0, 'the GPU flow never used the batched Smith-Waterman'
  monkeypatch.setattr(men.ExamplesGenerator, '_gpu', lambda self: tc.OracleEncoder(pi.to_params(self.options.pic_options, height=self.pileup_image_height)))
  monkeypatch.setattr(fpa.FastPassAligner, 'ssw_batch_min', 1 << 30, raising=False)       # host dvb_ssw_align per pair
  before = calls['batch']
  cpu_examples, cpu_cands = tc._run_cli(tmp_path, fa, bam_path, 'alt_cpu', extra=extra)
  assert calls['batch'] == before
  assert gpu_cands == cpu_cands and gpu_examples == cpu_examples
  shapes = [protos.parse_tf_example(r)['image/shape'][1] for r in gpu_examples]
  assert shapes and all(s == [100, 221, 9] for s in shapes)


@pytest.mark.gpu
def test_realigner_with_device_smith_waterman_equals_host():
  """The local realigner over BASELINE config 1's reads (NA12878 chr20:10,000,000-10,004,000): with `ssw_device` set, FastPassAligner
  sends its read-to-haplotype and haplotype-to-reference alignments through dvb_ssw_align_batch; every realigned read (position,
  CIGAR) equals the host flow's."""
  from deepvariant_b200 import bam, candidates as cand, fasta, realigner
  g = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
  table = bam.NativeBamTable(os.path.join(g, 'quickstart.chr20_10mb.bam'), bam.ReadRequirements(min_mapping_quality=5))
  ref = fasta.IndexedFastaReader(os.path.join(g, 'quickstart.chr20_10mb.fa.gz'))
  outs = []
  for device in (None, 0):
    rl = realigner.Realigner(ref, realigner.RealignerOptions())
    rl.ssw_device = device
    rows_out = []
    for p0 in range(10_000_000, 10_004_000, 1000):
      rows = cand.region_reads(table, 'chr20', p0, p0 + 1000, 1500, 609314161)
      rows_out += [(r.fragment_name, r.read_number, r.position, tuple(r.cigar)) for r in rl.realign_reads(table, 'chr20', rows, (p0, p0 + 1000))]
    outs.append(rows_out)
  assert len(outs[0]) > 1000 and outs[0] == outs[1]
