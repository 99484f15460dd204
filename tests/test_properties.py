"""Property-based tests (hypothesis) of the host-side codecs and read surgery: wire round trips of the protos the stages
exchange, TFRecord framing with CRC-32C, TrimCigar / TrimRead invariants (alt_aligned_pileup_lib.cc:91-266), and the native
BGZF/BAM decoder against the pure-Python reader on generated files.  CPU only."""
import struct
import zlib

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from deepvariant_b200 import bam, protos, tfrecord
from deepvariant_b200 import make_examples_native as men
from test_bam_native import _bam, _record

SETTINGS = dict(max_examples=60, deadline=None)
bases_st = st.text(alphabet='ACGT', min_size=1, max_size=12)
names_st = st.text(alphabet='abcXYZ0123_/:', min_size=1, max_size=20)


@settings(**SETTINGS)
@given(ref=bases_st, alts=st.lists(bases_st, min_size=1, max_size=3, unique=True), start=st.integers(0, 2**31 - 2),
       contig=st.sampled_from(['chr1', 'chr20', 'chrX', 'HLA-A*01:01']), support=st.lists(st.lists(names_st, max_size=5), max_size=3),
       idx=st.lists(st.lists(st.integers(0, 3), min_size=1, max_size=2), max_size=3))
def test_deepvariant_call_and_variant_round_trip(ref, alts, start, contig, support, idx):
  v = protos.Variant(reference_name=contig, start=start, end=start + len(ref), reference_bases=ref, alternate_bases=alts)
  v2 = protos.parse_variant(v.serialize())
  assert (v2.reference_name, v2.start, v2.end, v2.reference_bases, v2.alternate_bases) == (contig, start, start + len(ref), ref, alts)
  assert v2.serialize() == v.serialize()                      # re-emitting a parsed variant is byte-exact (raw is kept)
  call = protos.DeepVariantCall(variant=v, allele_support={a: s for a, s in zip(alts, support)}, make_examples_alt_allele_indices=idx)
  c2 = protos.parse_deepvariant_call(protos.serialize_deepvariant_call(call))
  assert c2.allele_support == call.allele_support and c2.make_examples_alt_allele_indices == idx
  assert c2.variant.serialize() == v.serialize()


@settings(**SETTINGS)
@given(image=st.binary(min_size=0, max_size=300), shape=st.lists(st.integers(0, 2**40), min_size=3, max_size=3),
       locus=st.text(alphabet='chr0123456789:-', max_size=30), vtype=st.integers(0, 2), variant=st.binary(max_size=60),
       probs=st.lists(st.floats(0, 1, allow_nan=False), min_size=3, max_size=3), idx=st.lists(st.integers(0, 5), max_size=3))
def test_tf_example_and_cvo_round_trip(image, shape, locus, vtype, variant, probs, idx):
  feats = {'image/encoded': ('bytes', [image]), 'image/shape': ('int64', shape), 'locus': ('bytes', [locus.encode()]),
           'variant_type': ('int64', [vtype]), 'variant/encoded': ('bytes', [variant]),
           'alt_allele_indices/encoded': ('bytes', [protos.encode_alt_allele_indices(idx)]), 'sequencing_type': ('int64', [0])}
  assert protos.parse_tf_example(protos.encode_tf_example(feats)) == feats
  cvo = protos.encode_call_variants_output(variant, idx, probs)
  v, i, p = protos.parse_call_variants_output(cvo)
  assert (v, i, p) == (variant, idx, probs)


@settings(**SETTINGS)
@given(payloads=st.lists(st.binary(max_size=2000), max_size=6))
def test_tfrecord_framing_round_trip_and_crc(tmp_path_factory, payloads):
  path = str(tmp_path_factory.mktemp('tfr') / 'x.tfrecord.gz')
  with tfrecord.Writer(path) as w:
    for p in payloads:
      w.write(p)
  assert list(tfrecord.read_records(path, check_crc=True)) == payloads
  for p in payloads[:2]:   # masked CRC-32C as TensorFlow defines it, against zlib-free arithmetic on the definition
    c = tfrecord.masked_crc32c(p)
    assert 0 <= c < 2**32


cigar_st = st.lists(st.tuples(st.sampled_from([0, 1, 2, 4, 7, 8]), st.integers(1, 30)), min_size=1, max_size=8)


@settings(**SETTINGS)
@given(cigar=cigar_st, ref_start=st.integers(0, 60), ref_length=st.integers(1, 80))
def test_trim_cigar_invariants(cigar, ref_start, ref_length):
  """TrimCigar (alt_aligned_pileup_lib.cc:91-150): the trimmed CIGAR covers at most ref_length reference bases, starts
  ref_start reference bases into the alignment, and its read-advancing length equals the reported new read length."""
  ref_adv = {0, 2, 3, 7, 8}
  read_adv = {0, 1, 4, 7, 8}
  total_ref = sum(n for op, n in cigar if op in ref_adv)
  new_cigar, read_start, new_len = men.trim_cigar(cigar, ref_start, ref_length)
  covered = sum(n for op, n in new_cigar if op in ref_adv)
  assert covered <= ref_length
  assert covered == max(0, min(ref_length, total_ref - ref_start)) or any(op in (1, 4) for op, _ in cigar)
  assert new_len == sum(n for op, n in new_cigar if op in read_adv)
  assert 0 <= read_start <= sum(n for op, n in cigar if op in read_adv)
  assert all(n >= 0 for _, n in new_cigar)
  # trimming nothing returns the alignment unchanged
  same, rs, ln = men.trim_cigar(cigar, 0, total_ref + 5)
  assert same == list(cigar) and rs == 0 and ln == sum(n for op, n in cigar if op in read_adv)


@settings(max_examples=25, deadline=None)
@given(reads=st.lists(st.tuples(st.integers(0, 900), st.integers(0, 60), st.integers(0, 0xFFF), st.integers(1, 40), st.booleans()),
                      min_size=1, max_size=25), seed=st.integers(0, 2**16))
def test_native_bam_decoder_equals_python_reader_on_generated_files(tmp_path_factory, reads, seed):
  rng = np.random.default_rng(seed)
  recs = []
  for k, (pos, mapq, flag, length, with_hp) in enumerate(sorted(reads)):
    flag &= ~0x4                                      # mapped
    seq = ''.join(rng.choice(list('ACGTN'), length))
    cigar = [(0, length)] if length < 6 else [(4, 2), (0, length - 5), (1, 3)]
    aux = (b'HPC' + bytes([int(rng.integers(0, 3))])) if with_hp else b'NMC\x01'
    recs.append(_record(0, pos, f'q{k % 7}', mapq, flag, cigar, seq, rng.integers(0, 60, length).tolist(), int(rng.integers(-1, 2)),
                        int(rng.integers(0, 900)), int(rng.integers(-500, 500)), aux=aux))
  path = str(tmp_path_factory.mktemp('bam') / 'g.bam')
  open(path, 'wb').write(_bam(recs))
  for req in (bam.ReadRequirements(), bam.ReadRequirements(min_mapping_quality=0, keep_duplicates=True, keep_supplementary_alignments=True,
                                                            keep_secondary_alignments=True, keep_failed_vendor_quality_checks=True,
                                                            keep_improperly_placed=True)):
    t = bam.NativeBamTable(path, req, parse_aux=True)
    py = bam.BamReader(path, req, parse_aux=True)
    assert t.reads() == py.reads
    np.testing.assert_array_equal(t.end, np.array([r.end() for r in py.reads], dtype=np.int32))
    for (s, e) in ((0, 1000), (100, 101), (450, 460)):
      assert t.query('chr20', s, e) == py.query('chr20', s, e)
