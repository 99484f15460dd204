"""The shared-memory example stream (deepvariant_b200/stream_examples.py, csrc/dvb_stream.cu) - the reference's
make_examples --stream_examples -> call_variants boundary (stream_examples.cc:94-176, stream_examples_kernel.cc:166-240,
fast_pipeline.cc:125-165): producers in child PROCESSES, the consumer here, buffers small enough to force the hand-over in the
middle of a region; the byte layout of a buffer is checked against the reference's record format by hand."""
import os
import struct
import subprocess
import sys

import numpy as np
import pytest

from deepvariant_b200 import stream_examples as se

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHAPE = (6, 11, 3)
IMAGE_BYTES = int(np.prod(SHAPE))

PRODUCER = r'''
import sys
import numpy as np
sys.path.insert(0, {root!r})
from deepvariant_b200 import stream_examples as se
prefix, shard, regions = sys.argv[1], int(sys.argv[2]), [int(x) for x in sys.argv[3].split(',')]
p = se.StreamProducer(prefix, shard)
k = 0
for n in regions:                       # one StartStreaming / EndStreaming pair per region, as WriteExamplesInRegion does
  p.start_streaming()
  for _ in range(n):
    rng = np.random.RandomState(1000 * shard + k)
    p.stream_example(bytes([10, 1, k % 3]), b'variant-%d-%d' % (shard, k) + bytes(rng.randint(0, 256, k % 7)), rng.randint(0, 256, {shape!r}).astype(np.uint8))
    k += 1
  p.end_streaming(n > 0)
p.signal_shard_finished()
'''


def _expected(shard, k):
  rng = np.random.RandomState(1000 * shard + k)
  variant = b'variant-%d-%d' % (shard, k) + bytes(rng.randint(0, 256, k % 7))
  return bytes([10, 1, k % 3]), variant, rng.randint(0, 256, SHAPE).astype(np.uint8)


@pytest.mark.parametrize('regions_per_shard', [['3,0,9,1', '5,4'], ['0', '7'], ['1,1,1', '0,0', '12']])
def test_producers_in_other_processes_reach_the_consumer(regions_per_shard):
  prefix = f'dvbtest_{os.getpid()}'
  n_shards = len(regions_per_shard)
  record = IMAGE_BYTES + 3 + 30 + 16
  orch = se.StreamOrchestrator(prefix, n_shards, buffer_size=3 * record)     # at most ~3 examples per hand-over
  try:
    procs = [subprocess.Popen([sys.executable, '-c', PRODUCER.format(root=ROOT, shape=SHAPE), prefix, str(s), regions_per_shard[s]]) for s in range(n_shards)]
    consumer = se.StreamConsumer(prefix, n_shards, SHAPE)
    got = {s: [] for s in range(n_shards)}
    batches = 0
    while True:
      out = consumer.next()
      if out is None:
        break
      images, variants, alts = out
      batches += 1
      assert 1 <= len(variants) <= 3
      for img, v, a in zip(images, variants, alts):
        shard = int(v.split(b'-')[1])
        got[shard].append((a, v, img.copy()))
    for p in procs:
      assert p.wait(timeout=60) == 0
    consumer.close()
    for s in range(n_shards):
      total = sum(int(x) for x in regions_per_shard[s].split(','))
      assert len(got[s]) == total                                # nothing lost, nothing duplicated, order kept within a shard
      for k, (a, v, img) in enumerate(got[s]):
        ea, ev, ei = _expected(s, k)
        assert a == ea and v == ev and np.array_equal(img, ei)
    assert batches >= sum(sum(int(x) for x in r.split(',')) for r in regions_per_shard) / 3
  finally:
    orch.remove()
  assert not [f for f in os.listdir('/dev/shm') if prefix in f]   # shm object and the three semaphores of every shard are gone


def test_buffer_bytes_are_the_reference_record_format():
  """{int32 len, alt indices}{int32 len, variant}{int32 len, image} ... int32 0, native-endian (stream_examples.cc:94-141)."""
  prefix = f'dvbfmt_{os.getpid()}'
  orch = se.StreamOrchestrator(prefix, 1, buffer_size=4096)
  try:
    p = se.StreamProducer(prefix, 0)
    img = np.arange(IMAGE_BYTES, dtype=np.uint8).reshape(SHAPE)
    p.start_streaming()
    p.stream_example(b'\x0a\x01\x00', b'VARIANT', img)
    p.stream_example(b'\x0a\x02\x00\x01', b'V2', img[::-1])
    p.end_streaming(True)
    raw = open(f'/dev/shm/{prefix}_shm_0', 'rb').read()
    want = (struct.pack('=i', 3) + b'\x0a\x01\x00' + struct.pack('=i', 7) + b'VARIANT' + struct.pack('=i', IMAGE_BYTES) + img.tobytes() +
            struct.pack('=i', 4) + b'\x0a\x02\x00\x01' + struct.pack('=i', 2) + b'V2' + struct.pack('=i', IMAGE_BYTES) + img[::-1].tobytes() + struct.pack('=i', 0))
    assert raw[:len(want)] == want
    c = se.StreamConsumer(prefix, 1, SHAPE)
    images, variants, alts = c.next()
    assert variants == [b'VARIANT', b'V2'] and alts == [b'\x0a\x01\x00', b'\x0a\x02\x00\x01'] and np.array_equal(images[1], img[::-1])
    p.signal_shard_finished()
    assert c.next() is None
    with pytest.raises(Exception):
      p.stream_example(b'a', b'v', np.zeros(5000, dtype=np.uint8))   # an example larger than the buffer can never be sent
    c.close()
    p.close()
  finally:
    orch.remove()


# ---- the stage CLIs over the stream: make_examples --stream_examples (two task processes) -> call_variants --stream_examples -----------

MAKE_EXAMPLES_TASK = r'''
import sys
sys.path.insert(0, {root!r}); sys.path.insert(0, {root!r} + '/tests')
import test_candidates as tc
from deepvariant_b200 import cli, make_examples_native as men, pileup_image as pi
# no GPU in the CPU suite: the encoder handle is the oracle-backed stand-in of the other CLI tests
men.ExamplesGenerator._gpu = lambda self: tc.OracleEncoder(pi.to_params(self.options.pic_options, height=self.pileup_image_height))
sys.exit(cli.make_examples(sys.argv[1:]))
'''


class _FakeNet:
  """Deterministic stand-in classifier: probabilities from the image bytes (so that any mix-up of images and records shows)."""

  def forward_host(self, images):
    s = images.reshape(len(images), -1).astype(np.int64).sum(1)
    p = np.stack([(s % 7 + 1), (s % 5 + 1), (s % 3 + 1)], 1).astype(np.float64)
    return (p / p.sum(1, keepdims=True)).astype(np.float32)


def test_stage_clis_over_the_stream_equal_the_staged_flow(tmp_path):
  import test_candidates as tc
  from deepvariant_b200 import call_variants as cv, records, tfrecord, protos
  fa, bam_path, _, _ = tc._planted_case(tmp_path)   # pylint: disable=protected-access
  common = ['--mode', 'calling', '--ref', fa, '--reads', bam_path, '--channel_list', 'BASE_CHANNELS,insert_size', '--regions', 'chr20:1001-5000', '--norealign_reads']
  # staged: two tasks write tf.Examples; the same stand-in classifier over them
  staged = {}
  for task in range(2):
    subprocess.check_call([sys.executable, '-c', MAKE_EXAMPLES_TASK.format(root=ROOT), *common, '--examples', str(tmp_path / 'ex.tfrecord@2.gz'), '--task', str(task)])
    for r in tfrecord.read_records(str(tmp_path / f'ex.tfrecord-0000{task}-of-00002.gz')):
      e = protos.parse_tf_example(r)
      img = np.frombuffer(e['image/encoded'][1][0], dtype=np.uint8).reshape(e['image/shape'][1])
      staged[(e['variant/encoded'][1][0], e['alt_allele_indices/encoded'][1][0])] = _FakeNet().forward_host(img[None])[0]
  assert len(staged) == 4
  # streamed: the same two tasks as producer processes, call_variants_from_stream as the consumer
  prefix = f'dvbcli_{os.getpid()}'
  orch = se.StreamOrchestrator(prefix, 2, buffer_size=2 * (100 * 221 * 7 + 4096))
  try:
    procs = [subprocess.Popen([sys.executable, '-c', MAKE_EXAMPLES_TASK.format(root=ROOT), *common, '--examples', str(tmp_path / 'unused@2.gz'), '--task', str(task),
                               '--stream_examples', '--shm_prefix', prefix]) for task in range(2)]
    out = str(tmp_path / 'cvo.tfrecord.gz')
    r = cv.call_variants_from_stream(prefix, 2, 'unused', out, image_shape=[100, 221, 7], writer_threads=1, net=_FakeNet())
    for p in procs:
      assert p.wait(timeout=120) == 0
  finally:
    orch.remove()
  assert r['n_examples'] == 4
  got = {}
  for path in r['paths']:
    for rec in tfrecord.read_records(path):
      variant, idx, probs = protos.parse_call_variants_output(rec)
      got[(variant, protos.encode_alt_allele_indices(idx))] = np.array(probs)
  assert set(k[1] for k in got) == set(k[1] for k in staged)
  by_alt_and_start = lambda d: {(protos.parse_variant(k[0]).start, k[1]): v for k, v in d.items()}   # the CVO's variant gains the MID call info
  a, b = by_alt_and_start(got), by_alt_and_start(staged)
  assert set(a) == set(b)
  for k in a:
    np.testing.assert_allclose(a[k], np.round(b[k].astype(np.float64), 10), atol=1e-9)
  assert not os.path.exists(str(tmp_path / 'unused-00000-of-00002.gz'))   # nothing was written beside the stream
