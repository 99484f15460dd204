"""call_variants stage driver (deepvariant/call_variants.py:766-1047) and the stage CLIs."""
import json
import os

import numpy as np
import pytest

from deepvariant_b200 import call_variants as cv
from deepvariant_b200 import protos, tfrecord

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def test_output_shard_naming():
  """call_variants.py:813-826: name.tfrecord.gz -> name-0000i-of-0000K.tfrecord.gz, K <= 16."""
  p = cv.output_shard_paths('/x/call_variants_output.tfrecord.gz', writer_threads=3)
  assert [os.path.basename(x) for x in p] == [f'call_variants_output-0000{i}-of-00003.tfrecord.gz' for i in range(3)]
  assert len(cv.output_shard_paths('/x/o.tfrecord.gz', writer_threads=64)) == 16
  assert cv.output_shard_paths('/x/o@2.tfrecord.gz') == ['/x/o-00000-of-00002.tfrecord.gz', '/x/o-00001-of-00002.tfrecord.gz']


def test_empty_input_writes_one_empty_shard(tmp_path):
  """call_variants.py:605-619 + golden.calling_examples_empty.tfrecord.gz behaviour."""
  empty = tmp_path / 'empty.tfrecord.gz'
  tfrecord.Writer(str(empty)).close()
  r = cv.call_variants(str(empty), 'random', str(tmp_path / 'out.tfrecord.gz'))
  assert r['n_examples'] == 0 and [os.path.basename(p) for p in r['paths']] == ['out-00000-of-00001.tfrecord.gz']
  assert list(tfrecord.read_records(r['paths'][0])) == []


@pytest.mark.gpu
def test_call_variants_end2end_on_reference_golden_examples(tmp_path):
  """Mirrors call_variants_test.py:91-200 (random weights; checks record count and fields), on the reference's own
  golden tf.Examples, and additionally checks the probabilities against the fp32 oracle."""
  import torch
  import cnn_oracle
  from deepvariant_b200 import modeling
  src = os.path.join(GOLDEN, 'golden.calling_examples.first3.tfrecord.gz')
  out = str(tmp_path / 'call_variants_output.tfrecord.gz')
  r = cv.call_variants(src, 'random:7', out, batch_size=2, writer_threads=2)
  assert r['n_examples'] == 3 and r['n_batches'] == 2 and len(r['paths']) == 2
  cvos = [protos.parse_call_variants_output(x) for p in r['paths'] for x in tfrecord.read_records(p, check_crc=True)]
  assert len(cvos) == 3
  examples = [protos.parse_tf_example(x) for x in tfrecord.read_records(src)]
  imgs = torch.from_numpy(np.stack([np.frombuffer(e['image/encoded'][1][0], np.uint8).reshape(100, 221, 7) for e in examples]))
  want = cnn_oracle.ReferenceModel(modeling.random_weights(7, 7)).forward(imgs).numpy()
  by_start = {protos.parse_variant(e['variant/encoded'][1][0]).start: i for i, e in enumerate(examples)}
  for variant, idx, probs in cvos:
    v = protos.parse_variant(variant)
    i = by_start[v.start]
    assert idx == [0] and len(probs) == 3 and abs(sum(probs) - 1) < 1e-9
    assert all(round(p, 10) == p for p in probs)                       # round_gls precision 10
    assert np.abs(np.array(probs) - want[i]).max() < 5e-3
    assert b'MID' in variant and b'deepvariant' in variant
    # everything but the MID entry is the example's variant
    assert v.reference_bases == protos.parse_variant(examples[i]['variant/encoded'][1][0]).reference_bases


@pytest.mark.gpu
def test_stage_clis_make_examples_then_call_variants(tmp_path):
  """run_deepvariant-style flow on a synthetic region: BAM is replaced by an in-memory reader via the Python API
  (the BAM path is exercised by tools/make_golden_fixtures.py where /root/reference exists)."""
  from test_make_examples_native import _region_fixture
  from deepvariant_b200 import make_examples_native as men
  ref, reads, cands, pic = _region_fixture()
  for c in cands:   # call_variants sets calls[0].info['MID'] (variantcall_utils.set_model_id): give every variant a call
    c.variant.raw = c.variant.serialize() + protos.f_bytes(11, protos.f_bytes(9, b'sample'))
  ex_path = str(tmp_path / 'make_examples.tfrecord-00000-of-00001.gz')
  gen = men.ExamplesGenerator(men.MakeExamplesOptions(pic_options=pic), {'main_sample': ex_path}, ref_reader=ref)
  gen.write_examples_in_region(cands, [reads], [0], 'main_sample', [0.0])
  gen.signal_shard_finished()
  from deepvariant_b200 import cli
  assert cli.call_variants(['--examples', str(tmp_path / 'make_examples.tfrecord@1.gz'), '--outfile',
                            str(tmp_path / 'cvo.tfrecord.gz'), '--checkpoint', 'random', '--writer_threads', '1']) == 0
  recs = list(tfrecord.read_records(str(tmp_path / 'cvo-00000-of-00001.tfrecord.gz')))
  assert len(recs) == 5
  _, idx, probs = protos.parse_call_variants_output(recs[3])
  assert idx == [0, 1] and abs(sum(probs) - 1) < 1e-9


def test_flags_for_calling_precedence(tmp_path):
  """apply_flags_for_calling (deepvariant/make_examples_core.py:3825-3920): command line > model.example_info.json > defaults."""
  import json
  from deepvariant_b200 import cli
  model = tmp_path / 'pacbio_model'
  model.mkdir()
  (model / 'saved_model.pb').write_bytes(b'')
  (model / 'model.example_info.json').write_text(json.dumps({
      'version': '1.10.0', 'shape': [100, 147, 10], 'channels': [1, 2, 3, 4, 5, 6, 7, 9, 10],
      'flags_for_calling': {'pileup_image_width': 147, 'sort_by_haplotypes': True, 'trim_reads_for_pileup': 'true', 'min_mapping_quality': 1,
                            'partition_size': 25000, 'alt_aligned_pileup': 'diff_channels', 'vsc_min_fraction_indels': 0.12,
                            'phase_reads': True, 'channel_list': 'BASE_CHANNELS,haplotype', 'track_ref_reads': 'true',
                            'call_small_model_examples': True, 'trained_small_model_path': '/x'}}))
  m = cli.apply_flags_for_calling({'min_mapping_quality': 7, 'task': 2}, str(model))
  assert m['pileup_image_width'] == 147 and m['sort_by_haplotypes'] is True and m['trim_reads_for_pileup'] is True
  assert m['partition_size'] == 25000 and m['alt_aligned_pileup'] == 'diff_channels' and m['channel_list'] == 'BASE_CHANNELS,haplotype'
  assert m['min_mapping_quality'] == 7 and m['task'] == 2                       # the command line wins
  assert m['pileup_image_height'] == 100 and m['min_base_quality'] == 10        # defaults
  assert m['phase_reads'] is True and m['track_ref_reads'] is True and abs(m['vsc_min_fraction_indels'] - 0.12) < 1e-12   # candidate generation
  assert sorted(m['_ignored_flags_for_calling']) == ['call_small_model_examples', 'trained_small_model_path']   # stages not built here
  assert cli.model_example_info_json_path(str(model / 'model.ckpt')) == str(model / 'model.example_info.json')   # json next to a ckpt
  ck = tmp_path / 'ckpt_dir'
  ck.mkdir()
  (ck / 'weights.example_info.json').write_text('{"flags_for_calling": {"pileup_image_width": 199}}')
  assert cli.apply_flags_for_calling({}, str(ck / 'model.ckpt'))['pileup_image_width'] == 199
  assert cli.apply_flags_for_calling({}, '')['pileup_image_width'] == 221
