"""Fused make_examples + call_variants (deepvariant_b200/fused.py): regions -> packed reads -> dvb_encode_classify_host ->
CallVariantsOutput shards, against the staged flow (tf.Example files between the stages)."""
import os

import numpy as np
import pytest

import oracle_lib
from deepvariant_b200 import fused, pileup_image as pi, synthetic
from subbatch_util import take_images


def _params():
  o = pi.default_options()
  o.channels = list(pi.PILEUP_CHANNELS_WITH_INSERT_SIZE)
  return pi.to_params(o)


def test_concat_packed_is_the_batch_of_all_images():
  """Stacking region batches (read tables appended, pair / CSR indices shifted) must encode to the images of the parts, in order;
  parts that share no reads, parts with zero-read images, a single part."""
  params = _params()
  whole = synthetic.make_batch(40, 'cpu').to_packed()
  parts = [take_images(whole, idx) for idx in (np.arange(0, 7), np.arange(7, 8), np.arange(8, 29), np.arange(29, 40))]
  other = synthetic.make_batch(5, 'cpu', chunk=3).to_packed()          # an independent read table
  cat = fused.concat_packed(parts + [other])
  assert cat.n_images == 45 and cat.n_reads == 4 * whole.n_reads + other.n_reads
  want = np.concatenate([oracle_lib.encode_batch(params, whole), oracle_lib.encode_batch(params, other)])
  np.testing.assert_array_equal(oracle_lib.encode_batch(params, cat), want)
  assert fused.concat_packed([parts[0]]) is parts[0]
  with pytest.raises(ValueError):
    fused.concat_packed([])


@pytest.mark.gpu
def test_fused_flow_writes_the_same_call_variants_outputs_and_vcf_as_the_staged_flow(tmp_path):
  """run_deepvariant (fused, the default) against run_deepvariant --staged on the planted genome: the same CallVariantsOutput
  records (variant, alt_allele_indices, rounded likelihoods) and a byte-identical VCF; with two shards the tasks run as two
  processes in parallel."""
  import test_candidates as tc
  from deepvariant_b200 import cli, protos, tfrecord
  fa, bam_path, genome, sites = tc._planted_case(tmp_path)
  common = ['--model_type', 'WGS', '--ref', fa, '--reads', bam_path, '--regions', 'chr20:1001-5000', '--customized_model', 'random:3']
  outs = {}
  for name, extra in (('staged', ['--staged']), ('fused', []), ('fused2', ['--num_shards', '2'])):
    d = str(tmp_path / name)
    vcf = os.path.join(d, 'out.vcf')
    assert cli.run_deepvariant(common + ['--output_dir', d, '--output_vcf', vcf] + extra) == 0
    cvos = sorted(r for p in tfrecord.resolve_input_paths(os.path.join(d, 'call_variants_output.tfrecord.gz')) for r in tfrecord.read_records(p))
    outs[name] = (open(vcf).read(), cvos)
  assert len(outs['staged'][1]) >= len(sites)
  assert outs['fused'][1] == outs['staged'][1]
  assert outs['fused2'][1] == outs['staged'][1]
  assert outs['fused'][0] == outs['staged'][0] and outs['fused2'][0] == outs['staged'][0]
  assert not os.path.exists(os.path.join(str(tmp_path / 'fused'), 'make_examples.tfrecord-00000-of-00001.gz'))   # no tf.Example round trip
