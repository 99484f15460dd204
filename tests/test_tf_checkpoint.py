"""TensorFlow tensor-bundle reader (deepvariant_b200/tf_checkpoint.py): `call_variants --checkpoint <SavedModel dir | ckpt prefix>`.
No released model ships with the reference, so the reader is exercised on bundles written by its own writer (same published layout):
table blocks with prefix compression and restarts, several data blocks, checksums, the object graph, layer-name matching."""
import os
import struct

import numpy as np
import pytest

from deepvariant_b200 import call_variants as cv, modeling, tf_checkpoint as tfc


def test_table_round_trip_with_prefix_compression_and_many_blocks(tmp_path):
  items = {b'': b'header'}
  for i in range(700):
    items[f'layer_with_weights-{i}/kernel/.ATTRIBUTES/VARIABLE_VALUE'.encode()] = os.urandom(i % 40)
  path = str(tmp_path / 't.index')
  tfc.write_table(path, items, block_size=512)
  assert tfc.read_table(path) == items
  raw = bytearray(open(path, 'rb').read())
  raw[10] ^= 0xff
  open(path, 'wb').write(raw)
  with pytest.raises(ValueError, match='checksum'):
    tfc.read_table(path)
  open(path, 'wb').write(b'not a table' * 10)
  with pytest.raises(ValueError, match='magic'):
    tfc.read_table(path)


def test_snappy_block_format():
  # literal "abcd", copy(offset 4, length 8) overlapping, literal "xyz"
  comp = bytes([15]) + bytes([3 << 2]) + b'abcd' + bytes([((8 - 4) << 2) | 1, 4]) + bytes([2 << 2]) + b'xyz'
  assert tfc.snappy_uncompress(comp) == b'abcdabcdabcdxyz'
  with pytest.raises(ValueError):
    tfc.snappy_uncompress(bytes([5]) + bytes([1, 9]))


def test_bundle_round_trip_dtypes_and_checksums(tmp_path):
  prefix = str(tmp_path / 'ckpt-1')
  t = {'a/kernel': np.arange(24, dtype=np.float32).reshape(2, 3, 4), 'b': np.array([1, 2, 3], dtype=np.int64), 'c': np.float32(2.5).reshape(())}
  tfc.write_bundle(prefix, t, {'conv2d/kernel': 'a/kernel'})
  b = tfc.Bundle(prefix)
  assert b.keys() == sorted(list(t) + [tfc.OBJECT_GRAPH_KEY])
  for k, v in t.items():
    np.testing.assert_array_equal(b.tensor(k), v)
    assert b.tensor(k).dtype == v.dtype
  assert b.object_graph_names() == {'conv2d/kernel': 'a/kernel'}
  data = bytearray(open(prefix + '.data-00000-of-00001', 'rb').read())
  data[-1] ^= 1
  open(prefix + '.data-00000-of-00001', 'wb').write(data)
  with pytest.raises(ValueError, match='checksum mismatch for tensor'):
    tfc.Bundle(prefix).tensor(sorted(t)[-1])


@pytest.mark.parametrize('first_index', [0, 94])
def test_inception_weights_round_trip_as_saved_model_and_as_ckpt(tmp_path, first_index):
  """first_index 94: the layer counter did not start at zero (a second model built in the same process, as
  keras_modeling.inceptionv3 does when it first builds the imagenet backbone)."""
  w = modeling.random_weights(7, seed=5)
  saved = tmp_path / 'wgs_model'
  (saved / 'variables').mkdir(parents=True)
  (saved / 'saved_model.pb').write_bytes(b'')
  tfc.save_inception_checkpoint(str(saved / 'variables' / 'variables'), w, first_layer_index=first_index)
  for path in (str(saved), str(saved / 'variables' / 'variables'), str(saved / 'variables' / 'variables.index')):
    got = cv.load_weights(path, 7)
    assert got.in_channels == 7 and len(got.convs) == 94
    for a, b in zip(got.convs, w.convs):
      for f in ('kernel', 'beta', 'moving_mean', 'moving_variance'):
        np.testing.assert_array_equal(getattr(a, f), getattr(b, f))
    np.testing.assert_array_equal(got.dense_kernel, w.dense_kernel)
    np.testing.assert_array_equal(got.dense_bias, w.dense_bias)
  assert modeling.pack_weights(got) == modeling.pack_weights(w)            # the classifier would get the very same blob
  with pytest.raises(ValueError, match='input channels'):
    cv.load_weights(str(saved), 6)


def test_topology_mismatch_is_reported(tmp_path):
  w = modeling.random_weights(7, seed=1)
  w.convs[10], w.convs[11] = w.convs[11], w.convs[10]                     # wrong creation order
  prefix = str(tmp_path / 'model.ckpt')
  tfc.save_inception_checkpoint(prefix, w)
  with pytest.raises(ValueError, match='topology expects'):
    tfc.load_inception_weights(prefix)
  with pytest.raises(NotImplementedError):
    cv.load_weights(str(tmp_path / 'nothing_here'), 7)
