"""Host-side mirror of the reference's call_variants stage around the CUDA classifier.

  GpuCnn                 the SavedModel call of predict_step (deepvariant/call_variants.py:904-932):
                         uint8 pileup images in, float32 genotype probabilities (p00, p0x, pxx) out
  round_gls              deepvariant/call_variants.py:248-285
  create_cvo             deepvariant/call_variants.py:353-399 (_create_cvo_proto, MID="deepvariant")
  call_variants          deepvariant/call_variants.py:766-1047 (examples TFRecords -> CVO TFRecords)

There is no CPU path: GpuCnn raises when libdvb.so or a B200 is missing.
"""
from __future__ import annotations

import ctypes as C
import struct
from typing import Iterable, List, Optional, Sequence, Tuple

import numpy as np

from deepvariant_b200 import _lib, modeling, protos

DEEP_VARIANT_MODEL_ID = 'deepvariant'  # call_variants.py:85
_GL_PRECISION = 10                       # call_variants.py:79
_DEFAULT_BATCH = 1024                    # --batch_size default, call_variants.py:115-117


class GpuCnn:
  """Owns one DvbCnn handle (per device)."""

  def __init__(self, weights: modeling.ModelWeights, image_shape: Sequence[int], device: int = 0,
               max_batch: int = 2048, precision: int = 0):
    self._lib = _lib.lib()
    h, w, c = [int(x) for x in image_shape]
    if c != weights.in_channels:
      raise ValueError(f'model expects {weights.in_channels} channels, images have {c}')
    blob = modeling.pack_weights(weights, precision)
    self.precision = precision
    self.shape = (h, w, c)
    self.device = device
    self.max_batch = max_batch
    handle = C.c_void_p()
    buf = (C.c_char * len(blob)).from_buffer_copy(blob)
    _lib.check(self._lib.dvb_cnn_create(C.cast(buf, C.c_void_p), len(blob), h, w, c, max_batch, precision, device,
                                        C.byref(handle)))
    self._h = handle
    self.flops_per_image = float(self._lib.dvb_cnn_flops_per_image(self._h))

  @classmethod
  def random_init(cls, image_shape: Sequence[int], device: int = 0, max_batch: int = 2048, seed: int = 0,
                  precision: int = 0) -> 'GpuCnn':
    """Random-init weights of the right architecture (no checkpoints ship with the reference)."""
    return cls(modeling.random_weights(int(image_shape[2]), seed), image_shape, device, min(max_batch, 4096), precision)

  def forward_device(self, images, probs, stream=None) -> None:
    """images: torch uint8 [n, H, W, C] on the device; probs: torch float32 [n, 3].  Asynchronous."""
    n = int(images.shape[0])
    sp = C.c_void_p(stream.cuda_stream) if stream is not None else C.c_void_p(0)
    _lib.check(self._lib.dvb_cnn_forward_device(self._h, C.c_void_p(images.data_ptr()), n, C.c_void_p(probs.data_ptr()), sp))

  def forward_host(self, images: np.ndarray) -> np.ndarray:
    images = np.ascontiguousarray(images, dtype=np.uint8)
    if images.shape[1:] != self.shape:
      raise ValueError(f'images {images.shape[1:]} != model input {self.shape}')
    probs = np.empty((images.shape[0], 3), dtype=np.float32)
    _lib.check(self._lib.dvb_cnn_forward_host(self._h, images.ctypes.data_as(C.c_void_p), images.shape[0],
                                              probs.ctypes.data_as(C.c_void_p)))
    return probs

  def debug_tensor(self, name: str, n: int) -> np.ndarray:
    h, w, c = C.c_int32(), C.c_int32(), C.c_int32()
    _lib.check(self._lib.dvb_cnn_debug_tensor(self._h, name.encode(), n, None, C.byref(h), C.byref(w), C.byref(c)))
    out = np.empty((n, h.value, w.value, c.value), dtype=np.float32)
    _lib.check(self._lib.dvb_cnn_debug_tensor(self._h, name.encode(), n, out.ctypes.data_as(C.c_void_p), C.byref(h),
                                              C.byref(w), C.byref(c)))
    return out

  @property
  def launch_count(self) -> int:
    return int(self._lib.dvb_cnn_launch_count(self._h))

  def roofline(self, ms_per_step: float, images_per_step: int, peaks: dict) -> dict:
    """Tensor-core roofline of the classifier: algorithmic conv FLOPs / device time of the CNN part."""
    tf = images_per_step * self.flops_per_image / (ms_per_step * 1e-3) / 1e12
    return {'bound': 'tensor', 'kernel': 'classifier forward = 94 tcgen05 implicit-GEMM convolutions (stem_conv1 / conv_rows / conv_gemm_persistent / conv_gemm_pair / conv_gemm kernels) + 12 pools + tail',
            'achieved': tf, 'peak': peaks['tflops_sustained'], 'unit': 'TFLOP/s', 'frac': tf / peaks['tflops_sustained'],
            'traffic': None, 'peak_source': peaks['source'] + ', sustained cuBLAS bf16',
            'flops_per_image': self.flops_per_image, 'ms_cnn_per_step': ms_per_step}

  def close(self):
    if getattr(self, '_h', None):
      self._lib.dvb_cnn_destroy(self._h)
      self._h = None

  def __del__(self):
    try:
      self.close()
    except Exception:  # pylint: disable=broad-except
      pass


def round_gls(gls: Sequence[float], precision: Optional[int] = None) -> List[float]:
  """deepvariant/call_variants.py:248-285."""
  gls = [float(g) for g in gls]
  if abs(sum(gls) - 1) > 1e-6:
    raise ValueError('Invalid genotype likelihoods do not sum to one: sum({}) = {}'.format(gls, sum(gls)))
  if precision is None:
    return gls
  min_ix = 0
  min_gl = gls[0]
  for ix, gl in enumerate(gls):
    if gl < min_gl:
      min_gl = gl
      min_ix = ix
  rounded_gls = [round(gl, precision) for gl in gls]
  rounded_gls[min_ix] = max(0.0, round(1 - sum(rounded_gls[:min_ix] + rounded_gls[min_ix + 1:]), precision))
  return rounded_gls


def _set_model_id(variant_encoded: bytes, model_id: str) -> bytes:
  """variantcall_utils.set_model_id (third_party/nucleus/util/variantcall_utils.py:235): sets
  calls[0].info['MID'] = [string_value model_id] inside a serialized Variant.
  Variant.calls = 11; VariantCall.info = map<string, ListValue> field 2; Value.string_value = 3."""
  entry = protos.f_bytes(1, b'MID') + protos.f_bytes(2, protos.f_bytes(1, protos.f_bytes(3, model_id.encode())))
  out = bytearray()
  done = False
  for fn, wt, val, raw in protos.iter_fields(variant_encoded):
    if fn == 11 and not done:
      call = bytearray()
      for f2, w2, v2, raw2 in protos.iter_fields(bytes(val)):
        if f2 == 2:
          key = b''
          for f3, w3, v3, _ in protos.iter_fields(bytes(v2)):
            if f3 == 1:
              key = bytes(v3)
          if key == b'MID':
            continue
        call += raw2
      call += protos.f_bytes(2, entry)
      out += protos.f_bytes(11, bytes(call))
      done = True
    else:
      out += raw
  if not done:
    raise IndexError('variant has no calls')  # reference: variant.calls[0] raises
  return bytes(out)


def create_cvo(variant_encoded: bytes, gls: Sequence[float], alt_allele_indices_encoded: bytes) -> bytes:
  """_create_cvo_proto (call_variants.py:353-399) without debug info -> serialized CallVariantsOutput."""
  variant = _set_model_id(variant_encoded, DEEP_VARIANT_MODEL_ID)
  out = bytearray()
  out += protos.f_bytes(1, variant)
  out += protos.f_bytes(2, alt_allele_indices_encoded)
  out += protos.f_bytes(3, struct.pack('<3d', *[float(g) for g in gls]))
  return bytes(out)


# ---- the stage driver (deepvariant/call_variants.py:766-1047) -------------------------------------

_MAX_WRITER_THREADS = 16   # call_variants.py:82


def load_weights(checkpoint_path: str, in_channels: int) -> modeling.ModelWeights:
  """Model weights for --checkpoint: a TensorFlow SavedModel directory or checkpoint prefix (tf_checkpoint.py reads the tensor
  bundle without TensorFlow), a .npz written by modeling.save_npz, or 'random[:seed]' (architecture-correct random init; no
  Inception weights ship with the reference)."""
  import re
  import sys
  m = re.fullmatch(r'random(?::(\d+))?', checkpoint_path)   # the exact token only: 'random_forest_model/' is a path, not a request for noise
  if m:
    print('call_variants: WARNING - --checkpoint random: architecture-correct RANDOM weights; the genotype calls are noise '
          '(smoke runs and benchmarks only)', file=sys.stderr)
    return modeling.random_weights(in_channels, int(m.group(1) or 0))
  if checkpoint_path.endswith('.npz'):
    w = modeling.load_npz(checkpoint_path)
    if w.in_channels != in_channels:
      raise ValueError(f'model has {w.in_channels} input channels, examples have {in_channels}')
    return w
  from deepvariant_b200 import tf_checkpoint
  if tf_checkpoint.is_tf_checkpoint(checkpoint_path):      # a released SavedModel directory or a model.ckpt / ckpt-N prefix
    return tf_checkpoint.load_inception_weights(checkpoint_path, in_channels)
  raise NotImplementedError(f'unsupported checkpoint format: {checkpoint_path} (a TensorFlow SavedModel directory / checkpoint prefix, '
                            'a .npz from modeling.save_npz, or random[:seed])')


def output_shard_paths(output_file: str, writer_threads: int = 0) -> List[str]:
  """Dynamic output sharding (call_variants.py:800-826): with a GPU present K = min(cpu_count, 16) writers
  unless the name is already sharded; name.tfrecord.gz -> name-0000i-of-0000K.tfrecord.gz."""
  import os
  from deepvariant_b200 import tfrecord
  if tfrecord.is_sharded_spec(output_file) or '-of-' in os.path.basename(output_file):
    return tfrecord.shard_paths(output_file)
  k = writer_threads if writer_threads else (os.cpu_count() or 1)
  k = max(1, min(k, _MAX_WRITER_THREADS))
  return tfrecord.shard_paths(output_file.replace('.tfrecord.gz', f'@{k}.tfrecord.gz'))


def write_empty_output_file(output_file: str) -> List[str]:
  """write_empty_output_file (call_variants.py:605-619): one empty shard."""
  from deepvariant_b200 import tfrecord
  paths = tfrecord.shard_paths(output_file.replace('.tfrecord.gz', '@1.tfrecord.gz'))
  for p in paths:
    tfrecord.Writer(p).close()
  return paths


def check_example_info(examples_info: dict, model_info: Optional[dict]) -> None:
  """The reference refuses to classify examples whose shape or channel list differs from the model's
  model.example_info.json (deepvariant/call_variants.py:724-763)."""
  if not model_info:
    return
  if [int(x) for x in model_info.get('shape', examples_info['shape'])] != [int(x) for x in examples_info['shape']]:
    raise ValueError(f'examples have shape {examples_info["shape"]}, the model was trained on {model_info["shape"]}')
  if 'channels' in model_info and 'channels' in examples_info and \
      [int(x) for x in model_info['channels']] != [int(x) for x in examples_info['channels']]:
    raise ValueError(f'examples have channels {examples_info["channels"]}, the model was trained on {model_info["channels"]}')


def model_example_info(checkpoint_path: str) -> Optional[dict]:
  """model.example_info.json beside a checkpoint / inside a SavedModel directory, or None."""
  import json
  import os
  for cand in (os.path.join(checkpoint_path, 'example_info.json'), os.path.join(checkpoint_path, 'model.example_info.json'),
               os.path.join(os.path.dirname(checkpoint_path), 'example_info.json'),
               os.path.join(os.path.dirname(checkpoint_path), 'model.example_info.json'), checkpoint_path + '.example_info.json'):
    if os.path.isfile(cand):
      return json.load(open(cand))
  return None


def call_variants_from_stream(shm_prefix: str, num_input_shards: int, checkpoint_path: str, output_file: str, image_shape: Optional[Sequence[int]] = None,
                              writer_threads: int = 0, device: int = 0, precision: int = 1, net=None) -> dict:
  """call_variants --stream_examples (deepvariant/call_variants.py:504-538 + stream_examples_kernel.cc): the examples come from the
  make_examples processes' shared-memory buffers (stream_examples.StreamConsumer) instead of TFRecord files; one classifier batch per
  drained buffer.  The image shape is the model's (model.example_info.json) unless given.  `net` (anything with forward_host(images) ->
  float32[n, 3]) replaces the GPU classifier in tests."""
  from deepvariant_b200 import records, stream_examples
  if image_shape is None:
    info = model_example_info(checkpoint_path)
    if not info:
      raise ValueError('--stream_examples needs the image shape: no example_info.json beside the checkpoint')
    image_shape = info['shape']
  shape = [int(x) for x in image_shape]
  own = net is None
  if own:
    net = GpuCnn(load_weights(checkpoint_path, shape[2]), shape, device=device, max_batch=2048, precision=precision)
  consumer = stream_examples.StreamConsumer(shm_prefix, num_input_shards, tuple(shape))
  out_paths = output_shard_paths(output_file, writer_threads)
  writers = [records.NativeCvoWriter(p, _GL_PRECISION) for p in out_paths]
  n_examples = n_batches = 0
  try:
    while True:
      batch = consumer.next()
      if batch is None:
        break
      images, variants, alts = batch
      probs = net.forward_host(images)
      writers[n_batches % len(writers)].write_batch(records.BatchMeta.from_lists(variants, alts), probs)
      n_examples += len(variants)
      n_batches += 1
  finally:
    for w in writers:
      w.close()
    consumer.close()
    if own:
      net.close()
  return {'n_examples': n_examples, 'n_batches': n_batches, 'paths': out_paths}


def call_variants(examples_filename: str, checkpoint_path: str, output_file: str, batch_size: int = _DEFAULT_BATCH,
                  writer_threads: int = 0, device: int = 0, max_batches: Optional[int] = None, reader_threads: int = 0,
                  precision: int = 1) -> dict:
  """examples TFRecords -> CallVariantsOutput TFRecords (main loop of deepvariant/call_variants.py:766-1047).  Returns
  {'n_examples', 'n_batches', 'paths'}.  precision: 1 = split-fp16 x3 (1e-5 of fp32, default), 0 = fp16 operands (3x faster).

  The records never become Python objects: records.NativeExamplesReader (C++ threads: gunzip, TFRecord framing + CRC,
  tf.Example field extraction, tf.data's interleave order) fills a pinned uint8 batch buffer, the classifier runs on it,
  and records.NativeCvoWriter (one C++ thread per output shard, batches dealt round-robin as the reference deals them
  to its writer processes, call_variants.py:1037-1047) rounds, builds and compresses the CallVariantsOutput protos.
  While the GPU works on batch k the reader fills batch k + 1 into the other half of the pinned buffer."""
  import json
  import os
  import torch
  from deepvariant_b200 import records, tfrecord
  paths_in = tfrecord.resolve_input_paths(examples_filename)
  reader = records.NativeExamplesReader(paths_in, threads=reader_threads)
  try:
    shape, image_bytes = reader.shape()
    if image_bytes == 0:
      return {'n_examples': 0, 'n_batches': 0, 'paths': write_empty_output_file(output_file)}
    if image_bytes != shape[0] * shape[1] * shape[2]:
      raise ValueError(f'image/encoded has {image_bytes} bytes, image/shape says {shape}')
    info_path = paths_in[0] + '.example_info.json'
    if os.path.exists(info_path):
      info = json.load(open(info_path))
      if [int(x) for x in info['shape']] != shape:
        raise ValueError(f'example_info.json shape {info["shape"]} != example image/shape {shape}')
      check_example_info(info, model_example_info(checkpoint_path))
    weights = load_weights(checkpoint_path, shape[2])
    # precision 1 (split-fp16 x3, within 1e-5 of fp32) is the default of the VCF-producing path: the probabilities are rounded
    # to 10 decimals and feed GQ / QUAL / PL; precision 0 (fp16 operands, 3e-4) is the throughput mode bench.py headlines.
    net = GpuCnn(weights, shape, device=device, max_batch=min(batch_size, 2048), precision=precision)
    out_paths = output_shard_paths(output_file, writer_threads)
    writers = [records.NativeCvoWriter(p, _GL_PRECISION) for p in out_paths]
    pinned = torch.empty((2, batch_size, image_bytes), dtype=torch.uint8).pin_memory()
    staging = pinned.numpy()
    dev = torch.device('cuda', device)
    images_dev = torch.empty((batch_size, image_bytes), dtype=torch.uint8, device=dev)
    probs_dev = torch.empty((batch_size, 3), dtype=torch.float32, device=dev)
    probs_host = torch.empty((batch_size, 3), dtype=torch.float32).pin_memory()
    stream = torch.cuda.current_stream(dev)

    n_examples = n_batches = 0
    half = 0
    meta = reader.next_into(staging[half])
    while meta is not None:
      n = meta.n
      images_dev[:n].copy_(pinned[half, :n], non_blocking=True)
      net.forward_device(images_dev[:n].view((n,) + tuple(shape)), probs_dev[:n], stream=stream)
      probs_host[:n].copy_(probs_dev[:n], non_blocking=True)
      done = max_batches is not None and n_batches + 1 >= max_batches
      nxt = None if done else reader.next_into(staging[half ^ 1])   # host work of batch k + 1 under the GPU work of batch k
      stream.synchronize()
      writers[n_batches % len(writers)].write_batch(meta, probs_host[:n].numpy())
      n_examples += n
      n_batches += 1
      meta = nxt
      half ^= 1
    for w in writers:
      w.close()
    net.close()
    return {'n_examples': n_examples, 'n_batches': n_batches, 'paths': out_paths}
  finally:
    reader.close()
