"""The C-ABI shared library loads and exports every symbol include/dvb.h declares (no GPU needed)."""
import ctypes
import os
import re

from deepvariant_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
  hdr = open(os.path.join(ROOT, 'include', 'dvb.h')).read()
  hdr = re.sub(r'/\*.*?\*/', '', hdr, flags=re.S)
  return sorted(set(re.findall(r'\b(dvb_[a-z0-9_]+)\s*\(', hdr)))


def test_library_is_built_and_exports_every_declared_symbol():
  assert os.path.exists(_lib.LIB_PATH), 'run __graft_entry__.build() first'
  l = ctypes.CDLL(_lib.LIB_PATH)
  declared = _declared_symbols()
  assert len(declared) >= 17
  for name in declared:
    assert hasattr(l, name), f'{name} declared in include/dvb.h but not exported'
  bound = {s[0] for s in _lib.SYMBOLS}
  assert set(declared) == bound, (set(declared) ^ bound)


def test_struct_layout_matches_header(tmp_path):
  """sizeof/offsetof as the C compiler sees include/dvb.h == the ctypes mirror."""
  import subprocess
  src = tmp_path / 'layout.c'
  src.write_text('''#include <stdio.h>
#include <stddef.h>
#include "dvb.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu %zu %zu\\n", sizeof(DvbPileupParams), sizeof(DvbBatch),
         offsetof(DvbPileupParams, random_seed), offsetof(DvbPileupParams, num_alt_channels),
         offsetof(DvbBatch, ref_stride), offsetof(DvbBatch, pair_begin), offsetof(DvbBatch, cigar));
  return 0;
}''')
  exe = tmp_path / 'layout'
  subprocess.check_call(['gcc', '-I', os.path.join(ROOT, 'include'), str(src), '-o', str(exe)])
  got = [int(x) for x in subprocess.check_output([str(exe)]).split()]
  P, B = _lib.DvbPileupParams, _lib.DvbBatch
  want = [ctypes.sizeof(P), ctypes.sizeof(B), P.random_seed.offset, P.num_alt_channels.offset,
          B.ref_stride.offset, B.pair_begin.offset, B.cigar.offset]
  assert got == want


def test_defaults_and_host_helpers_without_gpu():
  l = _lib.lib()
  p = _lib.DvbPileupParams()
  l.dvb_pileup_params_default(ctypes.byref(p))
  assert (p.width, p.height, p.reference_band_height, p.num_channels) == (221, 100, 5, 6)
  assert l.dvb_image_bytes(ctypes.byref(p)) == 132600
  import numpy as np
  t = np.zeros(100, dtype=np.int32)
  assert l.dvb_shuffle_table(100, 2101079370, 1, t.ctypes.data_as(ctypes.c_void_p)) == 0   # 1 = libstdc++
  assert t[:12].tolist() == [32, 69, 31, 60, 53, 68, 49, 39, 76, 54, 18, 82]
  assert l.dvb_shuffle_table(100, 2101079370, 0, t.ctypes.data_as(ctypes.c_void_p)) == 0   # 0 = libc++ (golden files)
  assert sorted(t.tolist()) == list(range(100)) and t[:12].tolist() != [32, 69, 31, 60, 53, 68, 49, 39, 76, 54, 18, 82]
  assert l.dvb_shuffle_table(100, 2101079370, 7, t.ctypes.data_as(ctypes.c_void_p)) != 0


def test_no_device_is_an_error_not_a_fallback():
  import torch
  if torch.cuda.is_available():
    return
  l = _lib.lib()
  p = _lib.DvbPileupParams()
  l.dvb_pileup_params_default(ctypes.byref(p))
  h = ctypes.c_void_p()
  st = l.dvb_encoder_create(ctypes.byref(p), 0, ctypes.byref(h))
  assert st == 6 and b'no CUDA device' in l.dvb_last_error()


def test_upload_phase_plan_of_the_fused_host_entry():
  """dvb_encode_classify_host uploads a large batch in phases of one classifier chunk; the plan is host arithmetic on the
  CSR arrays (no device): every read is uploaded by the first phase that references it, phases tile the images and pairs."""
  import numpy as np
  from deepvariant_b200 import synthetic
  l = _lib.lib()
  tb = synthetic.make_batch(100, 'cpu')
  cb = tb.as_ctypes()
  out = np.zeros((16, 6), dtype=np.int64)
  n = l.dvb_debug_upload_phases(ctypes.byref(cb), 32, out.ctypes.data_as(ctypes.c_void_p), 16)
  assert n == 4
  ph = out[:n]
  pair_begin = tb.tensors['pair_begin'].numpy()
  pair_read = tb.tensors['pair_read'].numpy()
  assert ph[:, 0].tolist() == [0, 32, 64, 96] and ph[:, 1].tolist() == [32, 64, 96, 100]
  assert (ph[:, 2] == pair_begin[ph[:, 0]]).all() and (ph[:, 3] == pair_begin[ph[:, 1]]).all()
  assert ph[0, 4] == 0 and (ph[1:, 4] == ph[:-1, 5]).all()                      # read ranges are contiguous
  for k in range(n):                                                             # every referenced read is on the device in time
    assert pair_read[ph[k, 2]:ph[k, 3]].max() < ph[k, 5]
  assert l.dvb_debug_upload_phases(ctypes.byref(cb), 100, out.ctypes.data_as(ctypes.c_void_p), 16) == 0    # fits one phase
  # a batch whose FIRST image references the LAST read front-loads the whole read table and stays correct
  tb.tensors['pair_read'][0] = tb.n_reads - 1
  cb = tb.as_ctypes()
  n = l.dvb_debug_upload_phases(ctypes.byref(cb), 32, out.ctypes.data_as(ctypes.c_void_p), 16)
  assert n == 4 and out[0, 5] == tb.n_reads and (out[1:n, 4] == tb.n_reads).all() and (out[1:n, 5] == tb.n_reads).all()
  tb.tensors['pair_read'][5] = tb.n_reads + 3                                    # out of range -> status, not a crash
  cb = tb.as_ctypes()
  assert l.dvb_debug_upload_phases(ctypes.byref(cb), 32, out.ctypes.data_as(ctypes.c_void_p), 16) == -1
