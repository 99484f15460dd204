"""ctypes access to the CPU oracle (oracle/dvb_oracle.cc) — TEST INFRASTRUCTURE ONLY.

Mirrors the call surface of deepvariant_b200.pileup_image.PileupImageEncoderNative so the
same test bodies can run against the oracle (CPU, `-m "not gpu"`) and against the CUDA
path (`-m gpu`).
"""
from __future__ import annotations

import ctypes as C
import dataclasses
import os
import subprocess
from typing import Optional, Sequence

import numpy as np

from deepvariant_b200 import _lib, packing
from deepvariant_b200 import pileup_image as pi

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_SO = os.path.join(_ROOT, 'oracle', '_build', 'libdvb_oracle.so')
_oracle = None


def build_oracle() -> str:
  src = os.path.join(_ROOT, 'oracle', 'dvb_oracle.cc')
  hdr = os.path.join(_ROOT, 'include', 'dvb.h')
  if (not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src)
      or os.path.getmtime(_SO) < os.path.getmtime(hdr)):
    subprocess.check_call(['make', '-s', '-C', os.path.join(_ROOT, 'oracle')])
  return _SO


def oracle():
  global _oracle
  if _oracle is None:
    l = C.CDLL(build_oracle())
    l.dvb_oracle_last_error.restype = C.c_char_p
    l.dvb_oracle_encode_batch.restype = C.c_int
    l.dvb_oracle_encode_batch.argtypes = [C.POINTER(_lib.DvbPileupParams), C.POINTER(_lib.DvbBatch), C.c_void_p]
    l.dvb_oracle_encode_read.restype = C.c_int
    l.dvb_oracle_encode_read.argtypes = [C.POINTER(_lib.DvbPileupParams), C.POINTER(_lib.DvbBatch), C.c_int32,
                                         C.c_int64, C.c_void_p, C.POINTER(C.c_int32)]
    l.dvb_oracle_encode_reference.restype = C.c_int
    l.dvb_oracle_encode_reference.argtypes = [C.POINTER(_lib.DvbPileupParams), C.c_void_p, C.c_void_p]
    l.dvb_oracle_shuffle_table.restype = C.c_int
    l.dvb_oracle_shuffle_table.argtypes = [C.c_int32, C.c_uint32, C.c_int32, C.c_int32, C.c_void_p]
    _oracle = l
  return _oracle


class OracleError(RuntimeError):

  def __init__(self, status):
    super().__init__(f'oracle status {status}: {oracle().dvb_oracle_last_error().decode()}')
    self.status = status


def encode_batch(params: _lib.DvbPileupParams, batch: packing.PackedBatch) -> np.ndarray:
  shape = (batch.n_images, params.height, params.width, params.num_channels + params.num_alt_channels)
  out = np.empty(shape, dtype=np.uint8)
  cb = batch.as_ctypes()
  st = oracle().dvb_oracle_encode_batch(C.byref(params), C.byref(cb), out.ctypes.data_as(C.c_void_p))
  if st:
    raise OracleError(st)
  return out


def shuffle_table(n: int, seed: int, max_reads: int, shuffle_stdlib: int = 0) -> np.ndarray:
  out = np.empty(n, dtype=np.int32)
  oracle().dvb_oracle_shuffle_table(n, seed, max_reads, shuffle_stdlib, out.ctypes.data_as(C.c_void_p))
  return out


class OraclePileupImageEncoder:
  """Same surface as PileupImageEncoderNative, computed by the oracle."""

  def __init__(self, options: pi.PileupImageOptions):
    if not (options.width % 2 == 1 and options.width >= 3):
      raise ValueError(f'Width must be odd; found {options.width}')
    self.options = options

  def _params(self, width: int, height: Optional[int] = None, band: Optional[int] = None):
    o = dataclasses.replace(self.options, width=width)
    if band is not None:
      o = dataclasses.replace(o, reference_band_height=band)
    return pi.to_params(o, height=height)

  def all_channels_enum(self, alt=''):
    return pi.all_channels_enum(self.options, alt)

  def encode_reference(self, ref_bases: str) -> np.ndarray:
    p = self._params(len(ref_bases), height=2, band=1)
    out = np.empty((1, len(ref_bases), p.num_channels), dtype=np.uint8)
    rb = np.frombuffer(ref_bases.encode(), dtype=np.uint8).copy()
    st = oracle().dvb_oracle_encode_reference(C.byref(p), rb.ctypes.data_as(C.c_void_p),
                                              out.ctypes.data_as(C.c_void_p))
    if st:
      raise OracleError(st)
    return out

  def encode_read(self, dv_call, ref_bases: str, read, image_start_pos: int, alt_alleles: Sequence[str]):
    p = self._params(len(ref_bases), height=2, band=1)
    p.num_alt_channels = 0
    spec = packing.image_spec_for(dv_call, ref_bases, [read], image_start_pos, list(alt_alleles), self.options)
    batch = packing.pack_images([spec], p)
    cb = batch.as_ctypes()
    out = np.zeros((1, len(ref_bases), p.num_channels), dtype=np.uint8)
    kept = C.c_int32(0)
    st = oracle().dvb_oracle_encode_read(C.byref(p), C.byref(cb), 0, 0, out.ctypes.data_as(C.c_void_p),
                                         C.byref(kept))
    if st:
      raise OracleError(st)
    return out if kept.value else None

  def build_pileup_for_one_sample(self, dv_call, ref_bases, reads, image_start_pos, alt_alleles):
    if len(ref_bases) != self.options.width:
      raise ValueError('ref_bases.size() != options.width')
    p = pi.to_params(self.options)
    spec = packing.image_spec_for(dv_call, ref_bases, list(reads), image_start_pos, list(alt_alleles),
                                  self.options)
    return encode_batch(p, packing.pack_images([spec], p))[0]
