"""Candidate generation for make_examples --mode calling (SURVEY.md 8(f) "next" row #2, host half).

Mirrors how deepvariant/make_examples_core.py drives the allele counter and the very-sensitive caller for ONE sample:

  regions_to_process           make_examples_core.py:799-888 (calling regions cut into --partition_size pieces, piece i -> task i mod N)
  region_reads                 RegionProcessor.region_reads_norealign (:2408-2477) + utils.reservoir_sample
                               (third_party/nucleus/util/utils.py:80-124) with np.random.RandomState(random_seed)
  candidates_in_region         RegionProcessor.candidates_in_region (:2832-2960): [track_ref_reads: first pass for the
                               candidate positions], AlleleCounter over the region, VariantCaller.calls_from_allele_counts
  sample_name_from_bam         assign_sample_name / extract_sample_name_from_sam_reader (:190-212, :470-505)

The arithmetic runs in C++ over the rows of the native BAM table (csrc/dvb_candidates.cu, dvb_candidates_in_region); this module
selects the rows, hands over the contig's bases and parses the DeepVariantCall protos that come back.  The realigner is not
restated: this is make_examples with --norealign_reads (the default for PACBIO / ONT; for WGS / WES the reference realigns
reads first, which moves a minority of candidates - see tools/check_candidates_golden.py for the measured agreement).
"""
from __future__ import annotations

import ctypes as C
import dataclasses
import gzip
import struct
from typing import Iterable, List, Optional, Sequence, Tuple

import numpy as np

from deepvariant_b200 import _lib, protos

DEFAULT_SAMPLE_NAME = 'default'          # dv_constants.DEFAULT_SAMPLE_NAME
MAKE_EXAMPLES_RANDOM_SEED = 609314161    # make_examples_options.py:981


@dataclasses.dataclass
class CandidateOptions:
  """The make_examples flags that reach AlleleCounterOptions / VariantCallerOptions (make_examples_options.py:293-343, 597-922;
  make_vc_options make_examples_core.py:214-248)."""
  min_mapping_quality: int = 5
  min_base_quality: int = 10
  keep_legacy_allele_counter_behavior: bool = False
  track_ref_reads: bool = False
  vsc_min_count_snps: int = 2
  vsc_min_count_indels: int = 2
  vsc_min_fraction_snps: float = 0.12
  vsc_min_fraction_indels: float = 0.06
  vsc_min_fraction_multiplier: float = 1.0
  vsc_min_indel_fraction_for_small_indels: float = 0.0
  vsc_min_indel_fraction_for_large_indels: float = 0.0
  vsc_small_indel_threshold: int = 0
  small_model_vaf_context_window_size: int = 0
  sample_name: str = ''
  max_reads_per_partition: int = 1500
  partition_size: int = 1000
  random_seed: int = MAKE_EXAMPLES_RANDOM_SEED

  def to_c(self) -> _lib.DvbCandidateOptions:
    o = _lib.DvbCandidateOptions()
    _lib.lib().dvb_candidate_options_default(C.byref(o))
    o.min_mapping_quality = self.min_mapping_quality
    o.min_base_quality = self.min_base_quality
    o.keep_legacy_behavior = int(self.keep_legacy_allele_counter_behavior)
    o.track_ref_reads = int(self.track_ref_reads)
    o.min_count_snps = self.vsc_min_count_snps
    o.min_count_indels = self.vsc_min_count_indels
    o.min_fraction_snps = self.vsc_min_fraction_snps
    o.min_fraction_indels = self.vsc_min_fraction_indels
    o.min_fraction_multiplier = self.vsc_min_fraction_multiplier
    o.vsc_min_indel_fraction_for_small_indels = self.vsc_min_indel_fraction_for_small_indels
    o.vsc_min_indel_fraction_for_large_indels = self.vsc_min_indel_fraction_for_large_indels
    o.vsc_small_indel_threshold = self.vsc_small_indel_threshold
    o.small_model_vaf_context_window_size = self.small_model_vaf_context_window_size
    o.sample_name = self.sample_name.encode()
    return o


def sample_name_from_bam(path: str) -> str:
  """SM of the @RG header lines: the only one, or the first when several differ; 'default' when there is none."""
  from deepvariant_b200 import bam
  text = bam.sam_header_text(path)      # BAM or CRAM
  samples = []
  for line in text.split('\n'):
    if line.startswith('@RG'):
      for field in line.rstrip('\r').split('\t')[1:]:
        if field.startswith('SM:') and field[3:]:
          samples.append(field[3:])
  return samples[0] if samples else DEFAULT_SAMPLE_NAME


END_OF_REGION = -1              # make_examples_core.py:125
END_OF_PARTITION = -2           # :129
MAX_PARTITION_LEN = 1000000     # :134
MAX_CANDIDATES_PER_PARTITION = 200   # regions_to_process, :874


def partition_by_candidates(regions: Sequence[Tuple[str, int, int]], candidate_positions: Sequence[int], max_size: int) -> List[Tuple[str, int, int]]:
  """make_examples_core.py:714-796: cuts the calling intervals so that none holds more than max_size candidates (nor spans more than
  MAX_PARTITION_LEN without one).  candidate_positions = the sorted positions of every interval in turn, each interval's run closed
  by END_OF_REGION (the merged output of the candidate_sweep shards)."""
  if max_size <= 0:
    raise ValueError('max_size must be > 0: {}'.format(max_size))
  out = []
  it = 0
  n = len(candidate_positions)
  for name, start, end in regions:
    count = 0
    p_start = p_end = start
    while it < n and candidate_positions[it] != END_OF_REGION and start <= candidate_positions[it] < end:
      if count == max_size or p_end - p_start >= MAX_PARTITION_LEN:
        for pos in range(p_start, p_end, MAX_PARTITION_LEN):
          out.append((name, pos, min(p_end, pos + MAX_PARTITION_LEN)))
        p_start = p_end
        p_end = p_start + 1
        count = 0
      else:
        p_end = int(candidate_positions[it]) + 1
        count += 1
      it += 1
    if it < n and candidate_positions[it] == END_OF_REGION:
      for pos in range(p_start, end, MAX_PARTITION_LEN):
        out.append((name, pos, min(end, pos + MAX_PARTITION_LEN)))
      it += 1
    else:
      raise ValueError('Terminating item is missing in candidates list')
  return out


def merge_ranges_from_files_sequential(position_arrays: Sequence[Sequence[int]]) -> List[int]:
  """make_examples_core.py:3247-3325: the candidate_sweep shards hold the positions of partitions i, i + N, i + 2N ... each closed
  by END_OF_PARTITION (and END_OF_REGION after the last partition of a calling region); reading one partition from every shard
  in turn restores genome order.  Raises AssertionError when the result is not increasing."""
  merged: List[int] = []
  index = [0] * len(position_arrays)
  shard = 0
  left = len(position_arrays)
  while left > 0:
    arr = position_arrays[shard]
    while index[shard] < len(arr):
      if arr[index[shard]] == END_OF_PARTITION:
        index[shard] += 1
        if index[shard] < len(arr) and arr[index[shard]] == END_OF_REGION:
          merged.append(END_OF_REGION)
          index[shard] += 1
        break
      if merged:
        assert arr[index[shard]] > merged[-1]
      merged.append(int(arr[index[shard]]))
      index[shard] += 1
    if index[shard] == len(arr):
      left -= 1
    for _ in range(len(position_arrays)):                      # move_to_the_next_non_exhausted_shard (:3222-3244)
      shard = (shard + 1) % len(position_arrays)
      if index[shard] < len(position_arrays[shard]):
        break
  return merged


def load_candidate_positions(path_spec: str) -> List[int]:
  """make_examples_core.py:3328-3339: every shard of `path@N` (or the one file); unreadable shards are skipped as there."""
  from deepvariant_b200 import tfrecord
  arrays = []
  for path in (tfrecord.shard_paths(path_spec) if tfrecord.is_sharded_spec(path_spec) else [path_spec]):
    try:
      arrays.append(np.fromfile(path, dtype=np.int32))
    except IOError:
      continue
  return merge_ranges_from_files_sequential(arrays)


def calling_intervals(contigs: Sequence[Tuple[str, int]], calling_regions=None) -> List[Tuple[str, int, int]]:
  """RangeSet.from_contigs(contigs).intersection(calling_regions) (make_examples_core.py:868-870): the calling regions - one
  (contig, start, end) or a list of them, in any order, overlapping ones merged - clipped to the contigs, in contig order."""
  if calling_regions is None:
    return [(name, 0, n) for name, n in contigs if n > 0]
  if calling_regions and isinstance(calling_regions[0], str):
    calling_regions = [calling_regions]
  out = []
  for name, n_bases in contigs:
    own = sorted((max(0, s), min(n_bases, e)) for c, s, e in calling_regions if c == name)
    merged: List[List[int]] = []
    for s, e in own:
      if s >= e:
        continue
      if merged and s < merged[-1][1]:
        merged[-1][1] = max(merged[-1][1], e)
      else:
        merged.append([s, e])
    out += [(name, s, e) for s, e in merged]
  return out


def regions_to_process(contigs: Sequence[Tuple[str, int]], partition_size: int, calling_region: Optional[Tuple[str, int, int]] = None,
                       task_id: Optional[int] = None, num_shards: Optional[int] = None,
                       candidates: Optional[Sequence[int]] = None) -> List[Tuple[str, int, int]]:
  """contigs = [(name, n_bases)] in reference order; calling_region = (contig, start, end) 0-based half-open, a list of them, or None.
  make_examples_core.py:799-888; with `candidates` (load_candidate_positions) the partitions are cut by candidate count instead of
  by length.  Shards take partitions round robin (the TFRecord case, :3439)."""
  if (task_id is None) != (num_shards is None):
    raise ValueError('Both task_id and num_shards must be present if either is', task_id, num_shards)
  if num_shards:
    if task_id < 0 or task_id >= num_shards:
      raise ValueError('task_id={} should be >= 0 and < num_shards={}'.format(task_id, num_shards))
  if partition_size <= 0:
    raise ValueError('max_size must be > 0: {}'.format(partition_size))
  intervals = calling_intervals(contigs, calling_region)
  if candidates is not None:
    pieces = partition_by_candidates(intervals, candidates, MAX_CANDIDATES_PER_PARTITION)
  else:
    pieces = [(name, pos, min(hi, pos + partition_size)) for name, lo, hi in intervals for pos in range(lo, hi, partition_size)]
  if num_shards:
    return [r for i, r in enumerate(pieces) if i % num_shards == task_id]
  return pieces


def reservoir_sample(items: Iterable, k: int, random=None) -> list:
  """utils.reservoir_sample (Algorithm R): same draws, same order of the retained elements."""
  if k < 0:
    raise ValueError('k must be nonnegative, but got {}'.format(k))
  if random is None:
    random = np.random
  sample = []
  for i, item in enumerate(items):
    if len(sample) < k:
      sample.append(item)
    else:
      j = random.randint(0, i + 1)
      if j < k:
        sample[j] = item
  return sample


def region_reads(table, contig: str, start: int, end: int, max_reads_per_partition: int = 1500,
                 random_seed: int = MAKE_EXAMPLES_RANDOM_SEED) -> np.ndarray:
  """Rows of the BAM table that the reference would hold in its InMemorySamReader for this region."""
  rows = table.query_indices(contig, start, end)
  if max_reads_per_partition > 0 and len(rows) > max_reads_per_partition:
    rows = np.asarray(reservoir_sample(rows.tolist(), max_reads_per_partition, np.random.RandomState(random_seed)), dtype=np.int64)
  return np.ascontiguousarray(rows, dtype=np.int64)


class NativeCandidates:
  """Result of one dvb_candidates_in_region call: raw DeepVariantCall records + parsed objects on demand."""

  def __init__(self, handle, keep: Optional[Tuple[int, int]] = None, interval: Optional[Tuple[int, int]] = None):
    lib = _lib.lib()
    try:
      sp = C.c_void_p()
      n_sites = int(lib.dvb_candidates_summary_counts(handle, C.byref(sp)))
      # AlleleCounter::SummaryCounts of the counted interval: [n_sites, 2] = (ref_supporting_read_count, total_read_count)
      self.summary_counts = np.frombuffer((C.c_char * (8 * n_sites)).from_address(sp.value), dtype=np.int32).reshape(n_sites, 2).copy() \
          if n_sites else np.zeros((0, 2), np.int32)
      self.interval = interval
      n = int(lib.dvb_candidates_count(handle))
      pos_ptr = C.c_void_p()
      lib.dvb_candidates_positions(handle, C.byref(pos_ptr))
      starts = np.frombuffer((C.c_char * (4 * n)).from_address(pos_ptr.value), dtype=np.int32).copy() if n else np.zeros(0, np.int32)
      data, begin = C.c_void_p(), C.c_void_p()
      _lib.check(lib.dvb_candidates_protos(handle, C.byref(data), C.byref(begin)))
      offs = np.frombuffer((C.c_char * (8 * (n + 1))).from_address(begin.value), dtype=np.int64).copy()
      blob = bytes((C.c_char * int(offs[-1])).from_address(data.value)) if n and offs[-1] else b''
      self.all_records: List[bytes] = [blob[int(offs[i]):int(offs[i + 1])] for i in range(n)]     # incl. the padding (phasing sees them)
      self.records: List[bytes] = [r for i, r in enumerate(self.all_records) if keep is None or keep[0] <= int(starts[i]) < keep[1]]
    finally:
      lib.dvb_candidates_free(handle)
    self._parsed: Optional[List[protos.DeepVariantCall]] = None

  def calls(self) -> List[protos.DeepVariantCall]:
    if self._parsed is None:
      self._parsed = [protos.parse_deepvariant_call(r) for r in self.records]
    return self._parsed


def _contig_buffer(ref_reader, contig: str):
  seq = ref_reader._contig(contig)   # pylint: disable=protected-access  (upper-cased bytes of the whole contig, cached)
  # the pointer is cached beside the bytes object it points into (wrapping a 60 MB contig in a c_char_p costs ~14 ms a call)
  cache = ref_reader.__dict__.setdefault('_dvb_ptr_cache', {})
  hit = cache.get(contig)
  if hit is None or hit[0] is not seq:
    hit = cache[contig] = (seq, C.cast(C.c_char_p(seq), C.c_void_p))
  return hit


def candidate_positions(table, ref_reader, contig: str, start: int, end: int, rows: np.ndarray, options: CandidateOptions) -> List[int]:
  """VariantCaller.get_candidate_positions (first pass of track_ref_reads)."""
  lib = _lib.lib()
  seq, ptr = _contig_buffer(ref_reader, contig)
  rows = np.ascontiguousarray(rows, dtype=np.int64)
  co = options.to_c()
  co.track_ref_reads = 0
  h = C.c_void_p()
  _lib.check(lib.dvb_candidate_positions(table.handle, ptr, len(seq), start, end, rows.ctypes.data_as(C.c_void_p), len(rows),
                                         C.byref(co), C.byref(h)))
  try:
    p = C.c_void_p()
    n = int(lib.dvb_candidates_positions(h, C.byref(p)))
    return np.frombuffer((C.c_char * (4 * n)).from_address(p.value), dtype=np.int32).tolist() if n else []
  finally:
    lib.dvb_candidates_free(h)


def candidates_in_region(table, ref_reader, contig: str, start: int, end: int, options: CandidateOptions,
                         rows: Optional[np.ndarray] = None, padding_pct: int = 0) -> NativeCandidates:
  """Candidates of [start, end) on `contig` from the reads `rows` (default: region_reads of the region).

  padding_pct > 0 (--phase_reads: dv_constants.PHASE_READS_REGION_PADDING_PCT = 20) counts alleles over the region expanded
  by that share of its length on both sides (ranges.expand, make_examples_core.py:2306-2318) - still from the reads that
  overlap the unpadded region - and keeps the candidates that start inside the region (filter_candidates_by_region)."""
  lib = _lib.lib()
  if rows is None:
    rows = region_reads(table, contig, start, end, options.max_reads_per_partition, options.random_seed)
  rows = np.ascontiguousarray(rows, dtype=np.int64)
  seq, ptr = _contig_buffer(ref_reader, contig)
  end = min(end, len(seq))
  region = (start, end)
  if padding_pct > 0:
    pad = int((end - start) * padding_pct / 100)
    start, end = max(start - pad, 0), min(end + pad, len(seq))
  positions = np.zeros(0, dtype=np.int32)
  if options.track_ref_reads:
    positions = np.asarray(candidate_positions(table, ref_reader, contig, start, end, rows, options), dtype=np.int32)
  co = options.to_c()
  h = C.c_void_p()
  _lib.check(lib.dvb_candidates_in_region(table.handle, contig.encode(), ptr, len(seq), start, end, rows.ctypes.data_as(C.c_void_p),
                                          len(rows), C.byref(co), positions.ctypes.data_as(C.c_void_p), len(positions), C.byref(h)))
  return NativeCandidates(h, keep=region if padding_pct > 0 else None, interval=(start, end))


def _value(buf):
  for fn, wt, val, _ in protos.iter_fields(buf):
    if fn == 7:
      return protos._to_signed32(val)   # pylint: disable=protected-access
    if fn == 2:
      return struct.unpack('<d', struct.pack('<Q', val))[0] if isinstance(val, int) else struct.unpack('<d', bytes(val))[0]
    if fn == 3:
      return bytes(val).decode()
  return None


def _read_support(buf):
  d = {'read_name': '', 'is_low_quality': 0, 'mapping_quality': 0, 'average_base_quality': 0, 'is_reverse_strand': 0, 'sample_name': ''}
  names = {1: 'read_name', 2: 'is_low_quality', 3: 'mapping_quality', 4: 'average_base_quality', 5: 'is_reverse_strand', 7: 'sample_name'}
  for fn, wt, val, _ in protos.iter_fields(buf):
    if fn in names:
      d[names[fn]] = bytes(val).decode() if wt == 2 else int(val)
  return d


def canonical_call(record: bytes) -> dict:
  """Semantic content of a DeepVariantCall, independent of map / field order."""
  out = {'ref': '', 'alts': [], 'start': 0, 'end': 0, 'contig': '', 'info': {}, 'call_set_name': '', 'genotype': [],
         'allele_support': {}, 'allele_support_ext': {}, 'ref_support': [], 'ref_support_ext': [], 'af_at_position': {}}
  for fn, wt, val, _ in protos.iter_fields(record):
    val = bytes(val) if wt == 2 else val
    if fn == 1:
      for f2, w2, v2, _ in protos.iter_fields(val):
        v2 = bytes(v2) if w2 == 2 else v2
        if f2 == 6:
          out['ref'] = v2.decode()
        elif f2 == 7:
          out['alts'].append(v2.decode())
        elif f2 == 13:
          out['end'] = int(v2)
        elif f2 == 14:
          out['contig'] = v2.decode()
        elif f2 == 16:
          out['start'] = int(v2)
        elif f2 == 11:
          for f3, w3, v3, _ in protos.iter_fields(v2):
            v3 = bytes(v3) if w3 == 2 else v3
            if f3 == 2:
              key, vals = '', []
              for f4, w4, v4, _ in protos.iter_fields(v3):
                if f4 == 1:
                  key = bytes(v4).decode()
                elif f4 == 2:
                  vals = [_value(bytes(v5)) for f5, w5, v5, _ in protos.iter_fields(bytes(v4)) if f5 == 1]
              out['info'][key] = vals
            elif f3 == 7:
              out['genotype'] = [protos._to_signed32(x) for x in protos.unpack_varints(v3)] if w3 == 2 else out['genotype'] + [protos._to_signed32(v3)]   # pylint: disable=protected-access
            elif f3 == 9:
              out['call_set_name'] = v3.decode()
    elif fn == 2:
      key, names = '', []
      for f2, w2, v2, _ in protos.iter_fields(val):
        if f2 == 1:
          key = bytes(v2).decode()
        elif f2 == 2:
          names = [bytes(v3).decode() for f3, w3, v3, _ in protos.iter_fields(bytes(v2)) if f3 == 1]
      out['allele_support'][key] = sorted(names)
    elif fn == 4:
      out['ref_support'].append(val.decode())
    elif fn == 5:
      key, infos = '', []
      for f2, w2, v2, _ in protos.iter_fields(val):
        if f2 == 1:
          key = bytes(v2).decode()
        elif f2 == 2:
          infos = [_read_support(bytes(v3)) for f3, w3, v3, _ in protos.iter_fields(bytes(v2)) if f3 == 1]
      out['allele_support_ext'][key] = sorted(infos, key=lambda d: d['read_name'])
    elif fn == 6:
      out['ref_support_ext'] = sorted((_read_support(bytes(v3)) for f3, w3, v3, _ in protos.iter_fields(val) if f3 == 1),
                                      key=lambda d: d['read_name'])
    elif fn == 7:
      k = v = 0
      for f2, w2, v2, _ in protos.iter_fields(val):
        if f2 == 1:
          k = int(v2)
        elif f2 == 2:
          v = int(v2)
      out['af_at_position'][str(k)] = v
  out['ref_support'].sort()
  return out



def debug_allele_counts(table, ref_reader, contig: str, start: int, end: int, rows, options: CandidateOptions,
                        candidate_positions: Sequence[int] = ()) -> list:
  """Test access to AlleleCounter.Counts(): per position {'ref': n, 'alleles': [[bases, type, low_quality, key, mapq, avg_bq, reverse]]}."""
  import json
  lib = _lib.lib()
  seq, ptr = _contig_buffer(ref_reader, contig)
  rows = np.ascontiguousarray(rows, dtype=np.int64)
  cp = np.ascontiguousarray(candidate_positions, dtype=np.int32)
  co = options.to_c()
  args = (table.handle, ptr, len(seq), start, end, rows.ctypes.data_as(C.c_void_p), len(rows), C.byref(co),
          cp.ctypes.data_as(C.c_void_p), len(cp))
  n = int(lib.dvb_debug_allele_counts(*args, None, 0))
  if n < 0:
    _lib.check(-n)
  buf = C.create_string_buffer(n + 1)
  lib.dvb_debug_allele_counts(*args, buf, n + 1)
  return json.loads(buf.value.decode())


# ---- allele counting on the device + exact calls on the flagged sites ---------------------------------------------------------------
def _dense_arrays(n: int):
  return np.zeros(6 * n, dtype=np.int32), np.zeros(n, dtype=np.uint8)


def split_dense_counts(counts: np.ndarray, n: int):
  """int32[6 n] -> (ref_count[n], subst[n, 4] by read base A, C, G, T, other[n])."""
  return counts[:n], counts[n:5 * n].reshape(n, 4), counts[5 * n:]


def debug_dense_counts_host(table, ref_reader, contig: str, start: int, end: int, rows, options: CandidateOptions, windowed: bool = True):
  """Test access: the device pass (walk + dense sink + flags) instantiated on the host."""
  lib = _lib.lib()
  seq, ptr = _contig_buffer(ref_reader, contig)
  rows = np.ascontiguousarray(rows, dtype=np.int64)
  counts, flags = _dense_arrays(end - start)
  co = options.to_c()
  _lib.check(lib.dvb_debug_allele_count_dense_host(table.handle, ptr, len(seq), start, end, rows.ctypes.data_as(C.c_void_p), len(rows),
                                                   C.byref(co), int(windowed), counts.ctypes.data_as(C.c_void_p), flags.ctypes.data_as(C.c_void_p)))
  return counts, flags


class GpuAlleleCounter:
  """The reads of a NativeBamTable resident in HBM (dvb_device_reads_create) + the allele-count / flag kernels.  Raises when no
  CUDA device is present: there is no CPU path behind this class (candidates_in_region is the host implementation)."""

  def __init__(self, table, device: int = 0):
    self._lib = _lib.lib()
    self.table = table
    h = C.c_void_p()
    _lib.check(self._lib.dvb_device_reads_create(table.handle, device, C.byref(h)))
    self._h = h

  def count_region(self, ref_reader, contig: str, start: int, end: int, rows, options: CandidateOptions):
    """-> (counts int32[6 len], flags uint8[len]); host arrays in and out (dvb_allele_count_host)."""
    seq, ptr = _contig_buffer(ref_reader, contig)
    rows = np.ascontiguousarray(rows, dtype=np.int64)
    counts, flags = _dense_arrays(end - start)
    co = options.to_c()
    _lib.check(self._lib.dvb_allele_count_host(self._h, self.table.handle, ptr, len(seq), start, end, rows.ctypes.data_as(C.c_void_p),
                                               len(rows), C.byref(co), counts.ctypes.data_as(C.c_void_p), flags.ctypes.data_as(C.c_void_p)))
    return counts, flags

  @property
  def launch_count(self) -> int:
    return int(self._lib.dvb_device_reads_launch_count(self._h))

  def close(self):
    if getattr(self, '_h', None):
      self._lib.dvb_device_reads_destroy(self._h)
      self._h = None

  def __del__(self):
    try:
      self.close()
    except Exception:  # pylint: disable=broad-except
      pass


def candidates_at_flagged_positions(table, ref_reader, contig: str, start: int, end: int, options: CandidateOptions, rows: np.ndarray,
                                    flags: np.ndarray) -> NativeCandidates:
  """The exact caller on the sites the device pass flagged: only the reads that overlap a flagged site (by one base on either
  side, or the allele-frequency context when requested) are walked on the host.  Equals candidates_in_region(rows) whenever
  `flags` is a superset of its sites (tests/test_candidates.py)."""
  if options.track_ref_reads:
    raise NotImplementedError('track_ref_reads needs the two-pass host counter (candidates_in_region)')
  lib = _lib.lib()
  seq, ptr = _contig_buffer(ref_reader, contig)
  emit = (np.nonzero(flags)[0] + start).astype(np.int32)
  rows = np.ascontiguousarray(rows, dtype=np.int64)
  if len(emit):
    margin = 1 + (options.small_model_vaf_context_window_size // 2 + 1 if options.small_model_vaf_context_window_size > 0 else 0)
    pos, rend = table.pos[rows].astype(np.int64), table.end[rows].astype(np.int64)
    # reads overlapping [p - margin, p + margin] for some flagged p: first flagged p >= pos - margin must be <= end - 1 + margin
    k = np.searchsorted(emit, pos - margin, side='left')
    keep = (k < len(emit)) & (emit[np.minimum(k, len(emit) - 1)] <= rend - 1 + margin)
    rows = np.ascontiguousarray(rows[keep])
  else:
    rows = rows[:0]
  co = options.to_c()
  h = C.c_void_p()
  _lib.check(lib.dvb_candidates_at_positions(table.handle, contig.encode(), ptr, len(seq), start, end, rows.ctypes.data_as(C.c_void_p),
                                             len(rows), C.byref(co), None, 0, emit.ctypes.data_as(C.c_void_p), len(emit), C.byref(h)))
  return NativeCandidates(h)


def candidates_in_region_gpu(counter: GpuAlleleCounter, ref_reader, contig: str, start: int, end: int, options: CandidateOptions,
                             rows: Optional[np.ndarray] = None) -> NativeCandidates:
  """candidates_in_region with the allele counting on the device: count + flag kernels over the region's reads, then the exact
  calls on the flagged sites."""
  table = counter.table
  if rows is None:
    rows = region_reads(table, contig, start, end, options.max_reads_per_partition, options.random_seed)
  end = min(end, ref_reader.n_bases(contig))
  _, flags = counter.count_region(ref_reader, contig, start, end, rows, options)
  return candidates_at_flagged_positions(table, ref_reader, contig, start, end, options, rows, flags)
