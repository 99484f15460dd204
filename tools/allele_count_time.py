"""Times the CUDA allele-count pass (dvb_allele_count_kernel + dvb_allele_flag_kernel) on a synthetic 30x coordinate-sorted BAM and
reports it against the HBM roofline, next to the host allele counter on the same reads.

  python tools/allele_count_time.py [--mbases 2] [--steps 20]

Algorithmic bytes per launch pair: per read 2 L (bases + qualities) + 4 n_cigar + 24 (position, mapping quality,
two CSR offsets, row index); per position 24 (counters: memset) + 24 (flag pass read) + 3 (indel byte write + read, flag byte) + 1
(reference base); atomics add 4 bytes per counted base (read-modify-write in L2, counted once)."""
import argparse
import ctypes as C
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))


def synthetic_bam(path, n_bases, depth=30, read_len=150, seed=3):
  import test_bam_native as tb
  rng = np.random.default_rng(seed)
  genome = rng.choice(np.frombuffer(b'ACGT', np.uint8), n_bases)
  n_reads = n_bases * depth // read_len
  starts = np.sort(rng.integers(0, n_bases - read_len - 20, n_reads))
  recs = []
  for i, pos in enumerate(starts.tolist()):
    seq = genome[pos:pos + read_len].copy()
    k = rng.random()
    if k < 0.5:
      seq[rng.integers(0, read_len)] = b'ACGT'[int(rng.integers(0, 4))]
    cigar = [(0, read_len)]
    if k > 0.98:
      at = int(rng.integers(20, read_len - 20))
      cigar = [(0, at), (1, 2), (0, read_len - at - 2)] if k > 0.99 else [(0, at), (2, 3), (0, read_len - at)]
    quals = rng.choice(np.array([2, 11, 25, 37, 37, 37, 37, 37], np.uint8), read_len)
    recs.append(tb._record(0, pos, f'r{i}', 60 if rng.random() < 0.9 else int(rng.integers(0, 60)), 0x10 if i % 2 else 0, cigar,
                           seq.tobytes().decode(), quals.tolist()))
  with open(path, 'wb') as f:
    f.write(tb._bam(recs, refs=(('chr1', n_bases),)))
  return genome.tobytes()


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument('--mbases', type=float, default=2.0)
  ap.add_argument('--steps', type=int, default=20)
  ap.add_argument('--warmup', type=int, default=3)
  a = ap.parse_args()
  import torch
  from deepvariant_b200 import _lib, bam, candidates as cand
  import test_candidates as tc
  n = int(a.mbases * 1e6)
  with tempfile.TemporaryDirectory() as d:
    t0 = time.time()
    genome = synthetic_bam(os.path.join(d, 's.bam'), n)
    table = bam.NativeBamTable(os.path.join(d, 's.bam'), bam.ReadRequirements(min_mapping_quality=5))
    print(f'[setup] {table.n_reads} reads over {n} bases in {time.time() - t0:.1f}s', file=sys.stderr)
    ref = tc.FakeRef([('chr1', genome)])
    o = cand.CandidateOptions()
    rows = np.arange(table.n_reads, dtype=np.int64)
    # host allele counter + caller on one core (the reference's own structure of work)
    t0 = time.time()
    host = cand.candidates_in_region(table, ref, 'chr1', 0, n, o, rows=rows)
    host_s = time.time() - t0
    out = {'reads': int(table.n_reads), 'positions': n, 'host_candidates': len(host.records), 'host_counter_reads_per_s_one_core': table.n_reads / host_s}
    if not torch.cuda.is_available():
      print(json.dumps(out))
      return
    lib = _lib.lib()
    counter = cand.GpuAlleleCounter(table)
    dev = torch.device('cuda', 0)
    rows_d = torch.from_numpy(rows).to(dev)
    ref_d = torch.from_numpy(np.frombuffer(genome, np.uint8).copy()).to(dev)
    counts_d = torch.empty(6 * n, dtype=torch.int32, device=dev)
    indel_d = torch.empty(n, dtype=torch.uint8, device=dev)
    flags_d = torch.empty(n, dtype=torch.uint8, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    co = o.to_c()
    stream = torch.cuda.current_stream(dev)

    def launch():
      _lib.check(lib.dvb_allele_count_device(counter._h, C.c_void_p(ref_d.data_ptr()), 0, n, n, 0, n, C.c_void_p(rows_d.data_ptr()), len(rows),
                                             C.byref(co), C.c_void_p(counts_d.data_ptr()), C.c_void_p(indel_d.data_ptr()),
                                             C.c_void_p(flags_d.data_ptr()), C.c_void_p(stream.cuda_stream)))
    for _ in range(a.warmup):
      launch()
    torch.cuda.synchronize()
    ms = []
    for _ in range(a.steps):
      flush.zero_()                       # L2 flush between timed iterations
      e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
      e0.record(stream)
      launch()
      e1.record(stream)
      torch.cuda.synchronize()
      ms.append(e0.elapsed_time(e1))
    want_counts, want_flags = cand.debug_dense_counts_host(table, ref, 'chr1', 0, n, rows, o)
    ok = bool(np.array_equal(counts_d.cpu().numpy(), want_counts) and np.array_equal(flags_d.cpu().numpy(), want_flags))
    n_cigar = int(table.cigar_begin[-1])
    counted = int(want_counts[:n].sum() + want_counts[n:5 * n].sum())
    alg = 2 * int(table.seq_begin[-1]) + 4 * n_cigar + 24 * table.n_reads + n * (24 + 24 + 3 + 1) + 4 * counted
    med = float(np.median(ms))
    peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json'))) if os.path.exists(os.path.join(ROOT, 'MEASURED_PEAKS.json')) else {}
    t0 = time.time()
    gpu = cand.candidates_in_region_gpu(counter, ref, 'chr1', 0, n, o, rows=rows)
    gpu_s = time.time() - t0
    out.update({'parity_with_host_instantiation': ok, 'ms_per_launch_pair': med, 'reads_per_s': table.n_reads / (med * 1e-3),
                'algorithmic_bytes': alg, 'achieved_GBps': alg / (med * 1e-3) / 1e9, 'flagged_positions': int((want_flags != 0).sum()),
                'gpu_flow_candidates_identical': gpu.records == host.records, 'gpu_flow_s': gpu_s, 'host_flow_s': host_s, 'peaks': peaks})
    print(json.dumps(out))


if __name__ == '__main__':
  main()
