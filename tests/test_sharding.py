"""Region sharding across ranks: independent units, no data-path collective.  The N>1 host logic is
exercised with a world_size-2 gloo group on CPU."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from deepvariant_b200 import make_examples_native as men
from deepvariant_b200.protos import DeepVariantCall, Variant


def _cands(n=97):
  out = []
  for i in range(n):
    start = 10_000 + i * 137
    out.append(DeepVariantCall(variant=Variant(reference_name='chr20' if i % 5 else 'chr21', start=start, end=start + 1,
                                               reference_bases='A', alternate_bases=['C'])))
  return out


def test_partitions_are_disjoint_and_exhaustive():
  cands = _cands()
  parts = men.partition_candidates(cands, 1000)
  assert sum(len(c) for _, c in parts) == len(cands)
  for (contig, k, origin), cs in parts:
    assert all(c.variant.reference_name == contig and (c.variant.start - origin) // 1000 == k for c in cs)
  for n in (1, 2, 3, 8):
    shards = [men.shard_partitions(parts, n, t) for t in range(n)]
    keys = [key for sh in shards for key, _ in sh]
    assert sorted(keys) == sorted(k for k, _ in parts) and len(set(keys)) == len(keys)
  only20 = men.partition_candidates(cands, 1000, region=('chr20', 10_000, 15_000))
  assert all(key[0] == 'chr20' for key, _ in only20) and sum(len(c) for _, c in only20) == sum(
      1 for c in cands if c.variant.reference_name == 'chr20' and c.variant.start < 15_000)


def _worker(rank, world, port, q):
  os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
  dist.init_process_group('gloo', rank=rank, world_size=world)
  cands = _cands()
  mine = men.shard_partitions(men.partition_candidates(cands, 1000), world, rank)
  starts = sorted(c.variant.start for _, cs in mine for c in cs)
  n = torch.tensor([len(starts)], dtype=torch.int64)
  dist.all_reduce(n)                      # bookkeeping only: the data path itself has no collective
  t = torch.tensor([0.001 * (rank + 1)], dtype=torch.float64)
  dist.all_reduce(t, op=dist.ReduceOp.MAX)  # bench.py's max-over-ranks timing
  gathered = [None] * world
  dist.all_gather_object(gathered, starts)
  q.put((rank, int(n.item()), float(t.item()), gathered))
  dist.barrier()
  dist.destroy_process_group()


def test_two_rank_gloo_sharding():
  ctx = mp.get_context('spawn')
  q = ctx.Queue()
  port = 29500 + os.getpid() % 2000
  procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
  for p in procs:
    p.start()
  res = [q.get(timeout=120) for _ in procs]
  for p in procs:
    p.join(timeout=60)
    assert p.exitcode == 0
  for rank, total, tmax, gathered in res:
    assert total == 97 and abs(tmax - 0.002) < 1e-12
    assert not set(gathered[0]) & set(gathered[1])
    assert sorted(gathered[0] + gathered[1]) == sorted(c.variant.start for c in _cands())


def test_shared_contigs_excludes_decoys_and_checks_the_build():
  """_ensure_consistent_contigs (make_examples_core.py:540-640): excluded names are dropped from the reference side, a contig is shared when
  name AND length agree, and less than 90 % shared bases is an error."""
  from deepvariant_b200 import cli
  ref = [('chr1', 1000), ('chr2', 800), ('chr1_KI270706v1_random', 50), ('chrUn_GL000195v1', 60), ('chrEBV', 70), ('HLA-A*01:01:01:01', 3), ('chrUn_KN707606v1_decoy', 9),
         ('chr6_GL000250v2_alt', 10)]
  reads = {'chr1': 1000, 'chr2': 800, 'chrEBV': 70, 'chr1_KI270706v1_random': 50}
  assert cli.shared_contigs(ref, reads) == [('chr1', 1000), ('chr2', 800)]
  assert cli.shared_contigs([('20', 100), ('GL000207.1', 5), ('hs37d5', 9), ('NC_007605', 3)], {'20': 100, 'hs37d5': 9}) == [('20', 100)]
  import pytest
  with pytest.raises(ValueError, match='common genome reference build'):
    cli.shared_contigs(ref, {'chr1': 999, 'chr2': 800})                       # chr1 has another length: 800 of 1800 bases shared
  assert cli.shared_contigs(ref, {'chr1': 1000}, min_coverage_fraction=0.5) == [('chr1', 1000)]
  with pytest.raises(ValueError):
    cli.shared_contigs(ref, {})
