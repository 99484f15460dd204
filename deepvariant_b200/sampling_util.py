"""Non-uniform down-sampling of a pileup's reads (SampleOptions.use_non_uniform_downsampling, DeepSomatic): at least
`non_uniform_downsampling_threshold` reads of every allele's supporters survive, the rest of the `max_reads` rows is drawn from
everything else.  Restates

  GetReadIndicesAllelePartition / DownsampleReadIndicesWithMinsPerAllele   deepvariant/pileup_image_native.cc:244-294
  ReservoirSampleImpl / SampleWithPartitionMinsImpl / ReservoirSample      deepvariant/sampling_util.h:55-160

The selection is a SET of read indices, visited in index order by BuildPileupForOneSample (:336-338), so on this side of the C ABI it
is a host-side filter of an image's pairs: the encoder then sees at most `max_reads` reads and never shuffles.

Randomness: std::mt19937_64(random_seed), copied per call, drawn through absl::Uniform<size_t>(absl::IntervalClosed, gen, 0, max).
Abseil is an un-vendored dependency of the reference; its uniform_int_distribution is restated from its published source
(absl/random/uniform_int_distribution.h: power-of-two ranges mask the low bits, otherwise Lemire's multiply-and-reject on one 64-bit
draw per attempt).  No golden file of the reference exercises this path: **parity of the random stream is unpinned**; the structure
(which partitions, the order of the draws, the error fall-back) is pinned by the reference's own distribution tests, restated in
tests/test_sampling_util.py by exhaustive enumeration.
"""
from __future__ import annotations

from typing import Callable, Dict, Iterable, List, Optional, Sequence, Set, Tuple


class Mt19937_64:
  """std::mt19937_64 (the C++ standard fixes the algorithm and its 10000th output, 9981545732273789042 for the default seed)."""
  NN, MM = 312, 156
  MATRIX_A, UM, LM = 0xB5026F5AA96619E9, 0xFFFFFFFF80000000, 0x7FFFFFFF
  MASK = (1 << 64) - 1

  def __init__(self, seed: int = 5489):
    mt = [0] * self.NN
    mt[0] = seed & self.MASK
    for i in range(1, self.NN):
      mt[i] = (6364136223846793005 * (mt[i - 1] ^ (mt[i - 1] >> 62)) + i) & self.MASK
    self.mt, self.mti = mt, self.NN

  def __call__(self) -> int:
    mt, NN, MM = self.mt, self.NN, self.MM
    if self.mti >= NN:
      for i in range(NN):
        x = (mt[i] & self.UM) | (mt[(i + 1) % NN] & self.LM)
        mt[i] = mt[(i + MM) % NN] ^ (x >> 1) ^ (self.MATRIX_A if x & 1 else 0)
      self.mti = 0
    x = mt[self.mti]
    self.mti += 1
    x ^= (x >> 29) & 0x5555555555555555
    x ^= (x << 17) & 0x71D67FFFEDA60000
    x ^= (x << 37) & 0xFFF7EEE000000000
    x ^= x >> 43
    return x & self.MASK


def absl_uniform_closed(gen: Callable[[], int], hi: int) -> int:
  """absl::Uniform<size_t>(absl::IntervalClosed, gen, 0, hi) over a 64-bit engine (uniform_int_distribution<uint64>::Generate)."""
  mask = (1 << 64) - 1
  r = hi & mask
  bits = gen()
  lim = (r + 1) & mask
  if (r & lim) == 0:
    return bits & r                       # the interval's length is a power of two (or the whole 64-bit range)
  product = bits * lim
  if (product & mask) < lim:
    threshold = ((mask - lim + 1) & mask) % lim          # 2^64 mod lim
    while (product & mask) < threshold:
      bits = gen()
      product = bits * lim
  return product >> 64


def reservoir_sample_impl(sample_size: int, index_provider: Callable[[int], int], population: Sequence[int]) -> Set[int]:
  """ReservoirSampleImpl (sampling_util.h:55-79); `population` is the sorted btree_set.  Note the draws that happen even when the
  sample is empty, and that a population of exactly sample_size elements takes the loop-free path."""
  population = list(population)
  if len(population) < sample_size:
    return set(population)
  sampled = population[:sample_size]
  for index in range(sample_size, len(population)):
    swap_index = index_provider(index)
    if swap_index < sample_size:
      sampled[swap_index] = population[index]
  return set(sampled)


def sample_with_partition_mins_impl(partition: Iterable[Iterable[int]], sample_size: int, min_per_partition: int,
                                    subset_provider: Callable[[Sequence[int], int], Set[int]]) -> Optional[Set[int]]:
  """SampleWithPartitionMinsImpl (sampling_util.h:81-115): None = the InvalidArgumentError (the minima alone exceed sample_size).
  `partition` is a btree_set of btree_sets: equal parts collapse, parts are visited in lexicographic order."""
  parts = sorted({tuple(sorted(p)) for p in partition})
  sampled: Set[int] = set()
  unsampled: Set[int] = set()
  for elements in parts:
    chosen = subset_provider(elements, min_per_partition)
    sampled |= chosen
    unsampled |= set(elements) - chosen
  remaining = sample_size - len(sampled)
  if remaining < 0:
    return None
  sampled |= subset_provider(sorted(unsampled), remaining)
  return sampled


def read_indices_allele_partition(allele_support: Dict[str, Sequence[str]], read_keys: Sequence[str]) -> List[Tuple[int, ...]]:
  """GetReadIndicesAllelePartition (pileup_image_native.cc:244-284): one part per allele of DeepVariantCall.allele_support (the reads
  it names, each read in the first allele that names it - map order - and a repeated key standing for its LAST read), plus the reads no
  allele names."""
  index_of: Dict[str, int] = {}
  for i, key in enumerate(read_keys):
    index_of[key] = i
  parts = []
  for allele in sorted(allele_support):       # protobuf map order is unspecified; which allele claims a read named twice depends on it
    part = []
    for name in allele_support[allele]:
      if name in index_of:
        part.append(index_of.pop(name))
    parts.append(tuple(sorted(part)))
  parts.append(tuple(sorted(index_of.values())))
  return parts


def downsample_read_indices_with_mins_per_allele(read_keys: Sequence[str], max_reads: int, allele_support: Dict[str, Sequence[str]],
                                                 min_per_allele: int, random_seed: int) -> Optional[List[int]]:
  """DownsampleReadIndicesWithMinsPerAllele (:286-294): the sorted sampled indices, or None when the thresholds cannot be met (the caller
  falls back to uniform down-sampling, :329-339)."""
  gen = Mt19937_64(random_seed)
  provider = lambda population, k: reservoir_sample_impl(k, lambda mx: absl_uniform_closed(gen, mx), population)   # noqa: E731
  out = sample_with_partition_mins_impl(read_indices_allele_partition(allele_support, read_keys), max_reads, min_per_allele, provider)
  return None if out is None else sorted(out)
