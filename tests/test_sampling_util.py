"""deepvariant_b200/sampling_util.py: the non-uniform down-sampling of DeepSomatic pileups (deepvariant/sampling_util.h,
pileup_image_native.cc:244-294).  The reference tests its two building blocks by exhaustive enumeration of the injected randomness
(deepvariant/sampling_util_test.cc:70-152); the same enumerations are restated here.  The engine is std::mt19937_64 (pinned by the C++
standard's own known answer); abseil's draw on top of it is restated from its published source and unpinned."""
import itertools
from collections import Counter
from fractions import Fraction

import pytest

from deepvariant_b200 import sampling_util as su
from deepvariant_b200 import make_examples_native as men
from deepvariant_b200.protos import DeepVariantCall, Read, Variant, parse_cigar_string


def test_mt19937_64_is_the_standard_engine():
  g = su.Mt19937_64()                       # [rand.predef]: the 10000th consecutive invocation of a default-constructed mt19937_64
  for _ in range(9999):
    g()
  assert g() == 9981545732273789042
  assert su.Mt19937_64(2101079370)() != su.Mt19937_64(2101079371)()


def test_reservoir_sample_is_uniform_over_all_index_providers():
  """ReservoirSampleIsUniform: 3 of {0..6}; the provider is asked for indices 3, 4, 5, 6 and may answer 0..index."""
  population = list(range(7))
  counts = Counter()
  for answers in itertools.product(range(4), range(5), range(6), range(7)):
    it = iter(answers)
    asked = []
    def provider(index, it=it, asked=asked):
      asked.append(index)
      return next(it)
    counts[frozenset(su.reservoir_sample_impl(3, provider, population))] += 1
    assert asked == [3, 4, 5, 6]
  assert len(counts) == 35 and set(counts.values()) == {4 * 5 * 6 * 7 // 35}          # every 3-subset, equally often


def test_reservoir_sample_edge_cases():
  assert su.reservoir_sample_impl(5, lambda i: 0, [1, 2, 3]) == {1, 2, 3}               # population smaller than the sample
  assert su.reservoir_sample_impl(3, lambda i: 1 / 0, [1, 2, 3]) == {1, 2, 3}           # exactly the sample size: no draw
  asked = []
  assert su.reservoir_sample_impl(0, lambda i: asked.append(i) or 0, [4, 5, 6]) == set() and asked == [0, 1, 2]   # draws happen all the same


def _all_subsets(s, k):
  return [frozenset(c) for c in itertools.combinations(sorted(s), k)]


def test_sample_with_partition_mins_distribution():
  """CheckSampleWithPartitionMinsDistribution: partitions {0,1,2} {3,4,5}, 4 elements, at least 1 per partition, uniform subset
  providers -> the sample is balanced (2 + 2) with probability 2/3."""
  partition = [{0, 1, 2}, {3, 4, 5}]
  balanced = Fraction(0)
  # the provider is called three times: 1 of part A, 1 of part B, 2 of the 4 unsampled elements
  for a in range(3):
    for b in range(3, 6):
      rest = sorted(set(range(6)) - {a, b})
      for extra in _all_subsets(rest, 2):
        calls = iter([{a}, {b}, set(extra)])
        got = su.sample_with_partition_mins_impl(partition, 4, 1, lambda pop, k, calls=calls: next(calls))
        assert got == {a, b} | set(extra)
        if len(got & {0, 1, 2}) % 2 == 0:
          balanced += Fraction(1, 3 * 3 * 6)
  assert balanced == Fraction(2, 3)


def test_partition_minima_that_exceed_the_sample_are_an_error():
  assert su.sample_with_partition_mins_impl([{0, 1, 2}, {3, 4, 5}], 3, 2, lambda pop, k: set(list(pop)[:k])) is None
  # equal parts collapse (a set of sets) and parts are visited in lexicographic order
  seen = []
  su.sample_with_partition_mins_impl([{5, 6}, set(), {1, 9}, set(), {1, 2}], 10, 1, lambda pop, k: seen.append(tuple(pop)) or set(list(pop)[:k]))
  assert seen[:4] == [(), (1, 2), (1, 9), (5, 6)]


def test_absl_uniform_stays_in_range_and_masks_powers_of_two():
  g = su.Mt19937_64(7)
  assert all(0 <= su.absl_uniform_closed(g, hi) <= hi for hi in (0, 1, 2, 3, 5, 6, 7, 94, 95, 96, 1000, (1 << 64) - 1) for _ in range(200))
  a, b = su.Mt19937_64(11), su.Mt19937_64(11)
  assert [su.absl_uniform_closed(a, 7) for _ in range(50)] == [b() & 7 for _ in range(50)]      # power-of-two range: the low bits of one draw
  draws = Counter(su.absl_uniform_closed(g, 4) for _ in range(20000))
  assert set(draws) == set(range(5)) and max(draws.values()) - min(draws.values()) < 600


def test_allele_partition_and_the_planner_hook():
  keys = [f'r{i}/1' for i in range(12)]
  support = {'C': ['r1/1', 'r3/1', 'zz/1'], 'G': ['r5/1', 'r3/1']}
  assert su.read_indices_allele_partition(support, keys) == [(1, 3), (5,), (0, 2, 4, 6, 7, 8, 9, 10, 11)]     # r3 is claimed by the first allele
  assert su.read_indices_allele_partition(support, keys + ['r1/1']) [0] == (3, 12)                              # a repeated key stands for its last read
  kept = su.downsample_read_indices_with_mins_per_allele(keys, 6, support, 2, 2101079370)
  assert len(kept) == 6 and kept == sorted(kept) and {1, 3} <= set(kept) and 5 in kept                          # both C reads (minimum 2), the only G read
  assert su.downsample_read_indices_with_mins_per_allele(keys, 6, support, 2, 2101079370) == kept               # a fresh engine per call
  assert su.downsample_read_indices_with_mins_per_allele(keys, 3, {'C': keys[:4], 'G': keys[4:8]}, 2, 1) is None  # 2 + 2 + 2 > 3: uniform fall-back
  assert su.downsample_read_indices_with_mins_per_allele(keys, 40, support, 2, 1) == list(range(12))             # fewer reads than rows: all of them
  # the planner applies it per image: at most `height - band` reads reach the encoder, in their original order
  import dataclasses
  from deepvariant_b200 import pileup_image as pi
  pic = dataclasses.replace(pi.default_options(pi.ReadRequirements(0, 0)), width=21, height=11, channels=list(pi.PILEUP_DEFAULT_CHANNELS), num_channels=6)
  reads = [Read(fragment_name=f'r{i}', read_number=1, position=100 + i % 3, cigar=parse_cigar_string('30M'), aligned_sequence=b'A' * 30, aligned_quality=bytes([30] * 30),
                mapping_quality=60, reference_name='chr1') for i in range(12)]
  call = DeepVariantCall(variant=Variant(reference_name='chr1', start=110, end=111, reference_bases='A', alternate_bases=['C', 'G']), allele_support=support)

  class Ref:
    def n_bases(self, c): return 10000
    def is_valid_interval(self, c, a, b): return True
    def query(self, c, a, b): return 'A' * (b - a)

  for on in (False, True):
    gen = men.ExamplesGenerator(men.MakeExamplesOptions(pic_options=dataclasses.replace(pic, multi_allelic_mode='NO_HET_ALT_IMAGES'), sample_options=[
        men.SampleOptions(use_non_uniform_downsampling=on, non_uniform_downsampling_threshold=2)]), test_mode=True, ref_reader=Ref())
    plans = gen.plan_region([call], reads, {})
    names = [r.fragment_name for r in plans[0].spec.reads]
    assert (names == [f'r{i}' for i in kept]) if on else (len(names) == 12)
