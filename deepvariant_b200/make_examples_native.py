"""Host-side mirror of deepvariant.python.make_examples_native (the per-region example driver).

  ExamplesGenerator(options, example_filenames, test_mode=False)
      .write_examples_in_region(candidates, reads_per_sample, sample_order, role,
                                mean_coverage_per_sample) -> (stats dict, image_shape[3])
      .signal_shard_finished()
  (deepvariant/python/make_examples_native_pybind.cc:56-109)

Restated host logic, each citing the reference (deepvariant/make_examples_native.cc):
  alt_allele_combinations          :191-267   encoded_variant_type            :301-321
  encode_alt_alleles               :350-374   encode_example                  :388-474
  get_reference_bases_for_pileup   :514-538   need_alt_alignment              :498-512
  create_and_write_examples_for_candidate :632-736 (as plan_region + finish_region)
  InMemoryReader.query             :802-810 + third_party/nucleus/util/utils.cc:172-240
  trim_cigar / trim_read / trim_reads / calculate_alignment_region
                                   deepvariant/alt_aligned_pileup_lib.cc:91-266

Where the reference encodes one candidate at a time on the CPU, this driver PLANS a whole region
(every candidate x alt-allele combination becomes one image spec), encodes all images in ONE CUDA
launch through libdvb.so, then serialises the tf.Examples.  There is no CPU encoder in this module.

Scope notes: single sample; alt-aligned pileups ('diff_channels' / 'base_channels')
are produced with zero alt channels — exactly what the reference emits for SNP candidates under
types_to_alt_align='indels' — and candidates that would need haplotype re-alignment (indels) are
counted in stats['n_needs_alt_alignment'] (SSW re-alignment is §8(f) "next" #3).  'rows' /
'single_row' layouts, training labels and the shared-memory stream are not implemented.
"""
from __future__ import annotations

import dataclasses
import json
import re
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from deepvariant_b200 import packing
from deepvariant_b200 import pileup_image as pi
from deepvariant_b200 import protos
from deepvariant_b200 import tfrecord
from deepvariant_b200.protos import DeepVariantCall, Read, Variant

K_DEFAULT_MINIMUM_READ_OVERLAP = 15  # make_examples_native.cc:75
VERSION = '1.10.0'


@dataclasses.dataclass
class SampleOptions:
  role: str = 'main_sample'
  name: str = ''
  pileup_height: int = 0
  order: List[int] = dataclasses.field(default_factory=lambda: [0])
  keep_only_window_spanning_reads: bool = False
  channels_enum_to_blank: List[int] = dataclasses.field(default_factory=list)   # deepvariant.proto SampleOptions: channel enums left blank in this sample's read rows
  use_non_uniform_downsampling: bool = False        # SampleOptions 19 / 20 (make_examples_somatic.py:189-202): keep at least `threshold` reads
  non_uniform_downsampling_threshold: int = 0       # of every allele's supporters when a pileup has more reads than rows


@dataclasses.dataclass
class MakeExamplesOptions:
  pic_options: pi.PileupImageOptions = dataclasses.field(default_factory=pi.default_options)
  sample_options: List[SampleOptions] = dataclasses.field(default_factory=lambda: [SampleOptions()])
  reference_filename: str = ''
  trim_reads_for_pileup: bool = False
  stream_examples: bool = False
  # realigner_options.aln_config (deepvariant/realigner/realigner.py:168-243 flag defaults): the aligner of the alt-aligned pileups
  aln_config: Dict[str, float] = dataclasses.field(default_factory=lambda: dict(
      match=4, mismatch=6, gap_open=8, gap_extend=2, kmer_size=32, max_num_of_mismatches=2, realignment_similarity_threshold=0.16934))


# ---- small pure functions ---------------------------------------------------------------------

_NON_CANONICAL = re.compile('[^ACGT]')


def encoded_variant_type(variant: Variant) -> int:
  """EncodedVariantType (make_examples_native.cc:301-321): 1 = SNP, 2 = indel, 0 = unknown."""
  if len(variant.reference_bases) == 1 and len(variant.alternate_bases) >= 1:
    if all(len(a) == 1 for a in variant.alternate_bases):
      return 1
  if len(variant.reference_bases) > 1:
    return 2
  if any(len(a) > 1 for a in variant.alternate_bases):
    return 2
  return 0


def alt_allele_combinations(candidate: DeepVariantCall, multi_allelic_mode: str) -> List[List[str]]:
  """AltAlleleCombinations / AltAlleleCombinationsFromIndices (make_examples_native.cc:191-267)."""
  variant = candidate.variant
  if multi_allelic_mode not in ('NO_HET_ALT_IMAGES', 'ADD_HET_ALT_IMAGES'):
    raise ValueError('multi_allelic_mode cannot be UNSPECIFIED')
  if candidate.make_examples_alt_allele_indices:
    alts = list(variant.alternate_bases)
    out = []
    for indices in candidate.make_examples_alt_allele_indices:
      if multi_allelic_mode == 'NO_HET_ALT_IMAGES':
        if len(indices) == 1:
          out.append([alts[indices[0]]])
      else:
        out.append([alts[i] for i in indices])
    return out
  if multi_allelic_mode == 'NO_HET_ALT_IMAGES':
    return [[alt] for alt in variant.alternate_bases]
  alts = [variant.reference_bases] + list(variant.alternate_bases)
  out = []
  for i in range(len(alts)):
    for j in range(i + 1, len(alts)):
      comb = []
      if i > 0:  # the ref allele is not used in combinations
        comb.append(alts[i])
      comb.append(alts[j])
      out.append(comb)
  return out


def encode_alt_alleles(variant: Variant, alt_combination: Sequence[str]) -> Tuple[bytes, List[int]]:
  """EncodeAltAlleles (make_examples_native.cc:350-374).  A later duplicate alt overwrites the
  index of an earlier one (flat_hash_map assignment)."""
  alt_indices = {}
  for i, alt in enumerate(variant.alternate_bases):
    alt_indices[alt] = i
  indices = [alt_indices.get(alt, 0) for alt in alt_combination]  # operator[] default-inserts 0
  return protos.encode_alt_allele_indices(indices), indices


def need_alt_alignment(variant: Variant, pic: pi.PileupImageOptions) -> bool:
  """NeedAltAlignment (make_examples_native.cc:498-512)."""
  if pic.alt_aligned_pileup == 'none' or not pic.alt_aligned_pileup:
    return False
  if pic.types_to_alt_align == 'all':
    return True
  if pic.types_to_alt_align == 'indels':
    return len(variant.reference_bases) > 1 or any(len(a) > 1 for a in variant.alternate_bases)
  return False


def read_overlaps_region(read: Read, contig: str, start: int, end: int) -> bool:
  """nucleus ReadOverlapsRegion (third_party/nucleus/util/utils.cc:172-188)."""
  return end > read.position and start < read.end() and contig == read.reference_name


# ---- trimming (alt_aligned_pileup_lib.cc) -----------------------------------------------------

_REF_ADVANCING = (0, 7, 8, 2, 3)      # M = X D N   (IsOperationRefAdvancing, :60-72)
_READ_ADVANCING = (0, 7, 1, 4, 8)     # M = I S X   (IsOperationReadAdvancing, :74-88)


def trim_cigar(cigar: Sequence[Tuple[int, int]], ref_start: int, ref_length: int):
  """TrimCigar (alt_aligned_pileup_lib.cc:91-150) -> (new_cigar, read_start, new_read_length)."""
  trim_remaining = ref_start
  ref_to_cover_remaining = ref_length
  read_start = 0
  new_read_length = 0
  new_cigar: List[Tuple[int, int]] = []
  for op, op_len in cigar:
    c_operation_length = op_len
    advances_ref = op in _REF_ADVANCING
    advances_read = op in _READ_ADVANCING
    ref_step = c_operation_length if advances_ref else 0
    if trim_remaining > 0:
      if ref_step <= trim_remaining:
        trim_remaining -= ref_step
        read_start += c_operation_length if advances_read else 0
        continue
      else:
        ref_step -= trim_remaining
        read_start += trim_remaining if advances_read else 0
        c_operation_length = ref_step
        trim_remaining = 0
    if trim_remaining == 0:
      if ref_step <= ref_to_cover_remaining:
        new_cigar.append((op, c_operation_length))
        ref_to_cover_remaining -= ref_step
        new_read_length += c_operation_length if advances_read else 0
      else:
        c_operation_length = ref_to_cover_remaining
        new_cigar.append((op, c_operation_length))
        new_read_length += c_operation_length if advances_read else 0
        ref_to_cover_remaining = 0
        break
  return new_cigar, read_start, new_read_length


def trim_read(read: Read, region_start: int, region_end: int) -> Read:
  """TrimRead (alt_aligned_pileup_lib.cc:152-236)."""
  read_start = read.position
  trim_left = max(region_start - read_start, 0)
  ref_length = region_end - max(region_start, read_start)
  if ref_length <= 0:
    raise ValueError('CHECK_GT(ref_length, 0) failed')
  new_cigar, read_trim, new_read_length = trim_cigar(read.cigar, trim_left, ref_length)
  if not (read_trim >= 0 and read_trim + new_read_length <= len(read.aligned_sequence)):
    raise ValueError('trimmed read exceeds aligned_sequence')
  new = dataclasses.replace(read, cigar=new_cigar,
                            aligned_sequence=read.aligned_sequence[read_trim:read_trim + new_read_length],
                            aligned_quality=read.aligned_quality[read_trim:read_trim + new_read_length])
  if trim_left != 0:
    new.position = region_start
  return new


def calculate_alignment_region(variant: Variant, half_width: int, contig_n_bases: int) -> Tuple[int, int]:
  """CalculateAlignmentRegion (alt_aligned_pileup_lib.cc:238-252)."""
  ref_end = variant.start + len(variant.reference_bases)
  return max(variant.start - half_width, 0), min(contig_n_bases, ref_end + half_width)


def trim_reads(reads: Sequence[Read], region_start: int, region_end: int, min_overlap: int):
  """TrimReads (alt_aligned_pileup_lib.cc:254-276) -> (trimmed reads, original alignment positions)."""
  out, original = [], []
  for read in reads:
    t = trim_read(read, region_start, region_end)
    cigar_len = sum(ln for op, ln in t.cigar if op in _REF_ADVANCING)
    if cigar_len >= min_overlap and len(t.aligned_sequence) > 0:
      original.append(read.position)
      out.append(t)
  return out, original


# ---- the generator ----------------------------------------------------------------------------

@dataclasses.dataclass
class ExamplePlan:
  """One (candidate, alt-allele combination): everything except the pixels."""
  spec: packing.ImageSpec
  variant: Variant
  alt_combination: List[str]
  variant_type: int
  alt_specs: List[packing.ImageSpec] = dataclasses.field(default_factory=list)   # alt-aligned pileups: one per alt allele (<= 2)


class ExamplesGenerator:

  def __init__(self, options: MakeExamplesOptions, example_filenames: Optional[Dict[str, str]] = None,
               test_mode: bool = False, device: int = 0, ref_reader=None, sink=None):
    self.options = options
    # Fused mode (deepvariant_b200/fused.py; the reference's counterpart is fast_pipeline, where make_examples streams to
    # call_variants without tf.Example files): the planned images go to `sink` (packed reads, or finished images for the
    # trimmed / alt-aligned route) instead of a TFRecord writer.
    self.sink = sink
    # --stream_examples (make_examples_native.cc:724-731, stream_examples.cc): a stream_examples.StreamProducer; the examples of a region
    # go into the shard's shared-memory buffer instead of a TFRecord
    self.stream = None
    self.ssw_device: Optional[int] = None    # CUDA device for the read-to-haplotype Smith-Waterman of alt-aligned pileups (None = host)
    # (min_mapping_quality, min_base_quality, keep_legacy_allele_counter_behavior, track_ref_reads) of the allele counter that
    # produced the candidates, when they come from the very-sensitive caller over the same reads the pileups show: the table path
    # then lets the encoder derive pair_support on the device (DvbBatch.allele_begin) instead of searching read names on the host.
    self.support_options: Optional[Tuple[int, int, bool, bool]] = None
    self.last_region_derived_support = False
    pic = options.pic_options
    self.half_width = (pic.width - 1) // 2
    if not options.sample_options:
      raise ValueError('MakeExamplesOptions.sample_options is empty')
    if len(options.sample_options) > 1 and pic.alt_aligned_pileup not in ('', 'none'):
      raise NotImplementedError('alt-aligned pileups of multi-sample images')
    sample = options.sample_options[0]
    # CalculatePileupImageHeight (pileup_image_native.cc:219-240): the samples' blocks are stacked (FillPileupArrayBySample,
    # pileup_image_native.h:313-336); pileup_image_height is ONE sample's block (the first sample's in multi-sample runs)
    self.sample_heights = [s.pileup_height or pic.height for s in options.sample_options]
    self.pileup_image_height = self.sample_heights[0]
    self._device = device
    self._encoder: Optional[pi.GpuEncoder] = None
    self._sample_encoders: Dict[tuple, pi.GpuEncoder] = {}
    self.ref_reader = ref_reader
    self.writers: Dict[str, tfrecord.Writer] = {}
    self._example_filenames = dict(example_filenames or {})
    self._wrote_info = False
    if test_mode:
      return
    if self.ref_reader is None:
      from deepvariant_b200 import fasta
      self.ref_reader = fasta.IndexedFastaReader(options.reference_filename)
    for role, path in self._example_filenames.items():
      self.writers[role] = tfrecord.Writer(path)

  # -- reference window --------------------------------------------------------------------------
  def get_reference_bases_for_pileup(self, variant: Variant) -> str:
    """GetReferenceBasesForPileup (make_examples_native.cc:514-538): N-padded at contig edges,
    '' when the clipped interval is invalid."""
    n_bases = self.ref_reader.n_bases(variant.reference_name)
    start = variant.start - self.half_width
    end = start + self.options.pic_options.width
    region_start, region_end = max(0, start), min(n_bases, end)
    if not self.ref_reader.is_valid_interval(variant.reference_name, region_start, region_end):
      return ''
    ref_bases = self.ref_reader.query(variant.reference_name, region_start, region_end)
    if start < 0:
      ref_bases = 'N' * abs(start) + ref_bases
    if end > n_bases:
      ref_bases = ref_bases + 'N' * (end - n_bases)
    return ref_bases

  def _non_uniform_subset(self, sample_index: int, candidate: DeepVariantCall, query, sort_positions):
    """BuildPileupForOneSample's non-uniform branch (pileup_image_native.cc:326-341): the reads DownsampleReadIndicesWithMinsPerAllele
    keeps, in index order - a host-side filter, after which the encoder has nothing left to down-sample; when the thresholds cannot be
    met the reads go on unfiltered and the encoder's uniform shuffle decides, as in the reference."""
    sample = self.options.sample_options[sample_index]
    if not sample.use_non_uniform_downsampling:
      return query, sort_positions
    from deepvariant_b200 import sampling_util
    pic = self.options.pic_options
    max_reads = self.sample_heights[sample_index] - pic.reference_band_height
    keep = sampling_util.downsample_read_indices_with_mins_per_allele([r.key() for r in query], max_reads, candidate.allele_support,
                                                                     sample.non_uniform_downsampling_threshold, pic.random_seed)
    if keep is None:
      return query, sort_positions
    return [query[i] for i in keep], ([sort_positions[i] for i in keep] if sort_positions else sort_positions)

  # -- planning (host) ---------------------------------------------------------------------------
  def plan_region(self, candidates: Sequence[DeepVariantCall], reads: Sequence[Read], stats: Dict[str, int]) -> List[ExamplePlan]:
    """CreateAndWriteExamplesForCandidate (make_examples_native.cc:632-736) for every candidate of the
    region, up to (not including) the pixel work."""
    pic = self.options.pic_options
    sample = self.options.sample_options[0]
    plans: List[ExamplePlan] = []
    for candidate in candidates:
      variant = candidate.variant
      image_start_pos = variant.start - self.half_width
      q_start = variant.start - pic.read_overlap_buffer_bp
      q_end = variant.end + pic.read_overlap_buffer_bp
      reference_bases = self.get_reference_bases_for_pileup(variant)
      if not reference_bases:
        continue  # at the edge of the contig, example cannot be created (:650-653)
      needs_alt = need_alt_alignment(variant, pic)
      if needs_alt:
        stats['n_needs_alt_alignment'] = stats.get('n_needs_alt_alignment', 0) + 1
      use_trimmed = self.options.trim_reads_for_pileup or needs_alt or sample.keep_only_window_spanning_reads
      query = [r for r in reads if read_overlaps_region(r, variant.reference_name, q_start, q_end)]
      sort_positions = None
      if use_trimmed:
        min_overlap = pic.width if sample.keep_only_window_spanning_reads else K_DEFAULT_MINIMUM_READ_OVERLAP
        a_start, a_end = calculate_alignment_region(variant, self.half_width, self.ref_reader.n_bases(variant.reference_name))
        query, sort_positions = trim_reads(query, a_start, a_end, min_overlap)
      query, sort_positions = self._non_uniform_subset(0, candidate, query, sort_positions)
      vtype = encoded_variant_type(variant)
      for alt_combination in alt_allele_combinations(candidate, pic.multi_allelic_mode):
        spec = packing.image_spec_for(candidate, reference_bases, query, image_start_pos, alt_combination, pic,
                                      sort_positions=sort_positions)
        plan = ExamplePlan(spec, variant, list(alt_combination), vtype)
        if needs_alt:
          plan.alt_specs = self.alt_aligned_specs(candidate, alt_combination, query, sort_positions, image_start_pos)
        plans.append(plan)
    return plans

  # -- alt-aligned pileups (CreateAltAlignedImages, make_examples_native.cc:553-626) -------------------------------------------
  def create_haplotype(self, variant: Variant, alt: str) -> Tuple[str, int, int]:
    """CreateHaplotype (:269-297): the reference window around the variant with `alt` in place of the reference allele."""
    n_bases = self.ref_reader.n_bases(variant.reference_name)
    var_end = variant.start + len(variant.reference_bases)
    ref_start = max(variant.start - self.half_width, 0)
    ref_end = min(n_bases, var_end + self.half_width)
    prefix = self.ref_reader.query(variant.reference_name, ref_start, variant.start) if ref_start < variant.start else ''
    suffix = self.ref_reader.query(variant.reference_name, var_end, ref_end) if ref_end > var_end else ''
    return prefix + alt + suffix, ref_start, ref_end

  def realign_reads_to_haplotype(self, haplotype: str, reads: Sequence[Read], contig: str, ref_start: int) -> List[Read]:
    """RealignReadsToHaplotype (alt_aligned_pileup_lib.cc:278-313; kRefAlignMargin = 0, so the aligner's reference IS the haplotype):
    forced alignment of every trimmed read; reads that cannot be placed come back empty."""
    from deepvariant_b200 import fast_pass_aligner
    aligner = fast_pass_aligner.FastPassAligner()
    a = self.options.aln_config
    aligner.set_options(kmer_size=a['kmer_size'], read_size=len(reads[0].aligned_sequence) if reads and len(reads[0].aligned_sequence) > 15 else 200,
                        max_num_of_mismatches=a['max_num_of_mismatches'], realignment_similarity_threshold=a['realignment_similarity_threshold'],
                        match=a['match'], mismatch=a['mismatch'], gap_open=a['gap_open'], gap_extend=a['gap_extend'], force_alignment=True)
    aligner.ssw_device = self.ssw_device     # batched Smith-Waterman on the GPU when the stage runs on one (cli.make_examples sets it)
    aligner.reference = haplotype
    aligner.region_position_in_chr = ref_start
    aligner.ref_prefix_len = aligner.ref_suffix_len = 0
    aligner.haplotypes = [haplotype]
    return aligner.align_reads(reads)

  def alt_aligned_specs(self, candidate: DeepVariantCall, alt_combination: Sequence[str], trimmed_reads: Sequence[Read],
                        original_start_positions: Optional[Sequence[int]], image_start_pos: int) -> List[packing.ImageSpec]:
    pic = self.options.pic_options
    specs: List[packing.ImageSpec] = []
    if len(alt_combination) > 2:
      raise ValueError('an alt combination has at most two alleles')
    for alt in alt_combination:
      haplotype, ref_start, _ = self.create_haplotype(candidate.variant, alt)
      if len(haplotype) < pic.width:
        break
      realigned = self.realign_reads_to_haplotype(haplotype, trimmed_reads, candidate.variant.reference_name, ref_start)
      keep = [i for i, r in enumerate(realigned) if r.aligned_sequence]
      specs.append(packing.image_spec_for(candidate, haplotype[:pic.width], [realigned[i] for i in keep], image_start_pos, alt_combination, pic,
                                          sort_positions=[original_start_positions[i] for i in keep]))
    return specs

  def plan_region_from_table(self, candidates: Sequence[DeepVariantCall], table, stats: Dict[str, int],
                             region: Optional[Tuple[str, int, int]] = None):
    """plan_region() over a bam.NativeBamTable: the per-candidate read query is an index computation on the table's
    position arrays and the per-image read lists are row numbers — no Read objects.
    `region` = (contig, start, end): only reads overlapping it are considered, as when the reference hands
    WriteExamplesInRegion the reads of one partition.  Returns (plans, table image specs).  Candidates that need
    trimmed reads (PACBIO / alt-aligned) are not handled here: use plan_region() with table.query()."""
    pic = self.options.pic_options
    sample = self.options.sample_options[0]
    if self.options.trim_reads_for_pileup or sample.keep_only_window_spanning_reads:
      raise NotImplementedError('the table path covers untrimmed reads; use plan_region() for trimmed pileups')
    plans: List[ExamplePlan] = []
    specs: List[packing.TableImageSpec] = []
    region_rows = table.query_indices(*region) if region is not None else None
    for candidate in candidates:
      variant = candidate.variant
      if need_alt_alignment(variant, pic):
        raise NotImplementedError('candidate needs alt-aligned (trimmed) reads; use plan_region()')
      image_start_pos = variant.start - self.half_width
      reference_bases = self.get_reference_bases_for_pileup(variant)
      if not reference_bases:
        continue
      q_start, q_end = variant.start - pic.read_overlap_buffer_bp, variant.end + pic.read_overlap_buffer_bp
      if region_rows is None:
        rows = table.query_indices(variant.reference_name, q_start, q_end)
      elif region[0] != variant.reference_name:
        rows = region_rows[:0]
      else:
        rows = region_rows[(table.pos[region_rows] < q_end) & (table.end[region_rows] > q_start)]
      key_to_local = {}
      for j, r in enumerate(rows):
        key_to_local.setdefault(table.names[int(table.name_begin[r]):int(table.name_begin[r + 1])].decode() + '/' +
                                str(int(table.read_number[r])), []).append(j)
      groups = packing.table_allele_groups(candidate, key_to_local, len(rows)) if pic.sort_by_alt_allele_support else None
      vtype = encoded_variant_type(variant)
      for alt_combination in alt_allele_combinations(candidate, pic.multi_allelic_mode):
        support = packing.table_support(candidate, key_to_local, len(rows), alt_combination)
        specs.append(packing.TableImageSpec(reference_bases, image_start_pos, variant.start, rows, support, groups))
        plans.append(ExamplePlan(None, variant, list(alt_combination), vtype))
    return plans, specs

  def plan_region_native(self, candidates: Sequence[DeepVariantCall], table, region: Tuple[str, int, int]):
    """plan_region_from_table() with the per-read work left to the C++ region packer: Python enumerates the images
    (candidate x alt combination), fetches reference windows and flattens allele_support; the read query, the
    read-name search and the gathers happen in dvb_pack_region_from_bam.  Returns (plans, packing.RegionImage list)."""
    pic = self.options.pic_options
    sample = self.options.sample_options[0]
    if self.options.trim_reads_for_pileup or sample.keep_only_window_spanning_reads:
      raise NotImplementedError('the table path covers untrimmed reads; use plan_region() for trimmed pileups')
    plans: List[ExamplePlan] = []
    images: List[packing.RegionImage] = []
    ref_index = {name: i for i, name in enumerate(table.references)}
    # device-side support: the alt alleles go to the encoder as read-allele keys and the read names stay on the host unread
    derive = self.support_options is not None
    if derive:
      keys_of = [[packing.read_allele_key(c.variant.reference_bases, alt) for alt in c.variant.alternate_bases] for c in candidates]
      derive = all(k is not None for ks in keys_of for k in ks)
    self.last_region_derived_support = derive
    for ci, candidate in enumerate(candidates):
      variant = candidate.variant
      if need_alt_alignment(variant, pic):
        raise NotImplementedError('candidate needs alt-aligned (trimmed) reads; use plan_region()')
      reference_bases = self.get_reference_bases_for_pileup(variant)
      if not reference_bases:
        continue
      rb = reference_bases.encode() if isinstance(reference_bases, str) else bytes(reference_bases)
      alts = list(variant.alternate_bases)
      enc = [[] if derive else [name.encode() for name in (candidate.allele_support.get(alt) or ())] for alt in alts]
      ref_run = self.canonical_run_after(variant.reference_name, variant.start) if derive else 0
      counts = np.array([len(e) for e in enc], dtype=np.int64)
      flat = [k for e in enc for k in e]
      blob = b''.join(flat)
      key_lens = np.fromiter((len(k) for k in flat), dtype=np.int64, count=len(flat))
      groups = np.repeat(np.arange(len(alts), dtype=np.uint8), counts) if pic.sort_by_alt_allele_support else None
      vtype = encoded_variant_type(variant)
      rid = ref_index.get(variant.reference_name, -2)
      for alt_combination in alt_allele_combinations(candidate, pic.multi_allelic_mode):
        classes = np.repeat(np.array([1 if alt in alt_combination else 2 for alt in alts], dtype=np.uint8), counts)
        images.append(packing.RegionImage(rid, variant.start, variant.end, variant.start - self.half_width, rb, blob, key_lens,
                                          classes, groups, len(alts)))
        if derive:
          images[-1].alleles = [(k[0], k[1], 1 if alt in alt_combination else 2, j) for j, (alt, k) in enumerate(zip(alts, keys_of[ci]))]
          images[-1].ref_run = ref_run
        plans.append(ExamplePlan(None, variant, list(alt_combination), vtype))
    return plans, images

  def canonical_run_after(self, contig: str, position: int) -> int:
    """Number of consecutive canonical (ACGT) in-contig reference bases after `position`: a deletion anchored there is a usable
    read allele iff it is not longer (MakeIndelReadAllele, allelecounter.cc:449-456).  Capped at 2^20."""
    n_bases = self.ref_reader.n_bases(contig)
    run, span = 0, 1024
    while run < (1 << 20):
      lo = position + 1 + run
      hi = min(n_bases, lo + span)
      if hi <= lo:
        break
      m = _NON_CANONICAL.search(self.ref_reader.query(contig, lo, hi).upper())
      if m is not None:
        return run + m.start()
      run += hi - lo
      span *= 8
    return run

  def pack_region_native(self, candidates: Sequence[DeepVariantCall], table, region: Tuple[str, int, int]):
    """(plans, PackedBatch) of a region through the C++ packer."""
    pic = self.options.pic_options
    plans, images = self.plan_region_native(candidates, table, region)
    rid = table.references.index(region[0]) if region[0] in table.references else -3
    params = pi.to_params(pic, height=self.pileup_image_height)   # host-side struct; no device needed to pack
    packed = packing.pack_region_native(table, images, rid, region[1], region[2], pic.read_overlap_buffer_bp,
                                        params, with_groups=bool(pic.sort_by_alt_allele_support))
    if self.last_region_derived_support:
      mq, bq, legacy, track = self.support_options
      packing.attach_alleles(packed, images, mq, bq, legacy, track, with_groups=bool(pic.sort_by_alt_allele_support))
    return plans, packed

  def write_examples_in_region_from_table(self, candidates: Sequence[DeepVariantCall], table, role: str,
                                          region: Optional[Tuple[str, int, int]] = None, native_packer: Optional[bool] = None):
    """WriteExamplesInRegion with the reads given as a native BAM table (one CUDA launch for the region).  With a
    region and an open table the C++ region packer builds the batch (native_packer=False forces the numpy packer;
    both give identical arrays, tests/test_bam_native.py)."""
    if role not in self.writers and self.sink is None:
      raise KeyError(f'Role {role} does not have a writer.')
    stats: Dict[str, int] = {}
    enc = self._gpu()
    if native_packer is None:
      native_packer = region is not None and getattr(table, '_handle', None) is not None
    if native_packer:
      plans, packed = self.pack_region_native(candidates, table, region)
    else:
      plans, specs = self.plan_region_from_table(candidates, table, stats, region)
      packed = packing.pack_images_from_table(specs, table, enc.params) if plans else None
    if self.sink is not None:
      self._count_examples(plans, stats)
      if plans:
        self.sink.add_packed(plans, packed)
      return stats, self.image_shape()
    images = enc.encode_host(packed) if plans else \
        np.zeros((0,) + enc.shape, dtype=np.uint8)
    for rec in self.finish_region(plans, images, stats):
      self.writers[role].write(rec)
    return stats, self.image_shape()

  # -- serialisation (host) ----------------------------------------------------------------------
  def image_shape(self) -> List[int]:
    pic = self.options.pic_options
    # CalculatePileupImageHeight (pileup_image_native.cc:218-240): 'rows' stacks the two alt-aligned pileups under the main one
    sections = {'rows': 3, 'single_row': 2}.get(pic.alt_aligned_pileup, 1)
    if len(self.options.sample_options) > 1:
      return [sum(self.sample_heights), pic.width, len(pic.channels)]
    return [self.pileup_image_height * sections, pic.width, len(pic.channels)]

  def encode_example(self, plan: ExamplePlan, image: np.ndarray, stats: Dict[str, int]) -> bytes:
    """EncodeExample (make_examples_native.cc:388-474), calling mode (no label)."""
    variant = plan.variant
    shape = self.image_shape()
    if list(image.shape) != shape:
      raise ValueError(f'image shape {image.shape} != {shape}')
    alt_indices_encoded, _ = encode_alt_alleles(variant, plan.alt_combination)
    locus = f'{variant.reference_name}:{variant.start + 1}-{variant.end}'
    features = {
        'alt_allele_indices/encoded': ('bytes', [alt_indices_encoded]),
        'image/encoded': ('bytes', [image.tobytes()]),
        'image/shape': ('int64', shape),
        'locus': ('bytes', [locus.encode()]),
        'sequencing_type': ('int64', [self.options.pic_options.sequencing_type]),
        'variant/encoded': ('bytes', [variant.serialize()]),
        'variant_type': ('int64', [plan.variant_type]),
    }
    stats['n_examples'] = stats.get('n_examples', 0) + 1   # UpdateStats (:330-348)
    if plan.variant_type == 2:
      stats['n_indels'] = stats.get('n_indels', 0) + 1
    else:
      stats['n_snps'] = stats.get('n_snps', 0) + 1
    return protos.encode_tf_example(features)

  @staticmethod
  def _count_examples(plans: Sequence[ExamplePlan], stats: Dict[str, int]) -> None:
    """UpdateStats (make_examples_native.cc:330-348) for plans that go to the fused sink instead of encode_example()."""
    for p in plans:
      stats['n_examples'] = stats.get('n_examples', 0) + 1
      key = 'n_indels' if p.variant_type == 2 else 'n_snps'
      stats[key] = stats.get(key, 0) + 1

  def finish_region(self, plans: Sequence[ExamplePlan], images: np.ndarray, stats: Dict[str, int]) -> List[bytes]:
    if self.stream is not None:
      # WriteExamplesInRegion with stream_examples (make_examples_native.cc:757-790): StartStreaming, one StreamExample per example
      # (alt_allele_indices, variant, image), EndStreaming(data_written)
      self.stream.start_streaming()
      for i, p in enumerate(plans):
        self.stream.stream_example(encode_alt_alleles(p.variant, p.alt_combination)[0], p.variant.serialize(), images[i])
      self.stream.end_streaming(bool(plans))
      self._count_examples(plans, stats)
      return []
    return [self.encode_example(p, images[i], stats) for i, p in enumerate(plans)]

  # -- the pybind entry point --------------------------------------------------------------------
  def _gpu(self) -> pi.GpuEncoder:
    if self._encoder is None:
      self._encoder = pi.GpuEncoder(pi.to_params(self.options.pic_options, height=self.pileup_image_height), self._device)
    return self._encoder

  def encode_plans(self, plans: Sequence[ExamplePlan]) -> np.ndarray:
    """All images of the region in one CUDA launch (dvb_encode_batch_host)."""
    enc = self._gpu()
    if not plans:
      return np.zeros((0,) + enc.shape, dtype=np.uint8)
    specs = [p.spec for p in plans]
    alt_at = []                       # per plan: indices of its alt-aligned images in the batch
    for p in plans:
      alt_at.append(list(range(len(specs), len(specs) + len(p.alt_specs))))
      specs += p.alt_specs
    images = enc.encode_host(packing.pack_images(specs, enc.params))
    return compose_alt_aligned(images, len(plans), alt_at, self.options.pic_options, [p.alt_combination for p in plans])

  # -- multi-sample images (DeepTrio / DeepSomatic layouts) -------------------------------------------------------------------
  def _sample_encoder(self, sample_index: int, mean_coverage: float):
    """The encoder handle of one sample's block: its own height, channels_enum_to_blank and mean coverage
    (BuildPileupForOneSample's sample_options / mean_coverage / channels_enum_to_blank arguments, pileup_image_native.cc:296-304)."""
    sample = self.options.sample_options[sample_index]
    key = (self.sample_heights[sample_index], tuple(sorted(sample.channels_enum_to_blank)), float(mean_coverage))
    if key not in self._sample_encoders:
      pic = dataclasses.replace(self.options.pic_options, channels_enum_to_blank=key[1], mean_coverage=key[2])
      self._sample_encoders[key] = self._make_encoder(pi.to_params(pic, height=key[0]))
    return key, self._sample_encoders[key]

  def _make_encoder(self, params):
    return pi.GpuEncoder(params, self._device)

  def plan_region_by_sample(self, candidates: Sequence[DeepVariantCall], reads_per_sample: Sequence[Sequence[Read]], sample_order: Sequence[int],
                            stats: Dict[str, int]):
    """CreateAndWriteExamplesForCandidate (make_examples_native.cc:632-736) with several samples: for every candidate and alt
    combination one BuildPileupForOneSample per sample of `sample_order`, all from the same DeepVariantCall.  Returns (plans,
    specs_of) with specs_of[i] = [(sample index, ImageSpec), ...] in stacking order."""
    pic = self.options.pic_options
    plans: List[ExamplePlan] = []
    specs_of: List[List[Tuple[int, packing.ImageSpec]]] = []
    for candidate in candidates:
      variant = candidate.variant
      image_start_pos = variant.start - self.half_width
      q_start = variant.start - pic.read_overlap_buffer_bp
      q_end = variant.end + pic.read_overlap_buffer_bp
      reference_bases = self.get_reference_bases_for_pileup(variant)
      if not reference_bases:
        continue
      vtype = encoded_variant_type(variant)
      queries = {}
      for alt_combination in alt_allele_combinations(candidate, pic.multi_allelic_mode):
        per_sample = []
        for this_sample in sample_order:
          sample = self.options.sample_options[this_sample]
          if this_sample not in queries:
            query = [r for r in reads_per_sample[this_sample] if read_overlaps_region(r, variant.reference_name, q_start, q_end)]
            sort_positions = None
            if self.options.trim_reads_for_pileup or sample.keep_only_window_spanning_reads:
              min_overlap = pic.width if sample.keep_only_window_spanning_reads else K_DEFAULT_MINIMUM_READ_OVERLAP
              a_start, a_end = calculate_alignment_region(variant, self.half_width, self.ref_reader.n_bases(variant.reference_name))
              query, sort_positions = trim_reads(query, a_start, a_end, min_overlap)
            queries[this_sample] = self._non_uniform_subset(this_sample, candidate, query, sort_positions)
          query, sort_positions = queries[this_sample]
          per_sample.append((this_sample, packing.image_spec_for(candidate, reference_bases, query, image_start_pos, alt_combination, pic,
                                                                 sort_positions=sort_positions)))
        plans.append(ExamplePlan(None, variant, list(alt_combination), vtype))
        specs_of.append(per_sample)
    return plans, specs_of

  def encode_plans_by_sample(self, specs_of, mean_coverage_per_sample: Optional[Sequence[float]] = None) -> np.ndarray:
    """FillPileupArrayBySample (pileup_image_native.h:313-336): each sample's block from its own encoder handle (one CUDA launch per
    distinct (height, blanked channels, mean coverage)), stacked in sample order."""
    shape = self.image_shape()
    out = np.zeros((len(specs_of),) + tuple(shape), dtype=np.uint8)
    groups: Dict[tuple, list] = {}
    for i, per_sample in enumerate(specs_of):
      row0 = 0
      for this_sample, spec in per_sample:
        cov = mean_coverage_per_sample[this_sample] if mean_coverage_per_sample else 0.0
        key, _ = self._sample_encoder(this_sample, cov)
        groups.setdefault(key, []).append((i, row0, spec))
        row0 += self.sample_heights[this_sample]
      if row0 != shape[0]:
        raise ValueError(f'the samples of sample_order stack to {row0} rows, the image has {shape[0]}')
    for key, items in groups.items():
      enc = self._sample_encoders[key]
      images = enc.encode_host(packing.pack_images([spec for _, _, spec in items], enc.params))
      for (i, row0, _), img in zip(items, images):
        out[i, row0:row0 + key[0]] = img
    return out

  def write_examples_in_region(self, candidates: Sequence[DeepVariantCall], reads_per_sample: Sequence[Sequence[Read]],
                               sample_order: Sequence[int], role: str,
                               mean_coverage_per_sample: Optional[Sequence[float]] = None):
    """WriteExamplesInRegion (make_examples_native.cc:742-793)."""
    if role not in self.writers and self.sink is None:
      raise KeyError(f'Role {role} does not have a writer.')
    stats: Dict[str, int] = {}
    if len(self.options.sample_options) > 1:
      plans, specs_of = self.plan_region_by_sample(candidates, reads_per_sample, sample_order, stats)
      images = self.encode_plans_by_sample(specs_of, mean_coverage_per_sample)
    else:
      reads = list(reads_per_sample[sample_order[0]]) if reads_per_sample else []
      plans = self.plan_region(candidates, reads, stats)
      images = self.encode_plans(plans)
    if self.sink is not None:
      self._count_examples(plans, stats)
      if plans:
        self.sink.add_images(plans, images)
      return stats, self.image_shape()
    for rec in self.finish_region(plans, images, stats):
      self.writers[role].write(rec)
    return stats, self.image_shape()

  def signal_shard_finished(self) -> None:
    if self.stream is not None:
      self.stream.signal_shard_finished()   # StreamExamples::SignalShardFinished
    for role, w in self.writers.items():
      w.close()
      if role not in self._example_filenames:
        continue              # streamed examples: no file, no example_info.json beside it
      write_example_info_json(self._example_filenames[role], self.image_shape(),
                              example_info_channels(self.options.pic_options))
    self.writers = {}


def compose_alt_aligned(images: np.ndarray, n_plans: int, alt_at: Sequence[Sequence[int]], pic,
                        alt_combinations: Optional[Sequence[Sequence[str]]] = None) -> np.ndarray:
  """FillPileupArray (pileup_image_native.h:214-308).  diff_channels / base_channels: the two extra channels of an example are
  channel 5 (base_differs_from_ref; 0 = read_base for base_channels) of its first and second alt-aligned pileup, row by row; with a
  single alt-aligned pileup both channels carry it; without any they stay zero.  rows / single_row: the alt-aligned pileups
  (GetAltImageRowIndices, pileup_image_native.cc:192-208: both; or the one of the longer alt) are stacked under the main pileup,
  zeros where there is none."""
  out = images[:n_plans]
  if pic.alt_aligned_pileup in ('rows', 'single_row'):
    h = images.shape[1]
    sections = 3 if pic.alt_aligned_pileup == 'rows' else 2
    stacked = np.zeros((n_plans, h * sections) + images.shape[2:], dtype=images.dtype)
    stacked[:, :h] = out
    for i, idx in enumerate(alt_at):
      if pic.alt_aligned_pileup == 'rows':
        wanted = [0, 1]
      else:
        combo = alt_combinations[i] if alt_combinations is not None else []
        wanted = [1] if len(combo) == 2 and len(combo[1]) > len(combo[0]) else [0]
      for j, w in enumerate(wanted):
        if w < len(idx):
          stacked[i, h * (j + 1):h * (j + 2)] = images[idx[w]]
    return stacked
  n_alt = sum(1 for c in pic.channels if c in pi.ALT_ALIGNED_PSEUDO_CHANNELS)
  if n_alt != 2 or not any(alt_at):
    return out
  c0 = images.shape[-1] - 2
  src = 5 if pic.alt_aligned_pileup == 'diff_channels' else 0
  for i, idx in enumerate(alt_at):
    if not idx:
      continue
    out[i, ..., c0] = images[idx[0], ..., src]
    out[i, ..., c0 + 1] = images[idx[1] if len(idx) > 1 else idx[0], ..., src]
  return out


def partition_candidates(candidates: Sequence[DeepVariantCall], partition_size: int, region: Optional[Tuple[str, int, int]] = None):
  """Groups candidates by genomic partition (make_examples' --partition_size regions,
  make_examples_core.processing_regions_from_options :3341): returns a sorted list of
  ((contig, partition_index, origin), [candidates])."""
  by_part = {}
  for c in candidates:
    v = c.variant
    if region and not (v.reference_name == region[0] and region[1] <= v.start < region[2]):
      continue
    origin = region[1] if region else 0
    by_part.setdefault((v.reference_name, (v.start - origin) // partition_size, origin), []).append(c)
  return sorted(by_part.items())


def shard_partitions(partitions, n_shards: int, task: int):
  """Partition i goes to task i mod N — the region sharding of `make_examples --task i` (scripts/run_deepvariant.py:457-497)
  and of the per-GPU ranks: disjoint, exhaustive, no exchange between shards."""
  return [p for i, p in enumerate(partitions) if i % n_shards == task]


def example_info_channels(pic: pi.PileupImageOptions) -> List[int]:
  """Channel enums for example_info.json (make_examples_core.py:3755-3774): computed channels, then
  the alt-aligned pseudo channels (9,10 / 20,21)."""
  return pi.all_channels_enum(pic, pic.alt_aligned_pileup if pic.alt_aligned_pileup in ('diff_channels', 'base_channels') else '')


def write_example_info_json(examples_path: str, shape: Sequence[int], channels: Sequence[int]) -> str:
  path = examples_path + '.example_info.json'
  with open(path, 'w') as f:
    json.dump({'version': VERSION, 'shape': list(shape), 'channels': list(channels)}, f)
  return path
