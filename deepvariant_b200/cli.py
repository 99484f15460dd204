"""Command-line surface kept from the reference (scripts/run_deepvariant.py:430-528 builds exactly these commands):

  python -m deepvariant_b200.cli make_examples --mode calling --ref REF --reads BAM --examples X@N.gz
         [--candidates CANDS.tfrecord.gz (output) --task i --regions chr:start-end --channel_list ... --pileup_image_width W
         --pileup_image_height H --min_mapping_quality q --min_base_quality q --partition_size 1000 --vsc_min_count_snps 2
         --vsc_min_fraction_snps 0.12 --vsc_min_count_indels 2 --vsc_min_fraction_indels 0.06 --track_ref_reads --phase_reads
         --sort_by_haplotypes --trim_reads_for_pileup --alt_aligned_pileup diff_channels --norealign_reads]
  python -m deepvariant_b200.cli call_variants --examples X@N.gz --outfile Y.tfrecord.gz --checkpoint M [--batch_size 1024]
  python -m deepvariant_b200.cli postprocess_variants --ref REF --infile Y.tfrecord.gz --outfile OUT.vcf[.gz]
  python -m deepvariant_b200.cli run_deepvariant --model_type WGS --ref REF --reads BAM --output_dir D [--output_vcf OUT.vcf.gz]

make_examples realigns the reads (deepvariant_b200/realigner.py; --norealign_reads turns it off), finds its candidates itself (allele
counter + very-sensitive caller, deepvariant_b200/candidates.py) and phases the reads when --phase_reads is set (direct_phasing.py); `--candidates_in` (not a reference flag) imports a DeepVariantCall
TFRecord written by another make_examples instead.
"""
from __future__ import annotations

import argparse
import os
import re
import sys
from typing import List

MODEL_DEFAULTS = {   # scripts/run_deepvariant.py + model.example_info.json flags_for_calling
    'WGS': dict(channel_list='BASE_CHANNELS,insert_size', pileup_image_width=221),
    'WES': dict(channel_list='BASE_CHANNELS,insert_size', pileup_image_width=221),
    'PACBIO': dict(channel_list='BASE_CHANNELS,haplotype,supplementary_alignment', pileup_image_width=147, sort_by_haplotypes=True,
                   trim_reads_for_pileup=True, alt_aligned_pileup='diff_channels', min_mapping_quality=1, partition_size=25000,
                   parse_sam_aux_fields=True, track_ref_reads=True, phase_reads=True, max_reads_per_partition=600, norealign_reads=True,
                   vsc_min_fraction_indels=0.12),       # flags_for_calling of the released PACBIO model's example_info.json
}


def _channels(spec: str) -> List[str]:
  from deepvariant_b200 import pileup_image as pi
  out: List[str] = []
  for c in spec.split(','):
    c = c.strip()
    if c == 'BASE_CHANNELS':
      out += pi.PILEUP_DEFAULT_CHANNELS
    elif c:
      out.append(c)
  return out


def parse_region(s: str):
  m = re.match(r'^([^:]+):([\d,]+)-([\d,]+)$', s)
  if not m:
    raise ValueError(f'bad region {s}')
  return m.group(1), int(m.group(2).replace(',', '')) - 1, int(m.group(3).replace(',', ''))


def parse_regions(text: str):
  """--regions as the reference takes it (make_examples_options.py: space-separated literals `chr20:10,000-20,000`, whole contigs
  `chr20`, and BED files): a list of 0-based half-open (contig, start, end)."""
  from deepvariant_b200 import postprocess_variants as pp
  out = []
  for token in text.split():
    # a region literal or a contig name first: a contig that happens to be called like a local file or directory is not a BED
    is_literal = re.fullmatch(r'[^:\s]+:[\d,]+(-[\d,]+)?', token) is not None
    if token.endswith(('.bed', '.bed.gz')) or (not is_literal and os.path.isfile(token)):
      if token.endswith('.gz'):
        import gzip
        out += [(p[0], int(p[1]), int(p[2])) for p in (l.split() for l in gzip.open(token, 'rt')) if len(p) >= 3 and not p[0].startswith(('#', 'track', 'browser'))]
      else:
        out += pp.read_bed(token)
    elif ':' in token:
      out.append(parse_region(token))
    else:
      out.append((token, 0, 1 << 40))
  return out


# The reference skips the decoys, alternate loci, unplaced / unlocalised scaffolds, EBV and HLA contigs of the common human references
# (deepvariant/exclude_contigs.py EXCLUDED_HUMAN_CONTIGS: 3402 names of hs37d5 and the GRCh38 analysis set).  Every one of those names
# has one of these shapes; the shapes are used instead of the table.
_EXCLUDED_CONTIG = re.compile(r'^(GL000\d{3}\.1|NC_007605|hs37d5|chrEBV|HLA-.*|chrUn_.*|chr(\d+|X|Y)_[A-Z]{2}\d+v\d+_(alt|random)|.*_decoy)$')


def shared_contigs(ref_contigs, reads_contigs, min_coverage_fraction: float = 0.9):
  """_ensure_consistent_contigs / common_contigs / validate_reference_contig_coverage (deepvariant/make_examples_core.py:540-640): the
  reference contigs (minus the excluded ones) that the reads' header has too, with the same length; an error when they cover less than
  `min_coverage_fraction` of the non-excluded reference bases (reads aligned to another build)."""
  ref_contigs = [(c, n) for c, n in ref_contigs if not _EXCLUDED_CONTIG.match(c)]
  shared = [(c, n) for c, n in ref_contigs if reads_contigs.get(c) == n]
  ref_bp, common_bp = sum(n for _, n in ref_contigs), sum(n for _, n in shared)
  coverage = common_bp / ref_bp if ref_bp else 0.0
  if not shared or coverage < min_coverage_fraction:
    raise ValueError(f'Reference contigs span {ref_bp} bases but only {common_bp} bases ({100 * coverage:.2f}%) were found in common among our input '
                     f'files. Check that the sources were created on a common genome reference build. Contig matches were: '
                     + ', '.join(f'"{c}" is {n} bp and {"matches" if reads_contigs.get(c) == n else "IS MISSING" if c not in reads_contigs else "has another length"}' for c, n in ref_contigs[:30]))
  return shared


class _NullWriter:
  """--stream_examples: finish_region hands the examples to the stream and returns no records."""

  def write(self, rec):
    raise AssertionError('a streamed example reached the TFRecord writer')

  def close(self):
    pass


def _lib_plane_enums():
  from deepvariant_b200 import _lib
  return set(_lib.PLANE_OF_CHANNEL)


RUNTIME_BY_REGION_COLUMNS = ('region', 'get reads', 'find candidates', 'make pileup images', 'write outputs', 'num reads', 'num candidates', 'num examples',
                             'small model generate examples', 'small model call examples', 'small model write variants', 'small model total')

MAKE_EXAMPLES_DEFAULTS = dict(
    task=0, regions='', channel_list='BASE_CHANNELS', pileup_image_width=221, pileup_image_height=100, min_mapping_quality=5,
    min_base_quality=10, partition_size=1000, sort_by_haplotypes=False, trim_reads_for_pileup=False, parse_sam_aux_fields=False,
    alt_aligned_pileup='none', min_shared_contigs_basepairs=0.9, stream_examples=False, shm_prefix='', population_vcfs='', mean_coverage_per_sample='', device=0, checkpoint='', checkpoint_json='', candidates='', candidates_in='', max_reads_per_partition=1500,
    sample_name='', vsc_min_count_snps=2, vsc_min_count_indels=2, vsc_min_fraction_snps=0.12, vsc_min_fraction_indels=0.06,
    vsc_min_fraction_multiplier=1.0, small_model_vaf_context_window_size=0, track_ref_reads=False, phase_reads=False,
    keep_legacy_allele_counter_behavior=False, normalize_reads=False, realign_reads=True, gvcf='', gvcf_gq_binsize=5, p_error=0.001,
    include_med_dp=False, haploid_contigs='', par_regions_bed='', candidate_positions='', runtime_by_region='', examples='', call_variants_outfile='', precision=1,
    call_batch_size=2048, variant_caller='very_sensitive_caller', proposed_variants='')      # --realign_reads defaults to true (make_examples_options.py:229)


def model_example_info_json_path(checkpoint: str, checkpoint_json: str = '') -> str:
  """get_model_example_info_json_path (deepvariant/make_examples_core.py:3780-3822): --checkpoint_json wins; a saved-model
  directory holds model.example_info.json (or example_info.json); for a ckpt path the file sits next to the checkpoint."""
  if checkpoint_json:
    return checkpoint_json
  if not checkpoint:
    return ''
  dirs = [checkpoint] if os.path.isdir(checkpoint) else [os.path.dirname(checkpoint)]
  for d in dirs:
    for name in ('model.example_info.json', 'example_info.json'):
      if os.path.exists(os.path.join(d, name)):
        return os.path.join(d, name)
    if os.path.isdir(d):
      for name in sorted(os.listdir(d)):
        if name.endswith('example_info.json'):
          return os.path.join(d, name)
  return ''


def apply_flags_for_calling(cli_values: dict, checkpoint: str, checkpoint_json: str = '') -> dict:
  """apply_flags_for_calling (deepvariant/make_examples_core.py:3825-3920): command line > the model's
  model.example_info.json 'flags_for_calling' > defaults.  Flags this implementation does not know (candidate generation,
  realigner, small model ...) belong to upstream stages and are reported, not applied."""
  import json
  merged = dict(MAKE_EXAMPLES_DEFAULTS)
  path = model_example_info_json_path(checkpoint, checkpoint_json)
  ignored = []
  if path:
    try:
      flags_map = json.load(open(path)).get('flags_for_calling', {})
    except (OSError, ValueError):
      flags_map = {}
    for k, v in flags_map.items():
      if k in MAKE_EXAMPLES_DEFAULTS:
        want = type(MAKE_EXAMPLES_DEFAULTS[k])
        merged[k] = (str(v).lower() in ('1', 'true', 'yes')) if want is bool and not isinstance(v, bool) else want(v)
      else:
        ignored.append(k)
  merged.update(cli_values)
  merged['_ignored_flags_for_calling'] = ignored
  return merged


def make_examples(argv):
  ap = argparse.ArgumentParser('make_examples', argument_default=argparse.SUPPRESS)
  ap.add_argument('--mode', default='calling', choices=['calling', 'candidate_sweep'])
  ap.add_argument('--runtime_by_region')          # TSV: seconds per stage and counts, one line per region (sharded like --examples)
  ap.add_argument('--candidate_positions')       # candidate_sweep: int32 positions out (sharded like --examples); calling: partitions cut by them
  ap.add_argument('--ref', required=True)
  ap.add_argument('--reads', required=True)
  ap.add_argument('--examples')                   # tf.Example shards out (the staged flow); optional with --call_variants_outfile
  ap.add_argument('--call_variants_outfile')      # FUSED flow (deepvariant_b200/fused.py; cf. the reference's fast_pipeline): pileups are encoded AND classified
                                                  # here, CallVariantsOutput shards out (name@N.tfrecord.gz), no tf.Example files; needs --checkpoint
  ap.add_argument('--precision', type=int, choices=[0, 1])
  ap.add_argument('--call_batch_size', type=int)
  ap.add_argument('--candidates')        # OUTPUT, as in the reference (make_examples_options.py:109): the DeepVariantCalls found
  ap.add_argument('--candidates_in')     # INPUT (not a reference flag): use these DeepVariantCalls instead of generating them
  # make_examples_options.py:1420-1490: 'very_sensitive_caller' (default) or 'vcf_candidate_importer' with --proposed_variants
  ap.add_argument('--variant_caller', choices=['very_sensitive_caller', 'vcf_candidate_importer'])
  ap.add_argument('--proposed_variants')
  ap.add_argument('--checkpoint')        # model directory / ckpt path: only its example_info.json flags are read here
  ap.add_argument('--checkpoint_json')
  ap.add_argument('--task', type=int)
  ap.add_argument('--regions')
  ap.add_argument('--channel_list')
  ap.add_argument('--pileup_image_width', type=int)
  ap.add_argument('--pileup_image_height', type=int)
  ap.add_argument('--min_mapping_quality', type=int)
  ap.add_argument('--min_base_quality', type=int)
  ap.add_argument('--partition_size', type=int)
  ap.add_argument('--max_reads_per_partition', type=int)
  ap.add_argument('--sample_name')
  ap.add_argument('--vsc_min_count_snps', type=int)
  ap.add_argument('--vsc_min_count_indels', type=int)
  ap.add_argument('--vsc_min_fraction_snps', type=float)
  ap.add_argument('--vsc_min_fraction_indels', type=float)
  ap.add_argument('--vsc_min_fraction_multiplier', type=float)
  ap.add_argument('--small_model_vaf_context_window_size', type=int)
  ap.add_argument('--track_ref_reads', action='store_true')
  ap.add_argument('--phase_reads', action='store_true')
  ap.add_argument('--keep_legacy_allele_counter_behavior', action='store_true')
  ap.add_argument('--normalize_reads', action='store_true')
  ap.add_argument('--gvcf')                       # non-variant site records (Variant protos), sharded like --examples
  ap.add_argument('--gvcf_gq_binsize', type=int)
  ap.add_argument('--p_error', type=float)
  ap.add_argument('--include_med_dp', action='store_true')
  ap.add_argument('--haploid_contigs')
  ap.add_argument('--par_regions_bed')           # pseudo-autosomal regions of the haploid contigs: diploid reference confidence there
  ap.add_argument('--realign_reads', dest='realign_reads', action='store_true')
  ap.add_argument('--norealign_reads', dest='realign_reads', action='store_false')
  ap.add_argument('--sort_by_haplotypes', action='store_true')
  ap.add_argument('--trim_reads_for_pileup', action='store_true')
  ap.add_argument('--parse_sam_aux_fields', action='store_true')
  ap.add_argument('--alt_aligned_pileup')
  ap.add_argument('--stream_examples', action='store_true')   # examples go to the shard's shared-memory buffer (the reference's fast_pipeline boundary)
  ap.add_argument('--shm_prefix')
  ap.add_argument('--min_shared_contigs_basepairs', type=float)   # make_examples_options.py: 0.9
  ap.add_argument('--population_vcfs')            # allele_frequency channel: one VCF for all contigs, or one per contig (space / comma separated)
  ap.add_argument('--mean_coverage_per_sample')   # mean_coverage channel (make_examples_options.py:573-580); the first value is this sample's
  ap.add_argument('--device', type=int)
  given = vars(ap.parse_args(argv))
  merged = apply_flags_for_calling(given, given.get('checkpoint', ''), given.get('checkpoint_json', ''))
  if merged['_ignored_flags_for_calling']:
    print('make_examples: flags_for_calling of upstream stages not applied: ' + ', '.join(merged['_ignored_flags_for_calling']), file=sys.stderr)
  a = argparse.Namespace(**merged)
  from deepvariant_b200 import bam, candidates as cand, fasta, make_examples_native as men, pileup_image as pi, protos, tfrecord
  pic = pi.default_options(pi.ReadRequirements(a.min_base_quality, a.min_mapping_quality))
  pic.channels = _channels(a.channel_list)
  if a.alt_aligned_pileup == 'diff_channels':
    pic.channels += ['diff_channels_alternate_allele_1', 'diff_channels_alternate_allele_2']
  pic.num_channels = len(pic.channels)
  pic.width, pic.height = a.pileup_image_width, a.pileup_image_height
  pic.sort_by_haplotypes = a.sort_by_haplotypes
  pic.alt_aligned_pileup = a.alt_aligned_pileup
  if a.mean_coverage_per_sample:
    pic.mean_coverage = float(str(a.mean_coverage_per_sample).replace(',', ' ').split()[0])
  # Channels whose values come from DeepVariantCall maps or per-base aux data (deepvariant_b200/channels.py) are planned from Read
  # objects, where that data lives; the table packers carry the classic per-read fields only.
  plane_channels = [c for c in pic.channels if pi.CHANNEL_ENUM.get(c) in _lib_plane_enums()]
  population = None
  if 'allele_frequency' in pic.channels:
    if not a.population_vcfs:
      raise ValueError('the allele_frequency channel needs --population_vcfs')
    from deepvariant_b200 import allele_frequency as af
    population = af.make_population_vcf_readers(str(a.population_vcfs).replace(',', ' ').split())
  opts = men.MakeExamplesOptions(pic_options=pic, reference_filename=a.ref, trim_reads_for_pileup=a.trim_reads_for_pileup)
  fused = bool(a.call_variants_outfile)
  if a.stream_examples and (fused or not a.shm_prefix or not a.examples):
    raise ValueError('--stream_examples needs --shm_prefix and --examples (its @N gives the shard count), and excludes --call_variants_outfile')
  if not fused and not a.examples:
    raise ValueError('make_examples needs --examples (staged flow) or --call_variants_outfile (fused flow)')
  shard_spec = a.call_variants_outfile if fused else a.examples
  n_shards = len(tfrecord.shard_paths(shard_spec))
  gen = men.ExamplesGenerator(opts, {'main_sample': tfrecord.shard_path(a.examples, a.task)} if a.examples and not fused and not a.stream_examples else {},
                              device=a.device)
  if a.stream_examples:
    from deepvariant_b200 import stream_examples
    gen.stream = stream_examples.StreamProducer(a.shm_prefix, a.task)
    gen.writers = {'main_sample': _NullWriter()}
  try:
    import torch
    if torch.cuda.is_available():
      gen.ssw_device = a.device      # read-to-haplotype Smith-Waterman of alt-aligned pileups: one batched launch per candidate haplotype
  except ImportError:
    pass
  fused_cnn = None
  if fused:
    # encoder -> classifier in one call per batch of regions; the CallVariantsOutput shard of this task is what call_variants
    # would have written for it (same records, same rounding; postprocess_variants sorts across shards either way)
    from deepvariant_b200 import call_variants as cv, fused as fz
    if not a.checkpoint:
      raise ValueError('--call_variants_outfile needs --checkpoint (SavedModel directory / checkpoint prefix / .npz / random[:seed])')
    shape = gen.image_shape()
    cv.check_example_info({'shape': shape, 'channels': men.example_info_channels(pic)}, cv.model_example_info(a.checkpoint))
    fused_cnn = cv.GpuCnn(cv.load_weights(a.checkpoint, shape[2]), shape, device=a.device, max_batch=min(a.call_batch_size, 2048), precision=a.precision)
    gen.sink = fz.FusedCaller(gen._gpu(), fused_cnn, tfrecord.shard_path(a.call_variants_outfile, a.task), batch_images=a.call_batch_size)  # pylint: disable=protected-access
  # Native block-parallel BAM decode into a Structure-of-Arrays read table (csrc/dvb_bam.cu); untrimmed pileups (WGS/WES)
  # are planned and packed straight from the table rows, trimmed ones (PACBIO, alt-aligned) from Read objects of table.query().
  regions = parse_regions(a.regions) if a.regions else None
  # Only the reads this task can touch are decoded (dvb_bam_open_regions; with a .bai beside the file and one contig, the decoder
  # seeks): --regions widened by what the stages reach beyond a partition - the phasing padding (20 % of a partition), the realigner's
  # windows and the pileup's read-overlap buffer.
  read_regions = None
  if regions and not a.candidates_in:
    margin = (a.partition_size // 5 if a.phase_reads else 0) + 1000
    read_regions = [(c, max(0, s - margin), e + margin) for c, s, e in regions]
  # MM / ML / MN (and Ultima's tp / t0) are parsed when a channel that needs them is requested, or with --parse_sam_aux_fields
  # (make_examples_options.py:445-452)
  if any(pi.CHANNEL_ENUM.get(c) in (23, 24, 28, 29, 30) for c in pic.channels):
    a.parse_sam_aux_fields = True
  # --reads may be a CRAM (decoded against --ref: csrc/dvb_cram.cu), as sam_reader.cc takes one with the FASTA
  cram_ref = fasta.IndexedFastaReader(a.ref) if bam.is_cram(a.reads) else None

  def open_reads(read_regions_):
    return bam.NativeBamTable(a.reads, bam.ReadRequirements(min_mapping_quality=a.min_mapping_quality), parse_aux=a.parse_sam_aux_fields, regions=read_regions_,
                              ref_reader=cram_ref)
  # Without --regions a task walks its partitions over the whole genome (the reference queries its indexed reader once per partition,
  # make_examples_core.py:2262-2288).  Holding every read of a 30x genome is not an option, so the table then covers one
  # window of the genome at a time: DVB_READ_WINDOW_BP bases (default 10 Mb, rounded to whole partitions), widened by the same margin
  # as --regions, re-opened through the .bai / the CRAM container headers when the walk leaves it.  Memory is one window's reads; each
  # task decodes each window it has a partition in once.
  window_bp = int(os.environ.get('DVB_READ_WINDOW_BP', '0'))
  indexed = cram_ref is not None or any(os.path.exists(p) for p in (a.reads + '.bai', os.path.splitext(a.reads)[0] + '.bai'))
  windowed = not regions and not a.candidates_in and (window_bp > 0 or indexed)
  if windowed:
    window_bp = max(a.partition_size, (window_bp or 10_000_000) // a.partition_size * a.partition_size)
  loaded = [None]            # (contig, first base, last base + 1) of the window `reader` holds
  reader = open_reads([(fasta.IndexedFastaReader(a.ref).contig_order[0], 0, 1)] if windowed else read_regions)

  def ensure_reads(contig, p0, p1):
    nonlocal reader
    if not windowed or (loaded[0] is not None and loaded[0][0] == contig and loaded[0][1] <= p0 and p1 <= loaded[0][2]):
      return
    lo = p0 // window_bp * window_bp
    hi = max(lo + window_bp, p1)          # a partition cut by candidate count (--candidate_positions) may be longer than a window
    margin = (max(a.partition_size, p1 - p0) // 5 if a.phase_reads else 0) + 1000
    reader.close()
    reader = open_reads([(contig, max(0, lo - margin), hi + margin)])
    loaded[0] = (contig, lo, hi)
  table_path = not a.trim_reads_for_pileup and a.alt_aligned_pileup == 'none' and not plane_channels

  def annotated(calls, contig):
    # make_examples_core.py:2380-2388: population allele frequencies of the candidates' alleles
    if population is None:
      return calls
    return af.add_allele_frequencies_to_candidates(calls, population.get(contig, population.get('*')), ref_for_af)
  if regions is not None and len(regions) > 1 and (a.candidates_in or a.mode == 'candidate_sweep'):
    raise NotImplementedError('several --regions together with --candidates_in / --mode candidate_sweep')
  region = regions[0] if regions else None
  totals = {}

  ref_for_af = fasta.IndexedFastaReader(a.ref) if population is not None else None

  def examples_in(cs, contig, p0, p1):
    cs = annotated(cs, contig)
    if table_path:
      stats, _ = gen.write_examples_in_region_from_table(cs, reader, 'main_sample', (contig, p0, p1))
    else:
      stats, _ = gen.write_examples_in_region(cs, [reader.query(contig, p0, p1)], [0], 'main_sample', [0.0])
    for key, val in stats.items():
      totals[key] = totals.get(key, 0) + val

  if a.candidates_in and a.gvcf:
    raise ValueError('--gvcf needs the allele counter: not available with --candidates_in')
  if a.candidates_in:
    cands = [protos.parse_deepvariant_call(r) for p in tfrecord.resolve_input_paths(a.candidates_in) for r in tfrecord.read_records(p)]
    for (contig, k, origin), cs in men.shard_partitions(men.partition_candidates(cands, a.partition_size, region), n_shards, a.task):
      p0 = origin + k * a.partition_size
      examples_in(cs, contig, p0, p0 + a.partition_size if not region else min(p0 + a.partition_size, region[2]))
  else:
    # Candidate generation (csrc/dvb_candidates.cu): allele counter + very-sensitive caller over each region of this task,
    # exactly the regions `--task i` of N gets in the reference (regions_to_process, make_examples_core.py:799-888).
    rl = None
    if a.realign_reads and a.phase_reads:
      raise NotImplementedError('--realign_reads together with --phase_reads (the PACBIO / ONT models run with --norealign_reads)')
    if a.normalize_reads and a.phase_reads:
      raise NotImplementedError('--normalize_reads together with --phase_reads')
    if a.realign_reads:
      from deepvariant_b200 import realigner
      ropts = realigner.RealignerOptions(normalize_reads=a.normalize_reads)                 # realigner.py:414-429
      ropts.ws.keep_legacy_behavior = a.keep_legacy_allele_counter_behavior                 # realigner.py:345-360
      rl = realigner.Realigner(fasta.IndexedFastaReader(a.ref), ropts)
      rl.ssw_device = gen.ssw_device     # read-to-haplotype / haplotype-to-reference Smith-Waterman in batched launches when the stage runs on a GPU
    ref = fasta.IndexedFastaReader(a.ref)
    copts = cand.CandidateOptions(
        min_mapping_quality=a.min_mapping_quality, min_base_quality=a.min_base_quality,
        keep_legacy_allele_counter_behavior=a.keep_legacy_allele_counter_behavior, track_ref_reads=a.track_ref_reads,
        vsc_min_count_snps=a.vsc_min_count_snps, vsc_min_count_indels=a.vsc_min_count_indels, vsc_min_fraction_snps=a.vsc_min_fraction_snps,
        vsc_min_fraction_indels=a.vsc_min_fraction_indels, vsc_min_fraction_multiplier=a.vsc_min_fraction_multiplier,
        small_model_vaf_context_window_size=a.small_model_vaf_context_window_size,
        sample_name=a.sample_name or cand.sample_name_from_bam(a.reads), max_reads_per_partition=a.max_reads_per_partition,
        partition_size=a.partition_size)
    if table_path and not a.normalize_reads and not a.phase_reads and a.variant_caller == 'very_sensitive_caller' and \
        os.environ.get('DVB_DEVICE_SUPPORT', '1') != '0' and \
        isinstance(gen._gpu(), pi.GpuEncoder):   # pylint: disable=protected-access  (a stand-in encoder, as in the CPU tests, cannot derive: names then)
      # the candidates come from the allele counter over the very reads the pileups show: the encoder's pre-pass derives the
      # (candidate, read) support classes on the device from the alt alleles instead of a read-name search on the host
      # (--normalize_reads counts from rewritten alignments that the pileups do not show: names stay authoritative there)
      gen.support_options = (a.min_mapping_quality, a.min_base_quality, bool(a.keep_legacy_allele_counter_behavior), bool(a.track_ref_reads))
    cand_writer = tfrecord.Writer(tfrecord.shard_path(a.candidates, a.task) if tfrecord.is_sharded_spec(a.candidates) else a.candidates) \
        if a.candidates else None
    contigs = shared_contigs([(c, ref.n_bases(c)) for c in ref.contig_order], dict(zip(reader.references, reader.reference_lengths)),
                             a.min_shared_contigs_basepairs)
    gvcf_writer = gvcf_options = gvcf_confidence = None
    if a.gvcf:
      # --gvcf: reference-confidence blocks of every region from the same allele counter (deepvariant_b200/gvcf.py;
      # variant_caller.py:256-468), one TFRecord of Variant protos per task
      from deepvariant_b200 import gvcf
      if a.gvcf_gq_binsize < 1:
        raise ValueError('--gvcf_gq_binsize must be a positive integer')
      gvcf_writer = tfrecord.Writer(tfrecord.shard_path(a.gvcf, a.task) if tfrecord.is_sharded_spec(a.gvcf) else a.gvcf)
      gvcf_options = gvcf.GvcfOptions(sample_name=copts.sample_name, p_error=a.p_error, gq_resolution=a.gvcf_gq_binsize, include_med_dp=a.include_med_dp,
                                      haploid_contigs=tuple(c for c in a.haploid_contigs.split(',') if c),
                                      par_regions=tuple(__import__('deepvariant_b200.postprocess_variants', fromlist=['read_bed']).read_bed(a.par_regions_bed))
                                      if a.par_regions_bed else ())
      gvcf_confidence = gvcf.ReferenceConfidence(gvcf_options)

    def write_gvcfs(found, contig, p0, p1):
      if gvcf_writer is None:
        return
      p1 = min(p1, ref.n_bases(contig))
      lo = p0 - found.interval[0]                 # summary_counts(left_padding, right_padding): without the phasing padding
      for block in gvcf.make_gvcfs(contig, p0, ref.query(contig, p0, p1), found.summary_counts[lo:lo + (p1 - p0)], gvcf_options, gvcf_confidence):
        gvcf_writer.write(gvcf.serialize_gvcf_record(block))
        totals['n_gvcf_records'] = totals.get('n_gvcf_records', 0) + 1

    if a.mode == 'candidate_sweep':
      # --mode candidate_sweep (make_examples_core.py:2117-2189, 3592-3605): the candidate positions of every partition of this task
      # (no realigner, one counter pass), each partition closed by END_OF_PARTITION and a calling region by END_OF_REGION; a later
      # calling run cuts its partitions by candidate count from the merged shards (--candidate_positions).
      if not a.candidate_positions:
        raise ValueError('--mode candidate_sweep needs --candidate_positions')
      import numpy as np
      path = tfrecord.shard_path(a.candidate_positions, a.task) if tfrecord.is_sharded_spec(a.candidate_positions) else a.candidate_positions
      region_ends = {(c, min(n, region[2]) if region and region[0] == c else n) for c, n in contigs}
      with open(path, 'wb') as f:
        for contig, p0, p1 in cand.regions_to_process(contigs, a.partition_size, region, a.task, n_shards):
          ensure_reads(contig, p0, p1)
          rows = cand.region_reads(reader, contig, p0, p1, copts.max_reads_per_partition, copts.random_seed)
          positions = cand.candidate_positions(reader, ref, contig, p0, p1, rows, copts) + [cand.END_OF_PARTITION]
          if (contig, p1) in region_ends:
            positions.append(cand.END_OF_REGION)
          f.write(np.asarray(positions, dtype=np.int32).tobytes())
          totals['n_candidate_positions'] = totals.get('n_candidate_positions', 0) + sum(1 for x in positions if x >= 0)
      gen.signal_shard_finished()
      print(f'make_examples task {a.task}: {totals}', file=sys.stderr)
      return 0
    proposed = None
    if a.variant_caller == 'vcf_candidate_importer':
      if not a.proposed_variants:
        raise SystemExit('--proposed_variants is required with --variant_caller vcf_candidate_importer')   # make_examples_options.py:1478-1486
      from deepvariant_b200 import vcf_candidate_importer
      proposed = vcf_candidate_importer.ProposedVcfReader(a.proposed_variants)
    elif a.proposed_variants:
      raise SystemExit('--proposed_variants needs --variant_caller vcf_candidate_importer')

    def find_candidates(table_, contig, p0, p1, rows_, padding_pct=0):
      if proposed is not None:
        return vcf_candidate_importer.calls_from_vcf(table_, ref, contig, p0, p1, copts, proposed, rows=rows_, padding_pct=padding_pct)
      return cand.candidates_in_region(table_, ref, contig, p0, p1, copts, rows=rows_, padding_pct=padding_pct)

    sweep = cand.load_candidate_positions(a.candidate_positions) if a.candidate_positions else None
    # --runtime_by_region: one TSV line per region with the seconds of each stage and its counts (make_examples_core.py:95-108,
    # 1348-1353, 3678-3707; docs/runtime-by-region.md).  Pileup encoding and the example writes are one fused step here
    # (pack -> CUDA encode -> tf.Example), so 'make pileup images' holds both and 'write outputs' the candidate / gVCF records.
    import time
    runtime_writer = None
    if a.runtime_by_region:
      runtime_path = tfrecord.shard_path(a.runtime_by_region, a.task) if tfrecord.is_sharded_spec(a.runtime_by_region) else a.runtime_by_region
      runtime_writer = open(runtime_path, 'w')
      runtime_writer.write('\t'.join(RUNTIME_BY_REGION_COLUMNS) + '\n')

    def mark(rt, stage):
      now = time.time()
      rt[stage] = round(rt.get(stage, 0) + now - rt['_t'], 3)        # trim_runtime: milliseconds
      rt['_t'] = now

    def region_body(contig, p0, p1, rt):
      ensure_reads(contig, p0, p1)
      if proposed is not None and gvcf_writer is None and not vcf_candidate_importer.region_has_proposed_variant(proposed, contig, p0, p1):
        return       # filter_regions_by_vcf (make_examples_core.py:3443-3478): nothing proposed here and no gVCF blocks to write
      rows = cand.region_reads(reader, contig, p0, p1, copts.max_reads_per_partition, copts.random_seed)
      rt['num reads'] = len(rows)
      if not len(rows) and proposed is None:      # (the importer still proposes its records, with no evidence: the reference's own early exit,
        if gvcf_writer is not None:               #  make_examples_core.py:2871-2875, tests a generator and never fires)               # no early exit with --gvcf: the region still gets its blocks (make_examples_core.py:2872-2875)
          write_gvcfs(find_candidates(reader, contig, p0, p1, rows, 20 if a.phase_reads else 0), contig, p0, p1)
        return
      if (rl is not None or a.normalize_reads) and len(rows):
        # --realign_reads: window selection, de Bruijn assembly, FastPassAligner (deepvariant_b200/realigner.py); the realigned reads
        # replace the region's reads for candidate generation AND pileups, as in_memory_sam_reader.replace_reads does
        # (make_examples_core.py:2290-2300).  --normalize_reads then left-normalises the indels of the region's reads
        # (deepvariant_b200/normalize_reads.py; make_examples_core.py:2900-2953).  The rewritten reads become a table derived natively from
        # the source table's rows (bam.scratch_table -> dvb_bam_derive) so that the candidate generator / packer can take them.
        region_read_list = rl.realign_reads(reader, contig, rows, (p0, p1)) if rl is not None else [reader.read(int(i)) for i in rows]
        count_reads = None
        if a.normalize_reads:
          from deepvariant_b200 import normalize_reads
          region_read_list, count_reads = normalize_reads.normalize_region_reads(
              region_read_list, lambda s, e, contig=contig: ref.query(contig, s, e).encode(), p0, p1, ref.n_bases(contig), a.min_mapping_quality)
        reqs = bam.ReadRequirements(min_mapping_quality=a.min_mapping_quality)
        refs = [(c, ref.n_bases(c)) for c in ref.contig_order]
        region_table = bam.scratch_table(region_read_list, refs, reqs, parse_aux=a.parse_sam_aux_fields)
        region_rows = region_table.query_indices(contig, p0, p1)
        mark(rt, 'get reads')
        if count_reads is None:
          found = find_candidates(region_table, contig, p0, p1, region_rows, 20 if a.phase_reads else 0)
        else:
          # a read whose only change is its heading indel is counted with the rewritten alignment but keeps its own in memory
          # (NormalizeAndAdd, allelecounter.cc:865-870): count from a second table
          count_table = bam.scratch_table(count_reads, refs, reqs)
          found = find_candidates(count_table, contig, p0, p1, count_table.query_indices(contig, p0, p1))
          count_table.close()
        mark(rt, 'find candidates')
        write_gvcfs(found, contig, p0, p1)
        totals['n_candidates'] = totals.get('n_candidates', 0) + len(found.records)
        rt['num candidates'] = len(found.records)
        if cand_writer is not None:
          for rec in found.records:
            cand_writer.write(rec)
        mark(rt, 'write outputs')
        if found.records:
          if table_path and not a.phase_reads:
            stats, _ = gen.write_examples_in_region_from_table(found.calls(), region_table, 'main_sample', (contig, p0, p1))
          else:
            stats, _ = gen.write_examples_in_region(annotated(found.calls(), contig), [[region_table.read(int(i)) for i in region_rows]], [0], 'main_sample', [0.0])
          for key, val in stats.items():
            totals[key] = totals.get(key, 0) + val
          rt['num examples'] = stats.get('n_examples', 0)
        mark(rt, 'make pileup images')
        region_table.close()
        return
      mark(rt, 'get reads')
      found = find_candidates(reader, contig, p0, p1, rows, 20 if a.phase_reads else 0)
      mark(rt, 'find candidates')
      write_gvcfs(found, contig, p0, p1)
      rt['num candidates'] = len(found.records)
      totals['n_candidates'] = totals.get('n_candidates', 0) + len(found.records)
      if cand_writer is not None:
        for rec in found.records:
          cand_writer.write(rec)
      mark(rt, 'write outputs')
      before = totals.get('n_examples', 0)
      if found.records:
        if a.phase_reads:
          # direct phasing over the padded region's candidates (deepvariant/direct_phasing.cc); HP of every region read is
          # replaced, as make_examples_core.py:2998-3111 does; the pileups then come from these Read objects.
          from deepvariant_b200 import direct_phasing
          reads = [reader.read(int(r)) for r in rows]
          phases = direct_phasing.phase_reads([cand.canonical_call(r) for r in found.all_records], [r.key() for r in reads])
          for r, ph in zip(reads, phases):
            r.hp_values = [ph]
          stats, _ = gen.write_examples_in_region(annotated(found.calls(), contig), [reads], [0], 'main_sample', [0.0])
          for key, val in stats.items():
            totals[key] = totals.get(key, 0) + val
        else:
          examples_in(found.calls(), contig, p0, p1)
      rt['num examples'] = totals.get('n_examples', 0) - before
      mark(rt, 'make pileup images')

    for contig, p0, p1 in cand.regions_to_process(contigs, a.partition_size, regions, a.task, n_shards, candidates=sweep):
      rt = {'_t': time.time(), 'region': f'{contig}:{p0 + 1}-{p1}', 'num reads': 0, 'num candidates': 0, 'num examples': 0}
      region_body(contig, p0, p1, rt)
      if runtime_writer is not None:
        runtime_writer.write('\t'.join(str(rt.get(k, 'NA')) for k in RUNTIME_BY_REGION_COLUMNS) + '\n')
    if runtime_writer is not None:
      runtime_writer.close()
    if cand_writer is not None:
      cand_writer.close()
    if gvcf_writer is not None:
      gvcf_writer.close()
  gen.signal_shard_finished()
  if fused:
    totals['n_call_variants_outputs'] = gen.sink.close()
    totals['n_classifier_batches'] = gen.sink.n_batches
    fused_cnn.close()
  print(f'make_examples task {a.task}: {totals}', file=sys.stderr)
  return 0


def call_variants(argv):
  ap = argparse.ArgumentParser('call_variants')
  ap.add_argument('--examples', default='')
  ap.add_argument('--stream_examples', action='store_true')   # examples from the make_examples processes' shared-memory buffers
  ap.add_argument('--shm_prefix', default='')
  ap.add_argument('--num_input_shards', type=int, default=0)
  ap.add_argument('--outfile', required=True)
  ap.add_argument('--checkpoint', required=True)
  ap.add_argument('--batch_size', type=int, default=1024)
  ap.add_argument('--writer_threads', type=int, default=0)
  ap.add_argument('--device', type=int, default=0)
  ap.add_argument('--precision', type=int, default=1, choices=[0, 1])   # 1: split-fp16 x3, 1e-5 of fp32 (default); 0: fp16 operands
  a = ap.parse_args(argv)
  from deepvariant_b200 import call_variants as cv
  if a.stream_examples:
    if not a.shm_prefix or a.num_input_shards < 1:
      raise ValueError('--stream_examples needs --shm_prefix and --num_input_shards')
    r = cv.call_variants_from_stream(a.shm_prefix, a.num_input_shards, a.checkpoint, a.outfile, writer_threads=a.writer_threads, device=a.device,
                                     precision=a.precision)
  else:
    if not a.examples:
      raise ValueError('call_variants needs --examples (or --stream_examples)')
    r = cv.call_variants(a.examples, a.checkpoint, a.outfile, a.batch_size, a.writer_threads, a.device, precision=a.precision)
  print(f'call_variants: {r["n_examples"]} examples in {r["n_batches"]} batches -> {len(r["paths"])} shard(s)', file=sys.stderr)
  return 0


def postprocess_variants(argv):
  """postprocess_variants --ref R.fa --infile call_variants_output.tfrecord.gz --outfile out.vcf[.gz] (deepvariant/postprocess_variants.py
  flags :60-330; the single-sample VCF path - no gVCF)."""
  ap = argparse.ArgumentParser('postprocess_variants')
  ap.add_argument('--ref', required=True)
  ap.add_argument('--infile', required=True)
  ap.add_argument('--outfile', required=True)
  ap.add_argument('--sample_name', default='')
  ap.add_argument('--qual_filter', type=float, default=1.0)
  ap.add_argument('--multi_allelic_qual_filter', type=float, default=1.0)
  ap.add_argument('--multiallelic_mode', default='product', choices=['min', 'product'])
  ap.add_argument('--only_keep_pass', action='store_true')
  ap.add_argument('--disable_haplotype_resolution', action='store_true')
  ap.add_argument('--group_variants', dest='group_variants', action='store_true', default=True)
  ap.add_argument('--nogroup_variants', dest='group_variants', action='store_false')
  ap.add_argument('--nonvariant_site_tfrecord_path', default='')     # make_examples --gvcf output, every shard
  ap.add_argument('--gvcf_outfile', default='')
  ap.add_argument('--haploid_contigs', default='')                   # e.g. chrX,chrY: heterozygous genotypes are ruled out there ...
  ap.add_argument('--par_regions_bed', default='')                   # ... except inside these pseudo-autosomal regions
  ap.add_argument('--cpus', type=int, default=min(os.cpu_count() or 1, 16))   # worker processes of the CVO -> Variant conversion (the reference's --cpus)
  a = ap.parse_args(argv)
  from deepvariant_b200 import fasta, postprocess_variants as pp
  ref = fasta.IndexedFastaReader(a.ref)
  r = pp.postprocess_variants(a.infile, a.outfile, [(c, ref.n_bases(c)) for c in ref.contig_order], a.sample_name, a.qual_filter,
                              a.multi_allelic_qual_filter, a.multiallelic_mode, a.only_keep_pass, a.disable_haplotype_resolution, a.group_variants,
                              a.nonvariant_site_tfrecord_path, a.gvcf_outfile, lambda c, p: ref.query(c, p, p + 1), a.haploid_contigs, a.par_regions_bed,
                              cpus=a.cpus)
  print(f'postprocess_variants: {r}', file=sys.stderr)
  return 0


def extra_args_to_dict(extra_args: str) -> dict:
  """scripts/run_deepvariant.py:343-357: "flag=value,flag2=true" -> {flag: value}; true / false become booleans.  A comma inside a
  value is kept when the piece after it has no '=' (split_extra_args, :330-340: `regions=chr1 chr2,other=x`)."""
  out, pieces = {}, []
  for piece in (extra_args or '').split(','):
    if '=' in piece or not pieces:
      pieces.append(piece)
    else:
      pieces[-1] += ',' + piece
  for piece in pieces:
    if not piece.strip():
      continue
    if '=' not in piece:
      raise ValueError(f'extra args: "{piece}" is not flag=value')
    name, value = piece.split('=', 1)
    name = name.strip().strip('-')
    out[name] = True if value.lower() == 'true' else False if value.lower() == 'false' else value
  return out


def _value_flags(args: List[str]) -> set:
  return {x[2:] for i, x in enumerate(args) if x.startswith('--') and i + 1 < len(args) and not args[i + 1].startswith('--')}


def apply_extra_args(args: List[str], extra: dict, value_flags) -> List[str]:
  """_update_kwargs_with_warning + _extend_command_by_args_dict (scripts/run_deepvariant.py:360-386) on an argv list: an extra flag
  replaces the one the wrapper set (with the reference's warning); booleans become --flag / --noflag."""
  args = list(args)
  for key in sorted(extra):
    value = extra[key]
    for name in ('--' + key, '--no' + key):
      while name in args:
        i = args.index(name)
        takes_value = name == '--' + key and key in value_flags
        old = args[i + 1] if takes_value else name == '--' + key
        if old != value:
          print(f'\nWarning: --{key} is previously set to {old}, now to {value}.', file=sys.stderr)
        del args[i:i + (2 if takes_value else 1)]
    if isinstance(value, bool):
      if value:
        args.append('--' + key)
      elif key in ('realign_reads', 'group_variants'):     # the two switches that default to true have a --no form; false is the default of the rest
        args.append('--no' + key)
    else:
      args += ['--' + key, str(value)]
  return args


def run_deepvariant(argv):
  ap = argparse.ArgumentParser('run_deepvariant')
  ap.add_argument('--model_type', required=True, choices=sorted(MODEL_DEFAULTS))
  ap.add_argument('--ref', required=True)
  ap.add_argument('--reads', required=True)
  ap.add_argument('--candidates_in', default='')   # optional: DeepVariantCalls exported by another make_examples
  ap.add_argument('--output_dir', required=True)   # the reference's --intermediate_results_dir
  ap.add_argument('--output_vcf', default='')       # scripts/run_deepvariant.py:84: when given, postprocess_variants runs too
  ap.add_argument('--output_gvcf', default='')      # scripts/run_deepvariant.py:90: make_examples --gvcf + the postprocess merge
  ap.add_argument('--sample_name', default='')
  ap.add_argument('--regions', default='')
  ap.add_argument('--num_shards', type=int, default=1)
  ap.add_argument('--customized_model', required=True)   # SavedModel dir / checkpoint prefix / .npz; the exact token random[:seed] = noise weights (loud warning)
  ap.add_argument('--precision', type=int, default=1, choices=[0, 1])
  ap.add_argument('--staged', action='store_true')         # the reference's three stages with tf.Example files in between; default = fused
  ap.add_argument('--num_gpus', type=int, default=0)        # 0 = every visible device; task i runs on device i mod num_gpus
  ap.add_argument('--jobs', type=int, default=0)            # tasks in flight; 0 = all of them, as `parallel -j num_shards` (scripts/run_deepvariant.py:457-462)
  ap.add_argument('--call_variants_extra_args', default='')  # scripts/run_deepvariant.py:112-118: "flag=value,..." - batch_size is honoured (classifier chunk, device memory)
  ap.add_argument('--make_examples_extra_args', default='')          # scripts/run_deepvariant.py:105-111, e.g. "variant_caller=vcf_candidate_importer,proposed_variants=X.vcf.gz"
  ap.add_argument('--postprocess_variants_extra_args', default='')   # :119-125
  ap.add_argument('--logging_dir', default='')              # scripts/run_deepvariant.py:141
  ap.add_argument('--runtime_report', action='store_true')  # scripts/run_deepvariant.py:149,744-758: make_examples --runtime_by_region into logging_dir
  a = ap.parse_args(argv)
  os.makedirs(a.output_dir, exist_ok=True)
  cv_extra = dict(kv.split('=', 1) for kv in a.call_variants_extra_args.split(',') if '=' in kv)
  unknown = sorted(set(cv_extra) - {'batch_size'})
  if unknown:
    raise ValueError(f'--call_variants_extra_args: unsupported flags {unknown} (batch_size is)')
  me_extra = extra_args_to_dict(a.make_examples_extra_args)
  pp_extra = extra_args_to_dict(a.postprocess_variants_extra_args)
  runtime_by_region = ''
  if a.logging_dir and a.runtime_report:
    os.makedirs(os.path.join(a.logging_dir, 'make_examples_runtime_by_region'), exist_ok=True)
    runtime_by_region = os.path.join(a.logging_dir, 'make_examples_runtime_by_region', f'make_examples_runtime@{a.num_shards}.tsv')
  d = MODEL_DEFAULTS[a.model_type]
  examples = os.path.join(a.output_dir, f'make_examples.tfrecord@{a.num_shards}.gz')
  nonvariants = os.path.join(a.output_dir, f'gvcf.tfrecord@{a.num_shards}.gz')
  cvo = os.path.join(a.output_dir, 'call_variants_output.tfrecord.gz')
  if a.output_gvcf and not a.output_vcf:
    raise ValueError('--output_gvcf needs --output_vcf')
  import glob
  for stale in glob.glob(os.path.join(a.output_dir, 'call_variants_output-?????-of-?????.tfrecord.gz')):
    os.remove(stale)          # shards of an earlier run with another shard count would be read by postprocess_variants as well
  n_gpus = a.num_gpus
  if n_gpus <= 0:
    try:
      import torch
      n_gpus = max(1, torch.cuda.device_count())
    except ImportError:
      n_gpus = 1
  task_args = []
  for task in range(a.num_shards):
    args = ['--mode', 'calling', '--ref', a.ref, '--reads', a.reads, '--task',
            str(task), '--channel_list', d['channel_list'], '--pileup_image_width', str(d['pileup_image_width']), '--device', str(task % n_gpus)]
    if a.staged:
      args += ['--examples', examples]
    else:
      args += ['--call_variants_outfile', os.path.join(a.output_dir, f'call_variants_output@{a.num_shards}.tfrecord.gz'),
               '--checkpoint', a.customized_model, '--precision', str(a.precision)]
      if 'batch_size' in cv_extra:
        args += ['--call_batch_size', str(int(cv_extra['batch_size']))]
    if a.candidates_in:
      args += ['--candidates_in', a.candidates_in]
    for flag in ('sort_by_haplotypes', 'trim_reads_for_pileup', 'parse_sam_aux_fields', 'track_ref_reads', 'phase_reads', 'norealign_reads'):
      if d.get(flag):
        args.append('--' + flag)
    for flag in ('alt_aligned_pileup', 'min_mapping_quality', 'partition_size', 'max_reads_per_partition', 'vsc_min_fraction_indels'):
      if flag in d:
        args += ['--' + flag, str(d[flag])]
    if a.regions:
      args += ['--regions', a.regions]
    if a.sample_name:
      args += ['--sample_name', a.sample_name]
    if a.output_gvcf:
      args += ['--gvcf', nonvariants]
    if runtime_by_region:
      args += ['--runtime_by_region', runtime_by_region]
    if me_extra:
      args = apply_extra_args(args, me_extra, _value_flags(args))
    task_args.append(args)
  # One process per task, task i on GPU i mod num_gpus, started together as the reference starts its make_examples shards
  # (scripts/run_deepvariant.py:457-462, 497); a single task runs in this process.
  if len(task_args) == 1:
    make_examples(task_args[0])
  else:
    import subprocess
    jobs = a.jobs if a.jobs > 0 else len(task_args)
    running, failed, queue = [], [], list(enumerate(task_args))
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([os.path.dirname(os.path.dirname(os.path.abspath(__file__))), os.environ.get('PYTHONPATH', '')]))
    while queue or running:
      while queue and len(running) < jobs:
        task, args = queue.pop(0)
        running.append((task, subprocess.Popen([sys.executable, '-m', 'deepvariant_b200.cli', 'make_examples'] + args, env=env)))
      task, proc = running.pop(0)
      if proc.wait() != 0:
        failed.append(task)
    if failed:
      raise RuntimeError(f'make_examples tasks {failed} failed')
  rc = 0
  if a.staged:
    rc = call_variants(['--examples', examples, '--outfile', cvo, '--checkpoint', a.customized_model, '--precision', str(a.precision)] +
                       (['--batch_size', str(int(cv_extra['batch_size']))] if 'batch_size' in cv_extra else []))
  if rc or not a.output_vcf:
    return rc
  args = ['--ref', a.ref, '--infile', cvo, '--outfile', a.output_vcf]     # the shards call_variants wrote are found by name
  if a.sample_name:
    args += ['--sample_name', a.sample_name]
  if a.output_gvcf:
    args += ['--nonvariant_site_tfrecord_path', nonvariants, '--gvcf_outfile', a.output_gvcf]
  if pp_extra:
    args = apply_extra_args(args, pp_extra, _value_flags(args))
  return postprocess_variants(args)


def main():
  stages = {'make_examples': make_examples, 'call_variants': call_variants, 'postprocess_variants': postprocess_variants,
            'run_deepvariant': run_deepvariant}
  if len(sys.argv) < 2 or sys.argv[1] not in stages:
    print(__doc__)
    return 2
  return stages[sys.argv[1]](sys.argv[2:])


if __name__ == '__main__':
  sys.exit(main())
