"""ctypes binding of the C ABI in include/dvb.h (libdvb.so, built in-tree by
__graft_entry__.build()).

The product path has NO CPU fallback: if the CUDA library is missing, `lib()` raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

DVB_MAX_CHANNELS = 16

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('DVB_LIB_PATH') or os.path.join(_HERE, 'csrc', 'libdvb.so')   # DVB_LIB_PATH: a differently compiled build of the same sources (A/B of compile-time choices)


class DvbPileupParams(C.Structure):
  _fields_ = [
      ('width', C.c_int32),
      ('height', C.c_int32),
      ('reference_band_height', C.c_int32),
      ('num_channels', C.c_int32),
      ('channels', C.c_int32 * DVB_MAX_CHANNELS),
      ('num_alt_channels', C.c_int32),
      ('base_color_offset_a_and_g', C.c_int32),
      ('base_color_offset_t_and_c', C.c_int32),
      ('base_color_stride', C.c_int32),
      ('allele_supporting_read_alpha', C.c_float),
      ('allele_unsupporting_read_alpha', C.c_float),
      ('other_allele_supporting_read_alpha', C.c_float),
      ('reference_matching_read_alpha', C.c_float),
      ('reference_mismatching_read_alpha', C.c_float),
      ('indel_anchoring_base_char', C.c_int32),
      ('reference_base_quality', C.c_int32),
      ('positive_strand_color', C.c_int32),
      ('negative_strand_color', C.c_int32),
      ('base_quality_cap', C.c_int32),
      ('mapping_quality_cap', C.c_int32),
      ('min_base_quality', C.c_int32),
      ('min_mapping_quality', C.c_int32),
      ('sort_by_haplotypes', C.c_int32),
      ('hp_tag_for_assembly_polishing', C.c_int32),
      ('sort_by_alt_allele_support', C.c_int32),
      ('random_seed', C.c_uint32),
      ('max_reads_per_image', C.c_int32),
      ('shuffle_stdlib', C.c_int32),
      ('mean_coverage', C.c_float),
      ('blank_channel_mask', C.c_uint32),
  ]


class DvbBatch(C.Structure):
  _fields_ = [
      ('n_images', C.c_int32),
      ('n_reads', C.c_int32),
      ('n_pairs', C.c_int64),
      ('n_bases', C.c_int64),
      ('n_cigar', C.c_int64),
      ('ref_bases', C.c_void_p),
      ('ref_stride', C.c_int32),
      ('image_start_pos', C.c_void_p),
      ('variant_start', C.c_void_p),
      ('pair_begin', C.c_void_p),
      ('pair_read', C.c_void_p),
      ('pair_support', C.c_void_p),
      ('pair_allele_group', C.c_void_p),
      ('read_pos', C.c_void_p),
      ('read_sort_pos', C.c_void_p),
      ('read_mapq', C.c_void_p),
      ('read_flags', C.c_void_p),
      ('read_fragment_length', C.c_void_p),
      ('read_hp', C.c_void_p),
      ('read_name_rank', C.c_void_p),
      ('read_seq_begin', C.c_void_p),
      ('read_cigar_begin', C.c_void_p),
      ('bases', C.c_void_p),
      ('quals', C.c_void_p),
      ('cigar', C.c_void_p),
      ('allele_begin', C.c_void_p),
      ('allele_type', C.c_void_p),
      ('allele_class', C.c_void_p),
      ('allele_group', C.c_void_p),
      ('allele_bases_begin', C.c_void_p),
      ('allele_bases', C.c_void_p),
      ('image_ref_run', C.c_void_p),
      ('image_group_default', C.c_void_p),
      ('n_alleles', C.c_int64),
      ('n_allele_bases', C.c_int64),
      ('support_min_mapping_quality', C.c_int32),
      ('support_min_base_quality', C.c_int32),
      ('support_flags', C.c_int32),
      ('pair_channel', C.c_void_p * 3),
      ('base_channel', C.c_void_p * 5),
  ]


# Optional members (device-side pair support): absent from PackedBatch.arrays = NULL.
ALLELE_ARRAYS = (
    ('allele_begin', 'int64'), ('allele_type', 'uint8'), ('allele_class', 'uint8'), ('allele_group', 'uint8'),
    ('allele_bases_begin', 'int64'), ('allele_bases', 'uint8'), ('image_ref_run', 'int32'), ('image_group_default', 'uint8'),
)
# Optional channel planes (include/dvb.h DVB_PAIR_PLANE_* / DVB_BASE_PLANE_*): PackedBatch.arrays['pair_channel_<k>'] uint8[n_pairs],
# ['base_channel_<k>'] uint8[n_bases]; channel enum -> (member, slot).
PLANE_OF_CHANNEL = {8: ('pair_channel', 0), 25: ('pair_channel', 1), 27: ('pair_channel', 2),
                    23: ('base_channel', 0), 24: ('base_channel', 1), 28: ('base_channel', 2), 29: ('base_channel', 3), 30: ('base_channel', 4)}
PLANE_ARRAYS = tuple(('pair_channel', k) for k in range(3)) + tuple(('base_channel', k) for k in range(5))
SUPPORT_KEEP_LEGACY, SUPPORT_TRACK_REF_READS, SUPPORT_REPEATED_KEYS = 1, 2, 4

# name -> (dtype string, per-what) for every array member of DvbBatch, in struct order.
BATCH_ARRAYS = (
    ('ref_bases', 'uint8'), ('image_start_pos', 'int32'), ('variant_start', 'int32'),
    ('pair_begin', 'int64'), ('pair_read', 'int32'), ('pair_support', 'uint8'),
    ('pair_allele_group', 'uint8'), ('read_pos', 'int32'), ('read_sort_pos', 'int32'),
    ('read_mapq', 'int32'), ('read_flags', 'uint8'), ('read_fragment_length', 'int32'),
    ('read_hp', 'int32'), ('read_name_rank', 'uint32'), ('read_seq_begin', 'int64'),
    ('read_cigar_begin', 'int64'), ('bases', 'uint8'), ('quals', 'uint8'), ('cigar', 'uint32'),
)

DVB_STATUS_NAMES = {
    0: 'DVB_OK', 1: 'DVB_ERR_INVALID_ARGUMENT', 2: 'DVB_ERR_UNSUPPORTED_CHANNEL',
    3: 'DVB_ERR_BAD_CIGAR', 4: 'DVB_ERR_TOO_MANY_READS', 5: 'DVB_ERR_CUDA',
    6: 'DVB_ERR_NO_DEVICE', 7: 'DVB_ERR_INTERNAL',
}


class DvbError(RuntimeError):

  def __init__(self, status: int, message: str):
    super().__init__(f'{DVB_STATUS_NAMES.get(status, status)}: {message}')
    self.status = status


class DvbReadRequirements(C.Structure):
  _fields_ = [(n, C.c_int32) for n in (
      'min_mapping_quality', 'keep_duplicates', 'keep_failed_vendor_quality_checks', 'keep_secondary_alignments',
      'keep_supplementary_alignments', 'keep_unaligned', 'keep_improperly_placed')]


class DvbReadTable(C.Structure):
  _fields_ = [
      ('n_reads', C.c_int32), ('n_refs', C.c_int32), ('n_bases', C.c_int64), ('n_cigar', C.c_int64),
      ('n_name_bytes', C.c_int64), ('n_records_seen', C.c_int64),
      ('ref_id', C.c_void_p), ('pos', C.c_void_p), ('end', C.c_void_p), ('mapq', C.c_void_p), ('flag', C.c_void_p),
      ('fragment_length', C.c_void_p), ('hp', C.c_void_p), ('read_number', C.c_void_p), ('number_reads', C.c_void_p),
      ('seq_begin', C.c_void_p), ('cigar_begin', C.c_void_p), ('name_begin', C.c_void_p),
      ('bases', C.c_void_p), ('quals', C.c_void_p), ('cigar', C.c_void_p), ('names', C.c_void_p),
      ('n_aux_bytes', C.c_int64), ('aux_begin', C.c_void_p), ('aux', C.c_void_p),
  ]


class DvbRegionCandidates(C.Structure):
  _fields_ = [
      ('n_images', C.c_int32), ('ref_id', C.c_void_p), ('variant_start', C.c_void_p), ('variant_end', C.c_void_p),
      ('image_start_pos', C.c_void_p), ('ref_bases', C.c_void_p), ('ref_stride', C.c_int32),
      ('support_begin', C.c_void_p), ('support_class', C.c_void_p), ('support_group', C.c_void_p),
      ('support_name_begin', C.c_void_p), ('support_names', C.c_void_p), ('group_default', C.c_void_p),
  ]


class DvbCandidateOptions(C.Structure):
  _fields_ = [
      ('min_mapping_quality', C.c_int32), ('min_base_quality', C.c_int32), ('keep_legacy_behavior', C.c_int32),
      ('track_ref_reads', C.c_int32), ('min_count_snps', C.c_int32), ('min_count_indels', C.c_int32),
      ('min_fraction_snps', C.c_float), ('min_fraction_indels', C.c_float), ('min_fraction_multiplier', C.c_float),
      ('vsc_min_indel_fraction_for_small_indels', C.c_float), ('vsc_min_indel_fraction_for_large_indels', C.c_float),
      ('vsc_small_indel_threshold', C.c_int32), ('small_model_vaf_context_window_size', C.c_int32),
      ('sample_name', C.c_char_p),
  ]


class DvbSswAlignment(C.Structure):
  _fields_ = [(n, C.c_int32) for n in ('sw_score', 'ref_begin', 'ref_end', 'query_begin', 'query_end', 'mismatches', 'cigar_len')]


class DvbExampleBatchMeta(C.Structure):
  _fields_ = [('variant_blob', C.c_void_p), ('variant_begin', C.c_void_p), ('alt_blob', C.c_void_p), ('alt_begin', C.c_void_p)]


_lib: Optional[C.CDLL] = None

# Every symbol include/dvb.h declares: (name, restype, argtypes).
SYMBOLS = (
    ('dvb_abi_version', C.c_int, []),
    ('dvb_last_error', C.c_char_p, []),
    ('dvb_pileup_params_default', None, [C.POINTER(DvbPileupParams)]),
    ('dvb_image_bytes', C.c_int64, [C.POINTER(DvbPileupParams)]),
    ('dvb_shuffle_table', C.c_int, [C.c_int32, C.c_uint32, C.c_int32, C.c_void_p]),
    ('dvb_encoder_create', C.c_int, [C.POINTER(DvbPileupParams), C.c_int, C.POINTER(C.c_void_p)]),
    ('dvb_encoder_destroy', None, [C.c_void_p]),
    ('dvb_encode_batch_device', C.c_int, [C.c_void_p, C.POINTER(DvbBatch), C.c_void_p, C.c_void_p, C.c_void_p]),
    ('dvb_encoder_check', C.c_int, [C.c_void_p, C.c_void_p]),
    ('dvb_encode_batch_host', C.c_int, [C.c_void_p, C.POINTER(DvbBatch), C.c_void_p, C.c_void_p]),
    ('dvb_encoder_launch_count', C.c_int64, [C.c_void_p]),
    ('dvb_channel_base_modification_plane', C.c_int, [C.c_void_p, C.c_int64, C.c_void_p]),
    ('dvb_channel_hmer_quality_plane', C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int32, C.c_void_p]),
    ('dvb_channel_t0_plane', C.c_int, [C.c_int64, C.c_char_p, C.c_int64, C.c_void_p]),
    ('dvb_channel_allele_frequency_color', C.c_int32, [C.c_float, C.c_float]),
    ('dvb_channel_allele_sample_probability_color', C.c_int32, [C.c_int32, C.c_float]),
    ('dvb_cnn_create', C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                 C.c_int32, C.c_int, C.POINTER(C.c_void_p)]),
    ('dvb_cnn_destroy', None, [C.c_void_p]),
    ('dvb_cnn_forward_device', C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]),
    ('dvb_cnn_forward_host', C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]),
    ('dvb_encode_classify_host', C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(DvbBatch), C.c_void_p, C.c_void_p]),
    ('dvb_cnn_launch_count', C.c_int64, [C.c_void_p]),
    ('dvb_cnn_flops_per_image', C.c_double, [C.c_void_p]),
    ('dvb_cnn_max_batch', C.c_int32, [C.c_void_p]),
    ('dvb_read_requirements_default', None, [C.POINTER(DvbReadRequirements)]),
    ('dvb_bam_open', C.c_int, [C.c_char_p, C.POINTER(DvbReadRequirements), C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    ('dvb_bam_open_regions', C.c_int, [C.c_char_p, C.POINTER(DvbReadRequirements), C.c_int, C.c_int, C.POINTER(C.c_char_p), C.c_void_p, C.c_void_p,
                                       C.c_int32, C.POINTER(C.c_void_p)]),
    ('dvb_cram_to_bam', C.c_int, [C.c_char_p, C.c_char_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32,
                                  C.POINTER(C.c_int64)]),
    ('dvb_bam_derive', C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)]),
    ('dvb_bam_table', C.c_int, [C.c_void_p, C.POINTER(DvbReadTable)]),
    ('dvb_bam_ref_name', C.c_char_p, [C.c_void_p, C.c_int32]),
    ('dvb_bam_ref_length', C.c_int32, [C.c_void_p, C.c_int32]),
    ('dvb_bam_close', None, [C.c_void_p]),
    ('dvb_pack_region_from_bam', C.c_int, [C.c_void_p, C.POINTER(DvbRegionCandidates), C.c_int32, C.c_int32, C.c_int32,
                                           C.c_int32, C.c_int32, C.POINTER(C.c_void_p)]),
    ('dvb_packed_region_batch', C.c_int, [C.c_void_p, C.POINTER(DvbBatch)]),
    ('dvb_packed_region_free', None, [C.c_void_p]),
    ('dvb_candidate_options_default', None, [C.POINTER(DvbCandidateOptions)]),
    ('dvb_candidates_in_region', C.c_int, [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_int64,
                                           C.POINTER(DvbCandidateOptions), C.c_void_p, C.c_int32, C.POINTER(C.c_void_p)]),
    ('dvb_candidate_positions', C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_int64,
                                          C.POINTER(DvbCandidateOptions), C.POINTER(C.c_void_p)]),
    ('dvb_candidates_count', C.c_int64, [C.c_void_p]),
    ('dvb_candidates_protos', C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]),
    ('dvb_candidates_positions', C.c_int64, [C.c_void_p, C.POINTER(C.c_void_p)]),
    ('dvb_candidates_summary_counts', C.c_int64, [C.c_void_p, C.POINTER(C.c_void_p)]),
    ('dvb_candidates_free', None, [C.c_void_p]),
    ('dvb_device_reads_create', C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_void_p)]),
    ('dvb_device_reads_destroy', None, [C.c_void_p]),
    ('dvb_device_reads_launch_count', C.c_int64, [C.c_void_p]),
    ('dvb_allele_count_device', C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_int64,
                                          C.POINTER(DvbCandidateOptions), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    ('dvb_allele_count_host', C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_int64,
                                        C.POINTER(DvbCandidateOptions), C.c_void_p, C.c_void_p]),
    ('dvb_candidates_at_positions', C.c_int, [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_int64,
                                              C.POINTER(DvbCandidateOptions), C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.POINTER(C.c_void_p)]),
    ('dvb_candidates_from_proposed', C.c_int, [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_int64,
                                               C.POINTER(DvbCandidateOptions), C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
                                               C.c_void_p, C.c_char_p, C.POINTER(C.c_void_p)]),
    ('dvb_debug_allele_count_dense_host', C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_int64,
                                                    C.POINTER(DvbCandidateOptions), C.c_int, C.c_void_p, C.c_void_p]),
    ('dvb_debug_allele_counts', C.c_int64, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_int64,
                                            C.POINTER(DvbCandidateOptions), C.c_void_p, C.c_int32, C.c_void_p, C.c_int64]),
    ('dvb_dbg_candidate_haplotypes', C.c_int64, [C.c_char_p, C.c_int64, C.c_char_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32,
                                                 C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_int64, C.c_void_p]),
    ('dvb_debug_read_allele_at', C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_int64, C.c_void_p, C.c_int64, C.c_int64,
                                           C.c_int64, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    ('dvb_encoder_last_pair_support', C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    ('dvb_ssw_align', C.c_int, [C.c_char_p, C.c_int64, C.c_char_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                C.POINTER(DvbSswAlignment), C.c_char_p, C.c_int64]),
    ('dvb_ssw_align_batch', C.c_int, [C.POINTER(C.c_char_p), C.c_void_p, C.POINTER(C.c_char_p), C.c_void_p, C.c_int32, C.c_int32, C.c_int32,
                                      C.c_int32, C.c_int32, C.c_int32, C.POINTER(DvbSswAlignment), C.c_char_p, C.c_int64]),
    ('dvb_fast_pass_scores', C.c_int, [C.c_char_p, C.c_int64, C.POINTER(C.c_char_p), C.c_void_p, C.c_int32, C.POINTER(C.c_char_p), C.c_void_p,
                                       C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    ('dvb_crc32c', C.c_uint32, [C.c_char_p, C.c_size_t]),
    ('dvb_masked_crc32c', C.c_uint32, [C.c_char_p, C.c_size_t]),
    ('dvb_crc32c_portable', C.c_uint32, [C.c_char_p, C.c_size_t]),
    ('dvb_examples_reader_open', C.c_int, [C.POINTER(C.c_char_p), C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_void_p)]),
    ('dvb_examples_reader_shape', C.c_int, [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    ('dvb_examples_reader_next', C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int64, C.POINTER(C.c_int32), C.POINTER(DvbExampleBatchMeta)]),
    ('dvb_examples_reader_close', None, [C.c_void_p]),
    ('dvb_cvo_writer_open', C.c_int, [C.c_char_p, C.c_int32, C.POINTER(C.c_void_p)]),
    ('dvb_cvo_writer_write_batch', C.c_int, [C.c_void_p, C.c_int32, C.POINTER(DvbExampleBatchMeta), C.c_void_p]),
    ('dvb_cvo_writer_close', C.c_int, [C.c_void_p, C.POINTER(C.c_int64)]),
    ('dvb_stream_open', C.c_int, [C.c_char_p, C.c_int32, C.c_int32, C.c_int64, C.POINTER(C.c_void_p)]),
    ('dvb_stream_close', None, [C.c_void_p]),
    ('dvb_stream_remove', C.c_int, [C.c_char_p, C.c_int32]),
    ('dvb_stream_buffer_size', C.c_int64, [C.c_void_p]),
    ('dvb_stream_start', C.c_int, [C.c_void_p]),
    ('dvb_stream_put', C.c_int, [C.c_void_p, C.c_char_p, C.c_int32, C.c_char_p, C.c_int32, C.c_void_p, C.c_int32]),
    ('dvb_stream_end', C.c_int, [C.c_void_p, C.c_int32]),
    ('dvb_stream_shard_finished', C.c_int, [C.c_void_p]),
    ('dvb_stream_wait_attached', C.c_int, [C.c_void_p, C.c_int64]),
    ('dvb_stream_next', C.c_int, [C.POINTER(C.c_void_p), C.c_int32, C.c_int64, C.c_void_p, C.c_int64, C.c_int64, C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                                  C.POINTER(DvbExampleBatchMeta), C.POINTER(C.c_int32)]),
    ('dvb_debug_round_gls', C.c_int, [C.POINTER(C.c_double), C.c_int32, C.POINTER(C.c_double)]),
    ('dvb_debug_upload_phases', C.c_int, [C.POINTER(DvbBatch), C.c_int64, C.c_void_p, C.c_int32]),
    ('dvb_cnn_debug_tensor', C.c_int, [C.c_void_p, C.c_char_p, C.c_int32, C.c_void_p, C.POINTER(C.c_int32),
                                       C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
)


def lib() -> C.CDLL:
  """Loads libdvb.so.  Raises (never falls back) when the CUDA library is absent."""
  global _lib
  if _lib is None:
    if not os.path.exists(LIB_PATH):
      raise RuntimeError(
          f'{LIB_PATH} not found: the CUDA extension is not built. Run '
          '`python -c "import __graft_entry__ as g; g.build()"` at the repo root. '
          'There is no CPU fallback for the product path.')
    l = C.CDLL(LIB_PATH)
    for name, restype, argtypes in SYMBOLS:
      fn = getattr(l, name)  # AttributeError if the .so does not export a declared symbol
      fn.restype = restype
      fn.argtypes = argtypes
    if l.dvb_abi_version() != 4:
      raise RuntimeError('libdvb.so ABI version mismatch')
    _lib = l
  return _lib


def check(status: int) -> None:
  if status != 0:
    msg = lib().dvb_last_error()
    raise DvbError(status, msg.decode() if msg else '')
