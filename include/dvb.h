/*
 * dvb.h — C ABI of the B200-native pileup-encode + CNN path (libdvb.so).
 *
 * This is the drop-in boundary for the one hot path of google/deepvariant v1.10.0
 * that this repository accelerates:
 *
 *   encoder  : replaces  PileupImageEncoderNative::BuildPileupForOneSample
 *              (deepvariant/pileup_image_native.cc:296-447), the per-candidate
 *              driver loop ExamplesGenerator::CreateAndWriteExamplesForCandidate
 *              (deepvariant/make_examples_native.cc:632-736) and the planar→HWC
 *              flatten FillPileupArray (deepvariant/pileup_image_native.h:214-308)
 *              — a whole batch of candidates per call instead of one.
 *   cnn      : replaces the SavedModel call in call_variants.predict_step
 *              (deepvariant/call_variants.py:904-932) = dv_utils.preprocess_images
 *              (deepvariant/dv_utils.py:356-380) + keras_modeling.inceptionv3
 *              (deepvariant/keras_modeling.py:246-336).
 *
 * Plain C: pointers + sizes, no torch / C++ types.  Every function returns a
 * DvbStatus (0 = OK) and never aborts the process (the reference CHECK-fails);
 * dvb_last_error() gives the message of the last failure on the calling thread.
 * Handles are per-device; calls are stream-ordered; the caller owns all buffers.
 */
#ifndef DVB_H_
#define DVB_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DVB_ABI_VERSION 4
#define DVB_MAX_CHANNELS 16

typedef enum DvbStatus {
  DVB_OK = 0,
  DVB_ERR_INVALID_ARGUMENT = 1,
  DVB_ERR_UNSUPPORTED_CHANNEL = 2,
  DVB_ERR_BAD_CIGAR = 3,          /* reference: LOG(FATAL) "Unrecognized CIGAR op", pileup_channel_lib.cc:251 */
  DVB_ERR_TOO_MANY_READS = 4,     /* > max_reads_per_image reads overlap one candidate */
  DVB_ERR_CUDA = 5,
  DVB_ERR_NO_DEVICE = 6,
  DVB_ERR_INTERNAL = 7
} DvbStatus;

/* DeepVariantChannelEnum values (deepvariant/protos/deepvariant.proto:1288-1343)
 * this library computes per read.  Anything else -> DVB_ERR_UNSUPPORTED_CHANNEL. */
enum {
  DVB_CH_READ_BASE = 1,
  DVB_CH_BASE_QUALITY = 2,
  DVB_CH_MAPPING_QUALITY = 3,
  DVB_CH_STRAND = 4,
  DVB_CH_READ_SUPPORTS_VARIANT = 5,
  DVB_CH_BASE_DIFFERS_FROM_REF = 6,
  DVB_CH_HAPLOTYPE_TAG = 7,
  DVB_CH_ALLELE_FREQUENCY = 8,          /* one value per (image, read): DvbBatch.pair_channel[DVB_PAIR_PLANE_ALLELE_FREQUENCY] (allele_frequency_channel.cc:57-118) */
  DVB_CH_READ_MAPPING_PERCENT = 11,     /* "Opt Channels" (deepvariant/pileup_channel_lib.h): whole-read statistics, */
  DVB_CH_AVG_BASE_QUALITY = 12,         /* one constant per read (channels/{read_mapping_percent,avg_base_quality,    */
  DVB_CH_IDENTITY = 13,                 /* identity,gap_compressed_identity,gc_content}_channel.cc)                    */
  DVB_CH_GAP_COMPRESSED_IDENTITY = 14,
  DVB_CH_GC_CONTENT = 15,
  DVB_CH_IS_HOMOPOLYMER = 16,           /* per base: inside a run of >= 3 equal bases (is_homopolymer_channel.cc:77-91) */
  DVB_CH_HOMOPOLYMER_WEIGHTED = 17,     /* per base: length of its run, capped at 30 (homopolymer_weighted_channel.cc:79-101) */
  DVB_CH_BLANK = 18,
  DVB_CH_INSERT_SIZE = 19,
  DVB_CH_MEAN_COVERAGE = 22,            /* blank per read; rows [0, band) = 255 and [band, band + int(mean_coverage)) = 200 painted over the finished
                                           image (pileup_image_native.cc:422-444) */
  DVB_CH_BASE_METHYLATION = 23,         /* per base: DvbBatch.base_channel[DVB_BASE_PLANE_5MC] (base_methylation_channel.cc:54-99) */
  DVB_CH_BASE_6MA = 24,                 /* per base: base_channel[DVB_BASE_PLANE_6MA] (base_6ma_channel.cc:54-99) */
  DVB_CH_READ_SUPPORTS_VARIANT_FUZZY = 25,   /* per (image, read): pair_channel[DVB_PAIR_PLANE_FUZZY_SUPPORT] (read_supports_variant_fuzzy_channel.cc:99-310) */
  DVB_CH_SUPPLEMENTARY_ALIGNMENT = 26,
  DVB_CH_ALLELE_SAMPLE_PROBABILITY = 27,     /* per (image, read): pair_channel[DVB_PAIR_PLANE_ALLELE_SAMPLE_PROBABILITY] (allele_sample_probability_channel.cc:48-101) */
  DVB_CH_HOMOPOLYMER_INSERTION_QUALITY = 28, /* per base, from the reads' tp tag: base_channel[DVB_BASE_PLANE_HMER_INSERTION] (homopolymer_indel_quality_channel.cc:127-183) */
  DVB_CH_HOMOPOLYMER_DELETION_QUALITY = 29,  /* base_channel[DVB_BASE_PLANE_HMER_DELETION] */
  DVB_CH_INTER_HOMOPOLYMER_INSERTION_QUALITY = 30   /* from the t0 tag: base_channel[DVB_BASE_PLANE_INTER_HMER_INSERTION] (inter_homopolymer_insertion_quality_channel.cc:75-127) */
};

/* Channel planes (DvbBatch.pair_channel / base_channel).  These channels are functions of data the pileup loop does not hold -
 * DeepVariantCall maps keyed by read-name strings (allele_frequency, allele_support sizes, ALT_PS phases) or per-base aux tags of
 * the alignment records (MM/ML base modifications, Ultima's tp / t0) - so the caller that owns that data hands the encoder the
 * channel's PIXEL VALUES: one byte per (image, read) pair, or one byte per base parallel to `bases` (0 where a read has no such
 * data: the reference leaves those pixels unwritten).  deepvariant_b200/channels.py restates the reference's value functions;
 * the encoder places the bytes exactly where FillReadBase would (same CIGAR walk, same last-writer-wins order). */
enum { DVB_PAIR_PLANE_ALLELE_FREQUENCY = 0, DVB_PAIR_PLANE_FUZZY_SUPPORT = 1, DVB_PAIR_PLANE_ALLELE_SAMPLE_PROBABILITY = 2, DVB_N_PAIR_PLANES = 3 };
enum { DVB_BASE_PLANE_5MC = 0, DVB_BASE_PLANE_6MA = 1, DVB_BASE_PLANE_HMER_INSERTION = 2, DVB_BASE_PLANE_HMER_DELETION = 3,
       DVB_BASE_PLANE_INTER_HMER_INSERTION = 4, DVB_N_BASE_PLANES = 5 };

/* read_flags bits */
enum {
  DVB_READ_REVERSE_STRAND = 1,  /* alignment.position.reverse_strand */
  DVB_READ_SUPPLEMENTARY = 2,   /* supplementary_alignment */
  DVB_READ_HAS_HP = 4,          /* info["HP"] present with >=1 value; read_hp holds values(0).int_value */
  DVB_READ_HP_MULTI = 8         /* info["HP"] has >1 value: haplotype channel uses 0 (haplotype_tag_channel.cc:83-86) */
};

/* CIGAR is packed like BAM: (length << 4) | op, op = M0 I1 D2 N3 S4 H5 P6 =7 X8. */

/* Mirrors the PileupImageOptions fields the hot path reads
 * (deepvariant/protos/deepvariant.proto PileupImageOptions; defaults in
 * deepvariant/pileup_image.py:36-74). */
typedef struct DvbPileupParams {
  int32_t width;                  /* odd, >= 3 (pileup_image_native.cc:114) */
  int32_t height;                 /* rows per image (single sample) */
  int32_t reference_band_height;
  int32_t num_channels;           /* channels computed per read */
  int32_t channels[DVB_MAX_CHANNELS];
  int32_t num_alt_channels;       /* 0, or 2 for alt-aligned diff/base channels appended last (zero-filled here) */
  int32_t base_color_offset_a_and_g;
  int32_t base_color_offset_t_and_c;
  int32_t base_color_stride;
  float allele_supporting_read_alpha;
  float allele_unsupporting_read_alpha;
  float other_allele_supporting_read_alpha;
  float reference_matching_read_alpha;
  float reference_mismatching_read_alpha;
  int32_t indel_anchoring_base_char;   /* '*' */
  int32_t reference_base_quality;
  int32_t positive_strand_color;
  int32_t negative_strand_color;
  int32_t base_quality_cap;
  int32_t mapping_quality_cap;
  int32_t min_base_quality;       /* read_requirements.min_base_quality */
  int32_t min_mapping_quality;    /* read_requirements.min_mapping_quality */
  int32_t sort_by_haplotypes;
  int32_t hp_tag_for_assembly_polishing;
  int32_t sort_by_alt_allele_support;
  uint32_t random_seed;           /* 2101079370 */
  int32_t max_reads_per_image;    /* capacity of the down-sampling tables; 0 -> 2048 */
  int32_t shuffle_stdlib;         /* which C++ standard library's std::shuffle DownsampleReadIndices uses (it is
                                     implementation-defined): DVB_SHUFFLE_LIBCXX (0, default) reproduces the reference's golden
                                     files (all 51 down-sampled examples of golden.allele_frequency_examples, row for row);
                                     DVB_SHUFFLE_LIBSTDCXX (1) is what a gcc/libstdc++ build of the reference does */
  float mean_coverage;            /* SampleOptions.mean_coverage (--mean_coverage_per_sample); only read with DVB_CH_MEAN_COVERAGE */
  uint32_t blank_channel_mask;    /* bit c set: channels[c] is in the sample's channels_enum_to_blank - its read pixels stay 0, the
                                     reference band is drawn as usual (pileup_channel_lib.cc:152-161; make_examples_native.cc:696-706) */
} DvbPileupParams;

enum { DVB_SHUFFLE_LIBCXX = 0, DVB_SHUFFLE_LIBSTDCXX = 1 };

/* One batch = the images of any number of candidates x alt-allele combinations.
 * All arrays are Structure-of-Arrays; for the *_device entry point every pointer
 * is a device pointer, for *_host every pointer is a host pointer.
 *
 * An "image" is one call of BuildPileupForOneSample: (dv_call, ref_bases, reads,
 * image_start_pos, alt_alleles).  Its reads are given as a CSR list of "pairs"
 * in the order InMemoryReader::Query returns them (make_examples_native.cc:802).
 * pair_support restates ReadSupportsVariantChannel::ReadSupportsAlt
 * (channels/read_supports_variant_channel.cc:75-104) as a class per (image, read):
 * 0 = ref/none, 1 = supports an alt of this image, 2 = supports another alt. */
typedef struct DvbBatch {
  int32_t n_images;
  int32_t n_reads;
  int64_t n_pairs;
  int64_t n_bases;   /* length of bases[] / quals[] */
  int64_t n_cigar;   /* length of cigar[] */
  /* per image */
  const uint8_t* ref_bases;        /* [n_images * ref_stride]; first `width` bytes of each record valid */
  int32_t ref_stride;              /* >= width */
  const int32_t* image_start_pos;  /* [n_images] variant.start - (width-1)/2 */
  const int32_t* variant_start;    /* [n_images] dv_call.variant.start */
  const int64_t* pair_begin;       /* [n_images + 1] */
  /* per pair */
  const int32_t* pair_read;        /* [n_pairs] row of the read table */
  const uint8_t* pair_support;     /* [n_pairs] 0/1/2 */
  const uint8_t* pair_allele_group;/* [n_pairs] or NULL; only read when sort_by_alt_allele_support */
  /* per read */
  const int32_t* read_pos;         /* alignment.position.position (after trimming) */
  const int32_t* read_sort_pos;    /* alignment position before trimming (pileup_image_native.cc:395-398) */
  const int32_t* read_mapq;
  const uint8_t* read_flags;
  const int32_t* read_fragment_length;
  const int32_t* read_hp;
  const uint32_t* read_name_rank;  /* dense rank of (fragment_name, read_number); equal keys share a rank */
  const int64_t* read_seq_begin;   /* [n_reads + 1] into bases / quals */
  const int64_t* read_cigar_begin; /* [n_reads + 1] into cigar */
  const uint8_t* bases;            /* ASCII aligned_sequence */
  const uint8_t* quals;            /* aligned_quality */
  const uint32_t* cigar;
  /* Optional (allele_begin == NULL: pair_support / pair_allele_group above are used as given).  The alt alleles of each
   * image in the allele counter's read-allele form (AlleleCount.read_alleles values, deepvariant.proto: a substitution is
   * its one read base; an insertion / deletion is the anchor base followed by the inserted / deleted bases).  The encoder
   * then derives pair_support and pair_allele_group on the device: each (image, read) pair is walked over the read's CIGAR
   * to the read allele AlleleCounter::Add (allelecounter.cc:880-978) records at variant_start, which is matched against
   * these keys - what DeepVariantCall.allele_support + ReadSupportsAlt (read_supports_variant_channel.cc:75-104) express
   * through read names.  Holds for candidates of the very-sensitive caller over the same reads (dvb_candidates_in_region). */
  const int64_t* allele_begin;       /* [n_images + 1] CSR into the allele arrays */
  const uint8_t* allele_type;        /* [n_alleles] 2 substitution, 3 insertion, 4 deletion (AlleleType) */
  const uint8_t* allele_class;       /* [n_alleles] 1 = an alt allele of this image, 2 = another alt of the candidate */
  const uint8_t* allele_group;       /* [n_alleles] alt index (sort_by_alt_allele_support), or NULL */
  const int64_t* allele_bases_begin; /* [n_alleles + 1] into allele_bases */
  const uint8_t* allele_bases;
  const int32_t* image_ref_run;      /* [n_images] canonical in-contig reference bases after variant_start (a deletion
                                        anchored there is usable iff it is not longer; allelecounter.cc:449-456) */
  const uint8_t* image_group_default;/* [n_images] pair_allele_group of reads that support no alt, or NULL (0) */
  int64_t n_alleles, n_allele_bases;
  int32_t support_min_mapping_quality;  /* AlleleCounterOptions.read_requirements.min_mapping_quality */
  int32_t support_min_base_quality;     /* ... min_base_quality */
  int32_t support_flags;                /* DVB_SUPPORT_* */
  /* Optional channel planes (see DVB_PAIR_PLANE_* / DVB_BASE_PLANE_* above); a plane is required exactly when params.channels names
   * its channel (else DVB_ERR_INVALID_ARGUMENT) and ignored otherwise. */
  const uint8_t* pair_channel[DVB_N_PAIR_PLANES];   /* each [n_pairs] */
  const uint8_t* base_channel[DVB_N_BASE_PLANES];   /* each [n_bases], parallel to bases / quals */
} DvbBatch;

enum {
  DVB_SUPPORT_KEEP_LEGACY = 1,     /* keep_legacy_allele_counter_behavior */
  DVB_SUPPORT_TRACK_REF_READS = 2, /* reference-matching reads also hold an entry at candidate positions */
  DVB_SUPPORT_REPEATED_KEYS = 4    /* some (fragment_name, read_number) occurs on more than one read of the batch: the later
                                      read's entry replaces the earlier one's (a std::map keyed by read name) */
};

typedef struct DvbEncoder DvbEncoder;
typedef struct DvbCnn DvbCnn;

int dvb_abi_version(void);
const char* dvb_last_error(void);

/* Fills *p with pileup_image.default_options() (pileup_image.py:36-74) and the six
 * default channels (dv_constants.py:45-52). */
void dvb_pileup_params_default(DvbPileupParams* p);

/* Bytes of one image: height * width * (num_channels + num_alt_channels). */
int64_t dvb_image_bytes(const DvbPileupParams* p);

/* DownsampleReadIndices (pileup_image_native.cc:153-165): iota(n) shuffled by std::shuffle with a fresh
 * std::mt19937_64(seed), as the chosen standard library implements it (shuffle_stdlib above).  Host only. */
int dvb_shuffle_table(int32_t n, uint32_t seed, int32_t shuffle_stdlib, int32_t* out);

/* device: CUDA ordinal.  Builds the down-sampling tables and uploads constants. */
int dvb_encoder_create(const DvbPileupParams* params, int device, DvbEncoder** out);
void dvb_encoder_destroy(DvbEncoder* enc);

/* All pointers in `batch`, `out` and `rows_kept` are device pointers on the encoder's device.
 * out: uint8[n_images][height][width][num_channels + num_alt_channels].
 * rows_kept: int32[n_images] or NULL — number of read rows each image received (reads that
 *   EncodeRead did not reject, capped at height - reference_band_height).
 * stream: cudaStream_t (NULL = default stream).  Asynchronous. */
int dvb_encode_batch_device(DvbEncoder* enc, const DvbBatch* batch, uint8_t* out,
                            int32_t* rows_kept, void* stream);

/* Device-side error word set by the kernel (bad CIGAR op, too many reads).
 * Synchronises `stream`, returns the status and clears it. */
int dvb_encoder_check(DvbEncoder* enc, void* stream);

/* Host buffers in, host buffer out: validates, stages H2D, encodes, copies D2H,
 * synchronises.  This is the call a reference maintainer would bind. */
int dvb_encode_batch_host(DvbEncoder* enc, const DvbBatch* batch, uint8_t* out_host,
                          int32_t* rows_kept_host /* or NULL */);

/* Number of kernel launches issued by this handle so far (bench bookkeeping). */
int64_t dvb_encoder_launch_count(const DvbEncoder* enc);

/* After a batch with allele keys (DvbBatch.allele_begin): the pair_support / pair_allele_group arrays the device derived for it
 * (host buffers of n_pairs bytes; group may be NULL).  Synchronises the device.  For tests and for callers that want the classes. */
int dvb_encoder_last_pair_support(DvbEncoder* enc, int64_t n_pairs, uint8_t* support, uint8_t* group);

/* ---- value functions of the plane-backed channels (host only; csrc/dvb_channels.cu) ---------------------------------------
 * What the caller puts into DvbBatch.base_channel / pair_channel, with the reference's float arithmetic. */
/* base_methylation_channel.cc:87-99 / base_6ma_channel.cc:87-99 ScaleColorVector(values, 255): values = the modification's ML bytes
 * per base of the aligned sequence (nucleus Read.base_modifications["5mC" | "6mA"]). */
int dvb_channel_base_modification_plane(const uint8_t* values, int64_t len, uint8_t* out);
/* homopolymer_indel_quality_channel.cc:127-183 HomoPolymerInDelQuality: tp = the read's tp tag as int8 per base (NULL = no tag),
 * is_deletion 0 = homopolymer_insertion_quality, 1 = homopolymer_deletion_quality. */
int dvb_channel_hmer_quality_plane(const uint8_t* seq, const uint8_t* qual, int64_t len, const int8_t* tp, int32_t is_deletion, uint8_t* out);
/* inter_homopolymer_insertion_quality_channel.cc:75-127 GetT0QualityValues: t0 = the t0 tag's text (phred + 33), t0_len 0 = no tag. */
int dvb_channel_t0_plane(int64_t len, const char* t0, int64_t t0_len, uint8_t* out);
/* allele_frequency_channel.cc:76-87 AlleleFrequencyColor. */
int32_t dvb_channel_allele_frequency_color(float allele_frequency, float min_non_zero_allele_frequency);
/* allele_sample_probability_channel.cc:87-101 ScaleColor(reads supporting the read's allele, all reads). */
int32_t dvb_channel_allele_sample_probability_color(int32_t value, float max_val);

/* ---- CNN (Inception-v3 + genotype softmax) ------------------------------- */

/* Weights blob layout: see deepvariant_b200/modeling.py (pack_weights): a flat
 * little-endian stream of per-layer folded conv weights (fp16, [Cout][KH][KW][Cin_pad])
 * and fp32 biases in network order, then the dense 2048x3 fp32 head.
 * precision: 0 = fp16 operands / fp32 accumulate (single pass),
 *            1 = split-fp16 x3 (fp32-grade products). */
int dvb_cnn_create(const void* weights, int64_t weights_bytes, int32_t height, int32_t width,
                   int32_t channels, int32_t max_batch, int32_t precision, int device, DvbCnn** out);
void dvb_cnn_destroy(DvbCnn* cnn);

/* images: device uint8 [n][H][W][C] (the encoder's output, consumed in place);
 * probs: device float [n][3].  Asynchronous on `stream`. */
int dvb_cnn_forward_device(DvbCnn* cnn, const uint8_t* images, int32_t n, float* probs, void* stream);

/* Host in / host out variant. */
int dvb_cnn_forward_host(DvbCnn* cnn, const uint8_t* images_host, int32_t n, float* probs_host);

/* The fused hot path a reference maintainer binds when make_examples and call_variants run in one process
 * (the reference's --fast_pipeline mode streams examples through shared memory instead,
 * deepvariant/fast_pipeline.cc + stream_examples.cc:94-156): host DvbBatch in, genotype probabilities out.
 * Stages the batch (one H2D copy), encodes on the device, classifies the images where they lie in HBM
 * (the 155 KB/example image/encoded bytes never cross PCIe), copies float[n_images][3] back, synchronises.
 * rows_kept_host: int32[n_images] or NULL.  enc and cnn must live on the same device. */
int dvb_encode_classify_host(DvbEncoder* enc, DvbCnn* cnn, const DvbBatch* batch_host, float* probs_host,
                             int32_t* rows_kept_host);

int64_t dvb_cnn_launch_count(const DvbCnn* cnn);
/* Images per internal chunk of a forward (the max_batch the handle was created with). */
int32_t dvb_cnn_max_batch(const DvbCnn* cnn);
/* FLOPs of one forward for one image (conv MACs x 2). */
double dvb_cnn_flops_per_image(const DvbCnn* cnn);

/* ---- file boundary helpers (host only) ------------------------------------- */
/* CRC-32C (Castagnoli) and TensorFlow's masked variant used by the TFRecord framing
 * (third_party/nucleus/io/example_writer.cc:99-115 -> tensorflow RecordWriter). */
uint32_t dvb_crc32c(const void* data, size_t n);            /* SSE4.2 crc32 instruction when the CPU has it */
uint32_t dvb_crc32c_portable(const void* data, size_t n);   /* slicing-by-8 tables; same value (tests compare the two) */
uint32_t dvb_masked_crc32c(const void* data, size_t n);

/* ---- BAM -> Structure-of-Arrays read table (host only; SURVEY.md 8(f) "next" row #1) ------------------------
 * Replaces nucleus SamReader::Iterate + ConvertToPb + ReadSatisfiesRequirements
 * (third_party/nucleus/io/sam_reader.cc:760-975, 1065-1135, 217-245; third_party/nucleus/util/utils.cc:255-266)
 * for the pileup path: the file is inflated block-parallel and every record that passes the filter is decoded
 * once into flat arrays — the layout DvbBatch's per-read arrays are gathered from — instead of one Read proto each. */
typedef struct DvbReadRequirements {   /* third_party/nucleus/protos/reads.proto ReadRequirements */
  int32_t min_mapping_quality;         /* make_examples default 5 (make_examples_options.py:957-964) */
  int32_t keep_duplicates;
  int32_t keep_failed_vendor_quality_checks;
  int32_t keep_secondary_alignments;
  int32_t keep_supplementary_alignments;
  int32_t keep_unaligned;
  int32_t keep_improperly_placed;
} DvbReadRequirements;

typedef struct DvbReadTable {          /* all pointers are owned by the DvbBam handle */
  int32_t n_reads;
  int32_t n_refs;
  int64_t n_bases, n_cigar, n_name_bytes;
  int64_t n_records_seen;              /* alignment records in the file before filtering */
  const int32_t* ref_id;               /* [n_reads] index into the header's reference list, -1 unmapped */
  const int32_t* pos;                  /* alignment.position.position (0-based) */
  const int32_t* end;                  /* ReadEnd: pos + sum of M/D/N/=/X lengths (utils.cc:222-240) */
  const uint8_t* mapq;
  const uint16_t* flag;                /* raw BAM FLAG */
  const int32_t* fragment_length;      /* isize */
  const int32_t* hp;                   /* HP aux tag (integer), INT32_MIN when absent or not parsed */
  const uint8_t* read_number;          /* 0 if FREAD1 or unpaired, else 1 (sam_reader.cc:786-793) */
  const uint8_t* number_reads;         /* 2 if paired else 1 */
  const int64_t* seq_begin;            /* [n_reads + 1] into bases / quals */
  const int64_t* cigar_begin;          /* [n_reads + 1] into cigar (empty for unmapped reads) */
  const int64_t* name_begin;           /* [n_reads + 1] into names */
  const uint8_t* bases;                /* ASCII from "=ACMGRSVTWYHKDBN" */
  const uint8_t* quals;                /* raw phred */
  const uint32_t* cigar;               /* BAM packing (len << 4 | op) */
  const char* names;                   /* concatenated QNAMEs, no terminators */
  int64_t n_aux_bytes;                 /* parse_hp & 2: the records' raw aux fields (BAM encoding), for tags the caller parses itself - */
  const int64_t* aux_begin;            /* [n_reads + 1] MM / ML / MN -> Read.base_modifications (sam_reader.cc:521-716), Ultima's tp / t0 */
  const uint8_t* aux;
} DvbReadTable;

typedef struct DvbBam DvbBam;
void dvb_read_requirements_default(DvbReadRequirements* r);
/* req may be NULL (defaults).  parse_hp: bit 0 extracts the HP aux tag, bit 1 keeps every record's raw aux bytes (DvbReadTable.aux).
 * threads <= 0: hardware concurrency. */
int dvb_bam_open(const char* path, const DvbReadRequirements* req, int parse_hp, int threads, DvbBam** out);
/* The same, restricted to the reads that overlap one of n_regions half-open [start, end) intervals (ReadOverlapsRegion,
 * third_party/nucleus/util/utils.cc:172-188) - what sam_reader.cc:Query returns for make_examples --regions.  When all intervals lie on
 * one contig and `path`.bai (or path with .bai for .bam) exists, decoding starts at the block the linear index gives for the first
 * interval and stops at the first read behind the last one (coordinate-sorted file); otherwise the file is scanned and filtered.
 * The file is read, inflated and parsed in 32-MB rounds either way (bounded memory besides the table itself). */
int dvb_bam_open_regions(const char* path, const DvbReadRequirements* req, int parse_hp, int threads, const char* const* contigs,
                         const int64_t* starts, const int64_t* ends, int32_t n_regions, DvbBam** out);
int dvb_bam_table(const DvbBam* bam, DvbReadTable* table);
/* CRAM 3.0 input: decodes `cram_path` (all of it, or with n_regions > 0 only the containers that overlap one of the half-open
 * intervals) into an uncompressed BAM at `bam_path`, which dvb_bam_open / dvb_bam_open_regions then read - what hts_open + sam_read1
 * give the reference for a CRAM (third_party/nucleus/io/sam_reader.cc:325-399, which also wants the FASTA there).  The reference
 * contigs the reads were compressed against come as whole upper- or lower-case sequences by name (only contigs the wanted containers
 * touch need to be present; a missing one is an error when a read needs it).  gzip and rANS 4x8 blocks; bzip2 / lzma blocks and the
 * 3.1 codecs are reported as errors.  *n_records_out = alignment records written. */
int dvb_cram_to_bam(const char* cram_path, const char* bam_path, const char* const* ref_names, const uint8_t* const* ref_bases,
                    const int64_t* ref_lens, int32_t n_refs, const char* const* region_contigs, const int64_t* region_starts,
                    const int64_t* region_ends, int32_t n_regions, int64_t* n_records_out);
/* A new table holding rows[0 .. n_rows) of `src` in that order; row i takes the alignment new_pos[i] + new_cigar[new_cigar_begin[i] ..
 * new_cigar_begin[i + 1]) (BAM packing) when that range is not empty and keeps its own otherwise (new_cigar_begin NULL: all kept).
 * Replaces in_memory_sam_reader.replace_reads (deepvariant/make_examples_core.py:2290-2300; third_party/nucleus/io/sam.py:357-361):
 * realigned / normalised reads as a table the candidate generator and the region packer take.  Close with dvb_bam_close. */
int dvb_bam_derive(const DvbBam* src, const int64_t* rows, int64_t n_rows, const int32_t* new_pos, const int64_t* new_cigar_begin,
                   const uint32_t* new_cigar, DvbBam** out);
const char* dvb_bam_ref_name(const DvbBam* bam, int32_t i);   /* NULL when out of range */
int32_t dvb_bam_ref_length(const DvbBam* bam, int32_t i);     /* l_ref of the header's reference i (-1 when out of range) */
void dvb_bam_close(DvbBam* bam);

/* ---- region packer: candidates + BAM table -> DvbBatch on the host (SURVEY.md 8(f) "next" row #1, second half) ---------
 * Restates what CreateAndWriteExamplesForCandidate does per candidate before the pixels
 * (deepvariant/make_examples_native.cc:632-736): the InMemoryReader::Query scan over the region's reads (:802-810,
 * third_party/nucleus/util/utils.cc:172-188), ReadSupportsAlt's read-name search
 * (deepvariant/channels/read_supports_variant_channel.cc:75-104) and the allele-group map
 * (deepvariant/pileup_image_native.cc:345-360), over the flat read table of a DvbBam.  The caller enumerates images
 * (candidate x alt combination), fetches the reference windows (FASTA) and flattens allele_support into entries. */
typedef struct DvbRegionCandidates {
  int32_t n_images;
  const int32_t* ref_id;              /* [n_images] contig of the candidate (index into the BAM header) */
  const int32_t* variant_start;       /* [n_images] */
  const int32_t* variant_end;         /* [n_images] */
  const int32_t* image_start_pos;     /* [n_images] variant.start - (width - 1) / 2 */
  const uint8_t* ref_bases;           /* [n_images * ref_stride] reference windows */
  int32_t ref_stride;
  /* read support, per image: entries in alt order (all names of alt 0, then alt 1, ...); the first entry naming a read
   * decides its class; keys are "fragment_name/read_number" */
  const int64_t* support_begin;       /* [n_images + 1] CSR into the entry arrays */
  const uint8_t* support_class;       /* [n_entries] 1 = this image's alt set, 2 = another alt of the variant */
  const uint8_t* support_group;       /* [n_entries] allele group (alt index) or NULL when not sorting by allele support */
  const int64_t* support_name_begin;  /* [n_entries + 1] into support_names */
  const char* support_names;
  const uint8_t* group_default;       /* [n_images] group of reads no alt names (= number of alts) or NULL */
} DvbRegionCandidates;

typedef struct DvbPackedRegion DvbPackedRegion;
int dvb_pack_region_from_bam(const DvbBam* bam, const DvbRegionCandidates* candidates, int32_t region_ref_id, int32_t region_start,
                             int32_t region_end, int32_t read_overlap_buffer_bp, int32_t width, DvbPackedRegion** out);
int dvb_packed_region_batch(const DvbPackedRegion* packed, DvbBatch* batch /* host pointers owned by `packed` */);
void dvb_packed_region_free(DvbPackedRegion* packed);

/* ---- candidate generation on the host (SURVEY.md 8(f) "next" row #2): allele counting + very-sensitive caller ------------
 * Replaces, for one sample and the make_examples defaults (--normalize_reads rewrites the reads before this call, normalize_reads.py; no complex / rejected alleles, no
 * methylation), the pybind modules deepvariant.python.allelecounter (AlleleCounter(ref, range, candidate_positions,
 * options).add(read, sample); deepvariant/python/allelecounter_pybind.cc, deepvariant/allelecounter.cc:880-978) and
 * deepvariant.python.variant_calling_multisample (VariantCaller.calls_from_allele_counts /
 * call_positions_from_allele_counts; deepvariant/variant_calling_multisample.cc:1000-1330) as make_examples_core.py
 * candidates_in_region (:2832-2960) drives them.  Reads are rows of a DvbBam table, given in the order the reference
 * would iterate them (InMemorySamReader.query(region) after reservoir sampling). */
typedef struct DvbCandidateOptions {
  int32_t min_mapping_quality;        /* AlleleCounterOptions.read_requirements (make_examples_options.py:293-310): 5 */
  int32_t min_base_quality;           /* 10 */
  int32_t keep_legacy_behavior;       /* --keep_legacy_allele_counter_behavior: 0 */
  int32_t track_ref_reads;            /* --track_ref_reads: 0 (1 for PACBIO / ONT) */
  int32_t min_count_snps;             /* --vsc_min_count_snps 2 */
  int32_t min_count_indels;           /* --vsc_min_count_indels 2 */
  float min_fraction_snps;            /* --vsc_min_fraction_snps 0.12 (float in VariantCallerOptions: compared as (double)0.12f) */
  float min_fraction_indels;          /* --vsc_min_fraction_indels 0.06 */
  float min_fraction_multiplier;      /* --vsc_min_fraction_multiplier 1.0 */
  float vsc_min_indel_fraction_for_small_indels;   /* 0 = unused */
  float vsc_min_indel_fraction_for_large_indels;
  int32_t vsc_small_indel_threshold;
  int32_t small_model_vaf_context_window_size;     /* > 0 fills DeepVariantCall.allele_frequency_at_position */
  const char* sample_name;            /* call_set_name and ReadSupport.sample_name; may be NULL */
} DvbCandidateOptions;
typedef struct DvbCandidates DvbCandidates;
void dvb_candidate_options_default(DvbCandidateOptions* options);
/* contig_bases = the whole contig, upper-cased (the reference reads its FASTA with keep_true_case = false); [start, end)
 * is the allele counter's interval; rows = BAM table rows.  candidate_positions (absolute, may be NULL) are the positions
 * whose reference-supporting reads are tracked when track_ref_reads is set (second pass of candidates_in_region). */
int dvb_candidates_in_region(const DvbBam* bam, const char* reference_name, const uint8_t* contig_bases, int64_t contig_n_bases,
                             int64_t start, int64_t end, const int64_t* rows, int64_t n_rows, const DvbCandidateOptions* options,
                             const int32_t* candidate_positions, int32_t n_candidate_positions, DvbCandidates** out);
/* First pass of track_ref_reads: positions with at least one selected alt allele (call_positions_from_allele_counts). */
int dvb_candidate_positions(const DvbBam* bam, const uint8_t* contig_bases, int64_t contig_n_bases, int64_t start, int64_t end,
                            const int64_t* rows, int64_t n_rows, const DvbCandidateOptions* options, DvbCandidates** out);
int64_t dvb_candidates_count(const DvbCandidates* candidates);
/* Serialized learning.genomics.deepvariant.DeepVariantCall records (deepvariant/protos/deepvariant.proto:262-317),
 * concatenated in coordinate order; begin = int64[count + 1].  Owned by `candidates`. */
int dvb_candidates_protos(const DvbCandidates* candidates, const uint8_t** data, const int64_t** begin);
/* variant.start of every candidate (or the positions of dvb_candidate_positions); returns the count. */
int64_t dvb_candidates_positions(const DvbCandidates* candidates, const int32_t** positions);
/* AlleleCounter::SummaryCounts (deepvariant/allelecounter.cc:986-1007) of the counter the candidates came from: int32
 * [2 * n] = (ref_supporting_read_count, total_read_count) of every position of [start, end), what VariantCaller.make_gvcfs
 * (deepvariant/variant_caller.py:256-413) turns into gVCF reference blocks.  Returns n.  Owned by `candidates`. */
int64_t dvb_candidates_summary_counts(const DvbCandidates* candidates, const int32_t** counts);
void dvb_candidates_free(DvbCandidates* candidates);
/* Test access to the allele counter (AlleleCounter::Counts()): JSON text, one object per position of [start, end):
 * {"ref": ref_supporting_read_count, "alleles": [[bases, AlleleType, is_low_quality, read key, mapping quality,
 * avg base quality, reverse strand], ...]} = the entries of AlleleCount.read_alleles in insertion order.  Returns the text
 * length (written, NUL-terminated, when cap is larger) or -DvbStatus. */
int64_t dvb_debug_allele_counts(const DvbBam* bam, const uint8_t* contig_bases, int64_t contig_n_bases, int64_t start, int64_t end,
                                const int64_t* rows, int64_t n_rows, const DvbCandidateOptions* options,
                                const int32_t* candidate_positions, int32_t n_candidate_positions, char* out, int64_t cap);

/* ---- local realigner, host side (the graph in native code) ----------------------------------
 * Candidate haplotypes of one window from the de Bruijn graph of its reference and reads (deepvariant/realigner/debruijn_graph.cc:
 * Build :224-248, AddEdges :250-300, Prune :302-345, CandidateHaplotypes :347-391): sorted, '\n'-separated, written to `out` when
 * they fit `cap`.  Returns the bytes needed (>= 1; 1 = a graph without a path), 0 when no k gives an acyclic graph (the caller
 * keeps the reference alone), or -DvbStatus.  reads: bases / qualities concatenated, read i at [read_begin[i], read_begin[i + 1]). */
int64_t dvb_dbg_candidate_haplotypes(const char* ref, int64_t ref_len, const char* bases, const uint8_t* quals, const int64_t* read_begin,
                                     const int32_t* mapq, int32_t n_reads, int32_t min_k, int32_t max_k, int32_t step_k, int32_t min_mapq,
                                     int32_t min_base_quality, int32_t min_edge_weight, int32_t max_num_paths, char* out, int64_t cap,
                                     int32_t* k_used);

/* Test access to the (candidate, read) support walk of the encoder's pre-pass (DvbBatch.allele_begin): the read allele of one read
 * at `target`, (a) walk_out: the last commit of the full allele-counter walk over [start, end) at that position, (b) at_out: what the
 * CIGAR-only walk used on the device finds.  Each int32[6] = {found, AlleleType, is_low_quality, anchor base, read offset, length}. */
int dvb_debug_read_allele_at(const uint8_t* seq, const uint8_t* qual, int32_t seq_len, const uint32_t* cigar, int32_t n_cigar, int64_t pos,
                             const uint8_t* contig, int64_t contig_len, int64_t start, int64_t end, int64_t target, int32_t min_base_quality,
                             int32_t keep_legacy, int32_t* walk_out, int32_t* at_out);

/* ---- allele counting on the device (SURVEY.md 8(f) "next" row #2, device half) ------------------------------------------------
 * The reads of a DvbBam table are uploaded once (Structure of Arrays in HBM); dvb_allele_count_* runs AlleleCounter::Add
 * (deepvariant/allelecounter.cc:880-978) with one thread per read and atomics into dense per-position counters, then flags the
 * positions that can carry a candidate: bit 0 = a substitution allele passes the count / ratio tests of
 * IsGoodAltAlleleWithReason (variant_calling_multisample.cc:175-196; 10 % slack on the ratio), bit 1 = an insertion or
 * deletion is anchored there.  The flags are a superset of CallVariant's positions; dvb_candidates_at_positions takes the exact
 * decision on the flagged sites from the reads that overlap them.
 *   counts = int32[6 * len]: ref_supporting_read_count[len], substitution counts [4 * len] by read base A, C, G, T
 *            (non-low-quality), other[len] = non-low-quality insertion / deletion / soft-clip entries;  flags = uint8[len]. */
typedef struct DvbDeviceReads DvbDeviceReads;
int dvb_device_reads_create(const DvbBam* bam, int device, DvbDeviceReads** out);     /* DVB_ERR_NO_DEVICE without a GPU */
void dvb_device_reads_destroy(DvbDeviceReads* reads);
int64_t dvb_device_reads_launch_count(const DvbDeviceReads* reads);
/* Device pointers, asynchronous on `stream`: ref_dev[0] is absolute position ref_origin, ref_avail bases resident (must cover
 * every read of rows_dev by one base on either side); indel_dev = uint8[len] scratch. */
int dvb_allele_count_device(DvbDeviceReads* reads, const uint8_t* ref_dev, int64_t ref_origin, int64_t ref_avail, int64_t contig_n_bases,
                            int64_t start, int64_t end, const int64_t* rows_dev, int64_t n_rows, const DvbCandidateOptions* options,
                            int32_t* counts_dev, uint8_t* indel_dev, uint8_t* flags_dev, void* stream);
/* Host pointers, synchronous: uploads the rows and the reference window, counts, flags, copies the results back. */
int dvb_allele_count_host(DvbDeviceReads* reads, const DvbBam* bam, const uint8_t* contig_bases, int64_t contig_n_bases, int64_t start,
                          int64_t end, const int64_t* rows, int64_t n_rows, const DvbCandidateOptions* options, int32_t* counts_host,
                          uint8_t* flags_host);
/* dvb_candidates_in_region restricted to `emit_positions` (absolute, sorted): rows need only hold the reads that overlap
 * those positions (by one base on either side, plus the allele-frequency context when it is requested). */
int dvb_candidates_at_positions(const DvbBam* bam, const char* reference_name, const uint8_t* contig_bases, int64_t contig_n_bases,
                                int64_t start, int64_t end, const int64_t* rows, int64_t n_rows, const DvbCandidateOptions* options,
                                const int32_t* candidate_positions, int32_t n_candidate_positions, const int32_t* emit_positions,
                                int32_t n_emit_positions, DvbCandidates** out);
/* --variant_caller vcf_candidate_importer: one DeepVariantCall per PROPOSED variant, with this region's read evidence attached
 * (replaces VariantCaller::CallsFromVcf -> ComputeVariant, deepvariant/variant_calling.cc:393-435, 493-541, bound by
 * deepvariant/vcf_candidate_importer.py:62-70).  The caller passes the VCF records that start inside [start, end), in file order:
 * variant v has the alleles allele_first[v] .. allele_first[v + 1] - 1 (the first one is its reference allele), allele a is
 * allele_chars[allele_begin[a] .. allele_begin[a + 1]).  A record whose reference allele contradicts the reads' is an error
 * (the reference QCHECK-fails); a non-canonical reference base drops the record. */
int dvb_candidates_from_proposed(const DvbBam* bam, const char* reference_name, const uint8_t* contig_bases, int64_t contig_n_bases,
                                 int64_t start, int64_t end, const int64_t* rows, int64_t n_rows, const DvbCandidateOptions* options,
                                 const int32_t* candidate_positions, int32_t n_candidate_positions, int32_t n_proposed,
                                 const int64_t* proposed_start, const int32_t* allele_first, const int64_t* allele_begin,
                                 const char* allele_chars, DvbCandidates** out);
/* Test access: the kernels' walk, sink and flag function instantiated on the host (windowed != 0: with the reference window
 * the device path uploads instead of the whole contig). */
int dvb_debug_allele_count_dense_host(const DvbBam* bam, const uint8_t* contig_bases, int64_t contig_n_bases, int64_t start, int64_t end,
                                      const int64_t* rows, int64_t n_rows, const DvbCandidateOptions* options, int windowed,
                                      int32_t* counts_host, uint8_t* flags_host);

/* ---- Smith-Waterman with libssw's tie-breaking (deepvariant/realigner/ssw.h: Aligner(match, mismatch, gap_open, gap_extend),
 * SetReferenceSequence, Align(query, Filter(), maskLen, &alignment)) - the aligner behind alt-aligned pileups and the realigner
 * (SURVEY.md 8(f) "next" row #3).  Host code.  cigar_out receives the ssw_cpp cigar_string ("8S4=1X4=1D5=17S"), NUL-terminated,
 * when cigar_cap > cigar_len. */
typedef struct DvbSswAlignment {
  int32_t sw_score, ref_begin, ref_end, query_begin, query_end, mismatches, cigar_len;
} DvbSswAlignment;
int dvb_ssw_align(const char* ref, int64_t ref_len, const char* query, int64_t query_len, int32_t match, int32_t mismatch, int32_t gap_open,
                  int32_t gap_extend, DvbSswAlignment* out, char* cigar_out, int64_t cigar_cap);
/* n alignments in one launch on `device` (csrc/dvb_ssw_gpu.cu): the two full-matrix Smith-Waterman scans of every (reference, query) pair
 * run on the GPU, a warp per alignment as an anti-diagonal wavefront; the banded traceback inside the window they delimit runs on the
 * host.  The reference side is FastPassAligner::SswAlignReadsToHaplotypes (deepvariant/realigner/fast_pass_aligner.cc:281-322) and
 * RealignReadsToHaplotype (deepvariant/alt_aligned_pileup_lib.cc:278-313), which call ssw once per (haplotype, read).  Results equal
 * dvb_ssw_align's field for field.  cigars = char[n][cigar_stride] (NUL-terminated; empty when the string does not fit - cigar_len says
 * how long it is). */
int dvb_ssw_align_batch(const char* const* refs, const int64_t* ref_lens, const char* const* queries, const int64_t* query_lens, int32_t n,
                        int32_t match, int32_t mismatch, int32_t gap_open, int32_t gap_extend, int32_t device, DvbSswAlignment* out,
                        char* cigars, int64_t cigar_stride);

/* FastPassAligner's exact k-mer pass (deepvariant/realigner/fast_pass_aligner.cc:165-279) for all haplotypes in one call:
 * hap_score = int32[n_haps]; position / score = int32[n_haps * n_reads] (position 65535 = read not placed on that haplotype). */
int dvb_fast_pass_scores(const char* reference, int64_t ref_len, const char* const* haplotypes, const int64_t* hap_lens, int32_t n_haps,
                         const char* const* reads, const int64_t* read_lens, int32_t n_reads, int32_t kmer_size, int32_t max_mismatches,
                         int32_t match, int32_t mismatch, int32_t ref_prefix_len, int32_t ref_suffix_len, int32_t* hap_score,
                         int32_t* position, int32_t* score);

/* ---- call_variants record I/O on the host ---------------------------------------------
 * Reader = call_variants.get_dataset (deepvariant/call_variants.py:449-538): the shards of the examples TFRecord
 * (gzip or plain) are read by `threads` workers and handed out in tf.data's deterministic interleave order
 * (cycle_length slots, one record per slot per turn; call_variants.py:83 uses 32).  Each record must hold exactly one
 * bytes value for image/encoded, variant/encoded and alt_allele_indices/encoded (parse_single_example with
 * FixedLenFeature((), string)).  Record CRCs are verified when verify_crc != 0 (TensorFlow's reader always does). */
typedef struct DvbExamplesReader DvbExamplesReader;
typedef struct DvbExampleBatchMeta {
  const uint8_t* variant_blob;    /* variant/encoded of the batch, concatenated */
  const int64_t* variant_begin;   /* [n + 1] */
  const uint8_t* alt_blob;        /* alt_allele_indices/encoded, concatenated */
  const int64_t* alt_begin;       /* [n + 1] */
} DvbExampleBatchMeta;
int dvb_examples_reader_open(const char* const* paths, int32_t n_paths, int32_t threads /* 0 = all cores */,
                             int32_t cycle_length /* 0 = 32 */, int32_t verify_crc, DvbExamplesReader** out);
/* image/shape and the byte size of image/encoded of the first record (all zero when there are no records). */
int dvb_examples_reader_shape(DvbExamplesReader* reader, int64_t shape[3], int64_t* image_bytes);
/* Copies the next up-to-max_n images into images_host (n * image_bytes, e.g. a pinned staging buffer); *n_out = 0 at the
 * end.  `meta` points into reader-owned memory that stays valid until the next call on this reader. */
int dvb_examples_reader_next(DvbExamplesReader* reader, int32_t max_n, uint8_t* images_host, int64_t image_bytes,
                             int32_t* n_out, DvbExampleBatchMeta* meta);
void dvb_examples_reader_close(DvbExamplesReader* reader);

/* Writer = write_variant_call / _create_cvo_proto / round_gls (deepvariant/call_variants.py:248-399): per record
 * CallVariantsOutput{variant with calls[0].info["MID"] = "deepvariant", alt_allele_indices, genotype_probabilities =
 * round_gls(float64(probs), gl_precision)} framed as a TFRecord, gzip when the path ends in ".gz".  One writer = one
 * output shard with its own thread (the reference's writer processes, call_variants.py:541-602); write_batch copies its
 * arguments and returns.  gl_precision < 0 = no rounding.  Errors (likelihoods not summing to 1 within 1e-6, a variant
 * without calls, I/O) are reported by the next write_batch or by close. */
typedef struct DvbCvoWriter DvbCvoWriter;
int dvb_cvo_writer_open(const char* path, int32_t gl_precision, DvbCvoWriter** out);
int dvb_cvo_writer_write_batch(DvbCvoWriter* writer, int32_t n, const DvbExampleBatchMeta* meta, const float* probs /* [n, 3] host */);
int dvb_cvo_writer_close(DvbCvoWriter* writer, int64_t* n_written /* may be NULL */);
/* Test access to the writer's round_gls (call_variants.py:248-285); precision < 0 = none. */
int dvb_debug_round_gls(const double gls[3], int32_t precision, double out[3]);

/* ---- the reference's shared-memory example stream, both ends (SURVEY.md 8(b) optional row; csrc/dvb_stream.cu) ----------------------
 * make_examples --stream_examples writes { int32 len, alt_allele_indices, int32 len, variant, int32 len, image } records, closed by
 * int32 0, into one POSIX shared-memory buffer per shard ("<prefix>_shm_<shard>") and hands it over through three named mutexes
 * ("<prefix>_buffer_empty_<shard>", "_items_available_", "_shard_finished_"; boost::interprocess named_mutex = a named semaphore of
 * count 1): deepvariant/stream_examples.cc:60-176, stream_examples_kernel.cc:166-240, fast_pipeline.cc:125-165,
 * fast_pipeline_utils.h:44-64.  A producer handle here can feed the reference's call_variants, a consumer handle can be fed by the
 * reference's make_examples; the orchestrator role creates (and dvb_stream_remove deletes) the objects as fast_pipeline does. */
enum { DVB_STREAM_ORCHESTRATOR = 0, DVB_STREAM_PRODUCER = 1, DVB_STREAM_CONSUMER = 2 };
typedef struct DvbStream DvbStream;
int dvb_stream_open(const char* shm_prefix, int32_t shard, int32_t role, int64_t buffer_size /* orchestrator only */, DvbStream** out);
void dvb_stream_close(DvbStream* stream);
int dvb_stream_remove(const char* shm_prefix, int32_t shard);
int64_t dvb_stream_buffer_size(const DvbStream* stream);
/* producer: StartStreaming / StreamExample / EndStreaming once per region, SignalShardFinished at the end of the shard */
int dvb_stream_start(DvbStream* stream);
int dvb_stream_put(DvbStream* stream, const void* alt_indices, int32_t alt_len, const void* variant, int32_t variant_len, const uint8_t* image,
                   int32_t image_len);
int dvb_stream_end(DvbStream* stream, int32_t data_written);
int dvb_stream_shard_finished(DvbStream* stream);
/* consumer, before its first dvb_stream_next: wait until the shard's producer has attached (it holds items_available and shard_finished,
 * or buffer_empty).  The reference relies on call_variants starting slowly instead. */
int dvb_stream_wait_attached(DvbStream* stream, int64_t timeout_ms);
/* consumer: StreamExamplesResource::Next over n shard handles, polling from shard index % n: the first ready buffer is drained - images
 * copied to images_host (images_cap bytes), records readable through *meta until the next call that drains the same shard.
 * *n_out = 0 with *all_finished = 1: every shard is done; *n_out = 0 with *all_finished = 0: a shard has just finished, call again. */
int dvb_stream_next(DvbStream* const* shards, int32_t n, int64_t index, uint8_t* images_host, int64_t images_cap, int64_t image_bytes,
                    int32_t* n_out, int32_t* shard_out, DvbExampleBatchMeta* meta, int32_t* all_finished);

/* Test access to the chunked-upload plan of dvb_encode_classify_host for phases of `sub` images: out = int64[cap][6] =
 * {image begin, image end, pair begin, pair end, first read uploaded, one past the last read uploaded}.  Returns the number
 * of phases (0 = the batch is uploaded in one piece) or -DvbStatus.  Host only, no device needed. */
int dvb_debug_upload_phases(const DvbBatch* batch_host, int64_t sub, int64_t* out, int32_t cap);

/* Debug / test access to an intermediate activation of the LAST forward (first `n` images of the
 * last chunk), converted to float NHWC: out_host = float[n][H][W][C].  Names follow
 * deepvariant_b200/modeling.py ("input", "s1".."s5", "p1", "p2", "mixed0".."mixed10", branch
 * tensors); "pooled" returns the 2048-wide pre-logits.  out_host may be NULL to query the shape. */
int dvb_cnn_debug_tensor(DvbCnn* cnn, const char* name, int32_t n, float* out_host, int32_t* h,
                         int32_t* w, int32_t* c);

#ifdef __cplusplus
}
#endif
#endif  /* DVB_H_ */
