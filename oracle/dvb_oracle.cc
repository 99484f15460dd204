// dvb_oracle.cc — CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
//
// A plain C++17 restatement of the reference's pileup-image algorithm, used ONLY by
// tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg
// as the checker the CUDA path is compared with.  Nothing under deepvariant_b200/
// may import, link or call it.
//
// It deliberately keeps the reference's *structure* (one planar ImageRow of C
// byte-vectors per read, a sequential CIGAR walk with last-writer-wins column
// writes, std::shuffle down-sampling with a fresh std::mt19937_64, stable_sort with
// the (hap, group, position, name) comparator, blank fill, then a byte-by-byte
// planar->HWC flatten) so that it is an independent check of the data-parallel
// CUDA formulation.  Each function cites the reference file:line it follows
// (google/deepvariant v1.10.0).
//
// Parity pin: tests/test_pileup_kat.py checks it against the known-answer vectors of
// deepvariant/pileup_image_test.py, pileup_image_native_test.cc and
// pileup_channel_lib_test.cc (transcribed as data), and tests/golden/ holds images of
// the reference's own golden.calling_examples.tfrecord.gz re-derived through it.
//
// Build: see oracle/Makefile (g++ -O2 -shared -fPIC).  std::shuffle is implementation-defined: both
// forms are restated (DvbPileupParams.shuffle_stdlib) — libc++'s, which the reference's golden files were
// made with (tests/golden/downsample_golden_*), by hand; libstdc++'s by calling this compiler's own.

#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <numeric>
#include <random>
#include <string>
#include <tuple>
#include <vector>

#include "../include/dvb.h"

namespace {

// deepvariant/channels/channel.h:78
constexpr float kMaxPixelValueAsFloat = 254;
// deepvariant/channels/channel.h:81
constexpr float kMaxFragmentLength = 1000;

thread_local std::string g_err;

struct ImageRow {  // deepvariant/pileup_image_native.h ImageRow
  std::vector<std::vector<unsigned char>> channel_data;
  ImageRow(int width, int num_channels)
      : channel_data(num_channels, std::vector<unsigned char>(width, 0)) {}
};

struct ReadView {  // the fields of nucleus Read the path touches
  int32_t pos, sort_pos, mapq, fraglen, hp;
  uint8_t flags;
  uint32_t name_rank;
  const uint8_t* bases;
  const uint8_t* quals;
  int64_t len;
  const uint32_t* cigar;
  int64_t n_cigar;
  const uint8_t* base_plane[DVB_N_BASE_PLANES];   // per-base channel values of this read (nullptr: plane absent)
};

ReadView GetRead(const DvbBatch& b, int32_t r) {
  ReadView v;
  v.pos = b.read_pos[r];
  v.sort_pos = b.read_sort_pos[r];
  v.mapq = b.read_mapq[r];
  v.fraglen = b.read_fragment_length[r];
  v.hp = b.read_hp[r];
  v.flags = b.read_flags[r];
  v.name_rank = b.read_name_rank[r];
  v.bases = b.bases + b.read_seq_begin[r];
  v.quals = b.quals + b.read_seq_begin[r];
  v.len = b.read_seq_begin[r + 1] - b.read_seq_begin[r];
  v.cigar = b.cigar + b.read_cigar_begin[r];
  v.n_cigar = b.read_cigar_begin[r + 1] - b.read_cigar_begin[r];
  for (int k = 0; k < DVB_N_BASE_PLANES; ++k) v.base_plane[k] = b.base_channel[k] ? b.base_channel[k] + b.read_seq_begin[r] : nullptr;
  return v;
}

// ---- channel colour functions ------------------------------------------------

// channels/read_base_channel.cc:56-73
int BaseColor(char base, const DvbPileupParams& o) {
  switch (base) {
    case 'A': return o.base_color_offset_a_and_g + o.base_color_stride * 3;
    case 'G': return o.base_color_offset_a_and_g + o.base_color_stride * 2;
    case 'T': return o.base_color_offset_t_and_c + o.base_color_stride * 1;
    case 'C': return o.base_color_offset_t_and_c + o.base_color_stride * 0;
    default: return 0;
  }
}

// channels/base_quality_channel.cc:59-66 (identical in mapping_quality_channel.cc:60-67,
// haplotype_tag_channel.cc:102-109): float compare, float divide, truncate, narrow to uint8.
std::uint8_t ScaleColor(int value, float max_val) {
  if (static_cast<float>(value) > max_val) {
    value = max_val;
  }
  return static_cast<int>(kMaxPixelValueAsFloat * (static_cast<float>(value) / max_val));
}

// channels/read_supports_variant_channel.cc:105-117
int SupportsAltColor(int read_supports_alt, const DvbPileupParams& o) {
  float alpha;
  if (read_supports_alt == 0) {
    alpha = o.allele_unsupporting_read_alpha;
  } else if (read_supports_alt == 1) {
    alpha = o.allele_supporting_read_alpha;
  } else {
    alpha = o.other_allele_supporting_read_alpha;
  }
  return static_cast<int>(kMaxPixelValueAsFloat * alpha);
}

// channels/base_differs_from_ref_channel.cc:59-65
int MatchesRefColor(bool base_matches_ref, const DvbPileupParams& o) {
  float alpha = base_matches_ref ? o.reference_matching_read_alpha
                                 : o.reference_mismatching_read_alpha;
  return static_cast<int>(kMaxPixelValueAsFloat * alpha);
}

// channels/insert_size_channel.cc:80-89
int NormalizeFragmentLength(int fragment_length_in) {
  int fragment_length = std::abs(fragment_length_in);
  if (static_cast<float>(fragment_length) > kMaxFragmentLength) {
    fragment_length = static_cast<int>(kMaxFragmentLength);
  }
  return static_cast<int>(kMaxPixelValueAsFloat *
                          (static_cast<float>(fragment_length) / kMaxFragmentLength));
}

// channels/haplotype_tag_channel.cc:74-100
int GetHPValueForHPChannel(const ReadView& read, int hp_tag_for_assembly_polishing) {
  if (!(read.flags & DVB_READ_HAS_HP)) return 0;
  if (read.flags & DVB_READ_HP_MULTI) return 0;
  int hp_value = read.hp;
  if (hp_tag_for_assembly_polishing == 2) {
    if (hp_value == 1) return 2;
    if (hp_value == 2) return 1;
  }
  return hp_value;
}

// ---- "Opt Channels": whole-read statistics (the reference caches them per read in a std::optional) --------------
// channels/read_mapping_percent_channel.cc:71-90 and channels/identity_channel.cc:65-93 (same arithmetic: matched bases
// (M, =) over the sequence length, in percent).
int ReadMappingPercent(const ReadView& read) {
  int match_len = 0;
  for (int64_t k = 0; k < read.n_cigar; ++k) {
    const int op = read.cigar[k] & 0xF;
    if (op == 0 || op == 7) match_len += static_cast<int>(read.cigar[k] >> 4);   // ALIGNMENT_MATCH, SEQUENCE_MATCH
  }
  float mapping_percent = (static_cast<float>(match_len) / static_cast<float>(read.len)) * 100;
  return static_cast<int>(mapping_percent);
}

// channels/avg_base_quality_channel.cc:70-86 (the reference LOG(FATAL)s outside [0, 93]; callers validate).
int AvgBaseQuality(const ReadView& read) {
  int base_qual_sum = 0;
  for (int64_t i = 0; i < read.len; ++i) base_qual_sum += read.quals[i];
  float avg_base_qual = (static_cast<float>(base_qual_sum) / static_cast<float>(read.len));
  return static_cast<int>(avg_base_qual);
}

// channels/gap_compressed_identity_channel.cc:67-103: insertions and deletions count as single events.
int GapCompressedIdentity(const ReadView& read) {
  int match_len = 0;
  int gap_compressed_len = 0;
  for (int64_t k = 0; k < read.n_cigar; ++k) {
    const int op = read.cigar[k] & 0xF;
    const int op_len = static_cast<int>(read.cigar[k] >> 4);
    switch (op) {
      case 0: case 7: match_len += op_len; gap_compressed_len += op_len; break;
      case 8: gap_compressed_len += op_len; break;
      case 1: gap_compressed_len += 1; break;
      case 2: gap_compressed_len += 1; break;
      default: break;
    }
  }
  float gap_compressed_identity = static_cast<float>(match_len) / static_cast<float>(gap_compressed_len) * 100;
  return static_cast<int>(gap_compressed_identity);
}

// channels/gc_content_channel.cc:77-87
int GcContent(const uint8_t* seq, int64_t len) {
  int gc_count{};
  for (int64_t i = 0; i < len; ++i) {
    if (seq[i] == 'G' || seq[i] == 'C') gc_count += 1;
  }
  return static_cast<int>((static_cast<float>(gc_count) / static_cast<float>(len)) * 100);
}

// channels/is_homopolymer_channel.cc:77-91: 1 for every base inside a run of three or more equal bases.
std::vector<std::uint8_t> IsHomopolymer(const uint8_t* seq, int64_t len) {
  std::vector<std::uint8_t> homopolymer(static_cast<size_t>(len), 0);
  for (int64_t i = 2; i < len; i++) {
    if (seq[i] == seq[i - 1] && seq[i - 1] == seq[i - 2]) {
      homopolymer[i] = 1;
      homopolymer[i - 1] = 1;
      homopolymer[i - 2] = 1;
    }
  }
  return homopolymer;
}

// channels/homopolymer_weighted_channel.cc:79-101: the length of the run each base belongs to (stored in a uint8).
std::vector<std::uint8_t> HomopolymerWeighted(const uint8_t* seq, int64_t len) {
  std::vector<std::uint8_t> homopolymer(static_cast<size_t>(len), 0);
  if (len == 0) return homopolymer;
  int current_weight = 1;
  for (int64_t i = 1; i < len; i++) {
    if (seq[i] == seq[i - 1]) {
      current_weight += 1;
    } else {
      for (int cw = current_weight; cw >= 1; cw--) homopolymer[i - cw] = current_weight;
      current_weight = 1;
    }
  }
  for (int cw = current_weight; cw >= 1; cw--) homopolymer[len - cw] = current_weight;
  return homopolymer;
}

// ScaleColorVector (is_homopolymer_channel.cc:93-105 / homopolymer_weighted_channel.cc:103-115)
void ScaleColorVector(std::vector<std::uint8_t>& v, float max_val) {
  for (auto& x : v) {
    int value = x;
    if (static_cast<float>(value) > max_val) value = max_val;
    x = static_cast<int>(kMaxPixelValueAsFloat * (static_cast<float>(value) / max_val));
  }
}

// Per-base channel vectors of one sequence, computed once (the reference caches them in a std::optional per read).
struct BaseVectors {
  std::vector<std::uint8_t> is_homopolymer, homopolymer_weighted;
  BaseVectors(const uint8_t* seq, int64_t len, const DvbPileupParams& o) {
    for (int c = 0; c < o.num_channels; ++c) {
      if (o.channels[c] == DVB_CH_IS_HOMOPOLYMER && is_homopolymer.empty()) {
        is_homopolymer = IsHomopolymer(seq, len);
        ScaleColorVector(is_homopolymer, 1);      // kMaxIsHomopolymer
      }
      if (o.channels[c] == DVB_CH_HOMOPOLYMER_WEIGHTED && homopolymer_weighted.empty()) {
        homopolymer_weighted = HomopolymerWeighted(seq, len);
        ScaleColorVector(homopolymer_weighted, 30);   // kMaxHomopolymerWeighted
      }
    }
  }
};

bool ChannelSupported(int ch) {
  switch (ch) {
    case DVB_CH_READ_BASE: case DVB_CH_BASE_QUALITY: case DVB_CH_MAPPING_QUALITY:
    case DVB_CH_STRAND: case DVB_CH_READ_SUPPORTS_VARIANT: case DVB_CH_BASE_DIFFERS_FROM_REF:
    case DVB_CH_HAPLOTYPE_TAG: case DVB_CH_BLANK: case DVB_CH_INSERT_SIZE:
    case DVB_CH_SUPPLEMENTARY_ALIGNMENT:
    case DVB_CH_READ_MAPPING_PERCENT: case DVB_CH_AVG_BASE_QUALITY: case DVB_CH_IDENTITY:
    case DVB_CH_GAP_COMPRESSED_IDENTITY: case DVB_CH_GC_CONTENT:
    case DVB_CH_IS_HOMOPOLYMER: case DVB_CH_HOMOPOLYMER_WEIGHTED:
    case DVB_CH_ALLELE_FREQUENCY: case DVB_CH_READ_SUPPORTS_VARIANT_FUZZY: case DVB_CH_ALLELE_SAMPLE_PROBABILITY:
    case DVB_CH_BASE_METHYLATION: case DVB_CH_BASE_6MA: case DVB_CH_HOMOPOLYMER_INSERTION_QUALITY:
    case DVB_CH_HOMOPOLYMER_DELETION_QUALITY: case DVB_CH_INTER_HOMOPOLYMER_INSERTION_QUALITY:
    case DVB_CH_MEAN_COVERAGE:
      return true;
    default:
      return false;
  }
}

// One FillReadBase call (channels/*_channel.cc FillReadBase), dispatched by enum like
// Channels::ChannelEnumToObject (pileup_channel_lib.cc:367-447).
// Plane slot of a plane-backed channel (include/dvb.h), or -1.
int PairPlaneOf(int channel_enum) {
  switch (channel_enum) {
    case DVB_CH_ALLELE_FREQUENCY: return DVB_PAIR_PLANE_ALLELE_FREQUENCY;
    case DVB_CH_READ_SUPPORTS_VARIANT_FUZZY: return DVB_PAIR_PLANE_FUZZY_SUPPORT;
    case DVB_CH_ALLELE_SAMPLE_PROBABILITY: return DVB_PAIR_PLANE_ALLELE_SAMPLE_PROBABILITY;
    default: return -1;
  }
}
int BasePlaneOf(int channel_enum) {
  switch (channel_enum) {
    case DVB_CH_BASE_METHYLATION: return DVB_BASE_PLANE_5MC;
    case DVB_CH_BASE_6MA: return DVB_BASE_PLANE_6MA;
    case DVB_CH_HOMOPOLYMER_INSERTION_QUALITY: return DVB_BASE_PLANE_HMER_INSERTION;
    case DVB_CH_HOMOPOLYMER_DELETION_QUALITY: return DVB_BASE_PLANE_HMER_DELETION;
    case DVB_CH_INTER_HOMOPOLYMER_INSERTION_QUALITY: return DVB_BASE_PLANE_INTER_HMER_INSERTION;
    default: return -1;
  }
}

// pair_values: the (image, read) pair's bytes of DvbBatch.pair_channel (the value the reference caches per read in a std::optional:
// allele_frequency_channel.cc:57-68, read_supports_variant_fuzzy_channel.cc:99-110; allele_sample_probability_channel.cc:48-79).
unsigned char FillReadBase(int channel_enum, char read_base, char ref_base, int base_quality,
                           const ReadView& read, int support_class, const DvbPileupParams& o, int read_index,
                           const BaseVectors& bv, const uint8_t* pair_values) {
  switch (channel_enum) {
    case DVB_CH_ALLELE_FREQUENCY: case DVB_CH_READ_SUPPORTS_VARIANT_FUZZY: case DVB_CH_ALLELE_SAMPLE_PROBABILITY:
      return pair_values[PairPlaneOf(channel_enum)];
    case DVB_CH_BASE_METHYLATION: case DVB_CH_BASE_6MA:                       // data[col] = vector.at(read_index) when the read has one
    case DVB_CH_HOMOPOLYMER_INSERTION_QUALITY: case DVB_CH_HOMOPOLYMER_DELETION_QUALITY:
    case DVB_CH_INTER_HOMOPOLYMER_INSERTION_QUALITY:                          // data[col] = vector[read_index]
      return read.base_plane[BasePlaneOf(channel_enum)][read_index];
    case DVB_CH_MEAN_COVERAGE:           // a BlankChannel per read (pileup_channel_lib.cc:418-421); painted in BuildPileupForOneSample
      return 0;
    case DVB_CH_IS_HOMOPOLYMER:          // data[col] = vector.at(read_index)
      return bv.is_homopolymer.at(static_cast<size_t>(read_index));
    case DVB_CH_HOMOPOLYMER_WEIGHTED:
      return bv.homopolymer_weighted.at(static_cast<size_t>(read_index));
    case DVB_CH_READ_BASE:
      return BaseColor(read_base, o);
    case DVB_CH_BASE_QUALITY:
      return ScaleColor(base_quality, o.base_quality_cap);
    case DVB_CH_MAPPING_QUALITY:
      return ScaleColor(read.mapq, o.mapping_quality_cap);
    case DVB_CH_STRAND:  // channels/strand_channel.cc:44-61
      return static_cast<std::uint8_t>((read.flags & DVB_READ_REVERSE_STRAND)
                                           ? o.negative_strand_color
                                           : o.positive_strand_color);
    case DVB_CH_READ_SUPPORTS_VARIANT:
      return static_cast<unsigned char>(SupportsAltColor(support_class, o));
    case DVB_CH_BASE_DIFFERS_FROM_REF:
      return MatchesRefColor(read_base == ref_base, o);
    case DVB_CH_HAPLOTYPE_TAG:
      return ScaleColor(GetHPValueForHPChannel(read, o.hp_tag_for_assembly_polishing), 2);
    case DVB_CH_INSERT_SIZE:
      return NormalizeFragmentLength(read.fraglen);
    case DVB_CH_READ_MAPPING_PERCENT:     // kMaxMappingPercent = 100
    case DVB_CH_IDENTITY:                 // kMaxIdentity = 100
      return ScaleColor(ReadMappingPercent(read), 100);
    case DVB_CH_AVG_BASE_QUALITY:         // kMaxAvgBaseQuality = 93
      return ScaleColor(AvgBaseQuality(read), 93);
    case DVB_CH_GAP_COMPRESSED_IDENTITY:
      return ScaleColor(GapCompressedIdentity(read), 100);
    case DVB_CH_GC_CONTENT:               // kMaxGcContent = 100
      return ScaleColor(GcContent(read.bases, read.len), 100);
    case DVB_CH_SUPPLEMENTARY_ALIGNMENT: {  // channels/supplementary_alignment_channel.cc:49-58
      float alpha = (read.flags & DVB_READ_SUPPLEMENTARY) ? o.allele_supporting_read_alpha
                                                          : o.allele_unsupporting_read_alpha;
      return static_cast<unsigned char>(kMaxPixelValueAsFloat * alpha);
    }
    case DVB_CH_BLANK:
    default:
      return 0;
  }
}

// One FillRefBase call (channels/*_channel.cc FillRefBase).
unsigned char FillRefBase(int channel_enum, char ref_base, const DvbPileupParams& o, const uint8_t* ref_bases, int col,
                          const BaseVectors& ref_bv) {
  switch (channel_enum) {
    case DVB_CH_IS_HOMOPOLYMER:          // the reference window treated as a read: ref_data[col] = vector.at(col)
      return ref_bv.is_homopolymer.at(static_cast<size_t>(col));
    case DVB_CH_HOMOPOLYMER_WEIGHTED:
      return ref_bv.homopolymer_weighted.at(static_cast<size_t>(col));
    case DVB_CH_READ_MAPPING_PERCENT: case DVB_CH_AVG_BASE_QUALITY: case DVB_CH_IDENTITY:
    case DVB_CH_GAP_COMPRESSED_IDENTITY:   // *_channel.cc FillRefBase: kMaxPixelValueAsFloat
      return static_cast<std::uint8_t>(kMaxPixelValueAsFloat);
    case DVB_CH_GC_CONTENT:                // gc_content_channel.cc:64-75: GC content of the reference window itself
      return ScaleColor(GcContent(ref_bases, o.width), 100);
    case DVB_CH_READ_BASE:
      return BaseColor(ref_base, o);
    case DVB_CH_BASE_QUALITY:      // base_quality_channel.cc:52-57
    case DVB_CH_MAPPING_QUALITY:   // mapping_quality_channel.cc:53-58 (uses base_quality_cap too)
      return ScaleColor(o.reference_base_quality, o.base_quality_cap);
    case DVB_CH_STRAND:
      return static_cast<std::uint8_t>(o.positive_strand_color);
    case DVB_CH_READ_SUPPORTS_VARIANT:
    case DVB_CH_READ_SUPPORTS_VARIANT_FUZZY:   // read_supports_variant_fuzzy_channel.cc:112-116
      return SupportsAltColor(0, o);
    case DVB_CH_ALLELE_FREQUENCY:              // AlleleFrequencyColor(0) = 0 (allele_frequency_channel.cc:70-74, 79-82)
    case DVB_CH_ALLELE_SAMPLE_PROBABILITY: case DVB_CH_BASE_METHYLATION: case DVB_CH_BASE_6MA:
    case DVB_CH_HOMOPOLYMER_INSERTION_QUALITY: case DVB_CH_HOMOPOLYMER_DELETION_QUALITY:
    case DVB_CH_INTER_HOMOPOLYMER_INSERTION_QUALITY: case DVB_CH_MEAN_COVERAGE:
      return 0;
    case DVB_CH_BASE_DIFFERS_FROM_REF:
      return MatchesRefColor(true, o);
    case DVB_CH_HAPLOTYPE_TAG:
      return ScaleColor(0, 2);
    case DVB_CH_INSERT_SIZE:       // insert_size_channel.cc:67-71
      return static_cast<std::uint8_t>(kMaxPixelValueAsFloat);
    case DVB_CH_SUPPLEMENTARY_ALIGNMENT:  // supplementary_alignment_channel.cc:60-63: float alpha narrowed to uchar
      return static_cast<unsigned char>(o.allele_unsupporting_read_alpha);
    case DVB_CH_BLANK:
    default:
      return 0;
  }
}

// ---- Channels::CalculateChannels + CalculateBaseLevelData ---------------------
// pileup_channel_lib.cc:91-261.  Returns 1 ok, 0 rejected (low-quality base at the call
// site), -1 unrecognized CIGAR op (the reference LOG(FATAL)s).
int CalculateChannels(std::vector<std::vector<unsigned char>>& data, const DvbPileupParams& o,
                      const ReadView& read, const uint8_t* ref_bases, int width,
                      int variant_start, int support_class, int image_start_pos, const uint8_t* pair_values) {
  const BaseVectors bv(read.bases, read.len, o);
  // action_per_cigar_unit, pileup_channel_lib.cc:126-165 (op in BAM numbering here).
  auto action = [&](int ref_i, int read_i, int op) -> bool {
    char read_base = 0;
    if (op == 1) {  // INSERT
      read_base = static_cast<char>(o.indel_anchoring_base_char);
    } else if (op == 2) {  // DELETE
      ref_i -= 1;
      read_base = static_cast<char>(o.indel_anchoring_base_char);
    } else if (op == 0 || op == 7 || op == 8) {  // M, =, X
      read_base = static_cast<char>(read.bases[read_i]);
    }
    size_t col = ref_i - image_start_pos;
    if (read_base && 0 <= col && col < static_cast<size_t>(width)) {
      uint8_t base_quality = read.quals[read_i];
      if (ref_i == variant_start && base_quality < o.min_base_quality) {
        return false;
      }
      char ref_base = static_cast<char>(ref_bases[col]);
      for (int c = 0; c < o.num_channels; ++c) {
        if (o.blank_channel_mask & (1u << c)) continue;   // channels_enum_to_blank.contains(channel_enum), pileup_channel_lib.cc:154
        data[c][col] = FillReadBase(o.channels[c], read_base, ref_base, base_quality, read,
                                    support_class, o, read_i, bv, pair_values);
      }
    }
    return true;
  };

  // CalculateBaseLevelData, pileup_channel_lib.cc:171-261.
  int ref_i = read.pos;
  int read_i = 0;
  bool ok = true;
  for (int64_t k = 0; k < read.n_cigar; ++k) {
    const int op = read.cigar[k] & 0xF;
    const int op_len = static_cast<int>(read.cigar[k] >> 4);
    switch (op) {
      case 0: case 7: case 8:
        for (int i = 0; i < op_len; i++) {
          ok = ok && action(ref_i, read_i, op);
          ref_i++;
          read_i++;
        }
        break;
      case 1: case 4:  // INSERT, CLIP_SOFT
        if (ref_i > 0) {
          ok = action(ref_i - 1, read_i, op);
        }
        read_i += op_len;
        break;
      case 2: case 3:  // DELETE, SKIP
        if (read_i > 0) {
          ok = action(ref_i, read_i - 1, op);
        }
        ref_i += op_len;
        break;
      case 5: case 6:  // CLIP_HARD, PAD
        break;
      default:
        return -1;
    }
    if (!ok) return 0;
  }
  return 1;
}

// PileupImageEncoderNative::EncodeRead, pileup_image_native.cc:477-510.
// status: 1 row produced, 0 nullptr, -1 bad cigar.
int EncodeRead(const DvbPileupParams& o, const ReadView& read, const uint8_t* ref_bases,
               int variant_start, int support_class, int image_start_pos, const uint8_t* pair_values,
               std::unique_ptr<ImageRow>* out) {
  ImageRow img_row(o.width, o.num_channels);
  if (read.mapq < o.min_mapping_quality) {
    return 0;
  }
  int st = CalculateChannels(img_row.channel_data, o, read, ref_bases, o.width, variant_start,
                             support_class, image_start_pos, pair_values);
  if (st != 1) return st;
  *out = std::make_unique<ImageRow>(img_row);
  return 1;
}

// PileupImageEncoderNative::EncodeReference + Channels::CalculateRefRows,
// pileup_image_native.cc:512-527, pileup_channel_lib.cc:263-293.
std::unique_ptr<ImageRow> EncodeReference(const DvbPileupParams& o, const uint8_t* ref_bases) {
  ImageRow img_row(o.width, o.num_channels);
  const BaseVectors ref_bv(ref_bases, o.width, o);
  for (int c = 0; c < o.num_channels; ++c) {
    for (int i = 0; i < o.width; ++i) {
      img_row.channel_data[c][i] = FillRefBase(o.channels[c], static_cast<char>(ref_bases[i]), o, ref_bases, i, ref_bv);
    }
  }
  return std::make_unique<ImageRow>(img_row);
}

// libc++'s std::shuffle (llvm-project libcxx/include/__algorithm/shuffle.h) over libc++'s
// uniform_int_distribution<ptrdiff_t> (libcxx/include/__random/uniform_int_distribution.h, __independent_bits_engine): walk
// the range from the front; for the remaining length d + 1 draw j uniformly in [0, d] by keeping the low
// ceil(log2(d + 1)) bits of ONE 64-bit engine output and rejecting values >= d + 1; swap(first, first + j).
// Third-party dependency of the reference (the C++ standard library its binaries link), absent from /root/reference;
// pinned by the reference's golden.allele_frequency_examples (tools/check_downsample_golden.py): all 51 of its
// down-sampled examples reproduce row for row with this form and none with libstdc++'s.
template <class Gen>
void LibcxxShuffle(std::vector<int>& v, Gen& g) {
  long d = static_cast<long>(v.size());
  if (d <= 1) return;
  long first = 0;
  for (--d; d >= 1; ++first, --d) {
    const uint64_t rp = static_cast<uint64_t>(d) + 1;
    size_t w = 64 - static_cast<size_t>(__builtin_clzll(rp)) - 1;
    if ((rp & (~uint64_t{0} >> (64 - w))) != 0) ++w;
    const uint64_t mask = w >= 64 ? ~uint64_t{0} : (~uint64_t{0} >> (64 - w));
    uint64_t u;
    do {
      u = g() & mask;
    } while (u >= rp);
    if (u != 0) std::swap(v[first], v[first + static_cast<long>(u)]);
  }
}

// DownsampleReadIndices, pileup_image_native.cc:153-165 (gen passed BY VALUE).
std::vector<int> DownsampleReadIndices(size_t n_reads, int max_reads, std::mt19937_64 gen, int shuffle_stdlib) {
  std::vector<int> read_indices(n_reads);
  std::iota(read_indices.begin(), read_indices.end(), 0);
  if (n_reads > static_cast<size_t>(max_reads)) {
    if (shuffle_stdlib == DVB_SHUFFLE_LIBSTDCXX) std::shuffle(read_indices.begin(), read_indices.end(), gen);   // g++ / libstdc++
    else LibcxxShuffle(read_indices, gen);
  }
  return read_indices;
}

// PileupImageEncoderNative::GetHapIndex, pileup_image_native.cc:449-475.
int GetHapIndex(const DvbPileupParams& o, const ReadView& read) {
  if (!o.sort_by_haplotypes || !(read.flags & DVB_READ_HAS_HP)) return 0;
  int hp_value = read.hp;
  if (o.hp_tag_for_assembly_polishing > 0 && hp_value == o.hp_tag_for_assembly_polishing) {
    return -1;
  } else if (hp_value < 0) {
    return 0;
  }
  return hp_value;
}

// <hap_index, allele_support_group, position, read (name rank), image_row>
using ReadPileupTuple = std::tuple<int, int, int, uint32_t, std::unique_ptr<ImageRow>>;

// SortImageRows, pileup_image_native.cc:75-102.  (fragment_name, read_number) is
// carried as its dense rank among the batch's reads, which orders identically.
bool SortImageRows(const ReadPileupTuple& a, const ReadPileupTuple& b) {
  if (std::get<0>(a) != std::get<0>(b)) return std::get<0>(a) < std::get<0>(b);
  if (std::get<1>(a) != std::get<1>(b)) return std::get<1>(a) < std::get<1>(b);
  if (std::get<2>(a) != std::get<2>(b)) return std::get<2>(a) < std::get<2>(b);
  return std::get<3>(a) < std::get<3>(b);
}

// BuildPileupForOneSample, pileup_image_native.cc:296-447 (uniform down-sampling branch).
int BuildPileupForOneSample(const DvbPileupParams& o, const DvbBatch& b, int32_t img,
                            std::vector<std::unique_ptr<ImageRow>>* rows_out) {
  const uint8_t* ref_bases = b.ref_bases + static_cast<int64_t>(img) * b.ref_stride;
  const int pileup_height = o.height;
  const int max_reads = pileup_height - o.reference_band_height;
  const int64_t p0 = b.pair_begin[img], p1 = b.pair_begin[img + 1];
  const size_t n_reads = static_cast<size_t>(p1 - p0);

  std::vector<std::unique_ptr<ImageRow>>& rows = *rows_out;
  rows.reserve(pileup_height);
  for (int i = 0; i < o.reference_band_height; i++) {
    rows.push_back(EncodeReference(o, ref_bases));
  }

  auto gen = std::mt19937_64(o.random_seed);
  std::vector<int> sampled_indices = DownsampleReadIndices(n_reads, max_reads, gen, o.shuffle_stdlib);

  std::vector<ReadPileupTuple> pileup_of_reads;
  for (int index : sampled_indices) {
    if (pileup_of_reads.size() >= static_cast<size_t>(max_reads)) break;
    const int64_t p = p0 + index;
    const ReadView read = GetRead(b, b.pair_read[p]);
    std::unique_ptr<ImageRow> image_row;
    uint8_t pair_values[DVB_N_PAIR_PLANES];
    for (int k = 0; k < DVB_N_PAIR_PLANES; ++k) pair_values[k] = b.pair_channel[k] ? b.pair_channel[k][p] : 0;
    int st = EncodeRead(o, read, ref_bases, b.variant_start[img], b.pair_support[p],
                        b.image_start_pos[img], pair_values, &image_row);
    if (st < 0) return DVB_ERR_BAD_CIGAR;
    if (st == 0) continue;
    int hap_idx = GetHapIndex(o, read);
    int allele_support_group = 0;
    if (o.sort_by_alt_allele_support && b.pair_allele_group) {
      allele_support_group = b.pair_allele_group[p];
    }
    pileup_of_reads.emplace_back(hap_idx, allele_support_group, read.sort_pos, read.name_rank,
                                 std::move(image_row));
  }

  std::stable_sort(pileup_of_reads.begin(), pileup_of_reads.end(), SortImageRows);
  for (auto& t : pileup_of_reads) rows.push_back(std::move(std::get<4>(t)));

  int empty_rows = pileup_height - static_cast<int>(rows.size());
  for (int i = 0; i < empty_rows; i++) {
    rows.push_back(std::make_unique<ImageRow>(o.width, o.num_channels));
  }

  // "Add average coverage information after reads are added and sorted", pileup_image_native.cc:422-444.
  for (int c = 0; c < o.num_channels; ++c) {
    if (o.channels[c] != DVB_CH_MEAN_COVERAGE) continue;
    for (int i = 0; i < std::min(static_cast<int>(o.mean_coverage) + o.reference_band_height, pileup_height); i++) {
      rows[i]->channel_data[c].assign(o.width, i < o.reference_band_height ? 255 : 200);   // kChannelValue255 / kChannelValue200
    }
    break;   // std::find: the first such channel
  }
  return DVB_OK;
}

// FillPileupArray, pileup_image_native.h:214-308, for alt_image = {{}, {}}: the
// alt-aligned channels (if any) are written as 0 (alt1 empty -> 0; alt2 empty -> copy
// of alt1 = 0).
int64_t FillPileupArray(const std::vector<std::unique_ptr<ImageRow>>& image, int num_alt_channels,
                        unsigned char* pileup_array, int64_t buffer_size, int64_t buffer_pos) {
  for (size_t row = 0; row < image.size(); row++) {
    const int width = static_cast<int>(image[row]->channel_data.empty()
                                           ? 0
                                           : image[row]->channel_data[0].size());
    for (int column = 0; column < width; column++) {
      for (size_t channel = 0; channel < image[row]->channel_data.size(); channel++) {
        if (buffer_pos >= buffer_size) return -1;
        pileup_array[buffer_pos++] = image[row]->channel_data[channel][column];
      }
      for (int a = 0; a < num_alt_channels; ++a) {
        if (buffer_pos >= buffer_size) return -1;
        pileup_array[buffer_pos++] = 0;
      }
    }
  }
  return buffer_pos;
}

// A plane-backed channel needs its plane (include/dvb.h: DVB_ERR_INVALID_ARGUMENT otherwise).
int ValidatePlanes(const DvbPileupParams& o, const DvbBatch& b) {
  for (int c = 0; c < o.num_channels; ++c) {
    const int pp = PairPlaneOf(o.channels[c]), bp = BasePlaneOf(o.channels[c]);
    if ((pp >= 0 && b.n_pairs > 0 && !b.pair_channel[pp]) || (bp >= 0 && b.n_bases > 0 && !b.base_channel[bp])) {
      g_err = "channel enum " + std::to_string(o.channels[c]) + " needs its channel plane";
      return DVB_ERR_INVALID_ARGUMENT;
    }
  }
  return DVB_OK;
}

int ValidateParams(const DvbPileupParams& o) {
  if (o.width < 1) { g_err = "width must be >= 1"; return DVB_ERR_INVALID_ARGUMENT; }
  if (o.num_channels < 1 || o.num_channels > DVB_MAX_CHANNELS) { g_err = "bad num_channels"; return DVB_ERR_INVALID_ARGUMENT; }
  if (o.height <= o.reference_band_height || o.reference_band_height < 0) { g_err = "bad height"; return DVB_ERR_INVALID_ARGUMENT; }
  for (int c = 0; c < o.num_channels; ++c) {
    if (!ChannelSupported(o.channels[c])) { g_err = "unsupported channel enum " + std::to_string(o.channels[c]); return DVB_ERR_UNSUPPORTED_CHANNEL; }
  }
  return DVB_OK;
}

}  // namespace

extern "C" {

const char* dvb_oracle_last_error(void) { return g_err.c_str(); }

// Same contract as dvb_encode_batch_host (include/dvb.h) — host pointers in, host out.
int dvb_oracle_encode_batch(const DvbPileupParams* params, const DvbBatch* batch, uint8_t* out) {
  if (!params || !batch || (!out && batch->n_images > 0)) { g_err = "null argument"; return DVB_ERR_INVALID_ARGUMENT; }
  int st = ValidateParams(*params);
  if (st) return st;
  st = ValidatePlanes(*params, *batch);
  if (st) return st;
  const int64_t image_bytes = static_cast<int64_t>(params->height) * params->width *
                              (params->num_channels + params->num_alt_channels);
  for (int32_t img = 0; img < batch->n_images; ++img) {
    std::vector<std::unique_ptr<ImageRow>> rows;
    st = BuildPileupForOneSample(*params, *batch, img, &rows);
    if (st) { g_err = "Unrecognized CIGAR op"; return st; }
    unsigned char* dst = out + img * image_bytes;
    std::memset(dst, 0, image_bytes);
    if (FillPileupArray(rows, params->num_alt_channels, dst, image_bytes, 0) != image_bytes) {
      g_err = "FillPileupArray size mismatch";
      return DVB_ERR_INTERNAL;
    }
  }
  return DVB_OK;
}

// PileupImageEncoderNative::EncodeRead for read row `pair` of image `img`:
// out = uint8[width][num_channels]; *kept = 0 when the reference returns nullptr.
int dvb_oracle_encode_read(const DvbPileupParams* params, const DvbBatch* batch, int32_t img,
                           int64_t pair, uint8_t* out, int32_t* kept) {
  int st = ValidateParams(*params);
  if (st) return st;
  const DvbPileupParams& o = *params;
  st = ValidatePlanes(o, *batch);
  if (st) return st;
  const ReadView read = GetRead(*batch, batch->pair_read[pair]);
  std::unique_ptr<ImageRow> row;
  uint8_t pair_values[DVB_N_PAIR_PLANES];
  for (int k = 0; k < DVB_N_PAIR_PLANES; ++k) pair_values[k] = batch->pair_channel[k] ? batch->pair_channel[k][pair] : 0;
  int r = EncodeRead(o, read, batch->ref_bases + static_cast<int64_t>(img) * batch->ref_stride,
                     batch->variant_start[img], batch->pair_support[pair],
                     batch->image_start_pos[img], pair_values, &row);
  if (r < 0) { g_err = "Unrecognized CIGAR op"; return DVB_ERR_BAD_CIGAR; }
  *kept = r;
  if (r == 1) {
    for (int col = 0; col < o.width; ++col)
      for (int c = 0; c < o.num_channels; ++c) out[col * o.num_channels + c] = row->channel_data[c][col];
  }
  return DVB_OK;
}

// PileupImageEncoderNative::EncodeReference: out = uint8[width][num_channels].
int dvb_oracle_encode_reference(const DvbPileupParams* params, const uint8_t* ref_bases, uint8_t* out) {
  int st = ValidateParams(*params);
  if (st) return st;
  auto row = EncodeReference(*params, ref_bases);
  for (int col = 0; col < params->width; ++col)
    for (int c = 0; c < params->num_channels; ++c) out[col * params->num_channels + c] = row->channel_data[c][col];
  return DVB_OK;
}

// DownsampleReadIndices table for n reads (independent of the product's dvb_shuffle_table).
int dvb_oracle_shuffle_table(int32_t n, uint32_t seed, int32_t max_reads, int32_t shuffle_stdlib, int32_t* out) {
  std::vector<int> idx = DownsampleReadIndices(static_cast<size_t>(n), max_reads, std::mt19937_64(seed), shuffle_stdlib);
  for (int i = 0; i < n; ++i) out[i] = idx[i];
  return DVB_OK;
}

}  // extern "C"
